"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

The reference ships no golden vectors for this path, so these are
produced from the reference's own code: the in-tree operator restatement
``multi_scale_deformable_attn_pytorch_2d`` (temporal_self_attention.py:293-348) and the
reference's module classes ``SpatialCrossAttention`` / ``MSDeformableAttention3D`` /
``TemporalSelfAttention`` / ``CustomMSDeformableAttention`` / ``BEVFormerEncoder`` (its
``get_reference_points`` and ``point_sampling``), imported through ``oracle/refshim`` (a
stand-in for the absent mmcv) and executed on their CPU branch.  Inputs, parameters and
outputs are all stored, so the tests do not depend on RNG streams.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

from oracle.refshim import load_reference  # noqa: E402
import apollo_vision_net_b200.synthetic as syn  # noqa: E402  (numpy-only input generator)


def _np(t):
    return t.detach().cpu().numpy()


def _randomize(module, gen):
    """Non-degenerate parameters (the reference init gives zero offset weights / uniform
    attention, spatial_cross_attention.py:259,272): N(0, 0.02) on the offset / weight Linears."""
    for name, p in module.named_parameters():
        if name.endswith('sampling_offsets.weight') or name.endswith('attention_weights.weight'):
            p.data = torch.randn(p.shape, generator=gen) * 0.02
        elif name.endswith('attention_weights.bias'):
            p.data = torch.randn(p.shape, generator=gen) * 0.5


def _state(module, prefix):
    return {prefix + k: _np(v) for k, v in module.state_dict().items()}


def op_case(ref, gen):
    B, M, Dh, Nq, P = 2, 4, 16, 37, 3
    levels = [(9, 13), (5, 7), (3, 4)]
    L = len(levels)
    Nk = sum(h * w for h, w in levels)
    value = torch.randn(B, Nk, M, Dh, generator=gen)
    loc = torch.rand(B, Nq, M, L, P, 2, generator=gen) * 1.3 - 0.15
    att = torch.softmax(torch.randn(B, Nq, M, L * P, generator=gen), -1).view(B, Nq, M, L, P)
    shapes = torch.tensor(levels)
    v, s, a = (t.clone().requires_grad_(True) for t in (value, loc, att))
    out = ref.msda_pytorch_2d(v, shapes, s, a)
    go = torch.randn(out.shape, generator=gen)
    out.backward(go)
    # float64 run of the same reference function: the yardstick for gradient tolerances
    v64, s64, a64 = (t.double().clone().requires_grad_(True) for t in (value, loc, att))
    out64 = ref.msda_pytorch_2d(v64, shapes, s64, a64)
    out64.backward(go.double())
    return dict(value=_np(value), loc=_np(loc), attn=_np(att), shapes=np.array(levels, np.int64),
                out=_np(out), grad_out=_np(go), grad_value=_np(v.grad), grad_loc=_np(s.grad),
                grad_attn=_np(a.grad), out64=_np(out64), grad_value64=_np(v64.grad),
                grad_loc64=_np(s64.grad), grad_attn64=_np(a64.grad))


def geometry_case(ref):
    bs, H, W, D = 2, 50, 50, 4
    l2i, img_shape = syn.camera_rig(0.5, bs=bs, jitter=4.0, seed=3)
    r3 = ref.BEVFormerEncoder.get_reference_points(H, W, 8.0, D, dim='3d', bs=bs, device='cpu',
                                                   dtype=torch.float32)
    r2 = ref.BEVFormerEncoder.get_reference_points(H, W, dim='2d', bs=bs, device='cpu',
                                                   dtype=torch.float32)
    metas = [dict(lidar2img=[l2i[b, i] for i in range(6)], img_shape=[img_shape] * 6)
             for b in range(bs)]

    class _Self:  # point_sampling only reads self.debug_nan
        pass
    uv, mask = ref.BEVFormerEncoder.point_sampling(_Self(), r3, syn.PC_RANGE, metas)
    lists = [m[0].sum(-1).nonzero().squeeze(-1) for m in mask]
    out = dict(lidar2img=l2i, img_shape=np.array(img_shape), pc_range=np.array(syn.PC_RANGE),
               bev_hw=np.array([H, W]), D=np.array(D), ref_3d=_np(r3), ref_2d=_np(r2),
               ref_cam=_np(uv), mask_packed=np.packbits(_np(mask).astype(np.uint8)),
               mask_shape=np.array(mask.shape), hit_count=np.array([len(x) for x in lists]))
    for i, x in enumerate(lists):
        out[f'hit_index_{i}'] = _np(x).astype(np.int32)
    return out


def sca_case(ref, gen):
    bs, H, W, C, heads, P, D = 2, 12, 14, 64, 4, 8, 4
    levels = [(12, 20), (6, 10)]
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    l2i, img_shape = syn.camera_rig(0.5, bs=bs, jitter=4.0, seed=5)
    r3 = ref.BEVFormerEncoder.get_reference_points(H, W, 8.0, D, dim='3d', bs=bs, device='cpu',
                                                   dtype=torch.float32)
    metas = [dict(lidar2img=[l2i[b, i] for i in range(6)], img_shape=[img_shape] * 6)
             for b in range(bs)]

    class _Self:
        pass
    uv, mask = ref.BEVFormerEncoder.point_sampling(_Self(), r3, syn.PC_RANGE, metas)
    mod = ref.SpatialCrossAttention(
        embed_dims=C, num_cams=6, pc_range=syn.PC_RANGE, dropout=0.1, batch_first=True,
        deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C, num_heads=heads,
                                  num_points=P, num_levels=len(levels), attn_logits_clamp=0.01))
    _randomize(mod, gen)
    mod.eval()
    query = torch.randn(bs, H * W, C, generator=gen).requires_grad_(True)
    feat = torch.randn(6, Nk, bs, C, generator=gen).requires_grad_(True)
    qpos = torch.randn(bs, H * W, C, generator=gen)
    shapes = torch.tensor(shapes_l)
    starts = torch.tensor(starts_l)
    out = mod(query, feat, feat, query_pos=qpos, reference_points_cam=uv, bev_mask=mask,
              spatial_shapes=shapes, level_start_index=starts)
    go = torch.randn(out.shape, generator=gen)
    out.backward(go)
    res = dict(query=_np(query), feat=_np(feat), query_pos=_np(qpos), ref_cam=_np(uv),
               mask=_np(mask), shapes=np.array(shapes_l, np.int64), starts=np.array(starts_l, np.int64),
               out=_np(out), grad_out=_np(go), grad_query=_np(query.grad), grad_feat=_np(feat.grad),
               cfg=np.array([bs, H, W, C, heads, P, D]))
    res.update(_state(mod, 'param.'))
    for n, p in mod.named_parameters():
        res['pgrad.' + n] = _np(p.grad)
    # the stand-alone MSDeformableAttention3D on a rebatched slice (its own call contract)
    da = mod.deformable_attention
    idx = mask[0][0].sum(-1).nonzero().squeeze(-1)
    q1 = query.detach()[:1, idx]
    r1 = uv[0][:1, idx]
    v1 = feat.detach()[0].permute(1, 0, 2)[:1]
    res['da_query'], res['da_ref'], res['da_value'] = _np(q1), _np(r1), _np(v1)
    res['da_out'] = _np(da(query=q1, key=v1, value=v1, reference_points=r1, spatial_shapes=shapes,
                           level_start_index=starts))
    return res


def tsa_case(ref, gen, with_prev):
    bs, H, W, C, heads, P = 2, 9, 11, 64, 4, 4
    mod = ref.TemporalSelfAttention(embed_dims=C, num_heads=heads, num_levels=1, num_points=P,
                                    attn_logits_clamp=0.6)
    _randomize(mod, gen)
    mod.eval()
    query = torch.randn(bs, H * W, C, generator=gen).requires_grad_(True)
    qpos = torch.randn(bs, H * W, C, generator=gen)
    r2 = ref.BEVFormerEncoder.get_reference_points(H, W, dim='2d', bs=bs, device='cpu',
                                                   dtype=torch.float32)
    shift = torch.tensor([[0.03, -0.02], [-0.05, 0.01]])
    hybrid = torch.stack([r2 + shift[:, None, None, :], r2], 1).reshape(bs * 2, H * W, 1, 2)
    prev = torch.randn(bs * 2, H * W, C, generator=gen).requires_grad_(True) if with_prev else None
    out = mod(query, prev, prev, query_pos=qpos, reference_points=hybrid,
              spatial_shapes=torch.tensor([[H, W]]), level_start_index=torch.tensor([0]))
    go = torch.randn(out.shape, generator=gen)
    out.backward(go)
    res = dict(query=_np(query), query_pos=_np(qpos), ref=_np(hybrid), out=_np(out), grad_out=_np(go),
               grad_query=_np(query.grad), cfg=np.array([bs, H, W, C, heads, P]))
    if with_prev:
        res['prev'] = _np(prev)
        res['grad_prev'] = _np(prev.grad)
    res.update(_state(mod, 'param.'))
    for n, p in mod.named_parameters():
        res['pgrad.' + n] = _np(p.grad)
    return res


def decoder_case(ref, gen):
    bs, H, W, C, heads, P, Nq = 2, 10, 10, 64, 4, 4, 60
    mod = ref.CustomMSDeformableAttention(embed_dims=C, num_heads=heads, num_levels=1, num_points=P,
                                          attn_logits_clamp=0.8)
    _randomize(mod, gen)
    mod.eval()
    query = torch.randn(Nq, bs, C, generator=gen).requires_grad_(True)
    qpos = torch.randn(Nq, bs, C, generator=gen)
    value = torch.randn(H * W, bs, C, generator=gen).requires_grad_(True)
    refp = torch.rand(bs, Nq, 1, 2, generator=gen).requires_grad_(True)
    out = mod(query, None, value, query_pos=qpos, reference_points=refp,
              spatial_shapes=torch.tensor([[H, W]]), level_start_index=torch.tensor([0]))
    go = torch.randn(out.shape, generator=gen)
    out.backward(go)
    res = dict(query=_np(query), query_pos=_np(qpos), value=_np(value), ref=_np(refp), out=_np(out),
               grad_out=_np(go), grad_query=_np(query.grad), grad_value=_np(value.grad),
               grad_ref=_np(refp.grad), cfg=np.array([bs, H, W, C, heads, P, Nq]))
    res.update(_state(mod, 'param.'))
    for n, p in mod.named_parameters():
        res['pgrad.' + n] = _np(p.grad)
    return res


def det_decoder_case(ref, gen):
    """The reference's DetectionTransformerDecoder (decoder.py:50-126) -- layer loop, 3-d reference
    point refinement through reg_branches, detaching, stacking of the intermediates -- driven with
    a small stub layer (the reference's CustomMSDeformableAttention followed by a LayerNorm), so
    that the decoder-level logic is pinned without mmdet's DetrTransformerDecoderLayer."""
    import torch.nn as nn
    bs, H, W, C, heads, P, Nq, NL = 2, 9, 7, 64, 4, 4, 23, 3

    class StubDecLayer(nn.Module):
        def __init__(self, embed_dims, num_heads, num_points):
            super().__init__()
            self.embed_dims = embed_dims
            self.attn = ref.CustomMSDeformableAttention(embed_dims=embed_dims, num_heads=num_heads,
                                                        num_levels=1, num_points=num_points)
            self.norm = nn.LayerNorm(embed_dims)

        def forward(self, query, key=None, value=None, query_pos=None, reference_points=None,
                    spatial_shapes=None, level_start_index=None, key_padding_mask=None, **kw):
            q = self.attn(query, key, value, query_pos=query_pos, reference_points=reference_points,
                          spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                          key_padding_mask=key_padding_mask)
            return self.norm(q)

    # the shim's TransformerLayerSequence builds its layers from the shim's TRANSFORMER_LAYER registry
    ref.LAYER.module_dict['StubDecLayer'] = StubDecLayer
    dec = ref.DetectionTransformerDecoder(
        transformerlayers=dict(type='StubDecLayer', embed_dims=C, num_heads=heads, num_points=P),
        num_layers=NL, return_intermediate=True)
    _randomize(dec, gen)
    dec.eval()
    reg = nn.ModuleList([nn.Linear(C, 10) for _ in range(NL)])
    for m in reg:
        m.weight.data = torch.randn(m.weight.shape, generator=gen) * 0.05
        m.bias.data = torch.randn(m.bias.shape, generator=gen) * 0.05
    query = torch.randn(Nq, bs, C, generator=gen).requires_grad_(True)
    qpos = torch.randn(Nq, bs, C, generator=gen)
    value = torch.randn(H * W, bs, C, generator=gen).requires_grad_(True)
    refp = torch.rand(bs, Nq, 3, generator=gen)
    inter, refs = dec(query, key=None, value=value, query_pos=qpos, reference_points=refp,
                      reg_branches=reg, spatial_shapes=torch.tensor([[H, W]]),
                      level_start_index=torch.tensor([0]))
    go = torch.randn(inter.shape, generator=gen)
    inter.backward(go)
    res = dict(query=_np(query), query_pos=_np(qpos), value=_np(value), ref=_np(refp),
               inter=_np(inter), refs=_np(refs), grad_out=_np(go), grad_query=_np(query.grad),
               grad_value=_np(value.grad), cfg=np.array([bs, H, W, C, heads, P, Nq, NL]))
    res.update(_state(dec, 'param.'))
    res.update(_state(reg, 'reg.'))
    return res


MAPTR_ORDER = ('self_attn', 'norm', 'self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')


def maptrv2_decoder_case(ref, gen):
    """The reference's MapTRv2Decoder + MapTRv2DecoupledDetrTransformerDecoderLayer
    (maptrv2/modules/decoder.py:10-213): inter-vector / intra-vector self-attention reshapes,
    deformable cross-attention on the BEV, 2-d reference refinement -- one-to-one and one-to-many
    vectors separated by the self-attention mask.  mmcv's MultiheadAttention / FFN /
    BaseTransformerLayer constructor come from the shim (third-party, restated)."""
    import torch.nn as nn
    bs, H, W, C, heads, P, V, Pn, NL = 2, 8, 6, 64, 4, 4, 6, 4, 2
    dec = ref.MapTRv2Decoder(
        transformerlayers=dict(
            type='MapTRv2DecoupledDetrTransformerDecoderLayer', num_vec=V, num_pts_per_vec=Pn,
            attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=heads, dropout=0.1),
                       dict(type='MultiheadAttention', embed_dims=C, num_heads=heads, dropout=0.1),
                       dict(type='CustomMSDeformableAttention', embed_dims=C, num_heads=heads,
                            num_levels=1, num_points=P)],
            feedforward_channels=2 * C, ffn_dropout=0.1, operation_order=MAPTR_ORDER),
        num_layers=NL, return_intermediate=True)
    _randomize(dec, gen)
    dec.eval()
    reg = nn.ModuleList([nn.Linear(C, 2) for _ in range(NL)])
    for m in reg:
        m.weight.data = torch.randn(m.weight.shape, generator=gen) * 0.05
        m.bias.data = torch.randn(m.bias.shape, generator=gen) * 0.05
    query = torch.randn(V * Pn, bs, C, generator=gen).requires_grad_(True)
    qpos = torch.randn(V * Pn, bs, C, generator=gen)
    value = torch.randn(H * W, bs, C, generator=gen).requires_grad_(True)
    refp = torch.rand(bs, V * Pn, 2, generator=gen)
    mask = torch.zeros(V, V, dtype=torch.bool)
    mask[V // 2:, :V // 2] = True
    mask[:V // 2, V // 2:] = True
    inter, refs = dec(query, key=None, value=value, query_pos=qpos, reference_points=refp,
                      reg_branches=reg, spatial_shapes=torch.tensor([[H, W]]),
                      level_start_index=torch.tensor([0]), self_attn_mask=mask, num_vec=V,
                      num_pts_per_vec=Pn)
    go = torch.randn(inter.shape, generator=gen)
    inter.backward(go)
    res = dict(query=_np(query), query_pos=_np(qpos), value=_np(value), ref=_np(refp), mask=_np(mask),
               inter=_np(inter), refs=_np(refs), grad_out=_np(go), grad_query=_np(query.grad),
               grad_value=_np(value.grad), cfg=np.array([bs, H, W, C, heads, P, V, Pn, NL]))
    res.update(_state(dec, 'param.'))
    res.update(_state(reg, 'reg.'))
    return res


ENC_ORDER = ('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')


def encoder_case(ref, gen):
    """The reference's BEVFormerEncoder + BEVFormerLayer (encoder.py:243-352, 354-519) built on its
    own MyCustomBaseTransformerLayer (custom_base_transformer_layer.py), two layers, with history
    and a CAN-bus shift: reference points, point_sampling from img_metas, the hybrid ref_2d stack
    (incl. the aliasing quirk, encoder.py:309-311), the [prev_bev, bev] value pair, TSA -> norm ->
    SCA -> norm -> FFN -> norm, all on the reference's CPU branch."""
    bs, H, W, C, heads = 2, 8, 10, 64, 8
    levels = [(6, 10), (3, 5)]
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    enc = ref.BEVFormerEncoder(
        transformerlayers=dict(
            type='BEVFormerLayer',
            attn_cfgs=[dict(type='TemporalSelfAttention', embed_dims=C, num_heads=heads, num_levels=1,
                            num_points=4),
                       dict(type='SpatialCrossAttention', embed_dims=C, num_cams=6, pc_range=syn.PC_RANGE,
                            deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                                      num_heads=heads, num_points=8,
                                                      num_levels=len(levels)))],
            feedforward_channels=2 * C, ffn_dropout=0.1,
            ffn_cfgs=dict(type='FFN', embed_dims=C, feedforward_channels=2 * C, num_fcs=2, ffn_drop=0.1,
                          act_cfg=dict(type='ReLU', inplace=True)),
            operation_order=ENC_ORDER),
        num_layers=2, pc_range=syn.PC_RANGE, num_points_in_pillar=4, return_intermediate=False)
    _randomize(enc, gen)
    enc.eval()
    l2i, img_shape = syn.camera_rig(0.05, bs=bs, jitter=4.0, seed=5)
    metas = [dict(lidar2img=[l2i[b, i] for i in range(6)], img_shape=[img_shape] * 6)
             for b in range(bs)]
    bevq = torch.randn(H * W, bs, C, generator=gen).requires_grad_(True)
    pos = torch.randn(H * W, bs, C, generator=gen)
    prev = torch.randn(H * W, bs, C, generator=gen).requires_grad_(True)
    feat = torch.randn(6, Nk, bs, C, generator=gen).requires_grad_(True)
    shift = torch.tensor([[0.012, -0.02], [-0.03, 0.007]])
    out = enc(bevq, feat, feat, bev_h=H, bev_w=W, bev_pos=pos, spatial_shapes=torch.tensor(shapes_l),
              level_start_index=torch.tensor(starts_l), prev_bev=prev, shift=shift, img_metas=metas)
    go = torch.randn(out.shape, generator=gen)
    out.backward(go)
    res = dict(bev_query=_np(bevq), bev_pos=_np(pos), prev_bev=_np(prev), feat=_np(feat), shift=_np(shift),
               lidar2img=l2i, img_shape=np.array(img_shape), levels=np.array(levels, np.int64),
               out=_np(out), grad_out=_np(go), grad_bev_query=_np(bevq.grad), grad_prev_bev=_np(prev.grad),
               grad_feat=_np(feat.grad), cfg=np.array([bs, H, W, C, heads]))
    res.update(_state(enc, 'param.'))
    for n, p in enc.named_parameters():
        res['pgrad.' + n] = _np(p.grad)
    return res


def bev_features_case(ref, gen):
    """PerceptionTransformer.get_bev_features (transformer.py:119-298): can_bus shift, prev_bev
    rotation (torchvision rotate, nearest), can_bus MLP, camera / level embeddings, flattening of
    the multi-level image features.  The encoder is replaced by a stub that records what it is
    called with: that call contract is what the B200 pre-processing must reproduce."""
    import torch.nn as nn
    captured = {}

    class CaptureEncoder(nn.Module):
        def __init__(self, **kw):
            super().__init__()

        def forward(self, bev_queries, key, value, **kw):
            captured.update(bev_queries=bev_queries, feat_flatten=key, **{
                k: kw[k] for k in ('bev_pos', 'spatial_shapes', 'level_start_index', 'prev_bev', 'shift')})
            return bev_queries.permute(1, 0, 2)

    ref.LAYER_SEQ.module_dict['CaptureEncoder'] = CaptureEncoder
    bs, num_cam, C, bev_h, bev_w = 2, 6, 64, 10, 12
    levels = [(6, 10), (3, 5), (2, 3)]
    trf = ref.PerceptionTransformer(num_feature_levels=len(levels), num_cams=num_cam, embed_dims=C,
                                    encoder=dict(type='CaptureEncoder'), decoder=None,
                                    rotate_center=[bev_w // 2, bev_h // 2])
    trf.init_weights()
    for prm in trf.parameters():                       # non-degenerate biases / norm too
        if prm.dim() == 1:
            prm.data = torch.randn(prm.shape, generator=gen) * 0.3
    trf.eval()
    mlvl = [torch.randn(bs, num_cam, C, h, w, generator=gen) for h, w in levels]
    mlvl[1][0, 2, 5, 1, 1] = float('nan')              # the reference sanitises non-finite features
    mlvl[2][1, 4, 7, 0, 2] = float('inf')
    bev_queries = torch.randn(bev_h * bev_w, C, generator=gen)
    bev_pos = torch.randn(bs, C, bev_h, bev_w, generator=gen)
    prev_bev = torch.randn(bs, bev_h * bev_w, C, generator=gen)
    can_bus = np.zeros((bs, 18))
    can_bus[0, :3] = [0.8, -0.35, 0.0]
    can_bus[0, -2:] = [0.21, 12.0]                     # ego angle (rad), rotation of prev_bev (deg)
    can_bus[1, :3] = [-1.3, 0.6, 0.0]
    can_bus[1, -2:] = [-0.4, -33.5]
    can_bus[:, 3:16] = np.asarray(torch.randn(bs, 13, generator=gen))
    metas = [dict(can_bus=list(can_bus[b])) for b in range(bs)]
    grid_length = (0.512, 0.6)
    out = trf.get_bev_features([m.clone() for m in mlvl], bev_queries, bev_h, bev_w,
                               grid_length=grid_length, bev_pos=bev_pos, prev_bev=prev_bev.clone(),
                               img_metas=metas)
    res = dict(cfg=np.array([bs, num_cam, C, bev_h, bev_w]), levels=np.array(levels, np.int64),
               bev_queries=_np(bev_queries), bev_pos=_np(bev_pos), prev_bev=_np(prev_bev),
               can_bus=can_bus, grid_length=np.array(grid_length),
               out_bev_queries=_np(captured['bev_queries']), out_feat_flatten=_np(captured['feat_flatten']),
               out_bev_pos=_np(captured['bev_pos']), out_spatial_shapes=_np(captured['spatial_shapes']),
               out_level_start_index=_np(captured['level_start_index']),
               out_prev_bev=_np(captured['prev_bev']), out_shift=_np(captured['shift']), out=_np(out))
    for i, m in enumerate(mlvl):
        res[f'feat_{i}'] = _np(m)
    res.update(_state(trf, 'param.'))
    return res


def dcnv3_case(gen):
    """DCNv3 through the reference's own pure-PyTorch ``dcnv3_core_pytorch`` (ops_dcnv3/functions/dcnv3_func.py:
    119-188), forward and autograd gradients: a 3x3 kernel with padding (the backbone's configuration,
    modules/dcnv3.py:216-345) and a strided, dilated, rectangular-kernel case without padding."""
    from oracle.refshim import load_reference_dcnv3
    fn = load_reference_dcnv3().dcnv3_core_pytorch
    res = {}
    for tag, (N, H, W, G, Cg, kh, kw, sh, sw, ph, pw, dh, dw, scale) in {
            'a': (2, 9, 11, 4, 8, 3, 3, 1, 1, 1, 1, 1, 1, 1.0),
            'b': (1, 13, 10, 2, 16, 3, 2, 2, 1, 0, 0, 2, 1, 0.7)}.items():
        Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
        Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
        K = kh * kw
        x = torch.randn(N, H, W, G * Cg, generator=gen, requires_grad=True)
        off = (torch.randn(N, Ho, Wo, G * K * 2, generator=gen) * 1.5).requires_grad_(True)
        msk = torch.softmax(torch.randn(N, Ho, Wo, G, K, generator=gen), -1).reshape(N, Ho, Wo, G * K).requires_grad_(True)
        go = torch.randn(N, Ho, Wo, G * Cg, generator=gen)
        out = fn(x, off, msk, kh, kw, sh, sw, ph, pw, dh, dw, G, Cg, scale)
        out.backward(go)
        res.update({f'{tag}.cfg': np.array([kh, kw, sh, sw, ph, pw, dh, dw, G, Cg], dtype=np.int64),
                    f'{tag}.offset_scale': np.array(scale, dtype=np.float64),
                    f'{tag}.input': _np(x), f'{tag}.offset': _np(off), f'{tag}.mask': _np(msk),
                    f'{tag}.grad_output': _np(go), f'{tag}.output': _np(out), f'{tag}.grad_input': _np(x.grad),
                    f'{tag}.grad_offset': _np(off.grad), f'{tag}.grad_mask': _np(msk.grad)})
    return res


def main():
    ref = load_reference()
    gen = torch.Generator().manual_seed(20261018)
    cases = {
        'op_small': op_case(ref, gen),
        'geometry_tiny': geometry_case(ref),
        'sca_small': sca_case(ref, gen),
        'tsa_prev': tsa_case(ref, gen, True),
        'tsa_first_frame': tsa_case(ref, gen, False),
        'decoder_small': decoder_case(ref, gen),
        'bev_features_small': bev_features_case(ref, gen),
        'det_decoder_small': det_decoder_case(ref, gen),
        'maptrv2_decoder_small': maptrv2_decoder_case(ref, gen),
        'encoder_small': encoder_case(ref, gen),
        'dcnv3_small': dcnv3_case(torch.Generator().manual_seed(20261019)),
    }
    only = sys.argv[1:]                                # optional: names of the cases to (re)write
    for name, arrays in cases.items():
        if only and name not in only:
            continue
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **arrays)
        print(f'{name}: {len(arrays)} arrays, {os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    main()
