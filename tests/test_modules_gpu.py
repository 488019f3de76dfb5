"""GPU parity of the fused kernels and the registry-named modules against (a) the golden
vectors produced by the unmodified reference and (b) the CPU oracle on larger seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import geometry_oracle as G
from oracle.modules_oracle import (OracleBEVFormerEncoder, OracleCustomMSDeformableAttention,
                                   OracleSpatialCrossAttention, OracleTemporalSelfAttention)
from tests import golden_util as gu
from tests.util import rel_l2, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
FWD, BWD = 1e-5, 1e-4        # north_star: forward 1e-5 relative in fp32, gradients 1e-4


def _randomize(m, seed=0):
    g = torch.Generator().manual_seed(seed)
    for n, p in m.named_parameters():
        if n.endswith('sampling_offsets.weight') or n.endswith('attention_weights.weight'):
            p.data = torch.randn(p.shape, generator=g) * 0.02
        elif n.endswith('attention_weights.bias'):
            p.data = torch.randn(p.shape, generator=g) * 0.5


# ------------------------------------------------------------------ geometry (bit-exact) ---
def _check_geometry(l2i, img_shape, H, W, D, bs, pc_range):
    import apollo_vision_net_b200 as pkg
    r3 = G.reference_points_3d(H, W, pc_range[5] - pc_range[2], D, bs=bs)
    uv, mask = G.point_sampling(r3, pc_range, l2i, img_shape[0], img_shape[1])
    lists, max_len = G.camera_hit_lists(mask)
    geo = pkg.bev_point_sampling(r3.to(DEV), pc_range, l2i, img_shape[0], img_shape[1])
    torch.cuda.synchronize()
    assert torch.equal(geo.bev_mask.cpu(), mask)                       # bit-exact
    assert torch.equal(geo.reference_points_cam.cpu(), uv)             # same op order => identical
    counts = geo.hit_count.cpu().tolist()
    assert counts == [len(x) for x in lists]
    for i, x in enumerate(lists):
        assert torch.equal(geo.hit_index[i, :counts[i]].cpu().long(), x)
        assert bool((geo.hit_index[i, counts[i]:] == -1).all())
    bits = geo.hit_bits.cpu()
    for i in range(mask.shape[0]):
        assert torch.equal(((bits >> i) & 1).bool(), mask[i].any(-1))
    return geo


def test_point_sampling_golden():
    g = gu.load('geometry_tiny')
    import apollo_vision_net_b200 as pkg
    H, W = (int(x) for x in g['bev_hw'])
    geo = pkg.bev_point_sampling(gu.T(g['ref_3d'], DEV), list(g['pc_range']), g['lidar2img'],
                                 int(g['img_shape'][0]), int(g['img_shape'][1]))
    assert np.array_equal(geo.bev_mask.cpu().numpy(), gu.unpack_mask(g))
    assert np.array_equal(geo.reference_points_cam.cpu().numpy(), g['ref_cam'])
    assert geo.hit_count.cpu().tolist() == g['hit_count'].tolist()
    for i in range(6):
        n = int(g['hit_count'][i])
        assert np.array_equal(geo.hit_index[i, :n].cpu().numpy(), g[f'hit_index_{i}'])


@pytest.mark.parametrize('H,W,scale,bs', [(50, 50, 0.5, 1), (200, 200, 1.0, 1), (37, 61, 0.5, 3),
                                          (400, 400, 1.0, 1)])
def test_point_sampling_bit_exact(H, W, scale, bs):
    import apollo_vision_net_b200.synthetic as syn
    l2i, img_shape = syn.camera_rig(scale, bs=bs, jitter=5.0, seed=H)
    _check_geometry(l2i, img_shape, H, W, 4, bs, syn.PC_RANGE)


def test_point_sampling_boundary_grazing():
    """Points that land exactly on z = eps, u = 0, u = 1, v = 0, v = 1: the strict inequalities of
    encoder.py:186,228-231 must exclude them.  Exactly representable arithmetic (integer camera
    matrix, power-of-two image size, pc_range mapping [0,1] -> integers)."""
    H, W, D = 8, 8, 2
    pc_range = [0.0, 0.0, 0.0, 8.0, 8.0, 8.0]
    # u = 64 * X / Z, v = 64 * Y / Z with image 256 x 256: X/Z = 4 -> u_norm = 1 exactly; X = 0 never
    # happens (cell centres), so add a second camera with an offset that puts centres on u = 0.
    l2i = np.zeros((1, 3, 4, 4), np.float32)
    l2i[0, 0] = [[64, 0, 0, 0], [0, 64, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    l2i[0, 1] = [[64, 0, 0, -32], [0, 64, 0, -96], [0, 0, 1, 0], [0, 0, 0, 1]]    # centres hit u=0 / v=0
    l2i[0, 2] = [[64, 0, 0, 0], [0, 64, 0, 0], [0, 0, 1, -2], [0, 0, 0, 1]]        # z - 2: z=2 -> 0 <= eps
    geo = _check_geometry(l2i, (256, 256, 3), H, W, D, 1, pc_range)
    assert 0 < int(geo.mask_u8.sum()) < geo.mask_u8.numel()


# ------------------------------------------------------------------ modules vs golden ------
def test_sca_module_golden():
    g = gu.load('sca_small')
    m = gu.build_sca(g, 'cuda', DEV)
    out, gq, gf = gu.run_sca(m, g, DEV)
    assert rel_err(out, g['out']) <= FWD
    assert rel_err(gq, g['grad_query']) <= BWD
    assert rel_err(gf, g['grad_feat']) <= BWD
    for n, p in m.named_parameters():
        assert rel_err(p.grad, gu.pgrads(g)[n]) <= BWD, n
    da = m.deformable_attention        # stand-alone contract: op-boundary path
    o = da(query=gu.T(g['da_query'], DEV), key=gu.T(g['da_value'], DEV), value=gu.T(g['da_value'], DEV),
           reference_points=gu.T(g['da_ref'], DEV), spatial_shapes=gu.T(g['shapes'], DEV),
           level_start_index=gu.T(g['starts'], DEV))
    assert rel_err(o, g['da_out']) <= FWD


@pytest.mark.parametrize('name', ['tsa_prev', 'tsa_first_frame'])
def test_tsa_module_golden(name):
    g = gu.load(name)
    m = gu.build_tsa(g, 'cuda', DEV)
    out, gq, gp = gu.run_tsa(m, g, DEV)
    assert rel_err(out, g['out']) <= FWD
    assert rel_err(gq, g['grad_query']) <= BWD
    if gp is not None:
        assert rel_err(gp, g['grad_prev']) <= BWD
    for n, p in m.named_parameters():
        assert rel_err(p.grad, gu.pgrads(g)[n]) <= BWD, n


def test_decoder_module_golden():
    g = gu.load('decoder_small')
    m = gu.build_decoder(g, 'cuda', DEV)
    out, gq, gv, gr = gu.run_decoder(m, g, DEV)
    assert rel_err(out, g['out']) <= FWD
    assert rel_err(gq, g['grad_query']) <= BWD
    assert rel_err(gv, g['grad_value']) <= BWD
    assert rel_err(gr, g['grad_ref']) <= BWD
    for n, p in m.named_parameters():
        assert rel_err(p.grad, gu.pgrads(g)[n]) <= BWD, n


# ------------------------------------------------------------------ modules vs oracle ------
def _sca_inputs(bs, H, W, levels, C, seed, scale=0.5):
    import apollo_vision_net_b200.synthetic as syn
    g = torch.Generator().manual_seed(seed)
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    l2i, img_shape = syn.camera_rig(scale, bs=bs, jitter=4.0, seed=seed)
    r3 = G.reference_points_3d(H, W, 8.0, 4, bs=bs)
    uv, mask = G.point_sampling(r3, syn.PC_RANGE, l2i, img_shape[0], img_shape[1])
    q = torch.randn(bs, H * W, C, generator=g)
    feat = torch.randn(6, Nk, bs, C, generator=g)
    go = torch.randn(bs, H * W, C, generator=g)
    return q, feat, go, uv, mask, torch.tensor(shapes_l), torch.tensor(starts_l)


@pytest.mark.parametrize('bs,H,W,levels', [
    (1, 50, 50, [(28, 48)]),                                      # BASELINE config 1 (tiny)
    (2, 30, 40, [(29, 50), (15, 25), (8, 13), (4, 7)]),           # 4 levels, bs=2 (batch-0 quirk)
])
def test_sca_module_vs_oracle(bs, H, W, levels):
    from apollo_vision_net_b200.modules import SpatialCrossAttention
    import apollo_vision_net_b200.synthetic as syn
    C = 256
    cfg = dict(embed_dims=C, pc_range=syn.PC_RANGE, batch_first=True,
               deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                         num_points=8, num_levels=len(levels)))
    o = OracleSpatialCrossAttention(**cfg)
    _randomize(o, 3)
    o.eval()
    m = SpatialCrossAttention(**cfg)
    m.load_state_dict(o.state_dict())
    m.to(DEV).eval()
    q, feat, go, uv, mask, shapes, starts = _sca_inputs(bs, H, W, levels, C, seed=11)
    q1, f1 = q.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    ref = o(q1, f1, f1, reference_points_cam=uv, bev_mask=mask, spatial_shapes=shapes,
            level_start_index=starts)
    ref.backward(go)
    q2, f2 = q.to(DEV).requires_grad_(True), feat.to(DEV).requires_grad_(True)
    out = m(q2, f2, f2, reference_points_cam=uv.to(DEV), bev_mask=mask.to(DEV),
            spatial_shapes=shapes.to(DEV), level_start_index=starts.to(DEV))
    out.backward(go.to(DEV))
    assert rel_err(out, ref) <= FWD
    assert rel_err(q2.grad, q1.grad) <= BWD
    assert rel_err(f2.grad, f1.grad) <= BWD
    og = dict(o.named_parameters())
    for n, p in m.named_parameters():
        assert rel_err(p.grad, og[n].grad) <= BWD, n


def test_tsa_module_vs_oracle_tiny_config():
    from apollo_vision_net_b200.modules import TemporalSelfAttention
    bs, H, W, C = 1, 50, 50, 256
    o = OracleTemporalSelfAttention(embed_dims=C, num_levels=1, num_points=4)
    _randomize(o, 5)
    o.eval()
    m = TemporalSelfAttention(embed_dims=C, num_levels=1, num_points=4)
    m.load_state_dict(o.state_dict())
    m.to(DEV).eval()
    g = torch.Generator().manual_seed(2)
    q = torch.randn(bs, H * W, C, generator=g)
    prev = torch.randn(bs * 2, H * W, C, generator=g)
    pos = torch.randn(bs, H * W, C, generator=g)
    go = torch.randn(bs, H * W, C, generator=g)
    r2 = G.reference_points_2d(H, W, bs=bs)
    hy = torch.stack([r2 + 0.013, r2], 1).reshape(bs * 2, H * W, 1, 2)
    kw = dict(spatial_shapes=torch.tensor([[H, W]]), level_start_index=torch.tensor([0]))
    q1, p1 = q.clone().requires_grad_(True), prev.clone().requires_grad_(True)
    ref = o(q1, p1, p1, query_pos=pos, reference_points=hy, **kw)
    ref.backward(go)
    q2, p2 = q.to(DEV).requires_grad_(True), prev.to(DEV).requires_grad_(True)
    out = m(q2, p2, p2, query_pos=pos.to(DEV), reference_points=hy.to(DEV),
            **{k: v.to(DEV) for k, v in kw.items()})
    out.backward(go.to(DEV))
    assert rel_err(out, ref) <= FWD
    assert rel_err(q2.grad, q1.grad) <= BWD
    assert rel_err(p2.grad, p1.grad) <= BWD
    og = dict(o.named_parameters())
    for n, p in m.named_parameters():
        assert rel_err(p.grad, og[n].grad) <= BWD, n


def test_maptrv2_cross_attention_config4():
    """BASELINE config 4: (50 + 300) vectors x 20 points = 7000 queries, L=1, P=4, BEV 50x50."""
    from apollo_vision_net_b200.modules import CustomMSDeformableAttention
    bs, H, W, C, Nq = 1, 50, 50, 256, 7000
    o = OracleCustomMSDeformableAttention(embed_dims=C, num_levels=1)
    _randomize(o, 8)
    o.eval()
    m = CustomMSDeformableAttention(embed_dims=C, num_levels=1)
    m.load_state_dict(o.state_dict())
    m.to(DEV).eval()
    g = torch.Generator().manual_seed(6)
    q = torch.randn(Nq, bs, C, generator=g)
    pos = torch.randn(Nq, bs, C, generator=g)
    val = torch.randn(H * W, bs, C, generator=g)
    rp = torch.rand(bs, Nq, 1, 2, generator=g)
    go = torch.randn(Nq, bs, C, generator=g)
    kw = dict(spatial_shapes=torch.tensor([[H, W]]), level_start_index=torch.tensor([0]))
    q1, v1 = q.clone().requires_grad_(True), val.clone().requires_grad_(True)
    ref = o(q1, None, v1, query_pos=pos, reference_points=rp, **kw)
    ref.backward(go)
    q2, v2 = q.to(DEV).requires_grad_(True), val.to(DEV).requires_grad_(True)
    out = m(q2, None, v2, query_pos=pos.to(DEV), reference_points=rp.to(DEV),
            **{k: v.to(DEV) for k, v in kw.items()})
    out.backward(go.to(DEV))
    assert rel_err(out, ref) <= FWD
    assert rel_err(q2.grad, q1.grad) <= BWD
    assert rel_err(v2.grad, v1.grad) <= BWD


def test_encoder_tiny_config1_vs_oracle():
    """BASELINE config 1: tiny encoder (3 layers, 50x50 BEV, 6 cams, 1 level, dim 256),
    forward + backward, against the oracle encoder (same state_dict)."""
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.synthetic as syn
    bs, H, W, C = 1, 50, 50, 256
    levels = syn.LEVELS_TINY
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    l2i, img_shape = syn.camera_rig(0.5, bs=bs)
    o = OracleBEVFormerEncoder(num_layers=3, pc_range=syn.PC_RANGE, num_points_in_pillar=4,
                               embed_dims=C, feedforward_channels=512, num_levels=1)
    _randomize(o, 12)
    o.eval()
    enc = pkg.build_transformer_layer_sequence(dict(
        type='BEVFormerEncoder', num_layers=3, pc_range=syn.PC_RANGE, num_points_in_pillar=4,
        return_intermediate=False,
        transformerlayers=dict(
            type='BEVFormerLayer',
            attn_cfgs=[dict(type='TemporalSelfAttention', embed_dims=C, num_levels=1),
                       dict(type='SpatialCrossAttention', pc_range=syn.PC_RANGE, embed_dims=C,
                            deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                                      num_points=8, num_levels=1))],
            feedforward_channels=512, ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    enc.load_state_dict(o.state_dict())
    enc.to(DEV).eval()
    g = torch.Generator().manual_seed(21)
    bevq = torch.randn(H * W, bs, C, generator=g)
    pos = torch.randn(H * W, bs, C, generator=g)
    prev = torch.randn(H * W, bs, C, generator=g)
    feat = torch.randn(6, Nk, bs, C, generator=g)
    go = torch.randn(bs, H * W, C, generator=g)
    shift = torch.tensor([[0.01, -0.02]])
    shapes, starts = torch.tensor(shapes_l), torch.tensor(starts_l)
    import copy
    o64 = copy.deepcopy(o).double()
    for prev_in in (prev, None):
        o.zero_grad()
        o64.zero_grad()
        enc.zero_grad()
        f1 = feat.clone().requires_grad_(True)
        ref = o(bevq, f1, f1, bev_h=H, bev_w=W, bev_pos=pos, spatial_shapes=shapes,
                level_start_index=starts, prev_bev=prev_in, shift=shift, lidar2img=l2i,
                img_h=img_shape[0], img_w=img_shape[1])
        ref.backward(go)
        # float64 run of the oracle: the truth; the fp32 oracle's own error is the yardstick
        # for a 3-layer-deep fp32 computation
        f0 = feat.double().requires_grad_(True)
        tru = o64(bevq.double(), f0, f0, bev_h=H, bev_w=W, bev_pos=pos.double(),
                  spatial_shapes=shapes, level_start_index=starts,
                  prev_bev=None if prev_in is None else prev_in.double(), shift=shift.double(),
                  lidar2img=l2i, img_h=img_shape[0], img_w=img_shape[1])
        tru.backward(go.double())
        f2 = feat.to(DEV).requires_grad_(True)
        metas = [dict(lidar2img=[l2i[b, i] for i in range(6)], img_shape=[img_shape] * 6)
                 for b in range(bs)]
        out = enc(bevq.to(DEV), f2, f2, bev_h=H, bev_w=W, bev_pos=pos.to(DEV),
                  spatial_shapes=shapes.to(DEV), level_start_index=starts.to(DEV),
                  prev_bev=None if prev_in is None else prev_in.to(DEV), shift=shift.to(DEV),
                  img_metas=metas)
        out.backward(go.to(DEV))

        def check(name, mine, oracle32, truth, tol):
            e, y = rel_err(mine, truth), rel_err(oracle32, truth)
            assert e <= max(tol, 5.0 * y), f'{name}: err {e:.3e} vs fp64 truth (fp32 oracle: {y:.3e})'

        def check_grad(name, mine, oracle32, truth):
            """Gradients through three layers: d out / d location is piecewise constant (it jumps
            when a sample crosses a pixel boundary), so rounding-level differences between the GPU
            and the CPU forward of the upstream layers put a handful of samples into a neighbouring
            pixel pair, and with only 2 500
            queries one such sample moves single entries of a weight gradient by percents of the
            largest entry (the fp32 oracle shows the same against its own fp64 run).  The bound is
            therefore tight on the gradient as a whole (2-norm) and loose on the single worst entry;
            the single-layer tests hold the 1e-4 max-norm bar, and the encoder's forward is bit-identical
            and its backward reproducible to 1e-6 from run to run and process to process (no races).
            The model instance is fixed by the per-test seed of conftest.py."""
            e2, y2 = rel_l2(mine, truth), rel_l2(oracle32, truth)
            assert e2 <= max(5e-3, 5.0 * y2), f'{name}: l2 err {e2:.3e} (fp32 oracle: {y2:.3e})'
            check(name, mine, oracle32, truth, 5e-2)

        check('out', out, ref, tru, 2e-5)
        check_grad('grad_feat', f2.grad, f1.grad, f0.grad)
        og, tg = dict(o.named_parameters()), dict(o64.named_parameters())
        for n, p in enc.named_parameters():
            check_grad(n, p.grad, og[n].grad, tg[n].grad)


def test_encoder_golden():
    """The CUDA BEVFormerEncoder (2 layers, history + CAN-bus shift, 2 feature levels, batch of 2 with
    jittered cameras) against the fixture produced by the UNMODIFIED reference encoder classes
    (encoder.py:243-519 on custom_base_transformer_layer.py), forward and gradients."""
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.synthetic as syn
    g = gu.load('encoder_small')
    bs, H, W, C, heads = (int(x) for x in g['cfg'])
    levels = [tuple(int(v) for v in r) for r in g['levels']]
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    enc = pkg.build_transformer_layer_sequence(dict(
        type='BEVFormerEncoder', num_layers=2, pc_range=syn.PC_RANGE, num_points_in_pillar=4,
        return_intermediate=False,
        transformerlayers=dict(
            type='BEVFormerLayer',
            attn_cfgs=[dict(type='TemporalSelfAttention', embed_dims=C, num_heads=heads, num_levels=1,
                            num_points=4),
                       dict(type='SpatialCrossAttention', pc_range=syn.PC_RANGE, embed_dims=C,
                            deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                                      num_heads=heads, num_points=8,
                                                      num_levels=len(levels)))],
            feedforward_channels=2 * C, ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    enc.load_state_dict(gu.params(g))
    enc.to(DEV).eval()
    l2i = g['lidar2img']
    img_shape = tuple(int(v) for v in g['img_shape'])
    metas = [dict(lidar2img=[l2i[b, i] for i in range(6)], img_shape=[img_shape] * 6) for b in range(bs)]
    bevq, prev, feat = (gu.T(g[k], DEV, True) for k in ('bev_query', 'prev_bev', 'feat'))
    out = enc(bevq, feat, feat, bev_h=H, bev_w=W, bev_pos=gu.T(g['bev_pos'], DEV),
              spatial_shapes=torch.tensor(shapes_l, device=DEV),
              level_start_index=torch.tensor(starts_l, device=DEV), prev_bev=prev,
              shift=gu.T(g['shift'], DEV), img_metas=metas)
    out.backward(gu.T(g['grad_out'], DEV))
    assert rel_err(out, g['out']) <= 1e-4
    assert rel_l2(bevq.grad, gu.T(g['grad_bev_query'])) <= 5e-3
    assert rel_l2(prev.grad, gu.T(g['grad_prev_bev'])) <= 5e-3
    assert rel_l2(feat.grad, gu.T(g['grad_feat'])) <= 5e-3
    pg = gu.pgrads(g)
    for n, p in enc.named_parameters():
        assert rel_l2(p.grad, gu.T(pg[n])) <= 5e-3, n


def test_fused_bf16_within_tolerance():
    """bf16 value path of the fused SCA kernel: within 1e-2 of the fp32 oracle (north_star)."""
    from apollo_vision_net_b200.modules import SpatialCrossAttention
    import apollo_vision_net_b200.synthetic as syn
    C = 256
    levels = [(29, 50), (15, 25)]
    cfg = dict(embed_dims=C, pc_range=syn.PC_RANGE, batch_first=True,
               deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                         num_points=8, num_levels=len(levels)))
    o = OracleSpatialCrossAttention(**cfg)
    _randomize(o, 3)
    o.eval()
    m = SpatialCrossAttention(**cfg)
    m.load_state_dict(o.state_dict())
    m.to(DEV).to(torch.bfloat16).eval()
    q, feat, go, uv, mask, shapes, starts = _sca_inputs(1, 30, 30, levels, C, seed=4)
    ref = o(q, feat, feat, reference_points_cam=uv, bev_mask=mask, spatial_shapes=shapes,
            level_start_index=starts)
    out = m(q.to(DEV).bfloat16(), feat.to(DEV).bfloat16(), feat.to(DEV).bfloat16(),
            reference_points_cam=uv.to(DEV), bev_mask=mask.to(DEV),
            spatial_shapes=shapes.to(DEV), level_start_index=starts.to(DEV))
    assert out.dtype == torch.bfloat16
    # compare the attention contribution (output minus the residual query) at bf16 resolution
    assert rel_err(out.float().cpu() - q.bfloat16().float(), ref - q) <= 3e-2
    assert rel_err(out, ref) <= 1e-2


@pytest.mark.parametrize('accum', ['fp16', 'fp32'])
def test_fused_bf16_backward_accumulators(accum, monkeypatch):
    """bf16 fused backward on identical (bf16-rounded) inputs: the scaled fp16 grad_value
    accumulator and the fp32 one both stay within bf16 resolution of the fp32 oracle gradient,
    also for tiny upstream gradients (the scale is derived on the device from max|g_out|)."""
    from apollo_vision_net_b200.fused_ops import QueueDeformAttnFunction
    from oracle.msda_oracle import msda_torch
    import apollo_vision_net_b200.fused_ops as fo
    monkeypatch.setattr(fo, '_accum_mode', [accum])
    g = torch.Generator().manual_seed(31)
    bs, H, W, M, Dh, L, P, Nq = 1, 24, 40, 8, 32, 1, 4, 3000
    value = torch.randn(bs, H * W, M, Dh, generator=g).bfloat16()
    offsets = torch.randn(bs, Nq, M, 1, L, P, 2, generator=g) * 3
    logits = torch.randn(bs, Nq, M, 1, L * P, generator=g)
    ref = torch.rand(bs, Nq, L, 2, generator=g)
    shapes = torch.tensor([[H, W]])
    starts = torch.tensor([0])
    for gscale in (1.0, 1e-6):
        g_out = (torch.randn(bs, Nq, M * Dh, generator=g) * gscale).bfloat16()
        v1 = value.float().requires_grad_(True)
        loc = ref[:, :, None, :, None, :] + offsets[:, :, :, 0] / torch.tensor([W, H]).view(1, 1, 1, 1, 1, 2)
        att = logits[:, :, :, 0].softmax(-1).view(bs, Nq, M, L, P)
        msda_torch(v1, shapes, loc, att).backward(g_out.float())
        v2 = value.to(DEV).requires_grad_(True)
        out = QueueDeformAttnFunction.apply(v2, shapes.to(DEV), starts.to(DEV), offsets.to(DEV),
                                            logits.to(DEV), ref.to(DEV), None, 0)
        out.backward(g_out.to(DEV))
        assert torch.isfinite(v2.grad).all()
        assert rel_err(v2.grad, v1.grad) <= 1e-2, (accum, gscale)


# ------------------------------------------------------------------ row-wise companions ----
@pytest.mark.parametrize('dtype,C', [(torch.float32, 256), (torch.bfloat16, 256), (torch.float32, 512),
                                     (torch.bfloat16, 512)])
def test_layernorm_and_linear_kernels(dtype, C):
    """LayerNorm fwd/bwd and the bias-gradient column sum against torch's fp32/fp64 CPU math."""
    from apollo_vision_net_b200.rowops import LayerNorm, Linear
    g = torch.Generator().manual_seed(5)
    rows = 3001
    x = torch.randn(rows, C, generator=g) * 2 + 0.5
    go = torch.randn(rows, C, generator=g)
    ln_ref = torch.nn.LayerNorm(C).double()
    ln_ref.weight.data = torch.randn(C, generator=g).double()
    ln_ref.bias.data = torch.randn(C, generator=g).double()
    ln = LayerNorm(C)
    ln.load_state_dict({k: v.float() for k, v in ln_ref.state_dict().items()})
    ln.to(DEV).to(dtype)
    xr = x.to(dtype).double().requires_grad_(True)
    yr = ln_ref(xr)
    yr.backward(go.to(dtype).double())
    xg = x.to(DEV).to(dtype).requires_grad_(True)
    y = ln(xg)
    y.backward(go.to(DEV).to(dtype))
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(y, yr) <= tol
    assert rel_err(xg.grad, xr.grad) <= (1e-4 if dtype == torch.float32 else 2e-2)
    assert rel_err(ln.weight.grad, ln_ref.weight.grad) <= (1e-4 if dtype == torch.float32 else 2e-2)
    assert rel_err(ln.bias.grad, ln_ref.bias.grad) <= (1e-4 if dtype == torch.float32 else 2e-2)
    lin = Linear(C, 128).to(DEV).to(dtype)
    x2 = x.to(DEV).to(dtype).requires_grad_(True)
    out = lin(x2)
    g2 = torch.randn(rows, 128, generator=g).to(DEV).to(dtype)
    out.backward(g2)
    assert rel_err(lin.bias.grad, g2.double().sum(0)) <= (1e-5 if dtype == torch.float32 else 1e-2)
    assert rel_err(lin.weight.grad, g2.double().t() @ x2.detach().double()) <= (1e-4 if dtype == torch.float32 else 2e-2)


def test_row_sharded_encoder_matches_full_on_one_gpu():
    """BEV row sharding (parallel.py) with the CUDA encoder: the ranks of a 3-way split, run one
    after the other on one GPU, reproduce the unsharded forward (no collective involved here;
    the gloo tests cover the all-gather)."""
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.synthetic as syn
    bs, H, W, C = 1, 23, 30, 256
    levels = [(29, 50), (15, 25)]
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    l2i, img_shape = syn.camera_rig(0.5, bs=bs)
    enc = pkg.build_transformer_layer_sequence(dict(
        type='BEVFormerEncoder', num_layers=2, pc_range=syn.PC_RANGE, num_points_in_pillar=4,
        transformerlayers=dict(
            type='BEVFormerLayer',
            attn_cfgs=[dict(type='TemporalSelfAttention', embed_dims=C, num_levels=1),
                       dict(type='SpatialCrossAttention', pc_range=syn.PC_RANGE, embed_dims=C,
                            deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                                      num_points=8, num_levels=len(levels)))],
            feedforward_channels=512, ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    _randomize(enc, 2)
    enc.to(DEV).eval()
    g = torch.Generator().manual_seed(9)
    bevq, pos, prev = (torch.randn(H * W, bs, C, generator=g).to(DEV) for _ in range(3))
    feat = torch.randn(6, Nk, bs, C, generator=g).to(DEV)
    kw = dict(bev_h=H, bev_w=W, bev_pos=pos, spatial_shapes=torch.tensor(shapes_l, device=DEV),
              level_start_index=torch.tensor(starts_l, device=DEV), prev_bev=prev,
              shift=torch.tensor([[0.01, 0.02]], device=DEV), lidar2img=l2i, img_shape=img_shape)
    with torch.no_grad():
        full = enc(bevq, feat, feat, **kw)
        parts = [enc(bevq, feat, feat, row_shard=(r, 3), **kw) for r in range(3)]
    got = torch.cat(parts, 1)
    assert got.shape == full.shape
    assert rel_err(got, full) <= 1e-5
    with pytest.raises(RuntimeError, match='prev_bev'):
        enc(bevq, feat, feat, row_shard=(0, 2), **dict(kw, prev_bev=None))


def test_row_sharded_encoder_backward_sums_to_the_full_gradients_on_one_gpu():
    """Training with sharded BEV rows: the ranks of a 3-way split, run one after the
    other on one GPU with their rows of the upstream gradient, produce partial gradients of the parameters,
    the image features, the history and the BEV queries whose SUM equals the unsharded backward (the sum is
    what parallel.replicated / allreduce_gradients compute over NCCL; gloo tests cover the collectives)."""
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.synthetic as syn
    from apollo_vision_net_b200.parallel import bev_query_range
    bs, H, W, C = 1, 24, 30, 256
    levels = [(29, 50), (15, 25)]
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    l2i, img_shape = syn.camera_rig(0.5, bs=bs)
    enc = pkg.build_transformer_layer_sequence(dict(
        type='BEVFormerEncoder', num_layers=2, pc_range=syn.PC_RANGE, num_points_in_pillar=4,
        transformerlayers=dict(
            type='BEVFormerLayer',
            attn_cfgs=[dict(type='TemporalSelfAttention', embed_dims=C, num_levels=1, dropout=0.0),
                       dict(type='SpatialCrossAttention', pc_range=syn.PC_RANGE, embed_dims=C, dropout=0.0,
                            deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                                      num_points=8, num_levels=len(levels)))],
            feedforward_channels=512, ffn_dropout=0.0,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    _randomize(enc, 2)
    enc.to(DEV).train()
    g = torch.Generator().manual_seed(9)
    bevq, pos, prev = (torch.randn(H * W, bs, C, generator=g).to(DEV) for _ in range(3))
    feat = torch.randn(6, Nk, bs, C, generator=g).to(DEV)
    go = torch.randn(bs, H * W, C, generator=g).to(DEV)
    kw = dict(bev_h=H, bev_w=W, spatial_shapes=torch.tensor(shapes_l, device=DEV),
              level_start_index=torch.tensor(starts_l, device=DEV),
              shift=torch.tensor([[0.01, 0.02]], device=DEV), lidar2img=l2i, img_shape=img_shape)
    params = [p for p in enc.parameters() if p.requires_grad]

    def run(shard):
        for p in params:
            p.grad = None
        q, ps, pv, f = (t.clone().requires_grad_(True) for t in (bevq, pos, prev, feat))
        out = enc(q, f, f, bev_pos=ps, prev_bev=pv, row_shard=shard, **kw)
        if shard is None:
            out.backward(go)
        else:
            q0, q1 = bev_query_range(H, W, *shard)
            out.backward(go[:, q0:q1])
        return [q.grad, ps.grad, pv.grad, f.grad] + [p.grad.clone() for p in params]

    full = run(None)
    total = None
    for r in range(3):
        part = run((r, 3))
        total = part if total is None else [a + b for a, b in zip(total, part)]
    for i, (a, b) in enumerate(zip(total, full)):
        assert rel_err(a, b) <= 2e-4, (i, rel_err(a, b))


# ------------------------------------------------------------------ more shapes / paths ----
@pytest.mark.parametrize('C,heads,D,P', [(128, 8, 4, 8), (64, 8, 2, 4), (256, 8, 1, 4)])
def test_sca_generic_head_dims_and_anchor_counts(C, heads, D, P):
    """The shared SCA of the voxel / hybrid configs: Dh = 16 / 8, D = num_points_in_voxel other
    than 4, bs = 2."""
    from apollo_vision_net_b200.modules import SpatialCrossAttention
    import apollo_vision_net_b200.synthetic as syn
    levels = [(15, 25), (8, 13)]
    cfg = dict(embed_dims=C, pc_range=syn.PC_RANGE, batch_first=True,
               deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C, num_heads=heads,
                                         num_points=P, num_levels=len(levels)))
    o = OracleSpatialCrossAttention(**cfg)
    _randomize(o, 1)
    o.eval()
    m = SpatialCrossAttention(**cfg)
    m.load_state_dict(o.state_dict())
    m.to(DEV).eval()
    bs, H, W = 2, 18, 22
    g = torch.Generator().manual_seed(13)
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    l2i, img_shape = syn.camera_rig(0.5, bs=bs, jitter=4.0, seed=2)
    r3 = G.reference_points_3d(H, W, 8.0, D, bs=bs)
    uv, mask = G.point_sampling(r3, syn.PC_RANGE, l2i, img_shape[0], img_shape[1])
    q = torch.randn(bs, H * W, C, generator=g)
    feat = torch.randn(6, Nk, bs, C, generator=g)
    go = torch.randn(bs, H * W, C, generator=g)
    shapes, starts = torch.tensor(shapes_l), torch.tensor(starts_l)
    q1, f1 = q.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    ref = o(q1, f1, f1, reference_points_cam=uv, bev_mask=mask, spatial_shapes=shapes,
            level_start_index=starts)
    ref.backward(go)
    q2, f2 = q.to(DEV).requires_grad_(True), feat.to(DEV).requires_grad_(True)
    out = m(q2, f2, f2, reference_points_cam=uv.to(DEV), bev_mask=mask.to(DEV),
            spatial_shapes=shapes.to(DEV), level_start_index=starts.to(DEV), bev_h=H, bev_w=W)
    out.backward(go.to(DEV))
    assert rel_err(out, ref) <= FWD
    assert rel_err(q2.grad, q1.grad) <= BWD
    assert rel_err(f2.grad, f1.grad) <= BWD


def test_decoder_box_reference_points_and_padding_mask():
    """4-d (box) reference points and key_padding_mask: the operator path of decoder.py:301-333."""
    from apollo_vision_net_b200.modules import CustomMSDeformableAttention
    bs, H, W, C, Nq = 2, 12, 9, 256, 40
    o = OracleCustomMSDeformableAttention(embed_dims=C, num_levels=1, attn_logits_clamp=2.0)
    _randomize(o, 4)
    o.eval()
    m = CustomMSDeformableAttention(embed_dims=C, num_levels=1, attn_logits_clamp=2.0)
    m.load_state_dict(o.state_dict())
    m.to(DEV).eval()
    g = torch.Generator().manual_seed(17)
    q = torch.randn(Nq, bs, C, generator=g)
    val = torch.randn(H * W, bs, C, generator=g)
    rp = torch.rand(bs, Nq, 1, 4, generator=g) * 0.5 + 0.25
    kpm = torch.rand(bs, H * W, generator=g) < 0.1
    go = torch.randn(Nq, bs, C, generator=g)
    kw = dict(spatial_shapes=torch.tensor([[H, W]]), level_start_index=torch.tensor([0]))
    q1, v1 = q.clone().requires_grad_(True), val.clone().requires_grad_(True)
    ref = o(q1, None, v1, reference_points=rp, key_padding_mask=kpm, **kw)
    ref.backward(go)
    q2, v2 = q.to(DEV).requires_grad_(True), val.to(DEV).requires_grad_(True)
    out = m(q2, None, v2, reference_points=rp.to(DEV), key_padding_mask=kpm.to(DEV),
            **{k: v.to(DEV) for k, v in kw.items()})
    out.backward(go.to(DEV))
    assert rel_err(out, ref) <= FWD
    assert rel_err(q2.grad, q1.grad) <= BWD
    assert rel_err(v2.grad, v1.grad) <= BWD


def test_maptrv2_decoder_stack_runs_config4():
    """MapTRv2Decoder (6 decoupled layers, 350 vectors x 20 points, one-to-many mask) built from
    the reference's config fragment (bev_tiny_det_mapv2.py:33-63): runs forward + backward on the
    fused cross-attention; checks shapes, finiteness and the refinement of the reference points."""
    import apollo_vision_net_b200 as pkg
    C, V, Pn, bs, H, W = 256, 350, 20, 1, 50, 50
    dec = pkg.build_transformer_layer_sequence(dict(
        type='MapTRv2Decoder', num_layers=6, return_intermediate=True,
        transformerlayers=dict(
            type='MapTRv2DecoupledDetrTransformerDecoderLayer', num_vec=V, num_pts_per_vec=Pn,
            attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                       dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                       dict(type='CustomMSDeformableAttention', embed_dims=C, num_levels=1)],
            feedforward_channels=512, ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    _randomize(dec, 6)
    dec.to(DEV).eval()
    g = torch.Generator().manual_seed(3)
    query = torch.randn(V * Pn, bs, C, generator=g).to(DEV).requires_grad_(True)
    qpos = torch.randn(V * Pn, bs, C, generator=g).to(DEV)
    bev = torch.randn(H * W, bs, C, generator=g).to(DEV).requires_grad_(True)
    refp = torch.rand(bs, V * Pn, 2, generator=g).to(DEV)
    mask = torch.zeros(V, V, dtype=torch.bool, device=DEV)
    mask[50:, :50] = True
    mask[:50, 50:] = True                                     # one2one / one2many separation
    reg = torch.nn.ModuleList([torch.nn.Linear(C, 2) for _ in range(6)]).to(DEV)
    inter, refs = dec(query, key=None, value=bev, query_pos=qpos, reference_points=refp,
                      reg_branches=reg, spatial_shapes=torch.tensor([[H, W]], device=DEV),
                      level_start_index=torch.tensor([0], device=DEV), self_attn_mask=mask,
                      num_vec=V, num_pts_per_vec=Pn)
    assert inter.shape == (6, V * Pn, bs, C) and refs.shape == (6, bs, V * Pn, 2)
    assert torch.isfinite(inter).all() and (refs >= 0).all() and (refs <= 1).all()
    inter[-1].float().pow(2).mean().backward()
    assert torch.isfinite(bev.grad).all() and bev.grad.abs().sum() > 0
    assert torch.isfinite(query.grad).all()


def test_maptrv2_decoder_golden():
    """MapTRv2Decoder with the CUDA cross-attention (hoisted value projections, fused kernel)
    against the fixture of the unmodified reference decoder (2 layers, one-to-many mask)."""
    g = gu.load('maptrv2_decoder_small')
    dec, reg = gu.build_maptrv2_decoder(g, 'b200', DEV)
    inter, refs, gq, gv = gu.run_maptrv2_decoder(dec, reg, g, DEV)
    assert rel_err(inter, g['inter']) <= 10 * FWD
    assert rel_err(refs, g['refs']) <= 10 * FWD
    assert rel_err(gq, g['grad_query']) <= 10 * BWD
    assert rel_err(gv, g['grad_value']) <= 10 * BWD


def test_detection_decoder_golden():
    """DetectionTransformerDecoder + CUDA CustomMSDeformableAttention (value projections hoisted
    into one batched GEMM) against the fixture of the unmodified reference class."""
    g = gu.load('det_decoder_small')
    dec, reg = gu.build_det_decoder(g, 'b200', DEV)
    inter, refs, gq, gv = gu.run_det_decoder(dec, reg, g, DEV)
    assert rel_err(inter, g['inter']) <= 10 * FWD      # three stacked layers
    assert rel_err(refs, g['refs']) <= 10 * FWD
    assert rel_err(gq, g['grad_query']) <= 10 * BWD
    assert rel_err(gv, g['grad_value']) <= 10 * BWD


def test_detection_decoder_vs_oracle_cross_attention():
    """DetectionTransformerDecoder (reference decoder.py:50-126): 3 DetrTransformerDecoderLayers,
    900-query style stack at a small size, 3-d reference points refined by reg_branches.  The
    CPU twin is the same stack with every cross-attention replaced by the oracle module (which
    projects the value per layer, as the reference does) -- outputs, refined points and gradients
    must agree within the fp32 tolerances."""
    import copy
    import apollo_vision_net_b200 as pkg
    C, Nq, bs, H, W = 256, 90, 2, 25, 20
    dec = pkg.build_transformer_layer_sequence(dict(
        type='DetectionTransformerDecoder', num_layers=3, return_intermediate=True,
        transformerlayers=dict(
            type='DetrTransformerDecoderLayer',
            attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                       dict(type='CustomMSDeformableAttention', embed_dims=C, num_levels=1)],
            feedforward_channels=512, ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    _randomize(dec, 21)
    dec.eval()
    twin = copy.deepcopy(dec)
    for layer in twin.layers:
        o = OracleCustomMSDeformableAttention(embed_dims=C, num_levels=1)
        o.load_state_dict(layer.attentions[1].state_dict())
        layer.attentions[1] = o.eval()
    reg = torch.nn.ModuleList([torch.nn.Linear(C, 10) for _ in range(3)])
    reg_gpu = copy.deepcopy(reg).to(DEV)
    dec.to(DEV)
    g = torch.Generator().manual_seed(23)
    query = torch.randn(Nq, bs, C, generator=g)
    qpos = torch.randn(Nq, bs, C, generator=g)
    bev = torch.randn(H * W, bs, C, generator=g)
    refp = torch.rand(bs, Nq, 3, generator=g)
    go = torch.randn(3, Nq, bs, C, generator=g)
    shapes, starts = torch.tensor([[H, W]]), torch.tensor([0])

    q1, b1 = query.clone().requires_grad_(True), bev.clone().requires_grad_(True)
    ref_out, ref_pts = twin(q1, key=None, value=b1, query_pos=qpos, reference_points=refp,
                            reg_branches=reg, spatial_shapes=shapes, level_start_index=starts)
    ref_out.backward(go)
    q2, b2 = query.to(DEV).requires_grad_(True), bev.to(DEV).requires_grad_(True)
    out, pts = dec(q2, key=None, value=b2, query_pos=qpos.to(DEV), reference_points=refp.to(DEV),
                   reg_branches=reg_gpu, spatial_shapes=shapes.to(DEV),
                   level_start_index=starts.to(DEV))
    out.backward(go.to(DEV))
    assert out.shape == (3, Nq, bs, C) and pts.shape == (3, bs, Nq, 3)
    assert rel_err(out, ref_out) <= 10 * FWD           # three stacked layers of fp32 GEMMs + softmax
    assert torch.allclose(pts.cpu(), ref_pts, atol=1e-5)
    assert rel_err(q2.grad, q1.grad) <= 10 * BWD and rel_err(b2.grad, b1.grad) <= 10 * BWD
    for (n, p), (_, pr) in zip(dec.layers[0].attentions[1].named_parameters(),
                               twin.layers[0].attentions[1].named_parameters()):
        assert rel_err(p.grad, pr.grad) <= 10 * BWD, n


def test_maptrv2_decoder_hoisted_value_projection_matches_per_layer():
    """The MapTRv2 decoder projects the BEV for all six cross-attentions in one batched GEMM before
    the layer loop; outputs, refined reference points and every
    gradient equal the per-layer projection of the reference (decoder.py:299-303)."""
    import apollo_vision_net_b200 as pkg
    C, V, Pn, bs, H, W = 256, 12, 20, 2, 30, 20
    dec = pkg.build_transformer_layer_sequence(dict(
        type='MapTRv2Decoder', num_layers=3, return_intermediate=True,
        transformerlayers=dict(
            type='MapTRv2DecoupledDetrTransformerDecoderLayer', num_vec=V, num_pts_per_vec=Pn,
            attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                       dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                       dict(type='CustomMSDeformableAttention', embed_dims=C, num_levels=1)],
            feedforward_channels=512, ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    _randomize(dec, 8)
    dec.to(DEV).eval()
    g = torch.Generator().manual_seed(11)
    query0 = torch.randn(V * Pn, bs, C, generator=g).to(DEV)
    qpos = torch.randn(V * Pn, bs, C, generator=g).to(DEV)
    bev0 = torch.randn(H * W, bs, C, generator=g).to(DEV)
    refp = torch.rand(bs, V * Pn, 2, generator=g).to(DEV)
    kpm = (torch.rand(bs, H * W, generator=g) < 0.05).to(DEV)
    reg = torch.nn.ModuleList([torch.nn.Linear(C, 2) for _ in range(3)]).to(DEV)
    go = torch.randn(3, V * Pn, bs, C, generator=g).to(DEV)

    def run(hoist):
        dec.hoist_value_proj = hoist
        dec.zero_grad()
        query = query0.clone().requires_grad_(True)
        bev = bev0.clone().requires_grad_(True)
        inter, refs = dec(query, key=None, value=bev, query_pos=qpos, reference_points=refp,
                          reg_branches=reg, key_padding_mask=kpm,
                          spatial_shapes=torch.tensor([[H, W]], device=DEV),
                          level_start_index=torch.tensor([0], device=DEV), num_vec=V, num_pts_per_vec=Pn)
        inter.backward(go)
        grads = {n: p.grad.clone() for n, p in dec.named_parameters() if p.grad is not None}
        return inter.detach(), refs, query.grad, bev.grad, grads

    a = run(True)
    b = run(False)
    assert rel_err(a[0], b[0]) <= FWD and torch.allclose(a[1], b[1], atol=1e-6)
    assert rel_err(a[2], b[2]) <= BWD and rel_err(a[3], b[3]) <= BWD
    assert a[4].keys() == b[4].keys() and any('value_proj.weight' in k for k in a[4])
    for k in a[4]:
        assert rel_err(a[4][k], b[4][k]) <= BWD, k


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_merged_projection_layout_matches_separate_tensors(dtype):
    """The fused Functions take offsets and logits either as two tensors or as the column blocks
    of ONE (bs, Nq, 3n) tensor (single GEMM over the concatenated weights, SURVEY.md section 8f
    rank 1): same forward bit for bit, same gradients, and the merged gradient comes back in the
    merged layout."""
    import apollo_vision_net_b200.fused_ops as fo
    import apollo_vision_net_b200.synthetic as syn
    from oracle import geometry_oracle as G
    g = torch.Generator().manual_seed(4)
    # spatial cross-attention
    bs, H, W, M, Dh, P, D = 2, 11, 13, 8, 32, 8, 4
    levels = [(29, 50), (15, 25)]
    L = len(levels)
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    shapes, starts = torch.tensor(shapes_l).to(DEV), torch.tensor(starts_l).to(DEV)
    l2i, img_shape = syn.camera_rig(0.5, bs=bs, jitter=4.0, seed=2)
    r3 = G.reference_points_3d(H, W, 8.0, D, bs=bs)
    geo = fo.bev_point_sampling(r3.to(DEV), syn.PC_RANGE, l2i, img_shape[0], img_shape[1])
    n = M * L * P
    value = torch.randn(bs * 6, Nk, M, Dh, generator=g).to(dtype).to(DEV)
    merged = torch.randn(bs, H * W, 3 * n, generator=g)
    merged[..., :2 * n] *= 5.0
    merged = merged.to(dtype).to(DEV)
    go = torch.randn(bs, H * W, M * Dh, generator=g).to(dtype).to(DEV)

    def run_sca(as_merged):
        v = value.clone().requires_grad_(True)
        mg = merged.clone().requires_grad_(True)
        if as_merged:
            out = fo.SpatialCrossAttnFunction.apply(v, shapes, starts, mg, None,
                                                    geo.reference_points_cam, geo.mask_u8,
                                                    geo.hit_bits, 6, W)
        else:
            out = fo.SpatialCrossAttnFunction.apply(
                v, shapes, starts, mg[..., :2 * n].reshape(bs, H * W, M, L, P, 2),
                mg[..., 2 * n:].reshape(bs, H * W, M, L * P), geo.reference_points_cam,
                geo.mask_u8, geo.hit_bits, 6, W)
        out.backward(go)
        return out.detach(), v.grad, mg.grad

    a, b = run_sca(True), run_sca(False)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert torch.equal(a[0], b[0])
    assert a[2].shape == merged.shape
    assert rel_err(a[1], b[1]) <= tol and rel_err(a[2], b[2]) <= tol

    # temporal self-attention layout (queue of 2) with reference-point gradients
    Q, Pt, Hb, Wb = 2, 4, 9, 7
    Nq = Hb * Wb
    tshape, tstart = torch.tensor([[Hb, Wb]]).to(DEV), torch.tensor([0]).to(DEV)
    nt = M * Q * Pt
    tv = torch.randn(Q, Nq, M, Dh, generator=g).to(dtype).to(DEV)
    tm = (torch.randn(1, Nq, 3 * nt, generator=g) * 2.0).to(dtype).to(DEV)
    ref = torch.rand(Q, Nq, 1, 2, generator=g).to(DEV)
    tgo = torch.randn(1, Nq, M * Dh, generator=g).to(dtype).to(DEV)

    def run_tsa(as_merged):
        v = tv.clone().requires_grad_(True)
        mg = tm.clone().requires_grad_(True)
        r = ref.clone().requires_grad_(True)
        if as_merged:
            out = fo.QueueDeformAttnFunction.apply(v, tshape, tstart, mg, None, r, 3.0, Wb)
        else:
            out = fo.QueueDeformAttnFunction.apply(
                v, tshape, tstart, mg[..., :2 * nt].reshape(1, Nq, M, Q, 1, Pt, 2),
                mg[..., 2 * nt:].reshape(1, Nq, M, Q, Pt), r, 3.0, Wb)
        out.backward(tgo)
        return out.detach(), v.grad, mg.grad, r.grad

    a, b = run_tsa(True), run_tsa(False)
    assert torch.equal(a[0], b[0])
    for x, y in zip(a[1:], b[1:]):
        assert rel_err(x, y) <= tol


@pytest.mark.parametrize('dtype,rows,Cin,C', [(torch.float32, 1000, 256, 256), (torch.bfloat16, 4099, 512, 256),
                                              (torch.float16, 77, 256, 512), (torch.float32, 1, 128, 128)])
def test_linear_add_layernorm_matches_composition(dtype, rows, Cin, C):
    """The fused post-norm tail y = LN(x W^T + b + residual) against the same three steps run
    separately (Linear, add, LayerNorm kernels): identical forward, gradients within rounding,
    and against plain torch in fp32."""
    import apollo_vision_net_b200.rowops as ro
    g = torch.Generator().manual_seed(31)
    lin = ro.Linear(Cin, C).to(DEV).to(dtype)
    norm = ro.LayerNorm(C).to(DEV).to(dtype)
    with torch.no_grad():
        norm.weight.copy_((torch.rand(C, generator=g) + 0.5).to(dtype))
        norm.bias.copy_(torch.randn(C, generator=g).to(dtype))
    x = torch.randn(2, rows, Cin, generator=g).to(dtype).to(DEV)
    res = torch.randn(2, rows, C, generator=g).to(dtype).to(DEV)
    go = torch.randn(2, rows, C, generator=g).to(dtype).to(DEV)

    def run(fused):
        for p in list(lin.parameters()) + list(norm.parameters()):
            p.grad = None
        x1, r1 = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
        y = ro.linear_add_layernorm(x1, lin, r1, norm) if fused else norm(lin(x1) + r1)
        y.backward(go)
        return [y.detach(), x1.grad, r1.grad, lin.weight.grad, lin.bias.grad, norm.weight.grad,
                norm.bias.grad]

    a, b = run(True), run(False)
    assert torch.equal(a[0], b[0])
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for u, v in zip(a[1:], b[1:]):
        assert u.shape == v.shape and rel_err(u, v) <= tol
    if dtype == torch.float32:
        x1, r1 = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
        y = torch.nn.functional.layer_norm(torch.nn.functional.linear(x1, lin.weight, lin.bias) + r1,
                                           (C,), norm.weight, norm.bias, norm.eps)
        assert rel_err(a[0], y) <= 1e-5


@pytest.mark.parametrize('N,O,I,dtype', [(1000, 256, 256, torch.bfloat16), (37, 192, 512, torch.bfloat16),
                                         (4099, 768, 256, torch.float16), (1, 8, 8, torch.bfloat16),
                                         (40000, 256, 256, torch.bfloat16), (20011, 512, 264, torch.float16)])
def test_linear_wgrad_tcgen05(N, O, I, dtype):
    """linear_wgrad (tcgen05.mma with the tile in tensor memory, rows split over the grid, bias sums
    fused): dW = dy^T x and db = column sums of dy against float64, with the error of the library GEMM
    on the same 16-bit inputs as the yardstick; the scratch must come back zeroed."""
    from apollo_vision_net_b200 import _lib
    from apollo_vision_net_b200.multi_scale_deformable_attn_function import _DTYPE_CODE
    g = torch.Generator().manual_seed(N + O + I)
    dy = torch.randn(N, O, generator=g).to(dtype).to(DEV)
    x = torch.randn(N, I, generator=g).to(dtype).to(DEV)
    dW = torch.full((O, I), float('nan'), dtype=dtype, device=DEV)
    db = torch.full((O,), float('nan'), dtype=dtype, device=DEV)
    ws = torch.zeros(int(_lib.lib().linear_wgrad_workspace_floats(O, I)), device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):                                   # twice: the scratch is reused without a memset
        _lib.call('linear_wgrad', dy.data_ptr(), x.data_ptr(), dW.data_ptr(), db.data_ptr(), ws.data_ptr(),
                  N, O, I, _DTYPE_CODE[dtype], st)
    torch.cuda.synchronize()
    rW, rb = dy.double().t() @ x.double(), dy.double().sum(0)
    lib_err = rel_err(dy.t() @ x, rW)
    assert rel_err(dW, rW) <= max(1.05 * lib_err, 1e-6)
    assert rel_err(db, rb) <= (5e-3 if dtype == torch.bfloat16 else 1e-3)
    assert float(ws.abs().max()) == 0.0
    # without a bias gradient
    _lib.call('linear_wgrad', dy.data_ptr(), x.data_ptr(), dW.data_ptr(), None, ws.data_ptr(), N, O, I,
              _DTYPE_CODE[dtype], st)
    assert rel_err(dW, rW) <= max(1.05 * lib_err, 1e-6)


def test_linear_backward_through_wgrad(monkeypatch):
    """APOLLO_B200_WGRAD=1 routes the Linear layers' weight / bias gradients through linear_wgrad."""
    import apollo_vision_net_b200.rowops as ro
    g = torch.Generator().manual_seed(3)
    lin = ro.Linear(256, 512).to(DEV).to(torch.bfloat16)
    x = torch.randn(2, 9000, 256, generator=g).to(torch.bfloat16).to(DEV)
    go = torch.randn(2, 9000, 512, generator=g).to(torch.bfloat16).to(DEV)

    def run():
        lin.weight.grad = lin.bias.grad = None
        x1 = x.clone().requires_grad_(True)
        lin(x1).backward(go)
        return x1.grad, lin.weight.grad, lin.bias.grad

    monkeypatch.setenv('APOLLO_B200_WGRAD', '0')
    base = run()
    n0 = ro._lib.launch_count()
    monkeypatch.setenv('APOLLO_B200_WGRAD', '1')
    got = run()
    assert ro._lib.launch_count() - n0 == 2                 # linear_wgrad + its finalize, no column sum
    for a, b in zip(got, base):
        assert rel_err(a, b) <= 1e-2


def test_value_proj_bias_grad_comes_from_the_accumulator_pass(monkeypatch):
    """bf16 model: unscale_cast (the last pass over grad_value) also returns its column sums, which the
    value projection's backward picks up instead of running a column-sum kernel over 185 k rows.  Same
    bias gradient with and without the shortcut; one launch fewer with it; a stale offer is not used."""
    from apollo_vision_net_b200.modules import SpatialCrossAttention
    import apollo_vision_net_b200.rowops as ro
    import apollo_vision_net_b200.synthetic as syn
    C = 256
    levels = [(29, 50), (15, 25), (8, 13), (4, 7)]
    cfg = dict(embed_dims=C, pc_range=syn.PC_RANGE, batch_first=True,
               deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C, num_points=8,
                                         num_levels=len(levels)))
    m = SpatialCrossAttention(**cfg)
    _randomize(m, 5)
    m.to(DEV).to(torch.bfloat16).eval()
    q, feat, go, uv, mask, shapes, starts = _sca_inputs(1, 40, 40, levels, C, seed=23)
    bf = torch.bfloat16

    def run():
        m.zero_grad()
        f = feat.to(DEV).to(bf).requires_grad_(True)
        n0 = ro._lib.launch_count()
        out = m(q.to(DEV).to(bf), f, f, reference_points_cam=uv.to(DEV), bev_mask=mask.to(DEV),
                spatial_shapes=shapes.to(DEV), level_start_index=starts.to(DEV))
        out.backward(go.to(DEV).to(bf))
        torch.cuda.synchronize()
        return (m.deformable_attention.value_proj.bias.grad.float().clone(), f.grad.float().clone(),
                ro._lib.launch_count() - n0)

    b1, f1, n1 = run()
    monkeypatch.setattr(ro, 'offer_bias_grad', lambda grad, colsum: None)
    b0, f0, n0 = run()
    assert n0 == n1 + 1
    assert rel_err(b1, b0) <= 1e-2 and rel_err(f1, f0) <= 1e-2
    # an offer for a tensor that has since been modified (or freed) must be ignored
    monkeypatch.undo()
    t = torch.randn(64, C, device=DEV).to(bf)
    ro.offer_bias_grad(t, torch.zeros(C, device=DEV, dtype=bf))
    t.add_(1.0)                                                        # version counter moves
    assert ro._take_bias_grad(t, C) is None
    t2 = torch.randn(64, C, device=DEV).to(bf)
    ro.offer_bias_grad(t2, torch.ones(C, device=DEV, dtype=bf))
    assert ro._take_bias_grad(t2.view(8, 8, C), C) is not None        # a view of the same bytes is fine
    assert ro._take_bias_grad(t2, C) is None                           # ... and an offer is used once


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_ffn_relu_backward_with_fused_bias_grad(dtype, monkeypatch):
    """FFN (Linear - ReLU - Linear + identity): the ReLU's backward kernel also sums its rows and the first
    Linear takes that as its bias gradient.  Against the same FFN with torch's ReLU."""
    import apollo_vision_net_b200.rowops as ro
    from apollo_vision_net_b200.modules.encoder import FFN
    g = torch.Generator().manual_seed(8)
    ffn = FFN(256, 512).to(DEV).to(dtype)
    x = torch.randn(1, 5000, 256, generator=g).to(dtype).to(DEV)
    go = torch.randn(1, 5000, 256, generator=g).to(dtype).to(DEV)

    def run():
        ffn.zero_grad()
        x1 = x.clone().requires_grad_(True)
        n0 = ro._lib.launch_count()
        y = ffn(x1)
        y.backward(go)
        return [y.detach(), x1.grad] + [p.grad.clone() for p in ffn.parameters()], ro._lib.launch_count() - n0

    fused, n_fused = run()
    monkeypatch.setattr(ro.ReLU, 'forward', lambda self, t: torch.nn.functional.relu(t, inplace=True))
    plain, n_plain = run()
    assert n_fused == n_plain                       # relu_bwd_colsum replaces the column-sum launch
    assert torch.equal(fused[0], plain[0])
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for a, b in zip(fused[1:], plain[1:]):
        assert rel_err(a, b) <= tol


@pytest.mark.parametrize('C', [384, 768, 640, 96])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_layernorm_unsupported_widths_fall_back_to_torch(C, dtype):
    """Widths the row kernels do not instantiate (anything but 128/256/512/1024) must take torch's
    layer_norm -- LayerNorm, the fused Linear+residual+LayerNorm tail and the FFN tail alike --
    instead of failing in the kernel dispatch."""
    import torch.nn.functional as F
    from apollo_vision_net_b200 import rowops
    g = torch.Generator().manual_seed(C)
    x = torch.randn(70, C, generator=g).to(dtype).to(DEV).requires_grad_(True)
    res = torch.randn(70, C, generator=g).to(dtype).to(DEV)
    ln = rowops.LayerNorm(C).to(DEV).to(dtype)
    lin = rowops.Linear(C, C).to(DEV).to(dtype)
    n0 = rowops._lib.launch_count()
    y = ln(x)
    z = rowops.linear_add_layernorm(x, lin, res, ln)
    (y.float().sum() + z.float().sum()).backward()
    assert torch.equal(y, F.layer_norm(x, (C,), ln.weight, ln.bias, ln.eps))
    assert torch.equal(z, F.layer_norm(lin(x) + res, (C,), ln.weight, ln.bias, ln.eps))
    assert x.grad is not None and torch.isfinite(x.grad.float()).all()
    # no LayerNorm kernel of ours ran (the Linear's bias column sum may)
    assert rowops._lib.launch_count() - n0 <= 4
