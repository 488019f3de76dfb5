"""Helpers to replay the golden fixtures (tests/golden/*.npz) through the oracle or CUDA modules."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    with np.load(os.path.join(GOLDEN_DIR, name + '.npz')) as z:
        return {k: z[k] for k in z.files}


def T(a, device='cpu', grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    if grad:
        t.requires_grad_(True)
    return t


def params(g):
    return {k[len('param.'):]: torch.from_numpy(v) for k, v in g.items() if k.startswith('param.')}


def pgrads(g):
    return {k[len('pgrad.'):]: v for k, v in g.items() if k.startswith('pgrad.')}


def unpack_mask(g):
    shape = tuple(int(x) for x in g['mask_shape'])
    n = int(np.prod(shape))
    return np.unpackbits(g['mask_packed'])[:n].reshape(shape).astype(bool)


def build_sca(g, kind, device='cpu'):
    bs, H, W, C, heads, P, D = (int(x) for x in g['cfg'])
    L = len(g['shapes'])
    cfg = dict(embed_dims=C, num_cams=6, dropout=0.1, batch_first=True,
               deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                         num_heads=heads, num_points=P, num_levels=L,
                                         attn_logits_clamp=0.01))
    if kind == 'oracle':
        from oracle.modules_oracle import OracleSpatialCrossAttention as cls
    else:
        from apollo_vision_net_b200.modules import SpatialCrossAttention as cls
    m = cls(**cfg)
    m.load_state_dict(params(g))
    return m.to(device).eval()


def build_tsa(g, kind, device='cpu'):
    bs, H, W, C, heads, P = (int(x) for x in g['cfg'])
    cfg = dict(embed_dims=C, num_heads=heads, num_levels=1, num_points=P, attn_logits_clamp=0.6)
    if kind == 'oracle':
        from oracle.modules_oracle import OracleTemporalSelfAttention as cls
    else:
        from apollo_vision_net_b200.modules import TemporalSelfAttention as cls
    m = cls(**cfg)
    m.load_state_dict(params(g))
    return m.to(device).eval()


def build_decoder(g, kind, device='cpu'):
    bs, H, W, C, heads, P, Nq = (int(x) for x in g['cfg'])
    cfg = dict(embed_dims=C, num_heads=heads, num_levels=1, num_points=P, attn_logits_clamp=0.8)
    if kind == 'oracle':
        from oracle.modules_oracle import OracleCustomMSDeformableAttention as cls
    else:
        from apollo_vision_net_b200.modules import CustomMSDeformableAttention as cls
    m = cls(**cfg)
    m.load_state_dict(params(g))
    return m.to(device).eval()


def run_sca(m, g, device='cpu'):
    q = T(g['query'], device, True)
    feat = T(g['feat'], device, True)
    out = m(q, feat, feat, query_pos=T(g['query_pos'], device),
            reference_points_cam=T(g['ref_cam'], device), bev_mask=T(g['mask'], device),
            spatial_shapes=T(g['shapes'], device), level_start_index=T(g['starts'], device))
    out.backward(T(g['grad_out'], device))
    return out, q.grad, feat.grad


def run_tsa(m, g, device='cpu'):
    bs, H, W, C, heads, P = (int(x) for x in g['cfg'])
    q = T(g['query'], device, True)
    prev = T(g['prev'], device, True) if 'prev' in g else None
    out = m(q, prev, prev, query_pos=T(g['query_pos'], device), reference_points=T(g['ref'], device),
            spatial_shapes=torch.tensor([[H, W]], device=device),
            level_start_index=torch.tensor([0], device=device))
    out.backward(T(g['grad_out'], device))
    return out, q.grad, (prev.grad if prev is not None else None)


def run_decoder(m, g, device='cpu'):
    bs, H, W, C, heads, P, Nq = (int(x) for x in g['cfg'])
    q = T(g['query'], device, True)
    v = T(g['value'], device, True)
    r = T(g['ref'], device, True)
    out = m(q, None, v, query_pos=T(g['query_pos'], device), reference_points=r,
            spatial_shapes=torch.tensor([[H, W]], device=device),
            level_start_index=torch.tensor([0], device=device))
    out.backward(T(g['grad_out'], device))
    return out, q.grad, v.grad, r.grad


class StubDecLayer(torch.nn.Module):
    """Decoder layer of the ``det_decoder_small`` fixture: deformable cross-attention + LayerNorm
    (same parameter names as the stub the fixture was generated with, make_golden.py)."""

    def __init__(self, embed_dims, num_heads, num_points, kind='oracle'):
        super().__init__()
        self.embed_dims = embed_dims
        if kind == 'oracle':
            from oracle.modules_oracle import OracleCustomMSDeformableAttention as cls
        else:
            from apollo_vision_net_b200.modules import CustomMSDeformableAttention as cls
        self.attn = cls(embed_dims=embed_dims, num_heads=num_heads, num_levels=1, num_points=num_points)
        self.norm = torch.nn.LayerNorm(embed_dims)
        self.attentions = [self.attn]           # lets the decoder hoist the value projections

    def forward(self, query, key=None, value=None, query_pos=None, reference_points=None,
                spatial_shapes=None, level_start_index=None, key_padding_mask=None, **kw):
        q = self.attn(query, key, value, query_pos=query_pos, reference_points=reference_points,
                      spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                      key_padding_mask=key_padding_mask, **kw)
        return self.norm(q)


def build_det_decoder(g, kind, device='cpu'):
    """The package's DetectionTransformerDecoder around stub layers, with the fixture's parameters;
    returns (decoder, reg_branches)."""
    from apollo_vision_net_b200.modules import DetectionTransformerDecoder
    from apollo_vision_net_b200.registry import TRANSFORMER_LAYER
    bs, H, W, C, heads, P, Nq, NL = (int(x) for x in g['cfg'])
    if TRANSFORMER_LAYER.get('StubDecLayer') is None:
        TRANSFORMER_LAYER.register_module(name='StubDecLayer', module=StubDecLayer)
    dec = DetectionTransformerDecoder(
        transformerlayers=dict(type='StubDecLayer', embed_dims=C, num_heads=heads, num_points=P, kind=kind),
        num_layers=NL, return_intermediate=True)
    dec.load_state_dict(params(g))
    reg = torch.nn.ModuleList([torch.nn.Linear(C, 10) for _ in range(NL)])
    reg.load_state_dict({k[len('reg.'):]: torch.from_numpy(v) for k, v in g.items() if k.startswith('reg.')})
    return dec.to(device).eval(), reg.to(device)


def run_det_decoder(dec, reg, g, device='cpu'):
    bs, H, W = (int(x) for x in g['cfg'][:3])
    query, value = T(g['query'], device, True), T(g['value'], device, True)
    inter, refs = dec(query, key=None, value=value, query_pos=T(g['query_pos'], device),
                      reference_points=T(g['ref'], device), reg_branches=reg,
                      spatial_shapes=torch.tensor([[H, W]], device=device),
                      level_start_index=torch.tensor([0], device=device))
    inter.backward(T(g['grad_out'], device))
    return inter, refs, query.grad, value.grad


MAPTR_ORDER = ('self_attn', 'norm', 'self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')


def build_maptrv2_decoder(g, kind, device='cpu'):
    """The package's MapTRv2Decoder (decoupled layers) with the fixture's parameters; ``kind`` picks
    the deformable cross-attention: the CPU oracle module or the CUDA module."""
    import apollo_vision_net_b200 as pkg
    from apollo_vision_net_b200.registry import ATTENTION
    bs, H, W, C, heads, P, V, Pn, NL = (int(x) for x in g['cfg'])
    xattn = 'CustomMSDeformableAttention'
    if kind == 'oracle':
        from oracle.modules_oracle import OracleCustomMSDeformableAttention
        xattn = 'OracleCustomMSDeformableAttention'
        if ATTENTION.get(xattn) is None:
            ATTENTION.register_module(name=xattn, module=OracleCustomMSDeformableAttention)
    dec = pkg.build_transformer_layer_sequence(dict(
        type='MapTRv2Decoder', num_layers=NL, return_intermediate=True,
        transformerlayers=dict(
            type='MapTRv2DecoupledDetrTransformerDecoderLayer', num_vec=V, num_pts_per_vec=Pn,
            attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=heads, dropout=0.1),
                       dict(type='MultiheadAttention', embed_dims=C, num_heads=heads, dropout=0.1),
                       dict(type=xattn, embed_dims=C, num_heads=heads, num_levels=1, num_points=P)],
            feedforward_channels=2 * C, ffn_dropout=0.1, operation_order=MAPTR_ORDER)))
    dec.load_state_dict(params(g))
    reg = torch.nn.ModuleList([torch.nn.Linear(C, 2) for _ in range(NL)])
    reg.load_state_dict({k[len('reg.'):]: torch.from_numpy(v) for k, v in g.items() if k.startswith('reg.')})
    return dec.to(device).eval(), reg.to(device)


def run_maptrv2_decoder(dec, reg, g, device='cpu'):
    bs, H, W, C, heads, P, V, Pn, NL = (int(x) for x in g['cfg'])
    query, value = T(g['query'], device, True), T(g['value'], device, True)
    inter, refs = dec(query, key=None, value=value, query_pos=T(g['query_pos'], device),
                      reference_points=T(g['ref'], device), reg_branches=reg,
                      spatial_shapes=torch.tensor([[H, W]], device=device),
                      level_start_index=torch.tensor([0], device=device),
                      self_attn_mask=T(g['mask'], device), num_vec=V, num_pts_per_vec=Pn)
    inter.backward(T(g['grad_out'], device))
    return inter, refs, query.grad, value.grad
