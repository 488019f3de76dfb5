"""Residual-block gradient junctions (rowops.Junction): the input of a post-norm block is both the residual of
its tail and the input of the branch's first Linear layer (custom_base_transformer_layer.py:142-161 around
temporal_self_attention.py:285-289 / spatial_cross_attention.py:171-173 / the FFN's identity add).  The
product adds the two gradients inside the first Linear's dX GEMM (beta = 1) instead of a separate kernel;
these tests pin that path to the plain autograd composition."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _block(ro, x, fc1, fc2, norm, junction):
    h = ro.linear(x, fc1.weight, fc1.bias, junction)
    h = torch.tanh(h)
    return ro.linear_add_layernorm(h, fc2, x, norm, p=0.0, junction=junction)


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-6), (torch.bfloat16, 1e-2)])
def test_junction_block_matches_plain_autograd(dtype, tol):
    import apollo_vision_net_b200.rowops as ro
    rows, C, Fd = 4096, 256, 512
    fc1, fc2 = ro.Linear(C, Fd).to(DEV, dtype), ro.Linear(Fd, C).to(DEV, dtype)
    norm = ro.LayerNorm(C).to(DEV, dtype)
    params = list(fc1.parameters()) + list(fc2.parameters()) + list(norm.parameters())
    x0 = torch.randn(rows, C, device=DEV, dtype=dtype)
    go = torch.randn(rows, C, device=DEV, dtype=dtype)
    grads = []
    for use in (False, True):
        x = x0.clone().requires_grad_(True)
        y = _block(ro, x * 1.0, fc1, fc2, norm, ro.Junction() if use else None)   # (x * 1: x is not a leaf)
        y.backward(go)
        grads.append([y.detach().clone(), x.grad.clone()] + [p.grad.clone() for p in params])
        for p in params:
            p.grad = None
    assert torch.equal(grads[0][0], grads[1][0])
    # truth for the input gradient: the fp64 composition
    xd = x0.double().requires_grad_(True)
    h = torch.tanh(F.linear(xd, fc1.weight.double(), fc1.bias.double()))
    yd = F.layer_norm(F.linear(h, fc2.weight.double(), fc2.bias.double()) + xd, (C,), norm.weight.double(),
                      norm.bias.double(), norm.eps)
    yd.backward(go.double())
    for a, b in zip(grads[0][1:], grads[1][1:]):
        assert rel_err(b, a) <= tol, rel_err(b, a)
    # the fused sum rounds once (fp32 accumulator of the GEMM), the separate add twice: it is not further
    # from the truth than the plain composition
    e_plain, e_junc = rel_err(grads[0][1], xd.grad), rel_err(grads[1][1], xd.grad)
    assert e_junc <= 1.05 * e_plain + 1e-7, (e_junc, e_plain)


def test_parked_gradient_that_nobody_collects_raises():
    import apollo_vision_net_b200.rowops as ro
    rows, C = 2048, 256
    fc1, fc2 = ro.Linear(C, C).to(DEV), ro.Linear(C, C).to(DEV)
    norm = ro.LayerNorm(C).to(DEV)
    x = torch.randn(rows, C, device=DEV, requires_grad=True)
    xx = x * 1.0
    tok = ro.Junction()
    h = ro.linear(xx, fc1.weight, fc1.bias, tok)              # registers as the junction's consumer ...
    y = ro.linear_add_layernorm(h.detach(), fc2, xx, norm, p=0.0, junction=tok)   # ... but is cut off
    with pytest.raises(RuntimeError, match='never collected'):
        y.sum().backward()


@pytest.mark.parametrize('pos_grad', [False, True])
@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-6), (torch.bfloat16, 1e-2)])
def test_paired_query_linear_matches_cat_composition(dtype, tol, pos_grad):
    import apollo_vision_net_b200.rowops as ro
    bs, n, C, O = 1, 4096, 256, 192
    lin = ro.Linear(2 * C, O).to(DEV, dtype)
    out_proj = ro.Linear(O, C).to(DEV, dtype)
    norm = ro.LayerNorm(C).to(DEV, dtype)
    paired0 = torch.randn(bs, n, C, device=DEV, dtype=dtype)
    q0 = torch.randn(bs, n, C, device=DEV, dtype=dtype)
    pos0 = torch.randn(bs, n, C, device=DEV, dtype=dtype)
    go = torch.randn(bs, n, C, device=DEV, dtype=dtype)
    params = list(lin.parameters()) + list(out_proj.parameters()) + list(norm.parameters())
    res = []
    for fused in (False, True):
        paired = paired0.clone().requires_grad_(True)
        q = q0.clone().requires_grad_(True)
        pos = pos0.clone().requires_grad_(pos_grad)
        qq = q * 1.0
        if fused:
            tok = ro.Junction()
            c = ro.paired_query_linear(paired, qq, pos, lin.weight, lin.bias, tok)
            y = ro.linear_add_layernorm(torch.sin(c), out_proj, qq, norm, p=0.0, junction=tok)
        else:
            c = ro.linear(torch.cat([paired, qq + pos], -1), lin.weight, lin.bias)
            y = ro.linear_add_layernorm(torch.sin(c), out_proj, qq, norm, p=0.0)
        y.backward(go)
        res.append([y.detach().clone(), paired.grad.clone(), q.grad.clone()] +
                   ([pos.grad.clone()] if pos_grad else []) + [p.grad.clone() for p in params])
        for p in params:
            p.grad = None
    assert torch.equal(res[0][0], res[1][0])
    for a, b in zip(res[0][1:], res[1][1:]):
        assert rel_err(b, a) <= tol, rel_err(b, a)


def test_unscale_cast_strided_equals_contiguous():
    """The column-block form of the accumulator conversion writes the same values (and returns the same column
    sums) as the contiguous form."""
    from apollo_vision_net_b200 import _lib
    from apollo_vision_net_b200.multi_scale_deformable_attn_function import _DTYPE_CODE
    import apollo_vision_net_b200.rowops as ro
    rows, C, blocks = 3000, 256, 3
    acc = (torch.randn(rows, C, device=DEV) * 8).to(torch.float16)
    scale = torch.tensor([4.0], device=DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    st = torch.cuda.current_stream(DEV).cuda_stream
    ws = ro._workspace(DEV, 8 * C)
    outs, sums = [], []
    wide = torch.full((rows, blocks * C), 7.0, device=DEV, dtype=torch.bfloat16)
    for ld, out in ((0, torch.empty(rows, C, device=DEV, dtype=torch.bfloat16)), (blocks * C, wide[:, C:2 * C])):
        s = torch.empty(C, device=DEV, dtype=torch.bfloat16)
        _lib.call('unscale_cast_strided', acc.data_ptr(), out.data_ptr(), scale.data_ptr(), acc.numel(),
                  _DTYPE_CODE[torch.bfloat16], None, 0, acc.numel(), 0, s.data_ptr(), ws.data_ptr(), C,
                  flag.data_ptr(), ld, st)
        outs.append(out.clone())
        sums.append(s)
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(outs[0], (acc.float() / 4.0).to(torch.bfloat16))
    assert rel_err(sums[1], sums[0]) <= 1e-2                                   # (fp32 atomics: order differs)
    assert torch.all(wide[:, :C] == 7.0) and torch.all(wide[:, 2 * C:] == 7.0)   # the neighbours are untouched
    assert int(flag.item()) == 0


def _small_encoder(num_layers, dtype):
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.synthetic as syn
    C = 256
    levels = [(15, 25), (8, 13)]
    enc = pkg.build_transformer_layer_sequence(dict(
        type='BEVFormerEncoder', num_layers=num_layers, pc_range=syn.PC_RANGE, num_points_in_pillar=4,
        transformerlayers=dict(
            type='BEVFormerLayer',
            attn_cfgs=[dict(type='TemporalSelfAttention', embed_dims=C, num_levels=1, dropout=0.0),
                       dict(type='SpatialCrossAttention', pc_range=syn.PC_RANGE, embed_dims=C, dropout=0.0,
                            deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                                      num_points=8, num_levels=len(levels)))],
            feedforward_channels=512, ffn_dropout=0.0,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    g = torch.Generator().manual_seed(3)
    for n, p in enc.named_parameters():
        if n.endswith('sampling_offsets.weight') or n.endswith('attention_weights.weight'):
            p.data = torch.randn(p.shape, generator=g) * 0.02
    return enc.to(DEV, dtype).train(), levels


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-5), (torch.bfloat16, 3e-2)])
def test_hoisted_value_projections_match_per_layer_projections(dtype, tol, monkeypatch):
    """Three-layer encoder, forward + backward: the hoisted SCA / TSA value projections with their gradients written
    side by side (one dX and one dW GEMM for all layers) against every layer projecting for itself."""
    import apollo_vision_net_b200.synthetic as syn
    from apollo_vision_net_b200.modules.encoder import BEVFormerEncoder
    enc, levels = _small_encoder(3, dtype)
    H, W, C = 20, 24, 256
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    l2i, img_shape = syn.camera_rig(0.5, bs=1)
    g = torch.Generator().manual_seed(11)
    bevq = torch.randn(H * W, 1, C, generator=g).to(DEV, dtype).requires_grad_(True)
    pos = torch.randn(H * W, 1, C, generator=g).to(DEV, dtype)
    prev = torch.randn(H * W, 1, C, generator=g).to(DEV, dtype)
    feat0 = torch.randn(6, Nk, 1, C, generator=g).to(DEV, dtype)
    go = torch.randn(1, H * W, C, generator=g).to(DEV, dtype)
    shapes, starts = torch.tensor(shapes_l, device=DEV), torch.tensor(starts_l, device=DEV)
    results = []
    for hoist in (True, False):
        if not hoist:
            monkeypatch.setattr(BEVFormerEncoder, '_hoisted_sca_values', lambda self, *a, **k: None)
            monkeypatch.setattr(BEVFormerEncoder, '_hoisted_tsa_values', lambda self, *a, **k: None)
        feat = feat0.clone().requires_grad_(True)
        bevq.grad = None
        for p in enc.parameters():
            p.grad = None
        out = enc(bevq, feat, feat, bev_h=H, bev_w=W, bev_pos=pos, spatial_shapes=shapes,
                  level_start_index=starts, prev_bev=prev, shift=torch.zeros(1, 2, device=DEV), lidar2img=l2i,
                  img_shape=img_shape)
        out.backward(go)
        named = {n: p.grad.clone() for n, p in enc.named_parameters() if 'value_proj' in n}
        results.append((out.detach().clone(), feat.grad.clone(), bevq.grad.clone(), named))
    (o1, f1, q1, n1), (o0, f0, q0, n0) = results
    assert torch.equal(o1, o0)                       # the forward is the same GEMMs
    assert rel_err(f1, f0) <= tol, rel_err(f1, f0)
    assert rel_err(q1, q0) <= tol, rel_err(q1, q0)
    assert set(n1) == set(n0) and len(n1) == 12      # 3 layers x (TSA, SCA) x (weight, bias)
    for k in n1:
        assert rel_err(n1[k], n0[k]) <= tol, (k, rel_err(n1[k], n0[k]))
