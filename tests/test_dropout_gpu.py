"""Dropout of the reference's training configuration (dropout 0.1 in TemporalSelfAttention /
SpatialCrossAttention, ffn_dropout 0.1: temporal_self_attention.py:54-66, 285-289;
spatial_cross_attention.py:171-173; bev_base_occ.py:127) fused into the row kernels with counter-based
masks that the backward recomputes.  Random masks cannot equal torch's, so parity is checked against a
torch composition that uses the SAME mask (read back through ``dropout_keep_mask``), plus the statistics
of the mask and its renewal per step."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_linear_dropout_add_layernorm_matches_torch_with_the_same_mask(dtype):
    import apollo_vision_net_b200.rowops as ro
    torch.manual_seed(0)
    rows, I, C, p = 4096, 256, 256, 0.1
    lin = ro.Linear(I, C).to(DEV, dtype)
    norm = ro.LayerNorm(C).to(DEV, dtype)
    x = torch.randn(rows, I, device=DEV, dtype=dtype, requires_grad=True)
    res = torch.randn(rows, C, device=DEV, dtype=dtype, requires_grad=True)
    go = torch.randn(rows, C, device=DEV, dtype=dtype)
    ro.advance_dropout_step(DEV)
    key = ro.dropout_state(DEV).clone()
    y = ro.linear_add_layernorm(x, lin, res, norm, p=p)
    site = ro._drop_site[0]
    y.backward(go)
    got = [y.detach(), x.grad.clone(), res.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone(),
           norm.weight.grad.clone(), norm.bias.grad.clone()]
    mask = ro.dropout_keep_mask((rows, C), dtype, key, site, p, DEV)
    keep = float(mask.float().mean())
    assert abs(keep - (1 - p)) < 0.01, keep
    for t in (x, res, lin.weight, lin.bias, norm.weight, norm.bias):
        t.grad = None
    # torch composition with that mask (fp32 reference for the 16-bit case)
    xd, rd = x.detach().double().requires_grad_(True), res.detach().double().requires_grad_(True)
    w, b = lin.weight.detach().double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    g, be = norm.weight.detach().double().requires_grad_(True), norm.bias.detach().double().requires_grad_(True)
    out = F.linear(xd, w, b)
    if dtype != torch.float32:
        out = out.to(dtype).double()            # the GEMM output is rounded before the dropout
    s = out * mask / (1 - p) + rd
    ref = F.layer_norm(s, (C,), g, be, norm.eps)
    ref.backward(go.double())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    want = [ref.detach(), xd.grad, rd.grad, w.grad, b.grad, g.grad, be.grad]
    for a, r, name in zip(got, want, ['y', 'dx', 'dres', 'dW', 'db', 'dgamma', 'dbeta']):
        assert rel_err(a, r) <= tol, (name, rel_err(a, r))


def test_relu_dropout_forward_backward_and_mask_statistics():
    import apollo_vision_net_b200.rowops as ro
    torch.manual_seed(1)
    rows, C, p = 8192, 1024, 0.1
    act = ro.ReLU(inplace=True)
    lin = ro.Linear(256, C).to(DEV, torch.bfloat16)
    x = torch.randn(rows, 256, device=DEV, dtype=torch.bfloat16)
    ro.advance_dropout_step(DEV)
    key = ro.dropout_state(DEV).clone()
    h = lin(x)
    pre = h.detach().clone()
    y = act.forward_dropout(h, p)
    site = ro._drop_site[0]
    mask = ro.dropout_keep_mask((rows, C), torch.bfloat16, key, site, p, DEV)
    want = (torch.relu(pre.float()) * mask / (1 - p)).to(torch.bfloat16)
    assert torch.equal(y.detach(), want)
    go = torch.randn(rows, C, device=DEV, dtype=torch.bfloat16)
    y.backward(go)
    # d bias of the Linear = column sums of go * relu' * mask / (1 - p)
    dpre = (go.float() * (pre.float() > 0) * mask / (1 - p)).to(torch.bfloat16).float()
    assert rel_err(lin.bias.grad, dpre.sum(0)) <= 2e-2
    assert rel_err(lin.weight.grad, dpre.t() @ x.float()) <= 2e-2
    # statistics: keep rate per column and per row
    m = mask.float()
    assert abs(float(m.mean()) - 0.9) < 2e-3
    assert float((m.mean(0) - 0.9).abs().max()) < 0.02 and float((m.mean(1) - 0.9).abs().max()) < 0.06
    # neighbouring elements are uncorrelated
    c = torch.corrcoef(torch.stack([m[:, :-1].flatten(), m[:, 1:].flatten()]))[0, 1]
    assert abs(float(c)) < 5e-3


def test_masks_change_with_the_step_and_the_site_but_repeat_for_the_backward():
    import apollo_vision_net_b200.rowops as ro
    ro.advance_dropout_step(DEV)
    k1 = ro.dropout_state(DEV).clone()
    m1 = ro.dropout_keep_mask((64, 256), torch.bfloat16, k1, 5, 0.1, DEV)
    assert torch.equal(m1, ro.dropout_keep_mask((64, 256), torch.bfloat16, k1, 5, 0.1, DEV))
    assert not torch.equal(m1, ro.dropout_keep_mask((64, 256), torch.bfloat16, k1, 6, 0.1, DEV))
    ro.advance_dropout_step(DEV)
    k2 = ro.dropout_state(DEV).clone()
    assert int(k2[1]) == int(k1[1]) + 1
    assert not torch.equal(m1, ro.dropout_keep_mask((64, 256), torch.bfloat16, k2, 5, 0.1, DEV))
    # fp32 tensors (4 elements per thread) see the same element-wise mask as 16-bit ones
    assert torch.equal(m1, ro.dropout_keep_mask((64, 256), torch.float32, k1, 5, 0.1, DEV))


def _tiny_encoder(dropout, dtype):
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.synthetic as syn
    bs, H, W, C = 1, 40, 40, 256          # (>= 1024 rows: every dropout takes the fused kernels)
    shapes_l, starts_l, Nk = syn.level_tables(syn.LEVELS_TINY)
    l2i, img_shape = syn.camera_rig(0.5, bs=bs)
    torch.manual_seed(3)
    enc = pkg.build_transformer_layer_sequence(dict(
        type='BEVFormerEncoder', num_layers=2, pc_range=syn.PC_RANGE, num_points_in_pillar=4,
        return_intermediate=False,
        transformerlayers=dict(
            type='BEVFormerLayer',
            attn_cfgs=[dict(type='TemporalSelfAttention', embed_dims=C, num_levels=1, dropout=dropout),
                       dict(type='SpatialCrossAttention', pc_range=syn.PC_RANGE, embed_dims=C, dropout=dropout,
                            deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                                      num_points=8, num_levels=1))],
            feedforward_channels=512, ffn_dropout=dropout,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    enc.init_weights() if hasattr(enc, 'init_weights') else None
    enc.to(DEV, dtype)
    g = torch.Generator().manual_seed(21)
    mk = lambda *sh: torch.randn(*sh, generator=g).to(DEV, dtype)       # noqa: E731
    feat = mk(6, Nk, bs, C)
    inputs = dict(bev_query=mk(H * W, bs, C), key=feat, value=feat, bev_h=H, bev_w=W, bev_pos=mk(H * W, bs, C),
                  spatial_shapes=torch.tensor(shapes_l, device=DEV), level_start_index=torch.tensor(starts_l, device=DEV),
                  prev_bev=mk(H * W, bs, C), shift=torch.tensor([[0.01, -0.02]], device=DEV),
                  lidar2img=l2i, img_shape=img_shape)
    return enc, inputs


def test_encoder_layer_with_reference_dropout_trains_and_is_deterministic_under_a_seed():
    """BEVFormerLayer with dropout 0.1 everywhere (the reference's values): the fused path runs (post-norm
    tails + FFN), two runs from the same seed / step agree, a different step does not, and eval mode is
    deterministic."""
    import apollo_vision_net_b200.rowops as ro
    from apollo_vision_net_b200 import _lib
    enc, inputs = _tiny_encoder(0.1, torch.bfloat16)
    enc.train()

    def run(seed_step):
        torch.manual_seed(5)
        ro.reseed_dropout(DEV, 1234)
        ro.dropout_state(DEV)[1] = seed_step
        ro._drop_site[0] = 0
        for prm in enc.parameters():
            prm.grad = None
        out = enc(**inputs)
        out.float().square().mean().backward()
        return out.detach().clone(), [prm.grad.clone() for prm in enc.parameters() if prm.grad is not None]

    n0 = _lib.launch_count()
    o1, g1 = run(7)
    assert _lib.launch_count() - n0 > 20                                # our kernels ran (fused path)
    o2, g2 = run(7)
    assert torch.equal(o1, o2)
    # (grad_value of the attention kernels is summed with reductions in arbitrary order: close, not equal)
    for a, b in zip(g1, g2):
        assert rel_err(a, b) <= 2e-2
    o3, _ = run(8)
    assert not torch.equal(o1, o3)
    enc.eval()
    with torch.no_grad():
        e1 = enc(**inputs)
        e2 = enc(**inputs)
    assert torch.equal(e1, e2)
    assert float((e1.float() - o1.float()).abs().max()) > 0            # training really dropped units
