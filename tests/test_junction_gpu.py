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
