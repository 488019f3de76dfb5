"""GPU parity of the operator boundary (C ABI msda_fwd / msda_bwd through the reference-named
autograd Functions) against the CPU oracle.  Tolerances are the ones north_star states:
forward 1e-5 relative in fp32 (1e-2 in bf16/fp16), gradients 1e-4 (atomic reordering)."""
import pytest
import torch

from oracle.msda_oracle import msda_torch, msda_torch_fwd_bwd
from tests.util import make_op_inputs, rel_err

pytestmark = pytest.mark.gpu

FWD_TOL = {torch.float32: 1e-5, torch.bfloat16: 1e-2, torch.float16: 2e-3}
BWD_TOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2, torch.float16: 2e-3}

CASES = [
    # B, levels, M, Dh, Nq, P
    (2, [(7, 9), (4, 5)], 8, 32, 50, 4),                       # small 2-level
    (6, [(28, 48)], 8, 32, 606, 8),                            # tiny SCA (config 1)
    (2, [(50, 50)], 8, 32, 2500, 4),                           # tiny TSA (config 1)
    (1, [(50, 50)], 8, 32, 7000, 4),                           # MapTRv2 decoder (config 4)
    (2, [(12, 20), (6, 10), (3, 5), (2, 3)], 8, 32, 333, 8),   # 4-level, ragged Nq
    (1, [(5, 7), (3, 3), (2, 2)], 3, 16, 17, 3),               # odd L, P, M
    (2, [(9, 11)], 8, 8, 100, 4),                              # Dh = 8 (hybrid configs)
    (1, [(9, 11)], 4, 64, 40, 2),
    (1, [(6, 5)], 2, 1, 9, 2),                                 # scalar fallback kernels
    (1, [(6, 5)], 2, 30, 9, 2),
    (1, [(6, 5)], 1, 71, 9, 3),
    (1, [(1, 1)], 1, 32, 5, 1),                                # 1x1 map
]


def _fn(dtype):
    import apollo_vision_net_b200 as pkg
    return (pkg.MultiScaleDeformableAttnFunction_fp32 if dtype == torch.float32
            else pkg.MultiScaleDeformableAttnFunction_fp16)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize('case', CASES)
def test_forward_backward_parity(case, dtype):
    B, levels, M, Dh, Nq, P = case
    value, shapes, starts, loc, att = make_op_inputs(B, levels, M, Dh, Nq, P, seed=1, dtype=dtype)
    g = torch.Generator().manual_seed(7)
    grad_out = torch.randn(B, Nq, M * Dh, generator=g).to(dtype)
    # oracle on the same (rounded) values: fp32 = the reference path, fp64 = the truth for grads
    v32 = value.float()
    ref_out = msda_torch(v32, shapes, loc, att)
    _, gv, gl, ga = msda_torch_fwd_bwd(v32.double(), shapes, loc.double(), att.double(),
                                       grad_out.double())
    dev = torch.device('cuda:0')
    v = value.to(dev).requires_grad_(True)
    lo = loc.to(dev).requires_grad_(True)
    at = att.to(dev).requires_grad_(True)
    out = _fn(dtype).apply(v, shapes.to(dev), starts.to(dev), lo, at, 64)
    assert out.shape == (B, Nq, M * Dh) and out.dtype == dtype
    assert rel_err(out, ref_out) <= FWD_TOL[dtype]
    out.backward(grad_out.to(dev))
    assert v.grad.dtype == dtype and lo.grad.dtype == torch.float32
    assert rel_err(v.grad, gv) <= BWD_TOL[dtype]
    assert rel_err(lo.grad, gl) <= BWD_TOL[dtype]
    assert rel_err(at.grad, ga) <= BWD_TOL[dtype]


def test_half_coordinates():
    """locations / weights in the value dtype (what custom_fwd(cast_inputs=fp16) produces)."""
    import apollo_vision_net_b200 as pkg
    value, shapes, starts, loc, att = make_op_inputs(2, [(10, 12)], 8, 32, 64, 4, seed=3,
                                                     dtype=torch.float16)
    loc16, att16 = loc.half(), att.half()
    ref = msda_torch(value.float(), shapes, loc16.float(), att16.float())
    dev = 'cuda:0'
    out = pkg.ms_deform_attn_forward(value.to(dev), shapes.to(dev), starts.to(dev),
                                     loc16.to(dev), att16.to(dev), im2col_step=64)
    assert rel_err(out, ref) <= 2e-3
    with torch.autocast('cuda', dtype=torch.float16):
        out2 = pkg.MultiScaleDeformableAttnFunction_fp16.apply(
            value.float().to(dev), shapes.to(dev), starts.to(dev), loc.to(dev), att.to(dev), 64)
    assert out2.dtype == torch.float16
    assert rel_err(out2, ref) <= 4e-3


def test_out_of_range_and_border_locations():
    """all-outside samples give exactly zero; samples on the -1 / W borders follow the
    (-1, W) open interval rule of the op."""
    import apollo_vision_net_b200 as pkg
    levels = [(4, 6)]
    value, shapes, starts, loc, att = make_op_inputs(1, levels, 2, 32, 8, 4, seed=5)
    loc[:, :4] = 3.0                                   # far outside
    H, W = levels[0]
    # x_pix = loc*W - 0.5: loc = -0.5/W -> x_pix = -1 (excluded), loc = (W+0.5)/W -> x_pix = W (excluded)
    loc[:, 4, :, :, 0, 0] = -0.5 / W
    loc[:, 5, :, :, 0, 0] = (W + 0.5) / W
    loc[:, 6, :, :, 0, 1] = -0.5 / H
    loc[:, 7, :, :, 0, 1] = (H + 0.5) / H
    ref = msda_torch(value, shapes, loc, att)
    dev = 'cuda:0'
    out = pkg.ms_deform_attn_forward(value.to(dev), shapes.to(dev), starts.to(dev), loc.to(dev),
                                     att.to(dev), im2col_step=64).cpu()
    assert torch.all(out[:, :4] == 0)
    assert rel_err(out, ref) <= 1e-5


def test_ext_module_contract():
    """ext_module.ms_deform_attn_backward fills caller-allocated zero buffers, returns None."""
    import apollo_vision_net_b200 as pkg
    value, shapes, starts, loc, att = make_op_inputs(2, [(6, 7), (3, 4)], 4, 32, 21, 2, seed=9)
    dev = 'cuda:0'
    a = [t.to(dev) for t in (value, shapes, starts, loc, att)]
    out = pkg.ext_module.ms_deform_attn_forward(*a, im2col_step=64)
    go = torch.randn_like(out)
    gv, gl, ga = torch.zeros_like(a[0]), torch.zeros_like(a[3]), torch.zeros_like(a[4])
    r = pkg.ext_module.ms_deform_attn_backward(*a, go.contiguous(), gv, gl, ga, im2col_step=64)
    assert r is None
    _, rv, rl, ra = msda_torch_fwd_bwd(value.double(), shapes, loc.double(), att.double(),
                                       go.cpu().double())
    assert rel_err(gv, rv) <= 1e-4 and rel_err(gl, rl) <= 1e-4 and rel_err(ga, ra) <= 1e-4


def test_empty_and_errors():
    import apollo_vision_net_b200 as pkg
    value, shapes, starts, loc, att = make_op_inputs(2, [(4, 4)], 8, 32, 0, 4)
    dev = 'cuda:0'
    out = pkg.ms_deform_attn_forward(value.to(dev), shapes.to(dev), starts.to(dev), loc.to(dev),
                                     att.to(dev))
    assert out.shape == (2, 0, 256)
    value, shapes, starts, loc, att = make_op_inputs(3, [(4, 4)], 8, 32, 5, 4)
    with pytest.raises(RuntimeError, match='im2col_step'):     # 3 % min(3, 2) != 0, as mmcv asserts
        pkg.ms_deform_attn_forward(value.to(dev), shapes.to(dev), starts.to(dev), loc.to(dev),
                                   att.to(dev), im2col_step=2)
    with pytest.raises(RuntimeError, match='CUDA tensor'):     # no CPU fallback
        pkg.ms_deform_attn_forward(value, shapes, starts, loc, att)


def test_linearity_at_base_size():
    """Size-independent property at BASELINE config 2 (base SCA op shape): the op is linear in
    `value`, and in `attention_weights`; out(v1 + v2) == out(v1) + out(v2)."""
    import apollo_vision_net_b200 as pkg
    levels = [(116, 200), (58, 100), (29, 50), (15, 25)]
    value, shapes, starts, loc, att = make_op_inputs(6, levels, 8, 32, 9690, 8, seed=11)
    dev = 'cuda:0'
    s, st, lo, at = shapes.to(dev), starts.to(dev), loc.to(dev), att.to(dev)
    v1 = value.to(dev)
    v2 = torch.randn_like(v1)
    f = pkg.ms_deform_attn_forward
    o1, o2, o12 = f(v1, s, st, lo, at), f(v2, s, st, lo, at), f(v1 + v2, s, st, lo, at)
    assert rel_err(o12, o1 + o2) <= 1e-5
    o_half = f(v1, s, st, lo, at * 0.5)
    assert rel_err(o_half * 2, o1) <= 1e-6
    # sub-sampled check against the oracle at full size
    idx = torch.arange(0, 9690, 97)
    ref = msda_torch(value, shapes, loc[:, idx], att[:, idx])
    assert rel_err(o1[:, idx.to(dev)], ref) <= 1e-5


def test_host_buffer_entry_points_match_oracle():
    """`msda_fwd_host` / `msda_fwd_bwd_host` (include/msda_b200.h): the op driven from pinned HOST
    buffers through the C ABI -- H2D, kernels, D2H and the stream synchronisation inside the call --
    against the CPU oracle (fp32: forward 1e-5, gradients 1e-4)."""
    import ctypes
    from apollo_vision_net_b200 import _lib
    from oracle.msda_oracle import msda_torch_fwd_bwd
    DEV = torch.device('cuda:0')
    B, M, Dh, Nq, P = 2, 8, 32, 300, 4
    levels = [(20, 30), (10, 15)]
    value, shapes, starts, loc, att = make_op_inputs(B, levels, M, Dh, Nq, P, seed=11)
    L, Nk = len(levels), value.shape[1]
    g = torch.Generator().manual_seed(12)
    go = torch.randn(B, Nq, M * Dh, generator=g)
    ref_out, ref_gv, ref_gl, ref_ga = msda_torch_fwd_bwd(value, shapes, loc, att, go)

    pin = lambda t: t.contiguous().pin_memory()
    value_h, shapes_h, starts_h, loc_h, att_h, go_h = (pin(t) for t in (value, shapes, starts, loc, att, go))
    out_h = torch.empty(B, Nq, M * Dh).pin_memory()
    gv_h, gl_h, ga_h = (torch.empty(t.shape).pin_memory() for t in (value, loc, att))
    need = int(_lib.lib().msda_host_scratch_bytes(B, Nk, M, Dh, L, Nq, P, _lib.F32, _lib.F32, 1))
    assert need > 0
    scratch = torch.empty(need, dtype=torch.uint8, device=DEV)
    stream = ctypes.c_void_p(torch.cuda.current_stream(DEV).cuda_stream)
    n0 = _lib.launch_count()
    _lib.call('msda_fwd_host', value_h.data_ptr(), shapes_h.data_ptr(), starts_h.data_ptr(), loc_h.data_ptr(),
              att_h.data_ptr(), out_h.data_ptr(), B, Nk, M, Dh, L, Nq, P, _lib.F32, _lib.F32,
              scratch.data_ptr(), need, stream)
    assert rel_err(out_h, ref_out) <= 1e-5            # the call returns after its own stream sync
    out_h.zero_()
    _lib.call('msda_fwd_bwd_host', value_h.data_ptr(), shapes_h.data_ptr(), starts_h.data_ptr(),
              loc_h.data_ptr(), att_h.data_ptr(), go_h.data_ptr(), out_h.data_ptr(), gv_h.data_ptr(),
              gl_h.data_ptr(), ga_h.data_ptr(), B, Nk, M, Dh, L, Nq, P, _lib.F32, _lib.F32,
              scratch.data_ptr(), need, stream)
    assert _lib.launch_count() - n0 >= 3
    assert rel_err(out_h, ref_out) <= 1e-5
    assert rel_err(gv_h, ref_gv) <= 1e-4
    assert rel_err(gl_h, ref_gl) <= 1e-4
    assert rel_err(ga_h, ref_ga) <= 1e-4
    # a scratch buffer that is too small is refused, not overrun
    need_fwd = int(_lib.lib().msda_host_scratch_bytes(B, Nk, M, Dh, L, Nq, P, _lib.F32, _lib.F32, 0))
    with pytest.raises(RuntimeError, match='scratch too small'):
        _lib.call('msda_fwd_host', value_h.data_ptr(), shapes_h.data_ptr(), starts_h.data_ptr(),
                  loc_h.data_ptr(), att_h.data_ptr(), out_h.data_ptr(), B, Nk, M, Dh, L, Nq, P, _lib.F32,
                  _lib.F32, scratch.data_ptr(), need_fwd - 1, stream)


def test_bf16_autocast_keeps_bf16():
    """Under torch.autocast(dtype=bfloat16) the module-level op-boundary path must stay in bf16: a
    bf16 value beyond the fp16 range (65504) must not be routed through an fp16 cast."""
    import apollo_vision_net_b200 as pkg
    from apollo_vision_net_b200.modules.deform_common import msda_apply
    value, shapes, starts, loc, att = make_op_inputs(1, [(8, 9)], 8, 32, 40, 4, seed=21)
    value = (value * 1.0e5).bfloat16()                  # |v| up to ~4e5 > 65504
    ref = msda_torch(value.float(), shapes, loc, att)
    dev = 'cuda:0'
    with torch.autocast('cuda', dtype=torch.bfloat16):
        out = msda_apply(value.to(dev), shapes.to(dev), starts.to(dev), loc.to(dev), att.to(dev), 64)
    assert out.dtype == torch.bfloat16
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) <= 2e-2
    with torch.autocast('cuda', dtype=torch.bfloat16):
        out2 = pkg.MultiScaleDeformableAttnFunction_bf16.apply(
            value.float().to(dev), shapes.to(dev), starts.to(dev), loc.to(dev), att.to(dev), 64)
    assert out2.dtype == torch.bfloat16 and rel_err(out2, ref) <= 2e-2


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_non_finite_value_only_reaches_samples_that_touch_it(dtype):
    """A non-finite value pixel must only affect samples with a VALID corner on it: samples outside
    the map, or whose clamped (zero-weight) corner address falls on that pixel, stay finite -- the
    reference kernel skips such corners instead of multiplying them by zero."""
    import apollo_vision_net_b200 as pkg
    H, W = 6, 8
    value, shapes, starts, loc, att = make_op_inputs(1, [(H, W)], 2, 32, 6, 2, seed=8, dtype=dtype)
    value[0, 0] = float('inf')                         # pixel (0, 0): a clamp target of the border
    value[0, W - 1] = float('nan')                     # pixel (0, W-1)
    loc[:, 0] = -2.0                                   # query 0: every sample far outside
    loc[:, 1, :, :, :, 0] = (W - 0.25) / W             # query 1: x_pix = W - 0.75 -> x1 = W is outside,
    loc[:, 1, :, :, :, 1] = 3.5 / H                    #   rows 3 / 3 (y_pix = 3.0): never touches row 0
    loc[:, 2, :, :, :, 0] = 4.5 / W                    # query 2: y_pix = -0.75 -> row -1 outside, row 0
    loc[:, 2, :, :, :, 1] = -0.25 / H                  #   at x = 4, 5: finite pixels only
    loc[:, 3:] = loc[:, 3:].clamp(0.3, 0.7)            # interior, rows >= 1
    dev = 'cuda:0'
    out = pkg.ms_deform_attn_forward(value.to(dev), shapes.to(dev), starts.to(dev), loc.to(dev),
                                     att.to(dev), im2col_step=64).cpu()
    assert torch.all(out[:, 0] == 0)
    assert torch.isfinite(out.float()).all(), 'a zero-weight corner leaked a non-finite value'
    v_ok = value.clone().float()
    v_ok[0, 0] = 0
    v_ok[0, W - 1] = 0
    ref = msda_torch(v_ok, shapes, loc, att)
    assert rel_err(out, ref) <= (1e-5 if dtype == torch.float32 else 1e-2)
