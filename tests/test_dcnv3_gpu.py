"""CUDA DCNv3 (csrc/dcnv3.cu on the MSDA operator kernels) against the oracle and the golden vectors of the
unmodified reference ``dcnv3_core_pytorch`` (ops_dcnv3/functions/dcnv3_func.py:119-188): forward and the
three gradients, fp32 (1e-5 / 1e-4, north_star's fp32 bar) and 16-bit inputs (compared with the oracle on the
same rounded inputs), padding / stride / dilation / rectangular kernels, the zero-offset initial state of
the reference module (positions exactly on pixel centres), non-vector channel counts."""
import os

import numpy as np
import pytest
import torch

from oracle.dcnv3_oracle import dcnv3_numpy, dcnv3_torch, output_size
from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'dcnv3_small.npz')


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_function_matches_reference_golden(tag):
    import apollo_vision_net_b200.dcnv3 as d
    z = np.load(GOLDEN)
    cfg = [int(v) for v in z[f'{tag}.cfg']]
    scale = float(z[f'{tag}.offset_scale'])
    t = {k: torch.from_numpy(z[f'{tag}.{k}']).to(DEV) for k in ('input', 'offset', 'mask', 'grad_output', 'output',
                                                                'grad_input', 'grad_offset', 'grad_mask')}
    x, off, msk = (t[k].clone().requires_grad_(True) for k in ('input', 'offset', 'mask'))
    out = d.DCNv3Function.apply(x, off, msk, *cfg, scale, 256)
    out.backward(t['grad_output'])
    assert rel_err(out, t['output']) <= 2e-5            # the golden carries the reference's normalise round trip
    assert rel_err(x.grad, t['grad_input']) <= 1e-4
    assert rel_err(off.grad, t['grad_offset']) <= 1e-4
    assert rel_err(msk.grad, t['grad_mask']) <= 1e-4


CASES = [
    # N, H, W, group, group_channels, kh, kw, stride, pad, dilation, offset_scale
    ('backbone_3x3', 2, 28, 36, 4, 16, 3, 3, 1, 1, 1, 1.0),
    ('strided_dilated', 1, 33, 21, 2, 32, 3, 3, 2, 2, 2, 2.0),
    ('rect_kernel', 2, 16, 19, 3, 8, 1, 5, 1, 0, 1, 0.5),
    ('odd_channels', 1, 12, 14, 2, 6, 3, 3, 1, 1, 1, 1.0),
]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16])
def test_function_matches_oracle(case, dtype):
    import apollo_vision_net_b200.dcnv3 as d
    _, N, H, W, G, Cg, kh, kw, st, pad, dil, scale = case
    if dtype != torch.float32 and Cg % 8 != 0:
        pytest.skip('16-bit inputs are covered with 16-byte channel lanes')
    Ho, Wo = output_size(H, kh, st, pad, dil), output_size(W, kw, st, pad, dil)
    K = kh * kw
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, H, W, G * Cg, generator=g).to(dtype)
    off = (torch.randn(N, Ho, Wo, G * K * 2, generator=g) * 2).to(dtype)
    msk = torch.softmax(torch.randn(N, Ho, Wo, G, K, generator=g), -1).reshape(N, Ho, Wo, G * K).to(dtype)
    go = torch.randn(N, Ho, Wo, G * Cg, generator=g).to(dtype)
    args = (kh, kw, st, st, pad, pad, dil, dil, G, Cg, scale)
    xd, od, md = (t.to(DEV).requires_grad_(True) for t in (x, off, msk))
    out = d.DCNv3Function.apply(xd, od, md, *args, 256)
    out.backward(go.to(DEV))
    xo, oo, mo = (t.double().requires_grad_(True) for t in (x, off, msk))
    ref = dcnv3_torch(xo, oo, mo, *args)
    ref.backward(go.double())
    ft, gt = {torch.float32: (1e-5, 1e-4), torch.bfloat16: (1e-2, 2e-2), torch.float16: (2e-3, 4e-3)}[dtype]
    assert rel_err(out, ref) <= ft
    assert rel_err(xd.grad, xo.grad) <= gt
    assert rel_err(md.grad, mo.grad) <= gt
    # The offset gradient is one-sided where a position falls exactly on a pixel row / column (bilinear
    # interpolation has a kink there); 16-bit offsets hit such positions all the time, and there the grid_sample
    # oracle's side depends on the rounding of its normalise round trip.  Compare everywhere else.
    o6 = off.double().view(N, Ho, Wo, G, K, 2)
    kk = torch.arange(K)
    cw, ch = (dil * (kw - 1)) // 2, (dil * (kh - 1)) // 2
    lw = (cw - pad + torch.arange(Wo) * st).view(1, 1, Wo, 1, 1) - cw * scale + ((kk // kh) * dil + o6[..., 0]) * scale
    lh = (ch - pad + torch.arange(Ho) * st).view(1, Ho, 1, 1, 1) - ch * scale + ((kk % kh) * dil + o6[..., 1]) * scale
    smooth = (((lw - lw.round()).abs() > 1e-6) & ((lh - lh.round()).abs() > 1e-6))[..., None].expand(-1, -1, -1, -1, -1, 2)
    assert smooth.float().mean() > (0.9 if dtype == torch.float32 else 0.5)
    got_off = od.grad.detach().cpu().double().view(N, Ho, Wo, G, K, 2)
    ref_off = oo.grad.view(N, Ho, Wo, G, K, 2)
    assert float(((got_off - ref_off).abs() * smooth).max() / ref_off.abs().max()) <= gt
    if dtype == torch.float32:
        assert rel_err(out, torch.from_numpy(dcnv3_numpy(x.numpy(), off.numpy(), msk.numpy(), *args))) <= 1e-5


def test_zero_offsets_sit_exactly_on_pixel_centres():
    """The reference module starts with zero offset weights (modules/dcnv3.py:296-299): every position is an
    integer pixel coordinate.  With a uniform mask the op is then a 3x3 box filter, and the one-sided
    derivative the reference kernel takes at an integer position (floor -> the pixel itself and its right /
    lower neighbour) must be reproduced, which needs the pixel-coordinate mode of the kernels."""
    import apollo_vision_net_b200.dcnv3 as d
    N, H, W, G, Cg = 1, 10, 12, 2, 8
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, H, W, G * Cg, generator=g)
    off = torch.zeros(N, H, W, G * 9 * 2)
    msk = torch.full((N, H, W, G * 9), 1.0 / 9)
    go = torch.randn(N, H, W, G * Cg, generator=g)
    args = (3, 3, 1, 1, 1, 1, 1, 1, G, Cg, 1.0)
    xd, od, md = (t.to(DEV).requires_grad_(True) for t in (x, off, msk))
    out = d.DCNv3Function.apply(xd, od, md, *args, 256)
    out.backward(go.to(DEV))
    box = torch.nn.functional.avg_pool2d(x.permute(0, 3, 1, 2), 3, 1, 1, count_include_pad=True).permute(0, 2, 3, 1)
    assert rel_err(out, box) <= 1e-6
    # gradient w.r.t. the offsets at integer positions: (v[x + 1] - v[x]) * mask * <grad_out, .>, right-sided
    xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (1, 2, 1, 2)).permute(0, 2, 3, 1).double()   # pad 1 left, 2 right
    exp = torch.zeros(N, H, W, G, 9, 2, dtype=torch.float64)
    gog = go.double().view(N, H, W, G, Cg)
    for i in range(3):          # kernel_w index
        for j in range(3):      # kernel_h index
            k = i * 3 + j
            v = xp[:, j:j + H, i:i + W].reshape(N, H, W, G, Cg)
            vx = xp[:, j:j + H, i + 1:i + 1 + W].reshape(N, H, W, G, Cg)
            vy = xp[:, j + 1:j + 1 + H, i:i + W].reshape(N, H, W, G, Cg)
            # a sample takes part only when -1 < position < size (dcnv3_im2col_cuda.cuh:262-263): the kernel points
            # that land on row / column -1 or H / W of the border pixels are skipped altogether
            lw = torch.arange(W).view(1, 1, W, 1) + i - 1
            lh = torch.arange(H).view(1, H, 1, 1) + j - 1
            take = ((lw >= 0) & (lw <= W - 1) & (lh >= 0) & (lh <= H - 1)).double()
            exp[..., k, 0] = ((vx - v) * gog).sum(-1) / 9 * take
            exp[..., k, 1] = ((vy - v) * gog).sum(-1) / 9 * take
    assert rel_err(od.grad.view(N, H, W, G, 9, 2), exp) <= 1e-5


def test_module_matches_reference_formulation():
    """DCNv3 module (reference constructor / parameter names) with the CUDA core against the same module with
    the oracle core, fp32."""
    import apollo_vision_net_b200.dcnv3 as d
    torch.manual_seed(3)
    m = d.DCNv3(channels=64, kernel_size=3, group=4, offset_scale=1.0, center_feature_scale=True)
    with torch.no_grad():
        m.offset.weight.normal_(0, 0.05)
        m.mask.weight.normal_(0, 0.05)
        m.center_feature_scale_proj_weight.normal_(0, 0.1)
    x = torch.randn(2, 14, 18, 64)
    assert sorted(n for n, _ in m.named_parameters()) == sorted(
        ['dw_conv.0.weight', 'dw_conv.0.bias', 'dw_conv.1.1.weight', 'dw_conv.1.1.bias', 'offset.weight', 'offset.bias',
         'mask.weight', 'mask.bias', 'input_proj.weight', 'input_proj.bias', 'output_proj.weight', 'output_proj.bias',
         'center_feature_scale_proj_weight', 'center_feature_scale_proj_bias'])
    # oracle arm: the same forward with dcnv3_torch as the core
    xi = m.input_proj(x)
    x1 = m.dw_conv(x.permute(0, 3, 1, 2))
    off = m.offset(x1)
    msk = torch.softmax(m.mask(x1).reshape(2, 14, 18, 4, -1), -1).reshape(2, 14, 18, -1)
    core = dcnv3_torch(xi, off, msk, 3, 3, 1, 1, 1, 1, 1, 1, 4, 16, 1.0)
    sc = torch.nn.functional.linear(x1, m.center_feature_scale_proj_weight, m.center_feature_scale_proj_bias).sigmoid()
    sc = sc[..., None].repeat(1, 1, 1, 1, 16).flatten(-2)
    ref = m.output_proj(core * (1 - sc) + xi * sc)
    got = m.to(DEV)(x.to(DEV))
    assert rel_err(got, ref) <= 2e-5


def test_errors_are_loud():
    import apollo_vision_net_b200.dcnv3 as d
    x = torch.randn(1, 8, 8, 32, device=DEV)
    off = torch.zeros(1, 8, 8, 4 * 9 * 2, device=DEV)
    msk = torch.zeros(1, 8, 8, 4 * 9, device=DEV)
    with pytest.raises(RuntimeError, match='does not follow from the geometry'):
        d.dcnv3_forward(x, off, msk, 3, 3, 1, 1, 0, 0, 1, 1, 4, 8, 1.0)          # pad 0 -> 6 x 6 output
    with pytest.raises(ValueError, match='wont match'):
        d.dcnv3_forward(x, off, msk, 3, 3, 1, 1, 1, 1, 1, 1, 4, 16, 1.0)
    with pytest.raises(RuntimeError, match='CPU'):
        d.dcnv3_forward(x.cpu(), off.cpu(), msk.cpu(), 3, 3, 1, 1, 1, 1, 1, 1, 4, 8, 1.0)
