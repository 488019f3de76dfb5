"""Tensor-core pass of the spatial cross-attention backward (csrc/coarse_scatter.cu): grad_value of the
coarse pyramid levels accumulated with tcgen05.mma from per-sample records instead of one L2 reduction
per corner (what it replaces: ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:106-146, mmcv's col2im).

Checked three ways: against the reduction path of the same kernel (offset / logit gradients must be
bit-identical -- the pass only moves grad_value), against an fp32 run of the kernel (fp32 reductions, no
tensor-core pass) as the truth for grad_value, and on small shapes against the CPU oracle
(oracle/msda_oracle.py, test infrastructure)."""
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _case(H, W, bs, levels, P, dtype, seed, scale=1.0, spread=4.0):
    import apollo_vision_net_b200.fused_ops as fo
    import apollo_vision_net_b200.synthetic as syn
    from apollo_vision_net_b200.modules.encoder import BEVFormerEncoder
    M, Dh, D, num_cam = 8, 32, 4, 6
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    L = len(shapes_l)
    shapes, starts = torch.tensor(shapes_l, device=DEV), torch.tensor(starts_l, device=DEV)
    l2i, img_shape = syn.camera_rig(scale, bs=bs)
    r3 = BEVFormerEncoder.get_reference_points(H, W, 8.0, D, dim='3d', bs=bs, device=DEV, dtype=torch.float32)
    geo = fo.bev_point_sampling(r3, syn.PC_RANGE, l2i, img_shape[0], img_shape[1], with_lists=True)
    g = torch.Generator(device='cpu').manual_seed(seed)
    value = torch.randn(bs * num_cam, Nk, M, Dh, generator=g).to(dtype).to(DEV)
    offsets = (torch.randn(bs, H * W, M, L, P, 2, generator=g) * spread).to(DEV)
    logits = torch.randn(bs, H * W, M, L * P, generator=g).to(DEV)
    go = torch.randn(bs, H * W, M * Dh, generator=g).to(dtype).to(DEV)
    return dict(value=value, shapes=shapes, starts=starts, offsets=offsets, logits=logits, geo=geo, go=go,
                num_cam=num_cam, W=W, starts_l=starts_l, Nk=Nk, L=L)


def _grads(c, value=None, go=None, lists=True):
    import apollo_vision_net_b200.fused_ops as fo
    geo = c['geo']
    v = (c['value'] if value is None else value).clone().requires_grad_(True)
    o = c['offsets'].clone().requires_grad_(True)
    lg = c['logits'].clone().requires_grad_(True)
    out = fo.SpatialCrossAttnFunction.apply(v, c['shapes'], c['starts'], o, lg, geo.reference_points_cam,
                                            geo.mask_u8, geo.hit_bits, c['num_cam'], c['W'],
                                            geo.lists() if lists else None)
    out.backward(c['go'] if go is None else go)
    return out.detach(), v.grad.float(), o.grad, lg.grad


CASES = [
    # H, W, bs, levels, P, dtype
    (200, 200, 1, 'base', 8, torch.bfloat16),          # BASELINE configs[1]: levels 2 + 3 = 1825 px on the tensor cores
    (50, 50, 2, 'base', 8, torch.float16),             # batch of 2 (batch-0 gating quirk), fp16 value
    (50, 50, 1, [(15, 25)], 8, torch.bfloat16),        # tiny config: the single level IS the patch
    (30, 30, 1, [(40, 64), (20, 32), (10, 16)], 4, torch.bfloat16),   # P = 4; 20x32 + 10x16 coarse
    (37, 23, 1, [(64, 96), (9, 13), (5, 7)], 8, torch.bfloat16),      # odd sizes: ragged K-steps
    (20, 20, 1, [(50, 60), (40, 50)], 8, torch.bfloat16),            # 5000 px do not fit: only the last level (2000)
    (20, 20, 1, [(50, 60), (45, 50)], 8, torch.bfloat16),            # no level fits: everything through reductions
]


@pytest.mark.parametrize('H,W,bs,levels,P,dtype', CASES)
def test_coarse_pass_against_reductions_and_fp32(H, W, bs, levels, P, dtype):
    import apollo_vision_net_b200.fused_ops as fo
    import apollo_vision_net_b200.synthetic as syn
    lv = syn.LEVELS_BASE if levels == 'base' else levels
    c = _case(H, W, bs, lv, P, dtype, seed=11)
    prev = fo.set_coarse_tensor_core_pass(True)
    try:
        out1, gv1, go1, gl1 = _grads(c)
        fo.set_coarse_tensor_core_pass(False)
        out0, gv0, go0, gl0 = _grads(c)
    finally:
        fo.set_coarse_tensor_core_pass(prev)
    # the main pass computes the same location / weight gradients either way
    assert torch.equal(out1, out0)
    assert torch.equal(go1, go0) and torch.equal(gl1, gl0)
    # grad_value: fp32 run of the kernel (fp32 value copy, fp32 reductions) is the truth
    _, truth, _, _ = _grads(c, value=c['value'].float(), go=c['go'].float())
    e1, e0 = rel_err(gv1, truth), rel_err(gv0, truth)
    assert e1 <= 8e-3, f'tensor-core pass: {e1:.3e} (reduction path {e0:.3e})'
    for l in range(c['L']):                              # per level, relative to the level's own scale
        s0 = c['starts_l'][l]
        s1 = c['starts_l'][l + 1] if l + 1 < c['L'] else c['Nk']
        el = rel_err(gv1[:, s0:s1], truth[:, s0:s1])
        assert el <= 8e-3, f'level {l}: {el:.3e}'
    assert not fo.grad_accumulator_overflowed(DEV)


def test_coarse_pass_with_fp32_accumulator_and_derived_lists():
    """fp32 accumulator mode (patch sums flushed with fp32 reductions) and hit lists derived from the bit
    field inside the Function (callers that only bring bev_mask)."""
    import apollo_vision_net_b200.fused_ops as fo
    import apollo_vision_net_b200.synthetic as syn
    c = _case(50, 50, 1, syn.LEVELS_BASE, 8, torch.bfloat16, seed=12)
    _, truth, _, _ = _grads(c, value=c['value'].float(), go=c['go'].float())
    with fo.grad_accumulator('fp32'):
        _, gv, _, _ = _grads(c, lists=False)
    assert rel_err(gv, truth) <= 6e-3
    _, gv16, _, _ = _grads(c, lists=False)
    assert rel_err(gv16, truth) <= 8e-3


def test_coarse_pass_camera_without_hits_and_repeated_pixels():
    """One camera sees nothing (empty hit list), and every sample of a row lands on the SAME coarse pixel
    (the record builder's read-modify-write of repeated entries)."""
    import apollo_vision_net_b200.fused_ops as fo
    import apollo_vision_net_b200.synthetic as syn
    c = _case(40, 40, 1, syn.LEVELS_BASE, 8, torch.bfloat16, seed=13, spread=0.0)   # zero offsets: repeated pixels
    geo = c['geo']
    bits = geo.hit_bits.clone()
    bits &= ~(1 << 2)                                      # camera 2 sees nothing
    geo2 = fo.BevGeometry(geo.reference_points_cam, geo.mask_u8, bits, None, None, geo.D)
    c['geo'] = geo2
    assert int(geo2.lists()[1][2]) == 0
    _, truth, _, _ = _grads(c, value=c['value'].float(), go=c['go'].float())
    _, gv, _, _ = _grads(c)
    assert rel_err(gv, truth) <= 8e-3
    cam2 = gv.view(1, 6, c['Nk'], 8, 32)[:, 2]
    assert float(cam2.abs().max()) == 0.0


def test_coarse_pass_matches_cpu_oracle_on_a_small_case():
    """Small enough for the CPU oracle: SCA through the reference's decomposition with
    multi_scale_deformable_attn_pytorch (oracle, fp32 autograd) vs the CUDA backward with the tensor-core
    pass, bf16 value.  Tolerance: bf16 rounding of value / gradient."""
    import apollo_vision_net_b200.synthetic as syn
    from oracle.msda_oracle import msda_torch
    c = _case(16, 16, 1, [(20, 32), (10, 16)], 8, torch.bfloat16, seed=14, spread=2.0)
    out, gv, go, gl = _grads(c)
    geo, shapes = c['geo'], c['shapes'].cpu()
    value = c['value'].float().cpu().requires_grad_(True)
    offsets = c['offsets'].cpu().requires_grad_(True)
    logits = c['logits'].cpu().requires_grad_(True)
    num_cam, HW, M, L, P, D = 6, 256, 8, c['L'], 8, 4
    attn = logits.softmax(-1).view(1, HW, M, L, P)
    norm = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
    off = (offsets / norm[None, None, None, :, None, :]).view(1, HW, M, L, P // D, D, 2)
    hit = geo.bev_mask.any(-1).cpu()
    count = hit.sum(0).clamp(min=1).float()
    total = torch.zeros(1, HW, M * 32)
    ref_cam = geo.reference_points_cam.cpu()
    for cam in range(num_cam):
        loc = (ref_cam[cam][:, :, None, None, None, :, :] + off).view(1, HW, M, L, P, 2)
        o = msda_torch(value[cam:cam + 1], shapes, loc, attn)
        total = total + o * hit[cam, 0][None, :, None]
    total = total / count[..., None]
    total.backward(c['go'].float().cpu())
    assert rel_err(out.float().cpu(), total.detach()) <= 1e-2
    assert rel_err(gv.cpu(), value.grad) <= 1e-2
    assert rel_err(gl.cpu(), logits.grad) <= 2e-2
    _ = syn
