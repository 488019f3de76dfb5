"""Oracle for DCNv3 (test infrastructure, never imported by the product).

Two independent CPU restatements of the operator the reference reaches through its compiled ``DCNv3``
extension (``projects/mmdet3d_plugin/bevformer/backbones/ops_dcnv3/functions/dcnv3_func.py:19-62``):

* :func:`dcnv3_torch` -- the grid_sample formulation, following the reference's own pure-PyTorch version
  ``dcnv3_core_pytorch`` (dcnv3_func.py:98-188): pad the input, reference points + dilation grid + offsets,
  normalised by the padded size, bilinear ``grid_sample`` with zero padding, weighted sum over the kernel
  points.  Differentiable, so autograd on it is the backward oracle.
* :func:`dcnv3_numpy` -- the loops of the CUDA kernel (``src/cuda/dcnv3_im2col_cuda.cuh:216-275`` and the
  bilinear helper :29-75) in float64: sampling positions in pixels of the UNPADDED input, a sample takes
  part when -1 < h < H and -1 < w < W, corners outside the map contribute zero.

Pinning: ``tests/golden/dcnv3_small.npz`` holds inputs, outputs and gradients of the UNMODIFIED
``dcnv3_core_pytorch`` executed from /root/reference (``oracle/refshim.load_reference_dcnv3``;
``tests/golden/make_golden.py``); ``tests/test_dcnv3_cpu.py`` checks both restatements against it.
"""
import numpy as np
import torch
import torch.nn.functional as F


def output_size(size, kernel, stride, pad, dilation):
    return (size + 2 * pad - (dilation * (kernel - 1) + 1)) // stride + 1


def dcnv3_torch(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                dilation_w, group, group_channels, offset_scale):
    """input (N, H, W, group * group_channels); offset (N, Ho, Wo, group * K * 2); mask (N, Ho, Wo, group * K)."""
    x = F.pad(input, [0, 0, pad_h, pad_h, pad_w, pad_w])          # as the reference writes it (:125-127)
    N, Hp, Wp, _ = x.shape
    _, Ho, Wo, _ = offset.shape
    K = kernel_h * kernel_w
    dev, dt = input.device, input.dtype
    # reference points of the output pixels in the padded map (:98-117)
    cy, cx = (dilation_h * (kernel_h - 1)) // 2 + 0.5, (dilation_w * (kernel_w - 1)) // 2 + 0.5
    ry = (cy + torch.arange(Ho, device=dev, dtype=torch.float32) * stride_h) / Hp
    rx = (cx + torch.arange(Wo, device=dev, dtype=torch.float32) * stride_w) / Wp
    ref = torch.stack(torch.broadcast_tensors(rx[None, :], ry[:, None]), -1).view(1, Ho, Wo, 1, 2)
    # dilation grid, kernel_w outside kernel_h (:120-143)
    gx = (-((dilation_w * (kernel_w - 1)) // 2) + torch.arange(kernel_w, device=dev, dtype=torch.float32) * dilation_w)
    gy = (-((dilation_h * (kernel_h - 1)) // 2) + torch.arange(kernel_h, device=dev, dtype=torch.float32) * dilation_h)
    grid = torch.stack([(gx[:, None] / Wp).expand(kernel_w, kernel_h), (gy[None, :] / Hp).expand(kernel_w, kernel_h)], -1)
    grid = grid.reshape(1, 1, 1, K, 2).repeat(1, 1, 1, group, 1)             # (1, 1, 1, group * K, 2)
    norm = torch.tensor([Wp, Hp], device=dev, dtype=torch.float32).view(1, 1, 1, 1, 2)
    loc = (ref + grid * offset_scale).to(dt) + offset.view(N, Ho, Wo, group * K, 2) * offset_scale / norm.to(dt)
    grids = (2 * loc - 1).view(N, Ho * Wo, group, K, 2).transpose(1, 2).flatten(0, 1)
    x_ = x.reshape(N, Hp * Wp, group * group_channels).transpose(1, 2).reshape(N * group, group_channels, Hp, Wp)
    sampled = F.grid_sample(x_, grids, mode='bilinear', padding_mode='zeros', align_corners=False)
    m = mask.view(N, Ho * Wo, group, K).transpose(1, 2).reshape(N * group, 1, Ho * Wo, K)
    out = (sampled * m).sum(-1).view(N, group * group_channels, Ho * Wo)
    return out.transpose(1, 2).reshape(N, Ho, Wo, -1).contiguous()


def dcnv3_numpy(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                dilation_w, group, group_channels, offset_scale):
    """float64 restatement of the forward CUDA kernel; arrays in, array out."""
    x = np.asarray(input, dtype=np.float64)
    off = np.asarray(offset, dtype=np.float64)
    msk = np.asarray(mask, dtype=np.float64)
    N, H, W, _ = x.shape
    _, Ho, Wo, _ = off.shape
    K = kernel_h * kernel_w
    x = x.reshape(N, H, W, group, group_channels)
    off = off.reshape(N, Ho, Wo, group, K, 2)
    msk = msk.reshape(N, Ho, Wo, group, K)
    cw, ch = (dilation_w * (kernel_w - 1)) >> 1, (dilation_h * (kernel_h - 1)) >> 1
    k = np.arange(K)
    i, j = k // kernel_h, k % kernel_h
    p0w = (cw - pad_w + np.arange(Wo) * stride_w)[None, None, :, None, None]
    p0h = (ch - pad_h + np.arange(Ho) * stride_h)[None, :, None, None, None]
    lw = p0w - cw * offset_scale + (i * dilation_w + off[..., 0]) * offset_scale
    lh = p0h - ch * offset_scale + (j * dilation_h + off[..., 1]) * offset_scale
    take = (lh > -1) & (lw > -1) & (lh < H) & (lw < W)
    h0, w0 = np.floor(lh).astype(np.int64), np.floor(lw).astype(np.int64)
    fh, fw = lh - h0, lw - w0
    out = np.zeros((N, Ho, Wo, group, group_channels))
    n_idx = np.arange(N)[:, None, None, None, None]
    g_idx = np.arange(group)[None, None, None, :, None]
    for dh, dw, wgt in ((0, 0, (1 - fh) * (1 - fw)), (0, 1, (1 - fh) * fw), (1, 0, fh * (1 - fw)), (1, 1, fh * fw)):
        hh, ww = h0 + dh, w0 + dw
        ok = take & (hh >= 0) & (hh <= H - 1) & (ww >= 0) & (ww <= W - 1)
        v = x[n_idx, np.clip(hh, 0, H - 1), np.clip(ww, 0, W - 1), g_idx]            # (N, Ho, Wo, group, K, Cg)
        out += ((wgt * msk * ok)[..., None] * v).sum(-2)
    return out.reshape(N, Ho, Wo, group * group_channels)
