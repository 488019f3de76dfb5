"""Oracle for the multi-scale deformable attention operator (test infrastructure).

Two independent CPU restatements of the operator the reference reaches through
``MultiScaleDeformableAttnFunction_fp32.apply`` (GPU, mmcv ``_ext``) and
``multi_scale_deformable_attn_pytorch`` (CPU):

* :func:`msda_torch` -- the grid_sample formulation, following the reference's
  in-tree restatement ``multi_scale_deformable_attn_pytorch_2d``
  (``projects/mmdet3d_plugin/bevformer/modules/temporal_self_attention.py:293-348``).
  Differentiable, so ``torch.autograd`` on it is the backward oracle.
* :func:`msda_numpy` / :func:`msda_numpy_backward` -- explicit bilinear
  gather / scatter in float64 with the pixel convention of the mmcv kernel
  (``x_pix = loc_x * W - 0.5``; a corner contributes only when it lies inside
  the map, which is grid_sample's ``padding_mode='zeros', align_corners=False``).

The third-party code the reference calls (mmcv-full==1.4.0,
``mmcv/ops/multi_scale_deform_attn.py``) is absent from /root/reference; its
published algorithm is what both restatements state, and parity is anchored on
the reference's own call sites (``spatial_cross_attention.py:394-399``,
``temporal_self_attention.py:262-268``, ``decoder.py:345-350``).
"""
import numpy as np
import torch
import torch.nn.functional as F


def _shape_list(spatial_shapes):
    if isinstance(spatial_shapes, torch.Tensor):
        spatial_shapes = spatial_shapes.detach().cpu().tolist()
    return [(int(h), int(w)) for h, w in spatial_shapes]


def msda_torch(value, spatial_shapes, sampling_locations, attention_weights):
    """out[b,q,m*Dh+c] = sum_{l,p} A[b,q,m,l,p] * bilinear(V_l[b,:,m,c], loc[b,q,m,l,p]).

    value (B, Nk, M, Dh); spatial_shapes (L, 2) as (h, w);
    sampling_locations (B, Nq, M, L, P, 2) normalised (x, y);
    attention_weights (B, Nq, M, L, P)  ->  (B, Nq, M*Dh).
    Follows temporal_self_attention.py:293-348 (per-level grid_sample on 2*loc-1,
    bilinear, zero padding, align_corners=False, then a weighted sum over L*P).
    """
    hw = _shape_list(spatial_shapes)
    B, Nk, M, Dh = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    assert len(hw) == L and sum(h * w for h, w in hw) == Nk
    grids = sampling_locations * 2 - 1
    per_level = []
    start = 0
    for lvl, (h, w) in enumerate(hw):
        # (B, h*w, M, Dh) -> (B*M, Dh, h, w): one image per (batch, head)
        img = value[:, start:start + h * w].permute(0, 2, 3, 1).reshape(B * M, Dh, h, w)
        start += h * w
        # (B, Nq, M, P, 2) -> (B*M, Nq, P, 2)
        g = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(B * M, Nq, P, 2)
        per_level.append(F.grid_sample(img, g, mode='bilinear',
                                       padding_mode='zeros', align_corners=False))
    # (B*M, Dh, Nq, L, P) -> (B*M, Dh, Nq, L*P)
    sampled = torch.stack(per_level, dim=-2).flatten(-2)
    wts = attention_weights.permute(0, 2, 1, 3, 4).reshape(B * M, 1, Nq, L * P)
    out = (sampled * wts).sum(-1)                      # (B*M, Dh, Nq)
    return out.view(B, M * Dh, Nq).permute(0, 2, 1).contiguous()


def msda_torch_fwd_bwd(value, spatial_shapes, sampling_locations,
                       attention_weights, grad_output):
    """Forward + autograd backward of :func:`msda_torch`.

    Returns (out, grad_value, grad_sampling_locations, grad_attention_weights),
    the tuple the reference's ``backward`` fills
    (multi_scale_deformable_attn_function.py:128-161).
    """
    v = value.detach().clone().requires_grad_(True)
    s = sampling_locations.detach().clone().requires_grad_(True)
    a = attention_weights.detach().clone().requires_grad_(True)
    out = msda_torch(v, spatial_shapes, s, a)
    out.backward(grad_output)
    return out.detach(), v.grad, s.grad, a.grad


def _corner_terms(hw, loc):
    """Per-level corner indices / weights in float64.

    Returns for each level: (idx[4], wgt[4], valid[4], dwdx[4], dwdy[4]) with
    arrays shaped like loc[..., lvl, :, 0].
    """
    out = []
    for lvl, (h, w) in enumerate(hw):
        x = loc[..., lvl, :, 0] * w - 0.5
        y = loc[..., lvl, :, 1] * h - 0.5
        x0 = np.floor(x)
        y0 = np.floor(y)
        fx = x - x0
        fy = y - y0
        x0 = x0.astype(np.int64)
        y0 = y0.astype(np.int64)
        corners = []
        for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1)):
            xi = x0 + dx
            yi = y0 + dy
            wx = fx if dx else 1.0 - fx
            wy = fy if dy else 1.0 - fy
            sx = 1.0 if dx else -1.0
            sy = 1.0 if dy else -1.0
            ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
            idx = np.where(ok, yi * w + xi, 0)
            corners.append((idx, wx * wy, ok, sx * wy, sy * wx))
        out.append(corners)
    return out


def msda_numpy(value, spatial_shapes, sampling_locations, attention_weights):
    """Explicit float64 forward (independent of grid_sample)."""
    hw = _shape_list(spatial_shapes)
    v = np.asarray(value, dtype=np.float64)
    loc = np.asarray(sampling_locations, dtype=np.float64)
    att = np.asarray(attention_weights, dtype=np.float64)
    B, Nk, M, Dh = v.shape
    _, Nq, _, L, P, _ = loc.shape
    out = np.zeros((B, Nq, M, Dh), dtype=np.float64)
    bi = np.arange(B)[:, None, None, None]
    mi = np.arange(M)[None, None, :, None]
    start = 0
    for lvl, corners in enumerate(_corner_terms(hw, loc)):
        h, w = hw[lvl]
        vl = v[:, start:start + h * w]                  # (B, hw, M, Dh)
        start += h * w
        for idx, wgt, ok, _, _ in corners:
            g = vl[bi, idx, mi]                         # (B, Nq, M, P, Dh)
            coef = (wgt * ok * att[..., lvl, :])[..., None]
            out += (g * coef).sum(axis=3)
    return out.reshape(B, Nq, M * Dh)


def msda_numpy_backward(value, spatial_shapes, sampling_locations,
                        attention_weights, grad_output):
    """Explicit float64 backward: (grad_value, grad_loc, grad_attn).

    grad_loc carries the chain-rule factors W_l (x) and H_l (y) of the pixel
    mapping, as the op's backward does at the boundary
    (multi_scale_deformable_attn_function.py:144-158 receives them ready-made).
    """
    hw = _shape_list(spatial_shapes)
    v = np.asarray(value, dtype=np.float64)
    loc = np.asarray(sampling_locations, dtype=np.float64)
    att = np.asarray(attention_weights, dtype=np.float64)
    B, Nk, M, Dh = v.shape
    _, Nq, _, L, P, _ = loc.shape
    go = np.asarray(grad_output, dtype=np.float64).reshape(B, Nq, M, 1, Dh)
    gv = np.zeros_like(v)
    gl = np.zeros_like(loc)
    ga = np.zeros_like(att)
    bi = np.arange(B)[:, None, None, None]
    mi = np.arange(M)[None, None, :, None]
    bfull = np.broadcast_to(bi, (B, Nq, M, P))
    mfull = np.broadcast_to(mi, (B, Nq, M, P))
    start = 0
    for lvl, corners in enumerate(_corner_terms(hw, loc)):
        h, w = hw[lvl]
        vl = v[:, start:start + h * w]
        for idx, wgt, ok, dwdx, dwdy in corners:
            g = vl[bi, idx, mi] * ok[..., None]         # (B, Nq, M, P, Dh)
            dot = (g * go).sum(-1)                      # (B, Nq, M, P)
            a = att[..., lvl, :]
            ga[..., lvl, :] += wgt * dot
            gl[..., lvl, :, 0] += a * dwdx * dot * w
            gl[..., lvl, :, 1] += a * dwdy * dot * h
            contrib = (a * wgt * ok)[..., None] * go    # (B, Nq, M, P, Dh)
            np.add.at(gv, (bfull, start + idx, mfull), contrib)
        start += h * w
    return gv, gl, ga
