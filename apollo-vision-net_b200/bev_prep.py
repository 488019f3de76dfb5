"""Host side of the pre-processing kernels in front of the BEV encoder (C ABI:
``bev_flatten_level``, ``bev_rotate_nearest``).

* :func:`can_bus_shift`      -- transformer.py:156-178: the BEV shift from the CAN-bus ego motion
  (host arithmetic on per-frame metadata, float64 like the reference's numpy code).
* :func:`rotate_prev_bev`    -- transformer.py:182-203: torchvision ``rotate`` (nearest) of every
  sample's previous BEV in ONE launch instead of a Python loop of per-sample rotations.
* :func:`flatten_features`   -- transformer.py:231-271: ``mlvl_feats`` to ``feat_flatten
  (num_cam, sum(hw), bs, C)`` with camera / level embeddings and non-finite values zeroed, one
  launch per level, no ``isfinite().all()`` host synchronisations.
"""
import math

import numpy as np
import torch
from torch.autograd.function import Function, once_differentiable

from . import _lib
from .multi_scale_deformable_attn_function import _DTYPE_CODE, _require_cuda, _stream_ptr


def can_bus_shift(can_bus, grid_length, bev_h, bev_w, use_shift=True):
    """``can_bus``: (bs, 18) rows (``img_metas[i]['can_bus']``).  Returns the (bs, 2) float64
    numpy array (shift_x, shift_y); the caller turns it into a tensor of the query dtype."""
    can_bus = np.asarray(can_bus, dtype=np.float64)
    delta_x, delta_y = can_bus[:, 0], can_bus[:, 1]
    ego_angle = can_bus[:, -2] / np.pi * 180
    grid_length_y, grid_length_x = grid_length[0], grid_length[1]
    translation_length = np.sqrt(delta_x ** 2 + delta_y ** 2)
    translation_angle = np.arctan2(delta_y, delta_x) / np.pi * 180
    bev_angle = ego_angle - translation_angle
    shift_y = translation_length * np.cos(bev_angle / 180 * np.pi) / grid_length_y / bev_h
    shift_x = translation_length * np.sin(bev_angle / 180 * np.pi) / grid_length_x / bev_w
    return np.stack([shift_x * use_shift, shift_y * use_shift], 1)


def _rotation_theta(angle_deg, center_xy, width, height):
    """torchvision ``rotate``: inverse affine matrix (python floats) of a rotation by ``angle``
    about ``center_xy`` (pixel coordinates), in image-centred coordinates."""
    cx = 1.0 * (center_xy[0] - width * 0.5)
    cy = 1.0 * (center_xy[1] - height * 0.5)
    rot = math.radians(-angle_deg)
    a, b, c, d = math.cos(rot), -math.sin(rot), math.sin(rot), math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m[2] += m[0] * (-cx) + m[1] * (-cy)
    m[5] += m[3] * (-cx) + m[4] * (-cy)
    m[2] += cx
    m[5] += cy
    return m


_grid_cache = {}


def _pixel_centres(n, device):
    key = (n, device)
    t = _grid_cache.get(key)
    if t is None:
        t = torch.linspace(-n * 0.5 + 0.5, n * 0.5 + 0.5 - 1, steps=n).to(device)   # host linspace, as torchvision's
        _grid_cache[key] = t
    return t


def rotate_prev_bev(prev_bev, angles_deg, bev_h, bev_w, center_xy):
    """prev_bev (bev_h*bev_w, bs, C) CUDA -> new tensor, sample i rotated by ``angles_deg[i]``
    degrees about ``center_xy`` (nearest neighbour, zeros outside).  History carries no gradient
    (the reference overwrites ``prev_bev`` in place)."""
    _require_cuda(prev_bev=prev_bev)
    HW, bs, C = prev_bev.shape
    assert HW == bev_h * bev_w, (prev_bev.shape, bev_h, bev_w)
    src = prev_bev.detach().contiguous()
    if src.dtype not in _DTYPE_CODE:
        raise RuntimeError(f'unsupported dtype {src.dtype}')
    # theta in fp32, then divided by the half sizes in fp32 -- the order torchvision uses
    theta = torch.tensor([_rotation_theta(float(a), center_xy, bev_w, bev_h) for a in angles_deg],
                         dtype=torch.float32).reshape(bs, 2, 3)
    theta = theta / torch.tensor([0.5 * bev_w, 0.5 * bev_h], dtype=torch.float32).view(1, 2, 1)
    theta = theta.reshape(bs, 6).to(src.device, non_blocking=True)
    xs, ys = _pixel_centres(bev_w, src.device), _pixel_centres(bev_h, src.device)
    out = torch.empty_like(src)
    with torch.cuda.device(src.device):
        _lib.call('bev_rotate_nearest', src.data_ptr(), out.data_ptr(), theta.data_ptr(), xs.data_ptr(),
                  ys.data_ptr(), bs, bev_h, bev_w, C, _DTYPE_CODE[src.dtype], _stream_ptr(src))
    return out


class _FlattenFunction(Function):
    @staticmethod
    def forward(ctx, cams_embeds, level_embeds, *feats):
        f0 = feats[0]
        bs, num_cam, C = f0.shape[:3]
        hws = [f.shape[3] * f.shape[4] for f in feats]
        Nk = sum(hws)
        out = torch.empty((num_cam, Nk, bs, C), dtype=f0.dtype, device=f0.device)
        ce = None if cams_embeds is None else cams_embeds.to(f0.dtype).contiguous()
        le = level_embeds.to(f0.dtype).contiguous()
        start = 0
        with torch.cuda.device(f0.device):
            for lvl, f in enumerate(feats):
                assert f.shape[:3] == (bs, num_cam, C) and f.dtype == f0.dtype
                fc = f.contiguous()
                _lib.call('bev_flatten_level', fc.data_ptr(), None if ce is None else ce.data_ptr(),
                          le[lvl].data_ptr(), out.data_ptr(), bs, num_cam, C, hws[lvl], Nk, start,
                          _DTYPE_CODE[f0.dtype], _stream_ptr(f0))
                start += hws[lvl]
        ctx.shapes = [tuple(f.shape) for f in feats]
        ctx.has_cams = cams_embeds is not None
        ctx.save_for_backward(*feats)
        ctx.embed_dtypes = (None if cams_embeds is None else cams_embeds.dtype, level_embeds.dtype)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        # not a hot path (once per frame, only when the image backbone trains): tensor ops
        feats = ctx.saved_tensors
        num_cam, Nk, bs, C = g.shape
        g_feats, g_lvl = [], []
        start = 0
        for f, shp in zip(feats, ctx.shapes):
            hw = shp[3] * shp[4]
            gl = g[:, start:start + hw]                                  # (num_cam, hw, bs, C)
            g_lvl.append(gl.float().sum(dim=(0, 1, 2)))
            gf = gl.permute(2, 0, 3, 1).reshape(shp)                     # (bs, num_cam, C, h, w)
            g_feats.append(torch.where(torch.isfinite(f), gf, torch.zeros_like(gf)))
            start += hw
        ce_dt, le_dt = ctx.embed_dtypes
        g_ce = g.float().sum(dim=(1, 2)).to(ce_dt) if ctx.has_cams else None
        g_le = torch.stack(g_lvl, 0).to(le_dt)
        return (g_ce, g_le, *g_feats)


def flatten_features(mlvl_feats, cams_embeds, level_embeds):
    """list of (bs, num_cam, C, h, w) CUDA tensors -> (feat_flatten (num_cam, sum(hw), bs, C),
    spatial_shapes (L, 2) int64, level_start_index (L,) int64, both on the device).  The level
    tables are built from Python ints (no device -> host traffic)."""
    _require_cuda(**{f'mlvl_feats[{i}]': f for i, f in enumerate(mlvl_feats)})
    if mlvl_feats[0].dtype not in _DTYPE_CODE:
        raise RuntimeError(f'unsupported feature dtype {mlvl_feats[0].dtype}')
    flat = _FlattenFunction.apply(cams_embeds, level_embeds, *mlvl_feats)
    shapes = [(int(f.shape[3]), int(f.shape[4])) for f in mlvl_feats]
    starts, s = [], 0
    for h, w in shapes:
        starts.append(s)
        s += h * w
    dev = mlvl_feats[0].device
    return (flat, torch.tensor(shapes, dtype=torch.long, device=dev),
            torch.tensor(starts, dtype=torch.long, device=dev))
