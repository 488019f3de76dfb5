"""CPU restatement of ``PerceptionTransformer.get_bev_features`` pre-processing
(reference: projects/mmdet3d_plugin/bevformer/modules/transformer.py:119-298).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Pinned by tests/golden/bev_features_small.npz,
which tests/golden/make_golden.py produced by running the UNMODIFIED reference class (with
torchvision's ``rotate``) through oracle/refshim.

Pieces (each is what the reference hands to ``self.encoder``):

* :func:`can_bus_shift`       -- transformer.py:156-178 (numpy float64 arithmetic, then fp32)
* :func:`rotate_prev_bev`     -- transformer.py:182-203; ``torchvision.transforms.functional.rotate``
  with its defaults (nearest, no expand, zero fill) restated from torchvision 0.26
  (``_get_inverse_affine_matrix`` + ``_gen_affine_grid`` + ``grid_sample(nearest, zeros,
  align_corners=False)``), which is a dependency absent from /root/reference
* :func:`flatten_features`    -- transformer.py:231-271 (non-finite values zeroed, camera and level
  embeddings, ``(num_cam, sum(HW), bs, C)`` layout, level tables)
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def can_bus_shift(can_bus, grid_length, bev_h, bev_w, use_shift=True, dtype=torch.float32):
    """(bs, 18) can_bus rows -> (bs, 2) shift = (shift_x, shift_y) in BEV-normalised units."""
    can_bus = np.asarray(can_bus, dtype=np.float64)
    delta_x, delta_y = can_bus[:, 0], can_bus[:, 1]
    ego_angle = can_bus[:, -2] / np.pi * 180
    grid_length_y, grid_length_x = grid_length[0], grid_length[1]
    translation_length = np.sqrt(delta_x ** 2 + delta_y ** 2)
    translation_angle = np.arctan2(delta_y, delta_x) / np.pi * 180
    bev_angle = ego_angle - translation_angle
    shift_y = translation_length * np.cos(bev_angle / 180 * np.pi) / grid_length_y / bev_h
    shift_x = translation_length * np.sin(bev_angle / 180 * np.pi) / grid_length_x / bev_w
    shift_y = shift_y * use_shift
    shift_x = shift_x * use_shift
    return torch.tensor(np.stack([shift_x, shift_y], 0), dtype=dtype).permute(1, 0)


def inverse_rotation_matrix(angle_deg, center_xy, width, height):
    """torchvision ``rotate``: centre relative to the image centre, inverse affine matrix of a
    rotation by ``-angle`` (python floats): [a, b, c, d, e, f] with x_src = a x + b y + c."""
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


def rotate_image_nearest(img, angle_deg, center_xy):
    """(C, H, W) float tensor -> rotated copy (nearest neighbour, zeros outside)."""
    C, H, W = img.shape
    theta = torch.tensor(inverse_rotation_matrix(angle_deg, center_xy, W, H), dtype=img.dtype).reshape(1, 2, 3)
    base = torch.empty(1, H, W, 3, dtype=img.dtype)
    base[..., 0].copy_(torch.linspace(-W * 0.5 + 0.5, W * 0.5 + 0.5 - 1, steps=W))
    base[..., 1].copy_(torch.linspace(-H * 0.5 + 0.5, H * 0.5 + 0.5 - 1, steps=H).unsqueeze(-1))
    base[..., 2].fill_(1)
    rescaled = theta.transpose(1, 2) / torch.tensor([0.5 * W, 0.5 * H], dtype=img.dtype)
    grid = base.view(1, H * W, 3).bmm(rescaled).view(1, H, W, 2)
    return F.grid_sample(img[None], grid, mode='nearest', padding_mode='zeros', align_corners=False)[0]


def rotate_prev_bev(prev_bev, angles_deg, bev_h, bev_w, center_xy):
    """prev_bev (HW, bs, C) -> rotated (HW, bs, C); sample i by ``angles_deg[i]`` (can_bus[-1])."""
    out = prev_bev.clone()
    for i in range(prev_bev.shape[1]):
        img = prev_bev[:, i].reshape(bev_h, bev_w, -1).permute(2, 0, 1)
        rot = rotate_image_nearest(img, float(angles_deg[i]), center_xy)
        out[:, i] = rot.permute(1, 2, 0).reshape(bev_h * bev_w, -1)
    return out


def flatten_features(mlvl_feats, cams_embeds, level_embeds):
    """list of (bs, num_cam, C, h, w) -> feat_flatten (num_cam, sum(hw), bs, C), spatial_shapes
    (L, 2) int64, level_start_index (L,) int64."""
    flat, shapes = [], []
    for lvl, feat in enumerate(mlvl_feats):
        bs, num_cam, c, h, w = feat.shape
        feat = torch.nan_to_num(feat, nan=0.0, posinf=0.0, neginf=0.0)
        feat = feat.flatten(3).permute(1, 0, 3, 2)
        if cams_embeds is not None:
            feat = feat + cams_embeds[:, None, None, :].to(feat.dtype)
        feat = feat + level_embeds[None, None, lvl:lvl + 1, :].to(feat.dtype)
        shapes.append((h, w))
        flat.append(feat)
    flat = torch.cat(flat, 2).permute(0, 2, 1, 3)
    shapes = torch.as_tensor(shapes, dtype=torch.long)
    starts = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    return flat, shapes, starts
