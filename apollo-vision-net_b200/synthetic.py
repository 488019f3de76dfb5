"""Synthetic inputs for the BEVFormer hot path (numpy only; no oracle, no CUDA).

The reference needs nuScenes files for real ``lidar2img`` matrices
(``projects/mmdet3d_plugin/datasets/nuscenes_dataset.py:220-249``); tests and
benchmarks here use a nuScenes-like 6-camera pinhole rig instead (SURVEY.md
section 8d): yaw {0, +-55, +-110, 180} degrees, fx = fy = 1266, cx = 816, cy = 491
(rear camera fx = 809, cx = 829, cy = 481), each camera 0.5 m along its viewing
direction from the lidar origin and 0.34 m below it, 900x1600 images padded to
928 rows, optionally rescaled the way ``RandomScaleImageMultiViewImage`` rescales
``lidar2img`` (``datasets/pipelines/transform_3d.py:292-321``).
"""
import math
import numpy as np

PC_RANGE = [-50.0, -50.0, -5.0, 50.0, 50.0, 3.0]

# (H_l, W_l) of the image feature levels, SURVEY.md section 8 shape table
LEVELS_BASE = [(116, 200), (58, 100), (29, 50), (15, 25)]
LEVELS_TINY = [(28, 48)]

_YAWS_DEG = [0.0, 55.0, -55.0, 110.0, -110.0, 180.0]


def _camera(yaw_deg, fx, cx, cy):
    """4x4 lidar->image matrix K.[R|t] for a camera looking along `yaw` in the lidar xy plane.

    Lidar frame: x right, y forward, z up.  Camera frame: x right, y down, z forward.
    """
    a = math.radians(yaw_deg)
    fwd = np.array([math.sin(a), math.cos(a), 0.0])
    right = np.array([math.cos(a), -math.sin(a), 0.0])
    down = np.array([0.0, 0.0, -1.0])
    R = np.stack([right, down, fwd], 0)            # rows: camera axes in lidar coords
    centre = 0.5 * fwd + np.array([0.0, 0.0, -0.34])
    t = -R @ centre
    ext = np.eye(4)
    ext[:3, :3] = R
    ext[:3, 3] = t
    K = np.eye(4)
    K[0, 0] = fx
    K[1, 1] = fx
    K[0, 2] = cx
    K[1, 2] = cy
    return K @ ext


def camera_rig(scale=1.0, bs=1, jitter=0.0, seed=0):
    """lidar2img (bs, 6, 4, 4) float32 and img_shape (H, W, 3) of the padded, scaled image.

    ``jitter`` (degrees) perturbs the yaw per sample so that batch elements differ.
    """
    rng = np.random.RandomState(seed)
    S = np.eye(4)
    S[0, 0] = scale
    S[1, 1] = scale
    out = np.zeros((bs, 6, 4, 4), dtype=np.float64)
    for b in range(bs):
        for i, yaw in enumerate(_YAWS_DEG):
            dy = float(rng.uniform(-jitter, jitter)) if (jitter > 0 and b > 0) else 0.0
            if i == 5:
                m = _camera(yaw + dy, 809.0, 829.0, 481.0)
            else:
                m = _camera(yaw + dy, 1266.0, 816.0, 491.0)
            out[b, i] = S @ m
    img_shape = (int(round(928 * scale)), int(round(1600 * scale)), 3)
    return out.astype(np.float32), img_shape


def level_tables(levels):
    """(spatial_shapes (L,2) int64 list, level_start_index (L,) list, Nk)."""
    starts = []
    s = 0
    for h, w in levels:
        starts.append(s)
        s += h * w
    return [list(x) for x in levels], starts, s
