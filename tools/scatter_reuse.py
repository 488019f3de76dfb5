"""How many grad_value updates of the SCA backward fall on the same (camera, head, pixel) slot?

CPU-only (numpy) analysis of the base workload of bench.py (200x200 BEV, 6 cameras, 4 levels, 8 heads,
8 points = 2 per Z-anchor): for every BEV tile shape it prints, per pyramid level, the number of corner
updates of a tile and the number of DISTINCT slots they hit -- the upper bound of what a tile-level
pre-reduction (shared memory or registers) can remove from the L2 reduction traffic that bounds
`sca_bwd`.  Geometry restated from encoder.py:123-241 of the reference; offsets =
ring bias + N(0, 0.32) (what `Linear(256 -> 512)` with N(0, 0.02) weights gives on unit-variance queries).

    python tools/scatter_reuse.py [--bev 200]
"""
import argparse
import importlib.util
import math
import os

import numpy as np

_spec = importlib.util.spec_from_file_location(
    'synthetic', os.path.join(os.path.dirname(__file__), '..', 'apollo-vision-net_b200', 'synthetic.py'))
syn = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(syn)

M, P, D = 8, 8, 4


def project(bev, lidar2img, img_shape):
    """reference_points_cam (6, HW, D, 2) and bev_mask (6, HW, D) of encoder.py:123-241 (float64)."""
    H = W = bev
    pc = syn.PC_RANGE
    zs = (np.linspace(0.5, 8 - 0.5, D) / 8)[:, None, None] * np.ones((D, H, W))
    xs = (np.linspace(0.5, W - 0.5, W) / W)[None, None, :] * np.ones((D, H, W))
    ys = (np.linspace(0.5, H - 0.5, H) / H)[None, :, None] * np.ones((D, H, W))
    pts = np.stack([xs * (pc[3] - pc[0]) + pc[0], ys * (pc[4] - pc[1]) + pc[1],
                    zs * (pc[5] - pc[2]) + pc[2], np.ones_like(xs)], -1).reshape(D, H * W, 4)
    cam = np.einsum('cij,dqj->cqdi', lidar2img.astype(np.float64), pts)      # (6, HW, D, 4)
    eps = 1e-5
    mask = cam[..., 2] > eps
    uv = cam[..., :2] / np.maximum(cam[..., 2:3], eps)
    uv[..., 0] /= img_shape[1]
    uv[..., 1] /= img_shape[0]
    mask &= (uv[..., 1] > 0) & (uv[..., 1] < 1) & (uv[..., 0] > 0) & (uv[..., 0] < 1)
    return uv, mask


def ring_bias():
    theta = np.arange(M) * (2 * math.pi / M)
    g = np.stack([np.cos(theta), np.sin(theta)], -1)
    g = g / np.abs(g).max(-1, keepdims=True)
    return g[:, None, :] * (np.arange(P) + 1)[None, :, None]                 # (M, P, 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--bev', type=int, default=200)
    args = ap.parse_args()
    bev = args.bev
    l2i, img_shape = syn.camera_rig(1.0)
    uv, mask = project(bev, l2i[0], img_shape)
    hit = mask.any(-1)                                                       # (6, HW)
    print(f'BEV {bev}x{bev}: (camera, query) pairs {int(hit.sum())}, per camera {hit.sum(1).tolist()}')
    rng = np.random.RandomState(0)
    bias = ring_bias()
    levels = syn.LEVELS_BASE
    tiles = [(2, 4), (8, 8), (16, 16), (32, 32)]
    qy, qx = np.divmod(np.arange(bev * bev), bev)
    # totals[tile][level] = [updates, unique slots per (tile, cam, head), unique per (row = q, cam, head)]
    totals = {t: np.zeros((len(levels), 2), dtype=np.int64) for t in tiles}
    row_unique = np.zeros(len(levels), dtype=np.int64)
    per_slot = [[] for _ in levels]
    for c in range(6):
        qs = np.nonzero(hit[c])[0]
        ref = uv[c, qs]                                                      # (n, D, 2)
        for li, (Hl, Wl) in enumerate(levels):
            off = bias[None] + rng.randn(len(qs), M, P, 2) * 0.32            # (n, M, P, 2) pixels
            z = np.arange(P) % D                                             # point p = k * D + z
            loc = ref[:, None, z, :] + off / np.array([Wl, Hl])              # (n, M, P, 2)
            x = loc[..., 0] * Wl - 0.5
            y = loc[..., 1] * Hl - 0.5
            x0 = np.floor(x).astype(np.int64)
            y0 = np.floor(y).astype(np.int64)
            keys, tq, th = [], [], []
            for dy in (0, 1):
                for dx in (0, 1):
                    xx, yy = x0 + dx, y0 + dy
                    ok = (xx >= 0) & (xx < Wl) & (yy >= 0) & (yy < Hl)
                    n_i, m_i, _ = np.nonzero(ok)
                    keys.append((yy[ok] * Wl + xx[ok]))
                    tq.append(qs[n_i])
                    th.append(m_i)
            pix = np.concatenate(keys)
            q = np.concatenate(tq)
            h = np.concatenate(th)
            slot = h * (Hl * Wl) + pix                                       # (head, pixel) of this camera
            nslot = M * Hl * Wl
            row_unique[li] += len(np.unique(q * nslot + slot))
            _, cnt = np.unique(slot, return_counts=True)          # contributors per (camera, head, pixel) slot
            per_slot[li].append(cnt)
            for t in tiles:
                tid = (qy[q] // t[0]) * ((bev + t[1] - 1) // t[1]) + qx[q] // t[1]
                totals[t][li, 0] += len(slot)
                totals[t][li, 1] += len(np.unique(tid * nslot + slot))
    upd = totals[tiles[0]][:, 0]
    print('level  updates(M)  distinct/updates within one (query, head) row')
    for li in range(len(levels)):
        print(f'  {li}     {upd[li] / 1e6:7.2f}     {row_unique[li] / upd[li]:.3f}')
    print('contributors per touched (camera, head, pixel) slot over the whole frame (an owner-computes backward '
          'would walk these lists):')
    for li, (Hl, Wl) in enumerate(levels):
        c = np.concatenate(per_slot[li])
        print(f'  L{li}: {len(c)} of {6 * M * Hl * Wl} slots touched, updates per slot mean {c.mean():.0f}  '
              f'median {np.median(c):.0f}  p99 {np.percentile(c, 99):.0f}  max {c.max()}')
    for t in tiles:
        tot = totals[t]
        n_tiles = ((bev + t[0] - 1) // t[0]) * ((bev + t[1] - 1) // t[1])
        line = '  '.join(f'L{li} {tot[li, 1] / tot[li, 0]:.3f}' for li in range(len(levels)))
        slots = tot[:, 1] / n_tiles / M          # distinct slots per tile and head, summed over cameras
        kb = '  '.join(f'L{li} {s * 64 / 1024:.0f} KB' for li, s in enumerate(slots))
        print(f'tile {t[0]}x{t[1]}: distinct/updates {line}  overall {tot[:, 1].sum() / tot[:, 0].sum():.3f}'
              f' | fp16 slots per (tile, head), all cameras: {kb}')


if __name__ == '__main__':
    main()
