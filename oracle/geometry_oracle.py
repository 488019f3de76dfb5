"""Oracle for the BEV geometry feeding spatial cross-attention (test infrastructure).

CPU/torch restatement of ``BEVFormerEncoder.get_reference_points`` and
``BEVFormerEncoder.point_sampling``
(``projects/mmdet3d_plugin/bevformer/modules/encoder.py:47-86`` and ``:89-241``),
without the img_metas plumbing: callers pass ``lidar2img`` (B, num_cam, 4, 4) and
the image height / width of camera 0 of sample 0, which is what the reference
normalises every camera with (``encoder.py:196-226``).
"""
import torch


def reference_points_3d(H, W, Z=8.0, num_points_in_pillar=4, bs=1,
                        device='cpu', dtype=torch.float32):
    """(bs, D, H*W, 3) pillar points in [0,1]^3; query index = y*W + x.

    encoder.py:62-72: zs = linspace(0.5, Z-0.5, D)/Z, xs = linspace(0.5, W-0.5, W)/W,
    ys likewise over H.
    """
    D = num_points_in_pillar
    zs = torch.linspace(0.5, Z - 0.5, D, dtype=dtype, device=device) / Z
    xs = torch.linspace(0.5, W - 0.5, W, dtype=dtype, device=device) / W
    ys = torch.linspace(0.5, H - 0.5, H, dtype=dtype, device=device) / H
    pts = torch.empty(D, H, W, 3, dtype=dtype, device=device)
    pts[..., 0] = xs.view(1, 1, W)
    pts[..., 1] = ys.view(1, H, 1)
    pts[..., 2] = zs.view(D, 1, 1)
    return pts.view(1, D, H * W, 3).repeat(bs, 1, 1, 1)


def reference_points_2d(H, W, bs=1, device='cpu', dtype=torch.float32):
    """(bs, H*W, 1, 2) BEV-plane cell centres (x, y) in [0,1]; encoder.py:75-86."""
    ys = torch.linspace(0.5, H - 0.5, H, dtype=dtype, device=device) / H
    xs = torch.linspace(0.5, W - 0.5, W, dtype=dtype, device=device) / W
    ref = torch.empty(H, W, 2, dtype=dtype, device=device)
    ref[..., 0] = xs.view(1, W)
    ref[..., 1] = ys.view(H, 1)
    return ref.view(1, H * W, 1, 2).repeat(bs, 1, 1, 1)


def point_sampling(ref_3d, pc_range, lidar2img, img_h, img_w):
    """Project pillar points into every camera.

    ref_3d (B, D, HW, 3) in [0,1]; lidar2img (B, num_cam, 4, 4) fp32.
    Returns reference_points_cam (num_cam, B, HW, D, 2) fp32 and
    bev_mask (num_cam, B, HW, D) bool.

    Follows encoder.py:147-239: scale to metres by pc_range (:149-154), append 1
    (:156-157), fp32 matmul with lidar2img (:176-183), mask = z > 1e-5 (:186),
    divide xy by max(z, 1e-5) (:187-188), divide by image W / H (:225-226), mask &=
    strict 0 < x,y < 1 (:228-231), permute (:238-239).
    """
    pts = ref_3d.clone().to(torch.float32)
    pts[..., 0:1] = pts[..., 0:1] * (pc_range[3] - pc_range[0]) + pc_range[0]
    pts[..., 1:2] = pts[..., 1:2] * (pc_range[4] - pc_range[1]) + pc_range[1]
    pts[..., 2:3] = pts[..., 2:3] * (pc_range[5] - pc_range[2]) + pc_range[2]
    pts = torch.cat([pts, torch.ones_like(pts[..., :1])], dim=-1)     # (B, D, HW, 4)
    B, D, HW, _ = pts.shape
    l2i = torch.as_tensor(lidar2img, dtype=torch.float32, device=pts.device)
    num_cam = l2i.shape[1]
    # (D, B, cam, HW, 4, 1) and (D, B, cam, HW, 4, 4), materialised like the
    # reference does with .repeat so that the matmul kernel choice is the same.
    p = pts.permute(1, 0, 2, 3).reshape(D, B, 1, HW, 4).repeat(1, 1, num_cam, 1, 1).unsqueeze(-1)
    m = l2i.view(1, B, num_cam, 1, 4, 4).repeat(D, 1, 1, HW, 1, 1)
    cam = torch.matmul(m, p).squeeze(-1)                              # (D, B, cam, HW, 4)
    eps = 1e-5
    mask = cam[..., 2:3] > eps
    uv = cam[..., 0:2] / torch.maximum(cam[..., 2:3], torch.ones_like(cam[..., 2:3]) * eps)
    uv[..., 0] /= float(img_w)
    uv[..., 1] /= float(img_h)
    mask = (mask & (uv[..., 1:2] > 0.0) & (uv[..., 1:2] < 1.0)
            & (uv[..., 0:1] < 1.0) & (uv[..., 0:1] > 0.0))
    mask = torch.nan_to_num(mask)
    uv = uv.permute(2, 1, 3, 0, 4)                                    # (cam, B, HW, D, 2)
    mask = mask.permute(2, 1, 3, 0, 4).squeeze(-1)                    # (cam, B, HW, D)
    return uv, mask


def camera_hit_lists(bev_mask):
    """Per-camera BEV-query index lists, from batch element 0 only.

    spatial_cross_attention.py:135-139: ``mask_per_img[0].sum(-1).nonzero()``.
    Returns (list of int64 index tensors, max_len).
    """
    lists = [m[0].sum(-1).nonzero().squeeze(-1) for m in bev_mask]
    return lists, max(len(x) for x in lists)
