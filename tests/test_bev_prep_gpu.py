"""GPU parity of the encoder's pre-processing: the B200
``PerceptionTransformer.get_bev_features`` front half against the golden vectors recorded from the
UNMODIFIED reference class (tests/golden/make_golden.py, ``bev_features_small``) and against the
oracle restatement on other shapes.  fp32 is bit-exact: same shift, same rotated pixels, same
flattened features."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import bev_features_oracle as B
from tests import golden_util as gu
from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


class _CaptureEncoder(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        self.captured = None

    def forward(self, bev_queries, key, value, **kw):
        self.captured = dict(bev_queries=bev_queries, feat_flatten=key, **kw)
        return bev_queries.permute(1, 0, 2)


def _register_capture():
    from apollo_vision_net_b200.registry import TRANSFORMER_LAYER_SEQUENCE as REG
    if REG.get('CaptureEncoder') is None:
        REG.register_module(name='CaptureEncoder', module=_CaptureEncoder)


def test_get_bev_features_golden():
    from apollo_vision_net_b200.modules import PerceptionTransformer
    _register_capture()
    g = gu.load('bev_features_small')
    bs, num_cam, C, bev_h, bev_w = (int(x) for x in g['cfg'])
    L = len(g['levels'])
    trf = PerceptionTransformer(num_feature_levels=L, num_cams=num_cam, embed_dims=C,
                                encoder=dict(type='CaptureEncoder'), decoder=None,
                                rotate_center=[bev_w // 2, bev_h // 2])
    trf.load_state_dict(gu.params(g))
    trf.to(DEV).eval()
    feats = [gu.T(g[f'feat_{i}'], DEV) for i in range(L)]
    metas = [dict(can_bus=list(g['can_bus'][b])) for b in range(bs)]
    out = trf.get_bev_features(feats, gu.T(g['bev_queries'], DEV), bev_h, bev_w,
                               grid_length=tuple(g['grid_length']), bev_pos=gu.T(g['bev_pos'], DEV),
                               prev_bev=gu.T(g['prev_bev'], DEV), img_metas=metas)
    cap = trf.encoder.captured
    assert np.array_equal(cap['shift'].cpu().numpy(), g['out_shift'])
    assert np.array_equal(cap['spatial_shapes'].cpu().numpy(), g['out_spatial_shapes'])
    assert np.array_equal(cap['level_start_index'].cpu().numpy(), g['out_level_start_index'])
    assert np.array_equal(cap['prev_bev'].cpu().numpy(), g['out_prev_bev'])          # same pixels picked
    assert np.array_equal(cap['feat_flatten'].detach().cpu().numpy(), g['out_feat_flatten'])  # incl. the zeroed NaN / inf
    assert np.array_equal(cap['bev_pos'].detach().cpu().numpy(), g['out_bev_pos'])
    assert rel_err(cap['bev_queries'], g['out_bev_queries']) <= 1e-6                 # can_bus MLP (cuBLAS GEMM)
    assert rel_err(out, g['out']) <= 1e-6


@pytest.mark.parametrize('bev_h,bev_w,C,dtype', [(50, 50, 256, torch.float32), (37, 61, 64, torch.float32),
                                                 (200, 200, 256, torch.bfloat16), (9, 7, 8, torch.float16)])
def test_rotate_prev_bev_matches_torchvision_semantics(bev_h, bev_w, C, dtype):
    from apollo_vision_net_b200.bev_prep import rotate_prev_bev
    g = torch.Generator().manual_seed(bev_h * 131 + bev_w)
    bs = 3
    prev = torch.randn(bev_h * bev_w, bs, C, generator=g).to(dtype)
    angles = [0.0, 17.3, -121.75]
    center = [bev_w // 2, bev_h // 2]
    got = rotate_prev_bev(prev.to(DEV), angles, bev_h, bev_w, center).cpu()
    want = B.rotate_prev_bev(prev.float(), angles, bev_h, bev_w, center).to(dtype)
    # nearest-neighbour picks are discrete: identical except (at most) a handful of pixels whose source
    # coordinate sits within rounding of a half-integer
    same = (got.view(bev_h * bev_w, bs, C) == want).all(-1)
    assert float((~same).float().mean()) <= 2e-4
    assert torch.equal(got[:, 0], prev[:, 0])                                          # angle 0: identity


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_flatten_features_forward_backward(dtype):
    from apollo_vision_net_b200.bev_prep import flatten_features
    g = torch.Generator().manual_seed(77)
    bs, num_cam, C = 2, 6, 256
    levels = [(29, 50), (15, 25), (8, 13), (4, 7)]
    feats = [torch.randn(bs, num_cam, C, h, w, generator=g).to(dtype) for h, w in levels]
    feats[0][1, 3, 100, 5, 7] = float('nan')
    ce = torch.randn(num_cam, C, generator=g).to(dtype)
    le = torch.randn(len(levels), C, generator=g).to(dtype)
    ref_in = [f.clone().float().requires_grad_(True) for f in feats]
    ce_r, le_r = ce.clone().float().requires_grad_(True), le.clone().float().requires_grad_(True)
    want, shapes_w, starts_w = B.flatten_features(ref_in, ce_r, le_r)
    go = torch.randn(want.shape, generator=g).to(dtype)
    want.backward(go.float())
    dev_in = [f.clone().to(DEV).requires_grad_(True) for f in feats]
    ce_d, le_d = ce.clone().to(DEV).requires_grad_(True), le.clone().to(DEV).requires_grad_(True)
    got, shapes, starts = flatten_features(dev_in, ce_d, le_d)
    got.backward(go.to(DEV))
    assert torch.equal(shapes.cpu(), shapes_w) and torch.equal(starts.cpu(), starts_w)
    if dtype == torch.float32:
        assert torch.equal(got.cpu(), want)
    else:
        assert rel_err(got, want) <= 1e-2
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for a, b in zip(dev_in, ref_in):
        assert rel_err(a.grad, b.grad) <= tol
    assert float(dev_in[0].grad[1, 3, 100, 5, 7]) == 0.0                                # zeroed input: no gradient
    assert rel_err(ce_d.grad, ce_r.grad) <= tol and rel_err(le_d.grad, le_r.grad) <= tol
