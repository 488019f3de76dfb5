"""Pins the CPU oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py), and -- when /root/reference is present -- against the
reference executed live.  No GPU, no product code."""
import numpy as np
import pytest
import torch

from oracle import geometry_oracle as G
from oracle.msda_oracle import (msda_numpy, msda_numpy_backward, msda_torch, msda_torch_fwd_bwd)
from oracle.refshim import load_reference, reference_available
from tests import golden_util as gu
from tests.util import rel_err


def test_op_oracle_matches_golden_bit_for_bit():
    g = gu.load('op_small')
    out, gv, gl, ga = msda_torch_fwd_bwd(gu.T(g['value']), gu.T(g['shapes']), gu.T(g['loc']),
                                         gu.T(g['attn']), gu.T(g['grad_out']))
    # same formula, same library kernels: identical on the machine that made the fixtures,
    # and within float rounding anywhere else
    assert rel_err(out, g['out']) <= 1e-6
    assert rel_err(gv, g['grad_value']) <= 1e-6
    assert rel_err(gl, g['grad_loc']) <= 1e-5
    assert rel_err(ga, g['grad_attn']) <= 1e-6


def test_explicit_bilinear_oracle_matches_golden_fp64():
    g = gu.load('op_small')
    levels = [tuple(x) for x in g['shapes'].tolist()]
    out = msda_numpy(g['value'], levels, g['loc'], g['attn'])
    assert rel_err(out, g['out64']) <= 1e-12
    gv, gl, ga = msda_numpy_backward(g['value'], levels, g['loc'], g['attn'], g['grad_out'])
    assert rel_err(gv, g['grad_value64']) <= 1e-12
    assert rel_err(gl, g['grad_loc64']) <= 1e-10
    assert rel_err(ga, g['grad_attn64']) <= 1e-12


def test_geometry_oracle_matches_golden_exactly():
    g = gu.load('geometry_tiny')
    H, W = (int(x) for x in g['bev_hw'])
    D = int(g['D'])
    bs = g['lidar2img'].shape[0]
    r3 = G.reference_points_3d(H, W, 8.0, D, bs=bs)
    r2 = G.reference_points_2d(H, W, bs=bs)
    assert np.array_equal(r3.numpy(), g['ref_3d'])
    assert np.array_equal(r2.numpy(), g['ref_2d'])
    uv, mask = G.point_sampling(r3, list(g['pc_range']), g['lidar2img'], int(g['img_shape'][0]),
                                int(g['img_shape'][1]))
    assert np.array_equal(mask.numpy(), gu.unpack_mask(g))           # bit-exact
    assert rel_err(uv, g['ref_cam']) <= 1e-6
    lists, max_len = G.camera_hit_lists(mask)
    assert [len(x) for x in lists] == g['hit_count'].tolist()
    for i, x in enumerate(lists):
        assert np.array_equal(x.numpy().astype(np.int32), g[f'hit_index_{i}'])


def test_sca_oracle_matches_golden():
    g = gu.load('sca_small')
    m = gu.build_sca(g, 'oracle')
    out, gq, gf = gu.run_sca(m, g)
    assert rel_err(out, g['out']) <= 1e-6
    assert rel_err(gq, g['grad_query']) <= 1e-5
    assert rel_err(gf, g['grad_feat']) <= 1e-5
    for n, p in m.named_parameters():
        assert rel_err(p.grad, gu.pgrads(g)[n]) <= 1e-5, n
    # stand-alone MSDeformableAttention3D contract
    da = m.deformable_attention
    o = da(query=gu.T(g['da_query']), key=gu.T(g['da_value']), value=gu.T(g['da_value']),
           reference_points=gu.T(g['da_ref']), spatial_shapes=gu.T(g['shapes']),
           level_start_index=gu.T(g['starts']))
    assert rel_err(o, g['da_out']) <= 1e-6


@pytest.mark.parametrize('name', ['tsa_prev', 'tsa_first_frame'])
def test_tsa_oracle_matches_golden(name):
    g = gu.load(name)
    m = gu.build_tsa(g, 'oracle')
    out, gq, gp = gu.run_tsa(m, g)
    assert rel_err(out, g['out']) <= 1e-6
    assert rel_err(gq, g['grad_query']) <= 1e-5
    if gp is not None:
        assert rel_err(gp, g['grad_prev']) <= 1e-5
    for n, p in m.named_parameters():
        assert rel_err(p.grad, gu.pgrads(g)[n]) <= 1e-5, n


def test_decoder_oracle_matches_golden():
    g = gu.load('decoder_small')
    m = gu.build_decoder(g, 'oracle')
    out, gq, gv, gr = gu.run_decoder(m, g)
    assert rel_err(out, g['out']) <= 1e-6
    assert rel_err(gq, g['grad_query']) <= 1e-5
    assert rel_err(gv, g['grad_value']) <= 1e-5
    assert rel_err(gr, g['grad_ref']) <= 1e-5


def test_detection_decoder_host_logic_matches_reference_golden():
    """DetectionTransformerDecoder of the package (layer loop, refinement of (x, y, z) reference
    points through reg_branches outputs 0:2 and 4, detach, stacking) against the fixture produced
    by the UNMODIFIED reference class (decoder.py:50-126); the cross-attention inside the stub
    layers is the CPU oracle here, the CUDA module in tests/test_modules_gpu.py."""
    g = gu.load('det_decoder_small')
    dec, reg = gu.build_det_decoder(g, 'oracle')
    inter, refs, gq, gv = gu.run_det_decoder(dec, reg, g)
    assert inter.shape == g['inter'].shape and refs.shape == g['refs'].shape
    assert rel_err(inter, g['inter']) <= 1e-6
    assert rel_err(refs, g['refs']) <= 1e-6
    assert rel_err(gq, g['grad_query']) <= 1e-5
    assert rel_err(gv, g['grad_value']) <= 1e-5


def test_encoder_oracle_matches_reference_golden():
    """OracleBEVFormerEncoder (the checker of the CUDA encoder tests) against the fixture produced by
    the UNMODIFIED reference BEVFormerEncoder / BEVFormerLayer / MyCustomBaseTransformerLayer: two
    layers, history, CAN-bus shift, two feature levels, forward and every gradient."""
    from oracle.modules_oracle import OracleBEVFormerEncoder
    import apollo_vision_net_b200.synthetic as syn
    g = gu.load('encoder_small')
    bs, H, W, C, heads = (int(x) for x in g['cfg'])
    levels = [tuple(int(v) for v in r) for r in g['levels']]
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    o = OracleBEVFormerEncoder(num_layers=2, pc_range=syn.PC_RANGE, num_points_in_pillar=4,
                               embed_dims=C, feedforward_channels=2 * C, num_levels=len(levels))
    o.load_state_dict(gu.params(g))
    o.eval()
    bevq, prev, feat = (gu.T(g[k], grad=True) for k in ('bev_query', 'prev_bev', 'feat'))
    out = o(bevq, feat, feat, bev_h=H, bev_w=W, bev_pos=gu.T(g['bev_pos']),
            spatial_shapes=torch.tensor(shapes_l), level_start_index=torch.tensor(starts_l),
            prev_bev=prev, shift=gu.T(g['shift']), lidar2img=g['lidar2img'],
            img_h=int(g['img_shape'][0]), img_w=int(g['img_shape'][1]))
    out.backward(gu.T(g['grad_out']))
    assert rel_err(out, g['out']) <= 1e-5
    assert rel_err(bevq.grad, g['grad_bev_query']) <= 1e-4
    assert rel_err(prev.grad, g['grad_prev_bev']) <= 1e-4
    assert rel_err(feat.grad, g['grad_feat']) <= 1e-4
    pg = gu.pgrads(g)
    for n, p in o.named_parameters():
        assert rel_err(p.grad, pg[n]) <= 1e-4, n


def test_maptrv2_decoder_host_logic_matches_reference_golden():
    """MapTRv2Decoder + MapTRv2DecoupledDetrTransformerDecoderLayer of the package (the
    inter-vector / intra-vector reshapes around the two self-attentions, cross-attention call
    contract, 2-d reference refinement) against the fixture produced by the UNMODIFIED reference
    classes (maptrv2/modules/decoder.py:10-213), one-to-many mask included; the deformable
    cross-attention is the CPU oracle here, the CUDA module in tests/test_modules_gpu.py."""
    g = gu.load('maptrv2_decoder_small')
    dec, reg = gu.build_maptrv2_decoder(g, 'oracle')
    inter, refs, gq, gv = gu.run_maptrv2_decoder(dec, reg, g)
    assert inter.shape == g['inter'].shape and refs.shape == g['refs'].shape
    assert rel_err(inter, g['inter']) <= 1e-5
    assert rel_err(refs, g['refs']) <= 1e-5
    assert rel_err(gq, g['grad_query']) <= 1e-4
    assert rel_err(gv, g['grad_value']) <= 1e-4


@pytest.mark.skipif(not reference_available(), reason='/root/reference not present on this box')
def test_oracle_equals_live_reference():
    """Where the reference sources exist (the build container), run them live against the oracle."""
    import warnings
    warnings.filterwarnings('ignore')
    ref = load_reference()
    torch.manual_seed(4)
    levels = [(6, 8), (3, 4)]
    value = torch.randn(2, 60, 4, 8)
    loc = torch.rand(2, 11, 4, 2, 3, 2) * 1.2 - 0.1
    att = torch.softmax(torch.randn(2, 11, 4, 6), -1).view(2, 11, 4, 2, 3)
    shapes = torch.tensor(levels)
    assert torch.equal(ref.msda_pytorch_2d(value, shapes, loc, att), msda_torch(value, shapes, loc, att))
    from oracle.modules_oracle import OracleTemporalSelfAttention
    r = ref.TemporalSelfAttention(embed_dims=32, num_heads=4, num_levels=1, num_points=2)
    o = OracleTemporalSelfAttention(embed_dims=32, num_heads=4, num_levels=1, num_points=2)
    torch.nn.init.normal_(r.sampling_offsets.weight, 0, 0.02)
    o.load_state_dict(r.state_dict())
    r.eval(), o.eval()
    q = torch.randn(1, 20, 32)
    ref2d = G.reference_points_2d(4, 5, bs=1)
    hy = torch.stack([ref2d, ref2d], 1).reshape(2, 20, 1, 2)
    kw = dict(query_pos=torch.randn(1, 20, 32), reference_points=hy,
              spatial_shapes=torch.tensor([[4, 5]]), level_start_index=torch.tensor([0]))
    assert torch.equal(r(q, None, None, **kw), o(q, None, None, **kw))


def test_bev_features_oracle_matches_golden_exactly():
    """get_bev_features pre-processing (transformer.py:119-298): the restatement reproduces what
    the reference hands to its encoder -- shift, rotated prev_bev (torchvision rotate, nearest),
    flattened features with camera / level embeddings -- bit for bit."""
    import torch.nn as nn
    from oracle import bev_features_oracle as B
    g = gu.load('bev_features_small')
    bs, num_cam, C, bev_h, bev_w = (int(x) for x in g['cfg'])
    levels = [tuple(int(v) for v in r) for r in g['levels']]
    prm = gu.params(g)
    shift = B.can_bus_shift(g['can_bus'], tuple(g['grid_length']), bev_h, bev_w)
    assert np.array_equal(shift.numpy(), g['out_shift'])
    prev = gu.T(g['prev_bev']).permute(1, 0, 2).contiguous()
    rot = B.rotate_prev_bev(prev, g['can_bus'][:, -1], bev_h, bev_w, [bev_w // 2, bev_h // 2])
    assert np.array_equal(rot.numpy(), g['out_prev_bev'])
    assert not np.array_equal(g['out_prev_bev'], prev.numpy())            # the rotation did something
    feats = [gu.T(g[f'feat_{i}']) for i in range(len(levels))]
    flat, shapes, starts = B.flatten_features(feats, prm['cams_embeds'], prm['level_embeds'])
    assert np.array_equal(flat.numpy(), g['out_feat_flatten'])
    assert np.array_equal(shapes.numpy(), g['out_spatial_shapes'])
    assert np.array_equal(starts.numpy(), g['out_level_start_index'])
    # can_bus MLP (transformer.py:206-210): Linear-ReLU-Linear-ReLU-LayerNorm added to every query
    mlp = nn.Sequential(nn.Linear(18, C // 2), nn.ReLU(), nn.Linear(C // 2, C), nn.ReLU())
    mlp.add_module('norm', nn.LayerNorm(C))
    mlp.load_state_dict({k[len('can_bus_mlp.'):]: v for k, v in prm.items() if k.startswith('can_bus_mlp.')})
    cb = mlp(torch.tensor(g['can_bus'], dtype=torch.float32))[None]
    q = gu.T(g['bev_queries']).unsqueeze(1).repeat(1, bs, 1) + cb
    assert np.allclose(q.detach().numpy(), g['out_bev_queries'], rtol=0, atol=1e-6)
