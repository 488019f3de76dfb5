"""The BENCHED configuration against the CPU oracle, directly (VERDICT r01 "what's weak" 1-2).

BASELINE.json configs[1]: BEVFormer-base encoder, 200x200 BEV, 6 cameras, 4 feature levels
(116x200 .. 15x25), 8 points x 4 Z-anchors, temporal self-attention with history.

* ``test_base_layer_fp32_vs_oracle``: ONE encoder layer at the full size, fp32, CUDA (fused kernels,
  merged projection, fused post-norm tail) against ``OracleBEVFormerEncoder`` on the CPU: forward to
  1e-5, every gradient to 1e-4 -- with the entries that lie downstream of a sample sitting on a pixel
  boundary identified explicitly (``d out / d location`` is piecewise constant, so a sample within
  rounding distance of a pixel edge legitimately picks either one-sided derivative) instead of a
  loosened tolerance.
* ``test_bench_model_bf16_vs_oracle_layer_by_layer``: the exact model ``bench.py`` times (bf16,
  6 layers, the whole step replayed as a CUDA graph) against the fp32 oracle run one layer at a
  time on the CUDA model's own layer inputs, 1e-2; plus the last layer's parameter gradients.
"""
import copy

import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
C, HEADS, PILLAR = 256, 8, 4


def _randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    for n, p in module.named_parameters():
        if n.endswith('sampling_offsets.weight') or n.endswith('attention_weights.weight'):
            p.data = (torch.randn(p.shape, generator=g) * 0.02).to(p.dtype)


def _encoder_cfg(num_layers, num_levels, pc_range):
    return dict(
        type='BEVFormerEncoder', num_layers=num_layers, pc_range=pc_range,
        num_points_in_pillar=PILLAR, return_intermediate=False,
        transformerlayers=dict(
            type='BEVFormerLayer',
            attn_cfgs=[dict(type='TemporalSelfAttention', embed_dims=C, num_points=4, num_levels=1),
                       dict(type='SpatialCrossAttention', pc_range=pc_range, embed_dims=C,
                            deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                                      num_points=8, num_levels=num_levels))],
            feedforward_channels=512, ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))


def _near_pixel_edge(pix, eps):
    """True where a pixel coordinate is within ``eps`` of an integer (a bilinear cell boundary)."""
    return (pix - pix.round()).abs() < eps


def _flagged_queries(o64, tsa_off, sca_query, ref2d_hybrid, ref_cam, bev_mask, levels, H, W, eps):
    """BEV queries that own at least one sample within ``eps`` pixels of a cell boundary, computed
    from the float64 oracle run: TSA samples (temporal_self_attention.py:239-245) and SCA samples of
    the cameras that see the query (spatial_cross_attention.py:361-376)."""
    HW = H * W
    layer = o64.layers[0]
    # TSA: off (1, HW, M, Q, 1, P, 2); loc = ref[b*Q + j] + off / (W, H); pix = loc * (W, H) - 0.5
    M, Q, P = HEADS, 2, 4
    off = tsa_off.view(1, HW, M, Q, 1, P, 2)
    ref = ref2d_hybrid.view(1, Q, HW, 1, 2).permute(0, 2, 1, 3, 4)              # (1, HW, Q, 1, 2)
    wh = torch.tensor([W, H], dtype=torch.float64)
    pix = (ref[:, :, None, :, :, None, :] + off / wh) * wh - 0.5
    flag = _near_pixel_edge(pix, eps).flatten(2).any(-1)[0]                        # (HW,)
    # SCA: off = sampling_offsets(query) (1, HW, M, L, P, 2), anchor z = p % D
    da = layer.attentions[1].deformable_attention
    L, P2, D = len(levels), 8, PILLAR
    off = F.linear(sca_query, da.sampling_offsets.weight, da.sampling_offsets.bias).view(1, HW, M, L, P2 // D, D, 2)
    hw = torch.tensor([[w, h] for h, w in levels], dtype=torch.float64)            # (L, 2) = (W_l, H_l)
    hit = bev_mask.any(-1)                                                        # (cam, 1, HW)
    for cam in range(ref_cam.shape[0]):
        r = ref_cam[cam].double()                                                 # (1, HW, D, 2)
        loc = r[:, :, None, None, None, :, :] + off / hw[None, None, None, :, None, None, :]
        pix = loc * hw[None, None, None, :, None, None, :] - 0.5
        near = _near_pixel_edge(pix, eps).flatten(2).any(-1)[0]
        flag |= near & hit[cam, 0]
    return flag


def test_base_layer_fp32_vs_oracle():
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.synthetic as syn
    from oracle import geometry_oracle as G
    from oracle.modules_oracle import OracleBEVFormerEncoder
    H = W = 200
    HW = H * W
    levels = syn.LEVELS_BASE
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    l2i, img_shape = syn.camera_rig(1.0, bs=1)
    torch.set_num_threads(max(1, (torch.get_num_threads())))
    o = OracleBEVFormerEncoder(num_layers=1, pc_range=syn.PC_RANGE, num_points_in_pillar=PILLAR,
                               embed_dims=C, feedforward_channels=512, num_levels=len(levels), dropout=0.0)
    _randomize(o, 0)
    o.eval()
    enc = pkg.build_transformer_layer_sequence(_encoder_cfg(1, len(levels), syn.PC_RANGE))
    enc.load_state_dict(o.state_dict())
    enc.to(DEV).eval()
    o64 = copy.deepcopy(o).double()

    g = torch.Generator().manual_seed(77)
    bevq = torch.randn(HW, 1, C, generator=g)
    pos = torch.randn(HW, 1, C, generator=g)
    prev = torch.randn(HW, 1, C, generator=g)
    feat = torch.randn(6, Nk, 1, C, generator=g)
    go = torch.randn(1, HW, C, generator=g)
    shift = torch.tensor([[0.005, 0.003]])
    shapes, starts = torch.tensor(shapes_l), torch.tensor(starts_l)

    def run_oracle(model, dt):
        q, p, f = (t.to(dt).clone().requires_grad_(True) for t in (bevq, prev, feat))
        out = model(q, f, f, bev_h=H, bev_w=W, bev_pos=pos.to(dt), spatial_shapes=shapes,
                    level_start_index=starts, prev_bev=p, shift=shift.to(dt), lidar2img=l2i,
                    img_h=img_shape[0], img_w=img_shape[1])
        out.backward(go.to(dt))
        return out.detach(), q.grad, p.grad, f.grad

    # float64 oracle = the truth (also records what the flagging needs); fp32 oracle = the yardstick
    rec = {}
    tsa, sca = o64.layers[0].attentions
    h1 = tsa.sampling_offsets.register_forward_hook(lambda m, i, out: rec.__setitem__('tsa_off', out.detach()))
    h2 = sca.register_forward_pre_hook(lambda m, a: rec.__setitem__('sca_query', a[0].detach()))
    t_out, t_gq, t_gp, t_gf = run_oracle(o64, torch.float64)
    h1.remove()
    h2.remove()
    r_out, r_gq, r_gp, r_gf = run_oracle(o, torch.float32)

    q2, p2, f2 = (t.to(DEV).requires_grad_(True) for t in (bevq, prev, feat))
    launches0 = pkg.launch_count()
    out = enc(q2, f2, f2, bev_h=H, bev_w=W, bev_pos=pos.to(DEV), spatial_shapes=shapes.to(DEV),
              level_start_index=starts.to(DEV), prev_bev=p2, shift=shift.to(DEV), lidar2img=l2i,
              img_shape=img_shape)
    out.backward(go.to(DEV))
    torch.cuda.synchronize()
    assert pkg.launch_count() - launches0 >= 6

    # forward: 1e-5 against the truth (the fp32 oracle's own distance is the floor)
    e, y = rel_err(out, t_out), rel_err(r_out, t_out)
    assert e <= max(1e-5, 3.0 * y), f'forward {e:.3e} (fp32 oracle {y:.3e})'

    # samples within 1e-4 px of a bilinear cell boundary -> their queries' rows are downstream of a
    # (possible) flip; everything else holds the 1e-4 bar entry by entry
    r3 = G.reference_points_3d(H, W, 8.0, PILLAR, bs=1)
    ref_cam, mask = G.point_sampling(r3, syn.PC_RANGE, l2i, img_shape[0], img_shape[1])
    ref2d = G.reference_points_2d(H, W, bs=1).double()
    hybrid = torch.stack([ref2d + shift.double()[:, None, None, :], ref2d + shift.double()[:, None, None, :]], 1)
    # (the aliasing "bug" of encoder.py:309-311: both queue entries carry the shifted reference)
    flag = _flagged_queries(o64, rec['tsa_off'], rec['sca_query'], hybrid.reshape(2, HW, 1, 2),
                            ref_cam, mask, levels, H, W, eps=1e-4)
    n_flag = int(flag.sum())
    assert n_flag <= 0.15 * HW, f'{n_flag} of {HW} queries flagged: the check would be vacuous'
    keep = (~flag).to(torch.float64)

    def rows_ok(name, mine, oracle32, truth):
        """(HW, 1, C) gradients of per-query tensors: rows of unflagged queries to 1e-4."""
        mine = mine.detach().double().cpu()
        scale = truth.abs().max()
        err = ((mine - truth).abs().amax(-1)[:, 0] * keep).max() / scale
        yard = ((oracle32.double() - truth).abs().amax(-1)[:, 0] * keep).max() / scale
        assert float(err) <= max(1e-4, 3.0 * float(yard)), \
            f'{name}: unflagged rows off by {float(err):.3e} (fp32 oracle {float(yard):.3e}; {n_flag} flagged rows)'
        # the flagged rows may pick the other one-sided derivative, but the tensor as a whole stays close
        assert rel_l2(mine, truth) <= max(2e-3, 3.0 * rel_l2(oracle32, truth)), name

    rows_ok('grad bev_query', q2.grad, r_gq, t_gq)
    rows_ok('grad prev_bev', p2.grad, r_gp, t_gp)
    # value-side and parameter gradients are sums over thousands of samples: a flipped sample moves
    # one term of each sum, far below the bar
    e, y = rel_err(f2.grad, t_gf), rel_err(r_gf, t_gf)
    assert e <= max(1e-4, 3.0 * y), f'grad feat {e:.3e} (fp32 oracle {y:.3e})'
    og, tg = dict(o.named_parameters()), dict(o64.named_parameters())
    for n, p in enc.named_parameters():
        e, y = rel_err(p.grad, tg[n].grad), rel_err(og[n].grad, tg[n].grad)
        assert e <= max(1e-4, 3.0 * y), f'{n}: {e:.3e} (fp32 oracle {y:.3e})'


def test_bench_model_bf16_vs_oracle_layer_by_layer():
    """bench.py's own model and step (bf16, 6 layers, CUDA-graph replay).  Layer i of the fp32 oracle
    runs on the CUDA model's input of layer i (its bf16 activations are exactly representable in
    fp32) with the CUDA model's bf16-rounded weights, so each comparison sees one layer's worth of
    bf16 arithmetic: 1e-2 of the layer output's scale (north_star)."""
    import bench
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.synthetic as syn
    from oracle.modules_oracle import OracleBEVFormerEncoder
    H = W = 200
    levels = syn.LEVELS_BASE
    NL = 6
    dtype = torch.bfloat16
    torch.manual_seed(0)
    enc = pkg.build_transformer_layer_sequence(bench.encoder_cfg(NL, len(levels), syn.PC_RANGE))
    bench.randomize(enc, 0)
    enc.to(DEV).to(dtype).train()
    for m in enc.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    params = [p for p in enc.parameters() if p.requires_grad]
    l2i, img_shape = syn.camera_rig(1.0, bs=1)
    l2i_dev = torch.as_tensor(l2i).to(DEV)
    host = bench.host_inputs(H, W, levels, seed=1, dtype=dtype, pin=False)
    d = {k: v.to(DEV) for k, v in host.items()}

    def step():
        for p in params:
            p.grad = None
        feat = d['feat'].detach().requires_grad_(True)
        out = enc(d['bev_query'], feat, feat, bev_h=H, bev_w=W, bev_pos=d['bev_pos'],
                  spatial_shapes=d['shapes'], level_start_index=d['starts'], prev_bev=d['prev_bev'],
                  shift=d['shift'], lidar2img=l2i_dev, img_shape=img_shape)
        loss = (out.float() * d['grad_w'].float()).sum() * (1.0 / out.numel())
        loss.backward()
        return out, loss, feat.grad

    # eager pass with the layer inputs recorded
    layer_in = []
    hooks = [l.register_forward_pre_hook(lambda m, a: layer_in.append(a[0].detach().clone())) for l in enc.layers]
    out_e, loss_e, gfeat_e = step()
    for h in hooks:
        h.remove()
    # (keep no reference to the eager autograd graph: its AccumulateGrad nodes belong to the default
    # stream and must not be reused inside the capture)
    out_e, gfeat_e, loss_e = out_e.detach().clone(), gfeat_e.detach().clone(), float(loss_e.detach())
    grads_e = {n: p.grad.detach().clone() for n, p in enc.named_parameters()}

    # the same step captured and replayed as a CUDA graph (what bench.py times): identical results
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for p in params:
        p.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o, l, gf = step()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(o, out_e), 'graph replay and eager forward differ'
    assert rel_err(gf, gfeat_e) <= 2e-2          # atomics reorder from run to run (bf16 result)
    assert abs(float(l.detach()) - loss_e) <= 1e-6 * abs(loss_e) + 1e-12

    # fp32 oracle, teacher-forced on the CUDA model's layer inputs
    o32 = OracleBEVFormerEncoder(num_layers=NL, pc_range=syn.PC_RANGE, num_points_in_pillar=PILLAR,
                                 embed_dims=C, feedforward_channels=512, num_levels=len(levels), dropout=0.0)
    o32.load_state_dict({k: v.float().cpu() for k, v in enc.state_dict().items()})
    o32.eval()
    # A bf16 model decides WHERE it samples from bf16-rounded Linear outputs: an offset of 5 px is only
    # known to 0.02 px, so a few percent of the samples sit in another bilinear cell than with fp32
    # offsets, and d out / d location (piecewise constant) is then another one-sided derivative.  The
    # oracle therefore rounds the outputs of its offset / weight Linears to bf16 as well (straight-
    # through for the gradient); everything else in it stays fp32.
    def as_bf16_model(mod, inp, out):
        return out + (out.bfloat16().float() - out).detach()
    for layer in o32.layers:
        tsa, sca = layer.attentions
        for lin in (tsa.sampling_offsets, tsa.attention_weights, sca.deformable_attention.sampling_offsets,
                    sca.deformable_attention.attention_weights):
            lin.register_forward_hook(as_bf16_model)
    ins = [t.float().cpu() for t in layer_in]
    last_in = ins[-1].clone().requires_grad_(True)
    ins[-1] = last_in
    feat32 = host['feat'].float()
    outs = o32(host['bev_query'].float(), feat32, feat32, bev_h=H, bev_w=W, bev_pos=host['bev_pos'].float(),
               spatial_shapes=host['shapes'], level_start_index=host['starts'],
               prev_bev=host['prev_bev'].float(), shift=host['shift'], lidar2img=l2i,
               img_h=img_shape[0], img_w=img_shape[1], layer_inputs=ins, return_all=True)
    cuda_outs = layer_in[1:] + [out_e]
    for i, (mine, ref) in enumerate(zip(cuda_outs, outs)):
        # a layer is ~10 chained bf16 operators (3 attentions' Linears, the sampling, FFN, 3 LayerNorms):
        # 1e-2 on the tensor as a whole (2-norm); the single worst of its 10^7 entries gets 3e-2
        e2, e = rel_l2(mine.float().cpu(), ref.detach()), rel_err(mine.float().cpu(), ref.detach())
        assert e2 <= 1e-2 and e <= 3e-2, f'layer {i}: bf16 output off by l2 {e2:.3e} / max {e:.3e} from the fp32 oracle'


    # last layer's backward: d loss / d out is known in closed form, so the oracle's parameter
    # gradients of layer 5 are comparable one to one (bf16: 2e-2 of each tensor's scale, 2-norm 1e-2)
    loss = (outs[-1] * host['grad_w'].float()).sum() * (1.0 / outs[-1].numel())
    loss.backward()
    og = dict(o32.named_parameters())
    errs = {}
    for n, gcuda in grads_e.items():
        if n.startswith(f'layers.{NL - 1}.'):
            errs[n] = rel_l2(gcuda.float().cpu(), og[n].grad)
    assert len(errs) >= 10
    print('last-layer parameter gradients, rel l2 vs the fp32 oracle:',
          {k.split('.', 2)[2]: round(v, 4) for k, v in errs.items()})
    # one layer of bf16 arithmetic: 2.5e-2 of each tensor's 2-norm.  Two groups are inherently noisier
    # in a bf16 model and get their own bar: the first FFN Linear (a pre-activation within bf16
    # resolution of zero flips the ReLU mask of that element) and the offset Linears (location
    # gradients are differences of neighbouring pixels of a bf16 value map)
    def bar(name):
        if 'ffns.0.layers.0.0' in name:
            return 8e-2
        if 'sampling_offsets' in name:
            return 6e-2
        return 2.5e-2
    bad = {k: v for k, v in errs.items() if v > bar(k)}
    assert not bad, f'bf16 parameter gradients off: {bad}'
