"""Out-of-bounds checks without a sanitizer: every kernel entry point is run with

* its INPUTS embedded in NaN-filled buffers -- a load one element outside a tensor turns the
  result into NaN even when it is multiplied by a zero weight (0 * NaN = NaN), so clamped
  out-of-image corners must stay inside the tensor too;
* its OUTPUTS (and accumulators / scratch) allocated between 0xA5-filled guard bands that are
  checked afterwards, which also proves every "empty" output element is written: the body
  starts as 0xA5 bytes (a NaN-free but absurd value) and the results are compared with an
  ordinary run.

Shapes are chosen ragged (odd widths, 1-pixel maps, query counts that do not fill a tile,
locations far outside the image) to exercise the tail paths of the tiled kernels.
"""
import math

import pytest
import torch

from tests.util import make_op_inputs, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
PAD = 4096          # elements per guard band (keeps 16-byte alignment for every dtype)


class GuardedTorch:
    """Stand-in for the ``torch`` module inside the host files: allocations come with guard
    bands and are recorded; everything else is forwarded."""

    def __init__(self):
        self.records = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def _alloc(self, shape, dtype, device, zero):
        shape = tuple(int(s) for s in shape)
        n = math.prod(shape)
        buf = torch.empty(n + 2 * PAD, dtype=dtype or torch.float32, device=device)
        buf.view(torch.uint8).fill_(0xA5)
        body = buf[PAD:PAD + n]
        if zero:
            body.zero_()
        self.records.append((buf, n))
        return body.view(shape)

    @staticmethod
    def _shape(size):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            return tuple(size[0])
        return tuple(size)

    def empty(self, *size, dtype=None, device=None):
        return self._alloc(self._shape(size), dtype, device, False)

    def zeros(self, *size, dtype=None, device=None):
        return self._alloc(self._shape(size), dtype, device, True)

    def empty_like(self, x):
        return self._alloc(x.shape, x.dtype, x.device, False)

    def zeros_like(self, x):
        return self._alloc(x.shape, x.dtype, x.device, True)

    def check(self):
        torch.cuda.synchronize()
        assert self.records, 'no allocation went through the guard'
        for buf, n in self.records:
            raw = buf.view(torch.uint8)
            e = buf.element_size()
            lo, hi = raw[:PAD * e], raw[(PAD + n) * e:]
            assert bool((lo == 0xA5).all()) and bool((hi == 0xA5).all()), \
                f'guard band overwritten around a {buf.dtype} buffer of {n} elements'


@pytest.fixture
def guard(monkeypatch):
    import apollo_vision_net_b200.fused_ops as fo
    import apollo_vision_net_b200.multi_scale_deformable_attn_function as fn
    import apollo_vision_net_b200.rowops as ro
    g = GuardedTorch()
    for mod in (fo, fn, ro):
        monkeypatch.setattr(mod, 'torch', g)
    monkeypatch.setattr(fo, '_scale_ws', {})
    monkeypatch.setattr(ro, '_workspaces', {})
    return g


def poisoned(t, requires_grad=False):
    """A contiguous CUDA copy of ``t`` with NaN (float) on both sides in memory."""
    t = t.to(DEV)
    if not t.is_floating_point():
        return t.contiguous()
    buf = torch.full((t.numel() + 2 * PAD,), float('nan'), dtype=t.dtype, device=DEV)
    body = buf[PAD:PAD + t.numel()]
    body.copy_(t.reshape(-1))
    return body.view(t.shape).requires_grad_(requires_grad)


def _finite(*ts):
    for t in ts:
        assert bool(torch.isfinite(t.float()).all())


OP_CASES = [
    # B, levels, M, Dh, Nq, P, dtype
    (2, [(7, 9), (4, 5)], 8, 32, 53, 4, torch.float32),
    (2, [(7, 9), (4, 5)], 8, 32, 53, 4, torch.bfloat16),
    (1, [(1, 1), (1, 3), (2, 1)], 8, 32, 19, 2, torch.float32),      # degenerate maps
    (1, [(1, 1), (1, 3), (2, 1)], 8, 32, 19, 2, torch.float16),
    (1, [(5, 7), (3, 3)], 3, 16, 17, 3, torch.float32),
    (1, [(6, 5)], 2, 30, 9, 2, torch.float32),                       # scalar fallback kernels
    (1, [(6, 5)], 1, 71, 9, 3, torch.bfloat16),
    (3, [(12, 20), (6, 10), (3, 5), (2, 3)], 8, 32, 257, 8, torch.bfloat16),
]


@pytest.mark.parametrize('case', OP_CASES)
def test_op_boundary_stays_in_bounds(case, guard):
    import apollo_vision_net_b200 as pkg
    B, levels, M, Dh, Nq, P, dtype = case
    # locations well outside [0, 1] as well: all four corners invalid, and partially valid ones
    value, shapes, starts, loc, att = make_op_inputs(B, levels, M, Dh, Nq, P, seed=3, dtype=dtype,
                                                     lo=-0.6, hi=1.6)
    loc.view(-1, 2)[::7] = torch.tensor([-3.0, 5.0])
    loc.view(-1, 2)[3::11] = torch.tensor([1.0, 0.0])
    go = torch.randn(B, Nq, M * Dh, generator=torch.Generator().manual_seed(5)).to(dtype)
    fn = (pkg.MultiScaleDeformableAttnFunction_fp32 if dtype == torch.float32
          else pkg.MultiScaleDeformableAttnFunction_fp16)

    def run(wrap):
        v, lo_, at = wrap(value, True), wrap(loc, True), wrap(att, True)
        out = fn.apply(v, shapes.to(DEV), starts.to(DEV), lo_, at, 64)
        out.backward(wrap(go))
        return out.detach(), v.grad, lo_.grad, at.grad

    plain = run(lambda t, rg=False: t.to(DEV).requires_grad_(rg))
    n_plain = len(guard.records)
    got = run(poisoned)
    assert len(guard.records) > n_plain
    guard.check()
    _finite(*got)
    assert torch.equal(got[0], plain[0])                       # forward is deterministic
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    for a, b in zip(got[1:], plain[1:]):
        assert rel_err(a, b) <= tol


def _sca_case(bs, H, W, levels, M, Dh, P, D, dtype, seed):
    from oracle import geometry_oracle as G
    import apollo_vision_net_b200.synthetic as syn
    g = torch.Generator().manual_seed(seed)
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    l2i, img_shape = syn.camera_rig(0.5, bs=bs, jitter=4.0, seed=seed)
    r3 = G.reference_points_3d(H, W, 8.0, D, bs=bs)
    L = len(levels)
    value = torch.randn(bs * 6, Nk, M, Dh, generator=g).to(dtype)
    offsets = (torch.randn(bs, H * W, M, L, P, 2, generator=g) * 6.0).to(dtype)
    logits = torch.randn(bs, H * W, M, L * P, generator=g).to(dtype)
    go = torch.randn(bs, H * W, M * Dh, generator=g).to(dtype)
    return (value, offsets, logits, go, r3, torch.as_tensor(l2i, dtype=torch.float32), img_shape,
            torch.tensor(shapes_l), torch.tensor(starts_l))


@pytest.mark.parametrize('bs,H,W,levels,M,Dh,P,D,dtype', [
    (1, 13, 17, [(28, 48)], 8, 32, 8, 4, torch.float32),
    (2, 13, 17, [(29, 50), (15, 25), (8, 13), (4, 7)], 8, 32, 8, 4, torch.bfloat16),
    (1, 9, 31, [(7, 5), (1, 1)], 8, 16, 4, 2, torch.float32),
    (1, 21, 3, [(11, 9)], 8, 8, 4, 1, torch.bfloat16),
    (1, 10, 10, [(28, 48)], 4, 64, 8, 4, torch.float16),
])
def test_fused_sca_stays_in_bounds(bs, H, W, levels, M, Dh, P, D, dtype, guard):
    import apollo_vision_net_b200.fused_ops as fo
    import apollo_vision_net_b200.synthetic as syn
    value, offsets, logits, go, r3, l2i, img_shape, shapes, starts = _sca_case(
        bs, H, W, levels, M, Dh, P, D, dtype, seed=9)

    def run(wrap):
        geo = fo.bev_point_sampling(wrap(r3.view(bs, D, H * W, 3)), syn.PC_RANGE, wrap(l2i),
                                    img_shape[0], img_shape[1], with_lists=True)
        v, of, lg = wrap(value, True), wrap(offsets, True), wrap(logits, True)
        out = fo.SpatialCrossAttnFunction.apply(v, shapes.to(DEV), starts.to(DEV), of, lg,
                                                geo.reference_points_cam, geo.mask_u8,
                                                geo.hit_bits, 6, W)
        out.backward(wrap(go))
        return out.detach(), v.grad, of.grad, lg.grad, geo

    plain = run(lambda t, rg=False: t.to(DEV).requires_grad_(rg) if rg else t.to(DEV))
    got = run(poisoned)
    guard.check()
    _finite(*got[:4])
    assert torch.equal(got[4].mask_u8, plain[4].mask_u8)
    assert torch.equal(got[4].hit_index, plain[4].hit_index)
    assert torch.equal(got[0], plain[0])
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for a, b in zip(got[1:4], plain[1:4]):
        assert rel_err(a, b) <= tol


@pytest.mark.parametrize('bs,Q,H,W,levels,M,Dh,P,dtype', [
    (1, 2, 13, 17, [(13, 17)], 8, 32, 4, torch.float32),             # TSA
    (2, 2, 7, 29, [(7, 29)], 8, 32, 4, torch.bfloat16),
    (1, 1, 1, 57, [(9, 5), (3, 2), (1, 1)], 8, 32, 4, torch.float32),  # decoder attention
    (2, 1, 1, 33, [(50, 50)], 8, 32, 4, torch.float16),
    (1, 2, 5, 5, [(5, 5)], 4, 16, 2, torch.float32),
])
def test_fused_queue_attention_stays_in_bounds(bs, Q, H, W, levels, M, Dh, P, dtype, guard):
    import apollo_vision_net_b200.fused_ops as fo
    import apollo_vision_net_b200.synthetic as syn
    g = torch.Generator().manual_seed(21)
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    L, Nq = len(levels), H * W
    value = torch.randn(bs * Q, Nk, M, Dh, generator=g).to(dtype)
    offsets = (torch.randn(bs, Nq, M, Q, L, P, 2, generator=g) * 4.0).to(dtype)
    logits = (torch.randn(bs, Nq, M, Q, L * P, generator=g) * 3.0).to(dtype)
    ref = torch.rand(bs * Q, Nq, L, 2, generator=g) * 1.4 - 0.2
    go = torch.randn(bs, Nq, M * Dh, generator=g).to(dtype)
    shapes, starts = torch.tensor(shapes_l).to(DEV), torch.tensor(starts_l).to(DEV)

    def run(wrap):
        v, of, lg = wrap(value, True), wrap(offsets, True), wrap(logits, True)
        out = fo.QueueDeformAttnFunction.apply(v, shapes, starts, of, lg, wrap(ref), 5.0,
                                               W if Q == 2 else 0)
        out.backward(wrap(go))
        return out.detach(), v.grad, of.grad, lg.grad

    plain = run(lambda t, rg=False: t.to(DEV).requires_grad_(rg) if rg else t.to(DEV))
    got = run(poisoned)
    guard.check()
    _finite(*got)
    assert torch.equal(got[0], plain[0])
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for a, b in zip(got[1:], plain[1:]):
        assert rel_err(a, b) <= tol


@pytest.mark.parametrize('rows,C,dtype', [(1, 256, torch.float32), (777, 256, torch.bfloat16),
                                          (4099, 512, torch.float32), (33, 256, torch.float16)])
def test_row_kernels_stay_in_bounds(rows, C, dtype, guard):
    import apollo_vision_net_b200.rowops as ro
    g = torch.Generator().manual_seed(2)
    x = torch.randn(rows, C, generator=g).to(dtype)
    gam = (torch.rand(C, generator=g) + 0.5).to(dtype)
    bet = torch.randn(C, generator=g).to(dtype)
    go = torch.randn(rows, C, generator=g).to(dtype)
    x1, g1, b1 = poisoned(x, True), poisoned(gam, True), poisoned(bet, True)
    y = ro.LayerNormFunction.apply(x1, g1, b1, 1e-5)
    y.backward(poisoned(go))
    s = ro.column_sum(poisoned(go))
    guard.check()
    _finite(y, x1.grad, g1.grad, b1.grad, s)
    x2 = x.to(DEV).float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(x2, (C,), gam.to(DEV).float(), bet.to(DEV).float(), 1e-5)
    ref.backward(go.to(DEV).float())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel_err(y, ref) <= tol
    assert rel_err(x1.grad, x2.grad) <= tol
    assert rel_err(s, go.to(DEV).float().sum(0)) <= tol


# ---- round 2 entry points: mha_fwd / mha_bwd (csrc/mha.cu) and dcnv3_fwd / dcnv3_bwd (csrc/dcnv3.cu) --------------
@pytest.fixture
def guard2(monkeypatch):
    import apollo_vision_net_b200.dcnv3 as dc
    import apollo_vision_net_b200.mha as mh
    g = GuardedTorch()
    for mod in (mh, dc):
        monkeypatch.setattr(mod, 'torch', g)
    monkeypatch.setattr(mh, '_mask_cache', {})
    return g


MHA_CASES = [
    # layout kind, a, b, c, heads, head_dim, dtype, impl, masked, dropout
    ('seq', 37, 3, 0, 4, 32, torch.bfloat16, 2, True, 0.0),          # ragged: last tile 5 rows, last key block 5 keys
    ('seq', 37, 3, 0, 4, 32, torch.float32, 1, True, 0.0),
    ('intra', 11, 20, 2, 8, 32, torch.float16, 2, False, 0.1),       # several problems per CTA, last CTA half full
    ('inter', 50, 3, 1, 8, 32, torch.bfloat16, 2, True, 0.1),
    ('batch', 2, 45, 0, 2, 64, torch.bfloat16, 2, True, 0.0),
    ('seq', 1, 5, 0, 8, 32, torch.bfloat16, 2, False, 0.0),          # one token
    ('seq', 33, 2, 0, 2, 16, torch.float32, 1, False, 0.1),          # FMA path, head_dim 16
    ('seq', 300, 1, 0, 8, 32, torch.bfloat16, 2, False, 0.0),        # few problems: half-filled CTAs
]


@pytest.mark.parametrize('case', MHA_CASES)
def test_self_attention_core_stays_in_bounds(case, guard2):
    import apollo_vision_net_b200.mha as m
    import apollo_vision_net_b200.rowops as ro
    kind, a, b, c, H, Dh, dtype, impl, masked, p = case
    lay = {'seq': lambda: m.sequence_first_layout(a, b), 'batch': lambda: m.batch_first_layout(a, b),
           'inter': lambda: m.inter_vector_layout(a, b, c), 'intra': lambda: m.intra_vector_layout(a, b, c)}[kind]()
    C = H * Dh
    rows = lay.G * lay.S
    g = torch.Generator().manual_seed(7)
    qk = torch.randn(rows, 2 * C, generator=g).to(dtype)
    v = torch.randn(rows, C, generator=g).to(dtype)
    go = torch.randn(rows, C, generator=g).to(dtype)
    mask = None
    if masked:
        mask = torch.zeros(lay.S, lay.S, dtype=torch.bool)
        mask[lay.S // 3:, :lay.S // 3] = True
        mask[:lay.S // 3, lay.S // 3:] = True
        mask = mask.to(DEV)

    def run(wrap):
        ro.reseed_dropout(DEV, 1234)                    # the same masks in both runs
        ro._drop_site[0] = 100
        q_, v_ = wrap(qk, True), wrap(v, True)
        out = m.self_attention_core(q_, v_, lay, H, mask, p, impl)
        out.backward(wrap(go))
        return out.detach(), q_.grad, v_.grad

    plain = run(lambda t, rg=False: t.to(DEV).requires_grad_(rg))
    n_plain = len(guard2.records)
    got = run(poisoned)
    assert len(guard2.records) > n_plain
    guard2.check()
    _finite(*got)
    for x, y in zip(got, plain):
        assert torch.equal(x, y)                         # deterministic kernels: bit-identical to the ordinary run


DCN_CASES = [
    # N, H, W, group, group_channels, kh, kw, stride, pad, dilation, offset_scale, dtype
    (2, 7, 9, 4, 8, 3, 3, 1, 1, 1, 1.0, torch.float32),
    (2, 7, 9, 4, 8, 3, 3, 1, 1, 1, 1.0, torch.bfloat16),
    (1, 1, 1, 2, 16, 3, 3, 1, 1, 1, 1.0, torch.float16),             # one-pixel map: every neighbour outside
    (1, 9, 6, 1, 6, 3, 3, 2, 2, 2, 3.0, torch.float32),              # scalar fallback kernels, far-away positions
    (1, 5, 8, 3, 8, 1, 5, 1, 0, 1, 0.5, torch.bfloat16),
]


@pytest.mark.parametrize('case', DCN_CASES)
def test_dcnv3_stays_in_bounds(case, guard2):
    import apollo_vision_net_b200.dcnv3 as d
    from oracle.dcnv3_oracle import output_size
    N, H, W, G, Cg, kh, kw, st, pad, dil, scale, dtype = case
    Ho, Wo = output_size(H, kh, st, pad, dil), output_size(W, kw, st, pad, dil)
    K = kh * kw
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, H, W, G * Cg, generator=g).to(dtype)
    off = (torch.randn(N, Ho, Wo, G * K * 2, generator=g) * 4).to(dtype)          # many positions outside the map
    msk = torch.softmax(torch.randn(N, Ho, Wo, G, K, generator=g), -1).reshape(N, Ho, Wo, G * K).to(dtype)
    go = torch.randn(N, Ho, Wo, G * Cg, generator=g).to(dtype)
    cfg = (kh, kw, st, st, pad, pad, dil, dil, G, Cg, scale)

    def run(wrap):
        x_, o_, m_ = wrap(x, True), wrap(off, True), wrap(msk, True)
        out = d.DCNv3Function.apply(x_, o_, m_, *cfg, 256)
        out.backward(wrap(go))
        return out.detach(), x_.grad, o_.grad, m_.grad

    plain = run(lambda t, rg=False: t.to(DEV).requires_grad_(rg))
    n_plain = len(guard2.records)
    got = run(poisoned)
    assert len(guard2.records) > n_plain
    guard2.check()
    _finite(*got)
    for a_, b_ in zip(got[:1] + got[2:], plain[:1] + plain[2:]):
        assert torch.equal(a_, b_)
    assert rel_err(got[1], plain[1]) <= (1e-6 if dtype == torch.float32 else 1e-2)   # grad_input: atomic order
