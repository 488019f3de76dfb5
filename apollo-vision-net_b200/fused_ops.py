"""Python host side of the fused B200 kernels (C ABI: bev_point_sampling, sca_*, tsa_*).

These replace the tensor plumbing between the Linear layers of the reference's attention
modules (no rebatch, no materialised sampling_locations, no scatter, no host sync):

* :func:`bev_point_sampling` -- ``BEVFormerEncoder.point_sampling`` (encoder.py:89-241) plus
  the per-camera ``nonzero()`` compaction of ``SpatialCrossAttention.forward``
  (spatial_cross_attention.py:135-139).
* :class:`SpatialCrossAttnFunction` -- spatial_cross_attention.py:135-170 and :342-396.
* :class:`QueueDeformAttnFunction` -- temporal_self_attention.py:204-279 (queue of 2) and
  decoder.py:299-350 (queue of 1).
"""
import ctypes

import torch
from torch.autograd.function import Function, once_differentiable

from . import _lib, rowops
from .multi_scale_deformable_attn_function import (_DTYPE_CODE, _require_cuda, _stream_ptr,
                                                   custom_bwd, custom_fwd)


class BevGeometry:
    """Result of :func:`bev_point_sampling`; everything stays on the device."""

    def __init__(self, ref_cam, mask_u8, hit_bits, hit_index, hit_count, D):
        self.reference_points_cam = ref_cam          # (num_cam, bs, HW, D, 2) fp32
        self.mask_u8 = mask_u8                       # (num_cam, bs, HW, D) uint8
        self.hit_bits = hit_bits                     # (bs, HW) int32 bit field over cameras
        self.hit_index = hit_index                   # (num_cam, HW) int32, -1 padded
        self.hit_count = hit_count                   # (num_cam,) int32
        self.D = D

    @property
    def bev_mask(self):
        """(num_cam, bs, HW, D) bool, the reference's ``bev_mask``."""
        return self.mask_u8.view(torch.bool)

    def lists(self):
        """(hit_index, hit_count) of THIS geometry's queries, derived once from the bit field when
        :func:`bev_point_sampling` was called without lists (or the geometry is a row slice)."""
        if self.hit_index is None:
            self.hit_index, self.hit_count = hit_lists(self.hit_bits, self.reference_points_cam.shape[0])
        return self.hit_index, self.hit_count

    def rows(self, q0, q1):
        """Geometry of the query slice [q0, q1) (BEV row sharding).  The camera gating of the
        reference reads batch element 0 (quirk 1), so the slice keeps that convention; the
        compacted hit lists of the slice are rebuilt on demand (:meth:`lists`)."""
        return BevGeometry(self.reference_points_cam[:, :, q0:q1].contiguous(),
                           self.mask_u8[:, :, q0:q1].contiguous(),
                           self.hit_bits[:, q0:q1].contiguous(), None, None, self.D)


def bev_point_sampling(ref_3d, pc_range, lidar2img, img_h, img_w, with_lists=True):
    """ref_3d (bs, D, HW, 3) fp32 CUDA; lidar2img (bs, num_cam, 4, 4) -> :class:`BevGeometry`.
    ``with_lists=False`` skips the ordered per-camera hit lists (``hit_index`` / ``hit_count``
    are then None); the fused spatial cross-attention only needs the bit field."""
    _require_cuda(ref_3d=ref_3d)
    dev = ref_3d.device
    ref_3d = ref_3d.to(torch.float32).contiguous()
    l2i = torch.as_tensor(lidar2img, dtype=torch.float32).to(dev).contiguous()
    bs, D, HW, _ = ref_3d.shape
    num_cam = l2i.shape[1]
    assert l2i.shape == (bs, num_cam, 4, 4), l2i.shape
    ref_cam = torch.empty((num_cam, bs, HW, D, 2), dtype=torch.float32, device=dev)
    mask = torch.empty((num_cam, bs, HW, D), dtype=torch.uint8, device=dev)
    hit_bits = torch.empty((bs, HW), dtype=torch.int32, device=dev)
    hit_index = torch.empty((num_cam, HW), dtype=torch.int32, device=dev) if with_lists else None
    hit_count = torch.empty((num_cam,), dtype=torch.int32, device=dev) if with_lists else None
    pc = (ctypes.c_double * 6)(*[float(x) for x in pc_range])
    with torch.cuda.device(dev):
        _lib.call('bev_point_sampling', ref_3d.data_ptr(), l2i.data_ptr(), ctypes.cast(pc, ctypes.c_void_p), float(img_h),
            float(img_w), bs, num_cam, HW, D, ref_cam.data_ptr(), mask.data_ptr(),
            hit_bits.data_ptr(), hit_index.data_ptr() if with_lists else None,
            hit_count.data_ptr() if with_lists else None, _stream_ptr(ref_3d))
    return BevGeometry(ref_cam, mask, hit_bits, hit_index, hit_count, D)


def hit_bits_from_mask(bev_mask):
    """(num_cam, bs, HW, D) bool -> (bs, HW) int32 camera bit field (for callers that bring
    their own ``bev_mask`` instead of calling :func:`bev_point_sampling`)."""
    num_cam = bev_mask.shape[0]
    any_d = bev_mask.any(-1).to(torch.int32)                       # (num_cam, bs, HW)
    w = (2 ** torch.arange(num_cam, device=bev_mask.device, dtype=torch.int64)).view(-1, 1, 1)
    bits = (any_d.to(torch.int64) * w).sum(0)
    bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits)
    return bits.to(torch.int32).contiguous()


def _check_value(value):
    if value.dtype not in _DTYPE_CODE:
        raise RuntimeError(f'unsupported value dtype {value.dtype}')
    return value.contiguous()


_scale_ws = {}
_overflow_flags = {}

# fp16 accumulator: replicas of the coarse tail of every value map (see sca_bwd in msda_b200.h).
# The coarse pyramid levels sit at the end of the pixel axis and hold ~1/16 of the pixels but half of
# the samples; 8-way replication cuts the rounding noise of their slots by ~sqrt(8).
_TAIL_FRACTION = 16
_TAIL_COPIES = 7

_accum_mode = ['fp32' if __import__('os').environ.get('APOLLO_B200_FP32_ACCUM', '0') == '1' else 'fp16']


def set_grad_accumulator(mode):
    """Accumulator of grad_value in the fused backward kernels for 16-bit value dtypes: ``'fp16'``
    (default: scaled fp16, half the L2 reduction traffic, overflow excluded by the scale bound of
    :func:`_accumulator`) or ``'fp32'`` (the reference's arithmetic: fp32 atomics).  fp32 models always
    use fp32.  Returns the previous mode.  (``APOLLO_B200_FP32_ACCUM=1`` selects 'fp32' at import.)"""
    if mode not in ('fp16', 'fp32'):
        raise ValueError(f"grad accumulator mode must be 'fp16' or 'fp32', got {mode!r}")
    prev, _accum_mode[0] = _accum_mode[0], mode
    return prev


class grad_accumulator:
    """``with grad_accumulator('fp32'): ...`` -- scoped :func:`set_grad_accumulator`."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = set_grad_accumulator(self.mode)

    def __exit__(self, *exc):
        set_grad_accumulator(self.prev)


def _overflow_flag(dev):
    f = _overflow_flags.get(dev.index)
    if f is None:
        f = torch.zeros(1, dtype=torch.int32, device=dev)
        _overflow_flags[dev.index] = f
    return f


def grad_accumulator_overflowed(device=None, reset=True):
    """True when an fp16 accumulator pass on ``device`` met a non-finite slot since the last reset
    (sticky device flag set by ``unscale_cast``).  Reading it synchronises; the training path never
    does.  The scale bound makes this unreachable for finite inputs -- it is the safety net."""
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    f = _overflow_flags.get(dev.index)
    if f is None:
        return False
    hit = bool(int(f.item()))
    if reset:
        f.zero_()
    return hit


_coarse_mode = [__import__('os').environ.get('APOLLO_B200_COARSE', '1') != '0']
_coarse_ws = {}


def set_coarse_tensor_core_pass(on):
    """Spatial cross-attention backward, 16-bit value dtypes with head_dim 32: accumulate grad_value of
    the coarse pyramid levels on the tensor cores (records + tcgen05 contraction, see ``sca_bwd`` in
    include/msda_b200.h) instead of one L2 reduction per corner.  Default on; returns the previous
    setting.  (``APOLLO_B200_COARSE=0`` turns it off at import.)"""
    prev, _coarse_mode[0] = _coarse_mode[0], bool(on)
    return prev


def hit_lists(hit_bits, num_cam):
    """(bs, HW) int32 camera bit field -> (hit_index (num_cam, HW) int32, hit_count (num_cam,) int32):
    the per-camera ordered query lists of batch element 0 (what :func:`bev_point_sampling` returns with
    ``with_lists=True``), for callers that only hold the bit field."""
    _require_cuda(hit_bits=hit_bits)
    hit_bits = hit_bits.contiguous()
    HW = hit_bits.shape[-1]
    idx = torch.empty((num_cam, HW), dtype=torch.int32, device=hit_bits.device)
    cnt = torch.empty((num_cam,), dtype=torch.int32, device=hit_bits.device)
    with torch.cuda.device(hit_bits.device):
        _lib.call('bev_hit_lists', hit_bits.data_ptr(), num_cam, HW, idx.data_ptr(), cnt.data_ptr(),
                  _stream_ptr(hit_bits))
    return idx, cnt


def _coarse_records(value, bs, num_cam, HW, M, Dh, P):
    """Record workspace of the tensor-core coarse pass (one buffer per device and stream, grown on
    demand, contents undefined between calls), or None when the pass does not apply."""
    if not _coarse_mode[0] or value.dtype == torch.float32:
        return None
    need = int(_lib.lib().sca_coarse_workspace_bytes(bs, num_cam, HW, M, Dh, P, _DTYPE_CODE[value.dtype]))
    if need <= 0:
        return None
    dev = value.device
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _coarse_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _coarse_ws[key] = ws
    return ws


def _accumulator(value, g_out, num_levels=1, rows_per_slot=0, replicas=True):
    """grad_value accumulator for the fused backward kernels.

    fp32 value: fp32 accumulator (red.global.add.v4.f32).  16-bit value: fp16 accumulator scaled by
    a power of two derived on the device from max|g_out| (no host sync) -- half the L2 sectors per
    update; :func:`set_grad_accumulator` selects fp32 instead.  ``rows_per_slot`` = how many
    (query, head) rows can add into one slot of a value map (the number of queries): a row's attention
    weights sum to one and bilinear weights are <= 1, so with a single scaled contribution bounded by
    ``min(4, 32768 / rows_per_slot)`` no slot can exceed 32768 -- fp16 cannot overflow whatever the
    signs and locations.  Multi-level value maps additionally get replicas of their coarse tail.
    Returns (buffer, code, scale_tensor, tail, tail_pixels).
    """
    import os
    dev = value.device
    if value.dtype == torch.float32 or _accum_mode[0] == 'fp32':
        return torch.zeros(value.shape, dtype=torch.float32, device=dev), _lib.F32, None, None, 0
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _scale_ws.get(key)
    if ws is None:
        ws = torch.zeros(64, dtype=torch.float32, device=dev)
        _scale_ws[key] = ws
    limit = min(4.0, 32768.0 / max(1, int(rows_per_slot)))
    # the launch that finds max|g_out| also clears the accumulator (no fill launch of its own)
    acc = torch.empty(value.shape, dtype=torch.float16, device=dev)
    fused_zero = (acc.numel() * 2) % 16 == 0 and acc.data_ptr() % 16 == 0
    with torch.cuda.device(dev):
        if fused_zero:
            _lib.call('grad_amax_scale_zero', g_out.data_ptr(), g_out.numel(), _DTYPE_CODE[g_out.dtype],
                      float(limit), ws.data_ptr(), acc.data_ptr(), acc.numel() * 2, _stream_ptr(value))
        else:
            acc.zero_()
            _lib.call('grad_amax_scale', g_out.data_ptr(), g_out.numel(), _DTYPE_CODE[g_out.dtype],
                      float(limit), ws.data_ptr(), _stream_ptr(value))
    maps, Nk, M, Dh = value.shape
    tail, tail_px = None, 0
    if replicas and num_levels > 1 and Nk >= 4 * _TAIL_FRACTION and (M * Dh) % 8 == 0 and \
            os.environ.get('APOLLO_B200_TAIL_REPLICAS', '1') != '0':
        tail_px = Nk // _TAIL_FRACTION
        tail = torch.zeros((_TAIL_COPIES, maps, tail_px, M, Dh), dtype=torch.float16, device=dev)
    return acc, _lib.F16, ws[16:17], tail, tail_px


def _finish_accumulator(acc, code, scale, value, tail=None, tail_px=0):
    if code == _lib.F32:
        return acc.to(value.dtype)
    maps, Nk, M, Dh = value.shape
    C = M * Dh
    # a value tensor that is one of several hoisted projections of the same input gets its gradient written
    # as a column block of their common matrix (rowops.GradSlots)
    block = rowops.grad_slot(value) if (tail is None and C % 8 == 0 and acc.numel() % C == 0) else None
    if block is not None:
        out, out_ld = block.view(maps, Nk, M, Dh), block.stride(0)
    else:
        out, out_ld = torch.empty(value.shape, dtype=value.dtype, device=value.device), 0
    # the same pass returns the column sums of the (maps * Nk, C) view: the bias gradient of the value
    # projection that produced `value` (rowops.offer_bias_grad hands it to that Linear's backward)
    with_sums = C % 8 == 0 and 256 % (C // 8) == 0 and acc.numel() % 8 == 0
    sums = torch.empty(C, dtype=value.dtype, device=value.device) if with_sums else None
    ws = rowops._workspace(value.device, C) if with_sums else None
    with torch.cuda.device(value.device):
        _lib.call('unscale_cast_strided', acc.data_ptr(), out.data_ptr(), scale.data_ptr(), acc.numel(),
                  _DTYPE_CODE[value.dtype], None if tail is None else tail.data_ptr(),
                  0 if tail is None else tail.shape[0], Nk * M * Dh, tail_px * M * Dh,
                  None if sums is None else sums.data_ptr(), None if ws is None else ws.data_ptr(), C,
                  _overflow_flag(value.device).data_ptr(), int(out_ld), _stream_ptr(value))
    if with_sums:
        rowops.offer_bias_grad(out, sums)
    return out


class _Coords:
    """Offsets / logits as the kernels take them: base pointers, per-query row strides (elements)
    and dtype.  They are consumed in fp32 or in the value dtype (no conversion pass for a bf16
    model); anything else is brought to fp32.

    Two calling conventions: separate tensors (the outputs of the ``sampling_offsets`` and
    ``attention_weights`` Linear layers), or -- ``logits is None`` -- ONE tensor ``(bs, Nq, 3*n)``
    whose columns ``[0, 2n)`` are the offsets and ``[2n, 3n)`` the logits of a query, i.e. the
    output of a single GEMM over the concatenated weights; the
    gradient then comes back as one tensor of the same layout.
    """

    def __init__(self, value, offsets, logits, tail_off, tail_log):
        self.merged = logits is None
        if self.merged:
            dt = offsets.dtype if offsets.dtype in (torch.float32, value.dtype) else torch.float32
            buf = offsets.to(dt).contiguous()
            bs, nq, width = buf.shape
            n = width // 3
            assert width == 3 * n and 2 * n == _prod(tail_off) and n == _prod(tail_log), \
                (buf.shape, tail_off, tail_log)
            self.keep = (buf,)
            self.off_ptr, self.log_ptr = buf.data_ptr(), buf.data_ptr() + 2 * n * buf.element_size()
            self.off_stride = self.log_stride = width
            self.n = n
        else:
            dt = offsets.dtype if (offsets.dtype == logits.dtype and
                                   offsets.dtype in (torch.float32, value.dtype)) else torch.float32
            o, lg = offsets.to(dt).contiguous(), logits.to(dt).contiguous()
            assert tuple(o.shape[2:]) == tuple(tail_off) and tuple(lg.shape[2:]) == tuple(tail_log), \
                (o.shape, lg.shape, tail_off, tail_log)
            self.keep = (o, lg)
            self.off_ptr, self.log_ptr = o.data_ptr(), lg.data_ptr()
            self.off_stride, self.log_stride = _prod(tail_off), _prod(tail_log)
        self.dtype = dt
        self.code = _DTYPE_CODE[dt]

    @staticmethod
    def grads(saved, merged):
        """Gradient buffers with the layout of the inputs: (tensors to return, off_ptr, log_ptr)."""
        if merged:
            buf, = saved
            g = torch.empty(buf.shape, dtype=buf.dtype, device=buf.device)
            n = buf.shape[-1] // 3
            return (g, None), g.data_ptr(), g.data_ptr() + 2 * n * g.element_size()
        o, lg = saved
        g_off = torch.empty(o.shape, dtype=o.dtype, device=o.device)
        g_log = torch.empty(lg.shape, dtype=lg.dtype, device=lg.device)
        return (g_off, g_log), g_off.data_ptr(), g_log.data_ptr()


def _prod(xs):
    n = 1
    for x in xs:
        n *= int(x)
    return n


class SpatialCrossAttnFunction(Function):
    """slots = (sum over hit cameras of MSDA(value_cam, ref_cam + offsets / (W, H), softmax(logits))) / count.

    value (bs*num_cam, Nk, M, Dh); offsets (bs, HW, M, L, P, 2) and logits (bs, HW, M, L*P)
    are the raw outputs of the ``sampling_offsets`` / ``attention_weights`` Linear layers on the
    BEV queries (computed once per query, not once per (camera, query) pair).
    """

    @staticmethod
    @custom_fwd(cast_inputs=None)
    def forward(ctx, value, spatial_shapes, level_start_index, offsets, logits, ref_cam,
                mask_u8, hit_bits, num_cam, bev_w=0, hit_lists_=None):
        _require_cuda(value=value, offsets=offsets, logits=logits, ref_cam=ref_cam,
                      mask=mask_u8, hit_bits=hit_bits)
        value = _check_value(value)
        shapes = spatial_shapes.to(torch.int64).contiguous()
        starts = level_start_index.to(torch.int64).contiguous()
        ref_cam = ref_cam.to(torch.float32).contiguous()
        mask_u8 = mask_u8.contiguous()
        hit_bits = hit_bits.contiguous()
        Bc, Nk, M, Dh = value.shape
        bs, HW = offsets.shape[:2]
        L = shapes.shape[0]
        D = ref_cam.shape[3]
        P = (offsets.shape[-1] // (3 * M * L)) if logits is None else offsets.shape[4]
        co = _Coords(value, offsets, logits, (M, L, P, 2), (M, L * P))
        assert Bc == bs * num_cam
        assert ref_cam.shape == (num_cam, bs, HW, D, 2) and hit_bits.shape == (bs, HW)
        slots = torch.empty((bs, HW, M * Dh), dtype=value.dtype, device=value.device)
        with torch.cuda.device(value.device):
            _lib.call('sca_fwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), co.off_ptr,
                co.log_ptr, ref_cam.data_ptr(), mask_u8.data_ptr(), hit_bits.data_ptr(),
                slots.data_ptr(), None, bs, num_cam, Nk, M, Dh, L, P, D, HW, int(bev_w or 0),
                _DTYPE_CODE[value.dtype], co.code, co.off_stride, co.log_stride, _stream_ptr(value))
        ctx.save_for_backward(value, shapes, starts, ref_cam, mask_u8, hit_bits, *co.keep)
        ctx.merged, ctx.dims = co.merged, (bs, HW, M, Dh, L, P, D, Nk)
        ctx.coord = (co.code, co.off_stride, co.log_stride)
        ctx.num_cam = num_cam
        # per-camera hit lists for the backward's tensor-core pass (non-differentiable int32 tensors
        # from bev_point_sampling; derived from the bit field when the caller did not bring them)
        ctx.hit_lists = hit_lists_
        ctx.bev_w = int(bev_w or 0)
        return slots

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, g_slots):
        value, shapes, starts, ref_cam, mask_u8, hit_bits, *coords = ctx.saved_tensors
        num_cam = ctx.num_cam
        bs, HW, M, Dh, L, P, D, Nk = ctx.dims
        code, so, sl = ctx.coord
        off_ptr = coords[0].data_ptr()
        log_ptr = off_ptr + 2 * (so // 3) * coords[0].element_size() if ctx.merged else coords[1].data_ptr()
        g_slots = g_slots.to(value.dtype).contiguous()
        records = _coarse_records(value, bs, num_cam, HW, M, Dh, P)
        lists = None
        if records is not None:
            lists = ctx.hit_lists
            if lists is None or lists[0] is None:
                lists = hit_lists(hit_bits, num_cam)
            assert lists[0].shape == (num_cam, HW) and lists[0].dtype == torch.int32 and lists[0].is_contiguous()
        # (with the tensor-core pass the coarse levels need no replicas of the accumulator tail, and the
        # head_dim-32 kernels do not implement them)
        g_value, acc_code, acc_scale, tail, tail_px = _accumulator(value, g_slots, num_levels=L, rows_per_slot=HW,
                                                                   replicas=records is None and Dh != 32)
        (g_off, g_log), g_off_ptr, g_log_ptr = _Coords.grads(coords, ctx.merged)
        with torch.cuda.device(value.device):
            _lib.call('sca_bwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), off_ptr,
                log_ptr, ref_cam.data_ptr(), mask_u8.data_ptr(), hit_bits.data_ptr(),
                g_slots.data_ptr(), g_value.data_ptr(), g_off_ptr, g_log_ptr,
                bs, num_cam, Nk, M, Dh, L, P, D, HW, ctx.bev_w, _DTYPE_CODE[value.dtype],
                code, so, sl, acc_code,
                None if acc_scale is None else acc_scale.data_ptr(),
                None if tail is None else tail.data_ptr(), 0 if tail is None else tail.shape[0],
                tail_px, None if records is None else records.data_ptr(),
                None if records is None else lists[0].data_ptr(),
                None if records is None else lists[1].data_ptr(), _stream_ptr(value))
        g_value = _finish_accumulator(g_value, acc_code, acc_scale, value, tail, tail_px)
        return (g_value, None, None, g_off, g_log, None, None, None, None, None, None)


class QueueDeformAttnFunction(Function):
    """out = mean over the queue of MSDA(value[b*Q+j], ref[b*Q+j] + offsets_j / (W, H), softmax(logits_j)).

    value (bs*Q, Nk, M, Dh); offsets (bs, Nq, M, Q, L, P, 2); logits (bs, Nq, M, Q, L*P);
    ref (bs*Q, Nq, L, 2).  Q = 2 is TemporalSelfAttention, Q = 1 CustomMSDeformableAttention.
    """

    @staticmethod
    @custom_fwd(cast_inputs=None)
    def forward(ctx, value, spatial_shapes, level_start_index, offsets, logits, ref, clamp,
                bev_w=0):
        _require_cuda(value=value, offsets=offsets, logits=logits, ref=ref)
        value = _check_value(value)
        ref = ref.to(torch.float32).contiguous()
        shapes = spatial_shapes.to(torch.int64).contiguous()
        starts = level_start_index.to(torch.int64).contiguous()
        BQ, Nk, M, Dh = value.shape
        bs, Nq = offsets.shape[:2]
        L = shapes.shape[0]
        Q = BQ // bs
        P = (offsets.shape[-1] // (3 * M * Q * L)) if logits is None else offsets.shape[5]
        co = _Coords(value, offsets, logits, (M, Q, L, P, 2), (M, Q, L * P))
        assert BQ == bs * Q
        assert ref.shape == (bs * Q, Nq, L, 2), (ref.shape, (bs * Q, Nq, L, 2))
        clamp = -1.0 if clamp is None else float(clamp)
        out = torch.empty((bs, Nq, M * Dh), dtype=value.dtype, device=value.device)
        with torch.cuda.device(value.device):
            _lib.call('tsa_fwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), co.off_ptr,
                co.log_ptr, ref.data_ptr(), out.data_ptr(), bs, Q, Nk, M, Dh, L, P, Nq,
                int(bev_w or 0), clamp, _DTYPE_CODE[value.dtype], co.code, co.off_stride,
                co.log_stride, _stream_ptr(value))
        ctx.save_for_backward(value, shapes, starts, ref, *co.keep)
        ctx.merged, ctx.dims = co.merged, (bs, Nq, M, Dh, Q, L, P, Nk)
        ctx.coord = (co.code, co.off_stride, co.log_stride)
        ctx.clamp = clamp
        ctx.bev_w = int(bev_w or 0)
        return out

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, g_out):
        value, shapes, starts, ref, *coords = ctx.saved_tensors
        bs, Nq, M, Dh, Q, L, P, Nk = ctx.dims
        code, so, sl = ctx.coord
        off_ptr = coords[0].data_ptr()
        log_ptr = off_ptr + 2 * (so // 3) * coords[0].element_size() if ctx.merged else coords[1].data_ptr()
        g_out = g_out.to(value.dtype).contiguous()
        g_value, acc_code, acc_scale, _, _ = _accumulator(value, g_out, rows_per_slot=Nq)
        (g_off, g_log), g_off_ptr, g_log_ptr = _Coords.grads(coords, ctx.merged)
        with torch.cuda.device(value.device):
            _lib.call('tsa_bwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), off_ptr,
                log_ptr, ref.data_ptr(), g_out.data_ptr(), g_value.data_ptr(),
                g_off_ptr, g_log_ptr, bs, Q, Nk, M, Dh, L, P, Nq, ctx.bev_w, ctx.clamp,
                _DTYPE_CODE[value.dtype], code, so, sl, acc_code,
                None if acc_scale is None else acc_scale.data_ptr(), _stream_ptr(value))
        g_value = _finish_accumulator(g_value, acc_code, acc_scale, value)
        g_ref = None
        if ctx.needs_input_grad[5]:
            # loc = ref + off / (W, H)  =>  d ref = sum over heads and points of d off * (W, H)
            wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).to(torch.float32)      # (L, 2)
            go = g_off[..., :2 * (so // 3)] if ctx.merged else g_off
            go = go.reshape(bs, Nq, M, Q, L, P, 2)
            g_ref = (go.float() * wh.view(1, 1, 1, 1, L, 1, 2)).sum(dim=(2, 5))            # (bs, Nq, Q, L, 2)
            g_ref = g_ref.permute(0, 2, 1, 3, 4).reshape(bs * Q, Nq, L, 2)
        return g_value, None, None, g_off, g_log, g_ref, None, None
