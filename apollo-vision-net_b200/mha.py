"""Small-sequence multi-head self-attention core on the sm_100a kernels of ``csrc/mha.cu``.

Replaces the ``torch.nn.MultiheadAttention`` core that mmcv's ``MultiheadAttention`` wraps for the
decoders' dense self-attentions (``projects/mmdet3d_plugin/maptrv2/modules/decoder.py:129-188``: the
inter-vector attention with the one-to-one / one-to-many mask and the intra-vector attention; mmdet's
``DetrTransformerDecoderLayer`` under ``bevformer/modules/decoder.py:50-126``).  The tokens are attended
IN PLACE in the layer's ``(num_query, bs, C)`` activations: a :class:`TokenLayout` says which rows form a
sequence, so neither grouping of the reference needs its permute + contiguous copies.

There is no CPU path: CPU tensors raise, a missing library raises (``_lib.lib()``).
"""
from collections import namedtuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from .multi_scale_deformable_attn_function import _DTYPE_CODE, _stream_ptr, custom_bwd, custom_fwd
from .rowops import _colsum_supported, _next_site, column_sum, dropout_state

# Token s of group g is row  s * seq_stride + (g // n_lo) * hi_stride + (g % n_lo) * lo_stride  of the
# (rows, C) activation matrix; G groups of S tokens cover every row exactly once.
TokenLayout = namedtuple('TokenLayout', 'G S seq_stride hi_stride lo_stride n_lo')

IMPL_AUTO, IMPL_FMA, IMPL_TENSOR_CORE = 0, 1, 2


def sequence_first_layout(S, B):
    """(S, B, C) activations, one sequence per batch column (mmcv's default, ``batch_first=False``)."""
    return TokenLayout(B, S, B, 0, 1, B)


def batch_first_layout(B, S):
    """(B, S, C) activations."""
    return TokenLayout(B, S, 1, 0, S, B)


def inter_vector_layout(num_vec, num_pts, bs):
    """(num_vec * num_pts, bs, C): the sequence runs over the vectors, one group per (point, sample)
    (reference decoder.py:131-148)."""
    return TokenLayout(num_pts * bs, num_vec, num_pts * bs, 0, 1, num_pts * bs)


def intra_vector_layout(num_vec, num_pts, bs):
    """(num_vec * num_pts, bs, C): the sequence runs over a vector's points, one group per (vector, sample)
    (reference decoder.py:149-185, without its permute + contiguous copies)."""
    return TokenLayout(num_vec * bs, num_pts, bs, num_pts * bs, 1, bs)


def supported_impl(layout, heads, head_dim, dtype):
    """0 = the kernels do not cover this case, 1 = FMA path, 2 = tensor-core path."""
    if dtype not in _DTYPE_CODE:
        return 0
    return int(_lib.lib().mha_impl(int(layout.G), int(heads), int(layout.S), int(head_dim), _DTYPE_CODE[dtype]))


_mask_cache = {}


def pack_mask(mask):
    """Boolean (S, S) ``attn_mask`` (True = may not attend) -> the two bit matrices the kernels read
    (rows of the mask and of its transpose).  Cached per mask tensor (and its version counter): the
    decoders pass the same mask to every layer of every step."""
    if mask is None:
        return None
    if mask.dim() != 2 or mask.shape[0] != mask.shape[1] or mask.dtype != torch.bool or not mask.is_cuda:
        raise ValueError('pack_mask expects a square boolean CUDA mask')
    key = (mask.data_ptr(), mask._version, mask.shape[0], mask.device.index)
    hit = _mask_cache.get(key)
    if hit is not None and hit[0] is mask:
        return hit[1]
    # tensors allocated while a CUDA graph is being captured belong to the graph's pool: pack (one tiny
    # captured launch) but do not keep them for later callers
    capturing = torch.cuda.is_current_stream_capturing()
    S = mask.shape[0]
    W = (S + 31) // 32
    m8 = mask.contiguous().view(torch.uint8)
    bits = torch.empty((S, W), dtype=torch.int32, device=mask.device)
    bits_t = torch.empty((S, W), dtype=torch.int32, device=mask.device)
    with torch.cuda.device(mask.device):
        _lib.call('mha_pack_mask', m8.data_ptr(), S, bits.data_ptr(), bits_t.data_ptr(), _stream_ptr(mask))
    if not capturing:
        if len(_mask_cache) > 64:
            _mask_cache.clear()
        _mask_cache[key] = (mask, (bits, bits_t))
    return bits, bits_t


class SelfAttentionCoreFunction(Function):
    """o = dropout_p(softmax(scale * q k^T + mask)) v per (group, head); q | k are the two column halves
    of ``qk`` (rows, 2C), v is (rows, C), o is (rows, C) -- token rows as the layout says."""

    @staticmethod
    @custom_fwd(cast_inputs=None)
    def forward(ctx, qk, v, layout, heads, packed_mask, p, impl):
        if not (qk.is_cuda and v.is_cuda):
            raise RuntimeError('self_attention_core has no CPU path')
        rows, C = v.shape
        if qk.shape != (rows, 2 * C) or qk.dtype != v.dtype or C % heads != 0:
            raise ValueError(f'qk {tuple(qk.shape)} / v {tuple(v.shape)} do not fit {heads} heads')
        if layout.G * layout.S != rows:
            raise ValueError(f'layout {layout} does not cover {rows} token rows')
        qk, v = qk.contiguous(), v.contiguous()
        Dh = C // heads
        bits, bits_t = packed_mask if packed_mask is not None else (None, None)
        out = torch.empty_like(v)
        lse = torch.empty((layout.G * heads, layout.S), dtype=torch.float32, device=v.device)
        ctx.p, ctx.site, key = float(p), 0, None
        if p > 0:
            ctx.site = _next_site()
            key = torch.empty(2, dtype=torch.int64, device=v.device)
        esz = qk.element_size()
        ctx.geom = (layout, heads, Dh, 1.0 / float(Dh) ** 0.5, int(impl))
        with torch.cuda.device(v.device):
            _lib.call('mha_fwd', qk.data_ptr(), qk.data_ptr() + C * esz, v.data_ptr(), out.data_ptr(),
                      lse.data_ptr(), 2 * C, 2 * C, C, C, None if bits is None else bits.data_ptr(),
                      layout.G, heads, layout.S, Dh, layout.seq_stride, layout.hi_stride, layout.lo_stride,
                      layout.n_lo, ctx.geom[3], _DTYPE_CODE[v.dtype], int(impl),
                      dropout_state(v.device).data_ptr() if p > 0 else None,
                      None if key is None else key.data_ptr(), ctx.site, float(p), _stream_ptr(v))
        ctx.masks = (bits, bits_t)
        ctx.save_for_backward(qk, v, out, lse, *([key] if key is not None else []))
        return out

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, dout):
        qk, v, out, lse = ctx.saved_tensors[:4]
        key = ctx.saved_tensors[4] if len(ctx.saved_tensors) > 4 else None
        layout, heads, Dh, scale, impl = ctx.geom
        bits, bits_t = ctx.masks
        C = v.shape[1]
        dout = dout.contiguous()
        dqk = torch.empty_like(qk)
        dv = torch.empty_like(v)
        delta = torch.empty_like(lse)
        esz = qk.element_size()
        with torch.cuda.device(v.device):
            _lib.call('mha_bwd', qk.data_ptr(), qk.data_ptr() + C * esz, v.data_ptr(), out.data_ptr(),
                      dout.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqk.data_ptr(),
                      dqk.data_ptr() + C * esz, dv.data_ptr(), 2 * C, 2 * C, C, C, 0, 0, 0,
                      None if bits is None else bits.data_ptr(), None if bits_t is None else bits_t.data_ptr(),
                      layout.G, heads, layout.S, Dh, layout.seq_stride, layout.hi_stride, layout.lo_stride,
                      layout.n_lo, scale, _DTYPE_CODE[v.dtype], impl,
                      None if key is None else key.data_ptr(), ctx.site, ctx.p, _stream_ptr(v))
        return dqk, dv, None, None, None, None, None


class SelfAttentionFunction(Function):
    """In-projection + attention core of an mmcv-convention self-attention as ONE autograd node:

        q | k = (x + pos) W[:2C]^T + b[:2C],   v = x W[2C:]^T + b[2C:],   o = core(q, k, v)

    (``torch.nn.MultiheadAttention``'s packed ``in_proj_weight`` / ``in_proj_bias``).  Against the composition
    of ``linear`` nodes and :class:`SelfAttentionCoreFunction` the backward has no parameter slices to zero-fill
    and accumulate, no gradient-accumulation adds on x, and one bias reduction: the core's backward writes
    dq | dk | dv into one (rows, 3C) buffer, whose column sums are the bias gradient, whose two column blocks
    give the two halves of dW (written straight into one (3C, C) tensor) and dx = dqk W[:2C] + dv W[2C:]
    (the second GEMM accumulates onto the first)."""

    @staticmethod
    @custom_fwd(cast_inputs=None)
    def forward(ctx, x, pos, weight, bias, layout, heads, packed_mask, p, impl):
        if not x.is_cuda:
            raise RuntimeError('self_attention has no CPU path')
        C = x.shape[-1]
        x2 = x.reshape(-1, C)
        rows = x2.shape[0]
        if layout.G * layout.S != rows or weight.shape != (3 * C, C) or C % heads != 0:
            raise ValueError(f'x {tuple(x.shape)}, in_proj_weight {tuple(weight.shape)}, {layout} do not fit')
        xp = x2 if pos is None else x2 + pos.reshape(-1, C)
        qk = torch.addmm(bias[:2 * C], xp, weight[:2 * C].t())
        v = torch.addmm(bias[2 * C:], x2, weight[2 * C:].t())
        Dh = C // heads
        bits, bits_t = packed_mask if packed_mask is not None else (None, None)
        out = torch.empty_like(v)
        lse = torch.empty((layout.G * heads, layout.S), dtype=torch.float32, device=x.device)
        ctx.p, ctx.site, key = float(p), 0, None
        if p > 0:
            ctx.site = _next_site()
            key = torch.empty(2, dtype=torch.int64, device=x.device)
        esz = qk.element_size()
        ctx.geom = (layout, heads, Dh, 1.0 / float(Dh) ** 0.5, int(impl))
        with torch.cuda.device(x.device):
            _lib.call('mha_fwd', qk.data_ptr(), qk.data_ptr() + C * esz, v.data_ptr(), out.data_ptr(),
                      lse.data_ptr(), 2 * C, 2 * C, C, C, None if bits is None else bits.data_ptr(),
                      layout.G, heads, layout.S, Dh, layout.seq_stride, layout.hi_stride, layout.lo_stride,
                      layout.n_lo, ctx.geom[3], _DTYPE_CODE[v.dtype], int(impl),
                      dropout_state(x.device).data_ptr() if p > 0 else None,
                      None if key is None else key.data_ptr(), ctx.site, float(p), _stream_ptr(x))
        ctx.masks = (bits, bits_t)
        ctx.has_pos = pos is not None
        ctx.x_shape = x.shape
        ctx.pos_shape = None if pos is None else pos.shape
        ctx.save_for_backward(x2, xp, weight, qk, v, out, lse, *([key] if key is not None else []))
        return out.view(x.shape)

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, dout):
        x2, xp, weight, qk, v, out, lse = ctx.saved_tensors[:7]
        key = ctx.saved_tensors[7] if len(ctx.saved_tensors) > 7 else None
        layout, heads, Dh, scale, impl = ctx.geom
        bits, bits_t = ctx.masks
        rows, C = v.shape
        dout = dout.reshape(rows, C).contiguous()
        dqkv = torch.empty((rows, 3 * C), dtype=v.dtype, device=v.device)
        delta = torch.empty_like(lse)
        esz = qk.element_size()
        with torch.cuda.device(v.device):
            _lib.call('mha_bwd', qk.data_ptr(), qk.data_ptr() + C * esz, v.data_ptr(), out.data_ptr(),
                      dout.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(),
                      dqkv.data_ptr() + C * esz, dqkv.data_ptr() + 2 * C * esz, 2 * C, 2 * C, C, C,
                      3 * C, 3 * C, 3 * C, None if bits is None else bits.data_ptr(), None if bits_t is None else bits_t.data_ptr(),
                      layout.G, heads, layout.S, Dh, layout.seq_stride, layout.hi_stride, layout.lo_stride,
                      layout.n_lo, scale, _DTYPE_CODE[v.dtype], impl,
                      None if key is None else key.data_ptr(), ctx.site, ctx.p, _stream_ptr(v))
        dqk, dv = dqkv[:, :2 * C], dqkv[:, 2 * C:]
        dx = dpos = dw = db = None
        need_x, need_pos, need_w, need_b = ctx.needs_input_grad[:4]
        if need_x or (need_pos and ctx.has_pos):
            dxp = dqk @ weight[:2 * C]
            if need_pos and ctx.has_pos:
                dpos = dxp.view(ctx.pos_shape)
            if need_x:
                dx = torch.addmm(dxp, dv, weight[2 * C:]).view(ctx.x_shape)
        if need_w:
            dw = torch.empty_like(weight)
            torch.mm(dqk.t(), xp, out=dw[:2 * C])
            torch.mm(dv.t(), x2, out=dw[2 * C:])
        if need_b:
            db = column_sum(dqkv, weight.dtype) if _colsum_supported(dqkv) else dqkv.sum(0).to(weight.dtype)
        return dx, dpos, dw, db, None, None, None, None, None


def self_attention(x, pos, in_proj_weight, in_proj_bias, layout, heads, attn_mask=None, p=0.0, impl=IMPL_AUTO):
    """``core(q, k, v)`` with q | k = (x + pos) Wqk^T + bqk and v = x Wv^T + bv from the packed in-projection
    parameters of ``torch.nn.MultiheadAttention``; x, pos: (..., C) activations whose rows the layout describes."""
    return SelfAttentionFunction.apply(x, pos, in_proj_weight, in_proj_bias, layout, int(heads),
                                       pack_mask(attn_mask), float(p), int(impl))


def self_attention_core(qk, v, layout, heads, attn_mask=None, p=0.0, impl=IMPL_AUTO):
    """Fused attention core.  ``qk`` (rows, 2C): the projected queries | keys; ``v`` (rows, C);
    ``attn_mask``: boolean (S, S) or None; ``p``: dropout on the attention weights (training)."""
    return SelfAttentionCoreFunction.apply(qk, v, layout, int(heads), pack_mask(attn_mask), float(p), int(impl))


def attention_keep_mask(layout, heads, key, site, p, device):
    """The keep mask the kernels used under (key, site, p), as a bool tensor (G * heads, S, S) (tests)."""
    P = layout.G * heads
    m = torch.empty((P, layout.S, layout.S), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.call('mha_keep_mask', m.data_ptr(), P, layout.S, key.data_ptr(), int(site), float(p),
                  torch.cuda.current_stream(device).cuda_stream)
    return m.bool()
