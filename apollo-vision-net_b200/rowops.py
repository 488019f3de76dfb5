"""Row-wise companions of the attention kernels inside a BEVFormer layer (SURVEY.md section 8f
rank 3): LayerNorm and Linear with sm_100a kernels for the pieces where the stock kernels were
the largest non-attention items of the measured step (LayerNorm backward, bias-gradient column
sums).  ``LayerNorm`` / ``Linear`` subclass the torch modules, so parameter names, initialisation
and ``state_dict`` layout are unchanged (the reference builds them through mmcv's
``build_norm_layer(dict(type='LN'))`` / ``nn.Linear``, custom_base_transformer_layer.py:142-161).

The GEMMs themselves stay on cuBLAS (library GEMMs, tensor cores): only the reductions and the
normalisation are ours.  CPU tensors and unsupported widths take the torch implementation --
these are not hot-path operators.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd.function import Function, once_differentiable

from . import _lib
from .multi_scale_deformable_attn_function import _DTYPE_CODE, _stream_ptr, custom_bwd, custom_fwd

_workspaces = {}


_WS_HEADER = 64      # floats: ticket counter of the "last CTA finishes" reduction (kept at zero)


def _workspace(device, floats):
    """Zero-initialised scratch, one per (device, stream): [64-float header | partial rows].  The
    kernels leave the header at zero, so the buffer is reusable without a memset."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < floats + _WS_HEADER:
        ws = torch.zeros(max(floats + _WS_HEADER, 1 << 20), dtype=torch.float32, device=device)
        _workspaces[key] = ws
    return ws


_LN_WIDTHS = (128, 256, 512, 1024)

# ---- dropout state (masks are counter-based and recomputed in the backward, see msda_b200.h) -------
_drop_state = {}
_drop_site = [0]


def dropout_state(device):
    """uint64 pair (seed, step) on ``device`` (stored as int64); the seed is drawn from torch's CPU
    generator at first use, so ``torch.manual_seed`` makes runs repeatable."""
    st = _drop_state.get(device.index)
    if st is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        st = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        _drop_state[device.index] = st
    return st


def advance_dropout_step(device):
    """New masks for the next training step: increments the step counter ON THE DEVICE (a captured CUDA
    graph replays the increment, so every replay draws fresh masks).  Called by the encoder / decoders at
    the start of a training forward."""
    dropout_state(device)[1:].add_(1)


def reseed_dropout(device, seed):
    st = dropout_state(device)
    st.copy_(torch.tensor([int(seed), 0], dtype=torch.int64))


def _next_site():
    _drop_site[0] = (_drop_site[0] + 1) & 0x7fffffff
    return _drop_site[0]


def dropout_keep_mask(shape, dtype, key, site, p, device):
    """The keep mask the fused kernels use for a tensor of ``shape`` / ``dtype`` under (key, site, p), as a
    bool tensor (tests)."""
    n = 1
    for d in shape:
        n *= int(d)
    m = torch.empty(n, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.call('dropout_keep_mask', m.data_ptr(), n, _DTYPE_CODE[dtype], key.data_ptr(), int(site), float(p),
                  torch.cuda.current_stream(device).cuda_stream)
    return m.view(*shape).bool()


def _ln_supported(x, weight, bias):
    if not x.is_cuda or weight is None or bias is None or x.dtype not in _DTYPE_CODE:
        return False
    # exactly the widths csrc/rowops.cu instantiates (ln_dispatch: C / 32 in {4, 8, 16, 32}); any
    # other width takes torch's layer_norm instead of failing in the kernel dispatch
    return x.shape[-1] in _LN_WIDTHS and weight.dtype == x.dtype and bias.dtype == x.dtype


class LayerNormFunction(Function):
    @staticmethod
    @custom_fwd(cast_inputs=None)
    def forward(ctx, x, weight, bias, eps):
        shape = x.shape
        C = shape[-1]
        x2 = x.reshape(-1, C).contiguous()
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        w, b = weight.contiguous(), bias.contiguous()
        with torch.cuda.device(x.device):
            _lib.call('ln_fwd', x2.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(),
                      mean.data_ptr(), rstd.data_ptr(), rows, C, float(eps), _DTYPE_CODE[x.dtype],
                      _stream_ptr(x))
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.shape = shape
        return y.view(shape)

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, dy):
        x2, w, mean, rstd = ctx.saved_tensors
        rows, C = x2.shape
        dy2 = dy.reshape(rows, C).to(x2.dtype).contiguous()
        dx = torch.empty_like(x2)
        dgb = torch.empty((2, C), dtype=x2.dtype, device=x2.device)
        nrows = _lib.lib().rowops_workspace_rows()
        ws = _workspace(x2.device, nrows * 2 * C)
        with torch.cuda.device(x2.device):
            _lib.call('ln_bwd', x2.data_ptr(), dy2.data_ptr(), w.data_ptr(), mean.data_ptr(),
                      rstd.data_ptr(), dx.data_ptr(), dgb.data_ptr(), ws.data_ptr(), rows, C,
                      _DTYPE_CODE[x2.dtype], _stream_ptr(x2))
        return dx.view(ctx.shape), dgb[0], dgb[1], None


class LayerNorm(nn.LayerNorm):
    """``torch.nn.LayerNorm`` over the last dimension with sm_100a forward / backward kernels."""

    def forward(self, x):
        if len(self.normalized_shape) == 1 and _ln_supported(x, self.weight, self.bias):
            return LayerNormFunction.apply(x, self.weight, self.bias, self.eps)
        return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)


class Junction:
    """The gradient junction at the input of a residual block, made explicit.

    A post-norm block uses its input twice: as the residual of its tail (``LinearAddLayerNormFunction``) and
    as the input of the branch's first Linear layer.  autograd would add the two gradients with a separate
    kernel (60 MB of traffic for a 20 MB tensor, 17 times per encoder step).  With a junction shared by the
    two nodes the tail's backward parks the residual gradient here instead of returning it, and the first
    Linear's backward -- which by data dependence always runs later in the same pass -- accumulates its
    ``dy @ W`` onto it through the GEMM's beta = 1 epilogue and returns the sum as THE gradient of the input.
    A gradient that was parked and never collected raises at the end of the backward pass."""

    __slots__ = ('ds', 'consumers', 'checked')

    def __init__(self):
        self.ds = None
        self.consumers = 0
        self.checked = False

    def park(self, ds):
        self.ds = ds
        if not self.checked:
            self.checked = True
            torch.autograd.Variable._execution_engine.queue_callback(self._check)

    def take(self):
        ds, self.ds = self.ds, None
        return ds

    def _check(self):
        self.checked = False
        if self.ds is not None:
            self.ds = None
            raise RuntimeError('a residual gradient parked at a Junction was never collected: the branch of the '
                               'block did not run its backward (detached branch?)')


def _accumulate_dx(junction, dy2, weight, shape):
    """``dy2 @ weight`` as the input gradient of a Linear layer, added onto a parked residual gradient when
    the layer's input is a block's junction (one GEMM with beta = 1, no separate add kernel)."""
    ds = junction.take() if junction is not None else None
    if ds is not None and ds.is_contiguous() and ds.dtype == dy2.dtype and ds.numel() == dy2.shape[0] * weight.shape[1]:
        return ds.view(dy2.shape[0], weight.shape[1]).addmm_(dy2, weight).view(shape)
    dx = (dy2 @ weight).view(shape)
    return dx if ds is None else dx + ds.view(shape)


class LinearAddLayerNormFunction(Function):
    """y = LayerNorm(x W^T + b + residual): the post-norm tail of an attention / FFN block as one
    autograd node.  Forward: cuBLAS GEMM, then one kernel for the residual add and the
    normalisation (the sum overwrites the GEMM output and is what the backward re-reads).
    Backward: one kernel for d sum, d gamma, d beta AND the bias gradient (column sums of d sum),
    then the two GEMMs; d residual is d sum itself."""

    @staticmethod
    @custom_fwd(cast_inputs=None)
    def forward(ctx, x, weight, bias, residual, gamma, beta, eps, p=0.0, junction=None):
        """``p`` > 0: y = LayerNorm(dropout(x W^T + b) + residual) with the mask drawn in the kernel.
        ``junction``: park the residual's gradient there (see :class:`Junction`) instead of returning it."""
        ctx.junction = junction if (junction is not None and junction.consumers > 0) else None
        C = weight.shape[0]
        x2 = x.reshape(-1, x.shape[-1])
        rows = x2.shape[0]
        res2 = residual.reshape(rows, C).contiguous()
        lin = torch.empty((rows, C), dtype=x.dtype, device=x.device)
        torch.addmm(bias, x2, weight.t(), out=lin)
        y = torch.empty_like(lin)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        g, b = gamma.contiguous(), beta.contiguous()
        ctx.p, ctx.site, key = float(p), 0, None
        with torch.cuda.device(x.device):
            if p > 0:
                ctx.site = _next_site()
                key = torch.empty(2, dtype=torch.int64, device=x.device)
                _lib.call('ln_residual_dropout_fwd', lin.data_ptr(), res2.data_ptr(), g.data_ptr(), b.data_ptr(),
                          lin.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, C,
                          float(eps), _DTYPE_CODE[x.dtype], dropout_state(x.device).data_ptr(), key.data_ptr(),
                          ctx.site, float(p), _stream_ptr(x))
            else:
                _lib.call('ln_residual_fwd', lin.data_ptr(), res2.data_ptr(), g.data_ptr(), b.data_ptr(),
                          lin.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, C,
                          float(eps), _DTYPE_CODE[x.dtype], _stream_ptr(x))
        ctx.save_for_backward(x2, weight, lin, g, mean, rstd, *([key] if key is not None else []))
        ctx.shapes = (x.shape, residual.shape)
        return y.view(residual.shape)

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, dy):
        x2, weight, s, g, mean, rstd, *key = ctx.saved_tensors
        rows, C = s.shape
        dy2 = dy.reshape(rows, C).to(s.dtype).contiguous()
        ds = torch.empty_like(s)
        out3 = torch.empty((3, C), dtype=s.dtype, device=s.device)
        nrows = _lib.lib().rowops_workspace_rows()
        ws = _workspace(s.device, nrows * 3 * C)
        dlin = ds                          # gradient of the Linear output (= d sum without dropout)
        with torch.cuda.device(s.device):
            if ctx.p > 0:
                dlin = torch.empty_like(s)
                _lib.call('ln_bwd_dxsum_dropout', s.data_ptr(), dy2.data_ptr(), g.data_ptr(), mean.data_ptr(),
                          rstd.data_ptr(), ds.data_ptr(), dlin.data_ptr(), out3.data_ptr(), ws.data_ptr(), rows,
                          C, _DTYPE_CODE[s.dtype], key[0].data_ptr(), ctx.site, ctx.p, _stream_ptr(s))
            else:
                _lib.call('ln_bwd_dxsum', s.data_ptr(), dy2.data_ptr(), g.data_ptr(), mean.data_ptr(),
                          rstd.data_ptr(), ds.data_ptr(), out3.data_ptr(), ws.data_ptr(), rows, C,
                          _DTYPE_CODE[s.dtype], _stream_ptr(s))
        x_shape, res_shape = ctx.shapes
        dx = (dlin @ weight).view(x_shape) if ctx.needs_input_grad[0] else None
        dw = weight_bias_grad(dlin, x2, weight, want_bias=False)[0] if ctx.needs_input_grad[1] else None
        dres = ds.view(res_shape)
        if ctx.junction is not None and ctx.needs_input_grad[3]:
            ctx.junction.park(ds)
            dres = None
        return dx, dw, out3[2], dres, out3[0], out3[1], None, None, None


def linear_add_layernorm(x, linear_mod, residual, norm, p=0.0, junction=None):
    """``norm(dropout_p(linear_mod(x)) + residual)`` through :class:`LinearAddLayerNormFunction` when
    the tensors qualify (CUDA, one dtype, supported width, affine LayerNorm over the last dim with a
    bias-carrying Linear); the plain composition otherwise.  ``p``: dropout probability of the training
    step (0 = none)."""
    w, b = linear_mod.weight, linear_mod.bias
    ok = (isinstance(norm, nn.LayerNorm) and len(norm.normalized_shape) == 1 and b is not None and
          x.is_cuda and x.dtype == w.dtype == residual.dtype and
          residual.shape[-1] == w.shape[0] and residual.numel() == x.numel() // x.shape[-1] * w.shape[0] and
          norm.weight is not None and norm.bias is not None and
          _ln_supported(residual, norm.weight, norm.bias))
    if not ok:
        out = linear_mod(x)
        return norm((F.dropout(out, p, True) if p > 0 else out) + residual)
    return LinearAddLayerNormFunction.apply(x, w, b, residual, norm.weight, norm.bias, norm.eps, float(p),
                                            junction)


# ---- bias gradients that a producer kernel already has -----------------------------------------
# The backward of the fused attention kernels ends with a pass over grad_value (unscale_cast); that pass
# also returns the column sums, which are exactly the bias gradient of the value projection whose output
# the attention consumed.  autograd only carries tensors, so the sums travel beside the gradient: the
# producer registers them under the gradient's storage, LinearFunction.backward looks its dy up.  An entry
# is honoured only while the registered tensor is alive (its memory cannot have been reused), still
# unmodified (version counter), and covers exactly dy's bytes.
_bias_grads = {}


# ---- side-by-side gradient blocks of hoisted projections ------------------------------------------
# Several Linear layers project the SAME input (the camera features, once per encoder layer).  When their outputs'
# gradients are written as column blocks of ONE (rows, layers * C) matrix, the input gradient is a single GEMM
# over K = layers * C and the weight gradients a single GEMM with layers * C output rows -- instead of one GEMM
# pair per layer plus accumulation.  The producer of such a gradient (fused_ops._finish_accumulator) asks here
# whether the tensor whose gradient it is about to write has a slot.
class GradSlots:
    """(rows, n * C) buffer, allocated when the first gradient arrives; block l belongs to output l."""

    def __init__(self, n, rows, C, dtype, device):
        self.n, self.rows, self.C, self.dtype, self.device = n, rows, C, dtype, device
        self.buffer = None
        self.filled = [False] * n

    def block(self, l):
        if self.buffer is None:
            self.buffer = torch.empty((self.rows, self.n * self.C), dtype=self.dtype, device=self.device)
        self.filled[l] = True
        return self.buffer[:, l * self.C:(l + 1) * self.C]


_grad_slots = {}


def register_grad_slot(out, slots, l):
    import weakref
    if len(_grad_slots) > 256:
        for k in [k for k, (r, _) in _grad_slots.items() if r() is None]:
            del _grad_slots[k]
    _grad_slots[(out.untyped_storage().data_ptr(), out.storage_offset())] = (weakref.ref(slots), l)


def grad_slot(t):
    """The column block ``t``'s gradient should be written into ((rows, C) view, row stride = the buffer's
    width), or None."""
    ent = _grad_slots.get((t.untyped_storage().data_ptr(), t.storage_offset())) if _grad_slots else None
    if ent is None:
        return None
    slots = ent[0]()
    if slots is None or t.numel() != slots.rows * slots.C or t.dtype != slots.dtype or not t.is_contiguous():
        return None
    return slots.block(ent[1])


def offer_bias_grad(grad, colsum):
    """``colsum`` = sum over all rows of ``grad`` viewed as (-1, colsum.numel())."""
    import weakref
    if len(_bias_grads) > 64:                               # entries nobody collected
        for k in [k for k, (r, *_) in _bias_grads.items() if r() is None]:
            del _bias_grads[k]
    # (a view's root keeps the memory alive and shares the version counter with it; the view object itself
    # may be gone by the time the consumer runs)
    root = grad._base if grad._base is not None else grad
    _bias_grads[(grad.untyped_storage().data_ptr(), grad.storage_offset())] = (
        weakref.ref(root), grad._version, grad.numel(), colsum, grad.dtype, _rows_layout(grad, colsum.numel()))


def _rows_layout(t, C):
    """Strides of the (rows, C) view of ``t``; None when ``t`` is not such a view."""
    try:
        return t.view(-1, C).stride()
    except RuntimeError:
        return None


def _take_bias_grad(dy, C):
    ent = _bias_grads.pop((dy.untyped_storage().data_ptr(), dy.storage_offset()), None) if _bias_grads else None
    if ent is None:
        return None
    ref, version, numel, colsum, dtype, layout = ent
    g = ref()
    # honoured only while the offered memory is alive and unmodified and the consumer's tensor covers exactly the
    # offered bytes: same start (the key), same number of elements, same (rows, C) layout -- which may be a column
    # block of a wider matrix (grad_slot())
    if (g is None or g._version != version or numel != dy.numel() or colsum.numel() != C or dy.dtype != dtype or
            layout is None or _rows_layout(dy, C) != layout):
        return None
    return colsum


class ReLUFunction(Function):
    """In-place ReLU whose backward also sums its result over the rows: the bias gradient of the Linear
    layer in front of it (offered to that layer's backward through :func:`offer_bias_grad`)."""

    @staticmethod
    def forward(ctx, x, p=0.0):
        """``p`` > 0: dropout(relu(x)) in one in-place pass (the FFN's Linear-ReLU-Dropout)."""
        ctx.scale = 1.0
        if p > 0:
            with torch.cuda.device(x.device):
                _lib.call('relu_dropout_fwd', x.data_ptr(), x.numel(), _DTYPE_CODE[x.dtype],
                          dropout_state(x.device).data_ptr(), None, _next_site(), float(p), _stream_ptr(x))
            ctx.scale = 1.0 / (1.0 - float(p))
            y = x
        else:
            y = torch.relu_(x)
        ctx.mark_dirty(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        y, = ctx.saved_tensors
        C = y.shape[-1]
        dy2 = dy.reshape(-1, C).to(y.dtype).contiguous()
        y2 = y.reshape(-1, C)
        dx = torch.empty_like(dy2)
        sums = torch.empty(C, dtype=y.dtype, device=y.device)
        ws = _workspace(y.device, _lib.lib().rowops_workspace_rows() * C)
        with torch.cuda.device(y.device):
            _lib.call('relu_bwd_colsum', dy2.data_ptr(), y2.data_ptr(), dx.data_ptr(), sums.data_ptr(),
                      ws.data_ptr(), dy2.shape[0], C, _DTYPE_CODE[y.dtype], _DTYPE_CODE[y.dtype],
                      float(ctx.scale), _stream_ptr(y))
        dx = dx.view(y.shape)
        offer_bias_grad(dx, sums)
        return dx, None


class ReLU(nn.ReLU):
    """``nn.ReLU(inplace=True)`` after a Linear layer: same forward; on CUDA the backward is one kernel
    that also yields that Linear's bias gradient."""

    def fused_ok(self, x):
        return (self.inplace and x.is_cuda and x.requires_grad and torch.is_grad_enabled() and x.is_contiguous()
                and x.dim() >= 2 and _colsum_supported(x.reshape(-1, x.shape[-1])))

    def forward(self, x):
        if self.fused_ok(x):
            return ReLUFunction.apply(x, 0.0)
        return super().forward(x)

    def forward_dropout(self, x, p):
        """dropout_p(relu(x)) -- one in-place kernel when the tensor qualifies."""
        if p > 0 and self.fused_ok(x) and x.numel() % 8 == 0:
            return ReLUFunction.apply(x, float(p))
        return F.dropout(self.forward(x), p, True) if p > 0 else self.forward(x)


def _wgrad_supported(dy2, x2, weight, want_bias=True):
    """Whether ``linear_wgrad`` (tcgen05 split-row kernel, csrc/wgrad.cu) takes this gradient:
    16-bit, many rows, a small O x I output.  OPT-IN: ``APOLLO_B200_WGRAD=1`` routes every such
    shape to it, ``=bias`` only those whose bias gradient it fuses; default off (the library's
    split-K GEMM plus our column-sum kernel is faster inside the measured step, see wgrad.cu)."""
    import os
    mode = os.environ.get('APOLLO_B200_WGRAD', '0')
    if mode == '0' or (mode == 'bias' and not want_bias):
        return False
    O, I = weight.shape
    return (dy2.is_cuda and dy2.dtype in (torch.bfloat16, torch.float16) and dy2.dtype == x2.dtype == weight.dtype and
            O % 8 == 0 and I % 8 == 0 and O * I <= 131072 and dy2.shape[0] >= 8192)


def weight_bias_grad(dy2, x2, weight, want_bias=True):
    """(dW, db) of ``y = x W^T + b`` from dy (N, O) and x (N, I): one pass over both through the
    tensor-core split-row kernel ``linear_wgrad`` where it applies, else the library GEMM plus the
    column-sum kernel."""
    if _wgrad_supported(dy2, x2, weight, want_bias):
        O, I = weight.shape
        dy2c = dy2 if dy2.is_contiguous() else dy2.contiguous()
        x2c = x2 if x2.is_contiguous() else x2.contiguous()
        dW = torch.empty((O, I), dtype=weight.dtype, device=weight.device)
        db = torch.empty((O,), dtype=weight.dtype, device=weight.device) if want_bias else None
        ws = _workspace(dy2.device, int(_lib.lib().linear_wgrad_workspace_floats(O, I)))
        with torch.cuda.device(dy2.device):
            _lib.call('linear_wgrad', dy2c.data_ptr(), x2c.data_ptr(), dW.data_ptr(),
                      None if db is None else db.data_ptr(), ws.data_ptr(), dy2c.shape[0], O, I,
                      _DTYPE_CODE[dy2.dtype], _stream_ptr(dy2))
        return dW, db
    dW = dy2.t() @ x2
    db = None
    if want_bias:
        dy2c = dy2 if dy2.is_contiguous() else dy2.contiguous()
        db = column_sum(dy2c, weight.dtype) if _colsum_supported(dy2c) else dy2.sum(0)
    return dW, db


def column_sum(x2, out_dtype=None):
    """Sum over the rows of a (rows, C) CUDA matrix (fp32 accumulation)."""
    rows, C = x2.shape
    out_dtype = out_dtype or x2.dtype
    out = torch.empty(C, dtype=out_dtype, device=x2.device)
    ws = _workspace(x2.device, _lib.lib().rowops_workspace_rows() * C)
    with torch.cuda.device(x2.device):
        _lib.call('colsum', x2.data_ptr(), out.data_ptr(), ws.data_ptr(), rows, C,
                  _DTYPE_CODE[x2.dtype], _DTYPE_CODE[out_dtype], _stream_ptr(x2))
    return out


def _colsum_supported(g2):
    if not g2.is_cuda or g2.dtype not in _DTYPE_CODE:
        return False
    vec = 4 if g2.dtype == torch.float32 else 8
    C = g2.shape[1]
    return C % vec == 0 and C // vec <= 256 and g2.shape[0] >= 1024


class LinearFunction(Function):
    """y = x W^T + b with cuBLAS GEMMs and our column-sum kernel for the bias gradient."""

    @staticmethod
    @custom_fwd(cast_inputs=None)
    def forward(ctx, x, weight, bias, junction=None):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.junction = None
        if junction is not None and ctx.needs_input_grad[0]:
            junction.consumers += 1
            ctx.junction = junction
        # write into a freshly allocated tensor of the final shape: the result must not be a
        # view (callers apply in-place activations to it, e.g. the FFN's ReLU(inplace=True))
        out = torch.empty(x.shape[:-1] + (weight.shape[0],), dtype=x.dtype, device=x.device)
        x2 = x.reshape(-1, x.shape[-1])
        o2 = out.view(-1, weight.shape[0])
        if bias is not None:
            torch.addmm(bias, x2, weight.t(), out=o2)
        else:
            torch.mm(x2, weight.t(), out=o2)
        return out

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _accumulate_dx(ctx.junction, dy2, weight, x.shape)
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        ready = _take_bias_grad(dy, weight.shape[0]) if want_b else None
        if ready is not None:                               # the producer of dy already summed its rows
            db = ready.to(weight.dtype)
            want_b = False
        if ctx.needs_input_grad[1]:
            dw, db2 = weight_bias_grad(dy2, x2, weight, want_bias=want_b)
            db = db2 if want_b else db
        elif want_b:
            dy2c = dy2 if dy2.is_contiguous() else dy2.contiguous()
            db = column_sum(dy2c, weight.dtype) if _colsum_supported(dy2c) else dy2.sum(0)
        return dx, dw, db, None


class PairedQueryLinearFunction(Function):
    """``[paired | query + pos] W^T + b`` -- the offsets / weights projection of TemporalSelfAttention, whose
    input is the previous BEV's cell next to the positioned query (temporal_self_attention.py:199-204).
    The backward computes only the halves of dX somebody needs (``paired`` is history: no gradient) and, when
    ``query`` is the block's residual, accumulates the query half onto the parked residual gradient
    (:class:`Junction`) -- unless ``pos`` wants its own gradient, which is that half alone."""

    @staticmethod
    @custom_fwd(cast_inputs=None)
    def forward(ctx, paired, query, pos, weight, bias, junction=None):
        C = query.shape[-1]
        rows = query.numel() // C
        # (one vectorised add and one cat: writing the sum straight into the right half of the buffer takes
        # torch's strided element-wise kernels, 52 us against 28 us for the pair at 40 000 x 256)
        cat = torch.cat([paired, query if pos is None else query + pos], -1)
        out = torch.empty(query.shape[:-1] + (weight.shape[0],), dtype=query.dtype, device=query.device)
        torch.addmm(bias, cat.view(rows, 2 * C), weight.t(), out=out.view(rows, weight.shape[0]))
        ctx.save_for_backward(cat, weight)
        ctx.shapes = (paired.shape, query.shape, None if pos is None else pos.shape)
        ctx.junction = None
        if junction is not None and ctx.needs_input_grad[1]:
            junction.consumers += 1
            ctx.junction = junction
        return out

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, dy):
        cat, weight = ctx.saved_tensors
        paired_shape, query_shape, pos_shape = ctx.shapes
        O = weight.shape[0]
        C = cat.shape[-1] // 2
        dy2 = dy.reshape(-1, O)
        d_paired = d_query = d_pos = dw = db = None
        if ctx.needs_input_grad[0]:
            d_paired = (dy2 @ weight[:, :C]).view(paired_shape)
        w_q = weight[:, C:]
        if pos_shape is not None and ctx.needs_input_grad[2]:
            dq = (dy2 @ w_q).view(query_shape)
            d_pos = dq
            if ctx.needs_input_grad[1]:
                ds = ctx.junction.take() if ctx.junction is not None else None
                d_query = dq if ds is None else dq + ds.view(query_shape)
        elif ctx.needs_input_grad[1]:
            d_query = _accumulate_dx(ctx.junction, dy2, w_q, query_shape)
        if ctx.needs_input_grad[3]:
            dw, db = weight_bias_grad(dy2, cat.view(-1, 2 * C), weight, want_bias=ctx.needs_input_grad[4])
        elif ctx.needs_input_grad[4]:
            dy2c = dy2 if dy2.is_contiguous() else dy2.contiguous()
            db = column_sum(dy2c, weight.dtype) if _colsum_supported(dy2c) else dy2.sum(0)
        return d_paired, d_query, d_pos, dw, db, None


def paired_query_linear(paired, query, pos, weight, bias, junction=None):
    """See :class:`PairedQueryLinearFunction`; the plain composition when the tensors do not qualify."""
    if (query.is_cuda and torch.is_grad_enabled() and query.dtype == weight.dtype == paired.dtype and
            query.dtype in _DTYPE_CODE and bias is not None and paired.shape == query.shape and
            (pos is None or (pos.shape == query.shape and pos.dtype == query.dtype))):
        return PairedQueryLinearFunction.apply(paired, query, pos, weight, bias, junction)
    q = query if pos is None else query + pos
    return linear(torch.cat([paired, q], -1), weight, bias)


def linear(x, weight, bias=None, junction=None):
    """``F.linear`` whose backward computes the bias gradient with the column-sum kernel.  ``junction``: ``x``
    is also the residual of the block's tail, whose gradient this layer's backward collects (:class:`Junction`)."""
    if x.is_cuda and torch.is_grad_enabled() and x.dtype == weight.dtype and x.dtype in _DTYPE_CODE:
        return LinearFunction.apply(x, weight, bias, junction)
    return F.linear(x, weight, bias)


class Linear(nn.Linear):
    """``torch.nn.Linear`` whose backward computes the bias gradient with the column-sum kernel."""

    def forward(self, x):
        return linear(x, self.weight, self.bias)
