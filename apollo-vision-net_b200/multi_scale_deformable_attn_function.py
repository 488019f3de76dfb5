"""Operator boundary of the hot path -- drop-in for the reference's
``projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py``.

Same public names, signatures and gradient tuples:

* ``MultiScaleDeformableAttnFunction_fp32.apply(value, value_spatial_shapes,
  value_level_start_index, sampling_locations, attention_weights, im2col_step)``
  (reference ``:88-161``) and the ``_fp16`` twin (``:13-85``);
* ``ext_module.ms_deform_attn_forward / ms_deform_attn_backward`` with the calling
  convention of the mmcv ``_ext`` functions the reference loads at ``:8-10`` and calls
  at ``:40-46`` / ``:72-82``.

Underneath, every call goes through the C ABI of ``libmsda_b200.so``
(``include/msda_b200.h``: ``msda_fwd`` / ``msda_bwd``), hand-written sm_100a kernels.
PyTorch only provides device memory and the current stream.  There is no CPU path:
CPU tensors raise, and a missing library raises.
"""
import types

import torch
from torch.autograd.function import Function, once_differentiable

from . import _lib

try:  # torch >= 2.4
    from torch.amp import custom_bwd as _custom_bwd, custom_fwd as _custom_fwd

    def custom_fwd(cast_inputs):
        return _custom_fwd(device_type='cuda', cast_inputs=cast_inputs)

    def custom_bwd(fn):
        return _custom_bwd(fn, device_type='cuda')
except ImportError:  # pragma: no cover
    from torch.cuda.amp import custom_bwd, custom_fwd  # noqa: F401

_DTYPE_CODE = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _stream_ptr(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(**tensors):
    for name, t in tensors.items():
        if t is None:                  # optional argument not given
            continue
        if not isinstance(t, torch.Tensor):
            raise TypeError(f'{name} must be a torch.Tensor, got {type(t)}')
        if not t.is_cuda:
            raise RuntimeError(
                f'{name} must be a CUDA tensor: the B200 deformable-attention op has no CPU '
                'implementation (the reference uses multi_scale_deformable_attn_pytorch there)')


def _dims(value, sampling_locations, attention_weights, spatial_shapes, level_start_index):
    if value.dim() != 4:
        raise RuntimeError(f'value must be (bs, num_keys, num_heads, head_dim), got {tuple(value.shape)}')
    if sampling_locations.dim() != 6 or sampling_locations.shape[-1] != 2:
        raise RuntimeError('sampling_locations must be (bs, num_queries, num_heads, num_levels, '
                           f'num_points, 2), got {tuple(sampling_locations.shape)}')
    B, Nk, M, Dh = value.shape
    Bq, Nq, Mq, L, P, _ = sampling_locations.shape
    if (Bq, Mq) != (B, M):
        raise RuntimeError('batch / head mismatch between value and sampling_locations: '
                           f'{tuple(value.shape)} vs {tuple(sampling_locations.shape)}')
    if tuple(attention_weights.shape) != (B, Nq, M, L, P):
        raise RuntimeError(f'attention_weights must be {(B, Nq, M, L, P)}, got '
                           f'{tuple(attention_weights.shape)}')
    if tuple(spatial_shapes.shape) != (L, 2) or level_start_index.numel() != L:
        raise RuntimeError('spatial_shapes must be (num_levels, 2) and level_start_index (num_levels,)')
    return B, Nk, M, Dh, L, Nq, P


def _prep(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    _require_cuda(value=value, value_spatial_shapes=spatial_shapes,
                  value_level_start_index=level_start_index,
                  sampling_locations=sampling_locations, attention_weights=attention_weights)
    if value.dtype not in _DTYPE_CODE:
        raise RuntimeError(f'unsupported value dtype {value.dtype}')
    value = value.contiguous()
    coord_dtype = sampling_locations.dtype
    if coord_dtype != torch.float32 and coord_dtype != value.dtype:
        coord_dtype = torch.float32
    loc = sampling_locations.to(coord_dtype).contiguous()
    attn = attention_weights.to(coord_dtype).contiguous()
    shapes = spatial_shapes.to(torch.int64).contiguous()
    starts = level_start_index.to(torch.int64).contiguous()
    return value, shapes, starts, loc, attn


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_locations,
                           attention_weights, im2col_step=64):
    """Same contract as mmcv ``_ext.ms_deform_attn_forward``: returns (bs, num_queries, M*Dh)."""
    value, shapes, starts, loc, attn = _prep(value, spatial_shapes, level_start_index,
                                             sampling_locations, attention_weights)
    B, Nk, M, Dh, L, Nq, P = _dims(value, loc, attn, shapes, starts)
    out = torch.empty((B, Nq, M * Dh), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        _lib.call('msda_fwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                                 loc.data_ptr(), attn.data_ptr(), out.data_ptr(),
                                 B, Nk, M, Dh, L, Nq, P, _DTYPE_CODE[value.dtype],
                                 _DTYPE_CODE[loc.dtype], int(im2col_step), _stream_ptr(value))
    return out


def _backward_raw(value, shapes, starts, loc, attn, grad_output, im2col_step):
    B, Nk, M, Dh, L, Nq, P = _dims(value, loc, attn, shapes, starts)
    grad_output = grad_output.to(value.dtype).contiguous()
    g_value = torch.zeros(value.shape, dtype=torch.float32, device=value.device)
    g_loc = torch.empty(loc.shape, dtype=torch.float32, device=value.device)
    g_attn = torch.empty(attn.shape, dtype=torch.float32, device=value.device)
    with torch.cuda.device(value.device):
        _lib.call('msda_bwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                                 loc.data_ptr(), attn.data_ptr(), grad_output.data_ptr(),
                                 g_value.data_ptr(), g_loc.data_ptr(), g_attn.data_ptr(),
                                 B, Nk, M, Dh, L, Nq, P, _DTYPE_CODE[value.dtype],
                                 _DTYPE_CODE[loc.dtype], int(im2col_step), _stream_ptr(value))
    return g_value, g_loc, g_attn


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                            grad_output, grad_value, grad_sampling_loc, grad_attn_weight,
                            im2col_step=64):
    """Same contract as mmcv ``_ext.ms_deform_attn_backward``: accumulates into the three
    caller-allocated (zero-initialised) gradient buffers and returns None."""
    _require_cuda(grad_output=grad_output, grad_value=grad_value,
                  grad_sampling_loc=grad_sampling_loc, grad_attn_weight=grad_attn_weight)
    value, shapes, starts, loc, attn = _prep(value, spatial_shapes, level_start_index,
                                             sampling_loc, attn_weight)
    g_value, g_loc, g_attn = _backward_raw(value, shapes, starts, loc, attn, grad_output,
                                           im2col_step)
    grad_value.add_(g_value.to(grad_value.dtype))
    grad_sampling_loc.add_(g_loc.to(grad_sampling_loc.dtype))
    grad_attn_weight.add_(g_attn.to(grad_attn_weight.dtype))
    return None


ext_module = types.SimpleNamespace(ms_deform_attn_forward=ms_deform_attn_forward,
                                   ms_deform_attn_backward=ms_deform_attn_backward)


def _forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
             attention_weights, im2col_step):
    ctx.im2col_step = im2col_step
    ctx.in_dtypes = (value.dtype, sampling_locations.dtype, attention_weights.dtype)
    value, shapes, starts, loc, attn = _prep(value, value_spatial_shapes,
                                             value_level_start_index, sampling_locations,
                                             attention_weights)
    output = ms_deform_attn_forward(value, shapes, starts, loc, attn, im2col_step=im2col_step)
    ctx.save_for_backward(value, shapes, starts, loc, attn)
    return output


def _backward(ctx, grad_output):
    value, shapes, starts, loc, attn = ctx.saved_tensors
    g_value, g_loc, g_attn = _backward_raw(value, shapes, starts, loc, attn, grad_output,
                                           ctx.im2col_step)
    dv, dl, da = ctx.in_dtypes
    return g_value.to(dv), None, None, g_loc.to(dl), g_attn.to(da), None


class MultiScaleDeformableAttnFunction_fp32(Function):
    """fp32 operator (inputs are cast to fp32 under autocast), reference ``:88-161``."""

    @staticmethod
    @custom_fwd(cast_inputs=torch.float32)
    def forward(ctx, value, value_spatial_shapes, value_level_start_index,
                sampling_locations, attention_weights, im2col_step):
        return _forward(ctx, value, value_spatial_shapes, value_level_start_index,
                        sampling_locations, attention_weights, im2col_step)

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, grad_output):
        return _backward(ctx, grad_output)


class MultiScaleDeformableAttnFunction_fp16(Function):
    """Half-precision operator (inputs are cast to fp16 under autocast), reference ``:13-85``.

    Dead code in the reference (never selected); here it is live: fp16 or
    bf16 value / output with fp32 accumulation, locations and weights in fp32 or the value dtype.
    """

    @staticmethod
    @custom_fwd(cast_inputs=torch.float16)
    def forward(ctx, value, value_spatial_shapes, value_level_start_index,
                sampling_locations, attention_weights, im2col_step):
        return _forward(ctx, value, value_spatial_shapes, value_level_start_index,
                        sampling_locations, attention_weights, im2col_step)

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, grad_output):
        return _backward(ctx, grad_output)


class MultiScaleDeformableAttnFunction_bf16(Function):
    """bf16 twin of ``_fp16`` (not in the reference): under ``torch.autocast(dtype=bfloat16)`` the
    inputs are cast to bf16, not to fp16 -- ``_fp16``'s decorator would silently turn a bf16 value
    into fp16 (range 65504) -- and outside autocast nothing is cast.  The module-level paths pick
    the Function by the value dtype (``deform_common.msda_apply``)."""

    @staticmethod
    @custom_fwd(cast_inputs=torch.bfloat16)
    def forward(ctx, value, value_spatial_shapes, value_level_start_index,
                sampling_locations, attention_weights, im2col_step):
        return _forward(ctx, value, value_spatial_shapes, value_level_start_index,
                        sampling_locations, attention_weights, im2col_step)

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, grad_output):
        return _backward(ctx, grad_output)
