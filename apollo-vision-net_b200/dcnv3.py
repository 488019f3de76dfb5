"""DCNv3 on the sm_100a kernels (``csrc/dcnv3.cu`` -> the MSDA operator kernels in pixel-coordinate mode).

Drop-in for ``projects/mmdet3d_plugin/bevformer/backbones/ops_dcnv3``: :data:`ext_module` exposes the two
entry points of the reference's compiled ``DCNv3`` extension (``dcnv3_forward`` / ``dcnv3_backward``,
src/dcnv3.h:21-59), :class:`DCNv3Function` has the reference Function's argument list
(functions/dcnv3_func.py:19-62) and :class:`DCNv3` the reference module's constructor, parameter names
and forward (modules/dcnv3.py:216-345), so checkpoints load unchanged.  No CPU path: CPU tensors raise, a
missing library raises.
"""
import types
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.init import constant_, xavier_uniform_

from . import _lib
from .multi_scale_deformable_attn_function import _DTYPE_CODE, _stream_ptr, custom_bwd, custom_fwd


def _geometry(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
              group, group_channels):
    if not (input.is_cuda and offset.is_cuda and mask.is_cuda):
        raise RuntimeError('DCNv3: not implemented on the CPU (as the reference, src/dcnv3.h:37)')
    if input.dtype not in _DTYPE_CODE or offset.dtype != input.dtype or mask.dtype != input.dtype:
        raise TypeError('DCNv3: input, offset and mask must share one of float32 / float16 / bfloat16')
    N, H_in, W_in, C = input.shape
    _, H_out, W_out, _ = offset.shape
    K = kernel_h * kernel_w
    if C != group * group_channels:
        raise ValueError(f'Input channels and group times group channels wont match: ({C} vs {group * group_channels}).')
    if tuple(offset.shape) != (N, H_out, W_out, group * K * 2) or tuple(mask.shape) != (N, H_out, W_out, group * K):
        raise ValueError('DCNv3: offset / mask shapes do not fit the kernel and the group count')
    return N, H_in, W_in, H_out, W_out


def _scratch(input, N, H_out, W_out, group, kernel_h, kernel_w):
    n = int(_lib.lib().dcnv3_scratch_floats(N, H_out, W_out, group, kernel_h, kernel_w, _DTYPE_CODE[input.dtype]))
    return torch.empty(n, dtype=torch.float32, device=input.device)


def dcnv3_forward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                  dilation_w, group, group_channels, offset_scale, im2col_step=256):
    """``DCNv3.dcnv3_forward`` of the reference extension; ``im2col_step`` is accepted and ignored."""
    N, H_in, W_in, H_out, W_out = _geometry(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                                            pad_w, dilation_h, dilation_w, group, group_channels)
    input, offset, mask = input.contiguous(), offset.contiguous(), mask.contiguous()
    out = torch.empty((N, H_out, W_out, group * group_channels), dtype=input.dtype, device=input.device)
    scratch = _scratch(input, N, H_out, W_out, group, kernel_h, kernel_w)
    with torch.cuda.device(input.device):
        _lib.call('dcnv3_fwd', input.data_ptr(), offset.data_ptr(), mask.data_ptr(), out.data_ptr(),
                  scratch.data_ptr(), N, H_in, W_in, H_out, W_out, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                  pad_w, dilation_h, dilation_w, group, group_channels, float(offset_scale),
                  _DTYPE_CODE[input.dtype], _stream_ptr(input))
    return out


def dcnv3_backward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                   dilation_w, group, group_channels, offset_scale, grad_output, im2col_step=256):
    """``DCNv3.dcnv3_backward``: (grad_input, grad_offset, grad_mask) in the dtype of the inputs."""
    N, H_in, W_in, H_out, W_out = _geometry(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                                            pad_w, dilation_h, dilation_w, group, group_channels)
    input, offset, mask, grad_output = (t.contiguous() for t in (input, offset, mask, grad_output))
    g_in = torch.zeros(input.shape, dtype=torch.float32, device=input.device)
    g_off = torch.empty(offset.shape, dtype=torch.float32, device=input.device)
    g_mask = torch.empty(mask.shape, dtype=torch.float32, device=input.device)
    scratch = _scratch(input, N, H_out, W_out, group, kernel_h, kernel_w)
    with torch.cuda.device(input.device):
        _lib.call('dcnv3_bwd', input.data_ptr(), offset.data_ptr(), mask.data_ptr(), grad_output.data_ptr(),
                  g_in.data_ptr(), g_off.data_ptr(), g_mask.data_ptr(), scratch.data_ptr(), N, H_in, W_in, H_out,
                  W_out, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                  group_channels, float(offset_scale), _DTYPE_CODE[input.dtype], _stream_ptr(input))
    dt = input.dtype
    return g_in.to(dt), g_off.to(dt), g_mask.to(dt)


# what `import DCNv3` gives the reference (functions/dcnv3_func.py:16)
ext_module = types.SimpleNamespace(dcnv3_forward=dcnv3_forward, dcnv3_backward=dcnv3_backward)


class DCNv3Function(Function):
    """Argument list and return convention of the reference's ``DCNv3Function`` (dcnv3_func.py:19-62)."""

    @staticmethod
    @custom_fwd(cast_inputs=None)
    def forward(ctx, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                dilation_w, group, group_channels, offset_scale, im2col_step):
        ctx.cfg = (kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                   group_channels, offset_scale)
        ctx.im2col_step = im2col_step
        output = ext_module.dcnv3_forward(input, offset, mask, *ctx.cfg, im2col_step)
        ctx.save_for_backward(input, offset, mask)
        return output

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, grad_output):
        input, offset, mask = ctx.saved_tensors
        grads = ext_module.dcnv3_backward(input, offset, mask, *ctx.cfg, grad_output.contiguous(), ctx.im2col_step)
        return (*grads, None, None, None, None, None, None, None, None, None, None, None, None)


class to_channels_first(nn.Module):
    def forward(self, x):
        return x.permute(0, 3, 1, 2)


class to_channels_last(nn.Module):
    def forward(self, x):
        return x.permute(0, 2, 3, 1)


def build_norm_layer(dim, norm_layer, in_format='channels_last', out_format='channels_last', eps=1e-6):
    layers = []
    if norm_layer == 'BN':
        if in_format == 'channels_last':
            layers.append(to_channels_first())
        layers.append(nn.BatchNorm2d(dim))
        if out_format == 'channels_last':
            layers.append(to_channels_last())
    elif norm_layer == 'LN':
        if in_format == 'channels_first':
            layers.append(to_channels_last())
        layers.append(nn.LayerNorm(dim, eps=eps))
        if out_format == 'channels_first':
            layers.append(to_channels_first())
    else:
        raise NotImplementedError(f'build_norm_layer does not support {norm_layer}')
    return nn.Sequential(*layers)


def build_act_layer(act_layer):
    if act_layer == 'ReLU':
        return nn.ReLU(inplace=True)
    if act_layer == 'SiLU':
        return nn.SiLU(inplace=True)
    if act_layer == 'GELU':
        return nn.GELU()
    raise NotImplementedError(f'build_act_layer does not support {act_layer}')


class DCNv3(nn.Module):
    """The reference's ``DCNv3`` module (modules/dcnv3.py:216-345): depth-wise conv + norm + activation feed the
    offset / mask Linears, the mask is soft-maxed over the kernel points of a group, the sampling core is
    :class:`DCNv3Function`, optionally blended with the projected input by a per-group centre-feature scale."""

    def __init__(self, channels=64, kernel_size=3, dw_kernel_size=None, stride=1, pad=1, dilation=1, group=4,
                 offset_scale=1.0, act_layer='GELU', norm_layer='LN', center_feature_scale=False):
        super().__init__()
        if channels % group != 0:
            raise ValueError(f'channels must be divisible by group, but got {channels} and {group}')
        per_group = channels // group
        dw_kernel_size = dw_kernel_size if dw_kernel_size is not None else kernel_size
        if per_group & (per_group - 1):
            warnings.warn('DCNv3: a power-of-two number of channels per group is more efficient (16-byte channel lanes)')
        self.offset_scale = offset_scale
        self.channels = channels
        self.kernel_size = kernel_size
        self.dw_kernel_size = dw_kernel_size
        self.stride = stride
        self.dilation = dilation
        self.pad = pad
        self.group = group
        self.group_channels = per_group
        self.center_feature_scale = center_feature_scale
        self.dw_conv = nn.Sequential(
            nn.Conv2d(channels, channels, kernel_size=dw_kernel_size, stride=1, padding=(dw_kernel_size - 1) // 2,
                      groups=channels),
            build_norm_layer(channels, norm_layer, 'channels_first', 'channels_last'),
            build_act_layer(act_layer))
        self.offset = nn.Linear(channels, group * kernel_size * kernel_size * 2)
        self.mask = nn.Linear(channels, group * kernel_size * kernel_size)
        self.input_proj = nn.Linear(channels, channels)
        self.output_proj = nn.Linear(channels, channels)
        self._reset_parameters()
        if center_feature_scale:
            self.center_feature_scale_proj_weight = nn.Parameter(torch.zeros((group, channels), dtype=torch.float))
            self.center_feature_scale_proj_bias = nn.Parameter(torch.zeros(group, dtype=torch.float))

    def _reset_parameters(self):
        constant_(self.offset.weight.data, 0.)
        constant_(self.offset.bias.data, 0.)
        constant_(self.mask.weight.data, 0.)
        constant_(self.mask.bias.data, 0.)
        xavier_uniform_(self.input_proj.weight.data)
        constant_(self.input_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    def forward(self, input):
        """input, output: (N, H, W, C)."""
        N, H, W, _ = input.shape
        x = self.input_proj(input)
        x_proj = x
        dtype = x.dtype
        x1 = self.dw_conv(input.permute(0, 3, 1, 2))
        offset = self.offset(x1)
        mask = self.mask(x1).reshape(N, H, W, self.group, -1)
        mask = F.softmax(mask, -1).reshape(N, H, W, -1).type(dtype)
        x = DCNv3Function.apply(x, offset, mask, self.kernel_size, self.kernel_size, self.stride, self.stride,
                                self.pad, self.pad, self.dilation, self.dilation, self.group, self.group_channels,
                                self.offset_scale, 256)
        if self.center_feature_scale:
            scale = F.linear(x1, self.center_feature_scale_proj_weight, self.center_feature_scale_proj_bias).sigmoid()
            scale = scale[..., None].repeat(1, 1, 1, 1, self.channels // self.group).flatten(-2)
            x = x * (1 - scale) + x_proj * scale
        return self.output_proj(x)
