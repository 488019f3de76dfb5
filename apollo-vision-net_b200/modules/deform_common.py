"""Pieces shared by the three deformable-attention modules: parameter layout, the reference's
initialisation, and the op-boundary (materialised locations) execution path."""
import math
import warnings

import torch
import torch.nn as nn

from ..multi_scale_deformable_attn_function import (MultiScaleDeformableAttnFunction_bf16,
                                                    MultiScaleDeformableAttnFunction_fp16,
                                                    MultiScaleDeformableAttnFunction_fp32)
from ..registry import BaseModule, constant_init, xavier_init
from ..rowops import Linear, linear, linear_add_layernorm


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError('invalid input for _is_power_of_2: {} (type: {})'.format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


def ring_bias(num_heads, reps, num_points):
    """Initial ``sampling_offsets.bias``: head h looks along angle 2*pi*h/M, point i at i+1 px
    (spatial_cross_attention.py:259-271, temporal_self_attention.py:115-128, decoder.py:215-226)."""
    theta = torch.arange(num_heads, dtype=torch.float32) * (2.0 * math.pi / num_heads)
    grid = torch.stack([theta.cos(), theta.sin()], -1)
    grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(num_heads, 1, 1, 2)
    grid = grid.repeat(1, reps, num_points, 1)
    for i in range(num_points):
        grid[:, :, i, :] *= i + 1
    return grid.view(-1)


class DeformAttnBase(BaseModule):
    """Holds ``sampling_offsets`` / ``attention_weights`` / ``value_proj`` (and optionally
    ``output_proj``) with the reference's names, so its checkpoints load unchanged
    (SURVEY.md appendix A)."""

    def _setup(self, embed_dims, num_heads, num_levels, num_points, im2col_step, batch_first,
               norm_cfg, attn_logits_clamp, debug_attn_nan, queue=1, with_output_proj=True):
        if embed_dims % num_heads != 0:
            raise ValueError(f'embed_dims must be divisible by num_heads, '
                             f'but got {embed_dims} and {num_heads}')
        if not _is_power_of_2(embed_dims // num_heads):
            warnings.warn("You'd better set embed_dims in MultiScaleDeformAttention to make the "
                          'dimension of each attention head a power of 2 which is more efficient '
                          'in our CUDA implementation.')
        self.norm_cfg = norm_cfg
        self.batch_first = batch_first
        self.fp16_enabled = False
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self.attn_logits_clamp = attn_logits_clamp
        self.debug_attn_nan = bool(debug_attn_nan)
        self._queue = queue
        self.sampling_offsets = Linear(embed_dims * queue,
                                          queue * num_heads * num_levels * num_points * 2)
        self.attention_weights = Linear(embed_dims * queue,
                                           queue * num_heads * num_levels * num_points)
        self.value_proj = Linear(embed_dims, embed_dims)
        self.output_proj = Linear(embed_dims, embed_dims) if with_output_proj else None

    def init_weights(self):
        constant_init(self.sampling_offsets, 0.)
        self.sampling_offsets.bias.data = ring_bias(
            self.num_heads, self.num_levels * self._queue, self.num_points).to(
                self.sampling_offsets.bias.device, self.sampling_offsets.bias.dtype)
        constant_init(self.attention_weights, val=0., bias=0.)
        xavier_init(self.value_proj, distribution='uniform', bias=0.)
        xavier_init(self.output_proj, distribution='uniform', bias=0.)
        self._is_init = True


    def project_coords(self, query, junction=None):
        """``sampling_offsets(query)`` and ``attention_weights(query)`` as ONE GEMM over the
        concatenated weights: returns (..., 3n) with the raw offsets
        of a query in columns [0, 2n) and its raw attention logits in [2n, 3n).  The parameters stay
        the reference's two Linear layers (checkpoint layout unchanged); the fused kernels read both
        column blocks in place and return one gradient tensor of the same layout, so the backward
        is one dX GEMM, one dW GEMM and one bias reduction instead of two of each.  ``junction``: the query is
        also the residual of the block (:class:`rowops.Junction`)."""
        w, b = self.coords_weight()
        return linear(query, w, b, junction)

    def coords_weight(self):
        """(weight, bias) of the merged offsets | logits projection."""
        w = torch.cat([self.sampling_offsets.weight, self.attention_weights.weight], 0)
        b = torch.cat([self.sampling_offsets.bias, self.attention_weights.bias], 0)
        return w, b


def finish_block(mod, output, identity, post_norm=None, junction=None):
    """Tail of an attention block: ``dropout(output_proj(output)) + identity`` (reference
    temporal_self_attention.py:285-289, spatial_cross_attention.py:171-173, decoder.py:353-358),
    optionally followed by the layer's next LayerNorm (``post_norm``).  With the norm given the
    projection, the dropout (mask drawn in the kernel and recomputed in the backward), the residual add
    and the norm run as one fused autograd node (:func:`rowops.linear_add_layernorm`); the result is then
    already normalised."""
    batch_first = getattr(mod, 'batch_first', True)
    drop = mod.dropout
    if post_norm is not None and batch_first:
        return linear_add_layernorm(output, mod.output_proj, identity, post_norm,
                                    p=drop.p if mod.training else 0.0, junction=junction)
    if post_norm is not None and output.dim() == 3 and output.shape[0] == 1:
        # sequence-first caller with one sample: (1, Nq, C) and (Nq, 1, C) hold the same rows in the same order
        return linear_add_layernorm(output.permute(1, 0, 2), mod.output_proj, identity, post_norm,
                                    p=drop.p if mod.training else 0.0)
    output = mod.output_proj(output)
    if not batch_first:
        output = output.permute(1, 0, 2)
    out = drop(output) + identity
    return out if post_norm is None else post_norm(out)


def msda_apply(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
               im2col_step):
    """Op-boundary call, as the reference selects it (spatial_cross_attention.py:389-396): fp32
    tensors use the ``_fp32`` Function, an fp16 value the ``_fp16`` one and a bf16 value the
    ``_bf16`` one (whose autocast decorator keeps bf16 instead of casting to fp16)."""
    if not value.is_cuda:
        raise RuntimeError('deformable attention needs CUDA tensors: this build has no CPU path '
                           '(the reference falls back to multi_scale_deformable_attn_pytorch)')
    fn = {torch.float32: MultiScaleDeformableAttnFunction_fp32,
          torch.bfloat16: MultiScaleDeformableAttnFunction_bf16}.get(value.dtype,
                                                                     MultiScaleDeformableAttnFunction_fp16)
    return fn.apply(value, spatial_shapes, level_start_index, sampling_locations,
                    attention_weights, im2col_step)
