"""B200-native ``CustomMSDeformableAttention`` -- the decoder-side deformable cross-attention
used by the detection decoder and by the MapTR / MapTRv2 decoders.

Drop-in for ``projects/mmdet3d_plugin/bevformer/modules/decoder.py:129-358``: same registry
name, constructor, forward keywords (``batch_first=False`` by default: query (Nq, bs, C), value
(HW, bs, C), reference_points (bs, Nq, L, 2)), parameter names, clamp + softmax, residual.
The sampling core is the fused queue kernel with a queue of one (C ABI ``tsa_fwd`` / ``tsa_bwd``).
"""
import torch
import torch.nn as nn

from torch.autograd.function import once_differentiable

from ..fused_ops import QueueDeformAttnFunction
from ..multi_scale_deformable_attn_function import custom_bwd, custom_fwd
from ..registry import ATTENTION
from ..rowops import GradSlots, _take_bias_grad, register_grad_slot, weight_bias_grad
from .deform_common import DeformAttnBase, finish_block, msda_apply


def inverse_sigmoid(x, eps=1e-5):
    """log(x / (1 - x)) with clamping (decoder.py:32-47)."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


class HoistedValueProjFunction(torch.autograd.Function):
    """``[x W_l^T + b_l for l in layers]`` before the layer loop, as one autograd node.

    Written as plain tensor ops (stack, baddbmm, ``out[l]``) the backward pays, per layer, a zero-filled
    (layers, rows, C) tensor for the select, a copy into it and an accumulation add of the whole stack
    (6 x 123 MB each way at the 200x200 BEV: ~0.6 ms per decoder step, ``profiles/r02_decoder.md``).  Here every
    layer's gradient is consumed where it arrives: dW_l = g_l^T x and db_l (the column sums the attention's
    last pass already produced, when offered) per layer, and dx accumulated by GEMMs with beta = 1."""

    @staticmethod
    @custom_fwd(cast_inputs=None)
    def forward(ctx, x, side_by_side, *params):
        """``side_by_side``: offer the consumers of the outputs column blocks of one (rows, n * C) matrix to write
        the outputs' gradients into (``rowops.GradSlots``; the fused attention backward takes them): the input
        gradient is then ONE GEMM over K = n * C and the weight gradients ONE GEMM with n * C output rows."""
        n = len(params) // 2
        weights, biases = params[:n], params[n:]
        # one GEMM per layer with the bias in its epilogue, back to back on one input that stays in L2 (a
        # batched GEMM would first materialise the broadcast bias as a (layers, rows, C) tensor: 108 us at 200x200)
        out = torch.empty((n, x.shape[0], weights[0].shape[0]), dtype=x.dtype, device=x.device)
        for i in range(n):
            torch.addmm(biases[i], x, weights[i].t(), out=out[i])
        ctx.save_for_backward(x, *weights)
        ctx.slots = None
        if side_by_side and x.is_cuda:
            ctx.slots = GradSlots(n, x.shape[0], weights[0].shape[0], x.dtype, x.device)
            for i in range(n):
                register_grad_slot(out[i], ctx.slots, i)
        return tuple(out[i] for i in range(n))

    @staticmethod
    @once_differentiable
    @custom_bwd
    def backward(ctx, *grads):
        x, *weights = ctx.saved_tensors
        n = len(weights)
        dx = None
        dws, dbs = [None] * n, [None] * n
        slots, ctx.slots = ctx.slots, None
        C = weights[0].shape[0]
        if (slots is not None and slots.buffer is not None and all(slots.filled) and
                all(g is not None and g.dim() == 2 and g.untyped_storage().data_ptr() ==
                    slots.buffer.untyped_storage().data_ptr() and g.storage_offset() == i * C and
                    g.stride() == slots.buffer.stride() and g.shape == (slots.rows, C)
                    for i, g in enumerate(grads))):
            # every gradient sits in its column block of the common matrix G (rows, n * C):
            # dx = G [W_0; ...; W_{n-1}] and [dW_0; ...; dW_{n-1}] = G^T x, one GEMM each
            G = slots.buffer
            if ctx.needs_input_grad[0]:
                dx = G @ torch.cat(list(weights), 0)
            if any(ctx.needs_input_grad[2 + i] for i in range(n)):
                dW = G.t() @ x
                for i in range(n):
                    if ctx.needs_input_grad[2 + i]:
                        dws[i] = dW[i * C:(i + 1) * C]
            for i, g in enumerate(grads):
                if ctx.needs_input_grad[2 + n + i]:
                    ready = _take_bias_grad(g, C)
                    dbs[i] = ready.to(weights[i].dtype) if ready is not None else g.sum(0)
            return (dx, None, *dws, *dbs)
        for i, (g, w) in enumerate(zip(grads, weights)):
            if g is None:
                continue
            g2 = g.reshape(-1, w.shape[0])
            if ctx.needs_input_grad[0]:
                dx = g2 @ w if dx is None else dx.addmm_(g2, w)
            want_w, want_b = ctx.needs_input_grad[2 + i], ctx.needs_input_grad[2 + n + i]
            ready = _take_bias_grad(g, w.shape[0]) if want_b else None
            if ready is not None:
                dbs[i] = ready.to(w.dtype)
                want_b = False
            if want_w:
                dws[i], db = weight_bias_grad(g2, x, w, want_bias=want_b)
                dbs[i] = db if want_b else dbs[i]
            elif want_b:
                dbs[i] = g2.sum(0)
        return (dx, None, *dws, *dbs)


def hoist_value_proj(attns, value, batch_first=False):
    """``[a.value_proj(value) for a in attns]`` hoisted in front of the layer loop.

    Every decoder layer projects the SAME BEV map with its own ``value_proj`` (decoder.py:299-303
    of the reference, once per layer inside the layer loop); the BEV does not change between the
    layers, so the projections of all layers can be computed before the loop, with a
    backward that accumulates the BEV gradient with beta = 1 GEMMs instead of gradient-accumulation adds
    (:class:`HoistedValueProjFunction`).  ``value`` is (HW, bs, C) (``batch_first=False``) or (bs, HW, C);
    returns a list of batch-major (bs, HW, C) tensors, one per attention module."""
    if not batch_first:
        value = value.permute(1, 0, 2)
    bs, n, C = value.shape
    outs = HoistedValueProjFunction.apply(value.reshape(bs * n, C), False, *[a.value_proj.weight for a in attns],
                                          *[a.value_proj.bias for a in attns])
    return [o.view(bs, n, -1) for o in outs]


def hoisted_projections(layers, args, kwargs):
    """Per-layer ``projected_value`` tensors for a decoder's layer loop, or None when the layers do
    not each hold exactly one seq-first ``CustomMSDeformableAttention`` with identically shaped
    ``value_proj`` parameters (then every layer projects its own value, as the reference does)."""
    value = kwargs.get('value', args[1] if len(args) > 1 else None)
    if len(layers) < 2 or not torch.is_tensor(value) or value.dim() != 3:
        return None
    found = []
    for layer in layers:
        mods = [a for a in getattr(layer, 'attentions', []) if isinstance(a, CustomMSDeformableAttention)]
        if len(mods) != 1 or mods[0].batch_first or mods[0].value_proj.bias is None:
            return None
        found.append(mods[0])
    w0 = found[0].value_proj.weight
    if any(a.value_proj.weight.shape != w0.shape for a in found) or value.dtype != w0.dtype:
        return None
    return hoist_value_proj(found, value)


@ATTENTION.register_module()
class CustomMSDeformableAttention(DeformAttnBase):
    accepts_post_norm = True        # forward(post_norm=LayerNorm): the layer's next norm runs inside the block

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64,
                 dropout=0.1, batch_first=False, norm_cfg=None, init_cfg=None,
                 attn_logits_clamp=None, debug_attn_nan=False):
        super().__init__(init_cfg)
        self.dropout = nn.Dropout(dropout)
        self._setup(embed_dims, num_heads, num_levels, num_points, im2col_step, batch_first,
                    norm_cfg, attn_logits_clamp, debug_attn_nan, queue=1, with_output_proj=True)
        self.init_weights()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, flag='decoder', post_norm=None, projected_value=None,
                **kwargs):
        if 'residual' in kwargs and identity is None:      # mmcv's deprecated_api_warning alias
            identity = kwargs.pop('residual')
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, num_query, _ = query.shape
        _, num_value, _ = value.shape
        M, L, P = self.num_heads, self.num_levels, self.num_points

        if projected_value is not None:
            # value_proj(value) of this layer, computed by the caller for all decoder layers in one
            # batched GEMM (hoist_value_proj below); batch-major (bs, num_value, C)
            assert projected_value.shape == value.shape, (projected_value.shape, value.shape)
            value = projected_value
        else:
            value = self.value_proj(value)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(bs, num_value, M, -1)
        coords = self.project_coords(query)             # offsets | logits of a query, one GEMM
        n = M * L * P

        if reference_points.shape[-1] == 2:
            if reference_points.shape[2] != L:          # a single reference broadcast over levels
                reference_points = reference_points.expand(-1, -1, L, -1)
            output = QueueDeformAttnFunction.apply(value, spatial_shapes, level_start_index,
                                                   coords, None, reference_points,
                                                   self.attn_logits_clamp, 0)
        elif reference_points.shape[-1] == 4:
            offsets = coords[..., :2 * n].reshape(bs, num_query, M, L, P, 2)
            logits = coords[..., 2 * n:].reshape(bs, num_query, M, L * P)
            if self.attn_logits_clamp is not None:
                c = float(self.attn_logits_clamp)
                logits = logits.clamp(min=-c, max=c)
            attn = logits.softmax(-1).view(bs, num_query, M, L, P)
            loc = reference_points[:, :, None, :, None, :2] \
                + offsets / P * reference_points[:, :, None, :, None, 2:] * 0.5
            output = msda_apply(value, spatial_shapes, level_start_index, loc, attn,
                                self.im2col_step)
        else:
            raise ValueError('Last dim of reference_points must be 2 or 4, '
                             f'but get {reference_points.shape[-1]} instead.')

        return finish_block(self, output.to(query.dtype), identity, post_norm)
