"""B200-native ``TemporalSelfAttention``.

Drop-in for ``projects/mmdet3d_plugin/bevformer/modules/temporal_self_attention.py`` (:24-289):
same registry name, constructor, forward keywords, parameter names and results (softmax per
queue entry, ``value[:bs]`` slice semantics, first-frame value duplication, optional logit
clamp).  The reference permutes offsets / weights into a ``bs*2`` batch, materialises
``sampling_locations``, calls the op and averages the two queue outputs with three more
permutes (:234-279); here one fused kernel (C ABI ``tsa_fwd`` / ``tsa_bwd``) consumes the raw
Linear outputs and writes the queue mean.
"""
import torch
import torch.nn as nn

from ..fused_ops import QueueDeformAttnFunction
from ..registry import ATTENTION
from ..rowops import Junction, paired_query_linear
from .deform_common import DeformAttnBase, finish_block, msda_apply


@ATTENTION.register_module()
class TemporalSelfAttention(DeformAttnBase):

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, num_bev_queue=2,
                 im2col_step=64, dropout=0.1, batch_first=True, norm_cfg=None, init_cfg=None,
                 attn_logits_clamp=None, debug_attn_nan=False):
        super().__init__(init_cfg)
        self.num_bev_queue = num_bev_queue
        self.dropout = nn.Dropout(dropout)
        self._setup(embed_dims, num_heads, num_levels, num_points, im2col_step, batch_first,
                    norm_cfg, attn_logits_clamp, debug_attn_nan, queue=num_bev_queue,
                    with_output_proj=True)
        self.init_weights()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, flag='decoder', bev_h=None, bev_w=None, row_slice=None,
                post_norm=None, **kwargs):
        """query (bs, HW, C) [batch_first]; value (bs*2, HW, C) = stack([prev_bev, bev], 1) or
        None (first frame); reference_points (bs*2, HW, L, 2) -> (bs, HW, C)."""
        if value is None:
            assert self.batch_first
            bs, len_bev, c = query.shape
            value = torch.stack([query, query], 1).reshape(bs * 2, len_bev, c)
        if identity is None:
            identity = query
        # batch-first callers keep the positional sum for the projection below, which writes it straight into
        # its concatenated input and meets the residual's gradient in its dX GEMM (rowops.Junction)
        raw_query, raw_pos = query, query_pos
        defer_pos = self.batch_first and query.is_cuda and torch.is_grad_enabled()
        if query_pos is not None and not defer_pos:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, num_query, embed_dims = query.shape
        _, num_value, _ = value.shape
        assert self.num_bev_queue == 2
        M, L, P, Q = self.num_heads, self.num_levels, self.num_points, self.num_bev_queue

        # value[:bs] pairs every query with the same cell of the first queue entry; under BEV row
        # sharding the queries are the slice ``row_slice`` of the value's cells
        paired = value[:bs] if row_slice is None else value[:bs, row_slice[0]:row_slice[1]]
        if defer_pos:
            tok = Junction() if (post_norm is not None and identity is raw_query) else None
            w, b = self.coords_weight()
            coords = paired_query_linear(paired, raw_query, raw_pos, w, b, tok)
        else:
            tok = None
            coords = self.project_coords(torch.cat([paired, query], -1))
        projected = kwargs.get('_tsa_projected_value')
        if projected is not None and projected.shape == value.shape:
            value = projected                      # hoisted in front of the layer loop (encoder._hoisted_tsa_values)
        else:
            value = self.value_proj(value)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.reshape(bs * Q, num_value, M, -1)
        n = M * Q * L * P

        grid_w = int(bev_w) if (bev_h and bev_w and int(bev_h) * int(bev_w) == num_query) else 0
        if reference_points.shape[-1] == 2:
            if reference_points.shape[2] != L:          # a single reference broadcast over levels
                reference_points = reference_points.expand(-1, -1, L, -1)
            output = QueueDeformAttnFunction.apply(value, spatial_shapes, level_start_index,
                                                   coords, None, reference_points,
                                                   self.attn_logits_clamp, grid_w)
        elif reference_points.shape[-1] == 4:
            offsets = coords[..., :2 * n].reshape(bs, num_query, M, Q, L, P, 2)
            logits = coords[..., 2 * n:].reshape(bs, num_query, M, Q, L * P)
            # box-shaped references (:246-250): rare path, run on the op boundary
            if self.attn_logits_clamp is not None:
                c = float(self.attn_logits_clamp)
                logits = logits.clamp(min=-c, max=c)
            attn = logits.softmax(-1).view(bs, num_query, M, Q, L, P)
            attn = attn.permute(0, 3, 1, 2, 4, 5).reshape(bs * Q, num_query, M, L, P).contiguous()
            off = offsets.permute(0, 3, 1, 2, 4, 5, 6).reshape(bs * Q, num_query, M, L, P, 2)
            loc = reference_points[:, :, None, :, None, :2] \
                + off / P * reference_points[:, :, None, :, None, 2:] * 0.5
            out = msda_apply(value, spatial_shapes, level_start_index, loc, attn, self.im2col_step)
            output = out.view(bs, Q, num_query, embed_dims).mean(1)
        else:
            raise ValueError('Last dim of reference_points must be 2 or 4, '
                             f'but get {reference_points.shape[-1]} instead.')

        return finish_block(self, output.to(query.dtype), identity, post_norm, junction=tok)
