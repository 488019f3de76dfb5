"""B200-native ``CustomMSDeformableAttention`` -- the decoder-side deformable cross-attention
used by the detection decoder and by the MapTR / MapTRv2 decoders.

Drop-in for ``projects/mmdet3d_plugin/bevformer/modules/decoder.py:129-358``: same registry
name, constructor, forward keywords (``batch_first=False`` by default: query (Nq, bs, C), value
(HW, bs, C), reference_points (bs, Nq, L, 2)), parameter names, clamp + softmax, residual.
The sampling core is the fused queue kernel with a queue of one (C ABI ``tsa_fwd`` / ``tsa_bwd``).
"""
import torch
import torch.nn as nn

from ..fused_ops import QueueDeformAttnFunction
from ..registry import ATTENTION
from .deform_common import DeformAttnBase, finish_block, msda_apply


def inverse_sigmoid(x, eps=1e-5):
    """log(x / (1 - x)) with clamping (decoder.py:32-47)."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


@ATTENTION.register_module()
class CustomMSDeformableAttention(DeformAttnBase):

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
                level_start_index=None, flag='decoder', post_norm=None, **kwargs):
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
