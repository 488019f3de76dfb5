"""B200-native ``SpatialCrossAttention`` / ``MSDeformableAttention3D``.

Drop-in for ``projects/mmdet3d_plugin/bevformer/modules/spatial_cross_attention.py``: same
registry names, constructor arguments, forward keyword convention, parameter names and
initialisation, and the same observable results -- including the reference's quirks (camera
hit lists taken from batch element 0 while the divisor is per sample, ``attn_logits_clamp``
stored but never applied, point index ``p = k*D + z``).

What changes is how the work is done.  The reference gathers, per camera, the queries that
see it (``nonzero()`` + a host sync, :135-139), copies them into a zero-padded
``(bs, num_cam, max_len, C)`` batch (:142-151), runs the offset / weight Linear layers on every
copy, materialises ``sampling_locations`` (:361-376), calls the op per padded row and scatters
the result back (:163-170).  Here the Linear layers run once on the ``HW`` BEV queries, and a
single fused kernel (C ABI ``sca_fwd`` / ``sca_bwd``) does camera gating, softmax, Z-anchor
locations, sampling, the sum over cameras and the division by the hit count.
"""
import torch
import torch.nn as nn

from ..fused_ops import SpatialCrossAttnFunction, hit_bits_from_mask
from ..registry import ATTENTION, BaseModule, build_attention, xavier_init
from ..rowops import Junction, Linear, linear_add_layernorm
from .deform_common import DeformAttnBase, finish_block, msda_apply


@ATTENTION.register_module()
class MSDeformableAttention3D(DeformAttnBase):
    """Deformable attention over ``num_Z_anchors`` reference points per query
    (reference :175-403).  ``output_proj`` is ``None`` as in the reference (:220)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=8, im2col_step=64,
                 dropout=0.1, batch_first=True, norm_cfg=None, init_cfg=None,
                 attn_logits_clamp=None, debug_attn_nan=False):
        super().__init__(init_cfg)
        self._setup(embed_dims, num_heads, num_levels, num_points, im2col_step, batch_first,
                    norm_cfg, attn_logits_clamp, debug_attn_nan, queue=1, with_output_proj=False)
        self.init_weights()

    def project(self, query, value):
        """The three Linear layers: value (B, Nk, M, Dh), raw offsets (B, Nq, M, L, P, 2) and raw
        attention logits (B, Nq, M, L*P).  ``attn_logits_clamp`` is NOT applied (reference :347)."""
        B, Nq, _ = query.shape
        M, L, P = self.num_heads, self.num_levels, self.num_points
        v = self.value_proj(value)
        off = self.sampling_offsets(query).view(B, Nq, M, L, P, 2)
        logits = self.attention_weights(query).view(B, Nq, M, L * P)
        return v, off, logits

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, **kwargs):
        """Stand-alone call with the reference's contract (:277-403): query (bs, Nq, C), value
        (bs, Nk, C), reference_points (bs, Nq, D, 2) -> (bs, Nq, C), no residual, no output
        projection.  Runs the op-boundary kernel on materialised sampling locations."""
        if value is None:
            value = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, num_query, _ = query.shape
        _, num_value, _ = value.shape
        M, L, P = self.num_heads, self.num_levels, self.num_points
        v, off, logits = self.project(query, value)
        if key_padding_mask is not None:
            v = v.masked_fill(key_padding_mask[..., None], 0.0)
        v = v.view(bs, num_value, M, -1)
        attn = logits.softmax(-1).view(bs, num_query, M, L, P)
        if reference_points.shape[-1] != 2:
            raise ValueError('Last dim of reference_points must be 2, '
                             f'but get {reference_points.shape[-1]} instead.')
        normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        D = reference_points.shape[2]
        assert P % D == 0
        off = (off / normalizer[None, None, None, :, None, :]).view(bs, num_query, M, L, P // D, D, 2)
        loc = (reference_points[:, :, None, None, None, :, :] + off).view(bs, num_query, M, L, P, 2)
        out = msda_apply(v, spatial_shapes, level_start_index, loc, attn, self.im2col_step)
        if not self.batch_first:
            out = out.permute(1, 0, 2)
        return out


@ATTENTION.register_module()
class SpatialCrossAttention(BaseModule):
    """Spatial cross-attention of BEVFormer (reference :28-173)."""

    def __init__(self, embed_dims=256, num_cams=6, pc_range=None, dropout=0.1, init_cfg=None,
                 batch_first=False,
                 deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=256,
                                           num_levels=4),
                 **kwargs):
        super().__init__(init_cfg)
        self.init_cfg = init_cfg
        self.dropout = nn.Dropout(dropout)
        self.pc_range = pc_range
        self.fp16_enabled = False
        self.deformable_attention = build_attention(deformable_attention)
        self.embed_dims = embed_dims
        self.num_cams = num_cams
        self.output_proj = Linear(embed_dims, embed_dims)
        self.batch_first = batch_first
        self.init_weight()

    def init_weight(self):
        xavier_init(self.output_proj, distribution='uniform', bias=0.)

    @staticmethod
    def _grid_w(bev_h, bev_w, num_query):
        """Width of the BEV grid when the caller says the queries are its row-major cells (lets
        the kernel tile the work in 2-D patches); 0 when unknown."""
        if bev_h and bev_w and int(bev_h) * int(bev_w) == num_query:
            return int(bev_w)
        return 0

    def forward(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, reference_points_cam=None,
                bev_mask=None, level_start_index=None, flag='encoder', bev_geometry=None,
                bev_h=None, bev_w=None, post_norm=None, **kwargs):
        """query (bs, HW, C); key = value (num_cam, Nk, bs, C); reference_points_cam
        (num_cam, bs, HW, D, 2); bev_mask (num_cam, bs, HW, D) -> (bs, HW, C).

        ``bev_geometry`` (optional, from :func:`fused_ops.bev_point_sampling`) carries the
        camera-hit bit field already on the device; without it the field is derived from
        ``bev_mask`` with a few tensor ops (still no host sync).
        """
        if key is None:
            key = query
        if value is None:
            value = key
        inp_residual = query if residual is None else residual
        if query_pos is not None:
            query = query + query_pos
        bs, num_query, _ = query.size()
        da = self.deformable_attention
        if not isinstance(da, MSDeformableAttention3D):
            raise TypeError('the fused spatial cross-attention needs an MSDeformableAttention3D')
        lists = None
        if bev_geometry is not None:
            hit_bits, mask_u8 = bev_geometry.hit_bits, bev_geometry.mask_u8
            reference_points_cam = bev_geometry.reference_points_cam
            if torch.is_grad_enabled():
                lists = bev_geometry.lists()       # the backward's tensor-core pass walks them
        else:
            mask_b = bev_mask.to(torch.bool)
            hit_bits = hit_bits_from_mask(mask_b)
            mask_u8 = mask_b.contiguous().view(torch.uint8)
        num_cams, l, bs_v, _ = value.shape
        projected = kwargs.get('_sca_projected_value')
        if projected is not None:                  # hoisted in front of the layer loop (encoder._hoisted_sca_values)
            v = projected.view(bs * self.num_cams, l, da.num_heads, -1)
        else:
            value = value.permute(2, 0, 1, 3).reshape(bs * self.num_cams, l, self.embed_dims)
            v = da.value_proj(value).view(bs * self.num_cams, l, da.num_heads, -1)
        # the query is also the block's residual: its two gradients meet in the projection's dX GEMM
        tok = Junction() if (post_norm is not None and query is inp_residual and torch.is_grad_enabled()) else None
        coords = da.project_coords(query, tok)     # offsets | logits of a query, one GEMM
        slots = SpatialCrossAttnFunction.apply(v, spatial_shapes, level_start_index, coords, None,
                                               reference_points_cam, mask_u8, hit_bits,
                                               self.num_cams, self._grid_w(bev_h, bev_w, num_query), lists)
        # (the residual is added without a permute whatever batch_first says, reference :171-173)
        if post_norm is not None:
            return linear_add_layernorm(slots.to(query.dtype), self.output_proj, inp_residual, post_norm,
                                        p=self.dropout.p if self.training else 0.0, junction=tok)
        out = self.dropout(self.output_proj(slots.to(query.dtype))) + inp_residual
        return out if post_norm is None else post_norm(out)
