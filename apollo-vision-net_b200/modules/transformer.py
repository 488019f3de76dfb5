"""``PerceptionTransformer`` (reference: projects/mmdet3d_plugin/bevformer/modules/transformer.py):
the caller of the BEV encoder -- CAN-bus shift, rotation of the previous BEV, CAN-bus embedding,
camera / level embeddings and flattening of the multi-level image features
(``get_bev_features``, :119-298), then the decoder call (``forward``, :300-401).

Same constructor keywords, parameter names (``level_embeds``, ``cams_embeds``, ``can_bus_mlp.*``,
``reference_points.*``, ``encoder.*``, ``decoder.*``) and call contract as the reference.  What
differs is how the pre-processing runs: one rotation launch for all
samples instead of a Python loop of torchvision calls, one flatten launch per level, and no
``torch.isfinite(x).all()`` guards (each is a device -> host synchronisation): non-finite image
features are zeroed inside the flatten kernel, unconditionally, which is what the guards do when
they fire.  The reference's NaN diagnostics / re-initialisation of non-finite embedding parameters
(:131-149, :216-229) are debugging aids and are not reproduced.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.nn.init import normal_

from ..bev_prep import can_bus_shift, flatten_features, rotate_prev_bev
from ..registry import TRANSFORMER, BaseModule, build_transformer_layer_sequence, xavier_init
from ..rowops import LayerNorm, Linear
from .decoder import CustomMSDeformableAttention
from .spatial_cross_attention import MSDeformableAttention3D
from .temporal_self_attention import TemporalSelfAttention


@TRANSFORMER.register_module()
class PerceptionTransformer(BaseModule):
    def __init__(self, num_feature_levels=4, num_cams=6, two_stage_num_proposals=300, encoder=None,
                 decoder=None, embed_dims=256, rotate_prev_bev=True, use_shift=True, use_can_bus=True,
                 can_bus_norm=True, can_bus_in_dataset=True, use_cams_embeds=True,
                 rotate_center=[100, 100], **kwargs):
        super().__init__(**kwargs)
        self.encoder = build_transformer_layer_sequence(encoder)
        self.decoder = build_transformer_layer_sequence(decoder) if decoder is not None else None
        self.embed_dims = embed_dims
        self.num_feature_levels = num_feature_levels
        self.num_cams = num_cams
        self.fp16_enabled = False
        self.rotate_prev_bev = rotate_prev_bev
        self.use_shift = use_shift
        self.use_can_bus = use_can_bus
        self.can_bus_norm = can_bus_norm
        self.can_bus_in_dataset = can_bus_in_dataset
        self.use_cams_embeds = use_cams_embeds
        self.two_stage_num_proposals = two_stage_num_proposals
        self.init_layers()
        self.rotate_center = rotate_center

    def init_layers(self):
        self.level_embeds = nn.Parameter(torch.Tensor(self.num_feature_levels, self.embed_dims))
        if self.use_cams_embeds:
            self.cams_embeds = nn.Parameter(torch.Tensor(self.num_cams, self.embed_dims))
        if self.use_can_bus:
            self.can_bus_mlp = nn.Sequential(
                Linear(18, self.embed_dims // 2), nn.ReLU(inplace=True),
                Linear(self.embed_dims // 2, self.embed_dims), nn.ReLU(inplace=True))
            if self.can_bus_norm:
                self.can_bus_mlp.add_module('norm', LayerNorm(self.embed_dims))
        if self.decoder is not None:
            self.reference_points = Linear(self.embed_dims, 3)

    def init_weights(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, (MSDeformableAttention3D, TemporalSelfAttention, CustomMSDeformableAttention)):
                m.init_weights()
        normal_(self.level_embeds)
        if self.use_cams_embeds:
            normal_(self.cams_embeds)
        if self.use_can_bus:
            for m in self.can_bus_mlp:
                if isinstance(m, nn.Linear):
                    xavier_init(m, distribution='uniform', bias=0.)
        if self.decoder is not None:
            xavier_init(self.reference_points, distribution='uniform', bias=0.)

    # ------------------------------------------------------------------ pre-processing ------
    def prepare_bev_inputs(self, mlvl_feats, bev_queries, bev_h, bev_w, grid_length=(0.512, 0.512),
                           bev_pos=None, prev_bev=None, img_metas=None):
        """Everything ``get_bev_features`` computes before it calls the encoder; returns the
        encoder's positional and keyword arguments (the call contract of transformer.py:273-286)."""
        bs = mlvl_feats[0].size(0)
        bev_queries = bev_queries.unsqueeze(1).repeat(1, bs, 1)                    # (HW, bs, C)
        bev_pos = bev_pos.flatten(2).permute(2, 0, 1)

        if self.can_bus_in_dataset:
            can_bus = np.asarray([each['can_bus'] for each in img_metas], dtype=np.float64)
            shift_np = can_bus_shift(can_bus, grid_length, bev_h, bev_w, self.use_shift)
        else:
            can_bus = None
            shift_np = np.zeros((bs, 2)) * self.use_shift
        shift = bev_queries.new_tensor(shift_np)                                     # (bs, 2)

        if prev_bev is not None:
            if prev_bev.shape[1] == bev_h * bev_w:
                prev_bev = prev_bev.permute(1, 0, 2)                                 # (HW, bs, C)
            if self.rotate_prev_bev:
                angles = []
                for i in range(bs):
                    cb = img_metas[i].get('can_bus', None) if img_metas is not None else None
                    # reference quirk (:191-194): only list / tuple can_bus rotates; anything else
                    # (e.g. an ndarray) silently falls back to no rotation
                    angles.append(float(cb[-1]) if isinstance(cb, (list, tuple)) and len(cb) > 0 else 0.0)
                prev_bev = rotate_prev_bev(prev_bev, angles, bev_h, bev_w, self.rotate_center)

        if self.use_can_bus:
            cb = bev_queries.new_tensor(np.asarray([each['can_bus'] for each in img_metas], dtype=np.float64))
            bev_queries = bev_queries + self.can_bus_mlp(cb)[None, :, :] * self.use_can_bus

        feat_flatten, spatial_shapes, level_start_index = flatten_features(
            list(mlvl_feats), self.cams_embeds if self.use_cams_embeds else None, self.level_embeds)
        return (bev_queries, feat_flatten, feat_flatten), dict(
            bev_h=bev_h, bev_w=bev_w, bev_pos=bev_pos, spatial_shapes=spatial_shapes,
            level_start_index=level_start_index, prev_bev=prev_bev, shift=shift)

    def get_bev_features(self, mlvl_feats, bev_queries, bev_h, bev_w, grid_length=[0.512, 0.512],
                         bev_pos=None, prev_bev=None, **kwargs):
        args, enc_kwargs = self.prepare_bev_inputs(mlvl_feats, bev_queries, bev_h, bev_w, grid_length,
                                                   bev_pos, prev_bev, kwargs.get('img_metas'))
        bev_embed = self.encoder(*args, **enc_kwargs, **kwargs)
        if kwargs.get('return_intermediate'):
            return bev_embed, args[1], enc_kwargs['spatial_shapes'], enc_kwargs['level_start_index']
        return bev_embed

    # ------------------------------------------------------------------ decoder call --------
    def forward(self, mlvl_feats, bev_queries, object_query_embed, bev_h, bev_w,
                grid_length=[0.512, 0.512], bev_pos=None, reg_branches=None, cls_branches=None,
                prev_bev=None, **kwargs):
        res = self.get_bev_features(mlvl_feats, bev_queries, bev_h, bev_w, grid_length=grid_length,
                                    bev_pos=bev_pos, prev_bev=prev_bev, **kwargs)
        ret_inter = bool(kwargs.get('return_intermediate'))
        if ret_inter:
            bev_embed, feat_flatten, spatial_shapes, level_start_index = res
        else:
            bev_embed = res
        bs = mlvl_feats[0].size(0)
        query_pos, query = torch.split(object_query_embed, self.embed_dims, dim=1)
        query_pos = query_pos.unsqueeze(0).expand(bs, -1, -1)
        query = query.unsqueeze(0).expand(bs, -1, -1)
        reference_points = self.reference_points(query_pos).sigmoid()
        init_reference_out = reference_points
        query = query.permute(1, 0, 2)
        query_pos = query_pos.permute(1, 0, 2)
        bev_embed = bev_embed.permute(1, 0, 2)
        inter_states, inter_references = self.decoder(
            query=query, key=None, value=bev_embed, query_pos=query_pos,
            reference_points=reference_points, reg_branches=reg_branches, cls_branches=cls_branches,
            spatial_shapes=torch.tensor([[bev_h, bev_w]], device=query.device),
            level_start_index=torch.tensor([0], device=query.device), **kwargs)
        if ret_inter:
            return (bev_embed, inter_states, init_reference_out, inter_references, feat_flatten,
                    spatial_shapes, level_start_index)
        return bev_embed, inter_states, init_reference_out, inter_references
