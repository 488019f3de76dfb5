"""B200-native hosts of the MapTRv2 decoder's deformable cross-attention.

Drop-in for ``projects/mmdet3d_plugin/maptrv2/modules/decoder.py``: ``MapTRv2Decoder`` (:10-58,
iterative reference refinement) and ``MapTRv2DecoupledDetrTransformerDecoderLayer`` (:61-213).
Only the ``cross_attn`` operation (``CustomMSDeformableAttention``, configured at
``configs/bevformer/bev_tiny_det_mapv2.py:53-56``) is on the hot path and runs on the fused
sm_100a kernel; the two dense self-attentions are out of scope and use
``torch.nn.MultiheadAttention`` behind mmcv's ``MultiheadAttention`` call convention.
"""
import copy

import torch
import torch.nn as nn

from ..registry import (ATTENTION, HAVE_MMCV, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE,
                        BaseModule, build_attention, build_transformer_layer)
from .decoder import hoisted_projections, inverse_sigmoid
from ..rowops import LayerNorm
from .encoder import FFN

if not HAVE_MMCV:
    @ATTENTION.register_module()
    class MultiheadAttention(BaseModule):
        """mmcv.cnn.bricks.transformer.MultiheadAttention call convention over nn.MultiheadAttention."""

        def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout=None,
                     dropout_layer=None, init_cfg=None, batch_first=False, **kwargs):
            super().__init__(init_cfg)
            if dropout is not None:
                attn_drop = dropout
                dropout_layer = dict(type='Dropout', drop_prob=dropout)
            self.embed_dims = embed_dims
            self.num_heads = num_heads
            self.batch_first = batch_first
            self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
            self.proj_drop = nn.Dropout(proj_drop)
            p = (dropout_layer or {}).get('drop_prob', 0.)
            self.dropout_layer = nn.Dropout(p) if p > 0 else nn.Identity()

        def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None,
                    attn_mask=None, key_padding_mask=None, **kwargs):
            if key is None:
                key = query
            if value is None:
                value = key
            if identity is None:
                identity = query
            if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
                key_pos = query_pos
            if query_pos is not None:
                query = query + query_pos
            if key_pos is not None:
                key = key + key_pos
            if self.batch_first:
                query, key, value = (t.transpose(0, 1) for t in (query, key, value))
            # need_weights=False: the averaged attention map mmcv discards (`[0]`) is never built,
            # and torch runs the fused scaled_dot_product_attention kernels
            out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                            key_padding_mask=key_padding_mask, need_weights=False)[0]
            if self.batch_first:
                out = out.transpose(0, 1)
            return identity + self.dropout_layer(self.proj_drop(out))


@TRANSFORMER_LAYER.register_module()
class MapTRv2DecoupledDetrTransformerDecoderLayer(BaseModule):
    """self_attn, norm, self_attn, norm, cross_attn, norm, ffn, norm (reference :61-213)."""

    def __init__(self, attn_cfgs, feedforward_channels, num_vec=50, num_pts_per_vec=20,
                 ffn_dropout=0.0, operation_order=None, act_cfg=dict(type='ReLU', inplace=True),
                 norm_cfg=dict(type='LN'), ffn_num_fcs=2, init_cfg=None, batch_first=False,
                 **kwargs):
        super().__init__(init_cfg)
        assert len(operation_order) == 8
        assert set(operation_order) == {'self_attn', 'norm', 'cross_attn', 'ffn'}
        self.operation_order = tuple(operation_order)
        self.pre_norm = self.operation_order[0] == 'norm'
        self.batch_first = batch_first
        self.num_attn = self.operation_order.count('self_attn') + self.operation_order.count('cross_attn')
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(self.num_attn)]
        self.attentions = nn.ModuleList()
        for cfg in attn_cfgs:
            cfg = copy.deepcopy(cfg)
            cfg['batch_first'] = batch_first
            self.attentions.append(build_attention(cfg))
        self.embed_dims = self.attentions[0].embed_dims
        self.ffns = nn.ModuleList([FFN(embed_dims=self.embed_dims,
                                       feedforward_channels=feedforward_channels,
                                       num_fcs=ffn_num_fcs, ffn_drop=ffn_dropout)
                                   for _ in range(self.operation_order.count('ffn'))])
        self.norms = nn.ModuleList([LayerNorm(self.embed_dims)
                                    for _ in range(self.operation_order.count('norm'))])
        self.num_vec = num_vec
        self.num_pts_per_vec = num_pts_per_vec

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        norm_index = attn_index = ffn_index = 0
        identity = query
        if attn_masks is None:
            attn_masks = [None for _ in range(self.num_attn)]
        elif isinstance(attn_masks, torch.Tensor):
            attn_masks = [copy.deepcopy(attn_masks) for _ in range(self.num_attn)]
        else:
            assert len(attn_masks) == self.num_attn
        V = int(kwargs.get('num_vec', self.num_vec))
        Pn = int(kwargs.get('num_pts_per_vec', self.num_pts_per_vec))
        self_attn_mask = kwargs.get('self_attn_mask', None)

        for op in self.operation_order:
            if op == 'self_attn':
                _, nb, nd = query.shape
                if attn_index == 0:
                    # sequence axis = vectors, batch axis = (point, sample)      (:131-148)
                    q = query.view(V, Pn, nb, nd).flatten(1, 2)
                    qp = query_pos.view(V, Pn, nb, nd).flatten(1, 2)
                    q = self.attentions[attn_index](q, q, q, identity if self.pre_norm else None,
                                                    query_pos=qp, key_pos=qp,
                                                    attn_mask=self_attn_mask,
                                                    key_padding_mask=query_key_padding_mask)
                    query = q.view(V, Pn, nb, nd).flatten(0, 1)
                else:
                    # sequence axis = points, batch axis = (vector, sample)      (:149-185)
                    q = query.view(V, Pn, nb, nd).permute(1, 0, 2, 3).contiguous().flatten(1, 2)
                    qp = query_pos.view(V, Pn, nb, nd).permute(1, 0, 2, 3).contiguous().flatten(1, 2)
                    q = self.attentions[attn_index](q, q, q, identity if self.pre_norm else None,
                                                    query_pos=qp, key_pos=qp,
                                                    attn_mask=attn_masks[attn_index],
                                                    key_padding_mask=query_key_padding_mask)
                    query = q.view(Pn, V, nb, nd).permute(1, 0, 2, 3).contiguous().flatten(0, 1)
                attn_index += 1
                identity = query
            elif op == 'norm':
                query = self.norms[norm_index](query)
                norm_index += 1
            elif op == 'cross_attn':
                query = self.attentions[attn_index](
                    query, key, value, identity if self.pre_norm else None, query_pos=query_pos,
                    key_pos=key_pos, attn_mask=attn_masks[attn_index],
                    key_padding_mask=key_padding_mask, **kwargs)
                attn_index += 1
                identity = query
            elif op == 'ffn':
                query = self.ffns[ffn_index](query, identity if self.pre_norm else None)
                ffn_index += 1
        return query


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class MapTRv2Decoder(BaseModule):
    """MapTRv2 decoder with iterative reference refinement (reference :10-58)."""

    def __init__(self, transformerlayers=None, num_layers=None, return_intermediate=False,
                 init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        self.num_layers = num_layers
        self.layers = nn.ModuleList([build_transformer_layer(c) for c in transformerlayers])
        self.embed_dims = self.layers[0].embed_dims
        self.return_intermediate = return_intermediate
        self.fp16_enabled = False
        # project the BEV for the cross-attentions of ALL layers in one batched GEMM before the loop
        self.hoist_value_proj = True

    def forward(self, query, *args, reference_points=None, reg_branches=None,
                key_padding_mask=None, **kwargs):
        """query (Nq, bs, C); reference_points (bs, Nq, 2) in [0, 1]; value=(HW, bs, C) in kwargs."""
        output = query
        if self.training and torch.is_grad_enabled() and query.is_cuda:
            from ..rowops import advance_dropout_step
            advance_dropout_step(query.device)     # fresh masks for the fused dropouts of this step
        intermediate, intermediate_refs = [], []
        projected = hoisted_projections(self.layers, args, kwargs) if self.hoist_value_proj else None
        for lid, layer in enumerate(self.layers):
            ref_in = reference_points[..., :2].unsqueeze(2)
            extra = {} if projected is None else {'projected_value': projected[lid]}
            output = layer(output, *args, reference_points=ref_in,
                           key_padding_mask=key_padding_mask, **kwargs, **extra)
            output = output.permute(1, 0, 2)
            if reg_branches is not None:
                tmp = reg_branches[lid](output)
                new_ref = (tmp + inverse_sigmoid(reference_points)).sigmoid()
                reference_points = new_ref.detach()
            output = output.permute(1, 0, 2)
            if self.return_intermediate:
                intermediate.append(output)
                intermediate_refs.append(reference_points)
        if self.return_intermediate:
            return torch.stack(intermediate), torch.stack(intermediate_refs)
        return output, reference_points
