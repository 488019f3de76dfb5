"""B200-native host of the detection decoder around the deformable cross-attention.

Drop-in for ``DetectionTransformerDecoder`` (``projects/mmdet3d_plugin/bevformer/modules/
decoder.py:50-126``): the layer sequence that calls ``CustomMSDeformableAttention`` (900 object
queries sampling the BEV) and refines 3-d reference points between the layers.  Its layers are
mmdet's ``DetrTransformerDecoderLayer`` (self_attn, norm, cross_attn, norm, ffn, norm --
``configs/bevformer/bev_tiny_det_map_apollo.py`` decoder section); without mmcv / mmdet installed
the same layer is provided here.  The cross-attention runs on the fused sm_100a
deformable-attention kernel, the dense self-attention over the object queries on the attention core of
``csrc/mha.cu``, and every ``norm`` is folded into the block in front of it.
"""
import copy

import torch
import torch.nn as nn

from ..registry import (HAVE_MMCV, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE, BaseModule,
                        build_attention, build_transformer_layer)
from ..rowops import LayerNorm
from .decoder import hoisted_projections, inverse_sigmoid
from .encoder import FFN
from .maptrv2_decoder import run_layer_ops      # also registers the MultiheadAttention shim without mmcv


if not HAVE_MMCV:
    @TRANSFORMER_LAYER.register_module()
    class DetrTransformerDecoderLayer(BaseModule):
        """mmdet's decoder layer on mmcv's ``BaseTransformerLayer`` call convention: the query is
        (Nq, bs, C); ``self_attn`` attends the queries to themselves, ``cross_attn`` receives
        key / value and every extra keyword (reference points, level tables)."""

        def __init__(self, attn_cfgs, feedforward_channels=None, ffn_dropout=0.0,
                     operation_order=None, act_cfg=dict(type='ReLU', inplace=True),
                     norm_cfg=dict(type='LN'), ffn_num_fcs=2, ffn_cfgs=None, init_cfg=None,
                     batch_first=False, **kwargs):
            super().__init__(init_cfg)
            assert operation_order is not None
            assert set(operation_order) <= {'self_attn', 'norm', 'ffn', 'cross_attn'}
            self.operation_order = tuple(operation_order)
            self.pre_norm = self.operation_order[0] == 'norm'
            self.batch_first = batch_first
            self.num_attn = (self.operation_order.count('self_attn') +
                             self.operation_order.count('cross_attn'))
            if isinstance(attn_cfgs, dict):
                attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(self.num_attn)]
            assert len(attn_cfgs) == self.num_attn
            self.attentions = nn.ModuleList()
            for cfg in attn_cfgs:
                cfg = copy.deepcopy(cfg)
                cfg['batch_first'] = batch_first
                self.attentions.append(build_attention(cfg))
            self.embed_dims = self.attentions[0].embed_dims
            ffn_cfgs = dict(ffn_cfgs or {})
            ffn_cfgs.pop('type', None)
            ffn_cfgs.setdefault('embed_dims', self.embed_dims)
            ffn_cfgs.setdefault('feedforward_channels', feedforward_channels or 4 * self.embed_dims)
            ffn_cfgs.setdefault('num_fcs', ffn_num_fcs)
            ffn_cfgs.setdefault('ffn_drop', ffn_dropout)
            self.ffns = nn.ModuleList([FFN(**copy.deepcopy(ffn_cfgs))
                                       for _ in range(self.operation_order.count('ffn'))])
            self.norms = nn.ModuleList([LayerNorm(self.embed_dims)
                                        for _ in range(self.operation_order.count('norm'))])

        def forward(self, query, key=None, value=None, query_pos=None, key_pos=None,
                    attn_masks=None, query_key_padding_mask=None, key_padding_mask=None, **kwargs):
            if attn_masks is None:
                attn_masks = [None for _ in range(self.num_attn)]
            elif isinstance(attn_masks, torch.Tensor):
                attn_masks = [copy.deepcopy(attn_masks) for _ in range(self.num_attn)]
            else:
                assert len(attn_masks) == self.num_attn

            def self_attention(attn, attn_index, query, identity, post_norm):
                extra = {} if post_norm is None else {'post_norm': post_norm}
                return attn(query, query, query, identity, query_pos=query_pos, key_pos=query_pos,
                            attn_mask=attn_masks[attn_index], key_padding_mask=query_key_padding_mask, **extra)

            return run_layer_ops(self, query, key, value, query_pos, key_pos, attn_masks,
                                 query_key_padding_mask, key_padding_mask, self_attention, kwargs)


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class DetectionTransformerDecoder(BaseModule):
    """Detection decoder with iterative refinement of (x, y, z) reference points
    (reference decoder.py:50-126): every layer's cross-attention samples the BEV around
    ``reference_points[..., :2]``; ``reg_branches[lid]`` moves x, y (outputs 0:2) and z
    (output 4) in logit space; the refined points are detached."""

    def __init__(self, transformerlayers=None, num_layers=None, return_intermediate=False,
                 init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        assert isinstance(transformerlayers, (list, tuple)) and len(transformerlayers) == num_layers
        self.num_layers = num_layers
        self.layers = nn.ModuleList([build_transformer_layer(c) for c in transformerlayers])
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = getattr(self.layers[0], 'pre_norm', False)
        self.return_intermediate = return_intermediate
        self.fp16_enabled = False
        # project the BEV for the cross-attentions of ALL layers in one batched GEMM before the loop
        self.hoist_value_proj = True

    def forward(self, query, *args, reference_points=None, reg_branches=None,
                key_padding_mask=None, **kwargs):
        """query (Nq, bs, C); reference_points (bs, Nq, 3) in [0, 1]; value (HW, bs, C)."""
        output = query
        if self.training and torch.is_grad_enabled() and query.is_cuda:
            from ..rowops import advance_dropout_step
            advance_dropout_step(query.device)     # fresh masks for the fused dropouts of this step
        intermediate, intermediate_refs = [], []
        projected = hoisted_projections(self.layers, args, kwargs) if self.hoist_value_proj else None
        for lid, layer in enumerate(self.layers):
            ref_in = reference_points[..., :2].unsqueeze(2)                  # (bs, Nq, 1, 2)
            extra = {} if projected is None else {'projected_value': projected[lid]}
            output = layer(output, *args, reference_points=ref_in,
                           key_padding_mask=key_padding_mask, **kwargs, **extra)
            output = output.permute(1, 0, 2)
            if reg_branches is not None:
                tmp = reg_branches[lid](output)
                assert reference_points.shape[-1] == 3
                xy = tmp[..., :2] + inverse_sigmoid(reference_points[..., :2])
                z = tmp[..., 4:5] + inverse_sigmoid(reference_points[..., 2:3])
                reference_points = torch.cat([xy, z], -1).sigmoid().detach()
            output = output.permute(1, 0, 2)
            if self.return_intermediate:
                intermediate.append(output)
                intermediate_refs.append(reference_points)
        if self.return_intermediate:
            return torch.stack(intermediate), torch.stack(intermediate_refs)
        return output, reference_points
