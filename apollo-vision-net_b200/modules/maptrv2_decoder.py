"""B200-native hosts of the MapTRv2 decoder's deformable cross-attention.

Drop-in for ``projects/mmdet3d_plugin/maptrv2/modules/decoder.py``: ``MapTRv2Decoder`` (:10-58,
iterative reference refinement) and ``MapTRv2DecoupledDetrTransformerDecoderLayer`` (:61-213).
The ``cross_attn`` operation (``CustomMSDeformableAttention``, configured at
``configs/bevformer/bev_tiny_det_mapv2.py:53-56``) runs on the fused sm_100a deformable-attention kernel;
the two dense self-attentions run on the small-sequence attention core of
``csrc/mha.cu`` behind mmcv's ``MultiheadAttention`` call convention, attended in place in the layer's
activations, and every ``norm`` is folded into the block in front of it.
"""
import copy

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..mha import (batch_first_layout, inter_vector_layout, intra_vector_layout, self_attention,
                   sequence_first_layout, supported_impl)
from ..registry import (ATTENTION, HAVE_MMCV, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE,
                        BaseModule, build_attention, build_transformer_layer)
from .decoder import hoisted_projections, inverse_sigmoid
from ..rowops import LayerNorm, linear, linear_add_layernorm
from .encoder import FFN

def _mask_ok(attn_mask, S):
    return attn_mask is None or (attn_mask.dtype == torch.bool and attn_mask.dim() == 2 and
                                 tuple(attn_mask.shape) == (S, S) and attn_mask.is_cuda)


class FusedSelfAttentionMixin:
    """The self-attention of an mmcv-convention ``MultiheadAttention`` (``self.attn`` = a
    ``torch.nn.MultiheadAttention``, ``self.dropout_layer``, ``self.proj_drop``) on the sm_100a kernels of
    ``csrc/mha.cu``: in-projection GEMMs (q | k from x + pos, v from x) and the fused attention core over the
    token layout, then ``identity + dropout(out_proj(.))`` -- with the layer's next LayerNorm folded into
    the same node when the caller hands it over (``post_norm``).  Parameters and their names are those of
    ``torch.nn.MultiheadAttention`` (checkpoints load unchanged)."""

    # True: the sm_100a core wherever it is the faster engine -- every 16-bit case (tensor-core kernels) and
    # fp32 sequences of up to 64 tokens; LONG fp32 sequences go to torch's attention, whose fp32 kernels are
    # 3-4x faster than the one-warp-per-row FMA path (profiles/r02_mha.md).  'always': the core for every
    # supported case (what the fp32 parity tests run).  False: torch.nn.MultiheadAttention throughout.
    use_fused_core = True
    fp32_fused_max_tokens = 64

    def fused_self_attention_ok(self, x, layout, attn_mask, key_padding_mask=None, pos=None):
        a = self.attn
        if not (self.use_fused_core and x.is_cuda and key_padding_mask is None and _mask_ok(attn_mask, layout.S)):
            return False
        if pos is not None and (pos.shape != x.shape or pos.dtype != x.dtype):      # broadcast encodings: torch path
            return False
        if x.dtype == torch.float32 and layout.S > self.fp32_fused_max_tokens and self.use_fused_core != 'always':
            return False
        if not getattr(a, '_qkv_same_embed_dim', True) or a.in_proj_bias is None or a.bias_k is not None \
                or a.add_zero_attn or x.dtype != a.in_proj_weight.dtype:
            return False
        if self.training and self.proj_drop.p > 0:
            return False
        return supported_impl(layout, a.num_heads, a.embed_dim // a.num_heads, x.dtype) != 0

    def fused_self_attention(self, x, pos, layout, attn_mask=None, identity=None, post_norm=None):
        """x, pos, identity: (..., C) activations whose flattened rows the layout describes."""
        a = self.attn
        if identity is None:
            identity = x
        p_attn = float(a.dropout) if self.training else 0.0
        # in-projection + attention core as one autograd node (mha.SelfAttentionFunction)
        o = self_attention(x, pos, a.in_proj_weight, a.in_proj_bias, layout, a.num_heads, attn_mask, p_attn)
        p_out = float(getattr(self.dropout_layer, 'p', 0.0)) if self.training else 0.0
        if post_norm is not None:
            return linear_add_layernorm(o.view(identity.shape), a.out_proj, identity, post_norm, p_out)
        out = linear(o, a.out_proj.weight, a.out_proj.bias).view(identity.shape)
        return identity + (F.dropout(out, p_out, True) if p_out > 0 else out)


if not HAVE_MMCV:
    @ATTENTION.register_module()
    class MultiheadAttention(FusedSelfAttentionMixin, BaseModule):
        """mmcv.cnn.bricks.transformer.MultiheadAttention call convention.  Self-attention (key and value
        are the query tensor, boolean 2-d ``attn_mask`` or none, no padding mask) runs on the fused sm_100a
        attention core; everything else goes through ``torch.nn.MultiheadAttention`` as in mmcv."""

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
                    attn_mask=None, key_padding_mask=None, post_norm=None, **kwargs):
            if key is None:
                key = query
            if value is None:
                value = key
            if identity is None:
                identity = query
            if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
                key_pos = query_pos
            if key is query and value is query and key_pos is query_pos and query.dim() == 3:
                layout = (batch_first_layout(query.shape[0], query.shape[1]) if self.batch_first
                          else sequence_first_layout(query.shape[0], query.shape[1]))
                if self.fused_self_attention_ok(query, layout, attn_mask, key_padding_mask, query_pos):
                    return self.fused_self_attention(query, query_pos, layout, attn_mask, identity, post_norm)
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
            out = identity + self.dropout_layer(self.proj_drop(out))
            return out if post_norm is None else post_norm(out)


def run_layer_ops(layer, query, key, value, query_pos, key_pos, attn_masks, query_key_padding_mask,
                  key_padding_mask, self_attention, kwargs):
    """The operation loop of an mmcv ``BaseTransformerLayer``-style decoder layer, with every ``norm`` that
    directly follows an attention / FFN block handed to that block (``post_norm``): the block's last
    Linear, its dropout, the residual add and the LayerNorm then run as one fused node.  A block that
    does not take the norm (foreign attention modules) is followed by the plain norm call.
    ``self_attention(attn, attn_index, query, identity, post_norm)`` runs one ``self_attn`` operation."""
    norm_index = attn_index = ffn_index = 0
    identity = query
    ops = layer.operation_order
    i = 0
    while i < len(ops):
        op = ops[i]
        post = None
        if op != 'norm' and i + 1 < len(ops) and ops[i + 1] == 'norm' and not layer.pre_norm:
            post = layer.norms[norm_index]
        took_norm = False
        if op == 'self_attn':
            attn = layer.attentions[attn_index]
            took_norm = post is not None and isinstance(attn, FusedSelfAttentionMixin)
            query = self_attention(attn, attn_index, query, identity if layer.pre_norm else None,
                                   post if took_norm else None)
            attn_index += 1
            identity = query
        elif op == 'norm':
            query = layer.norms[norm_index](query)
            norm_index += 1
        elif op == 'cross_attn':
            attn = layer.attentions[attn_index]
            took_norm = post is not None and getattr(attn, 'accepts_post_norm', False)
            extra = {'post_norm': post} if took_norm else {}
            query = attn(query, key, value, identity if layer.pre_norm else None, query_pos=query_pos,
                         key_pos=key_pos, attn_mask=attn_masks[attn_index],
                         key_padding_mask=key_padding_mask, **kwargs, **extra)
            attn_index += 1
            identity = query
        elif op == 'ffn':
            ffn = layer.ffns[ffn_index]
            took_norm = post is not None and isinstance(ffn, FFN)
            query = ffn(query, identity if layer.pre_norm else None, **({'post_norm': post} if took_norm else {}))
            ffn_index += 1
        if took_norm:
            norm_index += 1
            i += 1                      # the norm ran inside the block
        i += 1
    return query


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
        if attn_masks is None:
            attn_masks = [None for _ in range(self.num_attn)]
        elif isinstance(attn_masks, torch.Tensor):
            attn_masks = [copy.deepcopy(attn_masks) for _ in range(self.num_attn)]
        else:
            assert len(attn_masks) == self.num_attn
        V = int(kwargs.get('num_vec', self.num_vec))
        Pn = int(kwargs.get('num_pts_per_vec', self.num_pts_per_vec))
        self_attn_mask = kwargs.get('self_attn_mask', None)

        def self_attention(attn, attn_index, query, identity, post_norm):
            _, nb, nd = query.shape
            first = attn_index == 0
            mask = self_attn_mask if first else attn_masks[attn_index]
            # inter-vector: the sequence axis is the vectors, one group per (point, sample)   (:131-148)
            # intra-vector: the sequence axis is the points, one group per (vector, sample)   (:149-185)
            layout = inter_vector_layout(V, Pn, nb) if first else intra_vector_layout(V, Pn, nb)
            if (isinstance(attn, FusedSelfAttentionMixin) and not attn.batch_first and
                    attn.fused_self_attention_ok(query, layout, mask, query_key_padding_mask, query_pos)):
                # attended in place in the (V * Pn, nb, C) activations: no permute / contiguous copies
                return attn.fused_self_attention(query, query_pos, layout, mask, identity, post_norm)
            extra = {} if post_norm is None else {'post_norm': post_norm}
            if first:
                q = query.view(V, Pn, nb, nd).flatten(1, 2)
                qp = query_pos.view(V, Pn, nb, nd).flatten(1, 2)
                q = attn(q, q, q, identity, query_pos=qp, key_pos=qp, attn_mask=mask,
                         key_padding_mask=query_key_padding_mask, **extra)
                return q.view(V, Pn, nb, nd).flatten(0, 1)
            q = query.view(V, Pn, nb, nd).permute(1, 0, 2, 3).contiguous().flatten(1, 2)
            qp = query_pos.view(V, Pn, nb, nd).permute(1, 0, 2, 3).contiguous().flatten(1, 2)
            q = attn(q, q, q, identity, query_pos=qp, key_pos=qp, attn_mask=mask,
                     key_padding_mask=query_key_padding_mask, **extra)
            return q.view(Pn, V, nb, nd).permute(1, 0, 2, 3).contiguous().flatten(0, 1)

        return run_layer_ops(self, query, key, value, query_pos, key_pos, attn_masks, query_key_padding_mask,
                             key_padding_mask, self_attention, kwargs)


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
