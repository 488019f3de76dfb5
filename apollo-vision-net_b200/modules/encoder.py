"""B200-native ``BEVFormerEncoder`` / ``BEVFormerLayer``.

Drop-in for ``projects/mmdet3d_plugin/bevformer/modules/encoder.py``: same registry names,
constructor arguments and forward keywords.  The geometry (``point_sampling``, reference
:89-241) runs in one device kernel through the C ABI ``bev_point_sampling``; the per-camera
hit lists the reference extracts with ``nonzero()`` and a host synchronisation
(spatial_cross_attention.py:135-139) stay on the device as a bit field, so a whole encoder
pass enqueues without ever waiting for the GPU.

mmcv is absent from this image, so the layer (attentions + LayerNorm + FFN in
``operation_order``) is written out here with mmcv's parameter names
(``attentions.{i}``, ``norms.{i}``, ``ffns.{i}.layers.0.0 / .1``), cf.
custom_base_transformer_layer.py:35-258.
"""
import copy

import numpy as np
import torch
import torch.nn as nn

from .._lib import nvtx_range
from ..fused_ops import BevGeometry, bev_point_sampling
from ..rowops import Junction, LayerNorm, Linear, ReLU, advance_dropout_step, linear, linear_add_layernorm
from ..registry import (TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE, BaseModule, build_attention,
                        build_transformer_layer)


def _activation(act_cfg):
    """mmcv ``build_activation_layer`` for the activations the reference's configs use here."""
    cfg = dict(act_cfg or dict(type='ReLU', inplace=True))
    typ = cfg.pop('type', 'ReLU')
    if typ == 'ReLU':
        return ReLU(inplace=cfg.get('inplace', True))       # fused backward (rowops.ReLU)
    if typ == 'GELU':
        return nn.GELU()
    raise KeyError(f'unsupported activation {typ}')


class FFN(BaseModule):
    """mmcv FFN: Linear-ReLU-Dropout-Linear-Dropout plus identity."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        assert num_fcs >= 2
        self.embed_dims = embed_dims
        self.feedforward_channels = feedforward_channels
        layers = []
        in_channels = embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(Linear(in_channels, feedforward_channels),
                                        _activation(act_cfg), nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        self.add_identity = add_identity

    def forward(self, x, identity=None, post_norm=None):
        """``post_norm``: the layer's next LayerNorm; with it and two fcs the last Linear, its dropout,
        the identity add and the norm run as one fused node (result already normalised), and the first
        activation's dropout is fused into the ReLU pass."""
        res = x if identity is None else identity
        last_drop = self.layers[-1]
        if post_norm is not None and self.add_identity and len(self.layers) == 3:
            fc1, act, drop1 = self.layers[0]
            p1 = drop1.p if self.training else 0.0
            # x is also the residual: its two gradients meet in fc1's dX GEMM (rowops.Junction)
            tok = Junction() if (res is x and isinstance(fc1, Linear) and torch.is_grad_enabled()) else None
            h = linear(x, fc1.weight, fc1.bias, tok) if tok is not None else fc1(x)
            h = act.forward_dropout(h, p1) if hasattr(act, 'forward_dropout') else drop1(act(h))
            return linear_add_layernorm(h, self.layers[1], res, post_norm,
                                        p=last_drop.p if self.training else 0.0, junction=tok)
        out = self.layers(x)
        if self.add_identity:
            out = res + out
        return out if post_norm is None else post_norm(out)


@TRANSFORMER_LAYER.register_module()
class BEVFormerLayer(BaseModule):
    """self_attn (TSA) -> norm -> cross_attn (SCA) -> norm -> ffn -> norm (reference :354-519)."""

    def __init__(self, attn_cfgs, feedforward_channels=None, ffn_dropout=0.0, operation_order=None,
                 act_cfg=dict(type='ReLU', inplace=True), norm_cfg=dict(type='LN'), ffn_num_fcs=2,
                 ffn_cfgs=None, init_cfg=None, batch_first=True, **kwargs):
        super().__init__(init_cfg)
        assert operation_order is not None
        assert set(operation_order) <= {'self_attn', 'norm', 'ffn', 'cross_attn'}
        self.batch_first = batch_first
        self.operation_order = tuple(operation_order)
        self.pre_norm = self.operation_order[0] == 'norm'
        self.num_attn = self.operation_order.count('self_attn') + self.operation_order.count('cross_attn')
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(self.num_attn)]
        assert len(attn_cfgs) == self.num_attn
        self.attentions = nn.ModuleList()
        idx = 0
        for op in self.operation_order:
            if op in ('self_attn', 'cross_attn'):
                cfg = copy.deepcopy(attn_cfgs[idx])
                cfg['batch_first'] = self.batch_first      # custom_base_transformer_layer.py:129-132
                att = build_attention(cfg)
                att.operation_name = op
                self.attentions.append(att)
                idx += 1
        self.embed_dims = self.attentions[0].embed_dims
        ffn_cfgs = dict(ffn_cfgs or {})
        ffn_cfgs.pop('type', None)
        ffn_cfgs.setdefault('embed_dims', self.embed_dims)
        ffn_cfgs.setdefault('feedforward_channels', feedforward_channels or 2 * self.embed_dims)
        ffn_cfgs.setdefault('num_fcs', ffn_num_fcs)
        ffn_cfgs.setdefault('ffn_drop', ffn_dropout)
        self.ffns = nn.ModuleList([FFN(**copy.deepcopy(ffn_cfgs))
                                   for _ in range(self.operation_order.count('ffn'))])
        self.norms = nn.ModuleList([LayerNorm(self.embed_dims)
                                    for _ in range(self.operation_order.count('norm'))])
        self.fp16_enabled = False

    def forward(self, query, key=None, value=None, bev_pos=None, query_pos=None, key_pos=None,
                attn_masks=None, query_key_padding_mask=None, key_padding_mask=None, ref_2d=None,
                ref_3d=None, bev_h=None, bev_w=None, reference_points_cam=None, mask=None,
                spatial_shapes=None, level_start_index=None, prev_bev=None, **kwargs):
        norm_index = attn_index = ffn_index = 0
        identity = query
        tsa_shapes = kwargs.pop('_tsa_shapes', None)
        if tsa_shapes is None:
            tsa_shapes = (torch.tensor([[bev_h, bev_w]], device=query.device),
                          torch.tensor([0], device=query.device))
        # post-norm layers hand the LayerNorm that follows a block to the block, which fuses it
        # with its output projection and residual add; the 'norm' step is then skipped
        order = self.operation_order
        fused_norm = False
        for i, op in enumerate(order):
            post = None
            if op != 'norm' and not self.pre_norm and i + 1 < len(order) and order[i + 1] == 'norm':
                post = self.norms[norm_index]
            if op == 'norm' and fused_norm:
                fused_norm = False
                norm_index += 1
                continue
            fused_norm = post is not None
            if op == 'self_attn':
                with nvtx_range('BEVFormerLayer.self_attn'):
                    query = self.attentions[attn_index](
                        query, prev_bev, prev_bev, identity if self.pre_norm else None,
                        query_pos=bev_pos, key_pos=bev_pos, key_padding_mask=query_key_padding_mask,
                        reference_points=ref_2d, spatial_shapes=tsa_shapes[0],
                        level_start_index=tsa_shapes[1], bev_h=bev_h, bev_w=bev_w, post_norm=post,
                        **kwargs)
                    attn_index += 1
                    identity = query
            elif op == 'norm':
                query = self.norms[norm_index](query)
                norm_index += 1
            elif op == 'cross_attn':
                with nvtx_range('BEVFormerLayer.cross_attn'):
                    query = self.attentions[attn_index](
                        query, key, value, identity if self.pre_norm else None, query_pos=query_pos,
                        key_pos=key_pos, reference_points=ref_3d,
                        reference_points_cam=reference_points_cam, mask=mask,
                        key_padding_mask=key_padding_mask, spatial_shapes=spatial_shapes,
                        level_start_index=level_start_index, bev_h=bev_h, bev_w=bev_w, post_norm=post,
                        **kwargs)
                    attn_index += 1
                    identity = query
            elif op == 'ffn':
                with nvtx_range('BEVFormerLayer.ffn'):
                    query = self.ffns[ffn_index](query, identity if self.pre_norm else None,
                                                 post_norm=post)
                    ffn_index += 1
        return query


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class BEVFormerEncoder(BaseModule):
    """BEV encoder (reference :25-352)."""

    def __init__(self, transformerlayers=None, num_layers=None, pc_range=None,
                 num_points_in_pillar=4, return_intermediate=False, dataset_type='nuscenes',
                 init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        assert isinstance(transformerlayers, list) and len(transformerlayers) == num_layers
        self.num_layers = num_layers
        self.layers = nn.ModuleList([build_transformer_layer(c) for c in transformerlayers])
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm
        self.return_intermediate = return_intermediate
        self.num_points_in_pillar = num_points_in_pillar
        self.pc_range = pc_range
        self.fp16_enabled = False
        self._ref_cache = {}

    @staticmethod
    def get_reference_points(H, W, Z=8, num_points_in_pillar=4, dim='3d', bs=1, device='cuda',
                             dtype=torch.float):
        """Reference points for SCA (3d, (bs, D, HW, 3)) and TSA (2d, (bs, HW, 1, 2)),
        reference :47-86.  Generated on the host so the values are those of the reference's CPU
        path bit for bit, then moved to ``device``."""
        if dim == '3d':
            D = num_points_in_pillar
            zs = torch.linspace(0.5, Z - 0.5, D, dtype=dtype) / Z
            xs = torch.linspace(0.5, W - 0.5, W, dtype=dtype) / W
            ys = torch.linspace(0.5, H - 0.5, H, dtype=dtype) / H
            ref = torch.empty(D, H, W, 3, dtype=dtype)
            ref[..., 0] = xs.view(1, 1, W)
            ref[..., 1] = ys.view(1, H, 1)
            ref[..., 2] = zs.view(D, 1, 1)
            return ref.view(1, D, H * W, 3).repeat(bs, 1, 1, 1).to(device)
        if dim == '2d':
            ys = torch.linspace(0.5, H - 0.5, H, dtype=dtype) / H
            xs = torch.linspace(0.5, W - 0.5, W, dtype=dtype) / W
            ref = torch.empty(H, W, 2, dtype=dtype)
            ref[..., 0] = xs.view(1, W)
            ref[..., 1] = ys.view(H, 1)
            return ref.view(1, H * W, 1, 2).repeat(bs, 1, 1, 1).to(device)
        raise ValueError(dim)

    @staticmethod
    def _meta_geometry(img_metas):
        """lidar2img (B, num_cam, 4, 4) float32 and (H_img, W_img) of camera 0 / sample 0
        from the reference's ``img_metas`` (reference :123-141, :193-223)."""
        if isinstance(img_metas, list) and img_metas and isinstance(img_metas[0], list):
            img_metas = [m[-1] for m in img_metas]                 # (bs, len_queue) -> current frame
        l2i = np.asarray([np.asarray(m['lidar2img']) for m in img_metas], dtype=np.float32)
        shape0 = img_metas[0].get('img_shape', None)
        h = w = None
        if isinstance(shape0, (list, tuple)) and len(shape0) > 0:
            first = shape0[0]
            if isinstance(first, (list, tuple, np.ndarray)):
                h, w = int(first[0]), int(first[1])
            elif len(shape0) >= 2:
                h, w = int(shape0[0]), int(shape0[1])
        if not h or not w or h <= 0 or w <= 0:
            h, w = 1, 1                                            # the reference's last resort
        return l2i, h, w

    def point_sampling(self, reference_points, pc_range, img_metas=None, lidar2img=None,
                       img_shape=None, return_geometry=False, with_lists=True):
        """reference_points (bs, D, HW, 3) -> (reference_points_cam (num_cam, bs, HW, D, 2),
        bev_mask (num_cam, bs, HW, D) bool), reference :89-241; ``return_geometry`` also hands
        back the device-side hit bit field and compacted hit lists."""
        if lidar2img is None:
            lidar2img, h, w = self._meta_geometry(img_metas)
        else:
            h, w = int(img_shape[0]), int(img_shape[1])
        geo = bev_point_sampling(reference_points, pc_range, lidar2img, h, w, with_lists=with_lists)
        if return_geometry:
            return geo
        return geo.reference_points_cam, geo.bev_mask

    def _hoisted_sca_values(self, key, value, bs):
        """Every layer's SpatialCrossAttention projects the SAME camera features with its own ``value_proj``
        (spatial_cross_attention.py:338-343, once per layer).  When the features want a gradient the
        projections of all layers are taken in front of the layer loop as ONE autograd node
        (:class:`decoder.HoistedValueProjFunction`), whose backward accumulates the feature gradient with
        beta = 1 GEMMs -- otherwise autograd adds six 95 MB gradients with five separate kernels.  Returns a
        list of (bs * num_cams * Nk, C) tensors, or None (every layer then projects for itself)."""
        from .decoder import HoistedValueProjFunction
        from .spatial_cross_attention import MSDeformableAttention3D, SpatialCrossAttention
        if (len(self.layers) < 2 or not torch.is_tensor(value) or value.dim() != 4 or not value.is_cuda or
                not (torch.is_grad_enabled() and value.requires_grad) or (key is not None and key is not value)):
            return None
        found = []
        for layer in self.layers:
            mods = [a for a in getattr(layer, 'attentions', []) if isinstance(a, SpatialCrossAttention)]
            if len(mods) != 1 or not isinstance(mods[0].deformable_attention, MSDeformableAttention3D):
                return None
            vp = mods[0].deformable_attention.value_proj
            if vp.bias is None or vp.weight.dtype != value.dtype or vp.weight.shape != (value.shape[-1],) * 2:
                return None
            found.append(vp)
        num_cams, l, bs_v, C = value.shape
        flat = value.permute(2, 0, 1, 3).reshape(bs_v * num_cams * l, C)
        # (side by side: the fused SCA backward writes every layer's value gradient into its column block of one
        # matrix, so the feature gradient and the six weight gradients are one GEMM each)
        return list(HoistedValueProjFunction.apply(flat, True, *[vp.weight for vp in found],
                                                   *[vp.bias for vp in found]))

    def _hoisted_tsa_values(self, pair):
        """With history the TSA value -- the (bs * 2, HW, C) pair [prev_bev, bev_query] -- is the same tensor in every
        layer (reference :317-320 builds it once in front of the loop), each layer projecting it with its own
        ``value_proj``.  In training the projections are hoisted like the SCA ones: the six weight gradients
        (and the pair's gradient, when it wants one) become one GEMM each."""
        from .decoder import HoistedValueProjFunction
        from .temporal_self_attention import TemporalSelfAttention
        if (len(self.layers) < 2 or pair is None or not pair.is_cuda or not torch.is_grad_enabled() or
                not self.training):
            return None
        found = []
        for layer in self.layers:
            mods = [a for a in getattr(layer, 'attentions', []) if isinstance(a, TemporalSelfAttention)]
            if len(mods) != 1 or not mods[0].batch_first:
                return None
            vp = mods[0].value_proj
            if (vp.bias is None or vp.weight.dtype != pair.dtype or vp.weight.shape != (pair.shape[-1],) * 2 or
                    not vp.weight.requires_grad):
                return None
            found.append(vp)
        flat = pair.reshape(-1, pair.shape[-1])
        outs = HoistedValueProjFunction.apply(flat, True, *[vp.weight for vp in found], *[vp.bias for vp in found])
        return [o.view(pair.shape) for o in outs]

    def forward(self, bev_query, key, value, *args, bev_z=None, bev_h=None, bev_w=None,
                bev_pos=None, spatial_shapes=None, level_start_index=None, valid_ratios=None,
                prev_bev=None, shift=0., img_metas=None, lidar2img=None, img_shape=None,
                row_shard=None, **kwargs):
        """bev_query, bev_pos (HW, bs, C); key = value (num_cam, Nk, bs, C); prev_bev
        (HW, bs, C) or None; shift (bs, 2).  Returns (bs, HW, C) (or the stacked intermediates).

        ``row_shard=(rank, world)`` runs only this rank's contiguous block of BEV rows
        (``parallel.bev_row_range``) and returns (bs, rows*bev_w, C): BEV queries are independent
        units of SCA / TSA / FFN / LayerNorm, the value tensors (image features, the
        [prev_bev, bev] pair) stay replicated, and with history (``prev_bev`` given) no layer
        needs an exchange -- the caller all-gathers the rows once at encoder exit
        (``parallel.all_gather_bev_rows``).  Without history the TSA value is the current layer
        input, which would need a per-layer all-gather: that case is refused here."""
        output = bev_query
        intermediate = []
        bs = bev_query.size(1)
        dev, dt = bev_query.device, bev_query.dtype
        if self.training and torch.is_grad_enabled() and bev_query.is_cuda:
            advance_dropout_step(dev)          # new dropout masks for this step (device-side counter)
        ck = (bev_h, bev_w, bs, str(dev), self.num_points_in_pillar)
        if ck not in self._ref_cache:
            self._ref_cache = {ck: (
                self.get_reference_points(bev_h, bev_w, self.pc_range[5] - self.pc_range[2],
                                          self.num_points_in_pillar, dim='3d', bs=bs, device=dev,
                                          dtype=torch.float32),
                self.get_reference_points(bev_h, bev_w, dim='2d', bs=bs, device=dev,
                                          dtype=torch.float32),
                (torch.tensor([[bev_h, bev_w]], device=dev), torch.tensor([0], device=dev)))}
        ref_3d, ref_2d_base, tsa_shapes = self._ref_cache[ck]
        ref_2d = ref_2d_base.clone()

        geo = self.point_sampling(ref_3d, self.pc_range, img_metas, lidar2img, img_shape,
                                  return_geometry=True, with_lists=False)

        # "bug kept for reproducing the paper": shift_ref_2d aliases ref_2d (reference :309-311)
        shift_ref_2d = ref_2d
        if torch.is_tensor(shift):
            shift_ref_2d += shift[:, None, None, :].to(ref_2d.dtype)
        else:
            shift_ref_2d += shift

        bev_query = bev_query.permute(1, 0, 2)
        bev_pos = bev_pos.permute(1, 0, 2)
        _, len_bev, num_bev_level, _ = ref_2d.shape
        if prev_bev is not None:
            prev_bev = prev_bev.permute(1, 0, 2)
            prev_bev = torch.stack([prev_bev, bev_query], 1).reshape(bs * 2, len_bev, -1)
            hybird_ref_2d = torch.stack([shift_ref_2d, ref_2d], 1).reshape(
                bs * 2, len_bev, num_bev_level, 2)
        else:
            hybird_ref_2d = torch.stack([ref_2d, ref_2d], 1).reshape(
                bs * 2, len_bev, num_bev_level, 2)

        shard_h = bev_h
        if row_shard is not None:
            from ..parallel import bev_row_range
            if prev_bev is None:
                raise RuntimeError('row sharding needs prev_bev: without history the TSA value is '
                                   'the current layer input and every layer would need an all-gather')
            y0, y1 = bev_row_range(bev_h, int(row_shard[0]), int(row_shard[1]))
            q0, q1 = y0 * bev_w, y1 * bev_w
            shard_h = y1 - y0
            bev_query = bev_query[:, q0:q1].contiguous()
            bev_pos = bev_pos[:, q0:q1].contiguous()
            hybird_ref_2d = hybird_ref_2d[:, q0:q1].contiguous()
            geo = geo.rows(q0, q1)
            kwargs = dict(kwargs, row_slice=(q0, q1))

        projected = self._hoisted_sca_values(key, value, bs)
        tsa_projected = self._hoisted_tsa_values(prev_bev)
        for li, layer in enumerate(self.layers):
            if projected is not None:
                kwargs['_sca_projected_value'] = projected[li]
            if tsa_projected is not None:
                kwargs['_tsa_projected_value'] = tsa_projected[li]
            output = layer(bev_query, key, value, *args, bev_pos=bev_pos, ref_2d=hybird_ref_2d,
                           ref_3d=ref_3d, bev_h=shard_h, bev_w=bev_w, spatial_shapes=spatial_shapes,
                           level_start_index=level_start_index,
                           reference_points_cam=geo.reference_points_cam, bev_mask=geo.bev_mask,
                           bev_geometry=geo, prev_bev=prev_bev, _tsa_shapes=tsa_shapes, **kwargs)
            bev_query = output
            if self.return_intermediate:
                intermediate.append(output)
        if self.return_intermediate:
            return torch.stack(intermediate)
        return output
