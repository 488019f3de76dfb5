"""Oracle for the attention modules around the operator (test infrastructure).

mmcv-free CPU restatements of the reference's attention modules, with the same
parameter names (so one ``state_dict`` loads into the reference, the oracle and
the CUDA modules alike) and every observable quirk kept:

* ``OracleMSDeformableAttention3D``  <- spatial_cross_attention.py:175-403
* ``OracleSpatialCrossAttention``    <- spatial_cross_attention.py:28-173
* ``OracleTemporalSelfAttention``    <- temporal_self_attention.py:24-289
* ``OracleCustomMSDeformableAttention`` <- decoder.py:129-358
* ``OracleBEVFormerLayer`` / ``OracleBEVFormerEncoder`` <- encoder.py:243-352, 400-519
  with mmcv's FFN (Linear-ReLU-Dropout-Linear-Dropout + identity) and LayerNorm.

They always take the reference's CPU branch, i.e. the grid_sample operator
(:func:`oracle.msda_oracle.msda_torch`).
"""
import math
import torch
import torch.nn as nn

from .msda_oracle import msda_torch
from .geometry_oracle import reference_points_2d, reference_points_3d, point_sampling


def _xavier_uniform_(linear):
    nn.init.xavier_uniform_(linear.weight, gain=1.0)
    if linear.bias is not None:
        nn.init.constant_(linear.bias, 0.0)


def _ring_bias(num_heads, reps, num_points):
    """sampling_offsets.bias initialisation (spatial_cross_attention.py:259-271)."""
    ang = torch.arange(num_heads, dtype=torch.float32) * (2.0 * math.pi / num_heads)
    g = torch.stack([ang.cos(), ang.sin()], -1)
    g = (g / g.abs().max(-1, keepdim=True)[0]).view(num_heads, 1, 1, 2).repeat(1, reps, num_points, 1)
    for i in range(num_points):
        g[:, :, i, :] *= i + 1
    return g.reshape(-1)


class _DeformBase(nn.Module):
    def _init_common(self, reps):
        nn.init.constant_(self.sampling_offsets.weight, 0.0)
        self.sampling_offsets.bias.data = _ring_bias(self.num_heads, reps, self.num_points)
        nn.init.constant_(self.attention_weights.weight, 0.0)
        nn.init.constant_(self.attention_weights.bias, 0.0)
        _xavier_uniform_(self.value_proj)
        if getattr(self, 'output_proj', None) is not None:
            _xavier_uniform_(self.output_proj)


class OracleMSDeformableAttention3D(_DeformBase):
    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=8,
                 im2col_step=64, dropout=0.1, batch_first=True, norm_cfg=None,
                 init_cfg=None, attn_logits_clamp=None, debug_attn_nan=False):
        super().__init__()
        assert embed_dims % num_heads == 0
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.num_levels, self.num_points = num_levels, num_points
        self.im2col_step, self.batch_first = im2col_step, batch_first
        self.attn_logits_clamp = attn_logits_clamp      # stored, never applied (:252 vs :347)
        self.output_proj = None
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        self._init_common(self.num_levels)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, **kwargs):
        if value is None:
            value = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        B, Nq, _ = query.shape
        _, Nk, _ = value.shape
        M, L, P = self.num_heads, self.num_levels, self.num_points
        value = self.value_proj(value)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(B, Nk, M, -1)
        off = self.sampling_offsets(query).view(B, Nq, M, L, P, 2)
        att = self.attention_weights(query).view(B, Nq, M, L * P).softmax(-1).view(B, Nq, M, L, P)
        assert reference_points.shape[-1] == 2
        # Z-anchor expansion, spatial_cross_attention.py:361-376: point index p = k*D + z
        norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        D = reference_points.shape[2]
        ref = reference_points[:, :, None, None, None, :, :]
        off = off / norm[None, None, None, :, None, :]
        off = off.view(B, Nq, M, L, P // D, D, 2)
        loc = (ref + off).view(B, Nq, M, L, P, 2)
        out = msda_torch(value, spatial_shapes, loc, att)
        if not self.batch_first:
            out = out.permute(1, 0, 2)
        return out


class OracleSpatialCrossAttention(nn.Module):
    def __init__(self, embed_dims=256, num_cams=6, pc_range=None, dropout=0.1,
                 init_cfg=None, batch_first=False, deformable_attention=None, **kwargs):
        super().__init__()
        cfg = dict(deformable_attention or dict(embed_dims=256, num_levels=4))
        cfg.pop('type', None)
        self.deformable_attention = OracleMSDeformableAttention3D(**cfg)
        self.embed_dims, self.num_cams = embed_dims, num_cams
        self.pc_range, self.batch_first = pc_range, batch_first
        self.dropout = nn.Dropout(dropout)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        _xavier_uniform_(self.output_proj)

    def forward(self, query, key, value, residual=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                reference_points_cam=None, bev_mask=None, level_start_index=None,
                flag='encoder', **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        inp_residual = query if residual is None else residual
        slots = torch.zeros_like(query)
        if query_pos is not None:
            query = query + query_pos
        bs, num_query, _ = query.size()
        D = reference_points_cam.size(3)
        # index lists from batch element 0 only (:135-139)
        indexes = [m[0].sum(-1).nonzero().squeeze(-1) for m in bev_mask]
        max_len = max(len(e) for e in indexes)
        q_rb = query.new_zeros([bs, self.num_cams, max_len, self.embed_dims])
        r_rb = reference_points_cam.new_zeros([bs, self.num_cams, max_len, D, 2])
        for j in range(bs):
            for i in range(self.num_cams):
                idx = indexes[i]
                q_rb[j, i, :len(idx)] = query[j, idx]
                r_rb[j, i, :len(idx)] = reference_points_cam[i][j, idx]
        num_cams, l, bs, _ = key.shape
        key = key.permute(2, 0, 1, 3).reshape(bs * self.num_cams, l, self.embed_dims)
        value = value.permute(2, 0, 1, 3).reshape(bs * self.num_cams, l, self.embed_dims)
        q = self.deformable_attention(
            query=q_rb.view(bs * self.num_cams, max_len, self.embed_dims), key=key, value=value,
            reference_points=r_rb.view(bs * self.num_cams, max_len, D, 2),
            spatial_shapes=spatial_shapes, level_start_index=level_start_index
        ).view(bs, self.num_cams, max_len, self.embed_dims)
        for j in range(bs):
            for i, idx in enumerate(indexes):
                slots[j, idx] += q[j, i, :len(idx)]
        count = (bev_mask.sum(-1) > 0).permute(1, 2, 0).sum(-1)
        count = torch.clamp(count, min=1.0)
        slots = slots / count[..., None]
        slots = self.output_proj(slots)
        return self.dropout(slots) + inp_residual


class OracleTemporalSelfAttention(_DeformBase):
    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4,
                 num_bev_queue=2, im2col_step=64, dropout=0.1, batch_first=True,
                 norm_cfg=None, init_cfg=None, attn_logits_clamp=None, debug_attn_nan=False):
        super().__init__()
        assert embed_dims % num_heads == 0
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.num_levels, self.num_points = num_levels, num_points
        self.num_bev_queue, self.batch_first = num_bev_queue, batch_first
        self.im2col_step, self.attn_logits_clamp = im2col_step, attn_logits_clamp
        self.dropout = nn.Dropout(dropout)
        Q = num_bev_queue
        self.sampling_offsets = nn.Linear(embed_dims * Q, Q * num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims * Q, Q * num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        self._init_common(self.num_levels * self.num_bev_queue)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, flag='decoder', **kwargs):
        if value is None:
            assert self.batch_first
            bs, len_bev, c = query.shape
            value = torch.stack([query, query], 1).reshape(bs * 2, len_bev, c)   # :183-186
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, Nq, C = query.shape
        _, Nk, _ = value.shape
        M, L, P, Q = self.num_heads, self.num_levels, self.num_points, self.num_bev_queue
        assert Q == 2
        query = torch.cat([value[:bs], query], -1)                               # :203
        value = self.value_proj(value)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.reshape(bs * Q, Nk, M, -1)
        off = self.sampling_offsets(query).view(bs, Nq, M, Q, L, P, 2)
        att = self.attention_weights(query).view(bs, Nq, M, Q, L * P)
        if self.attn_logits_clamp is not None:
            c = float(self.attn_logits_clamp)
            att = att.clamp(min=-c, max=c)
        att = att.softmax(-1).view(bs, Nq, M, Q, L, P)
        att = att.permute(0, 3, 1, 2, 4, 5).reshape(bs * Q, Nq, M, L, P).contiguous()
        off = off.permute(0, 3, 1, 2, 4, 5, 6).reshape(bs * Q, Nq, M, L, P, 2)
        if reference_points.shape[-1] == 2:
            norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
            loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        else:
            loc = (reference_points[:, :, None, :, None, :2]
                   + off / P * reference_points[:, :, None, :, None, 2:] * 0.5)
        out = msda_torch(value, spatial_shapes, loc, att)                        # (bs*Q, Nq, C)
        out = out.permute(1, 2, 0).view(Nq, C, bs, Q).mean(-1).permute(2, 0, 1)  # :270-282
        out = self.output_proj(out)
        if not self.batch_first:
            out = out.permute(1, 0, 2)
        return self.dropout(out) + identity


class OracleCustomMSDeformableAttention(_DeformBase):
    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4,
                 im2col_step=64, dropout=0.1, batch_first=False, norm_cfg=None,
                 init_cfg=None, attn_logits_clamp=None, debug_attn_nan=False):
        super().__init__()
        assert embed_dims % num_heads == 0
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.num_levels, self.num_points = num_levels, num_points
        self.im2col_step, self.batch_first = im2col_step, batch_first
        self.attn_logits_clamp = attn_logits_clamp
        self.dropout = nn.Dropout(dropout)
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        self._init_common(self.num_levels)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, flag='decoder', **kwargs):
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, Nq, _ = query.shape
        _, Nk, _ = value.shape
        M, L, P = self.num_heads, self.num_levels, self.num_points
        value = self.value_proj(value)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(bs, Nk, M, -1)
        off = self.sampling_offsets(query).view(bs, Nq, M, L, P, 2)
        att = self.attention_weights(query).view(bs, Nq, M, L * P)
        if self.attn_logits_clamp is not None:
            c = float(self.attn_logits_clamp)
            att = att.clamp(min=-c, max=c)
        att = att.softmax(-1).view(bs, Nq, M, L, P)
        if reference_points.shape[-1] == 2:
            norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
            loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            loc = (reference_points[:, :, None, :, None, :2]
                   + off / P * reference_points[:, :, None, :, None, 2:] * 0.5)
        else:
            raise ValueError('Last dim of reference_points must be 2 or 4, '
                             f'but get {reference_points.shape[-1]} instead.')
        out = msda_torch(value, spatial_shapes, loc, att)
        out = self.output_proj(out)
        if not self.batch_first:
            out = out.permute(1, 0, 2)
        return self.dropout(out) + identity


class OracleFFN(nn.Module):
    """mmcv FFN(embed_dims, feedforward_channels, num_fcs=2, ReLU, ffn_drop, add_identity)."""

    def __init__(self, embed_dims=256, feedforward_channels=512, ffn_drop=0.1):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True),
                          nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))

    def forward(self, x, identity=None):
        out = self.layers(x)
        return (x if identity is None else identity) + out


class OracleBEVFormerLayer(nn.Module):
    """('self_attn','norm','cross_attn','norm','ffn','norm') layer, encoder.py:400-519."""

    def __init__(self, embed_dims=256, feedforward_channels=512, num_levels=4,
                 sca_points=8, tsa_points=4, num_cams=6, pc_range=None, dropout=0.1,
                 operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')):
        super().__init__()
        self.operation_order = tuple(operation_order)
        self.pre_norm = self.operation_order[0] == 'norm'
        atts = []
        for op in self.operation_order:
            if op == 'self_attn':
                atts.append(OracleTemporalSelfAttention(embed_dims=embed_dims, num_levels=1,
                                                        num_points=tsa_points, dropout=dropout))
            elif op == 'cross_attn':
                atts.append(OracleSpatialCrossAttention(
                    embed_dims=embed_dims, num_cams=num_cams, pc_range=pc_range, dropout=dropout,
                    batch_first=True,
                    deformable_attention=dict(embed_dims=embed_dims, num_points=sca_points,
                                              num_levels=num_levels)))
        self.attentions = nn.ModuleList(atts)
        self.ffns = nn.ModuleList([OracleFFN(embed_dims, feedforward_channels, dropout)
                                   for op in self.operation_order if op == 'ffn'])
        self.norms = nn.ModuleList([nn.LayerNorm(embed_dims)
                                    for op in self.operation_order if op == 'norm'])

    def forward(self, query, key=None, value=None, bev_pos=None, query_pos=None, ref_2d=None,
                ref_3d=None, bev_h=None, bev_w=None, reference_points_cam=None,
                spatial_shapes=None, level_start_index=None, prev_bev=None, bev_mask=None,
                **kwargs):
        ni = ai = fi = 0
        identity = query
        for op in self.operation_order:
            if op == 'self_attn':
                query = self.attentions[ai](
                    query, prev_bev, prev_bev, identity if self.pre_norm else None,
                    query_pos=bev_pos, reference_points=ref_2d,
                    spatial_shapes=torch.tensor([[bev_h, bev_w]], device=query.device),
                    level_start_index=torch.tensor([0], device=query.device))
                ai += 1
                identity = query
            elif op == 'norm':
                query = self.norms[ni](query)
                ni += 1
            elif op == 'cross_attn':
                query = self.attentions[ai](
                    query, key, value, identity if self.pre_norm else None,
                    query_pos=query_pos, reference_points=ref_3d,
                    reference_points_cam=reference_points_cam, bev_mask=bev_mask,
                    spatial_shapes=spatial_shapes, level_start_index=level_start_index)
                ai += 1
                identity = query
            elif op == 'ffn':
                query = self.ffns[fi](query, identity if self.pre_norm else None)
                fi += 1
        return query


class OracleBEVFormerEncoder(nn.Module):
    """Layer loop + geometry, encoder.py:243-352 (img_metas reduced to lidar2img + img h/w)."""

    def __init__(self, num_layers=6, pc_range=None, num_points_in_pillar=4, **layer_kwargs):
        super().__init__()
        self.pc_range = pc_range
        self.num_points_in_pillar = num_points_in_pillar
        self.layers = nn.ModuleList([OracleBEVFormerLayer(pc_range=pc_range, **layer_kwargs)
                                     for _ in range(num_layers)])

    def forward(self, bev_query, key, value, bev_h=None, bev_w=None, bev_pos=None,
                spatial_shapes=None, level_start_index=None, prev_bev=None, shift=None,
                lidar2img=None, img_h=None, img_w=None, layer_inputs=None, return_all=False):
        """bev_query, bev_pos (HW, bs, C); key=value (num_cam, Nk, bs, C); prev_bev (HW, bs, C) | None.

        ``layer_inputs`` (test aid): a list with one (bs, HW, C) tensor or None per layer; where
        given, that layer runs on it instead of the previous layer's output ("teacher forcing": a
        deep low-precision model is compared one layer at a time).  ``return_all`` returns the list
        of every layer's output instead of the last one."""
        bs = bev_query.size(1)
        ref_3d = reference_points_3d(bev_h, bev_w, self.pc_range[5] - self.pc_range[2],
                                     self.num_points_in_pillar, bs=bs,
                                     device=bev_query.device, dtype=bev_query.dtype)
        ref_2d = reference_points_2d(bev_h, bev_w, bs=bs, device=bev_query.device,
                                     dtype=bev_query.dtype)
        ref_cam, bev_mask = point_sampling(ref_3d, self.pc_range, lidar2img, img_h, img_w)
        shift_ref_2d = ref_2d            # aliasing kept on purpose, encoder.py:309-311
        if shift is not None:
            shift_ref_2d += shift[:, None, None, :]
        bev_query = bev_query.permute(1, 0, 2)
        bev_pos = bev_pos.permute(1, 0, 2)
        _, len_bev, nlvl, _ = ref_2d.shape
        if prev_bev is not None:
            prev_bev = prev_bev.permute(1, 0, 2)
            prev_bev = torch.stack([prev_bev, bev_query], 1).reshape(bs * 2, len_bev, -1)
            hybrid = torch.stack([shift_ref_2d, ref_2d], 1).reshape(bs * 2, len_bev, nlvl, 2)
        else:
            hybrid = torch.stack([ref_2d, ref_2d], 1).reshape(bs * 2, len_bev, nlvl, 2)
        out = bev_query
        outs = []
        for i, layer in enumerate(self.layers):
            if layer_inputs is not None and layer_inputs[i] is not None:
                bev_query = layer_inputs[i]
            out = layer(bev_query, key, value, bev_pos=bev_pos, ref_2d=hybrid, ref_3d=ref_3d,
                        bev_h=bev_h, bev_w=bev_w, spatial_shapes=spatial_shapes,
                        level_start_index=level_start_index, reference_points_cam=ref_cam,
                        bev_mask=bev_mask, prev_bev=prev_bev)
            bev_query = out
            outs.append(out)
        return outs if return_all else out
