"""Oracle for the decoders' dense self-attentions (test infrastructure, never imported by the product).

CPU restatement of what the reference computes at
``projects/mmdet3d_plugin/maptrv2/modules/decoder.py:129-188`` (two ``self_attn`` operations of a MapTRv2
decoder layer) and in mmdet's ``DetrTransformerDecoderLayer`` self-attention of the detection decoder
(``projects/mmdet3d_plugin/bevformer/modules/decoder.py:50-126`` builds the layers): mmcv's
``MultiheadAttention`` (mmcv-full==1.4.0, ``mmcv/cnn/bricks/transformer.py``; absent from /root/reference,
pinned at ``requirements.apollo_vnet.pip.txt:51``) adds the positional encodings to query and key and calls
``torch.nn.MultiheadAttention``, whose published algorithm is restated here explicitly:

    q, k = (x + pos) Wq^T + bq, (x + pos) Wk^T + bk;   v = x Wv^T + bv
    per head:  P = softmax(q k^T / sqrt(Dh) + mask);  P <- dropout(P);  o = P v
    out = identity + dropout(concat(o) Wo^T + bo)

Pinning: ``tests/test_mha_cpu.py`` checks :func:`mha_module_oracle` against ``torch.nn.MultiheadAttention``
itself (the third-party implementation the reference executes) on seeded inputs, with and without the
boolean mask, and the two token groupings against the reference's own view / permute / flatten sequence; the
golden fixture ``tests/golden/maptrv2_decoder_small.npz`` (unmodified reference decoder classes through
``oracle/refshim``) pins the whole layer.
"""
import torch


def token_rows(layout):
    """(G, S) int64: the activation row of token s of group g (layout = G, S, seq_stride, hi_stride,
    lo_stride, n_lo -- see apollo-vision-net_b200/mha.py; restated here so that the oracle shares no code
    with the product)."""
    G, S, seq_stride, hi_stride, lo_stride, n_lo = (int(v) for v in layout)
    g = torch.arange(G, dtype=torch.int64)
    s = torch.arange(S, dtype=torch.int64)
    return ((g // n_lo) * hi_stride + (g % n_lo) * lo_stride)[:, None] + s[None, :] * seq_stride


def attention_core_oracle(qk, v, layout, heads, attn_mask=None, keep=None, p=0.0):
    """qk (rows, 2C) = projected queries | keys, v (rows, C) -> o (rows, C), differentiable.
    ``attn_mask``: bool (S, S), True = may not attend.  ``keep``: bool (G * heads, S, S) dropout keep mask
    on the attention weights (None = no dropout), kept weights scaled by 1 / (1 - p)."""
    rows, C = v.shape
    G, S = int(layout[0]), int(layout[1])
    Dh = C // heads
    idx = token_rows(layout).reshape(-1)                                  # (G * S,)
    q = qk[idx, :C].view(G, S, heads, Dh).permute(0, 2, 1, 3)
    k = qk[idx, C:].view(G, S, heads, Dh).permute(0, 2, 1, 3)
    vv = v[idx].view(G, S, heads, Dh).permute(0, 2, 1, 3)
    scores = torch.matmul(q, k.transpose(-1, -2)) / float(Dh) ** 0.5       # (G, heads, S, S)
    if attn_mask is not None:
        scores = scores.masked_fill(attn_mask.view(1, 1, S, S), float('-inf'))
    prob = torch.softmax(scores, dim=-1)
    if keep is not None:
        prob = prob * keep.view(G, heads, S, S).to(prob.dtype) / (1.0 - p)
    o = torch.matmul(prob, vv).permute(0, 2, 1, 3).reshape(G * S, C)
    out = torch.zeros_like(v)
    return out.index_copy(0, idx, o)


def mha_module_oracle(x, pos, identity, in_proj_weight, in_proj_bias, out_weight, out_bias, layout, heads,
                      attn_mask=None):
    """Evaluation-mode forward of mmcv's MultiheadAttention used as self-attention over the token layout:
    x, pos, identity (rows, C)."""
    C = x.shape[-1]
    xp = x if pos is None else x + pos
    qk = torch.nn.functional.linear(xp, in_proj_weight[:2 * C], in_proj_bias[:2 * C])
    v = torch.nn.functional.linear(x, in_proj_weight[2 * C:], in_proj_bias[2 * C:])
    o = attention_core_oracle(qk, v, layout, heads, attn_mask)
    return identity + torch.nn.functional.linear(o, out_weight, out_bias)
