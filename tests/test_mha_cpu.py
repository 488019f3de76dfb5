"""Pins the self-attention oracle (oracle/mha_oracle.py) on the implementation the reference executes:
``torch.nn.MultiheadAttention`` behind mmcv's ``MultiheadAttention`` (positional encodings added to query
and key, identity added to the output), called at ``projects/mmdet3d_plugin/maptrv2/modules/decoder.py:129-188``
under two token groupings.  CPU only; the CUDA kernels are compared with the same oracle in
``tests/test_mha_gpu.py``."""
import pytest
import torch

from oracle.mha_oracle import attention_core_oracle, mha_module_oracle, token_rows


def _layouts(V, Pn, nb):
    import apollo_vision_net_b200.mha as m
    return m.inter_vector_layout(V, Pn, nb), m.intra_vector_layout(V, Pn, nb)


def _one2many_mask(V, one2one):
    mask = torch.zeros(V, V, dtype=torch.bool)
    mask[one2one:, :one2one] = True
    mask[:one2one, one2one:] = True
    return mask


def test_token_layouts_are_the_references_view_permute_flatten():
    """Row of token s of group g == where the reference's reshapes put that token (decoder.py:131-185)."""
    V, Pn, nb = 7, 5, 3
    inter, intra = _layouts(V, Pn, nb)
    rows = torch.arange(V * Pn * nb).view(V * Pn, nb)                      # row index of query[n, b]
    # first self-attention: query.view(V, Pn, nb).flatten(1, 2) -> (sequence V, batch Pn * nb)
    ref = rows.view(V, Pn, nb).flatten(1, 2)
    assert torch.equal(token_rows(inter), ref.t())
    # second: .view(V, Pn, nb).permute(1, 0, 2).contiguous().flatten(1, 2) -> (sequence Pn, batch V * nb)
    ref = rows.view(V, Pn, nb).permute(1, 0, 2).contiguous().flatten(1, 2)
    assert torch.equal(token_rows(intra), ref.t())
    for lay in (inter, intra):
        assert sorted(token_rows(lay).reshape(-1).tolist()) == list(range(V * Pn * nb))


@pytest.mark.parametrize('masked', [False, True])
@pytest.mark.parametrize('grouping', ['inter', 'intra'])
def test_module_oracle_equals_torch_multihead_attention(masked, grouping):
    torch.manual_seed(3)
    V, Pn, nb, C, H = 12, 5, 2, 64, 4
    mha = torch.nn.MultiheadAttention(C, H, dropout=0.1).double().eval()
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.1)
        mha.out_proj.bias.normal_(0, 0.1)
    query = torch.randn(V * Pn, nb, C, dtype=torch.float64)
    pos = torch.randn(V * Pn, nb, C, dtype=torch.float64)
    inter, intra = _layouts(V, Pn, nb)
    if grouping == 'inter':
        lay, S = inter, V
        q = query.view(V, Pn, nb, C).flatten(1, 2)
        qp = pos.view(V, Pn, nb, C).flatten(1, 2)
    else:
        lay, S = intra, Pn
        q = query.view(V, Pn, nb, C).permute(1, 0, 2, 3).contiguous().flatten(1, 2)
        qp = pos.view(V, Pn, nb, C).permute(1, 0, 2, 3).contiguous().flatten(1, 2)
    mask = _one2many_mask(S, S // 3) if masked else None
    # mmcv: query + query_pos, key + key_pos, value = the tokens, identity = the tokens
    ref = q + mha(q + qp, q + qp, q, attn_mask=mask, need_weights=False)[0]
    if grouping == 'inter':
        ref = ref.view(V, Pn, nb, C).flatten(0, 1)
    else:
        ref = ref.view(Pn, V, nb, C).permute(1, 0, 2, 3).contiguous().flatten(0, 1)
    got = mha_module_oracle(query.view(-1, C), pos.view(-1, C), query.view(-1, C), mha.in_proj_weight,
                            mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, lay, H, mask)
    assert torch.allclose(got.view_as(ref), ref, rtol=0, atol=1e-12)


def test_core_oracle_dropout_and_gradients_against_autograd_of_the_plain_formula():
    torch.manual_seed(5)
    G, S, H, Dh, p = 3, 9, 2, 8, 0.25
    C = H * Dh
    qk = torch.randn(G * S, 2 * C, dtype=torch.float64, requires_grad=True)
    v = torch.randn(G * S, C, dtype=torch.float64, requires_grad=True)
    keep = torch.rand(G * H, S, S) > p
    lay = (G, S, 1, 0, S, G)                                               # batch-first rows
    out = attention_core_oracle(qk, v, lay, H, None, keep, p)
    q = qk[:, :C].view(G, S, H, Dh).transpose(1, 2)
    k = qk[:, C:].view(G, S, H, Dh).transpose(1, 2)
    vv = v.view(G, S, H, Dh).transpose(1, 2)
    prob = torch.softmax(q @ k.transpose(-1, -2) / Dh ** 0.5, -1) * keep.view(G, H, S, S) / (1 - p)
    ref = (prob @ vv).transpose(1, 2).reshape(G * S, C)
    assert torch.allclose(out, ref, atol=1e-13)
