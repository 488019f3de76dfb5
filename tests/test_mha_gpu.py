"""CUDA parity of the decoders' self-attention core (csrc/mha.cu) against the CPU oracle
(oracle/mha_oracle.py, pinned on torch.nn.MultiheadAttention in tests/test_mha_cpu.py): the inter- and
intra-vector self-attentions of the MapTRv2 decoder layer (projects/mmdet3d_plugin/maptrv2/modules/
decoder.py:129-188) and the detection decoder's query self-attention, forward and backward, both kernel
families (tensor cores for 16-bit models, the fp32 FMA path), boolean mask, ragged sizes, dropout.

Tolerances: fp32 FMA path 1e-5 forward / 1e-4 gradients (north_star's bar for fp32).  16-bit paths are compared
with the oracle evaluated in fp64 on the SAME 16-bit-rounded inputs; what remains is the rounding of the
probabilities / dS to the model dtype before the second product and of the result: 2e-2 for bf16 (8 mantissa
bits), 3e-3 for fp16, relative to the tensor's max."""
import pytest
import torch

from oracle.mha_oracle import attention_core_oracle, mha_module_oracle
from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _mask(S, one2one):
    m = torch.zeros(S, S, dtype=torch.bool)
    m[one2one:, :one2one] = True
    m[:one2one, one2one:] = True
    return m


def _tols(dtype):
    return {torch.float32: (1e-5, 1e-4), torch.bfloat16: (2e-2, 2e-2), torch.float16: (3e-3, 3e-3)}[dtype]


def _run_core(layout, H, Dh, dtype, impl, mask=None, p=0.0, seed=0):
    import apollo_vision_net_b200.mha as m
    import apollo_vision_net_b200.rowops as ro
    g = torch.Generator().manual_seed(seed)
    C = H * Dh
    rows = layout.G * layout.S
    qk = (torch.randn(rows, 2 * C, generator=g) * 1.5).to(dtype)
    v = torch.randn(rows, C, generator=g).to(dtype)
    go = torch.randn(rows, C, generator=g).to(dtype)
    qk_d = qk.to(DEV).requires_grad_(True)
    v_d = v.to(DEV).requires_grad_(True)
    mask_d = None if mask is None else mask.to(DEV)
    keep = None
    if p > 0:
        ro.advance_dropout_step(DEV)
        key = ro.dropout_state(DEV).clone()
    out = m.self_attention_core(qk_d, v_d, layout, H, mask_d, p, impl)
    if p > 0:
        keep = m.attention_keep_mask(layout, H, key, ro._drop_site[0], p, DEV).cpu()
    out.backward(go.to(DEV))
    torch.cuda.synchronize()
    qk_o = qk.double().requires_grad_(True)
    v_o = v.double().requires_grad_(True)
    ref = attention_core_oracle(qk_o, v_o, tuple(layout), H, mask, keep, p)
    ref.backward(go.double())
    return (out, qk_d.grad, v_d.grad), (ref, qk_o.grad, v_o.grad), keep


CASES = [
    # name, layout factory args (kind, a, b, c), heads, head_dim, masked
    ('inter_vector_50', ('inter', 50, 20, 1), 8, 32, True),
    ('inter_vector_350', ('inter', 350, 20, 1), 8, 32, True),
    ('intra_vector_20', ('intra', 70, 20, 2), 8, 32, False),
    ('ragged_37', ('seq', 37, 3, 0), 4, 32, True),
    ('single_token', ('seq', 1, 5, 0), 8, 32, False),
    ('head_dim_64', ('batch', 2, 45, 0), 4, 64, True),
    ('det_queries_900', ('seq', 900, 1, 0), 8, 32, False),
]


def _layout(spec):
    import apollo_vision_net_b200.mha as m
    kind, a, b, c = spec
    if kind == 'inter':
        return m.inter_vector_layout(a, b, c), a
    if kind == 'intra':
        return m.intra_vector_layout(a, b, c), b
    if kind == 'seq':
        return m.sequence_first_layout(a, b), a
    return m.batch_first_layout(a, b), b


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize('dtype,impl', [(torch.float32, 1), (torch.bfloat16, 2), (torch.float16, 2),
                                        (torch.bfloat16, 1)],
                         ids=['fp32_fma', 'bf16_tensor_core', 'fp16_tensor_core', 'bf16_fma'])
def test_core_matches_oracle(case, dtype, impl):
    _, spec, H, Dh, masked = case
    layout, S = _layout(spec)
    if impl == 1 and dtype != torch.float32 and S > 400:
        pytest.skip('the 16-bit FMA path is covered at the smaller sizes')
    mask = _mask(S, max(1, S // 7)) if masked and S > 1 else None
    got, ref, _ = _run_core(layout, H, Dh, dtype, impl, mask)
    tf, tg = _tols(dtype)
    assert rel_err(got[0], ref[0]) <= tf
    if S == 1:          # softmax over one key: the exact q / k gradient is zero, what is left is rounding noise
        assert float(got[1].abs().max()) <= 1e-3 * tg * float(ref[2].abs().max())
    else:
        assert rel_err(got[1], ref[1]) <= tg
    assert rel_err(got[2], ref[2]) <= tg


def test_auto_dispatch_picks_tensor_cores_for_16_bit_and_fma_for_fp32():
    import apollo_vision_net_b200.mha as m
    lay = m.inter_vector_layout(350, 20, 1)
    assert m.supported_impl(lay, 8, 32, torch.bfloat16) == m.IMPL_TENSOR_CORE
    assert m.supported_impl(lay, 8, 32, torch.float16) == m.IMPL_TENSOR_CORE
    assert m.supported_impl(lay, 8, 32, torch.float32) == m.IMPL_FMA
    assert m.supported_impl(lay, 8, 16, torch.bfloat16) == m.IMPL_FMA
    assert m.supported_impl(lay, 8, 24, torch.float32) == 0
    assert m.supported_impl(m.sequence_first_layout(5000, 1), 8, 32, torch.bfloat16) == 0
    with pytest.raises(RuntimeError, match='tensor-core path'):
        _run_core(lay, 8, 32, torch.float32, m.IMPL_TENSOR_CORE)


def test_fully_masked_row_yields_zero_and_finite_gradients():
    import apollo_vision_net_b200.mha as m
    lay = m.sequence_first_layout(40, 2)
    mask = _mask(40, 6)
    mask[3, :] = True                                   # query 3 may attend to nothing
    for dtype, impl in ((torch.float32, 1), (torch.bfloat16, 2)):
        g = torch.Generator().manual_seed(1)
        qk = torch.randn(80, 512, generator=g).to(DEV, dtype).requires_grad_(True)
        v = torch.randn(80, 256, generator=g).to(DEV, dtype).requires_grad_(True)
        out = m.self_attention_core(qk, v, lay, 8, mask.to(DEV), 0.0, impl)
        out.float().square().sum().backward()
        assert torch.all(out.view(40, 2, 256)[3] == 0)
        assert torch.isfinite(out).all() and torch.isfinite(qk.grad).all() and torch.isfinite(v.grad).all()


@pytest.mark.parametrize('dtype,impl', [(torch.float32, 1), (torch.bfloat16, 2), (torch.bfloat16, 1)],
                         ids=['fp32_fma', 'bf16_tensor_core', 'bf16_fma'])
def test_attention_dropout_matches_oracle_with_the_same_mask(dtype, impl):
    """Dropout on the attention weights (nn.MultiheadAttention(dropout=0.1) in training): masks are
    counter-based and recomputed by both backward passes; parity against the oracle with the mask the
    kernels used, plus the mask's statistics and its renewal per step."""
    import apollo_vision_net_b200.mha as m
    lay = m.inter_vector_layout(60, 4, 2)
    p = 0.1
    got, ref, keep = _run_core(lay, 8, 32, dtype, impl, _mask(60, 10), p, seed=3)
    tf, tg = _tols(dtype)
    assert rel_err(got[0], ref[0]) <= tf
    assert rel_err(got[1], ref[1]) <= tg
    assert rel_err(got[2], ref[2]) <= tg
    rate = keep.float().mean().item()
    assert abs(rate - (1 - p)) < 0.01
    _, _, keep2 = _run_core(lay, 8, 32, dtype, impl, _mask(60, 10), p, seed=3)
    agree = (keep == keep2).float().mean().item()
    assert abs(agree - ((1 - p) ** 2 + p ** 2)) < 0.02          # independent masks step to step
    # neighbouring problems / rows / columns are uncorrelated
    k = keep.float() - rate
    for a, b in ((k[:-1], k[1:]), (k[:, :-1], k[:, 1:]), (k[:, :, :-1], k[:, :, 1:]), (k[:, :, :-8], k[:, :, 8:]),
                 (k[:, :-8], k[:, 8:])):
        assert abs((a * b).mean().item()) < 5e-3


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_multihead_attention_module_fused_equals_torch_path(dtype):
    """mmcv-convention MultiheadAttention: the fused path against the same module on torch.nn.MultiheadAttention
    and against the oracle (evaluation mode), outputs and parameter / input gradients; with the next LayerNorm
    folded in."""
    import apollo_vision_net_b200 as pkg
    from apollo_vision_net_b200.rowops import LayerNorm
    S, B, C, H = 50, 4, 256, 8
    attn = pkg.build_attention(dict(type='MultiheadAttention', embed_dims=C, num_heads=H, dropout=0.1))
    with torch.no_grad():
        attn.attn.in_proj_bias.normal_(0, 0.1)
    attn.to(DEV, dtype).eval()
    norm = LayerNorm(C).to(DEV, dtype)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(S, B, C, generator=g)
    pos = torch.randn(S, B, C, generator=g)
    go = torch.randn(S, B, C, generator=g)
    mask = _mask(S, 10)
    results = []
    for fused in ('always', False):
        attn.use_fused_core = fused
        xd = x.to(DEV, dtype).requires_grad_(True)
        pd = pos.to(DEV, dtype).requires_grad_(True)
        launches = pkg.launch_count()
        y = attn(xd, query_pos=pd, attn_mask=mask.to(DEV), post_norm=norm)
        y.backward(go.to(DEV, dtype))
        if fused:
            assert pkg.launch_count() - launches >= 5          # pack, fwd, ln, ln bwd, two bwd passes
        results.append([y, xd.grad, pd.grad, attn.attn.in_proj_weight.grad.clone(),
                        attn.attn.out_proj.weight.grad.clone(), attn.attn.in_proj_bias.grad.clone()])
        attn.zero_grad()
        norm.zero_grad()
    attn.use_fused_core = True
    tol = 2e-5 if dtype == torch.float32 else 3e-2
    for a, b in zip(*results):
        assert rel_err(a, b) <= tol
    if dtype == torch.float32:
        a = attn.attn
        import apollo_vision_net_b200.mha as m
        ref = mha_module_oracle(x.view(-1, C).double(), pos.view(-1, C).double(), x.view(-1, C).double(),
                                a.in_proj_weight.detach().cpu().double(), a.in_proj_bias.detach().cpu().double(),
                                a.out_proj.weight.detach().cpu().double(), a.out_proj.bias.detach().cpu().double(),
                                tuple(m.sequence_first_layout(S, B)), H, mask)
        ref = torch.nn.functional.layer_norm(ref, (C,), norm.weight.detach().cpu().double(),
                                             norm.bias.detach().cpu().double(), norm.eps)
        assert rel_err(results[0][0].view(-1, C), ref) <= 1e-5


def test_module_policy_long_fp32_sequences_take_the_library_path():
    """Default policy: 16-bit -> the core; fp32 -> the core up to 64 tokens, torch's attention beyond (its fp32
    kernels are faster than the FMA path); 'always' forces the core."""
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.mha as m
    attn = pkg.build_attention(dict(type='MultiheadAttention', embed_dims=256, num_heads=8)).to(DEV)
    x32 = torch.zeros(2, 1, 256, device=DEV)
    assert attn.fused_self_attention_ok(x32, m.sequence_first_layout(20, 4), None)
    assert not attn.fused_self_attention_ok(x32, m.sequence_first_layout(350, 4), None)
    attn.use_fused_core = 'always'
    assert attn.fused_self_attention_ok(x32, m.sequence_first_layout(350, 4), None)
    attn.use_fused_core = True
    attn.to(torch.bfloat16)
    assert attn.fused_self_attention_ok(x32.bfloat16(), m.sequence_first_layout(350, 4), None)
    assert not attn.fused_self_attention_ok(x32.bfloat16(), m.sequence_first_layout(350, 4), None,
                                            key_padding_mask=torch.zeros(4, 350, dtype=torch.bool, device=DEV))


@pytest.mark.parametrize('train', [False, True], ids=['eval', 'train_dropout'])
def test_maptrv2_layer_in_place_groupings_equal_the_permuting_path(train):
    """The layer with the fused self-attentions (tokens attended in place, norms folded) against the same
    layer on the reference's permute + contiguous + torch.nn.MultiheadAttention sequence.  In training mode
    only shapes / finiteness / renewal can be compared (different dropout streams)."""
    import apollo_vision_net_b200 as pkg
    C, V, Pn, nb, Hb = 256, 24, 20, 2, 16
    layer = pkg.build_transformer_layer(dict(
        type='MapTRv2DecoupledDetrTransformerDecoderLayer', num_vec=V, num_pts_per_vec=Pn,
        attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                   dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                   dict(type='CustomMSDeformableAttention', embed_dims=C, num_levels=1)],
        feedforward_channels=512, ffn_dropout=0.1,
        operation_order=('self_attn', 'norm', 'self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))
    g = torch.Generator().manual_seed(4)
    for n, prm in layer.named_parameters():
        if n.endswith('sampling_offsets.weight') or n.endswith('attention_weights.weight'):
            prm.data = torch.randn(prm.shape, generator=g) * 0.02
    layer.to(DEV).train(train)
    query = torch.randn(V * Pn, nb, C, generator=g).to(DEV)
    qpos = torch.randn(V * Pn, nb, C, generator=g).to(DEV)
    bev = torch.randn(Hb * Hb, nb, C, generator=g).to(DEV)
    refp = torch.rand(nb, V * Pn, 1, 2, generator=g).to(DEV)
    go = torch.randn(V * Pn, nb, C, generator=g).to(DEV)
    mask = _mask(V, 8).to(DEV)
    shapes = torch.tensor([[Hb, Hb]], device=DEV)
    starts = torch.tensor([0], device=DEV)
    outs = []
    for fused in ('always', False):
        for a in layer.attentions[:2]:
            a.use_fused_core = fused
        q = query.clone().requires_grad_(True)
        y = layer(q, key=None, value=bev, query_pos=qpos, reference_points=refp, spatial_shapes=shapes,
                  level_start_index=starts, self_attn_mask=mask, num_vec=V, num_pts_per_vec=Pn)
        y.backward(go)
        outs.append((y.detach(), q.grad.detach(), layer.attentions[1].attn.in_proj_weight.grad.clone()))
        layer.zero_grad()
    for a in layer.attentions[:2]:
        a.use_fused_core = True
    for a, b in zip(*outs):
        assert torch.isfinite(a).all() and a.shape == b.shape
        if not train:
            assert rel_err(a, b) <= 5e-5
