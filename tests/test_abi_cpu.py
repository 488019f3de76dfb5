"""CPU-side checks: the C-ABI library builds for sm_100a, loads, and exports every symbol
include/msda_b200.h declares; the product package never falls back to the oracle."""
import os
import re

import pytest


def test_library_builds_loads_and_exports_header_symbols():
    import apollo_vision_net_b200 as pkg
    pkg.build()
    lib = pkg._lib.lib()
    assert lib.msda_abi_version() == pkg._lib.ABI_VERSION == pkg._lib.header_abi_version()
    syms = pkg._lib.header_symbols()
    assert {'msda_fwd', 'msda_bwd', 'bev_point_sampling', 'sca_fwd', 'sca_bwd', 'tsa_fwd',
            'tsa_bwd', 'msda_fwd_host', 'msda_fwd_bwd_host'} <= set(syms)
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/msda_b200.h but not exported'
    assert set(pkg._lib._SIGNATURES) == set(syms)


def test_stale_binary_is_refused(monkeypatch):
    """A binary whose ABI version differs from the ctypes signatures must not be called through them
    (the .so travels outside version control)."""
    import apollo_vision_net_b200 as pkg
    pkg.build()
    monkeypatch.setattr(pkg._lib, '_lib', None)
    monkeypatch.setattr(pkg._lib, 'ABI_VERSION', pkg._lib.ABI_VERSION + 1)
    with pytest.raises(RuntimeError, match='stale binary'):
        pkg._lib.lib()
    monkeypatch.undo()
    pkg._lib._lib = None
    assert pkg._lib.lib().msda_abi_version() == pkg._lib.ABI_VERSION


def test_library_is_sm100a_only():
    import subprocess
    import apollo_vision_net_b200 as pkg
    pkg.build()
    out = subprocess.run(['cuobjdump', '-lelf', pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_\d+a?', out))
    assert archs == {'sm_100a'}, archs


def test_cpu_tensors_raise_no_fallback():
    import torch
    import apollo_vision_net_b200 as pkg
    from tests.util import make_op_inputs
    args = make_op_inputs(1, [(4, 4)], 2, 32, 3, 2)
    with pytest.raises(RuntimeError, match='CUDA tensor'):
        pkg.MultiScaleDeformableAttnFunction_fp32.apply(*args, 64)


def test_product_package_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg_dir = os.path.join(root, 'apollo-vision-net_b200')
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f


def test_host_scratch_size_is_consistent():
    import apollo_vision_net_b200 as pkg
    lib = pkg._lib.lib()
    n_fwd = lib.msda_host_scratch_bytes(2, 100, 8, 32, 1, 50, 4, 0, 0, 0)
    n_bwd = lib.msda_host_scratch_bytes(2, 100, 8, 32, 1, 50, 4, 0, 0, 1)
    assert 0 < n_fwd < n_bwd
    assert n_fwd >= 2 * 100 * 256 * 4 + 2 * 50 * 8 * 4 * 12 + 2 * 50 * 256 * 4


def test_hoisted_decoder_value_projection_equals_per_layer_linears():
    """hoist_value_proj (one batched GEMM before the decoder's layer loop) returns what each
    layer's own ``value_proj`` returns (reference decoder.py:299-303), with gradients for the BEV
    and for every layer's parameters.  Host logic only: plain torch on the CPU."""
    import torch
    from apollo_vision_net_b200.modules.decoder import CustomMSDeformableAttention, hoist_value_proj
    torch.manual_seed(5)
    attns = [CustomMSDeformableAttention(embed_dims=64, num_heads=4, num_levels=1) for _ in range(3)]
    for a in attns:
        torch.nn.init.normal_(a.value_proj.bias, std=0.5)
    bev = torch.randn(35, 2, 64, dtype=torch.float64)
    for a in attns:
        a.double()
    b1 = bev.clone().requires_grad_(True)
    ref = [a.value_proj(b1.permute(1, 0, 2)) for a in attns]
    go = [torch.randn_like(r) for r in ref]
    torch.autograd.backward(ref, go)
    ref_grads = [(a.value_proj.weight.grad.clone(), a.value_proj.bias.grad.clone()) for a in attns]
    for a in attns:
        a.zero_grad()
    b2 = bev.clone().requires_grad_(True)
    out = hoist_value_proj(attns, b2)
    assert len(out) == 3 and all(o.shape == (2, 35, 64) for o in out)
    torch.autograd.backward(out, go)
    for o, r in zip(out, ref):
        assert torch.allclose(o, r, rtol=1e-12, atol=1e-12)
    assert torch.allclose(b2.grad, b1.grad, rtol=1e-12, atol=1e-12)
    for a, (gw, gb) in zip(attns, ref_grads):
        assert torch.allclose(a.value_proj.weight.grad, gw, rtol=1e-12, atol=1e-12)
        assert torch.allclose(a.value_proj.bias.grad, gb, rtol=1e-12, atol=1e-12)
