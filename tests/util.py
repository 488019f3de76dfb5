"""Shared helpers for the parity tests (seeded synthetic inputs, tolerances)."""
import numpy as np
import torch


def make_op_inputs(B, levels, M, Dh, Nq, P, seed=0, dtype=torch.float32, lo=-0.1, hi=1.1):
    """value ~ N(0,1); locations ~ U(lo, hi) (about 17% out of range per axis);
    weights = softmax(N(0,1)) over L*P -- SURVEY.md section 8d."""
    g = torch.Generator().manual_seed(seed)
    L = len(levels)
    Nk = sum(h * w for h, w in levels)
    value = torch.randn(B, Nk, M, Dh, generator=g, dtype=torch.float32)
    loc = torch.rand(B, Nq, M, L, P, 2, generator=g, dtype=torch.float32) * (hi - lo) + lo
    att = torch.softmax(torch.randn(B, Nq, M, L * P, generator=g), -1).view(B, Nq, M, L, P)
    shapes = torch.tensor(levels, dtype=torch.int64).view(L, 2)
    starts = torch.cat([shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]])
    return value.to(dtype), shapes, starts, loc, att


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny): error relative to the scale of the reference tensor."""
    a = torch.as_tensor(np.asarray(a.detach().float().cpu() if isinstance(a, torch.Tensor) else a),
                        dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach().float().cpu() if isinstance(b, torch.Tensor) else b),
                        dtype=torch.float64)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2: insensitive to a handful of outliers (used where a few flipped
    samples move single entries of a deep gradient)."""
    a = torch.as_tensor(np.asarray(a.detach().float().cpu()), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach().float().cpu()), dtype=torch.float64)
    if b.numel() == 0:
        return 0.0
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
