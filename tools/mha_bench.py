"""Decoder self-attention core (csrc/mha.cu) against torch's scaled_dot_product_attention on the shapes of
BASELINE configs[2] / configs[3]: inter-vector (350 vectors x 20 point slots), intra-vector (20 points x 350
vectors), detection queries (900), bf16, forward and forward + backward, CUDA events, median.
The torch arm gets q / k / v already laid out (groups, heads, S, Dh) -- the reference additionally pays the
permute + contiguous copies around the second self-attention (decoder.py:149-185).
    python tools/mha_bench.py [--iters 30] [--bs 1]"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apollo_vision_net_b200 as pkg  # noqa: E402
import apollo_vision_net_b200.mha as m  # noqa: E402
import apollo_vision_net_b200.rowops as ro  # noqa: E402


def timed(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=30)
    ap.add_argument('--bs', type=int, default=1)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    C, H = 256, 8
    Dh = C // H
    V, Pn, bs = 350, 20, args.bs
    mask = torch.zeros(V, V, dtype=torch.bool, device=dev)
    mask[50:, :50] = True
    mask[:50, 50:] = True
    cases = [('inter_vector_350x20', m.inter_vector_layout(V, Pn, bs), mask),
             ('intra_vector_20x350', m.intra_vector_layout(V, Pn, bs), None),
             ('det_queries_900', m.sequence_first_layout(900, bs), None)]
    for name, lay, msk in cases:
        rows = lay.G * lay.S
        for dtype in (torch.bfloat16, torch.float32):
            for p in (0.0, 0.1):
                qk = torch.randn(rows, 2 * C, device=dev, dtype=dtype, requires_grad=True)
                v = torch.randn(rows, C, device=dev, dtype=dtype, requires_grad=True)
                go = torch.randn(rows, C, device=dev, dtype=dtype)
                q4 = torch.randn(lay.G, H, lay.S, Dh, device=dev, dtype=dtype, requires_grad=True)
                k4 = torch.randn(lay.G, H, lay.S, Dh, device=dev, dtype=dtype, requires_grad=True)
                v4 = torch.randn(lay.G, H, lay.S, Dh, device=dev, dtype=dtype, requires_grad=True)
                go4 = torch.randn(lay.G, H, lay.S, Dh, device=dev, dtype=dtype)
                tmask = None if msk is None else ~msk          # SDPA: True = may attend

                def ours_fwd():
                    with torch.no_grad():
                        return m.SelfAttentionCoreFunction.apply(qk, v, lay, H, m.pack_mask(msk), p, 0)

                def ours_fb():
                    o = m.self_attention_core(qk, v, lay, H, msk, p)
                    o.backward(go)
                    qk.grad = None
                    v.grad = None

                def torch_fwd():
                    with torch.no_grad():
                        return F.scaled_dot_product_attention(q4, k4, v4, attn_mask=tmask, dropout_p=p)

                def torch_fb():
                    o = F.scaled_dot_product_attention(q4, k4, v4, attn_mask=tmask, dropout_p=p)
                    o.backward(go4)
                    q4.grad = k4.grad = v4.grad = None

                res = dict(case=name, dtype=str(dtype).split('.')[-1], dropout=p, problems=lay.G * H, S=lay.S,
                           impl=m.supported_impl(lay, H, Dh, dtype),
                           ours_fwd_us=timed(ours_fwd, args.iters), ours_fwd_bwd_us=timed(ours_fb, args.iters),
                           torch_fwd_us=timed(torch_fwd, args.iters), torch_fwd_bwd_us=timed(torch_fb, args.iters))
                print(json.dumps(res), flush=True)


if __name__ == '__main__':
    main()
