"""DCNv3 operator timing at InternImage-like backbone shapes (six camera images of one frame), forward and
backward, CUDA events, L2 flushed between iterations; achieved bandwidth against the algorithmic bytes
(input + offset + mask + output; backward adds grad_output, the fp32 input-gradient accumulator written and
read, and the offset / mask gradients).
    python tools/dcnv3_bench.py [--iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apollo_vision_net_b200.dcnv3 as d  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
    shapes = [('stage1_116x200_c64', 6, 116, 200, 4, 16), ('stage2_58x100_c128', 6, 58, 100, 8, 16),
              ('stage3_29x50_c256', 6, 29, 50, 16, 16), ('stage1_232x400_c64', 6, 232, 400, 4, 16)]
    for name, N, H, W, G, Cg in shapes:
        for dtype in (torch.bfloat16, torch.float32):
            K = 9
            g = torch.Generator().manual_seed(0)
            x = torch.randn(N, H, W, G * Cg, generator=g).to(dev, dtype)
            off = torch.randn(N, H, W, G * K * 2, generator=g).to(dev, dtype)
            msk = torch.softmax(torch.randn(N, H, W, G, K, generator=g), -1).reshape(N, H, W, G * K).to(dev, dtype)
            go = torch.randn(N, H, W, G * Cg, generator=g).to(dev, dtype)
            cfg = (3, 3, 1, 1, 1, 1, 1, 1, G, Cg, 1.0)

            def timed(fn):
                for _ in range(3):
                    fn()
                ts = []
                for _ in range(args.iters):
                    flush.zero_()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    fn()
                    e.record()
                    torch.cuda.synchronize()
                    ts.append(s.elapsed_time(e) * 1e3)
                ts.sort()
                return ts[len(ts) // 2]

            es = x.element_size()
            fwd_bytes = (x.numel() + off.numel() + msk.numel() + go.numel()) * es
            bwd_bytes = fwd_bytes + 2 * x.numel() * 4 + (off.numel() + msk.numel()) * 4
            tf = timed(lambda: d.dcnv3_forward(x, off, msk, *cfg))
            tb = timed(lambda: d.dcnv3_backward(x, off, msk, *cfg, go))
            print(json.dumps(dict(shape=name, dtype=str(dtype).split('.')[-1], samples=N * H * W * G * K,
                                  fwd_us=round(tf, 1), bwd_us=round(tb, 1),
                                  fwd_gbs=round(fwd_bytes / tf / 1e3, 1), bwd_gbs=round(bwd_bytes / tb / 1e3, 1))),
                  flush=True)


if __name__ == '__main__':
    main()
