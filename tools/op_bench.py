"""Development micro-benchmark of the op-boundary kernels at BASELINE shapes (CUDA events,
L2 flushed between iterations).  Not the contract benchmark -- that is bench.py."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apollo_vision_net_b200 as pkg  # noqa: E402

LEVELS_BASE = [(116, 200), (58, 100), (29, 50), (15, 25)]
SHAPES = {
    'base_sca': dict(B=6, levels=LEVELS_BASE, M=8, Dh=32, Nq=9690, P=8),
    'base_tsa': dict(B=2, levels=[(200, 200)], M=8, Dh=32, Nq=40000, P=4),
    'tiny_sca': dict(B=6, levels=[(28, 48)], M=8, Dh=32, Nq=606, P=8),
    'maptr_dec': dict(B=1, levels=[(50, 50)], M=8, Dh=32, Nq=7000, P=4),
}


def algo_bytes(B, Nk, M, Dh, L, Nq, P, ev):
    value = B * Nk * M * Dh * ev
    loc = B * Nq * M * L * P * 8
    att = B * Nq * M * L * P * 4
    out = B * Nq * M * Dh * ev
    fwd = value + loc + att + out
    bwd = value + loc + att + out + 2 * (B * Nk * M * Dh * 4) + loc + att
    return fwd, bwd


def make(shape, dtype, dev, local):
    B, levels, M, Dh, Nq, P = (shape[k] for k in ('B', 'levels', 'M', 'Dh', 'Nq', 'P'))
    L = len(levels)
    Nk = sum(h * w for h, w in levels)
    g = torch.Generator(device=dev).manual_seed(0)
    value = torch.randn(B, Nk, M, Dh, generator=g, device=dev).to(dtype)
    if local:   # spatially coherent locations: query i looks near a smooth position
        t = torch.linspace(0, 1, Nq, device=dev).view(1, Nq, 1, 1, 1)
        cx = (t * 37.0) % 1.0
        cy = t
        c = torch.stack([cx, cy], -1).expand(B, Nq, M, L, P, 2)
        loc = c + (torch.rand(B, Nq, M, L, P, 2, generator=g, device=dev) - 0.5) * 0.08
    else:
        loc = torch.rand(B, Nq, M, L, P, 2, generator=g, device=dev) * 1.2 - 0.1
    att = torch.softmax(torch.randn(B, Nq, M, L * P, generator=g, device=dev), -1).view(B, Nq, M, L, P)
    shapes = torch.tensor(levels, dtype=torch.int64, device=dev)
    starts = torch.cat([shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]])
    return value, shapes, starts, loc.contiguous(), att.contiguous(), (B, Nk, M, Dh, L, Nq, P)


def timeit(fn, flush, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shapes', default='base_sca,base_tsa,tiny_sca,maptr_dec')
    ap.add_argument('--dtypes', default='bf16,fp32')
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    peak = 6529.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for name in args.shapes.split(','):
        for dn in args.dtypes.split(','):
            dtype = {'bf16': torch.bfloat16, 'fp32': torch.float32, 'fp16': torch.float16}[dn]
            for local in (False, True):
                value, shapes, starts, loc, att, dims = make(SHAPES[name], dtype, dev, local)
                B, Nk, M, Dh, L, Nq, P = dims
                fb, bb = algo_bytes(*dims, value.element_size())
                out = pkg.ms_deform_attn_forward(value, shapes, starts, loc, att)
                go = torch.randn_like(out)
                gv = torch.zeros(value.shape, dtype=torch.float32, device=dev)
                gl = torch.empty(loc.shape, dtype=torch.float32, device=dev)
                ga = torch.empty(att.shape, dtype=torch.float32, device=dev)
                from apollo_vision_net_b200 import _lib
                from apollo_vision_net_b200.multi_scale_deformable_attn_function import _DTYPE_CODE
                st = torch.cuda.current_stream().cuda_stream

                def fwd():
                    _lib.lib().msda_fwd(value.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                                        loc.data_ptr(), att.data_ptr(), out.data_ptr(), B, Nk, M, Dh,
                                        L, Nq, P, _DTYPE_CODE[dtype], 0, 64, st)

                def bwd():
                    gv.zero_()
                    _lib.lib().msda_bwd(value.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                                        loc.data_ptr(), att.data_ptr(), go.data_ptr(), gv.data_ptr(),
                                        gl.data_ptr(), ga.data_ptr(), B, Nk, M, Dh, L, Nq, P,
                                        _DTYPE_CODE[dtype], 0, 64, st)
                tf = timeit(fwd, flush, args.iters)
                tb = timeit(bwd, flush, args.iters)
                samples = B * Nq * M * L * P
                print(json.dumps(dict(shape=name, dtype=dn, locality='local' if local else 'uniform',
                                      fwd_us=round(tf, 1), bwd_us=round(tb, 1),
                                      fwd_gbs=round(fb / tf / 1e3, 1), bwd_gbs=round(bb / tb / 1e3, 1),
                                      fwd_frac=round(fb / tf / 1e3 / peak, 3),
                                      bwd_frac=round(bb / tb / 1e3 / peak, 3),
                                      fwd_gsamples_s=round(samples / tf / 1e3, 2),
                                      bwd_gsamples_s=round(samples / tb / 1e3, 2))), flush=True)


if __name__ == '__main__':
    main()
