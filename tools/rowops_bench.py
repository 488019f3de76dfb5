"""Development micro-benchmark of the row kernels (csrc/rowops.cu) at the base-encoder shapes: LayerNorm
forward / backward with the fused residual + dropout, the FFN's ReLU + dropout pass and its backward with
the bias sum, the accumulator conversion.  CUDA events, L2 flushed between iterations, C ABI called
directly (APOLLO_B200_LIB selects a variant build).  Prints one JSON line: microseconds (median) and
achieved GB/s over the bytes each kernel has to move."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apollo_vision_net_b200  # noqa: E402,F401
from apollo_vision_net_b200 import _lib  # noqa: E402
from apollo_vision_net_b200.multi_scale_deformable_attn_function import _DTYPE_CODE  # noqa: E402


NB = 4          # buffer sets per kernel: a timed region runs the kernel once on each (event resolution is ~2 us)


def timeit(fn, flush, iters, warmup=2):
    for _ in range(warmup):
        for i in range(NB):
            fn(i)
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(NB):
            fn(i)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3 / NB)
    ts.sort()
    return ts[len(ts) // 2]


def run(rows, iters, dev, flush):
    dt = torch.bfloat16
    code = _DTYPE_CODE[dt]
    C, F = 256, 512
    st = torch.cuda.current_stream(dev).cuda_stream
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*shape):
        return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32).to(dt)

    gamma, beta = rnd(C), rnd(C)
    rng = torch.tensor([1234, 7], dtype=torch.int64, device=dev)
    key = torch.tensor([1234, 7], dtype=torch.int64, device=dev)
    nrows = _lib.lib().rowops_workspace_rows()
    ws = torch.zeros(64 + nrows * 3 * F, dtype=torch.float32, device=dev)
    B = []
    for _ in range(NB):
        b = dict(x=rnd(rows, C), res=rnd(rows, C), dy=rnd(rows, C))
        b.update(s=torch.empty_like(b['x']), y=torch.empty_like(b['x']), dx=torch.empty_like(b['x']),
                 dxm=torch.empty_like(b['x']), mean=torch.zeros(rows, dtype=torch.float32, device=dev),
                 rstd=torch.ones(rows, dtype=torch.float32, device=dev), out3=torch.empty(3, C, dtype=dt, device=dev),
                 h=rnd(rows, F), dh=rnd(rows, F), dhx=torch.empty(rows, F, dtype=dt, device=dev),
                 csum=torch.empty(F, dtype=dt, device=dev), hc=torch.empty(rows, F, dtype=dt, device=dev))
        b['h0'] = b['h'].clone()
        B.append(b)
    P = lambda t: t.data_ptr()  # noqa: E731

    def ln_fwd_drop(i):
        b = B[i]
        _lib.call('ln_residual_dropout_fwd', P(b['x']), P(b['res']), P(gamma), P(beta), P(b['s']), P(b['y']),
                  P(b['mean']), P(b['rstd']), rows, C, 1e-5, code, P(rng), P(key), 3, 0.1, st)

    def ln_fwd_plain(i):
        b = B[i]
        _lib.call('ln_residual_fwd', P(b['x']), P(b['res']), P(gamma), P(beta), P(b['s']), P(b['y']),
                  P(b['mean']), P(b['rstd']), rows, C, 1e-5, code, st)

    def ln_bwd_drop(i):
        b = B[i]
        _lib.call('ln_bwd_dxsum_dropout', P(b['s']), P(b['dy']), P(gamma), P(b['mean']), P(b['rstd']), P(b['dx']),
                  P(b['dxm']), P(b['out3']), P(ws), rows, C, code, P(key), 3, 0.1, st)

    def ln_bwd_plain(i):
        b = B[i]
        _lib.call('ln_bwd_dxsum', P(b['s']), P(b['dy']), P(gamma), P(b['mean']), P(b['rstd']), P(b['dx']),
                  P(b['out3']), P(ws), rows, C, code, st)

    def relu_drop(i):
        _lib.call('relu_dropout_fwd', P(B[i]['h']), rows * F, code, P(rng), P(key), 5, 0.1, st)

    def relu_bwd(i):
        b = B[i]
        _lib.call('relu_bwd_colsum', P(b['dh']), P(b['h0']), P(b['dhx']), P(b['csum']), P(ws), rows, F, code, code,
                  1.0 / 0.9, st)

    sws = torch.zeros(64, dtype=torch.float32, device=dev)

    def amax(i):
        _lib.call('grad_amax_scale', P(B[i]['dy']), rows * C, code, 4.0, P(sws), st)

    def torch_copy(i):
        B[i]['hc'].copy_(B[i]['h0'])

    def torch_relu_(i):
        torch.relu_(B[i]['dh'])

    n1 = rows * C * 2
    n2 = rows * F * 2
    plan = [('ln_residual_dropout_fwd', ln_fwd_drop, 4 * n1), ('ln_residual_fwd', ln_fwd_plain, 4 * n1),
            ('ln_bwd_dxsum_dropout', ln_bwd_drop, 4 * n1), ('ln_bwd_dxsum', ln_bwd_plain, 3 * n1),
            ('relu_dropout_fwd', relu_drop, 2 * n2), ('relu_bwd_colsum', relu_bwd, 3 * n2),
            ('grad_amax_scale', amax, n1), ('torch_copy_rows_x_512', torch_copy, 2 * n2), ('torch_relu_inplace_rows_x_512', torch_relu_, 2 * n2)]
    out = {}
    for name, fn, nbytes in plan:
        us = timeit(fn, flush, iters)
        out[name] = {'us': round(us, 2), 'gbs': round(nbytes / us / 1e3, 1)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', default='40000', help='comma-separated row counts')
    ap.add_argument('--iters', type=int, default=15)
    ap.add_argument('--tag', default=os.environ.get('APOLLO_B200_LIB', 'product'))
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
    for rows in [int(r) for r in args.rows.split(',')]:
        out = {'tag': args.tag, 'rows': rows}
        out.update(run(rows, args.iters, dev, flush))
        torch.cuda.synchronize()
        print(json.dumps(out), flush=True)


if __name__ == '__main__':
    main()
