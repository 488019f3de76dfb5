"""Development micro-benchmark of the fused SCA / TSA kernels at the base-encoder shapes with
the synthetic camera rig (CUDA events, L2 flushed between iterations).  Also the target of the
ncu captures under profiles/."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apollo_vision_net_b200 as pkg  # noqa: E402
import apollo_vision_net_b200.synthetic as syn  # noqa: E402
from apollo_vision_net_b200 import _lib  # noqa: E402
from apollo_vision_net_b200.multi_scale_deformable_attn_function import _DTYPE_CODE  # noqa: E402


def timeit(fn, flush, iters, warmup=3):
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
    ap.add_argument('--bev', type=int, default=200)
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--which', default='sca,tsa')
    ap.add_argument('--accum', default='auto', help="'auto' (fp16 for 16-bit values) or 'fp32'")
    ap.add_argument('--no-tail', action='store_true', help='no replicas of the accumulator tail')
    ap.add_argument('--no-coarse', action='store_true', help='coarse levels through L2 reductions (no tensor-core pass)')
    ap.add_argument('--coord', default='same', help="'same' = offsets/logits in the value dtype, 'fp32'")
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    dtype = {'bf16': torch.bfloat16, 'fp32': torch.float32}[args.dtype]
    cdtype = dtype if args.coord == 'same' else torch.float32
    ccode = _DTYPE_CODE[cdtype]
    H = W = args.bev
    HW = H * W
    M, Dh, C = 8, 32, 256
    levels = syn.LEVELS_BASE
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    L, P, D = len(levels), 8, 4
    g = torch.Generator(device=dev).manual_seed(0)
    l2i, img_shape = syn.camera_rig(1.0)
    from apollo_vision_net_b200.modules import BEVFormerEncoder
    from apollo_vision_net_b200.modules.deform_common import ring_bias
    ref3d = BEVFormerEncoder.get_reference_points(H, W, 8.0, D, dim='3d', bs=1, device=dev, dtype=torch.float32)
    geo = pkg.bev_point_sampling(ref3d, syn.PC_RANGE, l2i, img_shape[0], img_shape[1])
    pairs = int(geo.hit_count.sum())
    shapes = torch.tensor(shapes_l, device=dev)
    starts = torch.tensor(starts_l, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    if 'sca' in args.which:
        value = torch.randn(6, Nk, M, Dh, generator=g, device=dev).to(dtype)
        bias = ring_bias(M, L, P).to(dev).view(1, 1, M, L, P, 2)
        offsets = (bias + 0.3 * torch.randn(1, HW, M, L, P, 2, generator=g, device=dev)).to(cdtype).contiguous()
        logits = torch.randn(1, HW, M, L * P, generator=g, device=dev).to(cdtype)
        slots = torch.empty(1, HW, C, device=dev, dtype=dtype)
        gs = torch.randn(1, HW, C, generator=g, device=dev).to(dtype)
        half_acc = dtype != torch.float32 and args.accum == 'auto'
        gv = torch.zeros(6, Nk, M, Dh, device=dev, dtype=torch.float16 if half_acc else torch.float32)
        sws = torch.zeros(64, device=dev)
        acode = 1 if half_acc else 0
        code = _DTYPE_CODE[dtype]
        rec_bytes = 0 if args.no_coarse else int(_lib.lib().sca_coarse_workspace_bytes(1, 6, HW, M, Dh, P, code))
        records = torch.empty(rec_bytes, dtype=torch.uint8, device=dev) if rec_bytes > 0 else None
        tail = None if True else torch.zeros(7, 6, Nk // 16, M, Dh, device=dev, dtype=torch.float16) \
            if (half_acc and not args.no_tail and records is None) else None
        goff = torch.empty_like(offsets)
        glog = torch.empty_like(logits)
        code = _DTYPE_CODE[dtype]

        def fwd():
            _lib.call('sca_fwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), offsets.data_ptr(),
                      logits.data_ptr(), geo.reference_points_cam.data_ptr(), geo.mask_u8.data_ptr(),
                      geo.hit_bits.data_ptr(), slots.data_ptr(), None, 1, 6, Nk, M, Dh, L, P, D, HW, W, code, ccode, 0, 0, st)

        def bwd():
            if half_acc:
                _lib.call('grad_amax_scale', gs.data_ptr(), gs.numel(), code, float(min(4.0, 32768.0 / HW)), sws.data_ptr(), st)
            _lib.call('sca_bwd', value.data_ptr(), shapes.data_ptr(), starts.data_ptr(), offsets.data_ptr(),
                      logits.data_ptr(), geo.reference_points_cam.data_ptr(), geo.mask_u8.data_ptr(),
                      geo.hit_bits.data_ptr(), gs.data_ptr(), gv.data_ptr(), goff.data_ptr(), glog.data_ptr(),
                      1, 6, Nk, M, Dh, L, P, D, HW, W, code, ccode, 0, 0, acode, sws[16:].data_ptr() if half_acc else None,
                      tail.data_ptr() if tail is not None else None, 7 if tail is not None else 0, Nk // 16,
                      records.data_ptr() if records is not None else None,
                      geo.hit_index.data_ptr() if records is not None else None,
                      geo.hit_count.data_ptr() if records is not None else None, st)
        res['sca_fwd_us'] = round(timeit(fwd, flush, args.iters), 1)
        res['sca_bwd_us'] = round(timeit(bwd, flush, args.iters), 1)
        res['sca_samples'] = pairs * M * L * P
    if 'tsa' in args.which:
        Q, Pt = 2, 4
        value = torch.randn(Q, HW, M, Dh, generator=g, device=dev).to(dtype)
        bias = ring_bias(M, Q, Pt).to(dev).view(1, 1, M, Q, 1, Pt, 2)
        offsets = (bias + 0.3 * torch.randn(1, HW, M, Q, 1, Pt, 2, generator=g, device=dev)).to(cdtype).contiguous()
        logits = torch.randn(1, HW, M, Q, Pt, generator=g, device=dev).to(cdtype)
        ref2d = BEVFormerEncoder.get_reference_points(H, W, dim='2d', bs=1, device=dev, dtype=torch.float32)
        ref = torch.stack([ref2d + 0.004, ref2d], 1).reshape(2, HW, 1, 2).contiguous()
        out = torch.empty(1, HW, C, device=dev, dtype=dtype)
        go = torch.randn(1, HW, C, generator=g, device=dev).to(dtype)
        gv = torch.zeros(Q, HW, M, Dh, device=dev, dtype=torch.float16 if (dtype != torch.float32 and args.accum == 'auto') else torch.float32)
        half_acc = gv.dtype == torch.float16
        sws = torch.zeros(64, device=dev)
        acode = 1 if half_acc else 0
        goff = torch.empty_like(offsets)
        glog = torch.empty_like(logits)
        tshape = torch.tensor([[H, W]], device=dev)
        tstart = torch.tensor([0], device=dev)
        code = _DTYPE_CODE[dtype]

        def tfwd():
            _lib.call('tsa_fwd', value.data_ptr(), tshape.data_ptr(), tstart.data_ptr(), offsets.data_ptr(),
                      logits.data_ptr(), ref.data_ptr(), out.data_ptr(), 1, Q, HW, M, Dh, 1, Pt, HW, W, -1.0, code, ccode, 0, 0, st)

        def tbwd():
            if half_acc:
                _lib.call('grad_amax_scale', go.data_ptr(), go.numel(), code, float(min(4.0, 32768.0 / HW)), sws.data_ptr(), st)
            _lib.call('tsa_bwd', value.data_ptr(), tshape.data_ptr(), tstart.data_ptr(), offsets.data_ptr(),
                      logits.data_ptr(), ref.data_ptr(), go.data_ptr(), gv.data_ptr(), goff.data_ptr(),
                      glog.data_ptr(), 1, Q, HW, M, Dh, 1, Pt, HW, W, -1.0, code, ccode, 0, 0, acode,
                      sws[16:].data_ptr() if half_acc else None, st)
        res['tsa_fwd_us'] = round(timeit(tfwd, flush, args.iters), 1)
        res['tsa_bwd_us'] = round(timeit(tbwd, flush, args.iters), 1)
    res.update(bev=args.bev, dtype=args.dtype, pairs=pairs, coarse=not args.no_coarse)
    print(json.dumps(res), flush=True)


if __name__ == '__main__':
    main()
