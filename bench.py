#!/usr/bin/env python
"""Benchmark of the hot path: BEVFormer-base encoder forward+backward on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A *step* is one pass of the hot path over one frame: the 6-layer BEVFormer-base encoder
(BASELINE.json configs[1]: 200x200 BEV, 6 cameras, 4 feature levels, 8 points x 4 Z-anchors,
bf16) forward + backward -- temporal self-attention, spatial cross-attention, LayerNorm, FFN,
with the image features requiring grad.  The metric is sampled points per second: the
(batch, query, head, level, point) bilinear samples the deformable attention takes in one
forward pass of the step (SCA over the cameras that see each query + TSA), divided by the
fwd+bwd step time.  `value` is measured with every input resident in HBM; `e2e` is the same
step driven from pinned HOST buffers (H2D of the frame's inputs, D2H of the BEV output and the
loss inside the timed region).  With N > 1 every rank runs its own frame (batch-level data
parallelism, the reference's only strategy, apis/mmdet_train.py:71-85) and the parameter
gradients are all-reduced over NCCL each step.

`--impl reference` times the reference's own CPU implementation of the same path (the oracle
port of its CPU branch: multi_scale_deformable_attn_pytorch inside SCA/TSA, fp32, all host
threads) on a bounded sample of the workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'BEVFormer-base encoder MSDeformAttn fwd+bwd sampled-points/s'
UNIT = 'points/s'
C, HEADS, FFN_DIM, SCA_POINTS, TSA_POINTS, PILLAR = 256, 8, 512, 8, 4, 4


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# --------------------------------------------------------------------------- workload ------
def encoder_cfg(num_layers, num_levels, pc_range, dropout=0.1):
    """The reference's training values: dropout 0.1 in TSA / SCA (temporal_self_attention.py:54-66,
    spatial_cross_attention.py:43-61) and ffn_dropout 0.1 (bev_base_occ.py:127)."""
    return dict(
        type='BEVFormerEncoder', num_layers=num_layers, pc_range=pc_range,
        num_points_in_pillar=PILLAR, return_intermediate=False,
        transformerlayers=dict(
            type='BEVFormerLayer',
            attn_cfgs=[
                dict(type='TemporalSelfAttention', embed_dims=C, num_points=TSA_POINTS, num_levels=1,
                     dropout=dropout),
                dict(type='SpatialCrossAttention', pc_range=pc_range, embed_dims=C, dropout=dropout,
                     deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=C,
                                               num_points=SCA_POINTS, num_levels=num_levels))],
            feedforward_channels=FFN_DIM, ffn_dropout=dropout,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm')))


def randomize(module, seed):
    """Reference init, then N(0, 0.02) on the offset / weight Linears so that offsets and the
    softmax are non-degenerate."""
    g = torch.Generator().manual_seed(seed)
    for n, p in module.named_parameters():
        if n.endswith('sampling_offsets.weight') or n.endswith('attention_weights.weight'):
            p.data = (torch.randn(p.shape, generator=g) * 0.02).to(p.dtype)


def host_inputs(bev_h, bev_w, levels, seed, dtype, pin):
    import apollo_vision_net_b200.synthetic as syn
    g = torch.Generator().manual_seed(seed)
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    HW = bev_h * bev_w

    def mk(*shape):
        t = torch.randn(*shape, generator=g).to(dtype)
        return t.pin_memory() if pin else t
    return dict(feat=mk(6, Nk, 1, C), bev_query=mk(HW, 1, C), bev_pos=mk(HW, 1, C),
                prev_bev=mk(HW, 1, C), grad_w=mk(1, HW, C),
                shift=torch.tensor([[0.005, 0.003]]),
                shapes=torch.tensor(shapes_l, dtype=torch.int64),
                starts=torch.tensor(starts_l, dtype=torch.int64))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed regions run."""
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.FIELDS}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower() == 'active':
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons),
                    samples=len(sm))


def leave(dist):
    """End of a multi-rank run.  Captured CUDA graphs that contain NCCL launches keep the communicator busy:
    destroy_process_group() after them was seen to hang (N = 2, driver 580.159) after the JSON line had been
    printed.  Every rank has finished its work once the barrier returns, so the processes exit directly."""
    sys.stdout.flush()
    sys.stderr.flush()
    try:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        os._exit(0)


# --------------------------------------------------------------------------- B200 arm ------
def run_b200(args):
    import torch.distributed as dist
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.synthetic as syn
    from apollo_vision_net_b200 import _lib

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    if args.gpus > 1 and world == 1:
        raise SystemExit('launch N > 1 with torch.distributed.run (one rank per GPU)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    pkg.build()

    dtype = torch.bfloat16
    bev_h = bev_w = args.bev
    levels = syn.LEVELS_BASE
    enc = pkg.build_transformer_layer_sequence(encoder_cfg(args.layers, len(levels), syn.PC_RANGE, args.dropout))
    randomize(enc, 0)
    enc.to(dev).to(dtype).train()                 # training mode: the reference's dropouts are active
    params = [p for p in enc.parameters() if p.requires_grad]
    l2i, img_shape = syn.camera_rig(1.0, bs=1)
    l2i_dev = torch.as_tensor(l2i).to(dev)
    # Batch-level data parallelism as DDP does it (bevformer/apis/mmdet_train.py:71-85): bucketed
    # all-reduce of the parameter gradients launched from autograd hooks while the backward is still
    # running; inside the captured step the NCCL launches are nodes of the CUDA graph.
    from apollo_vision_net_b200.parallel import BucketedGradReducer
    reducer = BucketedGradReducer(params, bucket_bytes=args.bucket_mb << 20) if world > 1 else None
    ddp = {'mode': 'single GPU' if world == 1 else
           f'bucketed all-reduce overlapped with the backward ({len(reducer.buckets)} buckets of '
           f'~{args.bucket_mb} MiB, averaged gradients left in p.grad)'}

    host = host_inputs(bev_h, bev_w, levels, seed=1 + rank, dtype=dtype, pin=True)
    devin = {k: v.to(dev) for k, v in host.items()}
    shapes, starts = devin['shapes'], devin['starts']
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    use_reducer = [reducer is not None]

    def encoder_step(feat, bev_query, bev_pos, prev_bev, grad_w, shift):
        for p in params:
            p.grad = None
        if reducer is not None:
            reducer.reset()
            reducer.enabled = use_reducer[0]
        feat.requires_grad_(True)
        out = enc(bev_query, feat, feat, bev_h=bev_h, bev_w=bev_w, bev_pos=bev_pos,
                  spatial_shapes=shapes, level_start_index=starts, prev_bev=prev_bev, shift=shift,
                  lidar2img=l2i_dev, img_shape=img_shape)
        loss = (out.float() * grad_w.float()).sum() * (1.0 / out.numel())
        loss.backward()
        if reducer is not None and use_reducer[0]:
            reducer.finish()
        return out, loss

    def flat_allreduce():                          # fallback when NCCL cannot be captured into the graph
        if world > 1 and not use_reducer[0]:
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)

    def step_eager():
        flush.zero_()                              # L2 flushed between iterations (inside the timed region)
        r = encoder_step(devin['feat'].detach(), devin['bev_query'], devin['bev_pos'],
                         devin['prev_bev'], devin['grad_w'], devin['shift'])
        flat_allreduce()
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_eager()
    # Per-kernel durations for the roofline object: the same step run eagerly (a replayed graph
    # cannot carry timing events), K steps, CUDA events on the launching stream around each of our
    # launches, L2 flushed every step exactly as in the graph-replay timed region below.
    clocks = ClockSampler(local)
    clocks.__enter__()                             # sampled over every timed region below
    timer = _lib.KernelTimer()
    launches0 = _lib.launch_count()
    _lib.set_timer(timer)
    ms_eager = timed(step_eager, args.steps)
    _lib.set_timer(None)
    launches = _lib.launch_count() - launches0
    kern = timer.summary()

    # The encoder never synchronises with the host (no nonzero(), no max_len), so the whole
    # forward + backward step is captured once into a CUDA graph and replayed: the 540 launches
    # of a step cost one graph launch on the host.  At N > 1 the bucketed gradient all-reduces are
    # captured with it (NCCL launches as graph nodes on a parallel branch).
    graph, static_out = None, {}
    for attempt in range(2 if (world > 1 and not args.no_graph) else (0 if args.no_graph else 1)):
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step_eager()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for p in params:
                p.grad = None
            graph = torch.cuda.CUDAGraph()
            # (thread-local capture mode: NCCL's watchdog thread may touch the device meanwhile)
            with torch.cuda.graph(graph, capture_error_mode='thread_local' if world > 1 else 'global'):
                o, l = encoder_step(devin['feat'].detach(), devin['bev_query'], devin['bev_pos'],
                                    devin['prev_bev'], devin['grad_w'], devin['shift'])
                static_out['out'], static_out['loss'] = o.detach(), l.detach()
            torch.cuda.synchronize()
            break
        except Exception as exc:                   # pragma: no cover - reported in the JSON line
            graph = None
            static_out['error'] = f'{type(exc).__name__}: {exc}'[:200]
            torch.cuda.synchronize()
            if world > 1 and use_reducer[0]:       # second attempt: collectives outside the graph
                use_reducer[0] = False
                ddp['mode'] = ('flat all-reduce after the graph replay (capturing the bucketed NCCL launches '
                               'failed: ' + static_out['error'][:80] + ')')

    def step_device():
        if graph is None:
            return step_eager()
        flush.zero_()
        graph.replay()
        flat_allreduce()
        return static_out['out'], static_out['loss']

    out_host = torch.empty((1, bev_h * bev_w, C), dtype=dtype).pin_memory()
    # Per-frame inputs travel from the host: the six cameras' feature maps and the ego-motion shift.
    # bev_query / bev_pos are parameters (bev_embedding, positional encoding), prev_bev is the previous
    # frame's output and grad_w stands for the decoders' upstream gradient: all live on the device in
    # the real model (dense_heads/bevformer_head.py:129-225) and stay resident here.
    h2d_keys = ('feat', 'shift')

    # End-to-end pipeline: the frame's inputs travel from pinned host memory over PCIe on a copy
    # stream into a staging set of device buffers while the previous frame computes; the compute
    # stream then moves them into the graph's static inputs (device-to-device), replays the step
    # and reads the BEV output and the loss back.  Every step's H2D and D2H happen inside the timed
    # region; only their overlap with the neighbouring step's compute is what the pipeline adds.
    copy_stream = torch.cuda.Stream()
    staging = {k: torch.empty_like(devin[k]) for k in h2d_keys}
    staged = torch.cuda.Event()
    consumed = torch.cuda.Event()
    pipe = {'primed': False, 'left': 1 << 30}

    d2h_stream = torch.cuda.Stream()
    computed = torch.cuda.Event()
    read_back = torch.cuda.Event()
    read_back.record(torch.cuda.current_stream())
    back = {'out': None, 'loss': None, 'loss_host': None, 'pending': False, 'value': None}

    def collect():
        read_back.synchronize()
        back['value'] = float(back['loss_host'])
        back['pending'] = False

    def issue_h2d():
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed)
            for k in h2d_keys:
                staging[k].copy_(host[k], non_blocking=True)
            staged.record(copy_stream)

    def step_e2e():
        main = torch.cuda.current_stream()
        flush.zero_()
        if not pipe['primed']:                     # first frame: nothing to overlap with
            consumed.record(main)
            issue_h2d()
            pipe['primed'] = True
        main.wait_event(staged)
        if graph is None:
            d = dict(devin, **{k: staging[k].clone() for k in h2d_keys})
        else:
            for k in h2d_keys:
                devin[k].copy_(staging[k], non_blocking=True)
            d = devin
        consumed.record(main)
        pipe['left'] -= 1
        if pipe['left'] > 0:
            issue_h2d()                            # next frame's inputs, overlapped with this step
        else:
            pipe['primed'] = False
        if graph is None:
            out, loss = encoder_step(d['feat'], d['bev_query'], d['bev_pos'], d['prev_bev'],
                                     d['grad_w'], d['shift'])
        else:
            graph.replay()
            out, loss = static_out['out'], static_out['loss']
        flat_allreduce()
        # D2H of the step's results on its own stream, one step behind the compute: the BEV output
        # and the loss are first moved (device-to-device) into staging buffers, so the next replay
        # may overwrite the graph's static outputs while PCIe is still carrying these; the host
        # reads step i's loss after it has queued step i+1.  The last step of a run drains the
        # pipeline before it returns, so every step's read-back lies inside the timed region.
        if back['out'] is None:
            back['out'] = torch.empty_like(out.detach())
            back['loss'] = torch.empty_like(loss.detach())
            back['loss_host'] = torch.empty(loss.shape, dtype=loss.dtype).pin_memory()
        main.wait_event(read_back)                 # the previous D2H has left the staging buffers
        back['out'].copy_(out.detach(), non_blocking=True)
        back['loss'].copy_(loss.detach(), non_blocking=True)
        computed.record(main)
        if back['pending']:
            collect()                              # host: result of the PREVIOUS step
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(computed)
            out_host.copy_(back['out'], non_blocking=True)
            back['loss_host'].copy_(back['loss'], non_blocking=True)
            read_back.record(d2h_stream)
        back['pending'] = True
        if not pipe['primed']:                     # last step of the run: drain
            main.wait_event(read_back)
            collect()
        return back['value']

    # sampled points of one forward pass (what the metric counts)
    ref3d = enc.get_reference_points(bev_h, bev_w, syn.PC_RANGE[5] - syn.PC_RANGE[2], PILLAR,
                                     dim='3d', bs=1, device=dev, dtype=torch.float32)
    geo = pkg.bev_point_sampling(ref3d, syn.PC_RANGE, l2i, img_shape[0], img_shape[1])
    pairs = int(geo.hit_count.sum().item())
    HW = bev_h * bev_w
    sca_points = pairs * HEADS * len(levels) * SCA_POINTS
    tsa_points = 2 * HW * HEADS * TSA_POINTS
    points_per_step = args.layers * (sca_points + tsa_points)

    for _ in range(args.warmup):
        step_device()
    ms = timed(step_device, args.steps)
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    torch.cuda.synchronize()
    pipe['primed'] = False                         # the timed region starts with an empty pipeline
    pipe['left'] = args.steps                      # ... and issues exactly `steps` H2D copies
    ms_e2e = timed(step_e2e, args.steps)
    # the loss the host read back from the last e2e step must be the loss the device computed
    # (with dropout every step draws new masks: compare with the device's copy of that same step's loss)
    loss_dev = float(back['loss'].float().item()) if back['loss'] is not None else float('nan')
    loss_read = back['value']
    read_ok = loss_read is not None and abs(loss_read - loss_dev) <= 1e-3 * abs(loss_dev) + 1e-30
    if not read_ok:
        print(f'WARNING: e2e read-back {loss_read} differs from the device loss {loss_dev}',
              file=sys.stderr, flush=True)
    clocks.__exit__(None, None, None)

    ms_per_step = ms / args.steps
    value = world * points_per_step * args.steps / (ms / 1e3)
    e2e_value = world * points_per_step * args.steps / (ms_e2e / 1e3)
    h2d = sum(host[k].numel() * host[k].element_size() for k in h2d_keys)
    d2h = out_host.numel() * out_host.element_size() + 4

    # ---- roofline of the dominant kernel (largest share of the step among our launches) ----
    peak, peak_src = measured_peak()
    Nk = sum(h * w for h, w in levels)
    L, P, M, Dh = len(levels), SCA_POINTS, HEADS, C // HEADS
    ev = 2                       # bf16 value / outputs / upstream gradients
    ec = 2                       # offsets / logits and their gradients arrive in the model dtype (bf16)
    ea = 4 if pkg.fused_ops._accum_mode[0] == 'fp32' else 2      # grad_value accumulator
    # compulsory traffic, every operand once: value + offsets + logits + ref_cam +
    # hit bits + out; backward adds g_out (= |out|), 2 x accumulator (zero-fill + final), g_offsets, g_logits
    sca_fwd_bytes = (6 * Nk * C * ev + HW * M * L * P * 3 * ec + 6 * HW * PILLAR * 2 * 4 + HW * 4
                     + HW * C * ev)
    sca_bwd_bytes = sca_fwd_bytes + 2 * 6 * Nk * C * ea + HW * M * L * P * 3 * ec
    tsa_s = 2 * 1 * TSA_POINTS
    tsa_fwd_bytes = 2 * HW * C * ev + HW * M * tsa_s * 3 * ec + 2 * HW * 2 * 4 + HW * C * ev
    tsa_bwd_bytes = tsa_fwd_bytes + 2 * 2 * HW * C * ea + HW * M * tsa_s * 3 * ec
    algo = dict(sca_fwd=sca_fwd_bytes, sca_bwd=sca_bwd_bytes, tsa_fwd=tsa_fwd_bytes,
                tsa_bwd=tsa_bwd_bytes)
    kernels = {}
    for name, st in kern.items():
        row = dict(launches_per_step=st['launches'] / args.steps, mean_us=round(st['mean_us'], 2))
        if name in algo:
            row['algorithmic_bytes'] = algo[name]
            row['achieved_gbs'] = round(algo[name] / st['mean_us'] / 1e3, 1)
            row['frac_of_measured_peak'] = round(algo[name] / st['mean_us'] / 1e3 / peak, 4)
        row['share_of_step'] = round(st['mean_us'] * st['launches'] / args.steps / (ms_eager / args.steps * 1e3), 4)
        kernels[name] = row
    dominant = max((k for k in kernels if k in algo), key=lambda k: kernels[k]['share_of_step'])
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')) as f:
            traffic = json.load(f).get(dominant)
    except Exception:
        pass
    roofline = dict(bound='hbm', kernel=dominant, achieved=kernels[dominant]['achieved_gbs'],
                    peak=peak, peak_source=peak_src, unit='GB/s',
                    frac=kernels[dominant]['frac_of_measured_peak'], traffic=traffic,
                    algorithmic_bytes_per_launch=algo[dominant],
                    mean_launch_us=kernels[dominant]['mean_us'])

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_reference_sample(steps=6, warmup=1)     # about 12 s of CPU work on the box

    extras = {}
    if not args.no_extras:
        # the north_star's own partition (BASELINE configs[4]: at N = 1 the unsharded frame, the reference of
        # the strong-scaling sweep), the det+map stand-in (configs[2]) and the MapTRv2 decoder (configs[3]) ride
        # on the same line as extra keys; the headline metric above is untouched
        graph_ok = graph is not None
        del graph
        graph = True if graph_ok else None
        torch.cuda.empty_cache()
        for name, fn in (('rowshard', lambda: measure_rowshard(args, rank, world, dev, max(5, args.steps // 2), 3,
                                                                bev=400, train_steps=3)),
                         ('detmap', lambda: measure_detmap(args, rank, world, dev, max(5, args.steps // 2), 3)),
                         ('maptrv2_decoder', lambda: measure_detmap(args, rank, world, dev, max(5, args.steps // 2), 3,
                                                                    decoder_only=True))):
            try:
                extras[name] = fn()
            except Exception as exc:               # pragma: no cover - reported in the JSON line
                extras[name] = {'error': f'{type(exc).__name__}: {exc}'[:300]}
                torch.cuda.synchronize()
            torch.cuda.empty_cache()
    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {
                'workload': f'BEVFormer-base encoder fwd+bwd, {bev_h}x{bev_w} BEV, 6 cams, 4 levels '
                            f'(116x200..15x25), 8 points x 4 Z-anchors, {args.layers} layers, '
                            'bs=1 frame per GPU, bf16 value / fp32 sampling arithmetic (BASELINE configs[1])',
                'points_per_step': points_per_step, 'sca_points_per_layer': sca_points,
                'tsa_points_per_layer': tsa_points, 'camera_query_pairs': pairs,
                'parallelism': f'dp{world}' if world > 1 else 'single',
                'ddp': ddp['mode'],
                'l2': 'flushed every step (192 MiB write inside the timed region); the step\'s '
                      'working set (> 1 GB of activations) also exceeds the 126 MB L2',
                'weights': 'random-init (reference init + N(0,0.02) offset/weight Linears)',
                'dropout': args.dropout,
                'dropout_note': 'training mode, the reference\'s dropout values (TSA / SCA 0.1, FFN 0.1 twice); masks '
                                'are drawn inside the fused row kernels and recomputed in the backward; the step counter '
                                'advances on the device, so every graph replay draws new masks',
                'cuda_graph': graph is not None,
                'cuda_graph_error': static_out.get('error'),
                'eager_ms_per_step': ms_eager / args.steps,
            },
            'frames_per_s': world * args.steps / (ms / 1e3),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps,
                    'h2d': 'per frame: 6-camera feature maps + ego-motion shift; device-resident: bev_query / '
                           'bev_pos (parameters), prev_bev (previous output), upstream gradient',
                    'loss_read_back': loss_read, 'loss_read_back_matches_device': read_ok},
            'gpu_launches': launches,
            'kernels': kernels,
            'roofline': roofline,
            'cpu_baseline': cpu_baseline,
            'clocks': clocks.summary(),
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        leave(dist)


# ------------------------------------------------------------------ BEV row sharding ------
def measure_rowshard(args, rank, world, dev, steps, warmup, bev=400, train_steps=0):
    """BASELINE configs[4]: the encoder at a 400x400 BEV with the query rows sharded over the ranks.

    Per frame and rank, inside the timed region: H2D of this rank's 1/N slice of the six cameras'
    feature maps (pinned host memory), ONE NCCL all-gather that gives every rank the full maps over
    NVLink (the "value features broadcast once per frame": no rank pulls more than 1/N of the frame over
    PCIe), the shard's forward (a replayed CUDA graph), ONE all-gather of the BEV rows.  Also measured:
    the same frame on one rank alone (unsharded, the strong-scaling reference) and, with
    ``train_steps``, forward + backward with the gradients of the replicated tensors and of the
    parameters summed over the row group (parallel.sharded_encoder_forward / allreduce_gradients)."""
    import torch.distributed as dist
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.synthetic as syn
    from apollo_vision_net_b200 import _lib
    from apollo_vision_net_b200.parallel import (all_gather_bev_rows, allreduce_gradients,
                                                 sharded_encoder_forward)
    dtype = torch.bfloat16
    levels = syn.LEVELS_BASE
    enc = pkg.build_transformer_layer_sequence(encoder_cfg(args.layers, len(levels), syn.PC_RANGE, 0.0))
    randomize(enc, 0)
    enc.to(dev).to(dtype).eval()
    l2i, img_shape = syn.camera_rig(1.0, bs=1)
    l2i_dev = torch.as_tensor(l2i).to(dev)
    host = host_inputs(bev, bev, levels, seed=1, dtype=dtype, pin=False)   # the same frame on every rank
    d = {k: v.to(dev) for k, v in host.items()}
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
    shard = (rank, world) if world > 1 else ((0, args.shard_sim) if args.shard_sim > 1 else None)
    # this rank's slice of the frame's features on the host (flat split: 6 cameras / N ranks)
    feat_flat = d['feat'].view(-1)
    n_feat = feat_flat.numel()
    chunk = (n_feat + world - 1) // world
    pad = torch.zeros(chunk * world, dtype=dtype)
    pad[:n_feat] = host['feat'].view(-1)
    my_host = pad[rank * chunk:(rank + 1) * chunk].clone().pin_memory()
    my_dev = [torch.empty(chunk, dtype=dtype, device=dev) for _ in range(2)]   # double-buffered staging
    gathered = torch.empty(chunk * world, dtype=dtype, device=dev)
    feat_static = d['feat']
    copy_stream = torch.cuda.Stream()
    staged = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    frame = {'i': 0, 'primed': False}

    def kwargs_of(feat):
        return dict(bev_h=bev, bev_w=bev, bev_pos=d['bev_pos'], spatial_shapes=d['shapes'],
                    level_start_index=d['starts'], prev_bev=d['prev_bev'], shift=d['shift'],
                    lidar2img=l2i_dev, img_shape=img_shape)

    def shard_forward(row_shard):
        with torch.no_grad():
            return enc(d['bev_query'], feat_static, feat_static, row_shard=row_shard, **kwargs_of(feat_static))

    def capture(fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        torch.cuda.synchronize()
        return g, out

    graph, static_out, graph_error = None, None, None
    if not args.no_graph:
        try:
            graph, static_out = capture(lambda: shard_forward(shard))
        except Exception as exc:                   # pragma: no cover - reported in the JSON line
            graph, graph_error = None, f'{type(exc).__name__}: {exc}'[:200]
            torch.cuda.synchronize()

    def issue_h2d(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            my_dev[slot].copy_(my_host, non_blocking=True)              # 1/N of the frame over PCIe
            staged[slot].record(copy_stream)

    # Frame pipeline (N > 1).  Everything a frame needs from outside the GPU is on its way while the frame
    # before it computes: its slice over PCIe (copy stream, double buffer), the other ranks' slices over
    # NVLink (asynchronous all-gather on NCCL's stream); and the frame's own BEV rows leave through an
    # asynchronous all-gather that overlaps the next frame.  Every frame still pays one H2D, one feature
    # all-gather and one row all-gather inside the timed region; the run is drained before the clock stops.
    pend = {'feat': None, 'rows': None}
    rows_send = {'buf': None, 'recv': None}

    def start_feat_gather():
        main = torch.cuda.current_stream()
        slot = frame['i'] & 1
        main.wait_event(staged[slot])
        pend['feat'] = (dist.all_gather_into_tensor(gathered, my_dev[slot], async_op=True), slot)
        frame['i'] += 1

    def take_features():
        main = torch.cuda.current_stream()
        if not frame['primed']:
            for sl in (0, 1):
                consumed[sl].record(main)
                issue_h2d(sl)
            start_feat_gather()
            frame['primed'] = True
        work, slot = pend['feat']
        work.wait()
        feat_flat.copy_(gathered[:n_feat])                              # (into the graph's static input)
        consumed[slot].record(main)
        issue_h2d(slot)                                                 # the slice of the frame after next
        start_feat_gather()                                             # the next frame's slices, during this frame

    def send_rows(out):
        if rows_send['buf'] is None:
            rows_send['buf'] = torch.empty_like(out)
            rows_send['recv'] = out.new_empty((world,) + tuple(out.shape))
        if pend['rows'] is not None:
            pend['rows'].wait()                                         # the previous frame's rows have arrived
        rows_send['buf'].copy_(out)                                     # the next replay may overwrite `out`
        pend['rows'] = dist.all_gather_into_tensor(rows_send['recv'], rows_send['buf'], async_op=True)

    def drain():
        if pend['rows'] is not None:
            pend['rows'].wait()
            pend['rows'] = None

    even = bev % world == 0

    def step():
        flush.zero_()
        if world > 1:
            take_features()
        if graph is None:
            out = shard_forward(shard)
        else:
            graph.replay()
            out = static_out
        if world > 1:
            if even:
                send_rows(out)
            else:
                out = all_gather_bev_rows(out, bev, bev)
        return out

    def timed(fn, n, after=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        if after is not None:
            after()
        e.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n

    for _ in range(max(warmup, 3)):
        step()
    drain()
    n0 = _lib.launch_count()
    if graph is not None:
        shard_forward(shard)                       # our launches per frame, counted on an eager pass
    launches = _lib.launch_count() - n0
    ms_sharded = timed(step, steps, after=drain)
    if world > 1 and pend['feat'] is not None:
        pend['feat'][0].wait()                     # (the gather issued for a frame that will not come)
    res = {'bev': bev, 'ms_per_frame': ms_sharded, 'frames_per_s': 1e3 / ms_sharded,
           'cuda_graph': graph is not None, 'cuda_graph_error': graph_error, 'launches_per_frame': launches,
           'h2d_bytes_per_frame_per_rank': chunk * 2 if world > 1 else 0,
           'collectives_per_frame': 'all-gather of the feature slices + all-gather of the BEV rows, both asynchronous '
                                    'and overlapped with the neighbouring frame' if world > 1 else 'none'}
    if world > 1:
        # the same frame on one rank alone: every rank times its own unsharded forward (graph replay, no
        # collective); efficiency = T(1) / (N * T(N)); the replicated share follows from T(N) = R + S / N
        del graph
        gfull, _ = capture(lambda: shard_forward(None)) if not args.no_graph else (None, None)

        def full_step():
            flush.zero_()
            if gfull is None:
                shard_forward(None)
            else:
                gfull.replay()
        for _ in range(3):
            full_step()
        ms_full = timed(full_step, max(3, steps // 2))
        res.update(ms_unsharded_one_gpu=ms_full, efficiency_vs_n1=ms_full / (world * ms_sharded),
                   speedup_vs_n1=ms_full / ms_sharded,
                   replicated_ms_estimate=max(0.0, (world * ms_sharded - ms_full) / (world - 1)),
                   replicated_note='work that does not shrink with N (value projections over the full maps, '
                                   'geometry, frame distribution, exit all-gather), from T(N) = R + S / N')
        del gfull
    if train_steps > 0:
        enc.train()
        params = [p for p in enc.parameters() if p.requires_grad]

        def train_step(sharded):
            flush.zero_()
            for p in params:
                p.grad = None
            feat = d['feat'].detach().requires_grad_(True)
            q = d['bev_query'].detach().requires_grad_(True)
            kw = kwargs_of(feat)
            if sharded and world > 1:
                out = sharded_encoder_forward(enc, q, feat, feat, **kw)
            else:
                out = enc(q, feat, feat, **kw)
            ((out.float() * d['grad_w'].float()).sum() * (1.0 / out.numel())).backward()
            if sharded and world > 1:
                allreduce_gradients(params)
        for _ in range(2):
            train_step(True)
        ms_train = timed(lambda: train_step(True), train_steps)
        res['train'] = {'ms_per_frame': ms_train, 'mode': 'eager forward + backward, gradients of the replicated '
                        'tensors all-reduced in the backward, parameter gradients in one flat all-reduce'}
        if world > 1:
            for _ in range(2):
                train_step(False)
            ms_train_full = timed(lambda: train_step(False), max(2, train_steps // 2))
            res['train'].update(ms_unsharded_one_gpu=ms_train_full,
                                efficiency_vs_n1=ms_train_full / (world * ms_train))
    return res


def run_rowshard(args):
    import torch.distributed as dist
    import apollo_vision_net_b200 as pkg
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    pkg.build()
    bev = args.bev if args.bev != 200 else 400
    res = measure_rowshard(args, rank, world, dev, args.steps, args.warmup, bev=bev,
                           train_steps=max(2, args.steps // 4))
    if rank == 0:
        print(json.dumps({
            'metric': 'BEVFormer encoder forward frames/s, BEV-query-row sharded', 'value': res['frames_per_s'],
            'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': res['ms_per_frame'], 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': f'BEVFormer-base encoder forward, {bev}x{bev} BEV, 6 cams, 4 levels, '
                                   f'{args.layers} layers, rows sharded over {world} rank(s), feature slices '
                                   'all-gathered once per frame, one all-gather of the rows at exit '
                                   '(BASELINE configs[4])',
                       'l2': 'flushed every step (192 MiB write inside the timed region)'},
            'rowshard': res, 'gpu_launches': res['launches_per_frame'] * args.steps}), flush=True)
    if world > 1:
        leave(dist)


# ------------------------------------------------ det + map stand-in (BASELINE configs[2]) ------
def measure_detmap(args, rank, world, dev, steps, warmup, decoder_only=False):
    """BASELINE configs[2] as a synthetic stand-in: BEVFormer-base encoder ->
    detection decoder (900 queries, 6 DetrTransformerDecoderLayers) + MapTRv2 decoder (350 vectors x 20
    points, one-to-many mask, 6 decoupled layers) on the same BEV -> L2 loss on both decoders' stacked
    outputs, forward + backward, one frame per rank, batch-level data parallelism (bucketed gradient
    all-reduce), the whole step replayed as a CUDA graph.  Backbone, heads' losses and assigners are out of
    scope; random-init weights, synthetic features.
    Reference: projects/configs/bevformer/bev_tiny_det_mapv2.py:5-63, bevformer/apis/mmdet_train.py:71-85.

    ``decoder_only``: BASELINE configs[3] -- the MapTRv2 decoder alone (deformable cross-attention of 7000
    queries on a given 200x200 BEV, 6 layers, one-to-many queries enabled), forward + backward."""
    import torch.distributed as dist
    import apollo_vision_net_b200 as pkg
    import apollo_vision_net_b200.synthetic as syn
    from apollo_vision_net_b200 import _lib
    from apollo_vision_net_b200.parallel import BucketedGradReducer
    dtype = torch.bfloat16
    levels = syn.LEVELS_BASE
    bev = 200
    HW = bev * bev
    NQ_DET, V, PN = 900, 350, 20
    torch.manual_seed(0)
    enc = pkg.build_transformer_layer_sequence(encoder_cfg(args.layers, len(levels), syn.PC_RANGE, args.dropout))
    det = pkg.build_transformer_layer_sequence(dict(
        type='DetectionTransformerDecoder', num_layers=6, return_intermediate=True,
        transformerlayers=dict(
            type='DetrTransformerDecoderLayer',
            attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=HEADS, dropout=args.dropout),
                       dict(type='CustomMSDeformableAttention', embed_dims=C, num_levels=1, dropout=args.dropout)],
            feedforward_channels=512, ffn_dropout=args.dropout,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    mapd = pkg.build_transformer_layer_sequence(dict(
        type='MapTRv2Decoder', num_layers=6, return_intermediate=True,
        transformerlayers=dict(
            type='MapTRv2DecoupledDetrTransformerDecoderLayer', num_vec=V, num_pts_per_vec=PN,
            attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=HEADS, dropout=args.dropout),
                       dict(type='MultiheadAttention', embed_dims=C, num_heads=HEADS, dropout=args.dropout),
                       dict(type='CustomMSDeformableAttention', embed_dims=C, num_levels=1, dropout=args.dropout)],
            feedforward_channels=512, ffn_dropout=args.dropout,
            operation_order=('self_attn', 'norm', 'self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    reg_det = torch.nn.ModuleList([torch.nn.Linear(C, 10) for _ in range(6)])
    reg_map = torch.nn.ModuleList([torch.nn.Linear(C, 2) for _ in range(6)])
    model = torch.nn.ModuleList([mapd, reg_map] if decoder_only else [enc, det, mapd, reg_det, reg_map])
    randomize(model, 0)
    model.to(dev).to(dtype).train()
    params = [p for p in model.parameters() if p.requires_grad]
    reducer = BucketedGradReducer(params, bucket_bytes=args.bucket_mb << 20) if world > 1 else None
    l2i, img_shape = syn.camera_rig(1.0, bs=1)
    l2i_dev = torch.as_tensor(l2i).to(dev)
    host = host_inputs(bev, bev, levels, seed=1 + rank, dtype=dtype, pin=False)
    d = {k: v.to(dev) for k, v in host.items()}
    g = torch.Generator().manual_seed(7 + rank)
    mk = lambda *sh: torch.randn(*sh, generator=g).to(dev, dtype)      # noqa: E731
    q_det, p_det, q_map, p_map = mk(NQ_DET, 1, C), mk(NQ_DET, 1, C), mk(V * PN, 1, C), mk(V * PN, 1, C)
    ref_det = torch.rand(1, NQ_DET, 3, generator=g).to(dev, dtype)
    ref_map = torch.rand(1, V * PN, 2, generator=g).to(dev, dtype)
    mask = torch.zeros(V, V, dtype=torch.bool, device=dev)
    mask[50:, :50] = True
    mask[:50, 50:] = True                              # one-to-one (50) / one-to-many (300) vectors
    bshape, bstart = torch.tensor([[bev, bev]], device=dev), torch.tensor([0], device=dev)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step():
        for p in params:
            p.grad = None
        if reducer is not None:
            reducer.reset()
        if decoder_only:
            value = d['prev_bev'].detach().requires_grad_(True)                           # a BEV (HW, 1, C)
        else:
            feat = d['feat'].detach().requires_grad_(True)
            bev_out = enc(d['bev_query'], feat, feat, bev_h=bev, bev_w=bev, bev_pos=d['bev_pos'],
                          spatial_shapes=d['shapes'], level_start_index=d['starts'], prev_bev=d['prev_bev'],
                          shift=d['shift'], lidar2img=l2i_dev, img_shape=img_shape)      # (1, HW, C)
            value = bev_out.permute(1, 0, 2)                                              # (HW, 1, C)
        map_out, _ = mapd(q_map, key=None, value=value, query_pos=p_map, reference_points=ref_map,
                          reg_branches=reg_map, spatial_shapes=bshape, level_start_index=bstart,
                          self_attn_mask=mask, num_vec=V, num_pts_per_vec=PN)
        loss = map_out.float().square().mean()
        if not decoder_only:
            det_out, _ = det(q_det, key=None, value=value, query_pos=p_det, reference_points=ref_det,
                             reg_branches=reg_det, spatial_shapes=bshape, level_start_index=bstart)
            loss = loss + det_out.float().square().mean()
        loss.backward()
        if reducer is not None:
            reducer.finish()
        return loss

    def capture():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in params:
            p.grad = None
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, capture_error_mode='thread_local' if world > 1 else 'global'):
            out = step()
        torch.cuda.synchronize()
        return gr, out

    graph, graph_error = None, None
    if not args.no_graph:
        try:
            graph, _ = capture()
        except Exception as exc:                       # pragma: no cover - reported in the JSON line
            graph, graph_error = None, f'{type(exc).__name__}: {exc}'[:200]
            torch.cuda.synchronize()

    def run():
        flush.zero_()
        if graph is None:
            step()
        else:
            graph.replay()

    n0 = _lib.launch_count()
    step()
    launches = _lib.launch_count() - n0
    for _ in range(max(warmup, 3)):
        run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        run()
    e.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / steps
    if reducer is not None:
        reducer.remove()
    what = ('MapTRv2 decoder alone (350 x 20 queries x 6 layers, one-to-many mask, deformable cross-attention on a '
            '200x200 BEV) -> L2 loss, fwd + bwd (BASELINE configs[3])') if decoder_only else (
        f'encoder ({args.layers} layers, 200x200) + detection decoder (900 queries x 6) + MapTRv2 '
        'decoder (350 x 20 queries x 6, one-to-many) -> L2 loss, fwd + bwd, one frame per rank')
    kernel_us = None
    if decoder_only:
        # our kernels inside this step, timed live with CUDA events around each C-ABI launch (eager steps; a replayed
        # graph cannot carry events): mha_fwd / mha_bwd = the self-attention core (mha_bwd = its two passes),
        # tsa_fwd / tsa_bwd = the deformable cross-attention (queue of one), ln_* / colsum = the row kernels
        timer = _lib.KernelTimer()
        _lib.set_timer(timer)
        try:
            for _ in range(3):
                step()
        finally:
            _lib.set_timer(None)
        kernel_us = {name: {'launches_per_step': st['launches'] / 3, 'mean_us': round(st['mean_us'], 2)}
                     for name, st in timer.summary().items()}
    torch_ms = None
    if decoder_only and world == 1:
        # the same step with the self-attentions on torch.nn.MultiheadAttention + the reference's permute copies
        # (what this repository ran before csrc/mha.cu): the gain of the attention core, measured side by side
        for layer in mapd.layers:
            for att in layer.attentions[:2]:
                att.use_fused_core = False
        try:
            graph = None
            if not args.no_graph:
                graph, _ = capture()
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            s.record()
            for _ in range(steps):
                run()
            e.record()
            torch.cuda.synchronize()
            torch_ms = s.elapsed_time(e) / steps
        except Exception:                              # pragma: no cover
            torch.cuda.synchronize()
    return {'workload': what, 'self_attention': 'csrc/mha.cu (mma.sync tensor-core core, tokens attended in place, '
                                                'LayerNorms folded into the blocks)',
            'ms_per_step_torch_self_attention': torch_ms, 'kernels': kernel_us,
            'ms_per_step': ms, 'frames_per_s': world * 1e3 / ms, 'cuda_graph': graph is not None,
            'cuda_graph_error': graph_error, 'our_launches_per_step': launches, 'dropout': args.dropout,
            'parallelism': f'dp{world}' if world > 1 else 'single'}


def run_detmap(args):
    import torch.distributed as dist
    import apollo_vision_net_b200 as pkg
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    pkg.build()
    res = measure_detmap(args, rank, world, dev, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps({
            'metric': 'det+map stand-in frames/s (encoder + detection decoder + MapTRv2 decoder, fwd+bwd)',
            'value': res['frames_per_s'], 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': res['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': res['workload'] + ' (BASELINE configs[2], synthetic stand-in)'},
            'detmap': res, 'gpu_launches': res['our_launches_per_step'] * args.steps}), flush=True)
    if world > 1:
        leave(dist)


# ---------------------------------------------------------------------- reference arm ------
def run_reference_sample(steps, warmup):
    """The reference's CPU path (oracle port) on a bounded sample of the workload: ONE of the
    six encoder layers of the base configuration (200x200 BEV, 4 levels, 6 cameras, 8x4 points),
    fp32, forward + backward, all host threads (BENCH_REF_BEV shrinks the BEV for quick checks)."""
    from oracle.modules_oracle import OracleBEVFormerEncoder
    import apollo_vision_net_b200.synthetic as syn
    torch.set_num_threads(os.cpu_count() or 1)
    bev = int(os.environ.get('BENCH_REF_BEV', '200'))
    levels = syn.LEVELS_BASE
    shapes_l, starts_l, Nk = syn.level_tables(levels)
    enc = OracleBEVFormerEncoder(num_layers=1, pc_range=syn.PC_RANGE, num_points_in_pillar=PILLAR,
                                 embed_dims=C, feedforward_channels=FFN_DIM, num_levels=len(levels),
                                 dropout=0.0)
    randomize(enc, 0)
    enc.train()
    host = host_inputs(bev, bev, levels, seed=1, dtype=torch.float32, pin=False)
    l2i, img_shape = syn.camera_rig(1.0, bs=1)
    from oracle import geometry_oracle as G
    r3 = G.reference_points_3d(bev, bev, 8.0, PILLAR, bs=1)
    _, mask = G.point_sampling(r3, syn.PC_RANGE, l2i, img_shape[0], img_shape[1])
    pairs = int(sum(len(x) for x in G.camera_hit_lists(mask)[0]))
    max_len = G.camera_hit_lists(mask)[1]
    # the reference computes 6 * max_len padded rows; the metric counts the real (cam, query) pairs
    points = pairs * HEADS * len(levels) * SCA_POINTS + 2 * bev * bev * HEADS * TSA_POINTS

    def step():
        enc.zero_grad()
        feat = host['feat'].clone().requires_grad_(True)
        out = enc(host['bev_query'], feat, feat, bev_h=bev, bev_w=bev, bev_pos=host['bev_pos'],
                  spatial_shapes=host['shapes'], level_start_index=host['starts'],
                  prev_bev=host['prev_bev'], shift=host['shift'], lidar2img=l2i,
                  img_h=img_shape[0], img_w=img_shape[1])
        loss = (out * host['grad_w']).sum() * (1.0 / out.numel())
        loss.backward()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    cpu = ''
    try:
        with open('/proc/cpuinfo') as f:
            cpu = next(l.split(':', 1)[1].strip() for l in f if l.startswith('model name'))
    except Exception:
        pass
    return dict(value=points * steps / dt, unit=UNIT, cores=torch.get_num_threads(), kind='port',
                sample=f'1 of 6 encoder layers, {bev}x{bev} BEV, full 4-level '
                       f'6-camera value maps, fp32 fwd+bwd, {steps} step(s), padded max_len={max_len}; '
                       f'oracle port of the reference CPU branch (multi_scale_deformable_attn_pytorch)',
                seconds_per_step=dt / steps, points_per_step=points, cpu_model=cpu)


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    base = run_reference_sample(steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': min(args.warmup, 1),
        'ms_per_step': base['seconds_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'BEVFormer-base encoder fwd+bwd (BASELINE configs[1]) -- reference CPU '
                               'path on a bounded sample: ' + base['sample']},
        'cpu_baseline': base,
        'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--bev', type=int, default=200)
    ap.add_argument('--layers', type=int, default=6)
    ap.add_argument('--workload', default='encoder', choices=['encoder', 'rowshard', 'detmap'],
                    help="'encoder' (default, the contract line) or 'rowshard' (400x400 forward, strong scaling)")
    ap.add_argument('--shard-sim', type=int, default=0,
                    help='rowshard on ONE GPU: run rank 0 of this many row shards, no collective '
                         '(profiling aid: what one rank of an N-way run executes)')
    ap.add_argument('--dropout', type=float, default=0.1,
                    help="dropout of TSA / SCA / FFN (reference training value 0.1; 0 = the parity-checked "
                         "deterministic configuration)")
    ap.add_argument('--bucket-mb', type=int, default=2, help='gradient bucket size of the data-parallel arm (MiB)')
    ap.add_argument('--no-extras', action='store_true',
                    help='N > 1: skip the row-shard / det+map measurements attached to the line')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='run the step eagerly instead of replaying a CUDA graph')
    ap.add_argument('--watchdog-s', type=int, default=1500,
                    help='abort the process (with a traceback of every thread on stderr) when the run takes longer')
    args = ap.parse_args()
    import faulthandler
    if args.watchdog_s > 0:          # a hung collective must end the run with a diagnosis, not sit on the GPUs
        faulthandler.dump_traceback_later(args.watchdog_s, exit=True)
    import torch
    torch.manual_seed(0)             # initial weights come from the global RNG: same model on every rank
    if args.impl == 'reference':
        args.steps = min(args.steps, 3)
        run_reference(args)
    elif args.workload == 'rowshard':
        run_rowshard(args)
    elif args.workload == 'detmap':
        run_detmap(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == '__main__':
    main()
