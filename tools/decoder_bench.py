"""MapTRv2 decoder (BASELINE configs[3]: 350 vectors x 20 points = 7000 queries, 6 decoupled layers,
one-to-many mask) forward + backward on one GPU, CUDA events, with the BEV value projections of the six
cross-attentions hoisted into one batched GEMM (default) and per layer (the reference's order).
    python tools/decoder_bench.py [--bev 50] [--dtype bf16] [--iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apollo_vision_net_b200 as pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--bev', type=int, default=50)
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--train', action='store_true', help='training mode (dropout 0.1 everywhere)')
    ap.add_argument('--torch-self-attn', action='store_true',
                    help='self-attentions on torch.nn.MultiheadAttention with the permute copies (the round-1 path)')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    dtype = {'bf16': torch.bfloat16, 'fp32': torch.float32}[args.dtype]
    torch.manual_seed(0)
    C, V, Pn, bs, H = 256, 350, 20, 1, args.bev
    dec = pkg.build_transformer_layer_sequence(dict(
        type='MapTRv2Decoder', num_layers=6, return_intermediate=True,
        transformerlayers=dict(
            type='MapTRv2DecoupledDetrTransformerDecoderLayer', num_vec=V, num_pts_per_vec=Pn,
            attn_cfgs=[dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                       dict(type='MultiheadAttention', embed_dims=C, num_heads=8, dropout=0.1),
                       dict(type='CustomMSDeformableAttention', embed_dims=C, num_levels=1)],
            feedforward_channels=512, ffn_dropout=0.1,
            operation_order=('self_attn', 'norm', 'self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))
    g = torch.Generator().manual_seed(1)
    for n, p in dec.named_parameters():
        if n.endswith('sampling_offsets.weight') or n.endswith('attention_weights.weight'):
            p.data = torch.randn(p.shape, generator=g) * 0.02
    dec.to(dev).to(dtype).train(args.train)
    if args.torch_self_attn:
        for layer in dec.layers:
            for a in layer.attentions[:2]:
                a.use_fused_core = False
    reg = torch.nn.ModuleList([torch.nn.Linear(C, 2) for _ in range(6)]).to(dev).to(dtype)
    query = torch.randn(V * Pn, bs, C, device=dev, dtype=dtype, requires_grad=True)
    qpos = torch.randn(V * Pn, bs, C, device=dev, dtype=dtype)
    bev = torch.randn(H * H, bs, C, device=dev, dtype=dtype, requires_grad=True)
    refp = torch.rand(bs, V * Pn, 2, device=dev, dtype=dtype)
    mask = torch.zeros(V, V, dtype=torch.bool, device=dev)
    mask[50:, :50] = True
    mask[:50, 50:] = True
    shapes = torch.tensor([[H, H]], device=dev)
    starts = torch.tensor([0], device=dev)
    go = torch.randn(6, V * Pn, bs, C, device=dev, dtype=dtype)
    params = list(dec.parameters()) + list(reg.parameters())

    def step():
        inter, _ = dec(query, key=None, value=bev, query_pos=qpos, reference_points=refp, reg_branches=reg,
                       spatial_shapes=shapes, level_start_index=starts, self_attn_mask=mask, num_vec=V,
                       num_pts_per_vec=Pn)
        inter.backward(go)
        query.grad = None
        bev.grad = None
        for prm in params:         # as a training loop does (zero_grad(set_to_none=True))
            prm.grad = None

    res = {'bev': H, 'dtype': args.dtype, 'train': args.train, 'torch_self_attn': args.torch_self_attn, 'queries': V * Pn, 'xattn_samples_per_layer': V * Pn * 8 * 4}
    for hoist in (True, False):
        dec.hoist_value_proj = hoist
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            step()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        res['hoisted_ms' if hoist else 'per_layer_ms'] = round(ts[len(ts) // 2], 3)
    # the same step replayed as a CUDA graph: GPU time without the host's launch overhead
    dec.hoist_value_proj = True
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        dec.zero_grad(set_to_none=True)
        reg.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            graph.replay()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        res['graph_replay_ms'] = round(ts[len(ts) // 2], 3)
    except Exception as exc:
        res['graph_error'] = f'{type(exc).__name__}: {exc}'[:160]
        torch.cuda.synchronize()
    # one cross-attention module alone, forward + backward, eager
    att = dec.layers[0].attentions[2]
    ref_in = refp.unsqueeze(2)
    q1 = query.detach().clone().requires_grad_(True)
    go1 = go[0]

    def xstep():
        o = att(q1, None, bev, query_pos=qpos, reference_points=ref_in, spatial_shapes=shapes,
                level_start_index=starts)
        o.backward(go1)
        q1.grad = None
        bev.grad = None
    for _ in range(5):
        xstep()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        xstep()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    res['one_cross_attention_ms'] = round(ts[len(ts) // 2], 3)
    print(json.dumps(res))


if __name__ == '__main__':
    main()
