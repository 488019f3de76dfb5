"""One MapTRv2 decoder training step (BASELINE configs[3] shape) between cudaProfilerStart / Stop, as an ncu target:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/decoder_launches.csv python tools/decoder_profile_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apollo_vision_net_b200 as pkg  # noqa: E402

dev = torch.device('cuda:0')
dtype = torch.bfloat16
torch.manual_seed(0)
C, V, Pn, bs, H = 256, 350, 20, 1, 200
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
dec.to(dev).to(dtype).train()
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
    for prm in params:             # as a training loop does (zero_grad(set_to_none=True)): no accumulate-into-grad adds
        prm.grad = None


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('ok')
