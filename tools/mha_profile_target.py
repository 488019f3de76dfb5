"""One forward + backward of the decoder self-attention core at a BASELINE configs[3] shape (ncu target).
    ncu --set full --clock-control none --import-source on -k regex:mha_ -c 3 python tools/mha_profile_target.py [inter|intra|det] [p]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apollo_vision_net_b200.mha as m  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else 'inter'
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
dev = torch.device('cuda:0')
C, H = 256, 8
lay = {'inter': m.inter_vector_layout(350, 20, 1), 'intra': m.intra_vector_layout(350, 20, 1),
       'det': m.sequence_first_layout(900, 1)}[case]
mask = None
if case == 'inter':
    mask = torch.zeros(350, 350, dtype=torch.bool, device=dev)
    mask[50:, :50] = True
    mask[:50, 50:] = True
rows = lay.G * lay.S
torch.manual_seed(0)
qk = torch.randn(rows, 2 * C, device=dev, dtype=torch.bfloat16, requires_grad=True)
v = torch.randn(rows, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
go = torch.randn(rows, C, device=dev, dtype=torch.bfloat16)
o = m.self_attention_core(qk, v, lay, H, mask, p)
o.backward(go)
torch.cuda.synchronize()
print('ok', float(o.float().abs().mean()))
