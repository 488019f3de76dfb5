"""Development aid: error of grad_value (tensor-core coarse pass on / off) against an fp32 run of the
kernel, per pyramid level (max error relative to the level's max, and relative 2-norm)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apollo_vision_net_b200.fused_ops as fo  # noqa: E402
import apollo_vision_net_b200.synthetic as syn  # noqa: E402
from tests.test_coarse_pass_gpu import _case, _grads  # noqa: E402


def errs(a, b):
    d = (a - b).double()
    return float(d.abs().max() / b.abs().max().clamp_min(1e-30)), float(d.norm() / b.double().norm().clamp_min(1e-30))


for name, kw in [('base200 spread4', dict(H=200, W=200, spread=4.0)), ('40 spread0', dict(H=40, W=40, spread=0.0)),
                 ('40 spread4', dict(H=40, W=40, spread=4.0)), ('100 spread1', dict(H=100, W=100, spread=1.0))]:
    c = _case(kw['H'], kw['W'], 1, syn.LEVELS_BASE, 8, torch.bfloat16, seed=13, spread=kw['spread'])
    _, truth, _, _ = _grads(c, value=c['value'].float(), go=c['go'].float())
    tb = truth.bfloat16().float()
    for on in (True, False):
        prev = fo.set_coarse_tensor_core_pass(on)
        _, gv, _, _ = _grads(c)
        fo.set_coarse_tensor_core_pass(prev)
        per = []
        for l in range(c['L']):
            s0 = c['starts_l'][l]
            s1 = c['starts_l'][l + 1] if l + 1 < c['L'] else c['Nk']
            per.append('L%d max %.2e l2 %.2e' % ((l,) + errs(gv[:, s0:s1], truth[:, s0:s1])))
        print(name, 'coarse' if on else 'reduce', '| '.join(per), flush=True)
    print(name, 'bf16 rounding of the truth: max %.2e l2 %.2e' % errs(tb, truth), flush=True)
