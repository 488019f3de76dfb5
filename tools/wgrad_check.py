"""Development check of linear_wgrad against torch (correctness + timing at the layer's shapes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apollo_vision_net_b200 import _lib
from apollo_vision_net_b200.multi_scale_deformable_attn_function import _DTYPE_CODE

dev = torch.device('cuda:0')
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(N, O, I, dtype, time_it=True):
    g = torch.Generator(device='cpu').manual_seed(N + O + I)
    dy = torch.randn(N, O, generator=g).to(dtype).to(dev)
    x = torch.randn(N, I, generator=g).to(dtype).to(dev)
    dW = torch.full((O, I), float('nan'), dtype=dtype, device=dev)
    db = torch.full((O,), float('nan'), dtype=dtype, device=dev)
    ws = torch.zeros(int(_lib.lib().linear_wgrad_workspace_floats(O, I)), device=dev)
    def ours():
        _lib.call('linear_wgrad', dy.data_ptr(), x.data_ptr(), dW.data_ptr(), db.data_ptr(), ws.data_ptr(), N, O, I,
                  _DTYPE_CODE[dtype], st)
    def theirs():
        return dy.t() @ x, dy.sum(0)
    ours(); ours()
    torch.cuda.synchronize()
    rW = dy.double().t() @ x.double()
    rb = dy.double().sum(0)
    eW = float((dW.double() - rW).abs().max() / rW.abs().max())
    eb = float((db.double() - rb).abs().max() / rb.abs().max().clamp_min(1e-30))
    tW, tb = theirs()
    eWt = float((tW.double() - rW).abs().max() / rW.abs().max())
    assert float(ws.abs().max()) == 0.0, 'workspace not re-zeroed'
    res = dict(N=N, O=O, I=I, dtype=str(dtype), err_dW=eW, err_db=eb, err_dW_torch=eWt)
    if time_it:
        for name, fn in (('ours_us', ours), ('torch_us', theirs)):
            ts = []
            for _ in range(10):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            res[name] = round(sorted(ts)[len(ts) // 2], 1)
    print(res, flush=True)


if __name__ == '__main__':
    bf = torch.bfloat16
    run(1000, 256, 256, bf, False)
    run(37, 192, 512, bf, False)
    run(4099, 768, 256, torch.float16, False)
    run(1, 8, 8, bf, False)
    for N, O, I in [(40000, 256, 256), (80000, 256, 256), (184950, 256, 256), (40000, 768, 256), (40000, 192, 512),
                    (40000, 512, 256), (40000, 256, 512)]:
        run(N, O, I, bf)
