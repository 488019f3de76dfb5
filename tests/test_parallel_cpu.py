"""world_size-2 (and 3) gloo tests of the BEV row-sharding host logic on CPU: the partition, the
sliced geometry and the single all-gather at encoder exit reproduce the unsharded result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


class _RowWiseEncoder(torch.nn.Module):
    """Stand-in with the encoder's calling convention whose layers are row-independent given
    replicated value tensors -- exactly the property the sharding relies on."""

    def __init__(self, C):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.w = torch.nn.Parameter(torch.randn(C, C, generator=g) * 0.1)

    def forward(self, bev_query, key, value, bev_h=None, bev_w=None, bev_pos=None, prev_bev=None,
                row_shard=None, **kw):
        from apollo_vision_net_b200.parallel import bev_query_range
        q = bev_query.permute(1, 0, 2)
        pos = bev_pos.permute(1, 0, 2)
        prev = prev_bev.permute(1, 0, 2)
        if row_shard is not None:
            q0, q1 = bev_query_range(bev_h, bev_w, *row_shard)
            q, pos, paired = q[:, q0:q1], pos[:, q0:q1], prev[:, q0:q1]
        else:
            paired = prev
        ctx = value.mean(dim=(0, 1))                       # replicated "value" summary (bs, C)
        out = q
        for _ in range(3):
            out = torch.tanh((out + pos + paired) @ self.w) + ctx[:, None, :]
        return out


def _worker(rank, world, port, bev_h, bev_w, ret, bs=2):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from apollo_vision_net_b200.parallel import sharded_encoder_forward
        C = 16
        g = torch.Generator().manual_seed(1)
        HW = bev_h * bev_w
        bevq, pos, prev = (torch.randn(HW, bs, C, generator=g) for _ in range(3))
        feat = torch.randn(6, 40, bs, C, generator=g)
        enc = _RowWiseEncoder(C)
        full = enc(bevq, feat, feat, bev_h=bev_h, bev_w=bev_w, bev_pos=pos, prev_bev=prev)
        out = sharded_encoder_forward(enc, bevq, feat, feat, bev_h=bev_h, bev_w=bev_w, bev_pos=pos,
                                      prev_bev=prev)
        ok = out.shape == full.shape and torch.equal(out, full)
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,bev_h,bev_w,bs', [(2, 10, 6, 2), (2, 7, 5, 2), (3, 8, 4, 2), (2, 10, 6, 1)])
def test_row_sharded_forward_equals_full(world, bev_h, bev_w, bs):
    """Even and uneven row splits; bs = 1 with an even split takes the all-gather's no-copy path
    (the gathered buffer is already in row order)."""
    ctx = mp.get_context('spawn')
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, bev_h, bev_w, ret, bs)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() == 1


def test_row_ranges_partition_the_grid():
    from apollo_vision_net_b200.parallel import bev_row_range
    for h in (1, 7, 200, 400):
        for world in (1, 2, 3, 4, 8):
            spans = [bev_row_range(h, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == h
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _train_worker(rank, world, port, bev_h, bev_w, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from apollo_vision_net_b200.parallel import allreduce_gradients, sharded_encoder_forward
        C, bs = 16, 2
        g = torch.Generator().manual_seed(1)
        HW = bev_h * bev_w
        bevq, pos, prev = (torch.randn(HW, bs, C, generator=g) for _ in range(3))
        feat = torch.randn(6, 40, bs, C, generator=g)
        gw = torch.randn(bs, HW, C, generator=g)

        def run(sharded):
            enc = _RowWiseEncoder(C)
            q = bevq.clone().requires_grad_(True)
            f = feat.clone().requires_grad_(True)
            pv = prev.clone().requires_grad_(True)
            kw = dict(bev_h=bev_h, bev_w=bev_w, bev_pos=pos, prev_bev=pv)
            out = sharded_encoder_forward(enc, q, f, f, **kw) if sharded else enc(q, f, f, **kw)
            (out * gw).sum().backward()                   # replicated consumer: the same loss on every rank
            if sharded:
                allreduce_gradients(enc.parameters())
            return out.detach(), enc.w.grad, f.grad, pv.grad, q.grad

        full = run(False)
        shard = run(True)
        from apollo_vision_net_b200.parallel import bev_query_range
        q0, q1 = bev_query_range(bev_h, bev_w, rank, world)
        def close(a, b):                                  # (sums in a different order: relative to the scale)
            return float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())

        ok = torch.equal(shard[0], full[0])
        ok = ok and close(shard[1], full[1])              # parameter gradient (summed over the row group)
        ok = ok and close(shard[2], full[2])              # replicated features
        # the history enters row-wise in this stand-in: every rank holds its rows' gradient after the sum
        ok = ok and close(shard[3], full[3])
        # queries (a replicated parameter of the real model): every rank ends with the full gradient
        ok = ok and close(shard[4], full[4]) and q1 > q0
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,bev_h,bev_w', [(2, 10, 6), (3, 8, 4), (2, 7, 5)])
def test_row_sharded_training_gradients_equal_unsharded(world, bev_h, bev_w):
    """Backward through the sharded encoder: output rows gathered, replicated-input gradients and parameter
    gradients summed over the row group -- equal to the unsharded encoder's."""
    ctx = mp.get_context('spawn')
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, bev_h, bev_w, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() == 1


def _ddp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from apollo_vision_net_b200.parallel import BucketedGradReducer
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(8, 32), torch.nn.Tanh(), torch.nn.Linear(32, 32), torch.nn.Tanh(),
                                  torch.nn.Linear(32, 4))
        red = BucketedGradReducer(net.parameters(), bucket_bytes=2048)
        assert len(red.buckets) >= 2
        xs = [torch.randn(5, 8, generator=torch.Generator().manual_seed(10 + r)) for r in range(world)]
        ok = True
        for _ in range(2):                                 # two steps: the reducer is reusable
            for p in net.parameters():
                p.grad = None
            red.reset()
            net(xs[rank]).square().sum().backward()
            red.finish()
            mine = [p.grad.clone() for p in net.parameters()]
            # reference: the average over the ranks' batches, computed locally
            want = [torch.zeros_like(p) for p in net.parameters()]
            for r in range(world):
                for p in net.parameters():
                    p.grad = None
                net(xs[r]).square().sum().backward()
                for w, p in zip(want, net.parameters()):
                    w += p.grad / world
            ok = ok and all(float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) + 1e-12 for a, b in zip(mine, want))
        red.remove()
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


def test_bucketed_grad_reducer_averages_like_ddp():
    ctx = mp.get_context('spawn')
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() == 1
