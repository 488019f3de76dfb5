"""Multi-GPU partitioning of the hot path: BEV-query row sharding.

BEV queries are independent units of SCA, TSA, LayerNorm and the FFN, so the encoder shards by
contiguous blocks of BEV rows (spatially compact stripes: good L2 locality for the gathers).  The
value tensors -- the six cameras' feature maps and the ``[prev_bev, bev]`` pair -- are
replicated; with history there is no per-layer exchange and the only collective on the data
path is ONE all-gather of the output rows at encoder exit (NCCL over NVLink / NVSwitch).
The reference has no counterpart (its only strategy is DDP, apis/mmdet_train.py:71-85); batch-level
data parallelism for training is :class:`BucketedGradReducer` (DDP's bucketed, backward-overlapped
all-reduce of the parameter gradients, usable inside a captured CUDA graph).

Training with sharded rows (SURVEY.md section 8e "Training"): :func:`sharded_encoder_forward` is
differentiable.  The exit all-gather hands every rank its own rows of the output gradient back (the
consumer of the BEV is replicated), the replicated inputs (image features, history) receive partial
gradients from every rank's rows, which :func:`replicated` sums over the group in the backward, and the
parameter gradients -- partial for the same reason -- are summed by :func:`allreduce_gradients`.

Everything here is host logic over ``torch.distributed``; it works with the ``gloo`` backend on
CPU tensors (tests) and ``nccl`` on CUDA tensors alike.
"""
import torch
import torch.distributed as dist


def bev_row_range(bev_h, rank, world):
    """Rows [y0, y1) of the BEV grid owned by ``rank``: contiguous, sizes differ by at most 1."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of {world}')
    base, extra = divmod(bev_h, world)
    y0 = rank * base + min(rank, extra)
    return y0, y0 + base + (1 if rank < extra else 0)


def bev_query_range(bev_h, bev_w, rank, world):
    y0, y1 = bev_row_range(bev_h, rank, world)
    return y0 * bev_w, y1 * bev_w


def all_gather_bev_rows(local, bev_h, bev_w, group=None):
    """local (bs, rows_r * bev_w, C) on every rank -> (bs, bev_h * bev_w, C) on every rank.

    One collective: ``all_gather_into_tensor`` when the rows divide evenly, otherwise a padded
    gather (ranks own at most one extra row)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local
    bs, n_local, C = local.shape
    counts = [(bev_row_range(bev_h, r, world)[1] - bev_row_range(bev_h, r, world)[0]) * bev_w
              for r in range(world)]
    assert counts[rank] == n_local, (counts, rank, n_local)
    n_max = max(counts)
    # gather along a leading rank dimension, then stitch the row blocks together
    send = local if n_local == n_max else torch.cat(
        [local, local.new_zeros(bs, n_max - n_local, C)], 1)
    send = send.contiguous()
    recv = send.new_empty((world,) + tuple(send.shape))
    if hasattr(dist, 'all_gather_into_tensor') and dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        parts = [recv[r] for r in range(world)]
        dist.all_gather(parts, send, group=group)
    if bs == 1 and all(c == n_max for c in counts):
        return recv.view(1, world * n_max, C)               # already in row order: no stitching copy
    return torch.cat([recv[r, :, :counts[r]] for r in range(world)], 1)


class _GatherRows(torch.autograd.Function):
    """all_gather_bev_rows with a backward: every rank consumes the full BEV in a replicated computation,
    so its gradient w.r.t. the full BEV is complete and the gradient of the local rows is that slice
    (``reduce=False``); when the consumers differ per rank (``reduce=True``) the slices are summed over
    the group first."""

    @staticmethod
    def forward(ctx, local, bev_h, bev_w, group, reduce):
        ctx.args = (bev_h, bev_w, group, reduce, local.shape[1])
        return all_gather_bev_rows(local, bev_h, bev_w, group=group)

    @staticmethod
    def backward(ctx, g_full):
        bev_h, bev_w, group, reduce, n_local = ctx.args
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if reduce:
            g_full = g_full.contiguous().clone()
            dist.all_reduce(g_full, group=group)
        q0, q1 = bev_query_range(bev_h, bev_w, rank, world)
        assert q1 - q0 == n_local
        return g_full[:, q0:q1].contiguous(), None, None, None, None


class _Replicated(torch.autograd.Function):
    """Identity on a tensor every rank of the row group holds in full; the backward SUMS the gradient over
    the group (each rank's rows contribute a partial gradient to the image features / the history)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, group=ctx.group)
        return g, None


def replicated(x, group=None):
    """Mark a replicated input of the row-sharded encoder (image features, ``prev_bev``): its gradient is
    all-reduced (sum) over the row group in the backward.  No-op without a process group / gradient."""
    if x is None or not dist.is_initialized() or dist.get_world_size(group) == 1 or not x.requires_grad:
        return x
    return _Replicated.apply(x, group)


def allreduce_gradients(params, group=None, average=False):
    """Sum (or average) the ``.grad`` of ``params`` over the group in ONE flat collective.  Row sharding:
    every rank holds the gradient contribution of its rows -- sum.  Data parallelism without overlap:
    average."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def sharded_encoder_forward(encoder, bev_query, key, value, *args, group=None, reduce_output_grad=False,
                            **kwargs):
    """Run ``encoder`` on this rank's BEV rows and all-gather the full BEV; differentiable (see the module
    docstring).  ``encoder`` is any callable with the ``BEVFormerEncoder.forward`` signature that honours
    ``row_shard``.  After ``backward()`` call :func:`allreduce_gradients` on the encoder's parameters."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return encoder(bev_query, key, value, *args, **kwargs)
    # every tensor the ranks hold in full receives, in the backward, the contribution of this rank's rows
    # only (the BEV queries also feed the temporal self-attention's value pair): summed over the group
    same = value is key
    key = replicated(key, group)
    value = key if same else replicated(value, group)
    bev_query = replicated(bev_query, group)
    for name in ('prev_bev', 'bev_pos'):
        if kwargs.get(name) is not None:
            kwargs[name] = replicated(kwargs[name], group)
    local = encoder(bev_query, key, value, *args, row_shard=(rank, world), **kwargs)
    if torch.is_grad_enabled() and local.requires_grad:
        return _GatherRows.apply(local, kwargs['bev_h'], kwargs['bev_w'], group, reduce_output_grad)
    return all_gather_bev_rows(local, kwargs['bev_h'], kwargs['bev_w'], group=group)


class BucketedGradReducer:
    """Batch-level data parallelism the way DDP does it (the reference's only strategy,
    bevformer/apis/mmdet_train.py:71-85): parameter gradients are averaged over the ranks in BUCKETS,
    each launched as an asynchronous all-reduce from a post-accumulate-grad hook as soon as the last
    gradient of the bucket exists, so the collectives overlap the rest of the backward; ``finish()`` waits
    and leaves the averaged gradients in ``p.grad`` (views of the bucket buffers -- no copy back).

    Buckets follow the REVERSE registration order of the parameters (the order the backward produces
    them); ``bucket_bytes`` is sized for launch latency and overlap, not for link count (NVSwitch).
    Works eagerly and inside ``torch.cuda.graph`` capture (the hooks run at capture time, the NCCL
    launches become nodes of the graph)."""

    def __init__(self, params, group=None, bucket_bytes=4 << 20, average=True):
        self.params = [p for p in params if p.requires_grad]
        self.group, self.average = group, average
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._where = {id(p): (b, i) for b, ps in enumerate(self.buckets) for i, p in enumerate(ps)}
        self._flat = [None] * len(self.buckets)
        self._pending = [0] * len(self.buckets)
        self._works = []
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.enabled = True
        self.reset()

    def reset(self):
        """Before every backward (gradients are expected to be freshly produced, ``p.grad = None``)."""
        self._pending = [len(b) for b in self.buckets]
        self._works = []

    def _hook(self, p):
        if not self.enabled:
            return
        b, _ = self._where[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b):
        ps = self.buckets[b]
        # (parameters without a gradient this step -- unused in the graph -- contribute zeros, like DDP's
        # find_unused_parameters; every rank runs the same model, so the collectives stay matched)
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
        if self.average and self.world > 1:
            flat.mul_(1.0 / self.world)
        self._flat[b] = flat
        work = dist.all_reduce(flat, group=self.group, async_op=True) if self.world > 1 else None
        self._works.append((b, work))

    def finish(self):
        """Wait for the collectives and point every ``p.grad`` at its slice of the reduced bucket."""
        for b, ps in enumerate(self.buckets):          # buckets that wait for a gradient that never came
            if self._pending[b] > 0 and any(p.grad is not None for p in ps):
                self._pending[b] = 0
                self._launch(b)
        for b, work in self._works:
            if work is not None:
                work.wait()
            off = 0
            for p in self.buckets[b]:
                n = p.numel()
                if p.grad is not None:
                    p.grad = self._flat[b][off:off + n].view_as(p)
                off += n
        self._works = []

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
