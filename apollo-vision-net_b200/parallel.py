"""Multi-GPU partitioning of the hot path: BEV-query row sharding.

BEV queries are independent units of SCA, TSA, LayerNorm and the FFN, so the encoder shards by
contiguous blocks of BEV rows (spatially compact stripes: good L2 locality for the gathers).  The
value tensors -- the six cameras' feature maps and the ``[prev_bev, bev]`` pair -- are
replicated; with history there is no per-layer exchange and the only collective on the data
path is ONE all-gather of the output rows at encoder exit (NCCL over NVLink / NVSwitch).
The reference has no counterpart (its only strategy is DDP, apis/mmdet_train.py:71-85); batch-level
data parallelism for training stays plain DDP over replicas of this encoder.

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


def sharded_encoder_forward(encoder, *args, group=None, **kwargs):
    """Run ``encoder`` on this rank's BEV rows and all-gather the full BEV.  ``encoder`` is any
    callable with the ``BEVFormerEncoder.forward`` signature that honours ``row_shard``."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return encoder(*args, **kwargs)
    local = encoder(*args, row_shard=(rank, world), **kwargs)
    return all_gather_bev_rows(local, kwargs['bev_h'], kwargs['bev_w'], group=group)
