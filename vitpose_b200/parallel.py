"""Data-parallel sharding of crops over the GPUs of one box, and the final result gather.

The path shards naturally (independent crops, BatchNorm in eval mode): one process per GPU, weights replicated,
no data-path collective. The only exchange is the gather of per-crop results at the end:

* ``strided_shard``  — the split ``DistributedSampler`` makes for test loaders
  (mmpose/datasets/samplers/distributed_sampler.py:34-41: pad to a multiple of world, take rank::world);
* ``contiguous_shard`` — rank r takes crops [r*n/G, (r+1)*n/G) (what bench.py uses);
* ``gather_strided`` — re-interleave and truncate exactly like ``collect_results_gpu``
  (mmpose/apis/test.py:179-223: zip(*parts), extend, [:size]) but on fixed-shape float tensors with one
  ``all_gather`` instead of two all_gathers of pickled, max-padded byte blobs;
* ``gather_contiguous`` — concatenation in rank order.

Training: the one real exchange step of the path is the gradient all-reduce —
``allreduce_gradients`` averages ``.grad`` of the replicated parameters over the ranks in a few large flat
buckets (what mmcv's DDP wrapper does for the reference, one NCCL all-reduce per bucket over NVLink/NVSwitch).

Works with any torch.distributed backend: NCCL over NVLink on the B200 box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def strided_shard(n_total, rank, world):
    """Indices this rank evaluates; padded by wrapping so every rank gets ceil(n/world)."""
    per_rank = (n_total + world - 1) // world
    total = per_rank * world
    idx = list(range(n_total))
    idx += idx[:total - n_total]
    if len(idx) < total:                       # n_total < padding (tiny inputs): keep wrapping
        while len(idx) < total:
            idx += list(range(n_total))[:total - len(idx)]
    return idx[rank:total:world]


def contiguous_shard(n_total, rank, world):
    lo = n_total * rank // world
    hi = n_total * (rank + 1) // world
    return lo, hi


def _all_gather_equal(t):
    rank, world = world_info()
    if world == 1:
        return [t]
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return parts


def gather_strided(local, n_total):
    """local: [ceil(n/world), ...] results of strided_shard order -> [n_total, ...] in dataset order."""
    parts = _all_gather_equal(local)
    inter = torch.stack(parts, dim=1)          # [per_rank, world, ...] == zip(*parts)
    return inter.reshape((-1,) + tuple(local.shape[1:]))[:n_total]


def gather_contiguous(local, counts=None):
    """local: this rank's [n_local, ...] block; counts: per-rank sizes when they differ (padded gather)."""
    rank, world = world_info()
    if world == 1:
        return local
    if counts is None:
        return torch.cat(_all_gather_equal(local), dim=0)
    m = max(counts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = _all_gather_equal(pad)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def allreduce_gradients(params, bucket_bytes=256 << 20):
    """Average ``p.grad`` over all ranks, in place. Gradients are packed into flat fp32 buckets of about
    ``bucket_bytes`` (parameter order), each reduced with one all_reduce(SUM) and scaled by 1/world.
    Every rank must pass the same parameters in the same order. Returns the number of collectives issued."""
    rank, world = world_info()
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return 0
    buckets, cur, size = [], [], 0
    for g in grads:
        nbytes = g.numel() * g.element_size()
        if cur and size + nbytes > bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
        cur.append(g)
        size += nbytes
    if cur:
        buckets.append(cur)
    for b in buckets:
        flat = torch.cat([g.reshape(-1) for g in b])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in b:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    return len(buckets)
