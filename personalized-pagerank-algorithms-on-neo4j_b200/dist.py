"""Multi-GPU plumbing: one process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch on GPUs,
gloo on CPU for the host-logic tests).  The path shards in two places:

* walks (FORA's walk phase, Monte-Carlo): every rank replays the deterministic push, runs the walks
  whose global index lies in its contiguous slice, and the raw fixed-point partial vectors are summed
  with ONE integer all-reduce.  Integer sums are exact, so the result is bit-identical to the 1-GPU run
  for any rank count.
* all-pair backward search: targets are block-partitioned; no communication until the final gather.

Independent single-source queries can also be dealt round-robin to ranks (replicas); that is reported
separately and is not a scaling claim.
"""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of `total` items owned by `rank` (same formula as the kernels' walk split)."""
    return (total * rank) // world, (total * (rank + 1)) // world


def target_blocks(n, world):
    """Block partition of the all-pair targets (J/Base_Whole_Graph.java:76: forEachNode)."""
    return [shard_range(n, r, world) for r in range(world)]


class _DevVec:
    """Zero-copy view of a device buffer for torch (via __cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}


def allreduce_result(graph, group=None):
    """Sum the raw uint64 result vector of the last query across ranks, in place on the device.

    int64 addition has the same bit pattern as uint64 addition (two's complement wrap), which is what
    NCCL's ncclInt64/ncclSum computes; torch has no uint64 reduction.
    """
    import torch
    import torch.distributed as dist
    ptr, n = graph.result_devptr()
    t = torch.as_tensor(_DevVec(ptr, n), device="cuda:%d" % graph.ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.synchronize(graph.ctx.device)
    return t


def fora_sharded(graph, src, alpha, eps, seed, rank, world, flags=0, out=None, group=None):
    """FORA with walks sharded over `world` GPUs; every rank ends with the full vector."""
    st = graph.fora_shard(src, alpha, eps, seed, rank, world, flags)
    if world > 1:
        allreduce_result(graph, group)
    return graph.result_export(0.0, out), st


def monte_carlo_sharded(graph, src, alpha, eps, seed, rank, world, out=None, group=None):
    st = graph.monte_carlo_shard(src, alpha, eps, seed, rank, world)
    if world > 1:
        allreduce_result(graph, group)
    return graph.result_export(st["omega"], out), st


def merge_pairs(parts, n):
    """Concatenate per-rank inverted indexes (src_ptr, tgt, val) whose target blocks ascend with rank."""
    ptr = np.zeros(n + 1, np.int64)
    for p, _, _ in parts:
        ptr[1:] += np.diff(p)
    np.cumsum(ptr[1:], out=ptr[1:])
    total = int(ptr[-1])
    tgt = np.zeros(total, parts[0][1].dtype if parts else np.int32)
    val = np.zeros(total, parts[0][2].dtype if parts else np.float64)
    pos = ptr[:-1].copy()
    for p, t, v in parts:  # rank order = ascending targets, so rows stay sorted by target
        cnt = np.diff(p)
        if t.size == 0:
            continue
        rows = np.repeat(np.arange(n), cnt)
        off = np.arange(t.size) - np.repeat(p[:-1], cnt)
        dst = pos[rows] + off
        tgt[dst] = t
        val[dst] = v
        pos += cnt
    return ptr, tgt, val


def gather_pairs(local, n, group=None):
    """All-gather the per-rank (src_ptr, tgt, val) results of a block-partitioned all-pair run."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    return merge_pairs(parts, n)
