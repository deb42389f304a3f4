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


def _sum_over_ranks(device, ptr, n, group=None):
    """In-place element-wise sum over ranks of n raw uint64 values at device address `ptr` (int64 bit pattern)."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(_DevVec(ptr, n), device="cuda:%d" % device)
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    else:                                   # gloo (tests: several ranks sharing one GPU): stage through the host
        h = t.cpu()
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
        t.copy_(h)
    torch.cuda.synchronize(device)


def fora_topk_sharded(graph, src, alpha, eps, k, seed, rank, world, flags=0, conf=None, group=None):
    """FORA top-k with every round's walks sharded over `world` GPUs (ppr_fora_topk_shard); one integer all-reduce
    of the estimate vector per delta-round.  Every rank returns the same (ids, values, stats)."""
    reduce = (lambda ptr, n: _sum_over_ranks(graph.ctx.device, ptr, n, group)) if world > 1 else None
    return graph.fora_topk_shard(src, alpha, eps, k, seed, rank, world, reduce, flags, conf)


def monte_carlo_sharded(graph, src, alpha, eps, seed, rank, world, out=None, group=None):
    st = graph.monte_carlo_shard(src, alpha, eps, seed, rank, world)
    if world > 1:
        allreduce_result(graph, group)
    return graph.result_export(st["omega"], out), st


def topk_rows(ptr, tgt, val, k):
    """Base_Whole_Graph's per-source cut (J/Base_Whole_Graph.java:133-148): keep the entries >= the k-th largest value of a
    row, sorted by value descending (stable: ties keep ascending targets).  Rows with fewer than k entries keep everything."""
    n = ptr.shape[0] - 1
    optr = np.zeros(n + 1, np.int64)
    ot, ov = [], []
    for v in range(n):
        t, x = tgt[ptr[v]:ptr[v + 1]], val[ptr[v]:ptr[v + 1]]
        if 1 <= k <= x.shape[0]:
            kth = np.partition(x, x.shape[0] - k)[x.shape[0] - k]
            keep = np.flatnonzero(x >= kth)
        else:
            keep = np.arange(x.shape[0])
        keep = keep[np.argsort(-x[keep].astype(np.float64) if x.dtype.kind == "f" else (np.iinfo(np.uint64).max - x[keep].astype(np.uint64)), kind="stable")]
        ot.append(t[keep]); ov.append(x[keep])
        optr[v + 1] = optr[v] + keep.shape[0]
    return optr, (np.concatenate(ot) if ot else tgt[:0]), (np.concatenate(ov) if ov else val[:0])


def merge_pairs(parts, n, k=-1):
    """Concatenate per-rank inverted indexes (src_ptr, tgt, val) whose target blocks ascend with rank.

    The parts must be FULL rows (computed with k < 0): a rank's own top-k cut is not the cut of the merged row.  With
    k >= 0 the cut is applied here, after the merge, over the union of the ranks' rows (as ppr_backward_all_pairs_sharded does)."""
    cnts = [np.diff(p[0]) for p in parts]
    ptr = np.zeros(n + 1, np.int64)
    np.cumsum(sum(cnts), out=ptr[1:])
    total = int(ptr[-1])
    tgt = np.empty(total, parts[0][1].dtype if parts else np.int32)
    val = np.empty(total, parts[0][2].dtype if parts else np.float64)
    start = ptr[:-1].copy()                       # where the next part's slice of row v begins
    for (p, t, v), cnt in zip(parts, cnts):       # rank order = ascending targets, so rows stay sorted by target
        if t.size:
            dst = np.repeat(start - p[:-1], cnt)  # row-wise shift from the part's position to the merged position
            dst += np.arange(t.size, dtype=np.int64)
            tgt[dst] = t
            val[dst] = v
        start += cnt
    if k >= 0:
        return topk_rows(ptr, tgt, val, k)
    return ptr, tgt, val


def gather_pairs(local, n, group=None, k=-1):
    """All-gather the per-rank (src_ptr, tgt, val) results of a block-partitioned all-pair run (full rows, k < 0 on every
    rank); k >= 0 applies the per-source top-k cut to the merged rows.

    With NCCL the three arrays travel as padded device tensors (one all_gather each, over NVLink); with gloo
    (CPU tests) as pickled objects.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if dist.get_backend(group) != "nccl":
        parts = [None] * world
        dist.all_gather_object(parts, local, group=group)
        return merge_pairs(parts, n, k)
    dev = torch.device("cuda", torch.cuda.current_device())
    ptr, tgt, val = local
    cnt = torch.tensor([tgt.shape[0]], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    cnts = [int(c.item()) for c in cnts]
    mx = max(max(cnts), 1)

    def gather(arr, length, np_dtype, t_dtype):
        buf = torch.zeros(length, dtype=t_dtype, device=dev)
        if arr.shape[0]:
            buf[:arr.shape[0]] = torch.from_numpy(np.ascontiguousarray(arr).view(np_dtype)).to(dev)
        out = torch.empty(world * length, dtype=t_dtype, device=dev)
        dist.all_gather_into_tensor(out, buf, group=group)
        return out.view(world, length)

    P = gather(ptr, n + 1, np.int64, torch.int64)
    T = gather(tgt, mx, np.int32, torch.int32)
    V = gather(val, mx, np.int64, torch.int64)
    # merge on the device (same arithmetic as merge_pairs): row v of the result is the concatenation, in rank order,
    # of row v of every part -- rank order = ascending target blocks, so rows stay sorted by target
    cnt = P[:, 1:] - P[:, :-1]                                   # [world, n]
    mptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(cnt.sum(0), 0, out=mptr[1:])
    before = torch.cumsum(cnt, 0) - cnt                          # pairs of row v held by lower ranks
    total = int(mptr[-1].item())
    mt = torch.empty(total, dtype=torch.int32, device=dev)
    mv = torch.empty(total, dtype=torch.int64, device=dev)
    for r in range(world):
        if cnts[r]:
            shift = mptr[:-1] + before[r] - P[r, :-1]            # part position -> merged position, per row
            dst = torch.repeat_interleave(shift, cnt[r], output_size=cnts[r])
            dst += torch.arange(cnts[r], dtype=torch.int64, device=dev)
            mt[dst] = T[r, :cnts[r]]
            mv[dst] = V[r, :cnts[r]]
    out = (mptr.cpu().numpy(), mt.cpu().numpy(), mv.cpu().numpy().view(val.dtype))
    return topk_rows(*out, k) if k >= 0 else out
