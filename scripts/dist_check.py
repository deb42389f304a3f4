"""Multi-GPU parity check (run under torchrun on a GPU box, one rank per GPU):
walk-sharded FORA / Monte-Carlo with an NCCL integer all-reduce must equal the 1-GPU vector and the
CPU oracle bit for bit; block-partitioned all-pair backward search must merge to the 1-GPU index.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
"""
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
from oracle import oracle as O  # checker only

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
ctx = pkg.Context(local)
ALPHA, EPS = 0.15, 0.5
ok = True
for name, (rp, ci) in (("GOT", pkg.graphio.load_got()[:2]), ("rmat14", pkg.rmat(1 << 14, 1 << 18, seed=42))):
    g = pkg.Graph(ctx, rp, ci)
    og = O.OGraph(rp, ci)
    deg = np.diff(rp)
    srcs = np.flatnonzero(deg > 0)[[3, 17]]
    for s in srcs:
        s = int(s)
        out, st = pkg.dist.fora_sharded(g, s, ALPHA, EPS, 7, rank, world)
        raw = g.last_raw(False)[0]
        ref, _ = O.b_fora(og, s, ALPHA, EPS, 7)
        good = bool(np.array_equal(raw, ref)) and int(raw.sum(dtype=np.uint64)) == pkg.FX_ONE
        out, st = pkg.dist.monte_carlo_sharded(g, s, ALPHA, 1.0, 7, rank, world)
        cnt = g.last_raw(False)[0]
        refc, rst = O.b_monte_carlo(og, s, ALPHA, 1.0, 7)
        good = good and bool(np.array_equal(cnt, refc)) and bool(np.array_equal(out, refc / rst["omega"]))
        ok = ok and good
        if rank == 0:
            print("%s src %d: sharded FORA + MC over %d GPUs bit-exact: %s (walks on rank0 %d of %d)" % (name, s, world, good, st["walks"], st["walks_total"]), flush=True)
    for s in srcs:      # FORA top-k: every round's walks sharded, one NCCL all-reduce per delta-round
        s = int(s)
        ids, vals, st = pkg.dist.fora_topk_sharded(g, s, ALPHA, EPS, 10, 7, rank, world)
        raw = g.last_raw(False)[0]
        ref, rst = O.b_fora_topk(og, s, ALPHA, EPS, 10, 7)
        rids = O.topk_select(ref, 10)
        good = bool(np.array_equal(raw, ref)) and set(ids.tolist()) == set(rids.tolist()) and st["rounds"] == rst["rounds"]
        ok = ok and good
        if rank == 0:
            print("%s src %d: sharded FORA top-10 over %d GPUs bit-exact: %s (%d rounds, %d ids)" % (name, s, world, good, st["rounds"], len(ids)), flush=True)
    lo, hi = pkg.dist.target_blocks(og.n, world)[rank]
    local_pairs = g.backward_all_pairs(ALPHA, 1e-3, -1, lo, hi, raw=True)[:3]
    ptr, tgt, val = pkg.dist.gather_pairs(local_pairs, og.n)
    rptr, rtgt, rval, _ = O.b_base_all_pairs(og, ALPHA, 1e-3)
    good = bool(np.array_equal(ptr, rptr) and np.array_equal(tgt, rtgt) and np.array_equal(val, rval))
    ok = ok and good
    if rank == 0:
        print("%s all-pair blocks over %d GPUs merge to the 1-GPU index: %s (%d pairs)" % (name, world, good, len(tgt)), flush=True)
    g.close()

# ---- the library's own collectives (ppr_comm_init + ppr_*_sharded: ncclAllReduce / ncclAllGather / ncclBroadcast behind the C ABI)
box = [pkg.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
ctx.comm_init(box[0], rank, world)
for name, (rp, ci) in (("GOT", pkg.graphio.load_got()[:2]), ("rmat14", pkg.rmat(1 << 14, 1 << 18, seed=42))):
    g = pkg.Graph(ctx, rp, ci)
    og = O.OGraph(rp, ci)
    srcs = [int(x) for x in np.flatnonzero(np.diff(rp) > 0)[[5, 23]]]
    good = True
    for s in srcs:
        out, st = g.fora_sharded(s, ALPHA, EPS, 7)
        ref, _ = O.b_fora(og, s, ALPHA, EPS, 7)
        good = good and bool(np.array_equal(g.last_raw(False)[0], ref)) and bool(np.array_equal(out, ref * O.FX_INV))
        # balance per shard: walk_cost / world (more push, fewer walks per GPU); the oracle takes the same constant
        out, st = g.fora_sharded(s, ALPHA, EPS, 7, flags=pkg._native.BALANCE_PER_SHARD)
        ref, _ = O.b_fora(og, s, ALPHA, EPS, 7, walk_cost=6.0 / world)
        good = good and bool(np.array_equal(g.last_raw(False)[0], ref))
        out, st = g.monte_carlo_sharded(s, ALPHA, 1.0, 7)
        refc, rst = O.b_monte_carlo(og, s, ALPHA, 1.0, 7)
        good = good and bool(np.array_equal(g.last_raw(False)[0], refc)) and bool(np.array_equal(out, refc / rst["omega"]))
        ids, vals, st = g.fora_topk_sharded(s, ALPHA, EPS, 10, 7)
        ref, rst = O.b_fora_topk(og, s, ALPHA, EPS, 10, 7)
        good = good and bool(np.array_equal(g.last_raw(False)[0], ref)) and set(ids.tolist()) == set(O.topk_select(ref, 10).tolist()) and st["rounds"] == rst["rounds"]
    rptr, rtgt, rval, _ = O.b_base_all_pairs(og, ALPHA, 1e-3)
    ptr, tgt, val, st = g.backward_all_pairs_sharded(ALPHA, 1e-3, -1, raw=True)
    good = good and bool(np.array_equal(ptr, rptr) and np.array_equal(tgt, rtgt) and np.array_equal(val, rval))
    kptr, ktgt, kval, _ = g.backward_all_pairs_sharded(ALPHA, 1e-3, 3, raw=True)      # top-3 rows: cut after the merge
    for v in range(og.n):
        rt, rv = rtgt[rptr[v]:rptr[v + 1]], rval[rptr[v]:rptr[v + 1]]
        keep = O.topk_select(rv, 3) if len(rv) else np.zeros(0, np.int32)
        good = good and bool(np.array_equal(ktgt[kptr[v]:kptr[v + 1]], rt[keep]) and np.array_equal(kval[kptr[v]:kptr[v + 1]], rv[keep]))
    ok = ok and good
    if rank == 0:
        print("%s library collectives (FORA, FORA per-shard balance, MC, FORA top-k, all-pair gather, top-k rows) over %d GPUs bit-exact: %s" % (name, world, good), flush=True)
    g.close()
if rank == 0:
    print("library communicator:", ctx.comm_info(), flush=True)

# walk scaling at a size where walks dominate: Monte-Carlo on R-MAT scale 20 (eps 0.5: 1.8e8 walks)
rp, ci = pkg.rmat(1 << 20, 1 << 24, seed=42)
g = pkg.Graph(ctx, rp, ci)
src = int(np.flatnonzero(np.diff(rp) > 0)[1234])
for it in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = g.monte_carlo_shard(src, ALPHA, EPS, 5, rank, world)
    t1 = time.perf_counter()
    pkg.dist.allreduce_result(g)
    torch.cuda.synchronize(); dist.barrier()
    t2 = time.perf_counter()
tt = torch.tensor([st["ms_walk"], (t2 - t1) * 1e3, (t2 - t0) * 1e3], device="cuda", dtype=torch.float64)
dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    W = st["walks_total"]
    print("MC scale-20 eps 0.5: %d walks over %d GPUs: walk kernel %.2f ms (max over ranks), allreduce(8 MB) %.2f ms, total %.2f ms -> %.3g walks/s"
          % (W, world, tt[0].item(), tt[1].item(), tt[2].item(), W / (tt[2].item() / 1e3)), flush=True)
    print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
g.close()
ctx.comm_destroy()
ctx.close()
dist.barrier()
dist.destroy_process_group()
