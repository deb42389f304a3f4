"""Measurements for the other BASELINE.json configs (run on a GPU box; under torchrun for N > 1):

  topk      config 3: LiveJournal-shape synthetic (n = 4.8 M, m = 69 M), FORA top-k = 500
  allpairs  config 4: all-pair backward search, symmetrised R-MAT with n = 100 000, m = 1.6 M
  walks     config 5 shape: Monte-Carlo / FORA walk phase on a large R-MAT (default scale 24; 25 needs ~40 GB of host RAM)
  sharded   config 5 through the library's own collectives (ppr_comm_init, ppr_*_sharded): one query at a time over all GPUs against
            the same query on one GPU (every rank also runs it alone), CUDA-event times, max over ranks

Prints one JSON object per config.  N > 1: walks are sharded with one NCCL all-reduce per query; all-pair
targets are block-partitioned with a gather at the end.
"""
import argparse, importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
ALPHA = 0.15

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["topk", "allpairs", "walks", "sharded"])
ap.add_argument("--queries", type=int, default=8)
ap.add_argument("--scale", type=int, default=24)
ap.add_argument("--edge-factor", type=int, default=16)
ap.add_argument("--thresholds", default="1e-3,5e-4,5e-5")
ap.add_argument("--cpu", type=int, default=1)
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
ctx = pkg.Context(local)


def sync():
    ctx.sync()
    if dist is not None:
        dist.barrier()


def srcs_of(rp, k, seed=1234):
    cand = np.flatnonzero(np.diff(rp) > 0)
    return [int(x) for x in cand[np.random.default_rng(seed).integers(0, len(cand), k)]]


out = {"config": a.what, "n_gpus": world}
if a.what == "topk":
    n, m, k = 4_800_000, 69_000_000, 500
    rp, ci = pkg.rmat(n, m, seed=42)
    g = pkg.Graph(ctx, rp, ci)
    S = srcs_of(rp, a.queries + 2)
    def topk(s, seed):     # N > 1: every round's walks sharded over the ranks, one NCCL all-reduce per delta-round
        if world > 1:
            return pkg.dist.fora_topk_sharded(g, s, ALPHA, 0.5, k, seed, rank, world)
        return g.fora_topk(s, ALPHA, 0.5, k, seed)

    for s in S[:2]:
        topk(s, 1)
    sync(); t0 = time.perf_counter()
    agg = {"ms_push": 0.0, "ms_walk": 0.0, "ms_select": 0.0, "ms_total": 0.0, "rounds": 0, "walks": 0, "edges": 0}
    sizes = []
    for i, s in enumerate(S[2:]):
        ids, vals, st = topk(s, 100 + i)
        sizes.append(len(ids))
        assert (np.diff(vals) <= 0).all() and len(ids) >= min(k, 1)
        for kk in agg:
            agg[kk] += st[kk]
    sync(); el = time.perf_counter() - t0
    out.update(workload="LJ-shape R-MAT n=4.8M m=69M, FORA top-k=500, eps 0.5" + (", walks sharded over %d GPUs (walks = this rank's share)" % world if world > 1 else ""), queries=a.queries, queries_per_s=a.queries / el,
               ms_per_query=el / a.queries * 1e3, returned=float(np.mean(sizes)), **{kk: v / a.queries for kk, v in agg.items()})
    if a.cpu and rank == 0:
        from oracle import oracle as O
        og = O.OGraph(rp, ci)
        t0 = time.perf_counter(); est, st = O.a_fora_topk(og, S[2], ALPHA, 0.5, k, 100); cpu = time.perf_counter() - t0
        ids, _, _ = g.fora_topk(S[2], ALPHA, 0.5, k, 100)
        ref = set(O.topk_select(est, k).tolist())
        out["cpu_port_s_per_query"] = cpu
        out["topk_overlap_vs_reference_order_replay"] = len(ref & set(ids.tolist())) / max(len(ref), 1)
elif a.what == "allpairs":
    n, m = 100_000, 1_600_000
    rp, ci = pkg.rmat(n, m, seed=42, symmetric=True)
    g = pkg.Graph(ctx, rp, ci)
    lo, hi = pkg.dist.target_blocks(n, world)[rank]
    res = {}
    for thr in [float(x) for x in a.thresholds.split(",")]:
        g.backward_all_pairs(ALPHA, thr, -1, lo, min(hi, lo + 64))   # warm-up
        sync(); t0 = time.perf_counter()
        ptr, tgt, val, st = g.backward_all_pairs(ALPHA, thr, -1, lo, hi, raw=True)
        t1 = time.perf_counter()
        pairs = len(tgt)
        if dist is not None:
            ptr, tgt, val = pkg.dist.gather_pairs((ptr, tgt, val), n)
        sync(); el = time.perf_counter() - t0
        res["%g" % thr] = {"seconds_total": el, "seconds_local_compute": t1 - t0, "kernel_ms": st["ms_push"], "pairs_total": int(len(tgt)), "pairs_local": pairs,
                           "pushes_local": st["pushes"], "edges_local": st["edges"], "batches": st["launches"],
                           "algorithmic_GBs": (48.0 * st["pushes"] + 24.0 * st["edges"]) / (st["ms_push"] / 1e3) / 1e9}
        if a.cpu and rank == 0 and world == 1:
            from oracle import oracle as O
            og = O.OGraph(rp, ci)
            sample = 200
            t0 = time.perf_counter(); O.a_base_all_pairs(og, ALPHA, thr, -1, 0, sample); c = time.perf_counter() - t0
            res["%g" % thr]["cpu_port_s_extrapolated"] = c / sample * n
            optr, otgt, oval, _ = O.b_base_all_pairs(og, ALPHA, thr, 0, sample)
            sptr, stgt, sval, _ = g.backward_all_pairs(ALPHA, thr, -1, 0, sample, raw=True)
            res["%g" % thr]["first_%d_targets_bit_exact" % sample] = bool(np.array_equal(optr, sptr) and np.array_equal(otgt, stgt) and np.array_equal(oval, sval))
    out.update(workload="symmetrised R-MAT n=100000 m=1.6M, all-pair backward search, targets block-partitioned", results=res)
elif a.what == "sharded":
    n = 1 << a.scale
    t0 = time.perf_counter()
    rp, ci = pkg.rmat(n, n * a.edge_factor, seed=42)
    gen_s = time.perf_counter() - t0
    g = pkg.Graph(ctx, rp, ci)
    if world > 1:
        box = [pkg.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], rank, world)
    S = srcs_of(rp, a.queries + 1)
    BAL = pkg._native.BALANCE_PER_SHARD

    def run(fn):     # warm-up on S[0], then the mean over the other sources; CUDA events on the library stream, max over ranks
        fn(S[0], 0)
        sync()
        acc = {}
        ctx.timer_record(0)
        for i, s in enumerate(S[1:]):
            st = fn(s, 1 + i)
            for k in ("ms_push", "ms_walk", "ms_comm", "walks_total", "walks", "steps", "edges", "rounds"):
                acc[k] = acc.get(k, 0) + st[k]
        ctx.timer_record(1)
        ms = ctx.timer_elapsed_ms(0, 1)
        sync()
        if world > 1:
            ms = float(ctx.comm_max([ms])[0])
        q = len(S) - 1
        r = {k: v / q for k, v in acc.items()}
        r["ms_per_query"] = ms / q
        return r

    res = {}
    res["monte_carlo_1gpu"] = run(lambda s, i: g.monte_carlo_shard(s, ALPHA, 0.5, 50 + i, 0, 1))
    res["fora_1gpu"] = run(lambda s, i: g.fora_shard(s, ALPHA, 0.5, 50 + i, 0, 1))
    if world > 1:
        res["monte_carlo_sharded"] = run(lambda s, i: g.monte_carlo_sharded(s, ALPHA, 0.5, 50 + i, resident=True)[1])
        res["fora_sharded"] = run(lambda s, i: g.fora_sharded(s, ALPHA, 0.5, 50 + i, resident=True)[1])
        res["fora_sharded_balance_per_shard"] = run(lambda s, i: g.fora_sharded(s, ALPHA, 0.5, 50 + i, flags=BAL, resident=True)[1])
        for k in ("monte_carlo", "fora"):
            res[k + "_speedup"] = res[k + "_1gpu"]["ms_per_query"] / res[k + "_sharded"]["ms_per_query"]
        res["fora_speedup_balance_per_shard"] = res["fora_1gpu"]["ms_per_query"] / res["fora_sharded_balance_per_shard"]["ms_per_query"]
        res["walks_only_speedup"] = res["monte_carlo_1gpu"]["ms_walk"] / res["monte_carlo_sharded"]["ms_walk"]
        res["comm"] = ctx.comm_info()
    res["monte_carlo_walks_per_s"] = res["monte_carlo_1gpu"]["walks_total"] / (res["monte_carlo_sharded" if world > 1 else "monte_carlo_1gpu"]["ms_per_query"] / 1e3)
    out.update(workload="R-MAT scale-%d edge factor %d (n=%d, m=%d), eps 0.5, %d non-dangling sources; sharded = one query at a time over %d GPU(s), ncclAllReduce inside the timed region"
                        % (a.scale, a.edge_factor, n, n * a.edge_factor, a.queries, world), rmat_gen_s=gen_s, results=res)
    if world > 1:
        ctx.comm_destroy()
else:
    n = 1 << a.scale
    rp, ci = pkg.rmat(n, n * a.edge_factor, seed=42)
    g = pkg.Graph(ctx, rp, ci)
    S = srcs_of(rp, 3)
    res = {}
    for name, eps in (("monte_carlo", 0.5), ("fora", 0.5)):
        best = None
        for it, s in enumerate(S):
            sync(); t0 = time.perf_counter()
            st = g.monte_carlo_shard(s, ALPHA, eps, 5, rank, world) if name == "monte_carlo" else g.fora_shard(s, ALPHA, eps, 5, rank, world)
            t1 = time.perf_counter()
            if dist is not None:
                pkg.dist.allreduce_result(g)
            sync(); t2 = time.perf_counter()
            if it:  # first is warm-up
                cur = {"walks_total": st["walks_total"], "walk_kernel_ms_rank0": st["ms_walk"], "push_ms": st["ms_push"], "allreduce_ms": (t2 - t1) * 1e3,
                       "total_ms": (t2 - t0) * 1e3, "walks_per_s": st["walks_total"] / (st["ms_walk"] / 1e3) * (1 if world == 1 else 1), "steps_rank0": st["steps"]}
                if dist is not None:
                    import torch
                    t = torch.tensor([st["ms_walk"]], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    cur["walk_kernel_ms_max"] = t.item(); cur["walks_per_s"] = st["walks_total"] / (t.item() / 1e3)
                if best is None or cur["total_ms"] < best["total_ms"]:
                    best = cur
        raw = g.last_raw(False)[0]
        best["sum_ok"] = bool(int(raw.sum(dtype=np.uint64)) == (pkg.FX_ONE if name == "fora" else best["walks_total"])) if (world == 1 or True) else None
        res[name] = best
    out.update(workload="R-MAT scale-%d (n=%d, m=%d) walks sharded over %d GPU(s)" % (a.scale, n, n * a.edge_factor, world), results=res)
if rank == 0:
    print(json.dumps(out), flush=True)
g.close(); ctx.close()
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
