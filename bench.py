#!/usr/bin/env python
"""bench.py -- FORA single-source PPR throughput on a synthetic R-MAT graph (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale 20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one batch of --queries-per-step FORA queries (alpha 0.15, eps 0.5, delta = p_f = 1/n as in
J/Algo_Conf.java:45-53) from distinct random sources.  Rank 0 prints ONE JSON line.

* value  : queries/s with the graph resident in HBM and results left on the device (C-ABI call
           ppr_fora_shard per query; only the 200-byte control block comes back).
* e2e    : queries/s through the reference-facing call (ppr_fora_batch = Fora_Whole_Graph.computeWholeGraphPPR
           per source): host source ids in, dense n-vector of doubles copied to pinned host memory, inside the timed region.
* roofline: dominant kernel k_push_query; algorithmic bytes 48*V_p + 24*E_p over its CUDA-event time.
* cpu_baseline / --impl reference: the line-by-line C++ restatement of the Java loops (oracle mode A; the
           reference itself is Java on Neo4j and cannot run in this image), timed on the box's host cores.
Multi-GPU (N > 1): queries are independent, so they are dealt to ranks with no data-path collective
("weak": every rank runs --queries-per-step queries per step); timing is max over ranks.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "personalized-pagerank-algorithms-on-neo4j_b200"
ALPHA, EPS = 0.15, 0.5
# Push/walk balance (stand-in for the reference's wall-clock balancer, J/Fora_Whole_Graph.java:93-103, which
# adapts to the machine it runs on): each arm uses the constant that matches its own measured costs.
GPU_WALK_COST = None   # library default: 6 while the graph fits in L2, 24 beyond (ppr_set_balance)
CPU_WALK_COST = 32.0   # host core: one walk ~ 32 relaxed edges (the reference's 400 ns / walk)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=20, help="R-MAT scale (configs[1] of BASELINE.json is scale 20)")
    ap.add_argument("--edge-factor", type=int, default=16)
    ap.add_argument("--queries-per-step", type=int, default=8)
    ap.add_argument("--eps", type=float, default=EPS)
    ap.add_argument("--walk-cost", type=float, default=None, help="override the push/walk balance constant")
    ap.add_argument("--also-scale", type=int, default=22, help="extra short measurement at this scale (0 = off)")
    ap.add_argument("--cpu-baseline", type=int, default=1)
    ap.add_argument("--ref-threads", type=int, default=0)
    return ap.parse_args()


def make_graph(pkg, scale, ef, seed=42):
    n = 1 << scale
    t0 = time.time()
    rp, ci = pkg.rmat(n, n * ef, seed=seed)
    return rp, ci, time.time() - t0


def pick_sources(rp, count, seed=1234):
    """Uniform over vertices with out-degree > 0 (a dangling source is answered by {s: 1} without any work;
    48 % of the R-MAT vertices are dangling)."""
    deg = np.diff(rp)
    cand = np.flatnonzero(deg > 0)
    rng = np.random.default_rng(seed)
    return cand[rng.integers(0, len(cand), count)].astype(np.int32)


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device):
        self.device, self.proc, self.rows = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_push_query launch, from `ncu --set full` captures of
# scripts/one_query.py at library defaults (profiles/r01_ncu_full_summary.txt); bytes per launch, keyed by R-MAT scale
NCU_PUSH_TRAFFIC = {20: 0.334086e9 + 48.295680e6, 22: 6.547190e9 + 1.555525e9}


def cpu_port_query(O, og, src, eps, seed, walk_cost):
    t0 = time.perf_counter()
    _, st = O.a_fora(og, int(src), ALPHA, eps, seed, walk_cost=walk_cost)
    return time.perf_counter() - t0, st


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path = oracle mode A (C++ port), all host threads."""
    if rank != 0:
        return
    pkg = importlib.import_module(PKG)
    from oracle import oracle as O
    O.lib()
    rp, ci, _ = make_graph(pkg, args.scale, args.edge_factor)
    og = O.OGraph(rp, ci)
    threads = args.ref_threads or max(1, min(os.cpu_count() or 1, 64))
    walk_cost = CPU_WALK_COST if args.walk_cost is None else args.walk_cost
    srcs = pick_sources(rp, (args.steps + args.warmup) * threads)

    def step(i):
        res = [None] * threads

        def work(t):
            res[t] = cpu_port_query(O, og, srcs[i * threads + t], args.eps, 1000 + i * threads + t, walk_cost)
        ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return time.perf_counter() - t0

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        step(i)
    el = time.perf_counter() - t0
    qps = args.steps * threads / el
    line = {
        "impl": "reference", "metric": "FORA SSPPR queries/s", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, threads),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": "%d queries per step, one per host thread (oracle mode A: C++ restatement of the Java loops; "
                                   "the Java/Neo4j reference cannot run in this image)" % threads},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, q):
    n = 1 << args.scale
    return {"workload": "R-MAT scale-%d (n=%d, m=%d, a/b/c/d=0.57/0.19/0.19/0.05, relabelled, seed 42) FORA single-source PPR "
                        "(BASELINE.json configs[1] shape)" % (args.scale, n, n * args.edge_factor),
            "alpha": ALPHA, "eps": args.eps, "delta": "1/n", "p_f": "1/n", "queries_per_step": q,
            "sources": "uniform over vertices with out-degree > 0, seed 1234", "cpu_walk_cost": CPU_WALK_COST,
            "l2": "L2 flushed (256 MiB write) before every step; working set at scale 20 is ~100 MB vs 126 MB L2"}


def measure_ours(pkg, ctx, g, rp, args, srcs, steps, warmup, q, rank, world, dist, sample_clocks):
    """Timed region of the product arm.  Returns dict with device-resident and e2e numbers + per-kernel sums."""
    n = g.n
    out = ctx.pinned((q, n), np.float64)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda:%d" % ctx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    acc = {"pushes": 0, "edges": 0, "walks": 0, "steps": 0, "ms_push": 0.0, "ms_walk": 0.0, "launches": 0, "iters": 0, "rounds": 0}

    def resident_step(i, record):
        ctx.flush_l2()
        for j in range(q):
            st = g.fora_shard(int(srcs[i * q + j]), ALPHA, args.eps, 1000 + i * q + j, 0, 1)
            if record:
                for k in acc:
                    acc[k] += st[k]

    for i in range(warmup):
        resident_step(i, False)
    barrier()
    sampler = ClockSampler(ctx.device) if sample_clocks else None
    if sampler:
        sampler.start()
    ctx.timer_record(0)
    for i in range(warmup, warmup + steps):
        resident_step(i, True)
    ctx.timer_record(1)
    ms = ctx.timer_elapsed_ms(0, 1)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(ms)

    # end to end: host source ids -> ppr_fora_batch -> dense doubles in pinned host memory
    def e2e_step(i):
        ctx.flush_l2()
        g.fora_batch(srcs[i * q:(i + 1) * q], ALPHA, args.eps, 1000 + i * q, out=out, want_stats=False)

    for i in range(warmup):
        e2e_step(i)
    barrier()
    ctx.timer_record(2)
    for i in range(warmup, warmup + steps):
        e2e_step(i)
    ctx.timer_record(3)
    ms_e2e = ctx.timer_elapsed_ms(2, 3)
    barrier()
    ms_e2e = max_over_ranks(ms_e2e)
    chk = float(out[q - 1].sum())

    # device-resident, pipelined (same batch call, results left on the device): what the overlap of two queries buys
    def pipe_step(i):
        ctx.flush_l2()
        g.fora_batch(srcs[i * q:(i + 1) * q], ALPHA, args.eps, 1000 + i * q, want_stats=False, resident=True)

    for i in range(warmup):
        pipe_step(i)
    barrier()
    ctx.timer_record(0)
    for i in range(warmup, warmup + steps):
        pipe_step(i)
    ctx.timer_record(1)
    ms_pipe = max_over_ranks(ctx.timer_elapsed_ms(0, 1))
    barrier()
    return {"ms": ms, "ms_e2e": ms_e2e, "ms_pipe": ms_pipe, "acc": acc, "clocks": clocks, "check_sum": chk}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    pkg = importlib.import_module(PKG)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    ctx = pkg.Context(local)
    if args.walk_cost is not None:
        ctx.set_balance(args.walk_cost, 64)
    walk_cost = GPU_WALK_COST if args.walk_cost is None else args.walk_cost
    cpu_walk_cost = CPU_WALK_COST if args.walk_cost is None else args.walk_cost
    q = args.queries_per_step
    rp, ci, gen_s = make_graph(pkg, args.scale, args.edge_factor)
    t0 = time.time()
    g = pkg.Graph(ctx, rp, ci)
    export_s = time.time() - t0
    total_q = (args.steps + args.warmup) * q
    all_srcs = pick_sources(rp, total_q * world)
    srcs = all_srcs[rank * total_q:(rank + 1) * total_q]   # ranks work on disjoint query sets, no collective on the data path

    m = measure_ours(pkg, ctx, g, rp, args, srcs, args.steps, args.warmup, q, rank, world, dist, rank == 0)
    acc = m["acc"]
    nq_rank = args.steps * q
    qps = nq_rank * world / (m["ms"] / 1e3)
    qps_e2e = nq_rank * world / (m["ms_e2e"] / 1e3)
    peak, peak_src = measured_peak()
    push_bytes = 48.0 * acc["pushes"] + 24.0 * acc["edges"]
    push_gbs = push_bytes / (acc["ms_push"] / 1e3) / 1e9 if acc["ms_push"] > 0 else 0.0
    walk_bytes = 20.0 * acc["steps"] + 16.0 * acc["walks"]
    walk_gbs = walk_bytes / (acc["ms_walk"] / 1e3) / 1e9 if acc["ms_walk"] > 0 else 0.0

    line = {
        "metric": "FORA SSPPR queries/s", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (2^-62 fixed point; f64 only at the boundary)", "data": "synthetic",
        "config": dict(workload_config(args, q), walk_cost=walk_cost if walk_cost is not None else "auto (6 in L2, 24 beyond)", parallelism="%d independent query streams (one per GPU)" % world),
        "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": 4 * q, "d2h_bytes_per_step": 8 * g.n * q,
                "api": "ppr_fora_batch (Fora_Whole_Graph.computeWholeGraphPPR per source), dense f64 result to pinned host memory; "
                       "the batch call keeps two queries in flight while the graph fits in L2, which is why e2e can exceed `value` "
                       "(measured one query at a time so that per-kernel CUDA-event times are unambiguous)"},
        "resident_pipelined": {"value": nq_rank * world / (m["ms_pipe"] / 1e3), "unit": "queries/s",
                               "api": "ppr_fora_batch with results left on the device"},
        "gpu_launches": int(acc["launches"]),
        "clocks": m["clocks"],
        "roofline": {"bound": "hbm", "kernel": "k_push_query", "achieved": push_gbs, "peak": peak, "unit": "GB/s",
                     "frac": push_gbs / peak, "traffic": NCU_PUSH_TRAFFIC.get(args.scale) if args.walk_cost is None else None,
                     "traffic_note": "DRAM bytes per launch from the ncu --set full capture in profiles/ (same scale, library defaults); "
                                     "algorithmic bytes per launch = %.3g" % (push_bytes / max(nq_rank, 1)),
                     "peak_source": peak_src,
                     "algorithmic_bytes": "48*V_p + 24*E_p", "V_p": acc["pushes"], "E_p": acc["edges"],
                     "launches": nq_rank, "avg_launch_ms": acc["ms_push"] / max(nq_rank, 1),
                     "share_of_step": acc["ms_push"] / m["ms"]},
        "kernels": {"k_push_query": {"ms": acc["ms_push"], "edges_per_s": acc["edges"] / (acc["ms_push"] / 1e3) if acc["ms_push"] else 0,
                                     "iters_per_query": acc["iters"] / max(nq_rank, 1), "rounds_per_query": acc["rounds"] / max(nq_rank, 1)},
                    "k_walk": {"ms": acc["ms_walk"], "walks_per_s": acc["walks"] / (acc["ms_walk"] / 1e3) if acc["ms_walk"] else 0,
                               "steps_per_s": acc["steps"] / (acc["ms_walk"] / 1e3) if acc["ms_walk"] else 0,
                               "algorithmic_GBs": walk_gbs, "frac_of_peak": walk_gbs / peak,
                               "walks_per_query": acc["walks"] / max(nq_rank, 1)}},
        "setup": {"rmat_gen_s": gen_s, "graph_export_s": export_s},
    }

    if rank == 0 and args.cpu_baseline:
        from oracle import oracle as O   # checker/baseline only: never on the product path
        O.lib()
        og = O.OGraph(rp, ci)
        el, st = cpu_port_query(O, og, srcs[args.warmup * q], args.eps, 1000 + args.warmup * q, cpu_walk_cost)
        line["cpu_baseline"] = {"value": 1.0 / el, "unit": "queries/s", "cores": 1, "kind": "port",
                                "sample": "1 query (first timed source) with oracle mode A, the C++ restatement of the reference's "
                                          "Java loops (walk_cost %g): %.1f s, %d edges relaxed, %d walks" % (cpu_walk_cost, el, st["edges"], st["walks"])}
    g.close()

    if rank == 0 and args.also_scale and args.also_scale != args.scale and world == 1:
        # the headline shape of BASELINE.json's metric string (R-MAT-22): short extra measurement, same code path
        try:
            a2 = argparse.Namespace(**vars(args))
            a2.scale = args.also_scale
            rp2, ci2, gen2 = make_graph(pkg, a2.scale, a2.edge_factor)
            g2 = pkg.Graph(ctx, rp2, ci2)
            q2, steps2, warm2 = 4, 2, 3
            s2 = pick_sources(rp2, (steps2 + warm2) * q2)
            m2 = measure_ours(pkg, ctx, g2, rp2, a2, s2, steps2, warm2, q2, 0, 1, None, False)
            a = m2["acc"]
            pg = (48.0 * a["pushes"] + 24.0 * a["edges"]) / (a["ms_push"] / 1e3) / 1e9
            line["rmat%d" % a2.scale] = {"value": steps2 * q2 / (m2["ms"] / 1e3), "unit": "queries/s",
                                         "e2e": steps2 * q2 / (m2["ms_e2e"] / 1e3), "queries": steps2 * q2,
                                         "push_GBs": pg, "push_frac": pg / peak,
                                         "walks_per_s": a["walks"] / (a["ms_walk"] / 1e3), "edges_per_query": a["edges"] / (steps2 * q2),
                                         "walks_per_query": a["walks"] / (steps2 * q2)}
            g2.close()
        except Exception as e:  # never lose the main line
            line["rmat%d" % args.also_scale] = {"error": str(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
