"""Sweep of the push/walk balance constant and the SCAN threshold (tuning aid; run on the GPU box)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rp, ci = pkg.rmat(1 << scale, 16 << scale, seed=42)
ctx = pkg.Context(0)
g = pkg.Graph(ctx, rp, ci)
cand = np.flatnonzero(np.diff(rp) > 0)
srcs = cand[np.random.default_rng(1234).integers(0, len(cand), nq)]
n = g.n
print("scale", scale, "queries", nq)
for thr_div in (1, 2, 4, 8):
    for wc in (3, 6, 12, 16, 24, 32, 64):
        ctx.set_balance(float(wc), 64)
        ctx.set_tuning(n // thr_div, -1)
        tp = tw = 0.0; rounds = 0; edges = 0; walks = 0
        for i, s in enumerate(srcs):
            st = g.fora_shard(int(s), 0.15, 0.5, 1000 + i, 0, 1)
            if i == 0:
                continue  # warm-up
            tp += st["ms_push"]; tw += st["ms_walk"]; rounds += st["rounds"]; edges += st["edges"]; walks += st["walks"]
        k = nq - 1
        print("scan_thr n/%d walk_cost %3d: push %.3f ms walk %.3f ms total %.3f ms  rounds %.1f  edges %.3g walks %.3g" %
              (thr_div, wc, tp / k, tw / k, (tp + tw) / k, rounds / k, edges / k, walks / k), flush=True)
