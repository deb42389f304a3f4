import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
n, m = 100_000, 1_600_000
rp, ci = pkg.rmat(n, m, seed=42, symmetric=True)
ctx = pkg.Context(0)
g = pkg.Graph(ctx, rp, ci)
for thr in (1e-3, 5e-5):
    for T in (128, 1024):
        ctx.set_bwd_batch(T)
        g.backward_all_pairs(0.15, thr, -1, 0, T)
        t0 = time.perf_counter(); ptr, tgt, val, st = g.backward_all_pairs(0.15, thr, -1, 0, 4096); el = time.perf_counter() - t0
        print("thr %g T %d: 4096 targets %.1f ms wall, kernel span %.1f ms, batches %d, iterations %d (%.1f per batch, %.1f us each), pushes %d edges %.3g pairs %d" %
              (thr, T, el * 1e3, st["ms_push"], st["launches"], st["iters"], st["iters"] / st["launches"], st["ms_push"] * 1e3 / max(st["iters"], 1), st["pushes"], st["edges"], len(tgt)), flush=True)
