import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
n, m = 100_000, 1_600_000
rp, ci = pkg.rmat(n, m, seed=42, symmetric=True)
ctx = pkg.Context(0)
g = pkg.Graph(ctx, rp, ci)
for thr in (1e-3, 5e-5):
    for T in (32, 64, 128, 256, 512, 1024, 4096):
        ctx.set_bwd_batch(T)
        hi = n if thr > 1e-4 else 20000
        g.backward_all_pairs(0.15, thr, -1, 0, 64)
        t0 = time.perf_counter(); ptr, tgt, val, st = g.backward_all_pairs(0.15, thr, -1, 0, hi); el = time.perf_counter() - t0
        print("thr %g T %4d targets %d: %.3f s  kernel %.1f ms batches %d edges %.3g -> %.3g edges/s" % (thr, T, hi, el, st["ms_push"], st["launches"], st["edges"], st["edges"] / (st["ms_push"] / 1e3)), flush=True)
