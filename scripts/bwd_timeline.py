import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
n, m = 100_000, 1_600_000
rp, ci = pkg.rmat(n, m, seed=42, symmetric=True)
ctx = pkg.Context(0)
g = pkg.Graph(ctx, rp, ci)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 5e-5
ctx.set_bwd_batch(T)
g.backward_all_pairs(0.15, thr, -1, 0, T)
t0 = time.perf_counter(); ptr, tgt, val, st = g.backward_all_pairs(0.15, thr, -1, 0, T); el = time.perf_counter() - t0
tl = g.debug_timeline().astype(np.int64)
print("T %d thr %g: wall %.2f ms, stream span %.2f ms, iterations %d" % (T, thr, el * 1e3, st["ms_push"], len(tl)))
prev = 0
for i, (a, b, c, f, e, _) in enumerate(tl):
    print("%3d items %8d edges %9d  A %7.1f us  B %7.1f us   %.1f Medges/s" % (i, f, e - prev, (b - a) / 1e3, (c - b) / 1e3, (e - prev) / max(c - b, 1) * 1e3)); prev = e
print("sum A %.2f ms sum B %.2f ms, first-to-last %.2f ms" % ((tl[:, 1] - tl[:, 0]).sum() / 1e6, (tl[:, 2] - tl[:, 1]).sum() / 1e6, (tl[-1, 2] - tl[0, 0]) / 1e6))
