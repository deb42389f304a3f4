"""ncu target: all-pair backward search on configs[3] for a target range (one k_bwd_cta launch per call after the first)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
thr = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-3
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
rp, ci = pkg.rmat(100000, 1600000, seed=42, symmetric=True)
ctx = pkg.Context(0)
g = pkg.Graph(ctx, rp, ci)
for rep in range(2):
    ptr, tgt, val, st = g.backward_all_pairs(0.15, thr, -1, 0, hi)
    print(rep, int(ptr[-1]), st["edges"], st["ms_push"], flush=True)
g.close(); ctx.close()
