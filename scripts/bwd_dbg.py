import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
mode = sys.argv[1] if len(sys.argv) > 1 else "r12"
if mode == "r12":
    rp, ci = pkg.rmat(1 << 12, 1 << 16, seed=42); hi = 4096; thr = 1e-3
elif mode == "r12deep":
    rp, ci = pkg.rmat(1 << 12, 1 << 16, seed=42); hi = 300; thr = 1e-6
else:
    rp, ci = pkg.rmat(100000, 1600000, seed=42, symmetric=True); hi = 6000; thr = 1e-3
ctx = pkg.Context(0)
g = pkg.Graph(ctx, rp, ci)
try:
    ptr, tgt, val, st = g.backward_all_pairs(0.15, thr, -1, 0, hi)
    print("OK", int(ptr[-1]), st["edges"], st["launches"], flush=True)
except Exception as e:
    print("FAIL", repr(e)[:300], flush=True)
