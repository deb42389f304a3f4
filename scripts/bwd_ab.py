"""All-pair backward search on BASELINE configs[3] (n = 100 000, m = 1.6 M symmetrised): wall / kernel / sort time per
threshold (tuning aid; PPR_BWD_BD / PPR_BWD_CPS override the CTA shape).  usage: python scripts/bwd_ab.py [thr ...]"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
thrs = [float(x) for x in sys.argv[1:]] or [1e-3, 5e-4, 5e-5]
rp, ci = pkg.rmat(100000, 1600000, seed=42, symmetric=True)
ctx = pkg.Context(0)
g = pkg.Graph(ctx, rp, ci)
g.backward_all_pairs(0.15, 1e-3, -1, 0, 2000)   # warm-up: module load, slot allocation
for thr in thrs:
    for rep in range(2):
        t0 = time.perf_counter()
        ptr, tgt, val, st = g.backward_all_pairs(0.15, thr, -1)
        wall = time.perf_counter() - t0
    gbs = (48.0 * st["pushes"] + 24.0 * st["edges"]) / (st["ms_push"] / 1e3) / 1e9
    print(json.dumps({"thr": thr, "bd": os.environ.get("PPR_BWD_BD"), "cps": os.environ.get("PPR_BWD_CPS"), "wall_s": round(wall, 4),
                      "kernel_ms": round(st["ms_push"], 2), "sort_ms": round(st["ms_select"], 2), "pairs": int(ptr[-1]), "edges": st["edges"],
                      "pushes": st["pushes"], "iters": st["iters"], "launches": st["launches"],
                      "Gedges_s": round(st["edges"] / st["ms_push"] / 1e6, 2), "algorithmic_GBs": round(gbs, 1), "frac": round(gbs / 6545.9, 4)}), flush=True)
g.close(); ctx.close()
