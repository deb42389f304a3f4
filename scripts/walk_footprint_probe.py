"""Walk throughput against the size of the edge array (R-MAT scale 20, m = 8..20 M): where does the L2 stop holding it?"""
import importlib, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
n = 1 << 20
ctx = pkg.Context(0)
for mm in (8, 10, 12, 14, 16, 18, 20):
    rp, ci = pkg.rmat(n, mm << 20, seed=42)
    g = pkg.Graph(ctx, rp, ci)
    cand = np.flatnonzero(np.diff(rp) > 0)
    srcs = cand[np.random.default_rng(1234).integers(0, len(cand), 8)]
    tw = steps = walks = 0
    for i, s in enumerate(srcs):
        st = g.fora_shard(int(s), 0.15, 0.5, 1, 0, 1)
        if i >= 2:
            tw += st["ms_walk"]; steps += st["steps"]; walks += st["walks"]
    print("m %2d M (edge array %3d MB): %.1f G moves/s, %.2f G walks/s" % (mm, mm * 4, steps / tw / 1e6, walks / tw / 1e6), flush=True)
    g.close()
