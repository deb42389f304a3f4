"""Push time against the SCAN/LIST switch point (frontier edges >= threshold -> dense sweep).  Run on the GPU box."""
import importlib, sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n = 1 << scale
rp, ci = pkg.rmat(n, 16 << scale, seed=42)
ctx = pkg.Context(0)
g = pkg.Graph(ctx, rp, ci)
cand = np.flatnonzero(np.diff(rp) > 0)
srcs = cand[np.random.default_rng(1234).integers(0, len(cand), nq + 2)]
for mult in (0.125, 0.25, 0.5, 1, 2, 4, 8, 16):
    ctx.set_tuning(int(n * mult))
    for s in srcs[:2]:
        g.fora_shard(int(s), 0.15, 0.5, 1, 0, 1)
    tp = tw = 0.0
    for s in srcs[2:]:
        st = g.fora_shard(int(s), 0.15, 0.5, 1, 0, 1)
        tp += st["ms_push"]; tw += st["ms_walk"]
    print("scale %d threshold %6.3f n: push %.3f ms, walk %.3f ms per query" % (scale, mult, tp / nq, tw / nq), flush=True)
