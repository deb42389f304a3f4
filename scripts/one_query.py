"""A handful of FORA queries from non-dangling sources on one R-MAT graph (ncu target: short, one GPU)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rp, ci = pkg.rmat(1 << scale, 16 << scale, seed=42)
ctx = pkg.Context(0)
g = pkg.Graph(ctx, rp, ci)
cand = np.flatnonzero(np.diff(rp) > 0)
srcs = cand[np.random.default_rng(1234).integers(0, len(cand), nq)]
for i, s in enumerate(srcs):
    st = g.fora_shard(int(s), 0.15, 0.5, 1000 + i, 0, 1)
    print(i, {k: st[k] for k in ("pushes", "edges", "iters", "rounds", "walks", "steps", "ms_push", "ms_walk")}, flush=True)
g.close(); ctx.close()
