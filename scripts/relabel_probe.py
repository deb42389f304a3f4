"""Does a degree-ordered vertex numbering help?  Same R-MAT graph under three labellings (as generated = random
permutation, in-degree descending, out-degree descending); FORA push / walk time per query.  Run on the GPU box."""
import importlib, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n = 1 << scale
rp, ci = pkg.rmat(n, 16 << scale, seed=42)
src = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
dst = ci.astype(np.int64)
cand = np.flatnonzero(np.diff(rp) > 0)
srcs = cand[np.random.default_rng(1234).integers(0, len(cand), nq + 2)]
ctx = pkg.Context(0)
indeg = np.bincount(dst, minlength=n)
outdeg = np.diff(rp)
for name, key in (("as generated", None), ("in-degree descending", indeg), ("out-degree descending", outdeg), ("in+out descending", indeg + outdeg)):
    if key is None:
        nid = np.arange(n, dtype=np.int64)
    else:
        order = np.argsort(-key, kind="stable")       # order[new] = old
        nid = np.empty(n, np.int64); nid[order] = np.arange(n)
    rp2, ci2 = pkg.graphio.csr_from_edges(n, nid[src], nid[dst])
    g = pkg.Graph(ctx, rp2, ci2)
    for s in srcs[:2]:
        g.fora_shard(int(nid[s]), 0.15, 0.5, 1, 0, 1)
    tp = tw = 0.0; walks = edges = 0
    for s in srcs[2:]:
        st = g.fora_shard(int(nid[s]), 0.15, 0.5, 1, 0, 1)
        tp += st["ms_push"]; tw += st["ms_walk"]; walks += st["walks"]; edges += st["edges"]
    print("scale %d %-22s push %.3f ms  walk %.3f ms per query  (%.1f M edges, %.1f M walks)" % (scale, name, tp / nq, tw / nq, edges / nq / 1e6, walks / nq / 1e6), flush=True)
    g.close()
