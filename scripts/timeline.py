"""Per-iteration timeline of the push kernel for a few queries (profiling aid; run on the GPU box)."""
import importlib, sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wc = float(sys.argv[3]) if len(sys.argv) > 3 else None
rp, ci = pkg.rmat(1 << scale, 16 << scale, seed=42)
ctx = pkg.Context(0)
if wc is not None:
    ctx.set_balance(wc, 64)
g = pkg.Graph(ctx, rp, ci)
deg = np.diff(rp)
cand = np.flatnonzero(deg > 0)
srcs = cand[np.random.default_rng(1234).integers(0, len(cand), nq + 2)]
for s in srcs[:2]:
    g.fora_shard(int(s), 0.15, 0.5, 1, 0, 1)
for s in srcs[2:]:
    st = g.fora_shard(int(s), 0.15, 0.5, 1, 0, 1)
    tl = g.debug_timeline().astype(np.int64)
    print(json.dumps({k: st[k] for k in ("pushes", "edges", "iters", "rounds", "walks", "ms_push", "ms_walk")}))
    prev_e = 0
    print(" it   frontier      edges    A_us    B_us  mode  Medges/s")
    for i, (t0, t1, t2, f, e, mode) in enumerate(tl):
        de = e - prev_e; prev_e = e
        extra = ""
        if mode >> 8:
            extra = "  warp busy mean %.1f us max %.1f us" % (((mode >> 8) & ((1 << 28) - 1)) / 1e3, (mode >> 36) / 1e3)
        print("%3d %10d %10d %7.1f %7.1f %4d %9.1f%s" % (i, f, de, (t1 - t0) / 1e3, (t2 - t1) / 1e3, mode & 255, de / max(t2 - t1, 1) * 1e3, extra))
    tot = (tl[-1, 2] - tl[0, 0]) / 1e6
    print("sum A %.3f ms, sum B %.3f ms, span %.3f ms" % ((tl[:, 1] - tl[:, 0]).sum() / 1e6, (tl[:, 2] - tl[:, 1]).sum() / 1e6, tot))
