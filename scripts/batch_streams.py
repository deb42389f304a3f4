"""ppr_fora_batch with one vs two queries in flight (tuning aid; GPU box)."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rp, ci = pkg.rmat(1 << scale, 16 << scale, seed=42)
ctx = pkg.Context(0)
g = pkg.Graph(ctx, rp, ci)
cand = np.flatnonzero(np.diff(rp) > 0)
srcs = cand[np.random.default_rng(1234).integers(0, len(cand), nq)].astype(np.int32)
out = ctx.pinned((nq, g.n), np.float64)
ref = None
for streams in (1, 2, 1, 2):
    ctx.set_batch_streams(streams)
    g.fora_batch(srcs[:4], 0.15, 0.5, 1000, out=out[:4], want_stats=False)
    ctx.sync(); t0 = time.perf_counter()
    g.fora_batch(srcs, 0.15, 0.5, 1000, out=out, want_stats=False)
    ctx.sync(); el = time.perf_counter() - t0
    if ref is None:
        ref = out.copy()
    print("scale %d streams %d: %d queries in %.1f ms -> %.1f q/s  identical to first run: %s" % (scale, streams, nq, el * 1e3, nq / el, bool(np.array_equal(ref, out))), flush=True)
