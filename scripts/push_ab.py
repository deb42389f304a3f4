"""A/B timing of the push kernel (tuning aid; run on the GPU box).

    PPRB200_LIB=.../variants/libpprb200_x.so python scripts/push_ab.py SCALE NQ [timeline]

Prints one JSON line: mean push / walk time per query, edges/s, algorithmic GB/s (48 V_p + 24 E_p) and the per-mode
sums of the in-kernel timeline of the last query (mode bits: 1 sweep start, 2 dense B, 4 single CTA, 8 small regime)."""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 16
show = len(sys.argv) > 3
rp, ci = pkg.rmat(1 << scale, 16 << scale, seed=42)
ctx = pkg.Context(0)
if os.environ.get("WALK_COST"):
    ctx.set_balance(float(os.environ["WALK_COST"]), 64)
g = pkg.Graph(ctx, rp, ci)
srcs = np.random.default_rng(1234).integers(0, 1 << scale, nq + 3)
acc = {k: 0.0 for k in ("pushes", "edges", "iters", "rounds", "walks", "steps", "ms_push", "ms_walk")}
for i, s in enumerate(srcs):
    st = g.fora_shard(int(s), 0.15, 0.5, 1000 + i, 0, 1)
    if i >= 3:
        for k in acc:
            acc[k] += st[k]
nd = np.flatnonzero(np.diff(rp) > 0)
g.fora_shard(int(nd[np.random.default_rng(7).integers(0, len(nd))]), 0.15, 0.5, 999, 0, 1)   # timeline of a non-dangling source
tl = g.debug_timeline().astype(np.int64)
out = {"lib": os.path.basename(pkg._native.LIB_PATH), "scale": scale, "nq": nq,
       "ms_push": acc["ms_push"] / nq, "ms_walk": acc["ms_walk"] / nq, "iters": acc["iters"] / nq,
       "Gedges_s": acc["edges"] / acc["ms_push"] / 1e6 if acc["ms_push"] else 0,
       "push_GBs": (48 * acc["pushes"] + 24 * acc["edges"]) / acc["ms_push"] / 1e6 if acc["ms_push"] else 0,
       "walks_per_s": acc["walks"] / acc["ms_walk"] * 1e3 if acc["ms_walk"] else 0}
out["frac"] = out["push_GBs"] / 6545.9
if len(tl):
    a, b, mode = (tl[:, 1] - tl[:, 0]) / 1e3, (tl[:, 2] - tl[:, 1]) / 1e3, tl[:, 5] & 255
    cls = {"dense(2,3)": (mode == 2) | (mode == 3), "list(0,1)": (mode == 0) | (mode == 1), "mini(4)": mode == 4, "small(8)": mode == 8}
    out["last_query"] = {k: {"n": int(m.sum()), "A_us": round(float(a[m].sum()), 1), "B_us": round(float(b[m].sum()), 1)} for k, m in cls.items()}
    out["last_query"]["span_ms"] = float(tl[-1, 2] - tl[0, 0]) / 1e6
print(json.dumps(out), flush=True)
if show:
    prev = 0
    for i, (t0, t1, t2, f, e, m) in enumerate(tl):
        print("%3d %10d %10d %7.1f %7.1f %4d" % (i, f, e - prev, (t1 - t0) / 1e3, (t2 - t1) / 1e3, m & 255)); prev = e
g.close(); ctx.close()
