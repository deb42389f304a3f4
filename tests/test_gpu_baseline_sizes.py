"""GPU (-m gpu): bit-exact parity against the oracle AT THE SIZES OF BASELINE.json's configs (SURVEY 8c checks 1, 3, 6).

test_gpu_parity.py covers GOT and R-MAT 2^12 / 2^16 exhaustively; this file runs one or two sources of each full-size
configuration through the same oracle (mode B: synchronous schedule, 2^-62 fixed point, Philox walk replay), so that the
code paths only large graphs select -- the L2-hint walk kernel, walk_cost 24, multi-batch reverse push, long dense push
rounds -- meet the oracle too.  Integer results, compared with ==.  CPU cost of the oracle: ~1 minute in total.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ALPHA, EPS, SEED = 0.15, 0.5, 7


def _nondangling(rp, k, seed=1234):
    nd = np.flatnonzero(np.diff(rp) > 0)
    return [int(x) for x in nd[np.random.default_rng(seed).integers(0, len(nd), k)]]


def _check_fora(g, og, O, s, walk_cost):
    ppr, st = g.fora(s, ALPHA, EPS, SEED)
    raw, r = g.last_raw()
    oraw, ost = O.b_fora(og, s, ALPHA, EPS, SEED, walk_cost=walk_cost)
    assert np.array_equal(raw, oraw)
    for k in ("pushes", "edges", "iters", "rounds", "walk_sources", "steps"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    assert st["walks"] == ost["walks"] and st["rmax_fx"] == ost["rmax_fx_final"] and st["rsum_fx"] == ost["rsum_fx"]
    assert int(raw.sum(dtype=np.uint64)) == O.FX_ONE and not r.any()
    return st


def test_scale20_fora_and_forward_push(pkg, ctx, O):
    """configs[1]: R-MAT scale-20.  One FORA query and one forward push, every bit of the result and every work counter."""
    rp, ci = pkg.rmat(1 << 20, 1 << 24, seed=42)
    g, og = pkg.Graph(ctx, rp, ci), O.OGraph(rp, ci)
    try:
        s = _nondangling(rp, 1)[0]
        st = _check_fora(g, og, O, s, 6.0)          # footprint 72 MB <= 96 MiB: the in-L2 balance constant
        assert st["edges"] > 50_000_000 and st["walks"] > 1_000_000
        _, _, _, pst = g.forward_push(s, ALPHA, 1e-7, want_residue=False)
        res, r = g.last_raw()
        ores, orr, ost = O.b_forward_push(og, s, ALPHA, 1e-7)
        assert np.array_equal(res, ores) and np.array_equal(r, orr)
        assert (pst["pushes"], pst["edges"], pst["iters"]) == (ost["pushes"], ost["edges"], ost["iters"])
        # Monte-Carlo counts (a smaller omega than eps = 0.5 would give: 1.1e7 walks) against the replay
        g.monte_carlo(s, ALPHA, 2.0, SEED)
        cnt = g.last_raw(False)[0]
        ocnt, ost = O.b_monte_carlo(og, s, ALPHA, 2.0, SEED)
        assert np.array_equal(cnt, ocnt)
    finally:
        g.close()


def test_walk_kernel_with_l2_hints(pkg, ctx, O, rmat16):
    """k_walk<true, 2> (eviction-priority hints) is what every graph beyond the L2 runs.  Forced on R-MAT 2^16 ..."""
    g, og = pkg.Graph(ctx, *rmat16), O.OGraph(*rmat16)
    try:
        for variant in (2, 1, 0, 3, 4, 5):   # 2 = hints + 32-bit row pointers (what graphs beyond the L2 run), 1 = hints, 0 = plain, 3-5 = edge-sector retention experiments
            ctx.set_walk_variant(variant)
            for s in _nondangling(rmat16[0], 2):
                _check_fora(g, og, O, s, 6.0)
                g.monte_carlo(s, ALPHA, 4.0, SEED)
                assert np.array_equal(g.last_raw(False)[0], O.b_monte_carlo(og, s, ALPHA, 4.0, SEED)[0])
    finally:
        ctx.set_walk_variant(-1)
        g.close()


def test_scale21_natural_large_graph_path(pkg, ctx, O):
    """... and selected naturally on R-MAT scale-21 (4m + 8n = 151 MB > 96 MiB), where the balance constant is 24 as well."""
    rp, ci = pkg.rmat(1 << 21, 1 << 25, seed=42)
    g, og = pkg.Graph(ctx, rp, ci), O.OGraph(rp, ci)
    try:
        st = _check_fora(g, og, O, _nondangling(rp, 1)[0], 24.0)
        assert st["rounds"] >= 2
    finally:
        g.close()


def test_config4_all_pairs_first_targets(pkg, ctx, O):
    """configs[3]: n = 100 000, m = 1.6 M symmetrised.  First 500 targets, thresholds 1e-3 and 5e-5, forced into >= 2 batches."""
    rp, ci = pkg.rmat(100000, 1600000, seed=42, symmetric=True)
    g, og = pkg.Graph(ctx, rp, ci), O.OGraph(rp, ci)
    try:
        ctx.set_bwd_batch(192)
        for thr in (1e-3, 5e-5):
            ptr, tgt, val, st = g.backward_all_pairs(ALPHA, thr, -1, 0, 500, raw=True)
            optr, otgt, oval, ost = O.b_base_all_pairs(og, ALPHA, thr, 0, 500)
            assert np.array_equal(ptr, optr) and np.array_equal(tgt, otgt) and np.array_equal(val, oval), thr
            assert (st["pushes"], st["edges"]) == (ost["pushes"], ost["edges"])
        ctx.set_bwd_batch(0)
        ptr, tgt, val, st = g.backward_all_pairs(ALPHA, 1e-3, -1, 250, 500, raw=True)    # a block that does not start at 0
        optr, otgt, oval, _ = O.b_base_all_pairs(og, ALPHA, 1e-3, 250, 500)
        assert np.array_equal(ptr, optr) and np.array_equal(tgt, otgt) and np.array_equal(val, oval)
    finally:
        ctx.set_bwd_batch(0)
        g.close()


def test_lj_shape_topk(pkg, ctx, O):
    """configs[2]: LiveJournal-shape (n = 4.8 M, m = 69 M), FORA top-k = 500: estimate vector, ids and rounds of one source."""
    rp, ci = pkg.rmat(4800000, 69000000, seed=42)
    g, og = pkg.Graph(ctx, rp, ci), O.OGraph(rp, ci)
    try:
        s = _nondangling(rp, 1)[0]
        ids, vals, st = g.fora_topk(s, ALPHA, EPS, 500, SEED)
        est = g.last_raw(False)[0]
        oest, ost = O.b_fora_topk(og, s, ALPHA, EPS, 500, SEED)
        assert np.array_equal(est, oest)
        assert st["rounds"] == ost["rounds"] and st["pushes"] == ost["pushes"] and st["edges"] == ost["edges"]
        assert set(ids.tolist()) == set(O.topk_select(oest, 500).tolist()) and len(ids) >= 500
        assert np.array_equal(vals, oest[ids] * O.FX_INV) and (np.diff(vals) <= 0).all()
    finally:
        g.close()
