"""CPU: oracle mode A / C against a second, independent restatement of the Java loops (tests/java_replay.py).

Two transcriptions of the same text agreeing bit for bit in double arithmetic is what is checked; the
reference itself still cannot run here (parity unpinned)."""
import numpy as np
import pytest

import java_replay as J

ALPHA = 0.15


@pytest.fixture(scope="module")
def g(O, got):
    return O.OGraph(got[0], got[1])


def dense(d, n):
    out = np.zeros(n)
    for v, x in d.items():
        out[v] = x
    return out


def test_forward_push_bitwise_on_every_got_source(O, g, got):
    rp, ci = got[0].tolist(), got[1].tolist()
    n = len(rp) - 1
    for s in range(n):
        for rmax in (7.554e-4, 1e-6):                        # FORA's GOT r_max, and a deep push
            res, r, rsum, pushes, edges = J.forward_push(rp, ci, s, ALPHA, rmax)
            ores, orr, orsum, st = O.a_forward_push(g, s, ALPHA, rmax)
            assert np.array_equal(dense(res, n), ores) and np.array_equal(dense(r, n), orr), (s, rmax)
            assert (pushes, edges) == (st["pushes"], st["edges"]) and rsum == orsum, (s, rmax)


def test_forward_push_bitwise_on_rmat(O, pkg):
    rp, ci = pkg.rmat(1 << 9, 1 << 12, seed=5)
    g = O.OGraph(rp, ci)
    rpl, cil = rp.tolist(), ci.tolist()
    deg = np.diff(rp)
    srcs = [int(np.argmax(deg)), int(np.flatnonzero(deg > 0)[3]), int(np.flatnonzero(deg == 0)[0])]
    for s in srcs:
        res, r, rsum, pushes, edges = J.forward_push(rpl, cil, s, ALPHA, 1e-7)
        ores, orr, orsum, st = O.a_forward_push(g, s, ALPHA, 1e-7)
        assert np.array_equal(dense(res, g.n), ores) and np.array_equal(dense(r, g.n), orr)
        assert (pushes, edges, rsum) == (st["pushes"], st["edges"], orsum)


def test_backward_search_and_all_pair_index_bitwise(O, g, got):
    rp, ci = got[0].tolist(), got[1].tolist()
    n = len(rp) - 1
    for thr, want in [(1e-3, (2667, 8658, 1195)), (5e-7, (5553, 16530, 1775))]:   # SURVEY a10 replay counts
        rows, pushes, edges = J.base_all_pairs(rp, ci, ALPHA, thr)
        assert (pushes, edges, sum(len(r) for r in rows.values())) == want
        ptr, tgt, val, st = O.a_base_all_pairs(g, ALPHA, thr)
        assert (st["pushes"], st["edges"]) == (pushes, edges)
        for v in range(n):
            row = rows.get(v, [])
            assert [t for t, _ in row] == tgt[ptr[v]:ptr[v + 1]].tolist()
            assert [p for _, p in row] == val[ptr[v]:ptr[v + 1]].tolist()          # same doubles
    rows, _, _ = J.base_all_pairs(rp, ci, ALPHA, 1e-3, k=3)                          # J/Base_Whole_Graph.java:132-147
    ptr, tgt, val, _ = O.a_base_all_pairs(g, ALPHA, 1e-3, k=3)
    for v in range(n):
        row = rows.get(v, [])
        assert [p for _, p in row] == val[ptr[v]:ptr[v + 1]].tolist()
        assert [t for t, _ in row] == tgt[ptr[v]:ptr[v + 1]].tolist()                # List.sort is stable: ties stay in target order
    # single targets, including one without in-edges (J/Backward_Search.java:46-49)
    indeg = np.bincount(got[1], minlength=n)
    outdeg = np.diff(got[0]).tolist()
    order = np.argsort(got[1], kind="stable")
    rp_in = np.concatenate([[0], np.cumsum(indeg)]).tolist()
    ri_in = np.repeat(np.arange(n), np.diff(got[0]))[order].tolist()
    for t in (int(np.argmax(indeg)), int(np.flatnonzero(indeg == 0)[0]), 0):
        res, pushes, edges = J.backward_search(rp_in, ri_in, outdeg, t, ALPHA, 1e-5)
        ores, st = O.a_backward_search(g, t, ALPHA, 1e-5)
        assert np.array_equal(dense(res, n), ores) and (pushes, edges) == (st["pushes"], st["edges"])


def test_power_method_hashmap_order(O, g, got):
    rp, ci, names = got[0].tolist(), got[1].tolist(), got[2]
    n = len(rp) - 1
    for nm in ("Aemon", "Tyrion", "Grenn"):
        s = names.index(nm)
        mine = dense(J.power_method(rp, ci, s, ALPHA), n)
        theirs = O.power_method(g, s, ALPHA, 100)
        # same 100 Jacobi rounds; the oracle adds in vertex order, java.util.HashMap in bucket order
        assert np.abs(mine - theirs).max() < 1e-15
        assert np.count_nonzero(mine) == np.count_nonzero(theirs)
    s = names.index("Aemon")
    mine = J.power_method(rp, ci, s, ALPHA)
    for who, val in [("Aemon", 0.399076), ("Grenn", 0.202946), ("Samwell", 0.169607)]:   # SURVEY 8c anchors
        assert abs(mine[names.index(who)] - val) < 5e-7


def test_kth_largest_matches(O):
    rng = np.random.default_rng(3)
    vals = rng.integers(1, 20, 50).astype(np.float64) / 16.0           # plenty of ties
    for k in (1, 2, 7, 50):
        assert O.kth_ppr(vals, k) == J.kth_largest(vals.tolist(), k)
    assert O.kth_ppr(vals, 51) is None and J.kth_largest(vals.tolist(), 51) is None
    assert O.kth_ppr(vals, -1) is None and J.kth_largest(vals.tolist(), -1) is None


SEED, EPS, K = 7, 0.5, 10


def test_fora_around_the_walks(O, g, got):
    """J/Fora_Whole_Graph.java:82-146: balance rounds, walk counts and per-walk increments, with the
    oracle's Philox walk plugged into the replay -- the whole estimate must come out the same doubles."""
    rp, ci = got[0].tolist(), got[1].tolist()
    n = len(rp) - 1
    deg = np.diff(got[0])
    for s in (0, 63, int(np.argmax(deg)), int(np.flatnonzero(deg == 0)[0])):
        walk = lambda v, j: O.walk(g, v, v, j, SEED, 0, ALPHA, True)[0]
        for cost in (32.0, 6.0):
            mine, st = J.fora(rp, ci, s, ALPHA, EPS, walk, cost)
            theirs, ost = O.a_fora(g, s, ALPHA, EPS, SEED, walk_cost=cost)
            assert np.array_equal(dense(mine, n), theirs), (s, cost)
            assert all(st[k] == ost[k] for k in ("pushes", "edges", "rounds", "walks", "rsum")), (s, cost, st, ost)
            assert st["rmax_final"] == ost["rmax_final"]


def test_fora_topk_around_the_walks(O, g, got):
    """J/Fora_Topk.java:102-184 with J/Forward_Push.java:144-250 underneath."""
    rp, ci = got[0].tolist(), got[1].tolist()
    n = len(rp) - 1
    deg = np.diff(got[0])
    for s in (0, 63, int(np.argmax(deg)), int(np.flatnonzero(deg == 0)[0])):
        for k in (10, 3):
            walk = lambda v, j, rnd: O.walk(g, v, v, j, SEED, rnd, ALPHA, False)[0]
            mine, st = J.fora_topk(rp, ci, s, ALPHA, EPS, k, walk)
            theirs, ost = O.a_fora_topk(g, s, ALPHA, EPS, k, SEED)
            assert np.array_equal(dense(mine, n), theirs), (s, k)
            assert all(st[x] == ost[x] for x in ("pushes", "edges", "rounds", "walks")), (s, k, st, ost)


def test_monte_carlo_count_and_scaling(O, g, got):
    n = len(got[0]) - 1
    for s in (0, 63):
        walk = lambda v, j: O.walk(g, v, v, j, SEED, 0, ALPHA, False)[0]
        mine, walks = J.monte_carlo(n, s, EPS, walk)                # J/Monte_Carlo.java:136-158
        theirs, st = O.a_monte_carlo(g, s, ALPHA, EPS, SEED)
        assert walks == st["walks"] and np.array_equal(dense(mine, n), theirs)


def test_conf_replay(O):
    for n, k in ((107, 10), (1 << 20, 500), (4847571, 500)):
        c, w = O.conf(n, k), J.conf_topk(n, k)
        assert (c["delta"], c["min_delta"], c["pfail"]) == (w["delta"], w["min_delta"], w["pfail"])
        c, w = O.conf(n), J.conf_whole(n)
        assert (c["delta"], c["pfail"]) == (w["delta"], w["pfail"])


def test_walk_replay_with_philox_draws(O, g, got):
    """The Java walk loop (java_replay.random_walk) fed with the draws the SPEC assigns to each step
    (oracle/ppr_oracle.cpp header): one Philox block per step, counter (start, j lo, j hi, step);
    nextDouble = the 53-bit fraction of (x0, x1), nextInt(d) = mulhi((x2<<32)|x3, d) of the same block;
    the mandatory first move of a no-zero-hop walk takes a block of its own."""
    rp, ci = got[0].tolist(), got[1].tolist()
    deg = np.diff(got[0])

    def replay(start, j, seed, tag, nzh):
        k = (seed + tag * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        st = {"step": 0, "blk": None}

        def block():
            x = O.philox([start, j & 0xFFFFFFFF, j >> 32, st["step"]], [k & 0xFFFFFFFF, k >> 32])
            st["step"] += 1
            return [int(v) for v in x]

        def next_double():
            st["blk"] = x = block()
            return ((x[0] << 21) | (x[1] >> 11)) / 2.0 ** 53

        def next_int(bound):
            x = st["blk"] if st["blk"] is not None else block()
            st["blk"] = None
            return (((x[2] << 32) | x[3]) * bound) >> 64

        return J.random_walk(rp, ci, start, ALPHA, next_double, next_int, no_zero_hop=nzh)

    dang = int(np.flatnonzero(deg == 0)[0])
    for start in (63, 0, int(np.argmax(deg)), dang):
        for nzh in (False, True):
            for j in list(range(300)) + [2 ** 32 + 5]:
                assert replay(start, j, SEED, 0, nzh) == O.walk(g, start, start, j, SEED, 0, ALPHA, nzh), (start, j, nzh)
    assert replay(63, 9, SEED, 3, True) == O.walk(g, 63, 63, 9, SEED, 3, ALPHA, True)
    # walks that pass through dangling vertices were among those (36 of GOT's 107 vertices are dangling)
    draws = iter([0.9, 0.9, 0.9, 0.1])
    assert J.random_walk(rp, ci, 63, ALPHA, lambda: next(draws), lambda b: 0)[1] == 3


def _random_digraph(rng, n, m, self_loops=True, multi=True):
    src = rng.integers(0, n, m)
    dst = rng.integers(0, n, m)
    if not self_loops:
        keep = src != dst
        src, dst = src[keep], dst[keep]
    if not multi:
        pairs = np.unique(np.stack([src, dst], 1), axis=0)
        src, dst = pairs[:, 0], pairs[:, 1]
    order = np.argsort(src, kind="stable")
    src, dst = src[order], dst[order]
    rp = np.zeros(n + 1, np.int64)
    np.add.at(rp, src + 1, 1)
    return np.cumsum(rp), dst.astype(np.int32)


def test_replay_on_random_small_digraphs(O):
    """Self-loops, parallel edges, dangling and isolated vertices, sources without out-edges, targets without in-edges:
    the two restatements must agree bit for bit on all of them (forward push, backward search, all-pair index, FORA)."""
    rng = np.random.default_rng(20240607)
    for trial in range(24):
        n = int(rng.integers(2, 40))
        m = int(rng.integers(1, 5 * n))
        rp, ci = _random_digraph(rng, n, m, self_loops=trial % 2 == 0, multi=trial % 3 != 0)
        if ci.shape[0] == 0:
            continue
        g = O.OGraph(rp, ci)
        rpl, cil = rp.tolist(), ci.tolist()
        for s in rng.integers(0, n, 3).tolist():
            rmax = float(10.0 ** -rng.integers(2, 8))
            res, r, rsum, pushes, edges = J.forward_push(rpl, cil, s, ALPHA, rmax)
            ores, orr, orsum, st = O.a_forward_push(g, s, ALPHA, rmax)
            assert np.array_equal(dense(res, n), ores) and np.array_equal(dense(r, n), orr), (trial, s, rmax)
            assert (pushes, edges, rsum) == (st["pushes"], st["edges"], orsum), (trial, s, rmax)
            walk = lambda v, j: O.walk(g, v, v, j, SEED, 0, ALPHA, True)[0]
            mine, fst = J.fora(rpl, cil, s, ALPHA, EPS, walk, 32.0)
            theirs, ost = O.a_fora(g, s, ALPHA, EPS, SEED)
            assert np.array_equal(dense(mine, n), theirs), (trial, s)
            assert (fst["rounds"], fst["walks"]) == (ost["rounds"], ost["walks"])
        thr = float(10.0 ** -rng.integers(2, 6))
        for k in (-1, 2):
            rows, pushes, edges = J.base_all_pairs(rpl, cil, ALPHA, thr, k)
            ptr, tgt, val, st = O.a_base_all_pairs(g, ALPHA, thr, k=k)
            assert (pushes, edges) == (st["pushes"], st["edges"]), (trial, thr, k)
            for v in range(n):
                row = rows.get(v, [])
                assert [p for _, p in row] == val[ptr[v]:ptr[v + 1]].tolist(), (trial, thr, k, v)
                assert [t for t, _ in row] == tgt[ptr[v]:ptr[v + 1]].tolist(), (trial, thr, k, v)   # ties keep target order (stable sort)


def test_compute_error_replay(O, pkg, g, got):
    """J/Gen_Util.java:259-326 three ways: the replay, the oracle (orc_topk_metrics / orc_max_err) and the harness' host path."""
    n = len(got[0]) - 1
    exact = O.power_method(g, 63, ALPHA)
    est, _ = O.a_fora(g, 63, ALPHA, EPS, SEED)
    est[5] += 0.25 if exact[5] == 0 else 0.0                     # an estimate on a vertex the ground truth does not hold is NOT counted
    for k in (3, 10, 50):
        gids = O.topk_select(exact, k).tolist()
        aids = O.topk_select(est, k).tolist()
        gtop = {v: float(exact[v]) for v in gids}
        prec = J.compute_error("PRECISION", None, gids, gtop, None, aids)
        ndcg = J.compute_error("NDCG", None, gids, gtop, None, aids)
        m = O.topk_metrics(exact, k, aids)
        assert prec == m["precision"] and ndcg == m["ndcg"]

        class Stub:
            def __init__(self, dense, ids, top=None):
                self.dense, self.ids, self.top = dense, ids, top
            def getTopKNodeIds(self, k):
                return np.asarray(self.ids, np.int32)
            def getTopK(self, k):
                return self.top
            def getWholeGraphPPR(self):
                return pkg.algo.PPRVector(self.dense)
        pm, al = Stub(exact, gids, gtop), Stub(est, aids)
        assert pkg.compute_error(al, pm, "PRECISION", k) == prec
        assert pkg.compute_error(al, pm, "NDCG", k) == ndcg
    gmap = {v: float(exact[v]) for v in range(n) if exact[v] > 0}
    amap = {v: float(est[v]) for v in range(n) if est[v] > 0}
    mx = J.compute_error("MAX_ERR", gmap, None, None, amap, None)
    assert mx == O.max_err(exact, est) == pkg.compute_error(Stub(est, []), Stub(exact, []), "MAX_ERR", -1)
    assert mx < 0.25
