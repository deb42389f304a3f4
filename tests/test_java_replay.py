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
        assert sorted(t for t, _ in row) == sorted(tgt[ptr[v]:ptr[v + 1]].tolist())
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
