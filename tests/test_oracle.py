"""CPU: the oracle against its external anchors, its invariants and the committed golden fixtures."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "got_golden.npz")
ALPHA, EPS, SEED, K = 0.15, 0.5, 7, 10


@pytest.fixture(scope="module")
def g(O, got):
    return O.OGraph(got[0], got[1])


def test_philox_known_answers(O):
    # Random123 kat_vectors, philox4x32-10
    assert [hex(x) for x in O.philox([0] * 4, [0] * 2)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in O.philox([0xffffffff] * 4, [0xffffffff] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_got_shape(got):
    rp, ci, names = got
    deg = np.diff(rp)
    assert len(names) == 107 and ci.shape[0] == 352           # SURVEY section 2 (fixture row)
    assert (deg == 0).sum() == 36 and deg.max() == 24 and names[int(deg.argmax())] == "Tyrion"
    indeg = np.bincount(ci, minlength=107)
    assert (indeg == 0).sum() == 7 and indeg.max() == 13 and names[int(indeg.argmax())] == "Robert"


def test_power_method_anchors(O, g, got):
    names = got[2]
    pm = O.power_method(g, names.index("Aemon"), ALPHA)        # SURVEY 8c sanity anchors
    for nm, val in [("Aemon", 0.399076), ("Grenn", 0.202946), ("Samwell", 0.169607), ("Craster", 0.038737), ("Karl", 0.032927)]:
        assert abs(pm[names.index(nm)] - val) < 5e-7
    assert abs(pm.sum() - (1 - 0.85 ** 100)) < 1e-12
    pm = O.power_method(g, names.index("Tyrion"), ALPHA)
    for nm, val in [("Tyrion", 0.391682), ("Margaery", 0.050106), ("Pycelle", 0.040263)]:
        assert abs(pm[names.index(nm)] - val) < 5e-7
    s = names.index("Grenn")                                   # dangling source
    pm = O.power_method(g, s, ALPHA)
    assert abs(pm[s] - (1 - 0.85 ** 100)) < 1e-12 and np.count_nonzero(pm) == 1


def test_conf_formulas(O):
    c = O.conf(107)
    assert c["delta"] == 1 / 107 and c["pfail"] == 1 / 107
    c = O.conf(107, 10)                                        # J/Algo_Conf.java:71-81, integer 107/10 = 10
    assert c["delta"] == 0.1 and c["min_delta"] == 1 / 107
    assert c["pfail"] == 1.0 / 107 / 107 / np.log(10.0)
    p = O.fora_params(107, 352, ALPHA, EPS)                    # SURVEY section 8 table, GOT row
    assert abs(p["rmax"] - 7.554e-4) < 1e-7 and abs(p["omega"] - 5741.6) < 0.1
    p = O.fora_params(1 << 22, 1 << 26, ALPHA, EPS)
    assert abs(p["rmax"] / 5.070e-9 - 1) < 1e-3 and abs(p["omega"] / 6.687e8 - 1) < 1e-3


def test_mode_a_push_invariants(O, g, got):
    rp = got[0]
    deg = np.diff(rp)
    exact = O.power_method(g, 0, ALPHA, 3000, 1e-17)
    for rmax in (1e-4, 1e-6, 1e-8):
        res, r, rsum, st = O.a_forward_push(g, 0, ALPHA, rmax)
        assert abs(res.sum() + r.sum() - 1) < 1e-12            # SURVEY 8c check 2
        assert (r[deg > 0] < rmax * deg[deg > 0]).all() and (r[deg == 0] == 0).all()
        assert np.abs(res - exact).max() <= r.sum() + 1e-12
    res, r, rsum, st = O.a_forward_push(g, 71, ALPHA, 1e-6)    # dangling source: {s: 1}, rsum 0
    assert res[71] == 1.0 and rsum == 0.0 and st["pushes"] == 0


def test_mode_b_push_invariants(O, g, got):
    deg = np.diff(got[0])
    for s in (0, 63, 5):
        for rmax in (1e-4, 1e-6, 1e-9):
            res, r, st = O.b_forward_push(g, s, ALPHA, rmax, trace=True)
            assert int(res.sum()) + int(r.sum()) == O.FX_ONE   # exact mass conservation
            rfx = max(1, int(rmax * 2.0 ** 62))
            thr = np.where(deg == 0, 1, rfx * deg.astype(object))
            assert all(int(r[v]) < thr[v] for v in range(len(deg)))
            assert sum(len(f) for f in st["frontiers"]) == st["pushes"] and len(st["frontiers"]) == st["iters"]
            assert all((np.diff(f) > 0).all() for f in st["frontiers"])
            a, _, _, _ = O.a_forward_push(g, s, ALPHA, rmax)   # both schedules approximate the same vector
            assert np.abs(res * O.FX_INV - a).max() < 2 * rmax * deg.max() + 1e-9


def test_walk_semantics(O, g, got):
    deg = np.diff(got[0])
    dang = int(np.flatnonzero(deg == 0)[0])
    assert O.walk(g, dang, dang, 0, SEED, 0, ALPHA, False) == (dang, 0)   # J/Monte_Carlo.java:70-72
    assert O.walk(g, dang, dang, 0, SEED, 0, ALPHA, True) == (dang, 0)
    steps0 = [O.walk(g, 63, 63, j, SEED, 0, ALPHA, False)[1] for j in range(4000)]
    steps1 = [O.walk(g, 63, 63, j, SEED, 0, ALPHA, True)[1] for j in range(4000)]
    assert abs(np.mean(steps0) - 0.85 / 0.15) < 0.35 and abs(np.mean(steps1) - (1 + 0.85 / 0.15)) < 0.35
    assert min(steps1) >= 1 and min(steps0) == 0
    assert abs(np.mean(np.array(steps0) == 0) - ALPHA) < 0.02
    # different seeds / tags give different streams
    a = [O.walk(g, 63, 63, j, 1, 0, ALPHA, True)[0] for j in range(64)]
    b = [O.walk(g, 63, 63, j, 2, 0, ALPHA, True)[0] for j in range(64)]
    c = [O.walk(g, 63, 63, j, 1, 1, ALPHA, True)[0] for j in range(64)]
    assert a != b and a != c


def test_monte_carlo(O, g):
    exact = O.power_method(g, 0, ALPHA, 3000, 1e-17)
    ppr, st = O.a_monte_carlo(g, 0, ALPHA, EPS, SEED)
    assert st["walks"] == 6889                                  # floor(6889.9)
    cnt, bst = O.b_monte_carlo(g, 0, ALPHA, EPS, SEED)
    assert int(cnt.sum()) == 6889 and (cnt / bst["omega"] == ppr).all()
    parts = sum(O.b_monte_carlo(g, 0, ALPHA, EPS, SEED, i, 3)[0] for i in range(3))
    assert (parts == cnt).all()
    big = exact > 1 / 107
    assert (np.abs(ppr - exact)[big] <= EPS * exact[big]).all()


def test_fora_guarantee_and_shards(O, g):
    bad = 0
    for s in (0, 63, 5, 100, 71):
        exact = O.power_method(g, s, ALPHA, 3000, 1e-17)
        f, st = O.b_fora(g, s, ALPHA, EPS, SEED, restart_at_source=True)
        assert int(f.sum()) == O.FX_ONE
        est = f * O.FX_INV
        big = exact > 1 / 107
        bad += int((np.abs(est - exact)[big] > EPS * exact[big]).sum())
        parts = sum(O.b_fora(g, s, ALPHA, EPS, SEED, restart_at_source=True, shard=i, nshards=4)[0] for i in range(4))
        assert (parts == f).all()
        fa, _ = O.a_fora(g, s, ALPHA, EPS, SEED, restart_at_source=True)
        assert abs(fa.sum() - 1) < 1e-9
    assert bad == 0


def test_topk_semantics(O, g):
    v = np.array([0.5, 0.1, 0.0, 0.3, 0.3, 0.2, 0.0])
    assert O.kth_ppr(v, 2) == 0.3 and O.kth_ppr(v, 8) is None and O.kth_ppr(v, 1) == 0.5   # J/Algo_Util.java:44,52
    assert list(O.topk_select(v, 2)) == [0, 3, 4]               # ties kept: more than k results
    assert list(O.topk_select(v, 6)) == [0, 3, 4, 5, 1]         # fewer positives than k: all of them
    exact = O.power_method(g, 0, ALPHA)
    ta, st = O.a_fora_topk(g, 0, ALPHA, EPS, K, SEED)
    tb, stb = O.b_fora_topk(g, 0, ALPHA, EPS, K, SEED)
    for t in (ta, tb * O.FX_INV):
        m = O.topk_metrics(exact, K, O.topk_select(t, K))
        assert m["precision"] >= 0.8 and m["ndcg"] >= 0.97
    assert 1 <= st["rounds"] <= 3 and 1 <= stb["rounds"] <= 3   # <= 3 rounds on GOT, k = 10
    td, _ = O.b_fora_topk(g, 71, ALPHA, EPS, K, SEED)           # dangling source
    assert int(td[71]) == O.FX_ONE and np.count_nonzero(td) == 1
    assert O.max_err(exact, exact) == 0.0


def test_all_pairs_counts(O, g):
    # replay of the reference semantics on GOT
    for thr, pushes, edges, pairs in [(1e-3, 2667, 8658, 1195), (5e-7, 5553, 16530, 1775)]:
        ptr, tgt, val, st = O.a_base_all_pairs(g, ALPHA, thr)
        assert (st["pushes"], st["edges"], len(tgt)) == (pushes, edges, pairs)
        assert (val >= thr).all() and ptr[-1] == pairs
    ptr, tgt, val, _ = O.a_base_all_pairs(g, ALPHA, 1e-3, k=3)
    for v in range(107):
        row = val[ptr[v]:ptr[v + 1]]
        assert (np.diff(row) <= 0).all()
    pa, ta, va, _ = O.a_base_all_pairs(g, ALPHA, 5e-5)
    pb, tb, vb, _ = O.b_base_all_pairs(g, ALPHA, 5e-5)
    A = {(v, int(t)): x for v in range(107) for t, x in zip(ta[pa[v]:pa[v + 1]], va[pa[v]:pa[v + 1]])}
    B = {(v, int(t)): x * O.FX_INV for v in range(107) for t, x in zip(tb[pb[v]:pb[v + 1]], vb[pb[v]:pb[v + 1]])}
    common = set(A) & set(B)
    assert len(common) >= 0.9 * max(len(A), len(B))
    assert max(abs(A[k] - B[k]) for k in common) < 5e-4         # both are within ~rmax of pi(v, t)
    parts = [O.b_base_all_pairs(g, ALPHA, 5e-5, lo, hi)[:3] for lo, hi in ((0, 40), (40, 107))]
    assert sum(len(p[1]) for p in parts) == len(tb)


def test_golden_fixtures_match_oracle(O, g):
    G = np.load(GOLD)
    assert (G["row_ptr"] == g.rp).all() and (G["col_idx"] == g.ci).all()
    for s in G["sources"]:
        s = int(s)
        assert np.array_equal(G["pm_%d" % s], O.power_method(g, s, ALPHA, 100))
        res, r, st = O.b_forward_push(g, s, ALPHA, 1e-6)
        assert np.array_equal(G["push_reserve_%d" % s], res) and np.array_equal(G["push_residue_%d" % s], r)
        assert np.array_equal(G["fora_%d" % s], O.b_fora(g, s, ALPHA, EPS, SEED)[0])
        assert np.array_equal(G["mc_%d" % s], O.b_monte_carlo(g, s, ALPHA, EPS, SEED)[0])
        t, _ = O.b_fora_topk(g, s, ALPHA, EPS, K, SEED)
        assert np.array_equal(G["topk_est_%d" % s], t) and np.array_equal(G["topk_ids_%d" % s], O.topk_select(t, K))
        assert np.array_equal(G["bwd_%d" % s], O.b_backward_search(g, s, ALPHA, 1e-3)[0])
    for thr in (1e-3, 5e-5):
        ptr, tgt, val, _ = O.b_base_all_pairs(g, ALPHA, thr)
        assert np.array_equal(G["ap_ptr_%g" % thr], ptr) and np.array_equal(G["ap_tgt_%g" % thr], tgt) and np.array_equal(G["ap_val_%g" % thr], val)
