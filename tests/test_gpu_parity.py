"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the golden fixtures.

Integer / fixed-point / index results are compared bit for bit.  Floating-point outputs are the exact
images of those integers (value * 2^-62), so they are compared with == as well; the only tolerance in
this file is the FORA (eps, delta, p_f) guarantee against the exact power iteration, written where it
is used.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "got_golden.npz")
ALPHA, EPS, SEED, K = 0.15, 0.5, 7, 10


@pytest.fixture(scope="module")
def G(pkg, ctx, got):
    g = pkg.Graph(ctx, got[0], got[1])
    yield g
    g.close()


@pytest.fixture(scope="module")
def OG(O, got):
    return O.OGraph(got[0], got[1])


@pytest.fixture(scope="module")
def R12(pkg, ctx, O, rmat12):
    g = pkg.Graph(ctx, *rmat12)
    yield g, O.OGraph(*rmat12)
    g.close()


@pytest.fixture(scope="module")
def R16(pkg, ctx, O, rmat16):
    g = pkg.Graph(ctx, *rmat16)
    yield g, O.OGraph(*rmat16)
    g.close()


def sources(og, k, seed=1234):
    rng = np.random.default_rng(seed)
    return [int(x) for x in rng.integers(0, og.n, size=k)]


# ------------------------------------------------------------------ graph export
def test_graph_export(G, OG, pkg, ctx):
    assert G.nodeCount() == 107 and G.relCount() == 352
    cp, ri = G.export_csc()
    src = np.repeat(np.arange(OG.n), np.diff(OG.rp))
    order = np.lexsort((src, OG.ci))
    assert np.array_equal(ri, src[order].astype(np.int32))
    assert np.array_equal(np.diff(cp), np.bincount(OG.ci, minlength=OG.n))
    with pytest.raises(pkg.PPRError) as e:
        pkg.Graph(ctx, np.array([0, 2, 1]), np.array([0, 1], np.int32))
    assert e.value.code == pkg._native.PPR_E_INVALID
    with pytest.raises(pkg.PPRError):
        pkg.Graph(ctx, np.array([0, 1]), np.array([5], np.int32))


# ------------------------------------------------------------------ forward push
def test_forward_push_got_all_sources(G, OG, O):
    for s in range(OG.n):
        reserve, residue, rsum, st = G.forward_push(s, ALPHA, 1e-6)
        raw_res, raw_r = G.last_raw()
        ores, orr, ost = O.b_forward_push(OG, s, ALPHA, 1e-6)
        assert np.array_equal(raw_res, ores) and np.array_equal(raw_r, orr), "source %d" % s
        assert (st["pushes"], st["edges"], st["iters"]) == (ost["pushes"], ost["edges"], ost["iters"])
        assert np.array_equal(reserve, ores * O.FX_INV) and np.array_equal(residue, orr * O.FX_INV)
        assert rsum == float(int(orr.sum())) * O.FX_INV
        assert int(raw_res.sum()) + int(raw_r.sum()) == O.FX_ONE


def test_forward_push_reference_grid(G, OG, O):
    for rmax in (1e-4, 1e-5, 1e-6, 1e-7, 1e-8):               # J/Gen_Util.java:464
        for s in (0, 63, 71):
            G.forward_push(s, ALPHA, rmax)
            raw_res, raw_r = G.last_raw()
            ores, orr, _ = O.b_forward_push(OG, s, ALPHA, rmax)
            assert np.array_equal(raw_res, ores) and np.array_equal(raw_r, orr)


def test_forward_push_golden(G):
    gold = np.load(GOLD)
    for s in gold["sources"]:
        G.forward_push(int(s), ALPHA, 1e-6)
        raw_res, raw_r = G.last_raw()
        assert np.array_equal(raw_res, gold["push_reserve_%d" % s]) and np.array_equal(raw_r, gold["push_residue_%d" % s])


def test_forward_push_frontier_sequence(G, OG, O, R12):
    for g, og, srcs, rmax in ((G, OG, (0, 63, 5), 1e-6), (R12[0], R12[1], sources(R12[1], 3), 1e-7)):
        for s in srcs:
            fr = g.push_trace(s, ALPHA, rmax)
            raw_res, raw_r = g.last_raw()
            ores, orr, ost = O.b_forward_push(og, s, ALPHA, rmax, trace=True)
            assert len(fr) == len(ost["frontiers"])
            for a, b in zip(fr, ost["frontiers"]):
                assert np.array_equal(a, b)
            assert np.array_equal(raw_res, ores) and np.array_equal(raw_r, orr)


def test_forward_push_rmat(R12, R16, O):
    for (g, og), rmaxs, nsrc in ((R12, (1e-5, 1e-7, 1e-9), 6), (R16, (1e-6, 1e-8), 4)):
        deg = np.diff(og.rp)
        for s in sources(og, nsrc) + [int(deg.argmax()), int(np.flatnonzero(deg == 0)[0])]:
            for rmax in rmaxs:
                _, _, _, st = g.forward_push(s, ALPHA, rmax, want_residue=False)
                raw_res, raw_r = g.last_raw()
                ores, orr, ost = O.b_forward_push(og, s, ALPHA, rmax)
                assert np.array_equal(raw_res, ores) and np.array_equal(raw_r, orr), (s, rmax)
                assert (st["pushes"], st["edges"], st["iters"]) == (ost["pushes"], ost["edges"], ost["iters"])


# ------------------------------------------------------------------ walks / Monte-Carlo
def test_walk_terminals(G, OG, O, R12):
    gold = np.load(GOLD)
    assert np.array_equal(G.walk_terminals(63, 63, 0, 256, SEED, 0, ALPHA, True), gold["walk_terminals_63"])
    for g, og in ((G, OG), R12):
        deg = np.diff(og.rp)
        starts = sources(og, 4) + [int(np.flatnonzero(deg == 0)[0])]
        for s in starts:
            for nzh in (False, True):
                for restart in (s, 0):
                    got = g.walk_terminals(s, restart, 5, 300, SEED, 2, ALPHA, nzh)
                    exp = [O.walk(og, s, restart, 5 + j, SEED, 2, ALPHA, nzh)[0] for j in range(300)]
                    assert list(got) == exp


def test_monte_carlo(G, OG, O, R12):
    for g, og, srcs in ((G, OG, (0, 63, 71)), (R12[0], R12[1], sources(R12[1], 2))):
        for s in srcs:
            ppr, st = g.monte_carlo(s, ALPHA, EPS, SEED)
            cnt, _ = g.last_raw(False)
            ocnt, ost = O.b_monte_carlo(og, s, ALPHA, EPS, SEED)
            assert np.array_equal(cnt, ocnt)
            assert st["walks_total"] == ost["walks"] and st["steps"] == ost["steps"] and st["omega"] == ost["omega"]
            assert np.array_equal(ppr, ocnt / ost["omega"])
            oa, _ = O.a_monte_carlo(og, s, ALPHA, EPS, SEED)
            assert np.array_equal(ppr, oa)
    gold = np.load(GOLD)
    G.monte_carlo(63, ALPHA, EPS, SEED)
    assert np.array_equal(G.last_raw(False)[0], gold["mc_63"])


def test_monte_carlo_shards(G, OG, O):
    full, _ = O.b_monte_carlo(OG, 0, ALPHA, 0.3, SEED)
    for world in (2, 3, 8):
        acc = np.zeros(OG.n, np.uint64)
        for r in range(world):
            G.monte_carlo_shard(0, ALPHA, 0.3, SEED, r, world)
            part = G.last_raw(False)[0]
            assert np.array_equal(part, O.b_monte_carlo(OG, 0, ALPHA, 0.3, SEED, r, world)[0])
            acc += part
        assert np.array_equal(acc, full)


# ------------------------------------------------------------------ FORA
def check_fora(g, og, O, s, eps, flags=0, walk_cost=6.0):
    ppr, st = g.fora(s, ALPHA, eps, SEED, flags)
    raw, r = g.last_raw()
    oraw, ost = O.b_fora(og, s, ALPHA, eps, SEED, walk_cost=walk_cost, restart_at_source=bool(flags & 1))
    assert np.array_equal(raw, oraw), (s, eps)
    for k in ("pushes", "edges", "iters", "rounds", "walk_sources", "steps"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    assert st["walks_total"] == ost["walks"] and st["walks"] == ost["walks"]
    assert st["rmax_fx"] == ost["rmax_fx_final"] and st["rsum_fx"] == ost["rsum_fx"]
    assert np.array_equal(ppr, oraw * O.FX_INV)
    return ppr, raw, r, st


def test_fora_got(G, OG, O):
    for s in range(0, OG.n, 3):
        ppr, raw, r, st = check_fora(G, OG, O, s, EPS)
        assert int(raw.sum()) == O.FX_ONE and not r.any()      # every unit of mass lands in the estimate
    for eps in (10.0, 5.0, 0.5, 0.1, 0.05):                    # J/Gen_Util.java:463
        check_fora(G, OG, O, 63, eps)
    gold = np.load(GOLD)
    for s in gold["sources"]:
        G.fora(int(s), ALPHA, EPS, SEED)
        assert np.array_equal(G.last_raw(False)[0], gold["fora_%d" % s])


def test_fora_rmat(R12, R16, O):
    for (g, og), nsrc in ((R12, 5), (R16, 2)):
        deg = np.diff(og.rp)
        for s in sources(og, nsrc) + [int(np.flatnonzero(deg == 0)[0])]:
            for flags in (0, 1):
                ppr, raw, r, st = check_fora(g, og, O, s, EPS, flags)
                assert int(raw.sum()) == O.FX_ONE


def test_fora_balance_knobs(pkg, ctx, G, OG, O):
    ctx.set_balance(4.0, 3)
    try:
        ppr, st = G.fora(63, ALPHA, EPS, SEED)
        oraw, ost = O.b_fora(OG, 63, ALPHA, EPS, SEED, walk_cost=4.0, max_rounds=3)
        assert np.array_equal(G.last_raw(False)[0], oraw) and st["rounds"] == ost["rounds"] <= 3
    finally:
        ctx.set_balance(-1.0, 64)


def test_frontier_modes_give_identical_bits(ctx, G, OG, R12, O):
    # LIST (returning atomics + crossing test) and SCAN (plain reductions + dense sweep) frontier discovery
    try:
        for thr in (0, 1 << 62, 64, -1):
            ctx.set_tuning(thr)
            for g, og, srcs in ((G, OG, (0, 63, 100)), (R12[0], R12[1], sources(R12[1], 2))):
                for s in srcs:
                    check_fora(g, og, O, s, EPS)
                    g.forward_push(s, ALPHA, 1e-7, want_residue=False)
                    raw_res, raw_r = g.last_raw()
                    ores, orr, ost = O.b_forward_push(og, s, ALPHA, 1e-7)
                    assert np.array_equal(raw_res, ores) and np.array_equal(raw_r, orr), (thr, s)
                    ids, _, st = g.fora_topk(s, ALPHA, EPS, K, SEED)
                    assert np.array_equal(g.last_raw(False)[0], O.b_fora_topk(og, s, ALPHA, EPS, K, SEED)[0])
    finally:
        ctx.set_tuning(-1)


def test_pull_direction_gives_identical_bits(pkg, ctx, G, OG, R12, R16, O, got, rmat12):
    # Fat iterations run in the pull direction (pull.cuh): every in-edge read once, amounts gathered from a dense array, no
    # atomics, phase A of the next iteration fused in.  Forced after EVERY dense iteration here (enter = exit = 0, sweep
    # threshold 0 so that every iteration is dense), then with the default switch points, then off: always the oracle's bits,
    # work counters and iteration counts.
    try:
        for enter, exit_, scan in ((0.0, 0.0, 0), (0.0, 0.0, -1), (0.05, 0.02, -1), (0.30, 0.20, -1), (-1.0, 0.0, -1)):
            ctx.set_pull(enter, exit_)
            ctx.set_tuning(scan)
            for g, og, srcs in ((G, OG, (0, 63, 100)), (R12[0], R12[1], sources(R12[1], 3)), (R16[0], R16[1], sources(R16[1], 1))):
                for s in srcs:
                    check_fora(g, og, O, s, EPS)
                    _, _, _, st = g.forward_push(s, ALPHA, 1e-7, want_residue=False)
                    raw_res, raw_r = g.last_raw()
                    ores, orr, ost = O.b_forward_push(og, s, ALPHA, 1e-7)
                    assert np.array_equal(raw_res, ores) and np.array_equal(raw_r, orr), (enter, scan, s)
                    assert (st["pushes"], st["edges"], st["iters"]) == (ost["pushes"], ost["edges"], ost["iters"])
                    g.fora_topk(s, ALPHA, EPS, K, SEED)
                    assert np.array_equal(g.last_raw(False)[0], O.b_fora_topk(og, s, ALPHA, EPS, K, SEED)[0])
        # long rows (more than 2048 in-edges, summed chunk-wise by several warps) and dangling mass returning to the source:
        # a star-like graph: vertex 0 receives an edge from everybody, half of the vertices have no other out-edge target
        ctx.set_pull(0.0, 0.0)
        ctx.set_tuning(0)
        n = 6000
        rng = np.random.default_rng(11)
        rows = [np.unique(np.concatenate(([0], rng.integers(0, n, rng.integers(0, 6))))) if v % 3 else np.zeros(0, np.int64) for v in range(n)]
        rows[0] = np.unique(rng.integers(1, n, 3000))
        rp = np.concatenate(([0], np.cumsum([len(r) for r in rows]))).astype(np.int64)
        ci = np.concatenate(rows).astype(np.int32)
        g, og = pkg.Graph(ctx, rp, ci), O.OGraph(rp, ci)
        try:
            for s in (0, 1, 2, 4):
                check_fora(g, og, O, s, EPS)
                g.forward_push(s, ALPHA, 1e-9, want_residue=False)
                raw_res, raw_r = g.last_raw()
                ores, orr, _ = O.b_forward_push(og, s, ALPHA, 1e-9)
                assert np.array_equal(raw_res, ores) and np.array_equal(raw_r, orr), s
        finally:
            g.close()
    finally:
        ctx.set_pull(-1.0, 0.0)     # the library default: off
        ctx.set_tuning(-1)


def test_single_cta_iterations_give_identical_bits(G, OG, R12, R16, O, monkeypatch):
    # small frontiers are iterated by one CTA (push_mini); PPR_NO_MINI=1 keeps every iteration grid-wide.
    # Both must give the oracle's bits (the rest of this file runs with the default, i.e. mini enabled).
    for no_mini in ("1", None):
        if no_mini:
            monkeypatch.setenv("PPR_NO_MINI", no_mini)
        else:
            monkeypatch.delenv("PPR_NO_MINI", raising=False)
        for g, og, srcs in ((G, OG, (0, 63, 100)), (R12[0], R12[1], sources(R12[1], 2)), (R16[0], R16[1], sources(R16[1], 1))):
            for s in srcs:
                check_fora(g, og, O, s, EPS)
                g.forward_push(s, ALPHA, 1e-7, want_residue=False)
                raw_res, raw_r = g.last_raw()
                ores, orr, ost = O.b_forward_push(og, s, ALPHA, 1e-7)
                assert np.array_equal(raw_res, ores) and np.array_equal(raw_r, orr), (no_mini, s)
        ids, _, _ = G.fora_topk(63, ALPHA, EPS, K, SEED)
        assert np.array_equal(G.last_raw(False)[0], O.b_fora_topk(OG, 63, ALPHA, EPS, K, SEED)[0])


def test_hub_slots_give_identical_bits(pkg, ctx, got, rmat12, O):
    # every vertex with an in-edge gets a shared-memory hub slot (hot_min_indeg = 1) and every iteration runs in
    # SCAN mode (threshold 0): the staged adds + per-CTA flush must reproduce the oracle's bits
    try:
        ctx.set_tuning(0, 1)
        for rp, ci in (got[:2], rmat12):
            g = pkg.Graph(ctx, rp, ci)
            og = O.OGraph(rp, ci)
            deg = np.diff(og.rp)
            for s in [int(x) for x in np.flatnonzero(deg > 0)[[0, 5, 11]]]:
                check_fora(g, og, O, s, EPS)
                g.forward_push(s, ALPHA, 1e-8, want_residue=False)
                raw_res, raw_r = g.last_raw()
                ores, orr, _ = O.b_forward_push(og, s, ALPHA, 1e-8)
                assert np.array_equal(raw_res, ores) and np.array_equal(raw_r, orr)
            ctx.set_tuning(1 << 62, -1)   # same graph (hub-encoded edges), LIST mode: decode path
            check_fora(g, og, O, int(np.flatnonzero(deg > 0)[3]), EPS)
            ctx.set_tuning(0, -1)
            g.close()
    finally:
        ctx.set_tuning(-1, 64)


def test_fora_guarantee(G, OG, R12, O):
    # FORA's guarantee (Dissertation p.14): |est - pi| <= eps * pi for every pi > delta = 1/n, w.p. 1 - 1/n.
    # Tolerance: exactly eps = 0.5 relative; the exact answer is the power iteration run to 1e-17.
    for g, og, srcs in ((G, OG, (0, 63, 5, 100)), (R12[0], R12[1], sources(R12[1], 3))):
        for s in srcs:
            ppr, _ = g.fora(s, ALPHA, EPS, SEED, 1)
            exact = O.power_method(og, s, ALPHA, 3000, 1e-17)
            big = exact > 1.0 / og.n
            assert (np.abs(ppr - exact)[big] <= EPS * exact[big]).all()


def test_fora_walk_shards(G, OG, R12, O):
    for g, og, s in ((G, OG, 63), (R12[0], R12[1], sources(R12[1], 1)[0])):
        g.fora(s, ALPHA, EPS, SEED)
        full = g.last_raw(False)[0]
        for world in (2, 4, 8):
            acc = np.zeros(og.n, np.uint64)
            for r in range(world):
                g.fora_shard(s, ALPHA, EPS, SEED, r, world)
                part = g.last_raw(False)[0]
                assert np.array_equal(part, O.b_fora(og, s, ALPHA, EPS, SEED, shard=r, nshards=world)[0])
                acc += part
            assert np.array_equal(acc, full)


def test_fora_topk_walk_shards(pkg, ctx, got, rmat12, O):
    # ppr_fora_topk_shard: every delta-round's walks split over shards, one integer sum per round through the caller's
    # reduce hook.  The shards run as threads here (one context + graph each, all on this GPU) and the hook sums the
    # estimate vectors through the host; ids, values, rounds and the raw estimate vector must equal the unsharded run.
    import threading
    import torch
    for (rp, ci), srcs, k in ((got[:2], (0, 63), K), (rmat12, sources(O.OGraph(*rmat12), 2), 50)):
        og = O.OGraph(rp, ci)
        g0 = pkg.Graph(ctx, rp, ci)
        for s in srcs:
            ids0, vals0, st0 = g0.fora_topk(s, ALPHA, EPS, k, SEED)
            raw0 = g0.last_raw(False)[0].copy()
            assert np.array_equal(raw0, O.b_fora_topk(og, s, ALPHA, EPS, k, SEED)[0])
            ids1, vals1, st1 = g0.fora_topk_shard(s, ALPHA, EPS, k, SEED, 0, 1)
            assert np.array_equal(ids0, ids1) and np.array_equal(vals0, vals1) and st0["rounds"] == st1["rounds"]
            for world in (2, 3):
                ctxs = [pkg.Context(0) for _ in range(world)]
                graphs = [pkg.Graph(c, rp, ci) for c in ctxs]
                bar = threading.Barrier(world)
                parts, out, calls = [None] * world, [None] * world, [0] * world

                def run(r):
                    def hook(ptr, n):
                        t = torch.as_tensor(pkg.dist._DevVec(ptr, n), device="cuda:0")
                        parts[r] = t.cpu()
                        bar.wait(60)
                        total = parts[0].clone()
                        for p_ in parts[1:]:
                            total += p_
                        bar.wait(60)
                        t.copy_(total)
                        torch.cuda.synchronize()
                        calls[r] += 1
                    try:
                        out[r] = graphs[r].fora_topk_shard(s, ALPHA, EPS, k, SEED, r, world, hook)
                    except BaseException as e:
                        out[r] = e
                        bar.abort()

                th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
                for t_ in th:
                    t_.start()
                for t_ in th:
                    t_.join(120)
                for r in range(world):
                    assert not isinstance(out[r], BaseException), out[r]
                    ids, vals, st = out[r]
                    assert np.array_equal(ids, ids0) and np.array_equal(vals, vals0), (world, r, s)
                    assert st["rounds"] == st0["rounds"] == calls[r]
                    assert np.array_equal(graphs[r].last_raw(False)[0], raw0)
                assert sum(o[2]["walks"] for o in out) == st0["walks"]
                for g_ in graphs:
                    g_.close()
                for c in ctxs:
                    c.close()
        g0.close()
    with pytest.raises(pkg.PPRError):      # more than one shard needs the hook
        G_ = pkg.Graph(ctx, *got[:2])
        try:
            G_.fora_topk_shard(0, ALPHA, EPS, K, SEED, 0, 2, None)
        finally:
            G_.close()


def test_fora_topk_output_grows_on_capacity_error(G, OG, O):
    # the binding offers k + ties slots first; a result that does not fit reports the needed size (PPR_E_CAPACITY)
    # and is fetched by a second call -- forced here with a first-call capacity of 3
    ids0, vals0, st0 = G.fora_topk(63, ALPHA, EPS, K, SEED)
    ids1, vals1, st1 = G.fora_topk(63, ALPHA, EPS, K, SEED, cap=3)
    assert len(ids0) >= K and np.array_equal(ids0, ids1) and np.array_equal(vals0, vals1)
    assert np.array_equal(G.last_raw(False)[0], O.b_fora_topk(OG, 63, ALPHA, EPS, K, SEED)[0])


def test_fora_batch_equals_single(R12, O):
    g, og = R12
    srcs = sources(og, 5)
    out, stats = g.fora_batch(srcs, ALPHA, EPS, 100)
    for q, s in enumerate(srcs):
        one, _ = g.fora(s, ALPHA, EPS, 100 + q)
        assert np.array_equal(out[q], one)
        assert np.array_equal(out[q], O.b_fora(og, s, ALPHA, EPS, 100 + q)[0] * O.FX_INV)
        assert stats[q]["ms_push"] > 0


# ------------------------------------------------------------------ top-k
def test_topk_select(ctx, O):
    rng = np.random.default_rng(5)
    v = rng.random(5000)
    v[rng.integers(0, 5000, 2500)] = 0.0
    v[10] = v[20] = v[30] = 0.77                               # ties at the cut
    for k in (1, 3, 10, 500, 2400, 4999):
        ids, vals, kth = ctx.topk(v, k)
        exp = O.topk_select(v, k)
        assert np.array_equal(ids, exp) and np.array_equal(vals, v[exp])
        okth = O.kth_ppr(v[v > 0], k)
        assert kth == (okth if okth is not None else 0.0)
    kcut = int((v > 0.77).sum()) + 1
    ids, _, kth = ctx.topk(v, kcut)
    assert kth == 0.77 and len(ids) == kcut + 2                # "might include more than k nodes"


def test_fora_topk(G, OG, R12, O):
    gold = np.load(GOLD)
    for g, og, srcs, k in ((G, OG, list(range(0, 107, 7)) + [71], K), (R12[0], R12[1], sources(R12[1], 3), 50)):
        for s in srcs:
            ids, vals, st = g.fora_topk(s, ALPHA, EPS, k, SEED)
            est = g.last_raw(False)[0]
            oest, ost = O.b_fora_topk(og, s, ALPHA, EPS, k, SEED)
            assert np.array_equal(est, oest), s
            oids = O.topk_select(oest, k)
            assert np.array_equal(ids, oids) and np.array_equal(vals, oest[oids] * O.FX_INV)
            assert st["rounds"] == ost["rounds"] and st["pushes"] == ost["pushes"] and st["edges"] == ost["edges"]
    for s in gold["sources"]:
        ids, _, _ = G.fora_topk(int(s), ALPHA, EPS, K, SEED)
        assert np.array_equal(ids, gold["topk_ids_%d" % s])


def test_fora_topk_overlap_with_reference_order(G, OG, O):
    # top-k set overlap against the reference-order (mode A) replay and the exact answer, as the harness reports it
    prec = []
    for s in range(0, 107, 5):
        ids, _, _ = G.fora_topk(s, ALPHA, EPS, K, SEED)
        exact = O.power_method(OG, s, ALPHA)
        prec.append(O.topk_metrics(exact, K, ids)["precision"])
        a, _ = O.a_fora_topk(OG, s, ALPHA, EPS, K, SEED)
        a_ids = set(O.topk_select(a, K).tolist())
        assert len(a_ids & set(ids.tolist())) >= 0.7 * min(len(a_ids), len(ids))
    assert np.mean(prec) >= 0.9


# ------------------------------------------------------------------ backward search / all pairs
def test_backward_search(G, OG, R12, O):
    for g, og, ts, rmax in ((G, OG, range(0, 107, 4), 1e-3), (G, OG, (5, 63), 5e-7), (R12[0], R12[1], sources(R12[1], 4), 1e-4)):
        for t in ts:
            res, st = g.backward_search(int(t), ALPHA, rmax)
            raw, r = g.last_raw()
            ores, orr, ost = O.b_backward_search(og, int(t), ALPHA, rmax)
            assert np.array_equal(raw, ores) and np.array_equal(r, orr), t
            assert (st["pushes"], st["edges"], st["iters"]) == (ost["pushes"], ost["edges"], ost["iters"])
            assert np.array_equal(res, ores * O.FX_INV)


def test_all_pairs(G, OG, R12, O):
    gold = np.load(GOLD)
    for thr in (1e-3, 5e-5):
        ptr, tgt, val, st = G.backward_all_pairs(ALPHA, thr, raw=True)
        assert np.array_equal(ptr, gold["ap_ptr_%g" % thr]) and np.array_equal(tgt, gold["ap_tgt_%g" % thr])
        assert np.array_equal(val, gold["ap_val_%g" % thr])
    g, og = R12
    optr, otgt, oval, ost = O.b_base_all_pairs(og, ALPHA, 1e-3)
    ptr, tgt, val, st = g.backward_all_pairs(ALPHA, 1e-3, raw=True)
    assert np.array_equal(ptr, optr) and np.array_equal(tgt, otgt) and np.array_equal(val, oval)
    assert (st["pushes"], st["edges"]) == (ost["pushes"], ost["edges"])
    # block partition of the targets (multi-GPU layout) merges to the same index
    from importlib import import_module
    D = import_module("personalized-pagerank-algorithms-on-neo4j_b200").dist
    parts = [g.backward_all_pairs(ALPHA, 1e-3, -1, lo, hi, raw=True)[:3] for lo, hi in D.target_blocks(og.n, 4)]
    mptr, mtgt, mval = D.merge_pairs(parts, og.n)
    assert np.array_equal(mptr, optr) and np.array_equal(mtgt, otgt) and np.array_equal(mval, oval)


def test_all_pairs_launch_split(R16, O, ctx):
    # n = 65536, 12000 targets: a probe launch over an evenly spaced sample (1184 targets, one CTA per search), the highest in-degree
    # targets as cluster teams, the rest heaviest first -- the index must not depend on how targets are split over launches or on
    # the pair buffer being regrown (forced batches below)
    g, og = R16
    hi = 12000
    optr, otgt, oval, ost = O.b_base_all_pairs(og, ALPHA, 1e-3, 0, hi)
    ptr, tgt, val, st = g.backward_all_pairs(ALPHA, 1e-3, -1, 0, hi, raw=True)
    assert np.array_equal(ptr, optr) and np.array_equal(tgt, otgt) and np.array_equal(val, oval)
    assert (st["pushes"], st["edges"]) == (ost["pushes"], ost["edges"])
    assert st["launches"] <= 3, st["launches"]         # probe, the heaviest targets as cluster teams, the rest
    try:
        ctx.set_bwd_batch(700)      # 18 launches of at most 700 targets
        ptr, tgt, val, st = g.backward_all_pairs(ALPHA, 1e-3, -1, 0, hi, raw=True)
        assert np.array_equal(ptr, optr) and np.array_equal(tgt, otgt) and np.array_equal(val, oval)
        assert (st["pushes"], st["edges"]) == (ost["pushes"], ost["edges"]) and st["launches"] >= 18
    finally:
        ctx.set_bwd_batch(0)


def test_all_pairs_multi_tile_frontiers(R12, O):
    # threshold 1e-6 on R-MAT 2^12: frontiers of thousands of vertices (several tiles per CTA, staged through the slot's
    # global rows), residues re-touched after being pushed (duplicates in the touched list)
    g, og = R12
    optr, otgt, oval, ost = O.b_base_all_pairs(og, ALPHA, 1e-6, 0, 300)
    ptr, tgt, val, st = g.backward_all_pairs(ALPHA, 1e-6, -1, 0, 300, raw=True)
    assert np.array_equal(ptr, optr) and np.array_equal(tgt, otgt) and np.array_equal(val, oval)
    assert (st["pushes"], st["edges"]) == (ost["pushes"], ost["edges"])


def test_all_pairs_cluster_teams(R12, R16, O, ctx, monkeypatch):
    # searches too large for one SM run as thread-block clusters (teams of 2 / 4 / 8 CTAs splitting a tile's edges, cursors in
    # distributed shared memory): forced for every launch here, on small frontiers (1e-3) and multi-tile ones (1e-6); then the
    # automatic plan (probe, heaviest targets as teams, rest one CTA each) on R-MAT 2^16
    g, og = R12
    ref = {thr: O.b_base_all_pairs(og, ALPHA, thr, 0, 300) for thr in (1e-3, 1e-6)}
    monkeypatch.setenv("PPR_BWD_FORCE_TEAM", "1")
    for team in ("2", "4", "8"):
        monkeypatch.setenv("PPR_BWD_TEAM", team)
        for thr in (1e-3, 1e-6):
            optr, otgt, oval, ost = ref[thr]
            ptr, tgt, val, st = g.backward_all_pairs(ALPHA, thr, -1, 0, 300, raw=True)
            assert np.array_equal(ptr, optr) and np.array_equal(tgt, otgt) and np.array_equal(val, oval), (team, thr)
            assert (st["pushes"], st["edges"]) == (ost["pushes"], ost["edges"])
    monkeypatch.delenv("PPR_BWD_FORCE_TEAM")
    monkeypatch.delenv("PPR_BWD_TEAM")
    g, og = R16
    optr, otgt, oval, ost = O.b_base_all_pairs(og, ALPHA, 1e-3, 0, 9000)
    ptr, tgt, val, st = g.backward_all_pairs(ALPHA, 1e-3, -1, 0, 9000, raw=True)
    assert np.array_equal(ptr, optr) and np.array_equal(tgt, otgt) and np.array_equal(val, oval)
    assert (st["pushes"], st["edges"]) == (ost["pushes"], ost["edges"]) and st["launches"] == 3   # probe, teams, rest


def test_all_pairs_dense_iterations(R12, G, OG, O, monkeypatch):
    # iterations whose frontier owns many in-edges add with plain reductions and find the next frontier by sweeping the residue row
    # (PPR_BWD_SCAN = the switch point in edges): every iteration that way (0), none (huge), a mix (256), as single CTAs and as teams
    g, og = R12
    refs = {thr: O.b_base_all_pairs(og, ALPHA, thr, 0, 300) for thr in (1e-3, 1e-6)}
    gref = O.b_base_all_pairs(OG, ALPHA, 1e-3)
    for scan in ("0", "256", "1000000000000"):
        monkeypatch.setenv("PPR_BWD_SCAN", scan)
        for team in (None, "2"):
            if team:
                monkeypatch.setenv("PPR_BWD_FORCE_TEAM", "1"); monkeypatch.setenv("PPR_BWD_TEAM", team)
            else:
                monkeypatch.delenv("PPR_BWD_FORCE_TEAM", raising=False); monkeypatch.delenv("PPR_BWD_TEAM", raising=False)
            for thr, (optr, otgt, oval, ost) in refs.items():
                ptr, tgt, val, st = g.backward_all_pairs(ALPHA, thr, -1, 0, 300, raw=True)
                assert np.array_equal(ptr, optr) and np.array_equal(tgt, otgt) and np.array_equal(val, oval), (scan, team, thr)
                assert (st["pushes"], st["edges"]) == (ost["pushes"], ost["edges"])
            ptr, tgt, val, st = G.backward_all_pairs(ALPHA, 1e-3, raw=True)       # GOT: dangling targets, tiny rows
            assert np.array_equal(ptr, gref[0]) and np.array_equal(tgt, gref[1]) and np.array_equal(val, gref[2]), (scan, team)


def test_backward_search_dense_graph(pkg, ctx, O):
    # n = 2048 with ~600 in-edges per vertex: one iteration of ONE target stages ceil(indeg / 256) chunk descriptors for
    # every vertex, far more than T * n -- the chunk buffer is sized from the in-degree histogram at export
    rng = np.random.default_rng(5)
    n, d = 2048, 600
    ci = np.sort(rng.integers(0, n, (n, d)), axis=1).astype(np.int32).ravel()
    rp = (np.arange(n + 1, dtype=np.int64) * d)
    g, og = pkg.Graph(ctx, rp, ci), O.OGraph(rp, ci)
    try:
        res, st = g.backward_search(7, ALPHA, 1e-7)
        raw, r = g.last_raw()
        ores, orr, ost = O.b_backward_search(og, 7, ALPHA, 1e-7)
        assert np.array_equal(raw, ores) and np.array_equal(r, orr)
        assert (st["pushes"], st["edges"], st["iters"]) == (ost["pushes"], ost["edges"], ost["iters"])
        optr, otgt, oval, ost = O.b_base_all_pairs(og, ALPHA, 1e-4, 0, 64)
        ptr, tgt, val, st = g.backward_all_pairs(ALPHA, 1e-4, -1, 0, 64, raw=True)
        assert np.array_equal(ptr, optr) and np.array_equal(tgt, otgt) and np.array_equal(val, oval)
    finally:
        g.close()


def test_all_pairs_topk_rows(G, OG, O):
    ptr, tgt, val, _ = G.backward_all_pairs(ALPHA, 1e-3, 3, raw=True)
    fptr, ftgt, fval, _ = O.b_base_all_pairs(OG, ALPHA, 1e-3)
    for v in range(OG.n):
        row_t, row_v = ftgt[fptr[v]:fptr[v + 1]], fval[fptr[v]:fptr[v + 1]]
        keep = O.topk_select(row_v, 3) if len(row_v) else np.zeros(0, np.int32)   # >= 3rd largest, descending, stable
        assert np.array_equal(tgt[ptr[v]:ptr[v + 1]], row_t[keep]) and np.array_equal(val[ptr[v]:ptr[v + 1]], row_v[keep])


# ------------------------------------------------------------------ power iteration
def test_power_iteration(G, OG, R12, O):
    # double arithmetic; identical summation order except for the dangling-mass term of the source,
    # tolerance 1e-13 absolute (values are <= 1)
    for g, og, srcs in ((G, OG, (0, 63, 71)), (R12[0], R12[1], sources(R12[1], 2))):
        for s in srcs:
            pm = g.power_iteration(s, ALPHA, 100)
            assert np.abs(pm - O.power_method(og, s, ALPHA, 100)).max() < 1e-13
    gold = np.load(GOLD)
    assert np.abs(G.power_iteration(0, ALPHA, 100) - gold["pm_0"]).max() < 1e-13


# ------------------------------------------------------------------ error behaviour
def test_error_codes(pkg, G):
    E = pkg._native
    for call in (lambda: G.fora(-1, ALPHA, EPS, 1), lambda: G.fora(107, ALPHA, EPS, 1), lambda: G.fora(0, 1.5, EPS, 1),
                 lambda: G.fora(0, ALPHA, 0.0, 1), lambda: G.forward_push(0, ALPHA, -1.0), lambda: G.fora_topk(0, ALPHA, EPS, 0, 1),
                 lambda: G.fora_shard(0, ALPHA, EPS, 1, 3, 2), lambda: G.backward_all_pairs(ALPHA, 1e-3, -1, 5, 2)):
        with pytest.raises(pkg.PPRError) as e:
            call()
        assert e.value.code == E.PPR_E_INVALID
    G.fora(0, ALPHA, EPS, 1)                                   # the context stays usable after errors


# ------------------------------------------------------------------ edge-case graphs
def _edge_graphs(pkg):
    G = pkg.graphio.csr_from_edges
    star_src = [0] * 1000 + list(range(1, 1001))            # hub with 1000 out-edges (4 chunks) and 1000 in-edges
    star_dst = list(range(1, 1001)) + [0] * 1000
    return {
        "self_loop": G(1, [0], [0]),
        "two_cycle": G(2, [0, 1], [1, 0]),
        "path_with_sink": G(5, [0, 1, 2, 3], [1, 2, 3, 4]),
        "duplicates_and_loops": G(4, [0, 0, 0, 1, 1, 2, 2, 3], [1, 1, 0, 2, 2, 0, 3, 3]),
        "star_1000": G(1001, star_src, star_dst),
        "hub_to_sinks": G(700, [0] * 699, list(range(1, 700))),   # one long row, every target dangling
    }


def test_edge_case_graphs(pkg, ctx, O):
    for name, (rp, ci) in _edge_graphs(pkg).items():
        g = pkg.Graph(ctx, rp, ci)
        og = O.OGraph(rp, ci)
        try:
            for s in sorted({0, og.n - 1, og.n // 2}):
                for rmax in (1e-3, 1e-9):
                    g.forward_push(s, ALPHA, rmax)
                    res, r = g.last_raw()
                    ores, orr, _ = O.b_forward_push(og, s, ALPHA, rmax)
                    assert np.array_equal(res, ores) and np.array_equal(r, orr), (name, s, rmax)
                    fr = g.push_trace(s, ALPHA, rmax)
                    _, _, ost = O.b_forward_push(og, s, ALPHA, rmax, trace=True)
                    assert len(fr) == len(ost["frontiers"]) and all(np.array_equal(a, b) for a, b in zip(fr, ost["frontiers"])), name
                for flags in (0, 1):
                    check_fora(g, og, O, s, EPS, flags)
                g.monte_carlo(s, ALPHA, 1.0, SEED)
                assert np.array_equal(g.last_raw(False)[0], O.b_monte_carlo(og, s, ALPHA, 1.0, SEED)[0]), name
                g.backward_search(s, ALPHA, 1e-4)
                bres, br = g.last_raw()
                obres, obr, _ = O.b_backward_search(og, s, ALPHA, 1e-4)
                assert np.array_equal(bres, obres) and np.array_equal(br, obr), name
                k = min(3, og.n)
                ids, vals, _ = g.fora_topk(s, ALPHA, EPS, k, SEED)
                oest, _ = O.b_fora_topk(og, s, ALPHA, EPS, k, SEED)
                assert np.array_equal(g.last_raw(False)[0], oest) and np.array_equal(ids, O.topk_select(oest, k)), name
                pm = g.power_iteration(s, ALPHA, 100)
                assert np.abs(pm - O.power_method(og, s, ALPHA, 100)).max() < 1e-13
            ptr, tgt, val, _ = g.backward_all_pairs(ALPHA, 1e-4, raw=True)
            optr, otgt, oval, _ = O.b_base_all_pairs(og, ALPHA, 1e-4)
            assert np.array_equal(ptr, optr) and np.array_equal(tgt, otgt) and np.array_equal(val, oval), name
        finally:
            g.close()


def test_graph_without_edges(pkg, ctx):
    g = pkg.Graph(ctx, np.zeros(4, np.int64), np.zeros(0, np.int32))
    try:
        reserve, residue, rsum, st = g.forward_push(1, ALPHA, 1e-6)       # dangling source: {s: 1}, rsum 0
        assert reserve[1] == 1.0 and reserve.sum() == 1.0 and rsum == 0.0 and st["pushes"] == 0
        ppr, _ = g.monte_carlo(2, ALPHA, 1.0, 1)
        assert np.count_nonzero(ppr) == 1 and ppr[2] > 0.99
        with pytest.raises(pkg.PPRError):                                 # FORA's rmax divides by m
            g.fora(0, ALPHA, EPS, 1)
    finally:
        g.close()
