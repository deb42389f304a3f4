"""GPU (-m gpu): the reference-facing host mirror (class names / harness protocol) and full-size properties."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ALPHA, EPS = 0.15, 0.5


@pytest.fixture(scope="module")
def G(pkg, ctx, got):
    g = pkg.Graph(ctx, got[0], got[1])
    yield g
    g.close()


def test_algorithm_classes_like_the_harness_uses_them(pkg, G, O, got):
    og = O.OGraph(got[0], got[1])
    ac = pkg.Algo_Conf(ALPHA, "name")
    n, m = G.nodeCount(), G.relCount()
    pm = ac.set_conf_power_method(None, G, "got.db")
    for algo, param in ((ac.set_conf_fora_whole_graph(n, m, None, G, "got.db"), 0.5), (ac.set_conf_fwdpush(n, m, None, G, "got.db"), 1e-6),
                        (ac.set_conf_mc(n, m, None, G, "got.db"), 0.5)):
        algo.computeWholeGraphPPR(63, param)
        pm.computeWholeGraphPPR(63, None)
        res = algo.getWholeGraphPPR()
        assert abs(sum(res.values()) - 1) < 1e-3 and all(v > 0 for v in res.values())
        err = pkg.compute_error(algo, pm, "MAX_ERR", -1)
        assert err == O.max_err(pm.reserve, res.dense) and err < 0.02
    tk = ac.set_conf_fora_topk(n, m, 10, None, G, "got.db")
    tk.computeTopKPPR(63, 10, 0.5)
    pm.computeTopKPPR(63, 10, None)
    exact = O.power_method(og, 63, ALPHA)
    assert np.array_equal(pm.getTopKNodeIds(10), O.topk_select(pm.reserve, 10))
    m_ = O.topk_metrics(pm.reserve, 10, tk.getTopKNodeIds(10))
    assert abs(pkg.compute_error(tk, pm, "PRECISION", 10) - m_["precision"]) < 1e-15
    assert abs(pkg.compute_error(tk, pm, "NDCG", 10) - m_["ndcg"]) < 1e-12
    assert np.abs(pm.reserve - exact).max() < 1e-13
    mc = ac.set_conf_mc(n, m, None, G, "got.db")
    assert mc.random_walk(71) == 71 and mc.random_walk_no_zero_hop(71) == 71   # dangling start (J/Monte_Carlo.java:70-72)
    assert 0 <= mc.random_walk_no_zero_hop(63) < n
    bs = pkg.Backward_Search(ALPHA, 1e-3, n, None, G)
    bs.backward_search_whole_graph(5)
    assert np.array_equal(bs.reserve, O.b_backward_search(og, 5, ALPHA, 1e-3)[0] * O.FX_INV)


def test_print_methods_on_device_results(pkg, G, O, got, capsys):
    """printWholeGraphResult / printTopKResult of the reference classes (J/Forward_Push.java:268-281, :366-391,
    J/Fora_Topk.java:60-79, J/Power_Method.java:103-116) over results that came from the device."""
    names = got[2]
    og = O.OGraph(got[0], got[1])
    ac = pkg.Algo_Conf(ALPHA, "name")
    s = names.index("Tyrion")
    fp = ac.set_conf_fwdpush(107, 352, names, G, "got.db")
    fp.computeTopKPPR(s, 5, 1e-6)
    capsys.readouterr()
    fp.printTopKResult(5)
    lines = capsys.readouterr().out.splitlines()
    res, _, _ = O.b_forward_push(og, s, ALPHA, 1e-6)
    top = np.argsort(-(res * O.FX_INV), kind="stable")[:5]
    assert lines[:2] == ["", "Forward-Push-Top5 PPR:"] and len(lines) == 7
    assert lines[2:] == ["@%s\t%s" % (names[v], pkg.algo._java_double(res[v] * O.FX_INV)) for v in top]
    assert lines[2].startswith("@Tyrion\t0.39")
    fp.printWholeGraphResult()
    lines = capsys.readouterr().out.splitlines()
    assert lines[0] == "Forward-Push Reserve(pi) & Residue(r):" and len(lines) == 1 + np.count_nonzero(res)
    assert all(len(ln.split("\t")) == 3 and ln[0] == "@" for ln in lines[1:])
    tk = ac.set_conf_fora_topk(107, 352, 10, names, G, "got.db")
    tk.computeTopKPPR(s, 10, 0.5)
    tk.printTopKResult(10)
    lines = capsys.readouterr().out.splitlines()
    assert lines[1] == "Fora-Top10 PPR:" and lines[2].startswith("@Tyrion\t") and 3 <= len(lines) <= 12
    vals = [float(ln.split("\t")[1]) for ln in lines[2:]]
    assert vals == sorted(vals, reverse=True)
    pm = ac.set_conf_power_method(names, G, "got.db")
    pm.computeTopKPPR(s, 3, None)
    pm.printTopKResult(3)
    pm.printWholeGraphResult()
    lines = capsys.readouterr().out.splitlines()
    assert lines[1] == "Exact-Top3 PPR:" and lines[2].startswith("@Tyrion\t0.3916") and lines[5] == "Exact-PPR:"


def test_base_whole_graph_store(pkg, G, O, got, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    og = O.OGraph(got[0], got[1])
    base = pkg.Algo_Conf(ALPHA).set_conf_base_whole_graph(107, 352, None, G, "got.db")
    base.preprocessing(1e-3, -1)
    assert os.path.isdir(os.path.join("BASE_ppr_results", "got.db", "0.001_-1")) and base.getPrepSize() > 0
    optr, otgt, oval, _ = O.b_base_all_pairs(og, ALPHA, 1e-3)
    for v in (0, 5, 63):
        base.computeWholeGraphPPR(v, None)
        row = base.getWholeGraphPPR()
        assert list(row.keys()) == [int(t) for t in otgt[optr[v]:optr[v + 1]]]
        assert list(row.values()) == [float(x) * O.FX_INV for x in oval[optr[v]:optr[v + 1]]]   # text round trip is exact
    base.deletePrepDir()


def test_harness_rows(pkg, ctx, got, tmp_path):
    t = pkg.PPR()
    t.ctx, t.dir_db = ctx, "got.db"
    t.adjM = pkg.Graph(ctx, got[0], got[1])
    t.set_configuration(ALPHA, EPS, "name", "Person", "Relation")
    with open(tmp_path / "rows.txt", "w") as fw:
        t.fw = fw
        r1 = t.algo_perf_test("FORA_WHOLE_GRAPH", 10, -1, 0.5, -1.0, False, "WHOLE_GRAPH")
        r2 = t.algo_perf_test("FORA_TOPK", 10, 10, 0.5, -1.0, False, "TOPK")
        r3 = t.algo_perf_test("FWDPUSH", 10, -1, 1e-6, -1.0, False, "WHOLE_GRAPH")
    lines = open(tmp_path / "rows.txt").read().splitlines()
    assert lines[0].startswith("0.5,") and lines[1].startswith("0.5,10,") and lines[2].startswith("1.0E-6,")
    assert r1["max_err"] < 0.05 and r2["precision"] > 0.8 and r2["ndcg"] > 0.97 and r3["max_err"] < 1e-4
    # the same rows with computeError evaluated in HBM (device_metrics): same queries (rng state restored), same seeds (fresh algorithm objects)
    for dev in (False, True):
        t.device_metrics = dev
        rows = []
        with open(tmp_path / ("rows_%d.txt" % dev), "w") as fw:
            t.fw = fw
            for args in (("FORA_WHOLE_GRAPH", 8, -1, 0.5, -1.0, False, "WHOLE_GRAPH"), ("FORA_TOPK", 8, 10, 0.5, -1.0, False, "TOPK"),
                         ("MC", 8, 10, 1.0, -1.0, False, "TOPK"), ("FWDPUSH", 8, -1, 1e-6, -1.0, False, "WHOLE_GRAPH")):
                t.rng = np.random.default_rng(99)
                rows.append(t.algo_perf_test(*args))
        if dev:
            for a, b in zip(host_rows, rows):
                for key in ("max_err", "precision", "ndcg"):
                    if key in a:
                        assert abs(a[key] - b[key]) < 1e-12, (a["algo"], key, a[key], b[key])
        else:
            host_rows = rows
    t.adjM.close()


def test_full_size_properties(pkg, ctx):
    """R-MAT scale-20 (config 2): no oracle at this size; size-independent properties instead."""
    rp, ci = pkg.rmat(1 << 20, 1 << 24, seed=42)
    g = pkg.Graph(ctx, rp, ci)
    deg = np.diff(rp)
    try:
        src = int(np.flatnonzero(deg > 0)[1234])
        # push: exact mass conservation and the exit condition r < rmax * d everywhere
        _, _, rsum, st = g.forward_push(src, ALPHA, 1e-7, want_residue=False)
        res, r = g.last_raw()
        assert int(res.sum(dtype=np.uint64)) + int(r.sum(dtype=np.uint64)) == pkg.FX_ONE
        rfx = int(1e-7 * 2.0 ** 62)
        nz = np.flatnonzero(r)
        assert (r[nz].astype(object) < np.where(deg[nz] == 0, 1, rfx * deg[nz].astype(object))).all()
        # idempotence: same call, same bits
        g.forward_push(src, ALPHA, 1e-7, want_residue=False)
        res2, r2 = g.last_raw()
        assert np.array_equal(res, res2) and np.array_equal(r, r2)
        # FORA: all mass ends in the estimate; shards sum to the unsharded vector; seeds matter
        ppr, st = g.fora(src, ALPHA, EPS, 11)
        raw = g.last_raw(False)[0]
        assert int(raw.sum(dtype=np.uint64)) == pkg.FX_ONE and st["walks"] == st["walks_total"] > 0
        acc = np.zeros_like(raw)
        for rk in range(2):
            g.fora_shard(src, ALPHA, EPS, 11, rk, 2)
            acc += g.last_raw(False)[0]
        assert np.array_equal(acc, raw)
        g.fora(src, ALPHA, EPS, 12)
        assert not np.array_equal(g.last_raw(False)[0], raw)
        # Monte-Carlo: counts add up to floor(omega)
        _, st = g.monte_carlo(src, ALPHA, 2.0, 3)
        assert int(g.last_raw(False)[0].sum(dtype=np.uint64)) == st["walks_total"] == int(np.floor(st["omega"]))
        # top-k of the FORA estimate is sorted, tie-inclusive and consistent with the dense vector
        ids, vals, _ = ctx.topk(ppr, 500)
        assert len(ids) >= 500 and (np.diff(vals) <= 0).all() and np.array_equal(vals, ppr[ids])
        assert (ppr[np.setdiff1d(np.arange(len(ppr)), ids)] < vals[-1]).all()
    finally:
        g.close()


def test_cpp_host_driver_on_got(pkg, tmp_path):
    """The compiled C++ `PPR` driver runs the reference's harness protocol on GOT and writes its result-file format."""
    import subprocess
    exe = os.path.join(os.path.dirname(pkg.__file__), "PPR")
    got = os.path.join(os.path.dirname(__file__), "golden", "got")
    r = subprocess.run([exe, "-alpha", "0.15", "-eps", "0.5", "-query", "5", "-k", "10", "-db", got, "-out", str(tmp_path)],
                       capture_output=True, text=True, timeout=600, cwd=str(tmp_path))      # the result stores of Test 3 are created under the cwd
    assert r.returncode == 0, r.stdout + r.stderr
    text = open(tmp_path / "got_AlgoPerfResults.txt").read()
    assert "Test 1. Whole-Graph test" in text and "1.1 FORA_WHOLE_GRAPH" in text and "2.1 FORA_TOPK" in text and "1.5 BASE_WHOLE_GRAPH" in text
    sec = text.split("1.1 FORA_WHOLE_GRAPH\n")[1].split("\n\n")[0].splitlines()
    assert [l.split(",")[0] for l in sec] == ["10.0", "5.0", "0.5", "0.1", "0.05"]        # J/Gen_Util.java:463
    errs = [float(l.split(",")[2].replace("E", "e")) for l in sec]
    assert errs[2] < 0.05 and errs[4] < errs[0]                      # eps = 0.5 row; error shrinks with eps
    sec = text.split("2.1 FORA_TOPK\n")[1].split("\n\n")[0].splitlines()
    prec = [float(l.split(",")[3].replace("E", "e")) for l in sec]
    assert min(prec[2:]) > 0.8
    sec = text.split("1.2 FWDPUSH\n")[1].split("\n\n")[0].splitlines()
    assert float(sec[-1].split(",")[2].replace("E", "e")) < 1e-6       # rmax 1e-8
    # Test 3 (J/Gen_Util.java:602-645): every source preprocessed into "<ALGO>_ppr_results/<db>/<param>/<id>.txt", queries answered from
    # the files, stores deleted afterwards (BASE keeps its index in memory).  Row: param, threshold, prep ms, store bytes, avg max error
    assert "Test 3. Preprocessing test" in text and "3.1 FORA_WHOLE_GRAPH" in text and "3.5 BASE_WHOLE_GRAPH" in text
    sec = text.split("3.2 FWDPUSH\n")[1].split("\n\n")[0].splitlines()
    assert [l.split(",")[0] for l in sec] == ["1.0E-4", "1.0E-5", "1.0E-6", "1.0E-7", "1.0E-8"] and all(l.split(",")[1] == "-1.0" for l in sec)
    assert all(int(l.split(",")[3]) > 107 * 4 for l in sec)            # bytes on disk: at least "s\t1.0\n" per source
    assert float(sec[-1].split(",")[4].replace("E", "e")) < 1e-6       # values survive the Double.toString round trip
    sec = text.split("3.1 FORA_WHOLE_GRAPH\n")[1].split("\n\n")[0].splitlines()
    assert float(sec[2].split(",")[4].replace("E", "e")) < 0.05
    assert not any(f.endswith(".txt") and f != "got_AlgoPerfResults.txt" for _, _, fs in os.walk(tmp_path) for f in fs)


def test_error_metrics_on_the_device(pkg, ctx, G, O, got, rmat12):
    """N4: Gen_Util.computeError (J/Gen_Util.java:259-326) evaluated in HBM (ppr_metrics_keep / ppr_metrics_compute) equals the
    oracle's orc_max_err / orc_topk_metrics on the downloaded vectors: MAX_ERR exactly, PRECISION exactly, NDCG to 1e-12."""
    for g, (rp, ci), srcs, k in ((G, got[:2], (63, 0, 71), 10), (None, rmat12, None, 50)):
        own = g is None
        if own:
            g = pkg.Graph(ctx, rp, ci)
            srcs = [int(x) for x in np.flatnonzero(np.diff(rp) > 0)[[3, 40, 400]]]
        try:
            for s in srcs:
                truth = g.power_iteration(s, ALPHA, 100)
                g.metrics_keep(0)
                for run in ("fora", "fwdpush", "mc", "topk"):
                    if run == "fora":
                        est, _ = g.fora(s, ALPHA, EPS, 7)
                    elif run == "fwdpush":
                        est = g.forward_push(s, ALPHA, 1e-6, want_residue=False)[0]
                    elif run == "mc":
                        est, _ = g.monte_carlo(s, ALPHA, 1.0, 7)
                    else:
                        g.fora_topk(s, ALPHA, EPS, k, 7)
                        est = g.last_raw(False)[0] * O.FX_INV
                    g.metrics_keep(1)
                    m = g.metrics_compute(k)
                    assert m["max_err"] == O.max_err(truth, est), (run, s)
                    ref = O.topk_metrics(truth, k, O.topk_select(est, k))
                    assert m["precision"] == ref["precision"], (run, s, m, ref)
                    assert abs(m["ndcg"] - ref["ndcg"]) < 1e-12, (run, s, m, ref)
                    assert 0.0 <= m["precision"] <= 1.0 and m["max_err"] < 0.05
            m0 = g.metrics_compute(0)
            assert m0["precision"] == 0.0 and m0["max_err"] == m["max_err"]
        finally:
            if own:
                g.close()
