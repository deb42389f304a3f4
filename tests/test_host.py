"""CPU: host logic -- graph export, result maps, preprocessing store, shard arithmetic, gloo world_size 2."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_csv_export(pkg, tmp_path):
    (tmp_path / "n.csv").write_text("﻿:ID,name\nb,Bee\na,Ay\nc,Cee\n", encoding="utf-8")
    (tmp_path / "r.csv").write_text(":START_ID,:END_ID,:TYPE\nb,a,R\nb,c,R\nb,a,R\nc,b,R\n")
    rp, ci, names = pkg.graphio.load_neo4j_csv(str(tmp_path / "n.csv"), str(tmp_path / "r.csv"))
    assert names == ["Bee", "Ay", "Cee"]                       # id = row index
    assert list(rp) == [0, 3, 3, 4] and list(ci) == [1, 1, 2, 0]   # ascending targets, duplicates kept
    with pytest.raises(ValueError):
        pkg.graphio.csr_from_edges(2, [0], [5])


def test_ppr_vector_is_a_sparse_map(pkg):
    v = pkg.PPRVector(np.array([0.0, 0.25, 0.0, 0.75]))
    assert dict(v) == {1: 0.25, 3: 0.75} and len(v) == 2 and v.get(0) is None and v.get(3) == 0.75
    with pytest.raises(KeyError):
        v[2]


def test_java_double_format(pkg):
    f = pkg.algo._java_double
    assert f(0.5) == "0.5" and f(1.0) == "1.0" and f(0.001) == "0.001"
    assert f(5e-4) == "5.0E-4" and f(1.2345e-7) == "1.2345E-7" and f(1e7) == "1.0E7"
    for x in (0.399076058, 3.3e-9, 12345678.9):
        assert float(f(x).replace("E", "e")) == x


def test_prep_store_roundtrip(pkg, tmp_path):
    A = pkg.algo
    d = str(tmp_path / "FWP_ppr_results" / "got.db" / "1.0E-6")
    A._prep_clean(d)
    vec = np.zeros(10)
    vec[[2, 7]] = [0.25, 3.5e-9]
    A._prep_write(d, 4, pkg.PPRVector(vec))
    assert open(os.path.join(d, "4.txt")).read() == "2\t0.25\n7\t3.5E-9\n"   # J/Forward_Push.java:321-329
    assert np.array_equal(A._prep_read(d, 4, 10), vec)
    assert A._dir_size(d) == len("2\t0.25\n7\t3.5E-9\n")


def test_algo_conf_mirror(pkg):
    ac = pkg.Algo_Conf(0.15)
    t = ac.set_conf_fora_topk(107, 352, 10, None, None, "got.db")
    assert t.delta == 0.1 and t.min_delta == 1 / 107 and t.pfail == 1.0 / 107 / 107 / np.log(10.0)
    f = ac.set_conf_fora_whole_graph(107, 352, None, None, "got.db")
    assert f.delta == f.pfail == 1 / 107 and f.rsum == 1.0
    assert pkg.Algo_Util().kth_ppr([0.1, 0.2], 3) is None and pkg.Algo_Util().kth_ppr([0.1, 0.2], 0) is None


def test_shard_arithmetic(pkg):
    D = pkg.dist
    for total in (0, 1, 7, 1000, 10 ** 10 + 7):
        for world in (1, 2, 3, 8):
            rs = [D.shard_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total and all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
    assert D.target_blocks(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]


def test_merge_pairs(pkg):
    D = pkg.dist
    a = (np.array([0, 1, 1, 3]), np.array([0, 0, 1], np.int32), np.array([.5, .25, .125]))   # targets 0..1
    b = (np.array([0, 0, 2, 2]), np.array([2, 3], np.int32), np.array([.75, .0625]))          # targets 2..3
    ptr, tgt, val = D.merge_pairs([a, b], 3)
    assert list(ptr) == [0, 1, 3, 5] and list(tgt) == [0, 2, 3, 0, 1] and list(val) == [.5, .75, .0625, .25, .125]


def _gloo_worker(rank, world, port, q):
    import importlib
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rp, ci, _ = pkg.graphio.load_got()
    g = O.OGraph(rp, ci)
    # walk shards: each rank replays its slice with the oracle; the integer all-reduce must equal the 1-rank answer
    part, _ = O.b_fora(g, 63, 0.15, 0.5, 7, shard=rank, nshards=world)
    t = torch.from_numpy(part.view(np.int64).copy())
    dist.all_reduce(t)
    full, _ = O.b_fora(g, 63, 0.15, 0.5, 7)
    ok1 = bool((t.numpy().view(np.uint64) == full).all())
    # all-pair target blocks: gather + merge equals the single-rank inverted index
    lo, hi = pkg.dist.target_blocks(g.n, world)[rank]
    local = O.b_base_all_pairs(g, 0.15, 1e-3, lo, hi)[:3]
    ptr, tgt, val = pkg.dist.gather_pairs(local, g.n)
    rptr, rtgt, rval, _ = O.b_base_all_pairs(g, 0.15, 1e-3)
    ok2 = np.array_equal(ptr, rptr) and np.array_equal(tgt, rtgt) and np.array_equal(val, rval)
    # top-k rows (k = 3) are cut AFTER the merge, over the union of the ranks' full rows (a rank's own cut is not the merged row's)
    kptr, ktgt, kval = pkg.dist.gather_pairs(local, g.n, k=3)
    for v in range(g.n):
        rt, rv = rtgt[rptr[v]:rptr[v + 1]], rval[rptr[v]:rptr[v + 1]]
        keep = O.topk_select(rv, 3) if len(rv) else np.zeros(0, np.int32)
        ok2 = ok2 and np.array_equal(ktgt[kptr[v]:kptr[v + 1]], rt[keep]) and np.array_equal(kval[kptr[v]:kptr[v + 1]], rv[keep])
    # FORA top-k with every round's walks sharded: one sum of the estimate vector per round, all ranks agree
    def reduce(est):
        t = torch.from_numpy(est.view(np.int64))       # shares memory with the oracle's vector
        dist.all_reduce(t)
    est, st = O.b_fora_topk(g, 63, 0.15, 0.5, 10, 7, shard=rank, nshards=world, reduce=reduce)
    ref, rst = O.b_fora_topk(g, 63, 0.15, 0.5, 10, 7)
    ok3 = bool(np.array_equal(est, ref)) and st["rounds"] == rst["rounds"] and st["walks"] < rst["walks"]
    ok2 = ok2 and ok3
    q.put((rank, ok1, ok2))
    dist.destroy_process_group()


def test_gloo_world2_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(60)
    assert sorted(res) == [(0, True, True), (1, True, True)]


def test_neo4j_store_reader_matches_csv(pkg):
    """N3: the topology read natively from the Neo4j 3.x store files equals the import CSVs."""
    store = os.path.join(ROOT, "tests", "golden", "got.db")
    rp, ci, ids = pkg.graphio.load_neo4j_store(store)
    rp2, ci2, names = pkg.graphio.load_got()
    assert np.array_equal(ids, np.arange(107)) and np.array_equal(rp, rp2) and np.array_equal(ci, ci2)


def test_bench_reference_arm_contract(pkg):
    """`bench.py --impl reference` (the oracle on the host cores) prints the contract line without a GPU; small scale here."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--scale", "12", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "FORA SSPPR queries/s" and line["unit"] == "queries/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["config"]["workload"].startswith("R-MAT scale-12") and line["config"]["l2"].startswith("inputs fit the L2")
    assert line["oracle_pinned"] is False


def test_edge_list_loader_and_csv_writer(pkg, tmp_path):
    """SNAP-style edge list -> CSR (ids relabelled in ascending order), undirected doubling, and the neo4j-admin CSV round trip."""
    (tmp_path / "g.txt").write_text("# Directed graph\n# FromNodeId\tToNodeId\n30\t10\n10\t20\n30\t20\n30\t10\n20\t20\n")
    rp, ci, ids = pkg.graphio.load_edge_list(str(tmp_path / "g.txt"))
    assert list(ids) == [10, 20, 30]
    assert list(rp) == [0, 1, 2, 5] and list(ci) == [1, 1, 0, 0, 1]           # parallel edge and self-loop kept
    rp, ci, ids = pkg.graphio.load_edge_list(str(tmp_path / "g.txt"), undirected=True)
    assert list(rp) == [0, 3, 6, 9] and list(ci) == [1, 2, 2, 0, 1, 2, 0, 0, 1]  # both directions; the self-loop stored once
    (tmp_path / "e.csv").write_text("1,2\n2,3\n")
    rp, ci, ids = pkg.graphio.load_edge_list(str(tmp_path / "e.csv"), delimiter=",")
    assert list(rp) == [0, 1, 2, 2] and list(ci) == [1, 2] and list(ids) == [1, 2, 3]
    with pytest.raises(ValueError):
        (tmp_path / "bad.txt").write_text("7\n")
        pkg.graphio.load_edge_list(str(tmp_path / "bad.txt"))
    grp, gci, names = pkg.graphio.load_got()
    pkg.graphio.write_neo4j_csv(grp, gci, str(tmp_path / "N.csv"), str(tmp_path / "R.csv"), names)
    rp2, ci2, names2 = pkg.graphio.load_neo4j_csv(str(tmp_path / "N.csv"), str(tmp_path / "R.csv"))
    assert np.array_equal(rp2, grp) and np.array_equal(ci2, gci) and names2 == list(names)
