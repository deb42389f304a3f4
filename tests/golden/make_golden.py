"""Regenerates tests/golden/got_golden.npz from the CPU oracle (oracle/ppr_oracle.cpp).

The reference cannot run in this image (Java/Neo4j, no JDK) and ships no golden vectors, so these
fixtures pin the ORACLE's behaviour (regression) and give the GPU tests committed vectors to compare
against on the GPU box, where /root/reference does not exist.  External anchors inside: the GOT
power-method values recorded in SURVEY.md 8c and the Random123 Philox known-answer vectors.

    python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

pkg = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")

ALPHA, EPS, SEED, K = 0.15, 0.5, 7, 10
SOURCES = [0, 63, 71, 5, 100]  # Aemon, Tyrion, Grenn (dangling), ...


def main():
    rp, ci, names = pkg.graphio.load_got()
    g = O.OGraph(rp, ci)
    out = {"row_ptr": rp, "col_idx": ci, "sources": np.array(SOURCES)}
    for s in SOURCES:
        out["pm_%d" % s] = O.power_method(g, s, ALPHA, 100)
        res, r, st = O.b_forward_push(g, s, ALPHA, 1e-6)
        out["push_reserve_%d" % s], out["push_residue_%d" % s] = res, r
        out["push_stats_%d" % s] = np.array([st["pushes"], st["edges"], st["iters"]])
        f, st = O.b_fora(g, s, ALPHA, EPS, SEED)
        out["fora_%d" % s] = f
        out["fora_stats_%d" % s] = np.array([st[k] for k in ("pushes", "edges", "iters", "rounds", "walks", "steps", "walk_sources")])
        c, st = O.b_monte_carlo(g, s, ALPHA, EPS, SEED)
        out["mc_%d" % s] = c
        t, st = O.b_fora_topk(g, s, ALPHA, EPS, K, SEED)
        out["topk_est_%d" % s] = t
        out["topk_ids_%d" % s] = O.topk_select(t, K)
        b, _, st = O.b_backward_search(g, s, ALPHA, 1e-3)
        out["bwd_%d" % s] = b
    for thr in (1e-3, 5e-5):
        ptr, tgt, val, st = O.b_base_all_pairs(g, ALPHA, thr)
        out["ap_ptr_%g" % thr], out["ap_tgt_%g" % thr], out["ap_val_%g" % thr] = ptr, tgt, val
    out["walk_terminals_63"] = np.array([O.walk(g, 63, 63, j, SEED, 0, ALPHA, True)[0] for j in range(256)], np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "got_golden.npz"), **out)
    print("wrote got_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
