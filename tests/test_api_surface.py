"""CPU: the reference's public method surface (tests/golden/api_surface.json, generated from the Java by
tests/golden/make_api_surface.py) against the three host mirrors of this repo -- the Python classes
(algo.py / harness.py), the C++ header (host/ppr_host.hpp) and the rewired Java sources (java/).  A user of
the reference must find every method of the in-scope classes, or an explicit waiver saying where it went."""
import inspect
import io
import json
import os
import re
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SURFACE = json.load(open(os.path.join(HERE, "golden", "api_surface.json")))

# Out of scope by SURVEY section 2 / section 8: the Neo4j competitor, and the three pure interfaces (their methods are
# checked through the classes that implement them).
SKIP_CLASSES = {"Neo4j_Method", "Preprocessing_Interface", "Topk_Util_Interface", "Whole_Graph_Util_Interface"}
WAIVED = {
    ("Algo_Conf", "set_conf_neo4j_method"): "Neo4j's built-in PageRank competitor is out of scope (SURVEY section 2)",
    ("Forward_Push", "forward_push_topk"): "only caller is Fora_Topk; the incremental push runs inside ppr_fora_topk (a3)",
    ("Forward_Push", "getNextQueueCopy"): "state of forward_push_topk, stays on the device between rounds",
    ("PPR", "run"): "body of the anonymous JVM shutdown-hook thread, not a method of PPR",
}
CPP_WAIVED = {
    ("PPR", "create_index"), ("PPR", "main"),                                         # no schema index; main() lives in host/ppr_main.cpp
    ("Gen_Util", "newAlgoObj"), ("Gen_Util", "algo_perf_test"),                       # split into whole_graph_test/topk_test/prep_test
}
CPP_PRINTS = set()
PY_ARITY_WAIVED = set()             # none needed: every mirrored method takes the reference's positional arguments
JAVA_CLASSES = ["Forward_Push", "Monte_Carlo", "Fora_Whole_Graph", "Fora_Topk", "Backward_Search", "Base_Whole_Graph", "Power_Method"]


def in_scope():
    for cls, methods in SURFACE.items():
        if cls in SKIP_CLASSES:
            continue
        for name, nargs, where in methods:
            if (cls, name) not in WAIVED:
                yield cls, name, nargs, where


def test_fixture_is_the_reference_surface():
    assert set(SURFACE) == {"Algo_Conf", "Algo_Util", "Backward_Search", "Base_Whole_Graph", "Fora_Topk", "Fora_Whole_Graph",
                            "Forward_Push", "Gen_Util", "Monte_Carlo", "Neo4j_Method", "PPR", "Power_Method",
                            "Preprocessing_Interface", "Topk_Util_Interface", "Whole_Graph_Util_Interface"}
    assert ["computeWholeGraphPPR", 2, "Forward_Push.java:63"] in SURFACE["Forward_Push"]
    assert ["backward_search_whole_graph", 1, "Backward_Search.java:38"] in SURFACE["Backward_Search"]
    assert all((c, n) in [(c2, m[0]) for c2, ms in SURFACE.items() for m in ms] for c, n in WAIVED)


def test_python_mirror_has_every_method(pkg):
    missing = []
    for cls, name, nargs, where in in_scope():
        owner = getattr(pkg.harness, cls, None) or getattr(pkg.algo, cls, None)
        if owner is None or not callable(getattr(owner, name, None)):
            missing.append((cls, name, where))
            continue
        # the method must take the reference's argument count positionally (extra optional arguments are fine)
        params = [p for p in inspect.signature(getattr(owner, name)).parameters.values() if p.name != "self"]
        required = sum(1 for p in params if p.default is p.empty and p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD))
        varargs = any(p.kind == p.VAR_POSITIONAL for p in params)
        positional = sum(1 for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD))
        if (cls, name) in PY_ARITY_WAIVED:
            continue
        if required > nargs or (positional < nargs and not varargs):
            missing.append((cls, name, where, "arity %d, reference %d" % (positional, nargs)))
    assert not missing, missing


def test_cpp_mirror_has_every_method():
    src = open(os.path.join(ROOT, "personalized-pagerank-algorithms-on-neo4j_b200", "host", "ppr_host.hpp")).read()
    bodies = {}
    for m in re.finditer(r"^struct (\w+)[^{;]*\{", src, re.M):
        end = src.find("\n};", m.end())
        bodies[m.group(1)] = src[m.end():end]
    missing = []
    for cls, name, nargs, where in in_scope():
        if (cls, name) in CPP_WAIVED or name in CPP_PRINTS:
            continue
        body = bodies.get(cls, "")
        for base in re.findall(r"struct %s\s*:\s*([^{]*)\{" % cls, src):      # inherited mixins (Prep_Store, Algo_Util, Gen_Util)
            for b in re.findall(r"(\w+)", base):
                body += bodies.get(b, "")
        if not re.search(r"\b%s\s*\(" % name, body):
            missing.append((cls, name, where))
    assert not missing, missing


def test_java_rewires_have_every_method():
    jdir = os.path.join(ROOT, "java", "joezie", "fora_neo4j")
    base = open(os.path.join(jdir, "Gpu_Algo_Base.java")).read()
    missing = []
    for cls in JAVA_CLASSES:
        src = open(os.path.join(jdir, cls + ".java")).read()
        assert src.count("{") == src.count("}") and src.count("(") == src.count(")"), cls
        assert re.search(r"package\s+joezie\.fora_neo4j\s*;", src) and re.search(r"class\s+%s\b" % cls, src)
        inherits = "extends Gpu_Algo_Base" in src
        for name, nargs, where in SURFACE[cls]:
            if (cls, name) in WAIVED and name != "getNextQueueCopy":
                continue
            pat = r"\b%s\s*\(" % name
            if not (re.search(pat, src) or (inherits and re.search(pat, base))):
                missing.append((cls, name, where))
    assert not missing, missing
    assert base.count("{") == base.count("}")


def test_print_methods_format(pkg):
    """'@name<TAB>value' lines, descending, Double.toString formatting, at most k rows (J/Forward_Push.java:268-281,
    :366-391; J/Fora_Topk.java:60-79; J/Base_Whole_Graph.java:188-193).  No device needed: results are set by hand."""
    A = pkg.algo
    names = ["Aemon", "Grenn", "Samwell", "Craster"]
    fp = A.Forward_Push(0.15, 1.0, 4, names, None)
    fp.reserve, fp.residue = np.array([0.5, 0.0, 0.25, 1e-4]), np.array([0.0, 0.125, 0.0, 2e-5])
    buf = io.StringIO()
    with redirect_stdout(buf):
        fp.printWholeGraphResult()
    assert buf.getvalue().splitlines() == ["Forward-Push Reserve(pi) & Residue(r):", "@Aemon\t0.5\t0.0", "@Samwell\t0.25\t0.0",
                                           "@Craster\t1.0E-4\t2.0E-5"]
    fp.topk_nodeIds = np.array([0, 2, 3])
    buf = io.StringIO()
    with redirect_stdout(buf):
        fp.printTopKResult(2)
    assert buf.getvalue().splitlines() == ["", "Forward-Push-Top2 PPR:", "@Aemon\t0.5", "@Samwell\t0.25"]
    ft = A.Fora_Topk(0.15, 1.0, 0.1, 0.5, 4, 6, None, 0.25, 2, None)
    ft.topk_nodeIds, ft.topk_vals = np.array([3, 1, 2]), np.array([0.4, 0.3, 0.3])
    buf = io.StringIO()
    with redirect_stdout(buf):
        ft.printTopKResult(0)
    assert buf.getvalue().splitlines() == ["", "Fora-Top2 PPR:", "@3\t0.4", "@1\t0.3"]          # no name column: ids
    bw = A.Base_Whole_Graph(0.15, 4, {0: "a", 1: "b"}, None, write_files=False)
    bw.ppr_src = {1: 0.125, 0: 0.5}
    buf = io.StringIO()
    with redirect_stdout(buf):
        bw.printWholeGraphResult()
        bw.printTopKResult(1)
    assert buf.getvalue().splitlines() == ["", "Base-Whole-Graph PPR:", "@b\t0.125", "@a\t0.5", "", "BASE-Top1 PPR:", "@a\t0.5"]
    bs = A.Backward_Search(0.15, 1e-3, 4, None, None)
    assert bs.getDuration() == 0
    assert fp.getNodeName(1) == "Grenn" and fp.getNodeName(9) == "9" and ft.getNodeName(2) == "2"
