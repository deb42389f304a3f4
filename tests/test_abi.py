"""CPU: the C-ABI library loads, exports every symbol include/pprb200.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(name):
    text = open(os.path.join(ROOT, "include", name)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ppr_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(pkg):
    L = pkg.load()
    syms, dbg = header_symbols("pprb200.h"), header_symbols("pprb200_debug.h")
    assert len(syms) >= 40 and len(dbg) >= 8
    for s in syms + dbg:
        assert hasattr(L, s), "libpprb200.so does not export " + s
    assert sorted(pkg._native.EXPORTS) == syms and sorted(pkg._native.DEBUG_EXPORTS) == dbg
    # measurement / test hooks and the synthetic-input generator are not part of the product header
    assert not set(syms) & set(dbg) and "ppr_rmat_generate" not in syms + dbg and not hasattr(L, "ppr_rmat_generate")
    G = pkg._native.load_gen()
    assert header_symbols("pprgen.h") == pkg._native.GEN_EXPORTS and hasattr(G, "ppr_rmat_generate")


def test_library_does_not_link_nccl_or_the_oracle(pkg):
    """NCCL is bound at run time (csrc/comm.hpp): single-GPU use must not need it, and a host process that already carries
    an NCCL (PyTorch) must not get a second one."""
    import subprocess
    out = subprocess.run(["ldd", pkg._native.LIB_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in out and "oracle" not in out
    out = subprocess.run(["ldd", pkg._native.GEN_PATH], capture_output=True, text=True).stdout
    assert "cuda" not in out and "pprb200" not in out


def test_library_is_sm100a_cuda(pkg):
    # the shipped binary carries sm_100a SASS (no PTX-only / other-arch build)
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", pkg._native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_stats_struct_layout(pkg):
    assert ctypes.sizeof(pkg._native.Stats) == 112 and ctypes.sizeof(pkg._native.Conf) == 24


def test_rmat_generator(pkg):
    rp, ci = pkg.rmat(1 << 10, 1 << 14, seed=3)
    rp2, ci2 = pkg.rmat(1 << 10, 1 << 14, seed=3)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
    assert rp[0] == 0 and rp[-1] == 1 << 14 and (np.diff(rp) >= 0).all() and ci.min() >= 0 and ci.max() < 1 << 10
    for v in range(1 << 10):
        assert (np.diff(ci[rp[v]:rp[v + 1]]) >= 0).all()
    rp3, ci3 = pkg.rmat(1 << 10, 1 << 14, seed=4)
    assert not np.array_equal(ci, ci3)
    # skew: R-MAT's heaviest vertex owns far more than the mean degree
    assert np.diff(rp).max() > 20 * 16
    rp, ci = pkg.rmat(1000, 16000, seed=3, symmetric=True)   # non power of two + symmetrised
    src = np.repeat(np.arange(1000), np.diff(rp))
    fwd = np.sort(src * 1000 + ci)
    bwd = np.sort(ci.astype(np.int64) * 1000 + src)
    assert np.array_equal(fwd, bwd)
    with pytest.raises(pkg.PPRError):
        pkg.rmat(1000, 15, symmetric=True)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="CPU-only behaviour")
def test_no_cpu_fallback(pkg):
    with pytest.raises(pkg.PPRError) as e:
        pkg.Context(0)
    assert e.value.code == pkg._native.PPR_E_CUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle(pkg):
    pkg_dir = os.path.dirname(pkg.__file__)
    for root, _, files in os.walk(pkg_dir):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ppr_oracle" not in txt, fn


def test_cpp_host_driver_builds_and_fails_loudly(pkg, tmp_path):
    """host/ppr_main.cpp: the reference's `PPR` driver (same options) compiled against the C ABI."""
    import subprocess
    exe = os.path.join(os.path.dirname(pkg.__file__), "PPR")
    assert os.path.exists(exe)
    out = subprocess.run([exe, "-help"], capture_output=True, text=True)
    assert out.returncode == 0
    for opt in ("-alpha", "-eps", "-query", "-k", "-node", "-label", "-rel", "-db", "-help"):
        assert opt in out.stdout                                   # J/PPR.java:157-166
    assert subprocess.run([exe, "-bogus"], capture_output=True).returncode == 2
    if not os.path.exists("/dev/nvidia0"):
        r = subprocess.run([exe, "-query", "1", "-db", os.path.join(ROOT, "tests", "golden", "got"), "-out", str(tmp_path)], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stdout


def test_cpp_host_print_rows(pkg, tmp_path):
    """The print loop of the C++ host classes (host/ppr_host.hpp print_rows, J/Forward_Push.java:268-281, :366-391) on hand-made
    results: same lines as the Python mirror prints (tests/test_api_surface.py).  Host-only code: no device call is made."""
    import subprocess
    pdir = os.path.dirname(pkg.__file__)
    src = tmp_path / "rows.cpp"
    src.write_text('''
#include "%s/host/ppr_host.hpp"
using namespace fora_neo4j;
int main() {
    Graph G;
    G.names = {"Aemon", "Grenn", "Samwell", "Craster"};
    std::vector<double> reserve = {0.5, 0.0, 0.25, 1e-4}, residue = {0.0, 0.125, 0.0, 2e-5};
    Algo_Util::print_rows(G, "Forward-Push Reserve(pi) & Residue(r):", Algo_Util::rows_of(reserve), SIZE_MAX, true, &residue);
    Algo_Util::print_rows(G, "\\nForward-Push-Top2 PPR:", Algo_Util::rows_of(std::vector<int64_t>{0, 2, 3}, reserve), 2);
    Algo_Util::Rows stored = {{1, 0.125}, {0, 0.5}, {9, 0.5}};
    Algo_Util::print_rows(G, "\\nBase-Whole-Graph PPR:", stored, SIZE_MAX, false);
    Algo_Util::print_rows(G, "\\nBASE-Top2 PPR:", stored, 2);
    return 0;
}
''' % pdir)
    exe = tmp_path / "rows"
    cc = subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-o", str(exe), str(src), "-L" + pdir, "-lpprb200", "-Wl,-rpath," + pdir],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.splitlines() == [
        "Forward-Push Reserve(pi) & Residue(r):", "@Aemon\t0.5\t0.0", "@Samwell\t0.25\t0.0", "@Craster\t1.0E-4\t2.0E-5",
        "", "Forward-Push-Top2 PPR:", "@Aemon\t0.5", "@Samwell\t0.25",
        "", "Base-Whole-Graph PPR:", "@Grenn\t0.125", "@Aemon\t0.5", "@9\t0.5",
        "", "BASE-Top2 PPR:", "@Aemon\t0.5", "@9\t0.5"]


def test_cpp_host_reads_the_neo4j_store(pkg, tmp_path):
    """host/ppr_host.hpp read_neo4j_store (15-/34-byte records) against read_import_csv on the GOT fixtures: the reference's own
    `-db target/got.db` argument works in the compiled driver without a JVM.  Host-only: no device call is made."""
    import subprocess
    pdir = os.path.dirname(pkg.__file__)
    src = tmp_path / "store.cpp"
    src.write_text('''
#include "%s/host/ppr_host.hpp"
using namespace fora_neo4j;
int main(int argc, char** argv) {
    std::vector<std::string> n1, n2;
    PPR::EdgeList e1, e2;
    PPR::read_neo4j_store(argv[1], n1, e1);
    PPR::read_import_csv(argv[2], argv[3], n2, e2);
    std::printf("%%zu %%zu %%zu %%zu %%d\\n", n1.size(), e1.size(), n2.size(), e2.size(), (int)(e1 == e2));
    std::printf("%%s %%s %%lld %%lld\\n", n1.front().c_str(), n2.front().c_str(), (long long)e1.front().first, (long long)e1.front().second);
    return 0;
}
''' % pdir)
    exe = tmp_path / "store"
    cc = subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-o", str(exe), str(src), "-L" + pdir, "-lpprb200", "-Wl,-rpath," + pdir],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-3000:]
    g = os.path.join(ROOT, "tests", "golden")
    out = subprocess.run([str(exe), os.path.join(g, "got.db"), os.path.join(g, "got", "GOT_Nodes.csv"), os.path.join(g, "got", "GOT_Rels.csv")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert lines[0] == "107 352 107 352 1"                  # same relationships in the same (relationship id = CSV row) order: SURVEY F10
    rp, ci, ids = pkg.graphio.load_neo4j_store(os.path.join(g, "got.db"))
    assert lines[1].split()[0] == str(int(ids[0]))
