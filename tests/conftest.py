import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "personalized-pagerank-algorithms-on-neo4j_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    so = os.path.join(ROOT, PKG, "libpprb200.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def got(pkg):
    rp, ci, names = pkg.graphio.load_got()
    return rp, ci, names


@pytest.fixture(scope="session")
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


def small_rmat(pkg, scale, ef=16, seed=42, **kw):
    n = 1 << scale
    return pkg.rmat(n, n * ef, seed=seed, **kw)


@pytest.fixture(scope="session")
def rmat12(pkg):
    return small_rmat(pkg, 12)


@pytest.fixture(scope="session")
def rmat16(pkg):
    return small_rmat(pkg, 16)
