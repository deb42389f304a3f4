"""B200-native Personalized-PageRank engine: host-side mirror of the reference's operator API over
libpprb200.so (CUDA, sm_100a).  The directory name contains '-', so import it with

    import importlib; ppr = importlib.import_module("personalized-pagerank-algorithms-on-neo4j_b200")

Importing never touches oracle/ (test infrastructure) and never falls back to the CPU: the native
library is loaded lazily by `_native.load()` and raises ImportError if it has not been built.
"""
from . import _native, algo, graphio, harness, dist  # noqa: F401
from ._native import Context, Graph, PPRError, rmat, load, FX_ONE, FX_INV, RESTART_AT_SOURCE, RESTART_AT_WALK_START  # noqa: F401
from .algo import (Algo_Conf, Algo_Util, Backward_Search, Base_Whole_Graph, Fora_Topk, Fora_Whole_Graph,  # noqa: F401
                   Forward_Push, Monte_Carlo, Power_Method, PPRVector)
from .harness import PPR, Gen_Util, compute_error  # noqa: F401
