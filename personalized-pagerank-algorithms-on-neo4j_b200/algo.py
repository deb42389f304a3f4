"""Host-side mirror of the reference's algorithm classes (same names, same method names, same argument
meaning), each a thin shell over one C-ABI call of libpprb200.so.

Reference: /root/reference/src/main/java/joezie/fora_neo4j/ (J/ below).  The Java host layer would look
exactly like this (see java/ and INTEGRATION.md); Python is used here because the image has no JDK.
`graphDb` arguments are accepted and ignored (Neo4j is only needed for node names, out of scope);
`adjM` is a `_native.Graph`.  Vertex ids are the mapped ids (for GOT: the Neo4j ids).

Result maps: `getWholeGraphPPR()` returns a `PPRVector`, a read-only Mapping over the dense result
that iterates only the non-zero entries, i.e. the entries the Java HashMap would hold.
"""
import math
import os
import shutil
import time
from collections.abc import Mapping

import numpy as np

from . import _native as N


class PPRVector(Mapping):
    """Sparse-map view {node id: ppr} of a dense result vector (keys = entries > 0)."""

    def __init__(self, dense):
        self.dense = np.asarray(dense)
        self._keys = None

    def keys_array(self):
        if self._keys is None:
            self._keys = np.flatnonzero(self.dense > 0)
        return self._keys

    def __getitem__(self, k):
        v = self.dense[k]
        if not v > 0:
            raise KeyError(k)
        return float(v)

    def get(self, k, default=None):
        if 0 <= k < len(self.dense) and self.dense[k] > 0:
            return float(self.dense[k])
        return default

    def __iter__(self):
        return iter(int(k) for k in self.keys_array())

    def __len__(self):
        return int(self.keys_array().shape[0])

    def values_array(self):
        return self.dense[self.keys_array()]


class Algo_Util:
    """J/Algo_Util.java: getNodeName, kth_ppr (k-th largest; None when k is out of range)."""

    def __init__(self, node_property="name", graphDb=None):
        self.node_property = node_property
        self.graphDb = graphDb

    def getNodeName(self, nodeId, graphDb=None):
        """J/Algo_Util.java:20-30.  Neo4j is not here: `graphDb` stands for the node-property column, i.e. a
        sequence or mapping id -> name (what `graphio.load_got` returns); without one the id itself is printed."""
        db = self.graphDb if graphDb is None else graphDb
        try:
            return str(db[int(nodeId)])
        except (TypeError, KeyError, IndexError):
            return str(int(nodeId))

    def _print_rows(self, header, ids, vals, limit=None, extra=None):
        """The print loops of the reference classes (e.g. J/Forward_Push.java:268-281, :366-391): entries sorted
        by value descending (List.sort is stable; ties keep map order = ascending id here), '@name<TAB>value'."""
        ids, vals = np.asarray(ids), np.asarray(vals, dtype=np.float64)
        order = np.argsort(-vals, kind="stable")
        if limit is not None:
            order = order[:max(int(limit), 0)]
        print(header)
        for i in order:
            line = "@" + self.getNodeName(ids[i]) + "\t" + _java_double(vals[i])
            if extra is not None:
                line += "\t" + _java_double(extra[int(ids[i])])
            print(line)

    def _print_dense(self, header, dense, limit=None, extra=None):
        ids = np.flatnonzero(np.asarray(dense) > 0)
        self._print_rows(header, ids, np.asarray(dense)[ids], limit, extra)

    def kth_ppr(self, pprs, k, adjM=None):
        pprs = np.asarray(pprs, dtype=np.float64)
        if k < 1 or k > pprs.shape[0]:
            return None  # J/Algo_Util.java:44,52
        ctx = (adjM or getattr(self, "adjM")).ctx
        pos = pprs[pprs > 0]
        if k > pos.shape[0]:  # k-th largest is a non-positive entry; selection on device covers positives only
            return float(np.sort(pprs)[::-1][k - 1])
        _, _, kth = ctx.topk(pprs, k)
        return kth


def _topk_from_dense(adjM, dense, k):
    """retrieveTopK + getTopKNodeIds (J/Forward_Push.java:394-429): ids with value >= k-th, sorted descending."""
    ids, vals, _ = adjM.ctx.topk(dense, k)
    return ids, vals


class Forward_Push(Algo_Util):
    """J/Forward_Push.java -- Whole_Graph_Util_Interface, Topk_Util_Interface, Preprocessing_Interface."""

    def __init__(self, alpha, rsum, node_amount, graphDb, adjM, node_property="name", dir_db="db"):
        super().__init__(node_property, graphDb)
        self.alpha, self.rsum, self.node_amount, self.adjM, self.dir_db = alpha, rsum, node_amount, adjM, dir_db
        self.reserve = self.residue = None
        self.topk_nodeIds = None
        self.preprocessing_dirName = os.path.join("FWP_ppr_results", dir_db)
        self.stats = None

    def computeWholeGraphPPR(self, nodeId_start, rmax):
        self.topk_nodeIds = None
        self.reserve, self.residue, self.rsum, self.stats = self.adjM.forward_push(int(nodeId_start), self.alpha, float(rmax))

    def getUpdatedRsum(self):
        return self.rsum

    def getResidueCopy(self):
        return PPRVector(self.residue.copy())

    def getReserveCopy(self):
        return PPRVector(self.reserve.copy())

    def getWholeGraphPPR(self):
        return PPRVector(self.reserve)

    def printWholeGraphResult(self):
        """J/Forward_Push.java:268-281: reserve and residue of every node that holds reserve."""
        self._print_dense("Forward-Push Reserve(pi) & Residue(r):", self.reserve, extra=self.residue)

    def computeTopKPPR(self, nodeId_start, k, rmax):
        self.computeWholeGraphPPR(nodeId_start, rmax)

    def getTopKNodeIds(self, k):
        if self.topk_nodeIds is None:
            self.topk_nodeIds, _ = _topk_from_dense(self.adjM, self.reserve, k)
        return self.topk_nodeIds

    def printTopKResult(self, k):
        """J/Forward_Push.java:366-391."""
        ids = self.getTopKNodeIds(k)
        self._print_rows("\nForward-Push-Top%d PPR:" % k, ids, self.reserve[ids], limit=k)

    # Preprocessing_Interface (J/Forward_Push.java:289-362): one text file per source, "id\tvalue\n"
    def preprocessing(self, dummy, rmax):
        self.preprocessing_dirName = os.path.join(self.preprocessing_dirName, repr(float(rmax)))
        _prep_clean(self.preprocessing_dirName)
        for v in range(self.node_amount):
            self.computeWholeGraphPPR(v, rmax)
            _prep_write(self.preprocessing_dirName, v, PPRVector(self.reserve))

    def readPreprocessedPPR(self, nodeId_start):
        self.reserve = _prep_read(self.preprocessing_dirName, nodeId_start, self.node_amount)

    def getPrepSize(self):
        return _dir_size(self.preprocessing_dirName)

    def deletePrepDir(self):
        shutil.rmtree(self.preprocessing_dirName, ignore_errors=True)


class Monte_Carlo(Algo_Util):
    """J/Monte_Carlo.java."""

    def __init__(self, alpha, node_amount, graphDb, pfail, delta, adjM, node_property="name", dir_db="db", seed=1):
        super().__init__(node_property, graphDb)
        self.alpha, self.node_amount, self.pfail, self.delta, self.adjM, self.dir_db = alpha, node_amount, pfail, delta, adjM, dir_db
        self.seed = seed
        self.ppr = None
        self.topk_nodeIds = None
        self._walk_counter = 0
        self.preprocessing_dirName = os.path.join("MC_ppr_results", dir_db)
        self.stats = None

    def _conf(self):
        return {"delta": self.delta, "pfail": self.pfail}

    def random_walk(self, nodeId_start):
        """J/Monte_Carlo.java:60-94 (one walk; walk index advances per call, Philox-keyed)."""
        j = self._walk_counter
        self._walk_counter += 1
        return int(self.adjM.walk_terminals(int(nodeId_start), int(nodeId_start), j, 1, self.seed, 0, self.alpha, False)[0])

    def random_walk_no_zero_hop(self, nodeId_start):
        """J/Monte_Carlo.java:96-133."""
        j = self._walk_counter
        self._walk_counter += 1
        return int(self.adjM.walk_terminals(int(nodeId_start), int(nodeId_start), j, 1, self.seed, 0, self.alpha, True)[0])

    def computeWholeGraphPPR(self, nodeId_start, epsilon):
        self.topk_nodeIds = None
        self.ppr, self.stats = self.adjM.monte_carlo(int(nodeId_start), self.alpha, float(epsilon), self.seed, conf=self._conf())
        self.seed += 1

    def getWholeGraphPPR(self):
        return PPRVector(self.ppr)

    def printWholeGraphResult(self):
        """J/Monte_Carlo.java:160-173."""
        self._print_dense("Monte-Carlo PPR:", self.ppr)

    def computeTopKPPR(self, nodeId_start, k, epsilon):
        self.computeWholeGraphPPR(nodeId_start, epsilon)

    def getTopKNodeIds(self, k):
        if self.topk_nodeIds is None:
            self.topk_nodeIds, _ = _topk_from_dense(self.adjM, self.ppr, k)
        return self.topk_nodeIds

    def printTopKResult(self, k):
        """J/Monte_Carlo.java:276-301."""
        ids = self.getTopKNodeIds(k)
        self._print_rows("\nMonte-Carlo-Top%d PPR:" % k, ids, self.ppr[ids], limit=k)

    def preprocessing(self, dummy, epsilon):
        self.preprocessing_dirName = os.path.join(self.preprocessing_dirName, repr(float(epsilon)))
        _prep_clean(self.preprocessing_dirName)
        for v in range(self.node_amount):
            self.computeWholeGraphPPR(v, epsilon)
            _prep_write(self.preprocessing_dirName, v, PPRVector(self.ppr))

    def readPreprocessedPPR(self, nodeId_start):
        self.ppr = _prep_read(self.preprocessing_dirName, nodeId_start, self.node_amount)

    def getPrepSize(self):
        return _dir_size(self.preprocessing_dirName)

    def deletePrepDir(self):
        shutil.rmtree(self.preprocessing_dirName, ignore_errors=True)


class Fora_Whole_Graph(Algo_Util):
    """J/Fora_Whole_Graph.java."""

    def __init__(self, alpha, rsum, pfail, delta, node_amount, rel_amount, graphDb, adjM, node_property="name",
                 dir_db="db", seed=1, flags=N.RESTART_AT_WALK_START):
        super().__init__(node_property, graphDb)
        self.alpha, self.rsum, self.pfail, self.delta = alpha, rsum, pfail, delta
        self.node_amount, self.rel_amount, self.adjM, self.dir_db = node_amount, rel_amount, adjM, dir_db
        self.seed, self.flags = seed, flags
        self.reserve = None
        self.preprocessing_dirName = os.path.join("FORA_ppr_results", dir_db)
        self.stats = None

    def computeWholeGraphPPR(self, nodeId_start, epsilon):
        self.reserve, self.stats = self.adjM.fora(int(nodeId_start), self.alpha, float(epsilon), self.seed, self.flags,
                                                  conf={"delta": self.delta, "pfail": self.pfail})
        self.seed += 1

    def getWholeGraphPPR(self):
        return PPRVector(self.reserve)

    def printWholeGraphResult(self):
        """J/Fora_Whole_Graph.java:55-68."""
        self._print_dense("Fora-Whole-Graph PPR:", self.reserve)

    def preprocessing(self, dummy, epsilon):
        self.preprocessing_dirName = os.path.join(self.preprocessing_dirName, repr(float(epsilon)))
        _prep_clean(self.preprocessing_dirName)
        for v in range(self.node_amount):
            self.computeWholeGraphPPR(v, epsilon)
            _prep_write(self.preprocessing_dirName, v, PPRVector(self.reserve))

    def readPreprocessedPPR(self, nodeId_start):
        self.reserve = _prep_read(self.preprocessing_dirName, nodeId_start, self.node_amount)

    def getPrepSize(self):
        return _dir_size(self.preprocessing_dirName)

    def deletePrepDir(self):
        shutil.rmtree(self.preprocessing_dirName, ignore_errors=True)


class Fora_Topk(Algo_Util):
    """J/Fora_Topk.java -- Topk_Util_Interface."""

    def __init__(self, alpha, rsum, pfail, delta, node_amount, rel_amount, graphDb, min_delta, k, adjM,
                 node_property="name", dir_db="db", seed=1, flags=N.RESTART_AT_WALK_START):
        super().__init__(node_property, graphDb)
        self.alpha, self.rsum, self.pfail, self.delta, self.min_delta, self.k = alpha, rsum, pfail, delta, min_delta, k
        self.node_amount, self.rel_amount, self.adjM, self.dir_db = node_amount, rel_amount, adjM, dir_db
        self.seed, self.flags = seed, flags
        self.topk_nodeIds = None
        self.topk_vals = None
        self.stats = None

    def computeTopKPPR(self, nodeId_start, dummy, eps):
        ids, vals, self.stats = self.adjM.fora_topk(int(nodeId_start), self.alpha, float(eps), self.k, self.seed, self.flags,
                                                    conf={"delta": self.delta, "pfail": self.pfail, "min_delta": self.min_delta})
        self.topk_nodeIds, self.topk_vals = ids, vals
        self.seed += 1

    def getTopKNodeIds(self, dummy=None):
        return self.topk_nodeIds

    def printTopKResult(self, dummy=None):
        """J/Fora_Topk.java:60-79: at most k lines although ties at the cut may have kept more ids."""
        self._print_rows("\nFora-Top%d PPR:" % self.k, self.topk_nodeIds, self.topk_vals, limit=self.k)


class Backward_Search:
    """J/Backward_Search.java."""

    def __init__(self, alpha, rmax, node_amount, graphDb, adjM):
        self.alpha, self.rmax, self.node_amount, self.adjM = alpha, rmax, node_amount, adjM
        self.reserve = None
        self.stats = None
        self.duration = 0

    def backward_search_whole_graph(self, nodeId_target):
        t0 = time.perf_counter_ns()
        self.reserve, self.stats = self.adjM.backward_search(int(nodeId_target), self.alpha, float(self.rmax))
        self.duration += time.perf_counter_ns() - t0

    def getDuration(self):
        """J/Backward_Search.java:104-106: nanoseconds spent in backward searches so far."""
        return self.duration

    def getReserve(self):
        return PPRVector(self.reserve)


class Base_Whole_Graph(Algo_Util):
    """J/Base_Whole_Graph.java -- All-Pair-Backward-Search ("BASE")."""

    def __init__(self, alpha, node_amount, graphDb, adjM, node_property="name", dir_db="db", write_files=True):
        super().__init__(node_property, graphDb)
        self.alpha, self.node_amount, self.adjM, self.dir_db = alpha, node_amount, adjM, dir_db
        self.write_files = write_files
        self.preprocessing_dirName = os.path.join("BASE_ppr_results", dir_db)
        self.ppr_all_pairs = None   # (src_ptr, tgt, val): inverted index v -> [(t, pi(v, t))]
        self.ppr_src = None
        self.topk_nodeIds = None
        self.stats = None

    def preprocessing(self, threshold, k):
        """J/Base_Whole_Graph.java:58-164.  k < 0: whole graph; k >= 0: top-k per source, sorted descending."""
        self.preprocessing_dirName = os.path.join(self.preprocessing_dirName, "%r_%d" % (float(threshold), int(k)))
        ptr, tgt, val, self.stats = self.adjM.backward_all_pairs(self.alpha, float(threshold), int(k))
        self.ppr_all_pairs = (ptr, tgt, val)
        if self.write_files:
            _prep_clean(self.preprocessing_dirName)
            for v in range(self.node_amount):  # only sources with at least one stored pair get a file
                lo, hi = ptr[v], ptr[v + 1]
                if hi > lo:
                    with open(os.path.join(self.preprocessing_dirName, "%d.txt" % v), "w") as f:
                        for t, x in zip(tgt[lo:hi], val[lo:hi]):
                            f.write("%d\t%s\n" % (t, _java_double(x)))

    def computeWholeGraphPPR(self, nodeId_start, dummy=None):
        """J/Base_Whole_Graph.java:167-186: the query is a lookup of the preprocessed row."""
        v = int(nodeId_start)
        self.topk_nodeIds = None
        if self.write_files:
            self.ppr_src = _prep_read_ordered(self.preprocessing_dirName, v)
        else:
            ptr, tgt, val = self.ppr_all_pairs
            self.ppr_src = dict(zip(map(int, tgt[ptr[v]:ptr[v + 1]]), map(float, val[ptr[v]:ptr[v + 1]])))

    def readPreprocessedPPR(self, nodeId_start):
        self.computeWholeGraphPPR(nodeId_start)

    def getWholeGraphPPR(self):
        return self.ppr_src

    def printWholeGraphResult(self):
        """J/Base_Whole_Graph.java:188-193: the stored row in stored order, not re-sorted."""
        print("\nBase-Whole-Graph PPR:")
        for t, x in self.ppr_src.items():
            print("@" + self.getNodeName(t) + "\t" + _java_double(x))

    def computeTopKPPR(self, nodeId_start, k, dummy=None):
        self.computeWholeGraphPPR(nodeId_start)

    def getTopKNodeIds(self, k):
        if self.topk_nodeIds is None:  # file rows are already sorted when preprocessing ran with k >= 0
            self.topk_nodeIds = np.fromiter(self.ppr_src.keys(), dtype=np.int32, count=len(self.ppr_src))
        return self.topk_nodeIds

    def printTopKResult(self, k):
        """J/Base_Whole_Graph.java:219-236."""
        self._print_rows("\nBASE-Top%d PPR:" % k, list(self.ppr_src.keys()), list(self.ppr_src.values()), limit=k)

    def getPrepSize(self):
        return _dir_size(self.preprocessing_dirName)

    def deletePrepDir(self):
        shutil.rmtree(self.preprocessing_dirName, ignore_errors=True)


class Power_Method(Algo_Util):
    """J/Power_Method.java -- the ground truth of the harness (100 synchronous iterations)."""

    def __init__(self, graphDb, alpha, adjM, node_property="name", dir_db="db"):
        super().__init__(node_property, graphDb)
        self.alpha, self.adjM = alpha, adjM
        self.reserve = None
        self.topk_res = None
        self.topk_nodeIds = None

    def computeWholeGraphPPR(self, nodeId_start, dummy=None):
        self.topk_nodeIds = self.topk_res = None
        self.reserve = self.adjM.power_iteration(int(nodeId_start), self.alpha, 100)

    def getWholeGraphPPR(self):
        return PPRVector(self.reserve)

    def computeTopKPPR(self, nodeId_start, k, dummy=None):
        self.computeWholeGraphPPR(nodeId_start)
        ids, vals = _topk_from_dense(self.adjM, self.reserve, k)
        self.topk_nodeIds, self.topk_res = ids, dict(zip(map(int, ids), map(float, vals)))

    def getTopK(self, k):
        return self.topk_res

    def getTopKNodeIds(self, k):
        return self.topk_nodeIds

    def printWholeGraphResult(self):
        """J/Power_Method.java:103-116."""
        self._print_dense("Exact-PPR:", self.reserve)

    def printTopKResult(self, k):
        """J/Power_Method.java:160-184."""
        self._print_rows("\nExact-Top%d PPR:" % k, self.topk_nodeIds, self.reserve[self.topk_nodeIds], limit=k)


class Algo_Conf:
    """J/Algo_Conf.java: parameter derivation + factories."""

    def __init__(self, alpha, node_property="name"):
        self.alpha, self.node_property = alpha, node_property

    def set_conf_power_method(self, graphDb, adjM, dir_db):
        return Power_Method(graphDb, self.alpha, adjM, self.node_property, dir_db)

    def set_conf_mc(self, node_amount, rel_amount, graphDb, adjM, dir_db):
        delta = pfail = 1.0 / float(node_amount)
        return Monte_Carlo(self.alpha, node_amount, graphDb, pfail, delta, adjM, self.node_property, dir_db)

    def set_conf_base_whole_graph(self, node_amount, rel_amount, graphDb, adjM, dir_db):
        return Base_Whole_Graph(self.alpha, node_amount, graphDb, adjM, self.node_property, dir_db)

    def set_conf_fora_whole_graph(self, node_amount, rel_amount, graphDb, adjM, dir_db):
        delta = pfail = 1.0 / float(node_amount)
        return Fora_Whole_Graph(self.alpha, 1.0, pfail, delta, node_amount, rel_amount, graphDb, adjM, self.node_property, dir_db)

    def set_conf_fwdpush(self, node_amount, rel_amount, graphDb, adjM, dir_db):
        return Forward_Push(self.alpha, 1.0, node_amount, graphDb, adjM, self.node_property, dir_db)

    def set_conf_fora_topk(self, node_amount, rel_amount, k, graphDb, adjM, dir_db):
        min_delta = 1.0 / float(node_amount)
        delta = 1.0 / float(k)
        q = node_amount // k  # integer division, J/Algo_Conf.java:76
        lg = math.log(q) if q > 0 else float("-inf")
        pfail = (1.0 / float(node_amount) / float(node_amount) / lg) if lg != 0 else float("inf")
        return Fora_Topk(self.alpha, 1.0, pfail, delta, node_amount, rel_amount, graphDb, min_delta, k, adjM,
                         self.node_property, dir_db)


# ------------------------------------------------------------------ preprocessing result store (row N2 of SURVEY 8f)
def _java_double(x):
    """Double.toString-compatible enough for round trips: shortest repr, Java's E notation."""
    from decimal import Decimal
    x = float(x)
    if x != 0 and (abs(x) < 1e-3 or abs(x) >= 1e7) and math.isfinite(x):
        sign, digits, exp = Decimal(repr(x)).as_tuple()   # shortest round-trip digits
        digits = list(digits)
        while len(digits) > 1 and digits[-1] == 0:
            digits.pop()
            exp += 1
        e10 = exp + len(digits) - 1
        frac = "".join(map(str, digits[1:])) or "0"
        return "%s%d.%sE%d" % ("-" if sign else "", digits[0], frac, e10)
    s = repr(x)
    return s if ("." in s or "e" in s or "n" in s) else s + ".0"


def _prep_clean(d):
    if os.path.isdir(d):
        shutil.rmtree(d)
    os.makedirs(d, exist_ok=True)


def _prep_write(d, v, vec):
    with open(os.path.join(d, "%d.txt" % v), "w") as f:
        ks, vs = vec.keys_array(), vec.values_array()
        for k, x in zip(ks, vs):
            f.write("%d\t%s\n" % (k, _java_double(x)))


def _prep_read(d, v, n):
    out = np.zeros(n)
    with open(os.path.join(d, "%d.txt" % int(v))) as f:
        for line in f:
            k, x = line.rstrip("\n").split("\t")
            out[int(k)] = float(x)
    return out


def _prep_read_ordered(d, v):
    out = {}
    path = os.path.join(d, "%d.txt" % int(v))
    if not os.path.exists(path):  # the Java prints the IOException and leaves the map empty
        return out
    with open(path) as f:
        for line in f:
            k, x = line.rstrip("\n").split("\t")
            out[int(k)] = float(x)
    return out


def _dir_size(d):
    total = 0
    for root, _, files in os.walk(d):
        for fn in files:
            total += os.path.getsize(os.path.join(root, fn))
    return total
