"""The reference's driver and perf harness (`PPR`, `Gen_Util`) re-hosted over the B200 library.

Mirrors /root/reference/src/main/java/joezie/fora_neo4j/PPR.java (options, defaults, lifecycle) and
Gen_Util.java (query sampling, timing loops, error metrics, parameter grids, result-file format), so
`python -m <package>.harness -alpha 0.15 -eps 0.5 -query 50 -k 10 -db tests/golden/got` produces the same
`<db>_AlgoPerfResults.txt` rows as `java PPR ...`.  The Neo4j competitor (NEO4J_METHOD) is out of
scope and is skipped.  `-db` names either a Neo4j 3.x store directory (its node and relationship stores are read natively,
`graphio.load_neo4j_store`) or a directory holding the Neo4j import CSVs (`*_Nodes.csv` / `*_Rels.csv`).
"""
import argparse
import glob
import math
import os
import time

import numpy as np

from . import _native as N
from . import algo as A
from . import graphio

MC, FWDPUSH, FORA_WHOLE_GRAPH, FORA_TOPK, POWER_METHOD, BASE_WHOLE_GRAPH, NEO4J_METHOD = (
    "MC", "FWDPUSH", "FORA_WHOLE_GRAPH", "FORA_TOPK", "POWER_METHOD", "BASE_WHOLE_GRAPH", "NEO4J_METHOD")
WHOLE_GRAPH, TOPK = "WHOLE_GRAPH", "TOPK"
PRECISION, NDCG, MAX_ERR = "PRECISION", "NDCG", "MAX_ERR"


def compute_error(algo, pm, err_type, k):
    """Gen_Util.computeError, J/Gen_Util.java:259-326."""
    if err_type in (PRECISION, NDCG):
        gnd_ids = pm.getTopKNodeIds(k)
        algo_ids = algo.getTopKNodeIds(k)
        if gnd_ids is None or algo_ids is None:
            return None
        gnd_set = set(int(x) for x in gnd_ids)
        if err_type == PRECISION:
            hit = float(sum(1 for x in algo_ids if int(x) in gnd_set))
            return hit / float(len(gnd_ids))
        gnd = pm.getTopK(k)
        zk = dcg = 0.0   # plain left-to-right accumulation as in the Java (Python's sum() compensates and differs in the last bit)
        for i, v in enumerate(gnd_ids, 1):
            zk += (math.pow(2.0, gnd[int(v)]) - 1.0) / math.log(i + 1.0) / math.log(2.0)
        for i, v in enumerate(algo_ids, 1):
            dcg += (math.pow(2.0, gnd.get(int(v), 0.0)) - 1.0) / math.log(i + 1.0) / math.log(2.0)
        return dcg / zk
    gnd = pm.getWholeGraphPPR()
    est = algo.getWholeGraphPPR()
    if isinstance(gnd, A.PPRVector) and isinstance(est, A.PPRVector):
        ks = gnd.keys_array()
        return float(np.max(np.abs(est.dense[ks] - gnd.dense[ks]))) if ks.size else 0.0
    return max((abs((est.get(v) or 0.0) - x) for v, x in gnd.items()), default=0.0)


class Gen_Util:
    """J/Gen_Util.java."""

    def __init__(self):
        self.ctx = None
        self.adjM = None
        self.graphDb = None
        self.dir_db = "got.db"
        self.alpha, self.epsilon = 0.15, 0.5
        self.node_property, self.label_type, self.rel_type = "name", "Person", "Relation"
        self.device_metrics = False   # True: computeError is evaluated in HBM (ppr_metrics_keep / ppr_metrics_compute) instead of on host vectors
        self.node_amount = self.rel_amount = 0
        self.fw = None
        self.rng = np.random.default_rng(1234)

    def newAlgoObj(self, algo, k):
        ac = A.Algo_Conf(self.alpha, self.node_property)
        if algo == MC:
            return ac.set_conf_mc(self.node_amount, self.rel_amount, self.graphDb, self.adjM, self.dir_db)
        if algo == POWER_METHOD:
            return ac.set_conf_power_method(self.graphDb, self.adjM, self.dir_db)
        if algo == BASE_WHOLE_GRAPH:
            return ac.set_conf_base_whole_graph(self.node_amount, self.rel_amount, self.graphDb, self.adjM, self.dir_db)
        if algo == FORA_WHOLE_GRAPH:
            return ac.set_conf_fora_whole_graph(self.node_amount, self.rel_amount, self.graphDb, self.adjM, self.dir_db)
        if algo == FWDPUSH:
            return ac.set_conf_fwdpush(self.node_amount, self.rel_amount, self.graphDb, self.adjM, self.dir_db)
        if algo == FORA_TOPK:
            return ac.set_conf_fora_topk(self.node_amount, self.rel_amount, k, self.graphDb, self.adjM, self.dir_db)
        print("Creating a new object of %s failed!" % algo)
        return None

    def getQueryNodes(self, query_num):
        return [int(x) for x in self.rng.integers(0, self.node_amount, size=query_num)]  # J/Gen_Util.java:99-107

    def algo_perf_test(self, algoType, query_num, k, param, threshold, to_be_preprocessed, testType):
        """J/Gen_Util.java:109-257.  Returns the row written to the results file (dict) for callers/tests."""
        fw = self.fw
        queryNodes = self.getQueryNodes(query_num)
        row = {"algo": algoType, "param": param, "k": k}
        if testType == TOPK:
            topk_t = self.newAlgoObj(algoType, k)
            pm_t = self.newAlgoObj(POWER_METHOD, -1)
            duration = 0
            sum_precision = sum_ndcg = 0.0
            if algoType == BASE_WHOLE_GRAPH:
                t0 = time.perf_counter_ns()
                topk_t.preprocessing(threshold, k)
                duration_prep = time.perf_counter_ns() - t0
                fw.write("%s,%d,%d,%d," % (A._java_double(threshold), k, duration_prep // 1000000, topk_t.getPrepSize()))
                row["prep_ms"] = duration_prep / 1e6
            else:
                fw.write("%s,%d," % (_param_str(param), k))
            for q in queryNodes:
                t0 = time.perf_counter_ns()
                topk_t.computeTopKPPR(q, k, param)
                duration += time.perf_counter_ns() - t0
                on_dev = self.device_metrics and algoType != BASE_WHOLE_GRAPH
                if on_dev:
                    self.adjM.metrics_keep(1)
                pm_t.computeTopKPPR(q, k, None)
                if on_dev:      # J/Gen_Util.java:264-300 on the device: both vectors are still in HBM
                    self.adjM.metrics_keep(0)
                    m = self.adjM.metrics_compute(k)
                    p, g = m["precision"], m["ndcg"]
                else:
                    p = compute_error(topk_t, pm_t, PRECISION, k)
                    g = compute_error(topk_t, pm_t, NDCG, k)
                sum_precision += p or 0.0
                sum_ndcg += g or 0.0
            avg = duration // query_num
            row.update(avg_ms=avg / 1e6, precision=sum_precision / query_num, ndcg=sum_ndcg / query_num)
            fw.write("%d,%s,%s\n" % (avg // 1000000, A._java_double(row["precision"]), A._java_double(row["ndcg"])))
        elif testType == WHOLE_GRAPH:
            algo_t = self.newAlgoObj(algoType, -1)
            pm_t = self.newAlgoObj(POWER_METHOD, -1)
            duration = 0
            sum_max_err = 0.0
            fw.write("%s," % _param_str(param))
            prep = to_be_preprocessed or algoType == BASE_WHOLE_GRAPH
            if prep:
                t0 = time.perf_counter_ns()
                algo_t.preprocessing(threshold, param)
                duration_prep = time.perf_counter_ns() - t0
                fw.write("%s,%d,%d," % (A._java_double(threshold), duration_prep // 1000000, algo_t.getPrepSize()))
                row["prep_ms"] = duration_prep / 1e6
            for q in queryNodes:
                t0 = time.perf_counter_ns()
                if to_be_preprocessed:
                    algo_t.readPreprocessedPPR(q)
                else:
                    algo_t.computeWholeGraphPPR(q, param)
                duration += time.perf_counter_ns() - t0
                on_dev = self.device_metrics and not prep
                if on_dev:
                    self.adjM.metrics_keep(1)
                pm_t.computeWholeGraphPPR(q, None)
                if on_dev:      # J/Gen_Util.java:306-321 on the device
                    self.adjM.metrics_keep(0)
                    sum_max_err += self.adjM.metrics_compute(0)["max_err"]
                else:
                    sum_max_err += compute_error(algo_t, pm_t, MAX_ERR, -1)
            avg = duration // query_num
            row.update(avg_ms=avg / 1e6, max_err=sum_max_err / query_num)
            if not to_be_preprocessed:
                fw.write("%d," % (avg // 1000000))
            fw.write("%s\n" % A._java_double(row["max_err"]))
            if to_be_preprocessed and algoType != BASE_WHOLE_GRAPH:
                algo_t.deletePrepDir()
        else:
            print("Error: Test type should be WHOLE_GRAPH or TOPK!")
            return None
        return row

    # parameter grids of the active test set (GOT), J/Gen_Util.java:454-478
    GRID = {
        BASE_WHOLE_GRAPH: [0.001, 5.0e-4, 5.0e-5, 1.0e-6, 5.0e-7],
        FORA_WHOLE_GRAPH: [10.0, 5.0, 0.5, 0.1, 0.05],
        FORA_TOPK: [10.0, 5.0, 0.5, 0.1, 0.05],
        MC: [1.0, 0.5, 0.3, 0.1, 0.05],
        FWDPUSH: [1.0e-4, 1.0e-5, 1.0e-6, 1.0e-7, 1.0e-8],
    }

    def algo_perf_batch_test(self, query_num, k, tests=(1, 2, 3), out_dir="."):
        """J/Gen_Util.java:328-647 (NEO4J_METHOD rows skipped)."""
        path = os.path.join(out_dir, self.dir_db + "_AlgoPerfResults.txt")
        rows = []
        with open(path, "a") as fw:
            self.fw = fw
            fw.write(time.strftime("%Y-%m-%d %H:%M:%S") + "\n")
            whole = [FORA_WHOLE_GRAPH, FWDPUSH, NEO4J_METHOD, MC, BASE_WHOLE_GRAPH]
            topk = [FORA_TOPK, FWDPUSH, NEO4J_METHOD, MC, BASE_WHOLE_GRAPH]
            if 1 in tests:
                fw.write("\nTest 1. Whole-Graph test\n")
                for i, algo in enumerate(whole, 1):
                    fw.write("\n1.%d %s\n" % (i, algo))
                    if algo == NEO4J_METHOD:
                        continue
                    for p in self.GRID[algo]:
                        if algo == BASE_WHOLE_GRAPH:
                            rows.append(self.algo_perf_test(algo, query_num, -1, -1, p, False, WHOLE_GRAPH))
                        else:
                            rows.append(self.algo_perf_test(algo, query_num, -1, p, -1.0, False, WHOLE_GRAPH))
            if 2 in tests:
                fw.write("\nTest 2. Top-k test\n")
                for i, algo in enumerate(topk, 1):
                    fw.write("\n2.%d %s\n" % (i, algo))
                    if algo == NEO4J_METHOD:
                        continue
                    for p in self.GRID[algo]:
                        if algo == BASE_WHOLE_GRAPH:
                            rows.append(self.algo_perf_test(algo, query_num, k, -1, p, False, TOPK))
                        else:
                            rows.append(self.algo_perf_test(algo, query_num, k, p, -1.0, False, TOPK))
            if 3 in tests:
                fw.write("\nTest 3. Preprocessing test\n")
                for i, algo in enumerate(whole, 1):
                    fw.write("\n3.%d %s\n" % (i, algo))
                    if algo == NEO4J_METHOD:
                        continue
                    for p in self.GRID[algo]:
                        if algo == BASE_WHOLE_GRAPH:
                            rows.append(self.algo_perf_test(algo, query_num, -1, -1, p, True, WHOLE_GRAPH))
                        else:
                            rows.append(self.algo_perf_test(algo, query_num, -1, p, -1.0, True, WHOLE_GRAPH))
            self.fw = None
        return path, rows


def _param_str(p):
    return A._java_double(p) if isinstance(p, float) else str(p)


class PPR(Gen_Util):
    """J/PPR.java: the driver."""

    def createDb(self, databaseDirectory, device=0):
        print("\nCreating or opening database %s..." % os.path.basename(os.path.normpath(databaseDirectory)))
        self.dir_db = os.path.basename(os.path.normpath(databaseDirectory))
        self.db_path = databaseDirectory
        self.ctx = N.Context(device)
        print("Database created or opened")

    def shutDown(self):
        print("\nShutting down database ...")
        if self.adjM is not None:
            self.adjM.close()
        if self.ctx is not None:
            self.ctx.close()

    def setupAdjMatrix(self):
        """J/PPR.java:136-152: one-time export to the device-resident CSR + CSC."""
        print("\nLoading graph from database...")
        t0 = time.perf_counter_ns()
        nodes = sorted(glob.glob(os.path.join(self.db_path, "*Nodes.csv")))
        rels = sorted(glob.glob(os.path.join(self.db_path, "*Rels.csv")))
        if os.path.exists(os.path.join(self.db_path, "neostore.nodestore.db")):   # a Neo4j 3.x store directory (target/got.db)
            row_ptr, col_idx, self.original_ids = graphio.load_neo4j_store(self.db_path)
            self.names = [str(i) for i in self.original_ids]
        elif nodes and rels:
            row_ptr, col_idx, self.names = graphio.load_neo4j_csv(nodes[0], rels[0])
        elif os.path.isfile(self.db_path):                                        # a plain edge list (SNAP text / edges.csv)
            row_ptr, col_idx, self.original_ids = graphio.load_edge_list(
                self.db_path, undirected=getattr(self, "undirected", False), delimiter="," if self.db_path.endswith(".csv") else None)
            self.names = [str(i) for i in self.original_ids]
        else:
            raise FileNotFoundError("neither a Neo4j store nor *_Nodes.csv / *_Rels.csv under %s" % self.db_path)
        self.adjM = N.Graph(self.ctx, row_ptr, col_idx)
        self.graphDb = self.names          # all the algorithm classes ask of graphDb is getNodeName (J/Algo_Util.java:20-30)
        print("\nFinish graph loading in %d(ms)" % ((time.perf_counter_ns() - t0) // 1000000))

    def set_configuration(self, alpha, epsilon, node_property, label_type, rel_type):
        self.alpha, self.epsilon = alpha, epsilon
        self.node_property, self.label_type, self.rel_type = node_property, label_type, rel_type
        self.node_amount = self.adjM.nodeCount()
        self.rel_amount = self.adjM.relCount()

    def create_index(self):
        print("\nCreating index ...\nIndex created")  # Neo4j schema index: no counterpart (names stay on the host)


def main(argv=None):
    """J/PPR.java:154-200 -- same option names and defaults."""
    ap = argparse.ArgumentParser(prog="PPR", add_help=False)
    ap.add_argument("-alpha", "--alpha", type=float, default=0.15)
    ap.add_argument("-eps", "--epsilon", type=float, default=0.5)
    ap.add_argument("-query", "--query", type=int, default=50)
    ap.add_argument("-k", "--k", type=int, default=10)
    ap.add_argument("-node", "--node_property", default="name")
    ap.add_argument("-label", "--label_type", default="Person")
    ap.add_argument("-rel", "--rel_type", default="Relation")
    ap.add_argument("-db", "--db_dir", default=os.path.dirname(graphio.got_paths()[0]))
    ap.add_argument("-help", "--help", action="help")
    ap.add_argument("--tests", default="1,2,3", help="subset of the three harness tests to run")
    ap.add_argument("--out", default=".")
    ap.add_argument("--undirected", action="store_true", help="-db names a plain edge-list file: store both directions of every pair")
    a = ap.parse_args(argv)
    t = PPR()
    t.undirected = a.undirected
    try:
        t.createDb(a.db_dir)
        t.setupAdjMatrix()
        t.set_configuration(a.alpha, a.epsilon, a.node_property, a.label_type, a.rel_type)
        t.create_index()
        path, _ = t.algo_perf_batch_test(a.query, a.k, tuple(int(x) for x in a.tests.split(",")), a.out)
        print("results appended to", path)
        t.shutDown()
    except Exception as e:  # J/PPR.java:196-199
        print("Algo performance batch test failed!")
        raise e


PPR.main = staticmethod(main)


if __name__ == "__main__":
    main()
