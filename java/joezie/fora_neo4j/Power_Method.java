package joezie.fora_neo4j;

import java.util.HashMap;

import org.neo4j.graphalgo.api.Graph;
import org.neo4j.graphdb.GraphDatabaseService;

/**
 * Drop-in replacement of the reference class (ground truth): the 100 Jacobi rounds of computeWholeGraphPPR
 * (J/Power_Method.java:44-101) are one ppr_power_iteration call (double arithmetic, deterministic pull over the CSC).
 * NOT COMPILED HERE (no JDK in the image).
 */
public class Power_Method extends Gpu_Algo_Base implements Whole_Graph_Util_Interface, Topk_Util_Interface {
    private static final int ITERATIONS = 100;   // J/Power_Method.java:57

    public Power_Method(GraphDatabaseService graphDb, Double alpha, Graph adjM, String node_property, String dir_db) {
        super(node_property, alpha, (int) adjM.nodeCount(), graphDb, adjM, "PM_ppr_results/" + dir_db);
    }

    @Override protected void query(Long nodeId_start, Object dummy) throws Throwable {
        reserve = gpu.powerMethod(nodeId_start, alpha, ITERATIONS);
    }

    @Override protected String wholeGraphHeader() { return "Exact-PPR:"; }
    @Override protected String topkHeaderPrefix() { return "Exact-Top"; }

    public HashMap<Long, Double> getTopK(int k) { return topk_res; }
}
