package joezie.fora_neo4j;

import org.neo4j.graphalgo.api.Graph;
import org.neo4j.graphdb.GraphDatabaseService;

/**
 * Drop-in replacement of the reference class of the same name (same constructor, interfaces and public methods).
 * The balance loop, the forward push and the random walks of computeWholeGraphPPR (J/Fora_Whole_Graph.java:82-146)
 * are ONE call into libpprb200.so (ppr_fora).  NOT COMPILED HERE (no JDK in the image); see INTEGRATION.md.
 */
public class Fora_Whole_Graph extends Gpu_Algo_Base implements Whole_Graph_Util_Interface, Preprocessing_Interface {
    private final int rel_amount;
    private final Double rsum, pfail, delta;

    public Fora_Whole_Graph(Double alpha, Double rsum, Double pfail, Double delta, int node_amount, int rel_amount,
                            GraphDatabaseService graphDb, Graph adjM, String node_property, String dir_db) {
        super(node_property, alpha, node_amount, graphDb, adjM, "FORA_ppr_results/" + dir_db);
        this.rsum = rsum; this.pfail = pfail; this.delta = delta; this.rel_amount = rel_amount;
    }

    @Override protected String wholeGraphHeader() { return "Fora-Whole-Graph PPR:"; }
    @Override protected String topkHeaderPrefix() { return "Fora-Whole-Graph-Top"; }   // the reference class has no top-k print

    @Override protected void query(Long nodeId_start, Object epsilon) throws Throwable {
        reserve = gpu.fora(nodeId_start, alpha, (Double) epsilon, delta, pfail, seed++);
    }
}
