package joezie.fora_neo4j;

import org.neo4j.graphalgo.api.Graph;
import org.neo4j.graphdb.GraphDatabaseService;

/**
 * Drop-in replacement of the reference class: the omega = 3 L / eps^2 / delta walks of computeWholeGraphPPR
 * (J/Monte_Carlo.java:136-158) run in one ppr_monte_carlo call (one walk per lane slot, Philox keyed by (seed, walk id)).
 * random_walk / random_walk_no_zero_hop (:60-133) stay available for callers that want single terminals
 * (ppr_walk_terminals, same draw order).  NOT COMPILED HERE (no JDK in the image).
 */
public class Monte_Carlo extends Gpu_Algo_Base implements Whole_Graph_Util_Interface, Preprocessing_Interface, Topk_Util_Interface {
    private final Double pfail, delta;
    private long walk_id = 0;

    public Monte_Carlo(Double alpha, int node_amount, GraphDatabaseService graphDb, Double pfail, Double delta,
                       Graph adjM, String node_property, String dir_db) {
        super(node_property, alpha, node_amount, graphDb, adjM, "MC_ppr_results/" + dir_db);
        this.pfail = pfail; this.delta = delta;
    }

    @Override protected void query(Long nodeId_start, Object epsilon) throws Throwable {
        reserve = gpu.monteCarlo(nodeId_start, alpha, (Double) epsilon, delta, pfail, seed++);
    }

    @Override protected String wholeGraphHeader() { return "Monte-Carlo PPR:"; }
    @Override protected String topkHeaderPrefix() { return "Monte-Carlo-Top"; }

    public Long random_walk(Long nodeId_start) { return walk(nodeId_start, false); }
    public Long random_walk_no_zero_hop(Long nodeId_start) { return walk(nodeId_start, true); }

    private Long walk(Long nodeId_start, boolean no_zero_hop) {
        try { return gpu.walkTerminal(nodeId_start, walk_id++, seed, alpha, no_zero_hop); }
        catch (Throwable t) { t.printStackTrace(); return nodeId_start; }
    }
}
