package joezie.fora_neo4j;

import java.util.HashMap;
import java.util.Map;
import java.util.Vector;

import org.neo4j.graphalgo.api.Graph;
import org.neo4j.graphdb.GraphDatabaseService;

/**
 * Drop-in replacement of the reference class: the delta-halving rounds of computeTopKPPR (J/Fora_Topk.java:102-184:
 * incremental push, walks, k-th estimate test) run inside ppr_fora_topk; k, delta, pfail and min_delta are the
 * values Algo_Conf.set_conf_fora_topk derived (J/Algo_Conf.java:71-81).  NOT COMPILED HERE (no JDK in the image).
 */
public class Fora_Topk extends Gpu_Algo_Base implements Topk_Util_Interface {
    private final int rel_amount, k;
    private final Double rsum, pfail, delta, min_delta;

    public Fora_Topk(Double alpha, Double rsum, Double pfail, Double delta, int node_amount, int rel_amount,
                     GraphDatabaseService graphDb, Double min_delta, int k, Graph adjM, String node_property, String dir_db) {
        super(node_property, alpha, node_amount, graphDb, adjM, "FORA_TOPK_ppr_results/" + dir_db);
        this.rsum = rsum; this.pfail = pfail; this.delta = delta; this.rel_amount = rel_amount; this.k = k; this.min_delta = min_delta;
    }

    @Override protected void query(Long nodeId_start, Object eps) throws Throwable {
        topk_res.clear();
        topk_nodeIds.clear();
        for (Map.Entry<Long, Double> e : gpu.foraTopkEntries(nodeId_start, alpha, (Double) eps, k, delta, pfail, min_delta, seed++)) {
            topk_res.put(e.getKey(), e.getValue());
            topk_nodeIds.add(e.getKey());
        }
        reserve = new HashMap<>(topk_res);
    }

    @Override public void computeTopKPPR(Long nodeId_start, int dummy, Object eps) {
        try { query(nodeId_start, eps); } catch (Throwable t) { t.printStackTrace(); }
    }
    @Override public Vector<Long> getTopKNodeIds(int dummy) { return topk_nodeIds; }
    @Override protected String wholeGraphHeader() { return "Fora-Top" + k + " PPR:"; }
    @Override protected String topkHeaderPrefix() { return "Fora-Top"; }
    @Override public void printTopKResult(int dummy) { printSorted("\nFora-Top" + k + " PPR:", topk_res, k); }   // J/Fora_Topk.java:60-79
}
