package joezie.fora_neo4j;

import java.util.ArrayList;
import java.util.Comparator;
import java.util.HashMap;
import java.util.List;
import java.util.Map;
import java.util.Queue;
import java.util.concurrent.ConcurrentLinkedQueue;

import org.neo4j.graphalgo.api.Graph;
import org.neo4j.graphdb.GraphDatabaseService;

/**
 * Drop-in replacement of the reference class: the FIFO loop of computeWholeGraphPPR (J/Forward_Push.java:63-142) is one
 * ppr_forward_push call (frontier-synchronous push on the GPU; same exit condition r(v)/d(v) < rmax for every v).
 * forward_push_topk (:144-250) has no counterpart here: its only caller, Fora_Topk, runs its rounds inside ppr_fora_topk.
 * NOT COMPILED HERE (no JDK in the image).
 */
public class Forward_Push extends Gpu_Algo_Base implements Whole_Graph_Util_Interface, Preprocessing_Interface, Topk_Util_Interface {
    private HashMap<Long, Double> residue = new HashMap<>();
    private Double rsum;

    public Forward_Push(Double alpha, Double rsum, int node_amount, GraphDatabaseService graphDb, Graph adjM, String node_property, String dir_db) {
        super(node_property, alpha, node_amount, graphDb, adjM, "FWP_ppr_results/" + dir_db);
        this.rsum = rsum;
    }

    @Override protected void query(Long nodeId_start, Object rmax) throws Throwable {
        double[] rs = new double[1];
        HashMap<Long, Double>[] out = gpu.forwardPush(nodeId_start, alpha, (Double) rmax, rs);
        reserve = out[0];
        residue = out[1];
        rsum = rs[0];
    }

    @Override protected String wholeGraphHeader() { return "Forward-Push Reserve(pi) & Residue(r):"; }
    @Override protected String topkHeaderPrefix() { return "Forward-Push-Top"; }

    @Override public void printWholeGraphResult() {   // J/Forward_Push.java:268-281: a third column with the residue
        List<Map.Entry<Long, Double>> rows = new ArrayList<>(reserve.entrySet());
        rows.sort(Map.Entry.<Long, Double>comparingByValue(Comparator.reverseOrder()));
        System.out.println(wholeGraphHeader());
        for (Map.Entry<Long, Double> e : rows)
            System.out.println("@" + getNodeName(e.getKey(), graphDb) + '\t' + e.getValue() + '\t' + residue.get(e.getKey()));
    }

    public Double getUpdatedRsum() { return rsum; }
    public HashMap<Long, Double> getResidueCopy() { return new HashMap<>(residue); }
    public HashMap<Long, Double> getReserveCopy() { return new HashMap<>(reserve); }
    public Queue<Long> getNextQueueCopy() { return new ConcurrentLinkedQueue<>(); }   // the GPU push leaves nothing parked
}
