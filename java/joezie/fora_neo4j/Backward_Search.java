package joezie.fora_neo4j;

import java.util.HashMap;

import org.neo4j.graphalgo.api.Graph;
import org.neo4j.graphdb.GraphDatabaseService;

/**
 * Drop-in replacement of the reference class: the reverse-push loop of backward_search_whole_graph
 * (J/Backward_Search.java:38-100) is one ppr_backward_search call.  Base_Whole_Graph does not loop over this class any
 * more (all targets go through ppr_backward_all_pairs); it stays for single-target callers.
 * NOT COMPILED HERE (no JDK in the image).
 */
public class Backward_Search {
    private final PprB200 gpu = PPR.gpu();
    private final Double alpha, rmax;
    private HashMap<Long, Double> reserve = new HashMap<>();
    private long duration = 0;

    public Backward_Search(Double alpha, Double rmax, int node_amount, GraphDatabaseService graphDb, Graph adjM) {
        this.alpha = alpha; this.rmax = rmax;
    }

    public void backward_search_whole_graph(Long nodeId_target) {
        long t0 = System.nanoTime();
        try { reserve = gpu.backwardSearch(nodeId_target, alpha, rmax); }
        catch (Throwable t) { t.printStackTrace(); }
        duration += System.nanoTime() - t0;
    }

    public HashMap<Long, Double> getReserve() { return reserve; }
    public long getDuration() { return duration; }
}
