package joezie.fora_neo4j;

import java.util.HashMap;

/**
 * Drop-in replacement of the reference class of the same name: same constructor, same interfaces, same
 * public methods; the push/walk loops of computeWholeGraphPPR (reference lines 82-146) become ONE call
 * into libpprb200.so.  NOT COMPILED HERE (no JDK in the image); see INTEGRATION.md.
 * The other algorithm classes change the same way (one delegating line each).
 */
public class Fora_Whole_Graph extends Algo_Util implements Whole_Graph_Util_Interface, Preprocessing_Interface {
    private final PprB200 gpu;             // shared by all algorithm objects of a run (created in PPR.setupAdjMatrix)
    private HashMap<Long, Double> reserve = new HashMap<>();
    private final int node_amount, rel_amount;
    private final Double alpha, rsum, pfail, delta;
    private long seed = 1;

    public Fora_Whole_Graph(Double alpha, Double rsum, Double pfail, Double delta, int node_amount, int rel_amount,
                            org.neo4j.graphdb.GraphDatabaseService graphDb, org.neo4j.graphalgo.api.Graph adjM,
                            String node_property, String dir_db) {
        super(node_property);
        this.alpha = alpha; this.rsum = rsum; this.pfail = pfail; this.delta = delta;
        this.node_amount = node_amount; this.rel_amount = rel_amount;
        this.gpu = PPR.gpu();              // the context that exported adjM
    }

    @Override public void computeWholeGraphPPR(Long nodeId_start, Object epsilon) {
        try { reserve = gpu.fora(nodeId_start, alpha, (Double) epsilon, delta, pfail, seed++); }
        catch (Throwable t) { t.printStackTrace(); }   // the reference swallows failures as well
    }

    @Override public HashMap<Long, Double> getWholeGraphPPR() { return reserve; }
    @Override public void printWholeGraphResult() { /* unchanged from the reference */ }
    @Override public void preprocessing(Double dummy, Object epsilon) { /* unchanged: loops computeWholeGraphPPR over forEachNode and writes <id>.txt */ }
    @Override public void readPreprocessedPPR(Long nodeId_start) { /* unchanged */ }
    @Override public Long getPrepSize() { return 0L; /* unchanged */ }
    @Override public void deletePrepDir() { /* unchanged */ }
}
