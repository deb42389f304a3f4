package joezie.fora_neo4j;

import java.nio.file.Files;
import java.nio.file.Path;
import java.nio.file.Paths;
import java.util.AbstractMap.SimpleEntry;
import java.util.ArrayList;
import java.util.HashMap;
import java.util.LinkedHashMap;
import java.util.List;
import java.util.Map;
import java.util.Vector;

import org.neo4j.graphalgo.api.Graph;
import org.neo4j.graphdb.GraphDatabaseService;

/**
 * Drop-in replacement of the reference class (All-Pair-Backward-Search): the n sequential backward searches and the
 * in-memory inverted index of preprocessing (J/Base_Whole_Graph.java:58-164) are one ppr_backward_all_pairs call, which
 * returns the index grouped by source (targets ascending, or the top-k rows sorted by value when k >= 0); this class
 * writes the same "<id>.txt" files.  Queries read those files, as in the reference (:167-218).
 * NOT COMPILED HERE (no JDK in the image).
 */
public class Base_Whole_Graph extends Gpu_Algo_Base implements Whole_Graph_Util_Interface, Preprocessing_Interface, Topk_Util_Interface {
    private LinkedHashMap<Long, Double> ppr_src = new LinkedHashMap<>();   // insertion order = file order

    public Base_Whole_Graph(Double alpha, int node_amount, GraphDatabaseService graphDb, Graph adjM, String node_property, String dir_db) {
        super(node_property, alpha, node_amount, graphDb, adjM, "BASE_ppr_results/" + dir_db);
    }

    @Override protected void query(Long nodeId_start, Object dummy) { readPreprocessedPPR(nodeId_start); }

    @Override public void preprocessing(Double threshold, Object k) {
        preprocessing_dirName += ("/" + threshold + "_" + k);   // J/Base_Whole_Graph.java:65
        Path dir = Paths.get(preprocessing_dirName);
        try {
            if (Files.isDirectory(dir)) deletePrepDir();
            Files.createDirectories(dir);
            PprB200.Pairs idx = gpu.backwardAllPairs(alpha, threshold, (Integer) k);
            for (int v = 0; v < node_amount; v++) {
                List<Map.Entry<Long, Double>> row = new ArrayList<>();
                for (long i = idx.srcPtr[v]; i < idx.srcPtr[v + 1]; i++)
                    row.add(new SimpleEntry<>(adjM.toOriginalNodeId(idx.tgt[(int) i]), idx.val[(int) i]));
                if (!row.isEmpty()) writeRow(dir.resolve(adjM.toOriginalNodeId(v) + ".txt"), row);   // only sources present in the index get a file (the reference iterates ppr_all_pairs, :112-116)
            }
        } catch (Throwable t) { t.printStackTrace(); }
    }

    @Override public void readPreprocessedPPR(Long nodeId_start) {
        ppr_src = readRow(Paths.get(preprocessing_dirName, nodeId_start + ".txt"), new LinkedHashMap<>());
        reserve = ppr_src;
    }

    @Override public HashMap<Long, Double> getWholeGraphPPR() { return ppr_src; }

    @Override public void computeTopKPPR(Long nodeId_start, int dummy1, Object dummy2) { readPreprocessedPPR(nodeId_start); }   // rows were cut to top-k at preprocessing time
    @Override public Vector<Long> getTopKNodeIds(int k) { return new Vector<>(ppr_src.keySet()); }
    @Override protected String wholeGraphHeader() { return "\nBase-Whole-Graph PPR:"; }
    @Override protected String topkHeaderPrefix() { return "BASE-Top"; }
    @Override public void printWholeGraphResult() {   // J/Base_Whole_Graph.java:188-193: stored order, not re-sorted
        System.out.println(wholeGraphHeader());
        for (Map.Entry<Long, Double> e : ppr_src.entrySet()) System.out.println("@" + getNodeName(e.getKey(), graphDb) + '\t' + e.getValue());
    }
    @Override public void printTopKResult(int k) { printSorted("\n" + topkHeaderPrefix() + k + " PPR:", ppr_src, k); }
}
