package joezie.fora_neo4j;

import java.io.BufferedReader;
import java.io.BufferedWriter;
import java.io.IOException;
import java.nio.charset.StandardCharsets;
import java.nio.file.Files;
import java.nio.file.Path;
import java.nio.file.Paths;
import java.util.ArrayList;
import java.util.Comparator;
import java.util.HashMap;
import java.util.List;
import java.util.Map;
import java.util.Vector;
import java.util.stream.Stream;

import org.neo4j.graphalgo.api.Graph;
import org.neo4j.graphdb.GraphDatabaseService;

/**
 * What the rewired algorithm classes share: the GPU context, the result maps the reference's getters hand out, the
 * top-k bookkeeping (kth largest, ties kept, sorted descending: J/Algo_Util.java:32-73 + retrieveTopK) done by
 * ppr_topk, and the on-disk result store "<ALGO>_ppr_results/<db>/<param>/<nodeId>.txt" with lines
 * "id \t Double.toString(value)" (J/Forward_Push.java:289-363, :432-443).
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK).  The Python classes in algo.py and the C++ classes in
 * host/ppr_host.hpp implement the same behaviour and are what the tests run.
 */
abstract class Gpu_Algo_Base extends Algo_Util {
    protected final PprB200 gpu;
    protected final GraphDatabaseService graphDb;
    protected final Graph adjM;
    protected final int node_amount;
    protected final Double alpha;
    protected HashMap<Long, Double> reserve = new HashMap<>();
    protected HashMap<Long, Double> topk_res = new HashMap<>();
    protected Vector<Long> topk_nodeIds = new Vector<>();
    protected String preprocessing_dirName;
    protected long seed = 1;                       // one Philox key per query: results are reproducible run to run

    protected Gpu_Algo_Base(String node_property, Double alpha, int node_amount, GraphDatabaseService graphDb, Graph adjM, String prep_root) {
        super(node_property);
        this.gpu = PPR.gpu();                      // the context that exported adjM in PPR.setupAdjMatrix
        this.graphDb = graphDb;
        this.adjM = adjM;
        this.node_amount = node_amount;
        this.alpha = alpha;
        this.preprocessing_dirName = prep_root;
    }

    /** one single-source query; fills {@link #reserve} */
    protected abstract void query(Long nodeId_start, Object param) throws Throwable;

    public void computeWholeGraphPPR(Long nodeId_start, Object param) {
        try { query(nodeId_start, param); }
        catch (Throwable t) { t.printStackTrace(); }          // the reference's compute methods never throw either
    }

    public HashMap<Long, Double> getWholeGraphPPR() { return reserve; }

    /** first line of printWholeGraphResult, e.g. "Fora-Whole-Graph PPR:" (J/Fora_Whole_Graph.java:65) */
    protected abstract String wholeGraphHeader();
    /** "Forward-Push-Top" in "\nForward-Push-Top10 PPR:" (J/Forward_Push.java:384) */
    protected abstract String topkHeaderPrefix();

    public void printWholeGraphResult() { printSorted(wholeGraphHeader(), reserve, Integer.MAX_VALUE); }

    // ---- top-k over the whole-graph result (Forward_Push / Monte_Carlo / Power_Method / Base_Whole_Graph)
    public void computeTopKPPR(Long nodeId_start, int k, Object param) {
        computeWholeGraphPPR(nodeId_start, param);
        retrieveTopK(k);
    }

    protected void retrieveTopK(int k) {
        topk_res.clear();
        topk_nodeIds.clear();
        try {
            for (Map.Entry<Long, Double> e : gpu.topk(reserve, k)) { topk_res.put(e.getKey(), e.getValue()); topk_nodeIds.add(e.getKey()); }
        } catch (Throwable t) { t.printStackTrace(); }
    }

    public Vector<Long> getTopKNodeIds(int k) { return topk_nodeIds; }

    public void printTopKResult(int k) {
        if (topk_res.isEmpty()) retrieveTopK(k);
        printSorted("\n" + topkHeaderPrefix() + k + " PPR:", topk_res, k);
    }

    /** the print loops of the reference (e.g. J/Forward_Push.java:366-391): value-descending (stable), "@name \t value", at most limit rows */
    protected void printSorted(String header, Map<Long, Double> res, int limit) {
        List<Map.Entry<Long, Double>> rows = new ArrayList<>(res.entrySet());
        rows.sort(Map.Entry.<Long, Double>comparingByValue(Comparator.reverseOrder()));
        System.out.println(header);
        int count = 0;
        for (Map.Entry<Long, Double> e : rows) {
            if (count >= limit) break;     // ties at the cut may have kept more than k entries
            count++;
            System.out.println("@" + getNodeName(e.getKey(), graphDb) + '\t' + e.getValue());
        }
    }

    // ---- result store (Preprocessing_Interface)
    public void preprocessing(Double threshold, Object param) {
        preprocessing_dirName += "/" + param;
        Path dir = Paths.get(preprocessing_dirName);
        try {
            if (Files.isDirectory(dir)) clearDirectory(dir); else Files.createDirectories(dir);
        } catch (IOException e) { e.printStackTrace(); }
        adjM.forEachNode(nodeIdM -> {
            Long nodeId_start = adjM.toOriginalNodeId(nodeIdM);
            computeWholeGraphPPR(nodeId_start, param);
            writeRow(dir.resolve(nodeId_start + ".txt"), reserve.entrySet());
            if (node_amount >= 10 && (nodeIdM + 1) % Math.max(1, node_amount / 10) == 0) System.out.println("Progress: " + (100L * (nodeIdM + 1) / node_amount) + "%");
            return true;
        });
    }

    protected static void writeRow(Path file, Iterable<Map.Entry<Long, Double>> row) {
        try (BufferedWriter w = Files.newBufferedWriter(file, StandardCharsets.US_ASCII)) {
            for (Map.Entry<Long, Double> e : row) { w.write(e.getKey().toString()); w.write('\t'); w.write(e.getValue().toString()); w.write('\n'); }
        } catch (IOException e) {
            System.out.println("Write to file " + file + " failed!");
            e.printStackTrace();
        }
    }

    public void readPreprocessedPPR(Long nodeId_start) {
        reserve = readRow(Paths.get(preprocessing_dirName, nodeId_start + ".txt"), new HashMap<>());
    }

    protected static <M extends Map<Long, Double>> M readRow(Path file, M into) {
        try (BufferedReader r = Files.newBufferedReader(file, StandardCharsets.US_ASCII)) {
            for (String line = r.readLine(); line != null; line = r.readLine()) {
                int tab = line.indexOf('\t');
                into.put(Long.parseLong(line.substring(0, tab)), Double.parseDouble(line.substring(tab + 1)));
            }
        } catch (IOException e) {
            System.out.println("Read from file " + file + " failed!");
            e.printStackTrace();
        }
        return into;
    }

    public Long getPrepSize() {
        try (Stream<Path> files = Files.walk(Paths.get(preprocessing_dirName))) {
            return files.filter(Files::isRegularFile).mapToLong(p -> p.toFile().length()).sum();
        } catch (IOException e) { return 0L; }
    }

    public void deletePrepDir() {
        Path dir = Paths.get(preprocessing_dirName);
        try { clearDirectory(dir); Files.deleteIfExists(dir); }
        catch (IOException e) {
            System.out.println("Clear directory " + preprocessing_dirName + " failed!");
            e.printStackTrace();
        }
    }

    private static void clearDirectory(Path dir) throws IOException {
        try (Stream<Path> all = Files.walk(dir)) {
            for (Path p : (Iterable<Path>) all.sorted(Comparator.reverseOrder())::iterator) if (!p.equals(dir)) Files.delete(p);
        }
    }
}
