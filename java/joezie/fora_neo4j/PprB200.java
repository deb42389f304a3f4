package joezie.fora_neo4j;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import org.neo4j.graphalgo.api.Graph;
import org.neo4j.graphdb.Direction;

/**
 * Panama (java.lang.foreign, JDK 22+) binding of libpprb200.so -- the C ABI in include/pprb200.h.
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK): this is the binding a maintainer of the reference
 * adds so that the existing classes (Forward_Push, Monte_Carlo, Fora_Whole_Graph, Fora_Topk,
 * Backward_Search, Base_Whole_Graph, Power_Method) keep their names and interfaces and delegate their
 * inner loops to the GPU.  The Python ctypes mirror (_native.py) binds the very same symbols and is
 * what the test-suite exercises.
 */
public final class PprB200 implements AutoCloseable {
    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("pprb200.lib", "libpprb200.so"), Arena.global());
    private static final ValueLayout.OfInt I32 = ValueLayout.JAVA_INT;
    private static final ValueLayout.OfLong I64 = ValueLayout.JAVA_LONG;
    private static final ValueLayout.OfDouble F64 = ValueLayout.JAVA_DOUBLE;
    private static final AddressLayout PTR = ValueLayout.ADDRESS;

    private static MethodHandle fn(String name, FunctionDescriptor d) { return LINKER.downcallHandle(LIB.find(name).orElseThrow(), d); }

    private static final MethodHandle INIT = fn("ppr_init", FunctionDescriptor.of(I32, I32, PTR));
    private static final MethodHandle SHUTDOWN = fn("ppr_shutdown", FunctionDescriptor.of(I32, PTR));
    private static final MethodHandle LAST_ERROR = fn("ppr_last_error", FunctionDescriptor.of(PTR, PTR));
    private static final MethodHandle GRAPH_CREATE = fn("ppr_graph_create", FunctionDescriptor.of(I32, PTR, I64, I64, PTR, PTR, PTR));
    private static final MethodHandle GRAPH_DESTROY = fn("ppr_graph_destroy", FunctionDescriptor.of(I32, PTR));
    private static final MethodHandle FORWARD_PUSH = fn("ppr_forward_push", FunctionDescriptor.of(I32, PTR, I32, F64, F64, PTR, PTR, PTR, PTR));
    private static final MethodHandle MONTE_CARLO = fn("ppr_monte_carlo", FunctionDescriptor.of(I32, PTR, I32, F64, F64, I64, PTR, PTR, PTR));
    private static final MethodHandle FORA = fn("ppr_fora", FunctionDescriptor.of(I32, PTR, I32, F64, F64, I64, I32, PTR, PTR, PTR));
    private static final MethodHandle FORA_TOPK = fn("ppr_fora_topk", FunctionDescriptor.of(I32, PTR, I32, F64, F64, I32, I64, I32, PTR, PTR, PTR, I32, PTR, PTR));
    private static final MethodHandle FORA_TOPK_SHARD = fn("ppr_fora_topk_shard", FunctionDescriptor.of(I32, PTR, I32, F64, F64, I32, I64, I32, PTR, I32, I32, PTR, PTR, PTR, PTR, I32, PTR, PTR));
    /** ppr_reduce_fn: int32 (*)(void* user, uint64 dev_ptr, int64 n) */
    private static final FunctionDescriptor REDUCE_FN = FunctionDescriptor.of(I32, PTR, I64, I64);
    private static final MethodHandle TOPK = fn("ppr_topk", FunctionDescriptor.of(I32, PTR, PTR, I64, I32, PTR, PTR, I32, PTR, PTR));
    private static final MethodHandle POWER = fn("ppr_power_iteration", FunctionDescriptor.of(I32, PTR, I32, F64, I32, PTR));
    private static final MethodHandle BACKWARD = fn("ppr_backward_search", FunctionDescriptor.of(I32, PTR, I32, F64, F64, PTR, PTR));
    private static final MethodHandle ALL_PAIRS = fn("ppr_backward_all_pairs", FunctionDescriptor.of(I32, PTR, F64, F64, I32, I32, I32, PTR, PTR));
    private static final MethodHandle PAIRS_COUNT = fn("ppr_pairs_count", FunctionDescriptor.of(I64, PTR));
    private static final MethodHandle PAIRS_EXPORT = fn("ppr_pairs_export", FunctionDescriptor.of(I32, PTR, PTR, PTR, PTR, PTR));
    private static final MethodHandle PAIRS_DESTROY = fn("ppr_pairs_destroy", FunctionDescriptor.of(I32, PTR));
    private static final MethodHandle WALK_TERMINALS = fn("ppr_walk_terminals", FunctionDescriptor.of(I32, PTR, I32, I32, I64, I64, I64, I64, F64, I32, PTR));

    private final Arena arena = Arena.ofShared();
    private final MemorySegment ctx;
    private MemorySegment graph = MemorySegment.NULL;
    private int n;
    private long[] originalIds;   // mapped id -> Neo4j id (Graph.toOriginalNodeId)
    private Graph adjM;

    public PprB200(int device) {
        try {
            MemorySegment out = arena.allocate(PTR);
            check((int) INIT.invokeExact(device, out), MemorySegment.NULL);
            ctx = out.get(PTR, 0);
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    private void check(int rc, MemorySegment c) throws Throwable {
        if (rc != 0) {
            MemorySegment msg = ((MemorySegment) LAST_ERROR.invokeExact(c)).reinterpret(512);
            throw new IllegalStateException("pprb200 error " + rc + ": " + msg.getString(0));
        }
    }

    /** One-time export of the HeavyGraph topology (replaces nothing in PPR.setupAdjMatrix; runs right after it). */
    public void exportGraph(Graph adjM) throws Throwable {
        this.adjM = adjM;
        n = (int) adjM.nodeCount();
        long[] rowPtr = new long[n + 1];
        for (int v = 0; v < n; v++) rowPtr[v + 1] = rowPtr[v] + adjM.degree(v, Direction.OUTGOING);
        int m = (int) rowPtr[n];
        int[] colIdx = new int[m];
        originalIds = new long[n];
        for (int v = 0; v < n; v++) {
            originalIds[v] = adjM.toOriginalNodeId(v);
            final int base = (int) rowPtr[v];
            final int[] k = {0};
            adjM.forEachRelationship(v, Direction.OUTGOING, (s, t, rel) -> { colIdx[base + k[0]++] = t; return true; });
            java.util.Arrays.sort(colIdx, base, base + k[0]);   // canonical order: ascending target, duplicates kept
        }
        MemorySegment rp = arena.allocateFrom(I64, rowPtr), ci = arena.allocateFrom(I32, colIdx), out = arena.allocate(PTR);
        check((int) GRAPH_CREATE.invokeExact(ctx, (long) n, (long) m, rp, ci, out), ctx);
        graph = out.get(PTR, 0);
    }

    private java.util.HashMap<Long, Double> toMap(MemorySegment dense) {
        java.util.HashMap<Long, Double> res = new java.util.HashMap<>();
        for (int v = 0; v < n; v++) { double x = dense.getAtIndex(F64, v); if (x > 0) res.put(originalIds[v], x); }
        return res;
    }

    /** Fora_Whole_Graph.computeWholeGraphPPR: delta / pfail as the class holds them (Algo_Conf), seed per query. */
    public java.util.HashMap<Long, Double> fora(long nodeIdStart, double alpha, double eps, double delta, double pfail, long seed) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(F64, n), conf = a.allocate(F64, 3);
            conf.setAtIndex(F64, 0, delta); conf.setAtIndex(F64, 1, pfail); conf.setAtIndex(F64, 2, 0.0);
            check((int) FORA.invokeExact(graph, adjM.toMappedNodeId(nodeIdStart), alpha, eps, seed, 0, conf, out, MemorySegment.NULL), ctx);
            return toMap(out);
        }
    }

    /** Forward_Push.computeWholeGraphPPR: returns {reserve, residue}; rsum[0] receives getUpdatedRsum(). */
    public java.util.HashMap<Long, Double>[] forwardPush(long nodeIdStart, double alpha, double rmax, double[] rsum) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment res = a.allocate(F64, n), rr = a.allocate(F64, n), rs = a.allocate(F64);
            check((int) FORWARD_PUSH.invokeExact(graph, adjM.toMappedNodeId(nodeIdStart), alpha, rmax, res, rr, rs, MemorySegment.NULL), ctx);
            rsum[0] = rs.get(F64, 0);
            @SuppressWarnings("unchecked") java.util.HashMap<Long, Double>[] out = new java.util.HashMap[] {toMap(res), toMap(rr)};
            return out;
        }
    }

    /** Monte_Carlo.computeWholeGraphPPR */
    public java.util.HashMap<Long, Double> monteCarlo(long nodeIdStart, double alpha, double eps, double delta, double pfail, long seed) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(F64, n), conf = a.allocate(F64, 3);
            conf.setAtIndex(F64, 0, delta); conf.setAtIndex(F64, 1, pfail); conf.setAtIndex(F64, 2, 0.0);
            check((int) MONTE_CARLO.invokeExact(graph, adjM.toMappedNodeId(nodeIdStart), alpha, eps, seed, conf, out, MemorySegment.NULL), ctx);
            return toMap(out);
        }
    }

    /** Fora_Topk.computeTopKPPR + getTopKNodeIds: ids sorted by estimate, ties included (may exceed k). */
    public java.util.Vector<Long> foraTopk(long nodeIdStart, double alpha, double eps, int k, double delta, double pfail, double minDelta, long seed) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment ids = a.allocate(I32, n), vals = a.allocate(F64, n), cnt = a.allocate(I32), conf = a.allocate(F64, 3);
            conf.setAtIndex(F64, 0, delta); conf.setAtIndex(F64, 1, pfail); conf.setAtIndex(F64, 2, minDelta);
            check((int) FORA_TOPK.invokeExact(graph, adjM.toMappedNodeId(nodeIdStart), alpha, eps, k, seed, 0, conf, ids, vals, n, cnt, MemorySegment.NULL), ctx);
            java.util.Vector<Long> out = new java.util.Vector<>();
            for (int i = 0; i < cnt.get(I32, 0); i++) out.add(originalIds[ids.getAtIndex(I32, i)]);
            return out;
        }
    }

    /** Fora_Topk as (node id, estimate) entries in result order (estimate descending, ties included). */
    public java.util.List<java.util.Map.Entry<Long, Double>> foraTopkEntries(long nodeIdStart, double alpha, double eps, int k, double delta, double pfail, double minDelta, long seed) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            int cap = (int) Math.min((long) n, Math.max(4L * k, k + 4096L));
            for (;;) {
                MemorySegment ids = a.allocate(I32, cap), vals = a.allocate(F64, cap), cnt = a.allocate(I32), conf = a.allocate(F64, 3);
                conf.setAtIndex(F64, 0, delta); conf.setAtIndex(F64, 1, pfail); conf.setAtIndex(F64, 2, minDelta);
                int rc = (int) FORA_TOPK.invokeExact(graph, adjM.toMappedNodeId(nodeIdStart), alpha, eps, k, seed, 0, conf, ids, vals, cap, cnt, MemorySegment.NULL);
                if (rc == -5 /* PPR_E_CAPACITY: more ties than room; *count_out holds the needed size */) { cap = cnt.get(I32, 0); continue; }
                check(rc, ctx);
                java.util.List<java.util.Map.Entry<Long, Double>> out = new java.util.ArrayList<>();
                for (int i = 0; i < cnt.get(I32, 0); i++) out.add(new java.util.AbstractMap.SimpleEntry<>(originalIds[ids.getAtIndex(I32, i)], vals.getAtIndex(F64, i)));
                return out;
            }
        }
    }

    /** Algo_Util.kth_ppr + retrieveTopK + getTopKNodeIds over a result map: entries >= the k-th largest value, sorted descending. */
    public java.util.List<java.util.Map.Entry<Long, Double>> topk(java.util.HashMap<Long, Double> res, int k) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment dense = a.allocate(F64, n), ids = a.allocate(I32, n), vals = a.allocate(F64, n), cnt = a.allocate(I32);
            for (java.util.Map.Entry<Long, Double> e : res.entrySet()) dense.setAtIndex(F64, adjM.toMappedNodeId(e.getKey()), e.getValue());
            check((int) TOPK.invokeExact(ctx, dense, (long) n, k, ids, vals, n, cnt, MemorySegment.NULL), ctx);
            java.util.List<java.util.Map.Entry<Long, Double>> out = new java.util.ArrayList<>();
            for (int i = 0; i < cnt.get(I32, 0); i++) out.add(new java.util.AbstractMap.SimpleEntry<>(originalIds[ids.getAtIndex(I32, i)], vals.getAtIndex(F64, i)));
            return out;
        }
    }

    /** Monte_Carlo.random_walk / random_walk_no_zero_hop: terminal of walk `walkId` from nodeIdStart (restart at the start vertex, as the reference). */
    public long walkTerminal(long nodeIdStart, long walkId, long seed, double alpha, boolean noZeroHop) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(I32);
            int s = adjM.toMappedNodeId(nodeIdStart);
            check((int) WALK_TERMINALS.invokeExact(graph, s, s, walkId, 1L, seed, 0L, alpha, noZeroHop ? 1 : 0, out), ctx);
            return originalIds[out.get(I32, 0)];
        }
    }

    /** Sums n int64 values at a device address over all ranks, in place (e.g. ncclAllReduce(ptr, ptr, n, ncclInt64, ncclSum, comm, stream) + stream sync). */
    public interface DeviceSum { int sum(long devPtr, long n); }

    /** Fora_Topk with every delta-round's walks sharded over GPUs: same result on every rank, identical to foraTopk. */
    public java.util.Vector<Long> foraTopkSharded(long nodeIdStart, double alpha, double eps, int k, long seed, int shard, int nshards, DeviceSum reduce) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            int cap = (int) Math.min((long) n, Math.max(4L * k, k + 4096L));
            MemorySegment ids = a.allocate(I32, cap), vals = a.allocate(F64, cap), cnt = a.allocate(I32);
            MethodHandle target = java.lang.invoke.MethodHandles.lookup().findVirtual(DeviceSum.class, "sum", java.lang.invoke.MethodType.methodType(int.class, long.class, long.class)).bindTo(reduce);
            target = java.lang.invoke.MethodHandles.dropArguments(target, 0, MemorySegment.class);   // (user, devPtr, n) -> sum(devPtr, n)
            MemorySegment hook = nshards > 1 ? LINKER.upcallStub(target, REDUCE_FN, a) : MemorySegment.NULL;
            check((int) FORA_TOPK_SHARD.invokeExact(graph, adjM.toMappedNodeId(nodeIdStart), alpha, eps, k, seed, 0, MemorySegment.NULL, shard, nshards, hook, MemorySegment.NULL, ids, vals, cap, cnt, MemorySegment.NULL), ctx);
            java.util.Vector<Long> out = new java.util.Vector<>();
            for (int i = 0; i < cnt.get(I32, 0); i++) out.add(originalIds[ids.getAtIndex(I32, i)]);
            return out;
        }
    }

    /** Power_Method.computeWholeGraphPPR (100 iterations in the reference) */
    public java.util.HashMap<Long, Double> powerMethod(long nodeIdStart, double alpha, int iterations) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(F64, n);
            check((int) POWER.invokeExact(graph, adjM.toMappedNodeId(nodeIdStart), alpha, iterations, out), ctx);
            return toMap(out);
        }
    }

    /** Backward_Search.backward_search_whole_graph */
    public java.util.HashMap<Long, Double> backwardSearch(long nodeIdTarget, double alpha, double rmax) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(F64, n);
            check((int) BACKWARD.invokeExact(graph, adjM.toMappedNodeId(nodeIdTarget), alpha, rmax, out, MemorySegment.NULL), ctx);
            return toMap(out);
        }
    }

    /** Base_Whole_Graph.preprocessing, compute part: the inverted index v -> [(t, pi(v,t))] as CSR arrays. */
    public static final class Pairs { public long[] srcPtr; public int[] tgt; public double[] val; }
    public Pairs backwardAllPairs(double alpha, double threshold, int k) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(PTR);
            check((int) ALL_PAIRS.invokeExact(graph, alpha, threshold, k, 0, n, out, MemorySegment.NULL), ctx);
            MemorySegment p = out.get(PTR, 0);
            int cnt = (int) (long) PAIRS_COUNT.invokeExact(p);
            MemorySegment sp = a.allocate(I64, n + 1), tg = a.allocate(I32, Math.max(cnt, 1)), vl = a.allocate(F64, Math.max(cnt, 1));
            check((int) PAIRS_EXPORT.invokeExact(p, sp, tg, vl, MemorySegment.NULL), ctx);
            int ignored = (int) PAIRS_DESTROY.invokeExact(p);
            Pairs r = new Pairs();
            r.srcPtr = sp.toArray(I64); r.tgt = tg.asSlice(0, 4L * cnt).toArray(I32); r.val = vl.asSlice(0, 8L * cnt).toArray(F64);
            return r;
        }
    }

    @Override public void close() {
        try {
            if (!graph.equals(MemorySegment.NULL)) { int rc = (int) GRAPH_DESTROY.invokeExact(graph); }
            int rc = (int) SHUTDOWN.invokeExact(ctx);
        } catch (Throwable ignored) { }
        arena.close();
    }
}
