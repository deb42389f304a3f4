// ppr_oracle.cpp -- CPU oracle for the Personalized-PageRank hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// The shipped library (libpprb200.so) never links, imports or calls this file.
//
// PARITY UNPINNED: the reference (joezie/Personalized-PageRank-Algorithms-on-Neo4j) ships no tests,
// no golden vectors, cannot be built here (no JDK; J/Fora_Topk.java:136 has a syntax error; the
// neo4j-graph-algorithms jar is undeclared in pom.xml) and is unseeded (ThreadLocalRandom).  This file
// restates its algorithms from the source text; the only external anchors are the GOT power-method
// values recorded in SURVEY.md section 8c and the Random123 Philox4x32-10 known-answer vectors.
// A second, independently written restatement (tests/java_replay.py, plain Python) is held equal to
// mode A bit for bit by tests/test_java_replay.py -- a transcription check, not a pin.
//
// `J/` = /root/reference/src/main/java/joezie/fora_neo4j/.
//
// Three modes:
//   A  reference-order replay in double arithmetic (FIFO asynchronous push, same formulas, same
//      edge cases).  Semantic twin of the Java; also the timed CPU baseline ("port").
//   B  the synchronous-schedule / fixed-point / Philox specification that the CUDA path must match
//      bit for bit (see SPEC below).
//   C  exact answer: J/Power_Method.java (100 Jacobi rounds) and an iterate-to-tolerance variant.
//
// ---------------------------------------------------------------------------------------------
// SPEC (mode B) -- shared by this oracle and csrc/ (restated there, not shared code)
//   * Fixed point: value x is stored as the integer x * 2^62 in a uint64 ("FX").  FX_ONE = 2^62.
//   * alpha_q64 = (uint64) (alpha * 2^64)   (exact power-of-two scaling, truncation).
//   * Forward push of vertex v holding residue x, out-degree d:
//        a = mulhi64(x, alpha_q64);  reserve[v] += a;  y = x - a;
//        d > 0 : q = y / d  (integer),  r[v] = y - q*d  (the <d leftover units stay in r[v]),
//                every out-edge (v,u) adds q to r[u];
//        d == 0: q = y goes to r[source];  r[v] = 0.
//     Mass is conserved exactly: sum(reserve) + sum(r) == FX_ONE at all times.
//   * Threshold: rmax_fx = max(2^64/alpha_q64 + 2, (uint64)(rmax * 2^62))  (floor: a push must move >= 1 unit to reserve);
//        thr(u) = d(u)==0 ? 1 : saturating(rmax_fx * d(u)).     (J/Forward_Push.java:132: r/d >= rmax;
//        d==0 gives +inf >= rmax, i.e. any positive residue on a dangling vertex is pushed)
//   * Synchronous schedule: iteration k's frontier is {v : r[v] >= thr(v)}; all frontier vertices
//     read their residue first (phase A), then all scatter (phase B).  Integer adds commute, so the
//     state after each iteration is schedule independent.
//   * Walks: Philox4x32-10, key = seed + tag*0x9E3779B97F4A7C15 split (lo,hi), counter =
//     (start vertex, walk_index lo32, walk_index hi32, step).  One Philox block per step:
//        stop  iff  ((x0<<21)|(x1>>11)) < ceil(alpha * 2^53)          (J/Monte_Carlo.java:76)
//        move to out[cur][ mulhi64((x2<<32)|x3, d(cur)) ]              (J/Monte_Carlo.java:84-85)
//     dangling cur -> restart vertex, no draw used (J/Monte_Carlo.java:87-90); dangling start ->
//     returns start without any draw (J/Monte_Carlo.java:70-72).  no_zero_hop: step 0 is a mandatory
//     move using only (x2,x3) (J/Monte_Carlo.java:110-112), termination draws start at step 1.
// ---------------------------------------------------------------------------------------------
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <vector>

typedef uint64_t u64;
typedef int64_t i64;
typedef int32_t i32;
typedef uint32_t u32;

namespace {

const u64 FX_ONE = 1ull << 62;
const double FX_SCALE = 4611686018427387904.0;       // 2^62
const double FX_INV = 1.0 / 4611686018427387904.0;   // 2^-62 (exact)

inline u64 mulhi64(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) >> 64); }

struct Graph {
    i64 n, m;
    const i64* rp;  // row_ptr[n+1]
    const i32* ci;  // col_idx[m]
    std::vector<i64> cp;  // CSC col_ptr (built on demand; sources ascending inside a column)
    std::vector<i32> ri;
    Graph(i64 n_, i64 m_, const i64* rp_, const i32* ci_) : n(n_), m(m_), rp(rp_), ci(ci_) {}
    inline i64 deg(i64 v) const { return rp[v + 1] - rp[v]; }
    void build_csc() {
        if (!cp.empty()) return;
        cp.assign(n + 1, 0);
        ri.assign(m, 0);
        for (i64 e = 0; e < m; e++) cp[ci[e] + 1]++;
        for (i64 v = 0; v < n; v++) cp[v + 1] += cp[v];
        std::vector<i64> pos(cp.begin(), cp.end() - 1);
        for (i64 v = 0; v < n; v++)
            for (i64 e = rp[v]; e < rp[v + 1]; e++) ri[pos[ci[e]]++] = (i32)v;
    }
    inline i64 indeg(i64 v) const { return cp[v + 1] - cp[v]; }
};

// ------------------------------------------------------------------ Philox4x32-10 (Random123)
struct Philox4 { u32 x[4]; };
inline Philox4 philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1) {
    const u32 M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; r++) {
        u64 p0 = (u64)M0 * c0, p1 = (u64)M1 * c2;
        u32 n0 = (u32)(p1 >> 32) ^ c1 ^ k0;
        u32 n1 = (u32)p1;
        u32 n2 = (u32)(p0 >> 32) ^ c3 ^ k1;
        u32 n3 = (u32)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x[0] = c0; o.x[1] = c1; o.x[2] = c2; o.x[3] = c3;
    return o;
}

inline u64 walk_key(u64 seed, u64 tag) { return seed + tag * 0x9E3779B97F4A7C15ull; }
inline u64 stop_threshold(double alpha) { return (u64)std::ceil(std::ldexp(alpha, 53)); }
inline u64 alpha_q64_of(double alpha) { return (u64)std::ldexp(alpha, 64); }

// J/Monte_Carlo.java:60-94 (random_walk) and :96-133 (random_walk_no_zero_hop).
// `restart` is the vertex a walk jumps to from a dangling vertex: the Java always uses the walk's own
// start (F12 in SURVEY.md); passing the query source instead gives the unbiased estimator.
i32 walk(const Graph& g, i32 start, i32 restart, u64 j, u64 key, u64 tstop, bool no_zero_hop, i64* steps) {
    i64 d0 = g.deg(start);
    if (d0 == 0) return start;
    u32 k0 = (u32)key, k1 = (u32)(key >> 32);
    u32 c1 = (u32)j, c2 = (u32)(j >> 32);
    i32 cur = start;
    u32 step = 0;
    if (no_zero_hop) {
        Philox4 x = philox4x32_10((u32)start, c1, c2, step++, k0, k1);
        u64 rnd = ((u64)x.x[2] << 32) | x.x[3];
        cur = g.ci[g.rp[start] + (i64)mulhi64(rnd, (u64)d0)];
        if (steps) (*steps)++;
    }
    for (;;) {
        Philox4 x = philox4x32_10((u32)start, c1, c2, step++, k0, k1);
        u64 u53 = ((u64)x.x[0] << 21) | (x.x[1] >> 11);
        if (u53 < tstop) break;
        i64 d = g.deg(cur);
        if (d > 0) {
            u64 rnd = ((u64)x.x[2] << 32) | x.x[3];
            cur = g.ci[g.rp[cur] + (i64)mulhi64(rnd, (u64)d)];
        } else {
            cur = restart;
        }
        if (steps) (*steps)++;
    }
    return cur;
}

// ------------------------------------------------------------------ parameter derivation
// J/Algo_Conf.java:29-63 : delta = pfail = 1/n, rsum = 1.
// J/Algo_Conf.java:71-81 : top-k: delta0 = 1/k, min_delta = 1/n, pfail' = 1/n/n/ln(n / k) with INTEGER n/k.
struct Conf { double delta, pfail, min_delta; };
Conf conf_whole(i64 n) { Conf c; c.delta = 1.0 / (double)n; c.pfail = 1.0 / (double)n; c.min_delta = c.delta; return c; }
Conf conf_topk(i64 n, i64 k) {
    Conf c;
    c.min_delta = 1.0 / (double)n;
    c.delta = 1.0 / (double)k;
    i64 q = (k != 0) ? n / k : 0;
    c.pfail = 1.0 / (double)n / (double)n / std::log((double)q);
    return c;
}

// ------------------------------------------------------------------ k-th largest / top-k
// J/Algo_Util.java:32-73: k-th largest; null when k is out of range.  (k == 0 throws in the Java;
// treated as "null" here.)
bool kth_largest(std::vector<double> v, i64 k, double* out) {
    if (k <= 0 || k > (i64)v.size()) return false;
    std::nth_element(v.begin(), v.begin() + (k - 1), v.end(), [](double a, double b) { return a > b; });
    *out = v[k - 1];
    return true;
}
bool kth_largest_u64(std::vector<u64> v, i64 k, u64* out) {
    if (k <= 0 || k > (i64)v.size()) return false;
    std::nth_element(v.begin(), v.begin() + (k - 1), v.end(), [](u64 a, u64 b) { return a > b; });
    *out = v[k - 1];
    return true;
}

// retrieveTopK (J/Forward_Push.java:413-429, J/Fora_Topk.java:186-199) + getTopKNodeIds sort
// (J/Fora_Topk.java:82-99).  "entries" are the vertices with a positive value (= keys present in
// the Java map).  Ties keep ascending id (Java leaves tie order to HashMap iteration).
template <typename T>
i64 topk_select(const T* vals, i64 n, i64 k, i32* ids_out, i64 cap) {
    std::vector<T> pos;
    std::vector<i32> idx;
    for (i64 v = 0; v < n; v++) if (vals[v] > 0) { pos.push_back(vals[v]); idx.push_back((i32)v); }
    T kth = 0;
    bool have = false;
    if (k > 0 && k <= (i64)pos.size()) {
        std::vector<T> tmp(pos);
        std::nth_element(tmp.begin(), tmp.begin() + (k - 1), tmp.end(), [](T a, T b) { return a > b; });
        kth = tmp[k - 1];
        have = true;
    }
    std::vector<i64> keep;
    for (i64 i = 0; i < (i64)pos.size(); i++) if (!have || pos[i] >= kth) keep.push_back(i);
    std::stable_sort(keep.begin(), keep.end(), [&](i64 a, i64 b) { return pos[a] > pos[b]; });
    i64 cnt = (i64)keep.size();
    for (i64 i = 0; i < cnt && i < cap; i++) ids_out[i] = idx[keep[i]];
    return cnt;
}

// ------------------------------------------------------------------ mode C: power method
// J/Power_Method.java:44-101.  Entries are visited in ascending id (the Java visits HashMap order).
void power_method(const Graph& g, i32 s, double alpha, int iters, double tol, double* reserve, int* iters_done) {
    i64 n = g.n;
    std::vector<double> r(n, 0.0), nr(n, 0.0);
    std::fill(reserve, reserve + n, 0.0);
    r[s] = 1.0;
    int it = 0;
    for (; it < iters; it++) {
        std::fill(nr.begin(), nr.end(), 0.0);
        double left = 0.0;
        for (i64 v = 0; v < n; v++) {
            double rc = r[v];
            if (rc > 0) {
                reserve[v] += rc * alpha;
                double remain = rc * (1 - alpha);
                left += remain;
                i64 d = g.deg(v);
                if (d == 0) nr[s] += remain;
                else {
                    double avg = remain / (double)d;
                    for (i64 e = g.rp[v]; e < g.rp[v + 1]; e++) nr[g.ci[e]] += avg;
                }
            }
        }
        r.swap(nr);
        if (tol > 0 && left < tol) { it++; break; }
    }
    if (iters_done) *iters_done = it;
}

// ------------------------------------------------------------------ mode A: forward push
struct PushStats { i64 pushes, edges; };

// J/Forward_Push.java:63-142
void a_forward_push(const Graph& g, i32 s, double alpha, double rmax, double* reserve, double* residue,
                    double* rsum_field, PushStats* st) {
    i64 n = g.n;
    std::fill(reserve, reserve + n, 0.0);
    std::fill(residue, residue + n, 0.0);
    double rsum_local = 1.0;
    i64 d_s = g.deg(s);
    if (d_s == 0) { reserve[s] = 1.0; *rsum_field = 0.0; return; }
    std::vector<char> inq(n, 0);
    std::deque<i32> Q;
    Q.push_back(s); inq[s] = 1; residue[s] = 1.0;
    while (!Q.empty()) {
        i32 v = Q.front(); Q.pop_front();
        inq[v] = 0;
        double rc = residue[v];
        residue[v] = 0.0;
        reserve[v] += rc * alpha;
        rsum_local -= rc * alpha;
        i64 d = g.deg(v);
        if (st) st->pushes++;
        if (d == 0) {
            double ns = residue[s] + rc * (1.0 - alpha);
            residue[s] = ns;
            if (d_s > 0 && ns / (double)d_s >= rmax && !inq[s]) { Q.push_back(s); inq[s] = 1; }
            continue;  // NB: the rsum field is not refreshed on this path (J/Forward_Push.java:114)
        }
        double avg = ((1.0 - alpha) * rc) / (double)d;
        for (i64 e = g.rp[v]; e < g.rp[v + 1]; e++) {
            i32 u = g.ci[e];
            double nr = residue[u] + avg;
            residue[u] = nr;
            i64 du = g.deg(u);
            if (nr / (double)du >= rmax && !inq[u]) { Q.push_back(u); inq[u] = 1; }
        }
        if (st) st->edges += d;
        *rsum_field = rsum_local;
    }
}

// J/Forward_Push.java:144-250 -- incremental two-threshold push used by FORA top-k.
struct FwdTopkState {
    std::vector<double> residue, reserve;
    std::vector<i32> Q_next;
    double rsum;
};
void a_forward_push_topk(const Graph& g, i32 s, double alpha, FwdTopkState& S, std::deque<i32>& Q, double min_rmax,
                         bool first, double rmax, PushStats* st) {
    i64 n = g.n;
    i64 d_s = g.deg(s);
    if (d_s == 0) { S.reserve[s] = 1.0; S.rsum = 0.0; return; }
    if (first) S.residue[s] = 1.0;
    S.Q_next.clear();
    double rsum_local = S.rsum;
    std::vector<char> inq(n, 0), inq_next(n, 0);
    for (i32 v : Q) inq[v] = 1;
    while (!Q.empty()) {
        i32 v = Q.front(); Q.pop_front();
        inq[v] = 0;
        double rc = S.residue[v];
        i64 d = g.deg(v);
        if (rc / (double)d >= rmax) {
            S.residue[v] = 0.0;
            S.reserve[v] += rc * alpha;
            rsum_local -= rc * alpha;
            if (st) st->pushes++;
            if (d == 0) {
                double ns = S.residue[s] + rc * (1 - alpha);
                S.residue[s] = ns;
                if (d_s > 0 && ns / (double)d_s >= rmax && !inq[s]) { Q.push_back(s); inq[s] = 1; }
                else if (d_s > 0 && ns / (double)d_s >= min_rmax && !inq_next[s]) { inq_next[s] = 1; S.Q_next.push_back(s); }
                continue;
            }
            double avg = ((1.0 - alpha) * rc) / (double)d;
            for (i64 e = g.rp[v]; e < g.rp[v + 1]; e++) {
                i32 u = g.ci[e];
                double nr = S.residue[u] + avg;
                S.residue[u] = nr;
                i64 du = g.deg(u);
                if (nr / (double)du >= rmax && !inq[u]) { Q.push_back(u); inq[u] = 1; }
                else if (nr / (double)du >= min_rmax && !inq_next[u]) { inq_next[u] = 1; S.Q_next.push_back(u); }
            }
            if (st) st->edges += d;
        } else if (rc / (double)d >= min_rmax && !inq_next[v]) {
            inq_next[v] = 1; S.Q_next.push_back(v);
        }
    }
    S.rsum = rsum_local;
}

// J/Monte_Carlo.java:136-158
i64 a_monte_carlo(const Graph& g, i32 s, double alpha, double eps, u64 seed, double* ppr, i64* steps) {
    Conf c = conf_whole(g.n);
    double omega = 3 * std::log(2 / c.pfail) / eps / eps / c.delta;
    std::vector<i64> cnt(g.n, 0);
    u64 key = walk_key(seed, 0), tstop = stop_threshold(alpha);
    i64 nw = 0;
    for (i64 i = 1; (double)i <= omega; i++) { cnt[walk(g, s, s, (u64)(i - 1), key, tstop, false, steps)]++; nw++; }
    for (i64 v = 0; v < g.n; v++) ppr[v] = (double)cnt[v] / omega;
    return nw;
}

struct ForaStats { i64 pushes, edges, rounds, walks, steps; double rmax_final, rsum; };

// J/Fora_Whole_Graph.java:82-146.  The wall-clock balance loop (:93-103, 400 ns/walk) is replaced by a
// deterministic cost model: modelled push time = (vertices popped + edges relaxed) "edge units",
// modelled walk time = walk_cost * rsum * omega edge units.  Like the Java, every balance round
// restarts the push from scratch with rmax halved and the push cost accumulates over rounds.
void a_fora(const Graph& g, i32 s, double alpha, double eps, u64 seed, double walk_cost, int max_rounds,
            bool restart_at_source, double* reserve, ForaStats* fs) {
    i64 n = g.n;
    Conf c = conf_whole(n);
    double rsum_local = 1.0;
    double rmax_local = eps * std::sqrt(c.delta / 3.0 / (double)g.m / std::log(2.0 / c.pfail)) / (1.0 - alpha);
    double omega_local = (eps + 2.0) * std::log(2.0 / c.pfail) / eps / eps / c.delta;
    std::vector<double> residue(n, 0.0);
    PushStats ps = {0, 0};
    double rsum_field = 1.0;
    int rounds = 0;
    double cost = 0.0;
    while (cost < walk_cost * rsum_local * omega_local && rounds < max_rounds) {
        rsum_field = 1.0;  // new Forward_Push(alpha, 1.0, ...)
        PushStats one = {0, 0};
        a_forward_push(g, s, alpha, rmax_local, reserve, residue.data(), &rsum_field, &one);
        ps.pushes += one.pushes; ps.edges += one.edges;
        cost += (double)(one.pushes + one.edges);
        rsum_local = rsum_field * (1 - alpha);
        rmax_local /= 2.0;
        rounds++;
    }
    i64 num_random_walk = (i64)(omega_local * rsum_local);
    u64 key = walk_key(seed, 0), tstop = stop_threshold(alpha);
    i64 walks = 0, steps = 0;
    for (i64 v = 0; v < n; v++) {
        double rc = residue[v];
        if (!(rc > 0)) continue;  // zero residues give omega_i = 0 in the Java
        reserve[v] += rc * alpha;
        rc *= (1.0 - alpha);
        i64 omega_i = (i64)std::ceil(rc / rsum_local * (double)num_random_walk);
        double a_i = rc / rsum_local * (double)num_random_walk / (double)omega_i;
        double incr = a_i / (double)num_random_walk * rsum_local;
        i32 restart = restart_at_source ? s : (i32)v;
        for (i64 j = 0; j < omega_i; j++) {
            i32 t = walk(g, (i32)v, restart, (u64)j, key, tstop, true, &steps);
            reserve[t] += incr;
        }
        walks += omega_i > 0 ? omega_i : 0;
    }
    if (fs) { fs->pushes = ps.pushes; fs->edges = ps.edges; fs->rounds = rounds; fs->walks = walks; fs->steps = steps;
              fs->rmax_final = rmax_local * 2.0; fs->rsum = rsum_local; }
}

// J/Fora_Topk.java:102-184 (reading :136 as if the missing ';' were there).
void a_fora_topk(const Graph& g, i32 s, double alpha, double eps_in, i64 k, u64 seed, bool restart_at_source,
                 double* reserve_out, ForaStats* fs) {
    i64 n = g.n;
    Conf c = conf_topk(n, k);
    double epsilon = eps_in * 0.5;
    double delta_local = c.delta;
    double omega_local = (epsilon + 2.0) * std::log(2.0 / c.pfail) / epsilon / epsilon / c.delta;
    double min_rmax = epsilon * std::sqrt(c.min_delta / 3 / (double)g.m / std::log(2 / c.pfail));
    double rsum_local = 1.0;
    FwdTopkState S;
    S.residue.assign(n, 0.0); S.reserve.assign(n, 0.0); S.rsum = rsum_local;
    std::deque<i32> Q; Q.push_back(s);
    bool first = true;
    std::vector<double> reserve(n, 0.0);
    PushStats ps = {0, 0};
    i64 walks = 0, steps = 0; int round = 0;
    u64 tstop = stop_threshold(alpha);
    while (delta_local >= c.min_delta) {
        double rmax_local = epsilon * std::sqrt(delta_local / 3.0 / (double)g.m / std::log(2.0 / c.pfail));
        omega_local = (epsilon + 2.0) * std::log(2.0 / c.pfail) / epsilon / epsilon / delta_local;
        if (g.deg(s) == 0) { std::fill(reserve.begin(), reserve.end(), 0.0); reserve[s] = 1.0; rsum_local = 0.0; break; }
        rmax_local *= std::sqrt((double)g.m * rmax_local) * 3.0;
        a_forward_push_topk(g, s, alpha, S, Q, min_rmax, first, rmax_local, &ps);
        first = false;
        rsum_local = S.rsum;
        reserve = S.reserve;
        Q.assign(S.Q_next.begin(), S.Q_next.end());
        double rsum_random_walk = rsum_local * (1.0 - alpha);
        i64 num_random_walk = (i64)(omega_local * rsum_random_walk);
        u64 key = walk_key(seed, (u64)round);
        for (i64 v = 0; v < n; v++) {
            double rc = S.residue[v];
            if (!(rc > 0)) continue;
            i64 omega_i = (i64)std::ceil(rc * (double)num_random_walk);
            double a_i = rc * (double)num_random_walk / (double)omega_i;
            double incr = a_i / (double)num_random_walk;
            i32 restart = restart_at_source ? s : (i32)v;
            for (i64 j = 0; j < omega_i; j++) {
                i32 t = walk(g, (i32)v, restart, (u64)j, key, tstop, false, &steps);
                reserve[t] += incr;
            }
            walks += omega_i > 0 ? omega_i : 0;
        }
        round++;
        std::vector<double> vals;
        for (i64 v = 0; v < n; v++) if (reserve[v] > 0) vals.push_back(reserve[v]);
        double kth = 0.0;
        if (!kth_largest(vals, k, &kth)) kth = 0.0;
        if (kth >= (1 + epsilon) * delta_local || delta_local <= c.min_delta) break;
        else delta_local = std::max(c.min_delta, delta_local / 4.0);
    }
    std::copy(reserve.begin(), reserve.end(), reserve_out);
    if (fs) { fs->pushes = ps.pushes; fs->edges = ps.edges; fs->rounds = round; fs->walks = walks; fs->steps = steps;
              fs->rmax_final = 0; fs->rsum = rsum_local; }
}

// J/Backward_Search.java:38-100
void a_backward_search(Graph& g, i32 t, double alpha, double rmax, double* reserve, double* residue, PushStats* st) {
    g.build_csc();
    i64 n = g.n;
    std::fill(reserve, reserve + n, 0.0);
    std::fill(residue, residue + n, 0.0);
    if (g.indeg(t) == 0) { reserve[t] = 1.0; return; }
    std::vector<char> inq(n, 0);
    std::deque<i32> Q;
    Q.push_back(t); inq[t] = 1; residue[t] = 1.0;
    while (!Q.empty()) {
        i32 v = Q.front(); Q.pop_front();
        inq[v] = 0;
        double rc = residue[v];
        residue[v] = 0.0;
        reserve[v] += rc * alpha;
        double avg = (1.0 - alpha) * rc;
        if (st) { st->pushes++; st->edges += g.indeg(v); }
        for (i64 e = g.cp[v]; e < g.cp[v + 1]; e++) {
            i32 u = g.ri[e];
            double nr = residue[u] + avg / (double)g.deg(u);
            residue[u] = nr;
            if (nr > rmax && !inq[u]) { Q.push_back(u); inq[u] = 1; }
        }
    }
}

// ------------------------------------------------------------------ mode B
inline u64 thr_fwd(u64 rmax_fx, i64 d) {
    if (d == 0) return 1;
    unsigned __int128 p = (unsigned __int128)rmax_fx * (u64)d;
    return (p >> 64) ? ~0ull : (u64)p;
}
// Smallest usable threshold: below 1/alpha units mulhi64(x, alpha_q64) is 0, a push would move no mass into the
// reserve and a cycle could circulate forever.  (2^-62 / alpha ~ 1e-18: far below any meaningful rmax.)
inline u64 rmax_floor_fx(u64 alpha_q64) { return ~0ull / alpha_q64 + 2; }
inline u64 rmax_to_fx(double rmax, u64 alpha_q64) {
    const u64 lo = rmax_floor_fx(alpha_q64);
    double x = rmax * FX_SCALE;
    if (!(x >= 1.0)) return lo;
    if (x >= 18446744073709551615.0) return ~0ull;
    return std::max(lo, (u64)x);
}

struct Trace {  // frontier sequence of a synchronous push (ascending ids inside an iteration)
    std::vector<i64> iter_ptr;
    std::vector<i32> verts;
    Trace() { iter_ptr.push_back(0); }
};

struct BState {
    std::vector<u64> r, reserve;
    u64 reserve_total;
    i64 pushes, edges, iters;
    BState(i64 n) : r(n, 0), reserve(n, 0), reserve_total(0), pushes(0), edges(0), iters(0) {}
};

// One round: push synchronously until no vertex is above threshold for rmax_fx.
void b_push_round(const Graph& g, i32 s, u64 alpha_q64, u64 rmax_fx, BState& S, Trace* tr) {
    i64 n = g.n;
    std::vector<i32> F, next;
    for (i64 v = 0; v < n; v++) if (S.r[v] >= thr_fwd(rmax_fx, g.deg(v))) F.push_back((i32)v);
    std::vector<u64> q;
    while (!F.empty()) {
        if (tr) { tr->verts.insert(tr->verts.end(), F.begin(), F.end()); tr->iter_ptr.push_back((i64)tr->verts.size()); }
        q.resize(F.size());
        for (size_t i = 0; i < F.size(); i++) {  // phase A
            i32 v = F[i];
            u64 x = S.r[v];
            u64 a = mulhi64(x, alpha_q64);
            u64 y = x - a;
            i64 d = g.deg(v);
            u64 qq, left;
            if (d == 0) { qq = y; left = 0; } else { qq = y / (u64)d; left = y - qq * (u64)d; }
            S.r[v] = left;
            S.reserve[v] += a;
            S.reserve_total += a;
            q[i] = qq;
            S.pushes++; S.edges += d;
        }
        next.clear();
        u64 dang = 0;
        for (size_t i = 0; i < F.size(); i++) {  // phase B
            i32 v = F[i];
            i64 d = g.deg(v);
            if (d == 0) { dang += q[i]; continue; }
            for (i64 e = g.rp[v]; e < g.rp[v + 1]; e++) {
                i32 u = g.ci[e];
                u64 old = S.r[u], nw = old + q[i];
                S.r[u] = nw;
                u64 th = thr_fwd(rmax_fx, g.deg(u));
                if (old < th && nw >= th) next.push_back(u);
            }
        }
        if (dang > 0) {
            u64 old = S.r[s], nw = old + dang;
            S.r[s] = nw;
            u64 th = thr_fwd(rmax_fx, g.deg(s));
            if (old < th && nw >= th) next.push_back(s);
        }
        std::sort(next.begin(), next.end());
        F.swap(next);
        S.iters++;
    }
}

struct BForaStats { i64 pushes, edges, iters, rounds, walks, steps, walk_sources; u64 rmax_fx_final, rsum_fx; };

// Walk phase of FORA in fixed point (J/Fora_Whole_Graph.java:116-140).
void b_fora_walks(const Graph& g, i32 s, double alpha, u64 alpha_q64, double omega, u64 seed, u64 tag,
                  bool restart_at_source, BState& S, i64 walk_lo_frac_num, i64 walk_frac_den, BForaStats* fs,
                  bool add_push_part) {
    // walk_lo_frac_num/den select the shard [W*num/den, W*(num+1)/den) of the global walk index space
    // (multi-GPU replay); den == 1 runs everything.  add_push_part=false leaves out the non-walk terms
    // (alpha*r and dangling-start mass) so shards can be summed.
    i64 n = g.n;
    std::vector<i32> sv; std::vector<u64> sy;
    u64 rsum_fx = 0;
    for (i64 v = 0; v < n; v++) {
        u64 x = S.r[v];
        if (x == 0) continue;
        u64 a = mulhi64(x, alpha_q64);
        if (add_push_part) { S.reserve[v] += a; }
        u64 y = x - a;
        S.r[v] = 0;
        sv.push_back((i32)v); sy.push_back(y);
        rsum_fx += y;
    }
    double rsum_d = (double)rsum_fx * FX_INV;
    i64 N = (i64)(omega * rsum_d);
    u64 key = walk_key(seed, tag), tstop = stop_threshold(alpha);
    // first pass: per-source walk counts and the global prefix
    std::vector<i64> w(sv.size(), 0);
    i64 W = 0;
    for (size_t i = 0; i < sv.size(); i++) {
        if (N <= 0) continue;
        if (g.deg(sv[i]) == 0) { if (add_push_part) S.reserve[sv[i]] += sy[i]; continue; }
        double yd = (double)sy[i] * FX_INV;
        double t1 = yd / rsum_d;
        double t2 = t1 * (double)N;
        w[i] = (i64)std::ceil(t2);
        W += w[i];
    }
    i64 lo = (i64)(((__int128)W * walk_lo_frac_num) / walk_frac_den);
    i64 hi = (i64)(((__int128)W * (walk_lo_frac_num + 1)) / walk_frac_den);
    i64 base = 0, steps = 0, nsrc = 0;
    for (size_t i = 0; i < sv.size(); i++) {
        if (w[i] == 0) continue;
        nsrc++;
        u64 inc = sy[i] / (u64)w[i], rem = sy[i] - inc * (u64)w[i];
        i32 restart = restart_at_source ? s : sv[i];
        i64 j0 = std::max<i64>(0, lo - base), j1 = std::min<i64>(w[i], hi - base);
        for (i64 j = j0; j < j1; j++) {
            i32 t = walk(g, sv[i], restart, (u64)j, key, tstop, true, &steps);
            S.reserve[t] += inc + (j == 0 ? rem : 0);
        }
        base += w[i];
    }
    if (fs) { fs->walks = W; fs->steps = steps; fs->rsum_fx = rsum_fx; fs->walk_sources = nsrc; }
}

void b_fora(const Graph& g, i32 s, double alpha, double eps, u64 seed, double walk_cost, int max_rounds,
            bool restart_at_source, u64* reserve_out, BForaStats* fs, Trace* tr, i64 shard, i64 nshards) {
    i64 n = g.n;
    Conf c = conf_whole(n);
    double L = std::log(2.0 / c.pfail);
    double rmax = eps * std::sqrt(c.delta / 3.0 / (double)g.m / L) / (1.0 - alpha);
    double omega = (eps + 2.0) * L / eps / eps / c.delta;
    u64 aq = alpha_q64_of(alpha);
    BState S(n);
    BForaStats st; memset(&st, 0, sizeof st);
    if (g.deg(s) == 0) {
        if (shard == 0) S.reserve[s] = FX_ONE;
    } else {
        S.r[s] = FX_ONE;
        int rounds = 0;
        u64 rfx = 0;
        for (;;) {
            rfx = rmax_to_fx(rmax, aq);
            b_push_round(g, s, aq, rfx, S, tr);
            u64 rsum_fx = FX_ONE - S.reserve_total;
            double rsum_w = (double)rsum_fx * FX_INV * (1.0 - alpha);
            rmax = rmax * 0.5;
            rounds++;
            double cost = (double)(S.pushes + S.edges);
            if (!(cost < walk_cost * rsum_w * omega) || rounds >= max_rounds) break;
        }
        st.rounds = rounds; st.rmax_fx_final = rfx;
        if (shard != 0) std::fill(S.reserve.begin(), S.reserve.end(), 0);
        b_fora_walks(g, s, alpha, aq, omega, seed, 0, restart_at_source, S, shard, nshards, &st, shard == 0);
    }
    st.pushes = S.pushes; st.edges = S.edges; st.iters = S.iters;
    std::copy(S.reserve.begin(), S.reserve.end(), reserve_out);
    if (fs) *fs = st;
}

// FORA top-k in fixed point (J/Fora_Topk.java:102-184).  The push of each round is the synchronous
// push to convergence at that round's rmax (equivalent frontier set to the Java's Q/Q_next bookkeeping
// whenever rmax >= min_rmax).
// Sharded form (ppr_fora_topk_shard): this shard runs the walks whose global index (sources in ascending id, then
// walk number) lies in [W*shard/nshards, W*(shard+1)/nshards); only shard 0's estimate starts from the push
// reserve; `reduce` sums the estimate vector over the shards once per round, after which every shard selects and
// decides on identical data.
typedef int (*orc_reduce_fn)(void* user, u64* est, i64 n);
void b_fora_topk(const Graph& g, i32 s, double alpha, double eps_in, i64 k, u64 seed, bool restart_at_source,
                 u64* est_out, BForaStats* fs, i64 shard = 0, i64 nshards = 1, orc_reduce_fn reduce = nullptr, void* user = nullptr) {
    i64 n = g.n;
    Conf c = conf_topk(n, k);
    double epsilon = eps_in * 0.5;
    double delta_local = c.delta;
    double L = std::log(2.0 / c.pfail);
    u64 aq = alpha_q64_of(alpha), tstop = stop_threshold(alpha);
    BState S(n);
    std::vector<u64> est(n, 0);
    BForaStats st; memset(&st, 0, sizeof st);
    int round = 0;
    if (g.deg(s) == 0) {
        if (shard == 0) est[s] = FX_ONE;
        if (nshards > 1 && reduce) reduce(user, est.data(), n);
    } else {
        S.r[s] = FX_ONE;
        while (delta_local >= c.min_delta) {
            double rmax_local = epsilon * std::sqrt(delta_local / 3.0 / (double)g.m / L);
            double omega_local = (epsilon + 2.0) * L / epsilon / epsilon / delta_local;
            rmax_local *= std::sqrt((double)g.m * rmax_local) * 3.0;
            u64 rfx = rmax_to_fx(rmax_local, aq);
            b_push_round(g, s, aq, rfx, S, nullptr);
            st.rmax_fx_final = rfx;
            u64 rsum_fx = FX_ONE - S.reserve_total;
            double rsum_local = (double)rsum_fx * FX_INV;
            double rsum_random_walk = rsum_local * (1.0 - alpha);
            i64 N = (i64)(omega_local * rsum_random_walk);
            if (shard == 0) est = S.reserve; else std::fill(est.begin(), est.end(), 0);
            u64 key = walk_key(seed, (u64)round);
            // first pass: walk counts per source and their total (the shard boundaries are fractions of it)
            std::vector<i64> wv(n, 0);
            i64 W = 0;
            for (i64 v = 0; v < n; v++) {
                u64 x = S.r[v];
                if (x == 0 || N <= 0) continue;
                double xd = (double)x * FX_INV;
                i64 w = (i64)std::ceil(xd * (double)N);
                if (w <= 0) continue;
                if (g.deg(v) == 0) { if (shard == 0) est[v] += x; continue; }
                wv[v] = w; W += w;
            }
            i64 lo = (i64)(((__int128)W * shard) / nshards), hi = (i64)(((__int128)W * (shard + 1)) / nshards);
            i64 base = 0;
            for (i64 v = 0; v < n; v++) {
                i64 w = wv[v];
                if (w == 0) continue;
                u64 x = S.r[v];
                u64 inc = x / (u64)w, rem = x - inc * (u64)w;
                i32 restart = restart_at_source ? s : (i32)v;
                i64 j0 = std::max<i64>(0, lo - base), j1 = std::min<i64>(w, hi - base);
                for (i64 j = j0; j < j1; j++) {
                    i32 t = walk(g, (i32)v, restart, (u64)j, key, tstop, false, &st.steps);
                    est[t] += inc + (j == 0 ? rem : 0);
                }
                if (j1 > j0) st.walks += j1 - j0;
                st.walk_sources++;
                base += w;
            }
            if (nshards > 1 && reduce) reduce(user, est.data(), n);
            round++;
            std::vector<u64> vals;
            for (i64 v = 0; v < n; v++) if (est[v] > 0) vals.push_back(est[v]);
            u64 kth = 0;
            if (!kth_largest_u64(vals, k, &kth)) kth = 0;
            double kth_d = (double)kth * FX_INV;
            if (kth_d >= (1 + epsilon) * delta_local || delta_local <= c.min_delta) break;
            else delta_local = std::max(c.min_delta, delta_local / 4.0);
        }
    }
    st.rounds = round; st.pushes = S.pushes; st.edges = S.edges; st.iters = S.iters;
    st.rsum_fx = FX_ONE - S.reserve_total;
    std::copy(est.begin(), est.end(), est_out);
    if (fs) *fs = st;
}

// Reverse push in fixed point (J/Backward_Search.java:38-100), synchronous schedule.
//   phase A: x = r[v]; r[v] = 0; a = mulhi64(x, alpha_q64); reserve[v] += a; y = x - a
//   phase B: every in-neighbour u gets inc = mulhi64(y, invdeg[u]), invdeg[u] = (2^64-1) / d_out(u)
//   frontier: r[u] > rmax_fx (strict, absolute), rmax_fx = (uint64)(rmax * 2^62)
void b_backward(Graph& g, const std::vector<u64>& invdeg, i32 t, u64 aq, u64 rmax_fx, u64* reserve, u64* r,
                PushStats* st, i64* iters) {
    i64 n = g.n;
    std::fill(reserve, reserve + n, 0);
    std::fill(r, r + n, 0);
    if (g.indeg(t) == 0) { reserve[t] = FX_ONE; return; }
    r[t] = FX_ONE;
    std::vector<i32> F(1, t), next;
    std::vector<u64> y;
    while (!F.empty()) {
        y.resize(F.size());
        for (size_t i = 0; i < F.size(); i++) {
            i32 v = F[i];
            u64 x = r[v]; r[v] = 0;
            u64 a = mulhi64(x, aq);
            reserve[v] += a;
            y[i] = x - a;
            if (st) { st->pushes++; st->edges += g.indeg(v); }
        }
        next.clear();
        for (size_t i = 0; i < F.size(); i++) {
            i32 v = F[i];
            for (i64 e = g.cp[v]; e < g.cp[v + 1]; e++) {
                i32 u = g.ri[e];
                u64 inc = mulhi64(y[i], invdeg[u]);
                u64 old = r[u], nw = old + inc;
                r[u] = nw;
                if (old <= rmax_fx && nw > rmax_fx) next.push_back(u);
            }
        }
        std::sort(next.begin(), next.end());
        F.swap(next);
        if (iters) (*iters)++;
    }
}

inline u64 thr_ge_fx(double threshold) {  // smallest FX value whose real value is >= threshold
    if (!(threshold > 0)) return 0;
    double x = std::ceil(threshold * FX_SCALE);
    if (x >= 18446744073709551615.0) return ~0ull;
    return (u64)x;
}
inline u64 rmax_abs_fx(double rmax, u64 alpha_q64) {  // largest FX value whose real value is <= rmax (floored, see rmax_floor_fx)
    const u64 lo = rmax_floor_fx(alpha_q64);
    if (!(rmax > 0)) return lo;
    double x = std::floor(rmax * FX_SCALE);
    if (x >= 18446744073709551615.0) return ~0ull;
    return std::max(lo, (u64)x);
}

}  // namespace

// =============================================================================================
// C ABI (ctypes)
// =============================================================================================
extern "C" {

void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    Philox4 o = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    memcpy(out, o.x, 16);
}

// conf: which = 0 whole-graph (MC/FWD/FORA), 1 = FORA top-k.  out = {delta, pfail, min_delta}
void orc_conf(int which, i64 n, i64 k, double* out) {
    Conf c = which ? conf_topk(n, k) : conf_whole(n);
    out[0] = c.delta; out[1] = c.pfail; out[2] = c.min_delta;
}
// out = {rmax, omega} per J/Fora_Whole_Graph.java:86-87
void orc_fora_params(i64 n, i64 m, double alpha, double eps, double* out) {
    Conf c = conf_whole(n);
    out[0] = eps * std::sqrt(c.delta / 3.0 / (double)m / std::log(2.0 / c.pfail)) / (1.0 - alpha);
    out[1] = (eps + 2.0) * std::log(2.0 / c.pfail) / eps / eps / c.delta;
}

int orc_power_method(i64 n, i64 m, const i64* rp, const i32* ci, i32 s, double alpha, int iters, double tol,
                     double* out) {
    Graph g(n, m, rp, ci);
    int done = 0;
    power_method(g, s, alpha, iters, tol, out, &done);
    return done;
}

i32 orc_walk(i64 n, i64 m, const i64* rp, const i32* ci, i32 start, i32 restart, u64 j, u64 seed, u64 tag, double alpha,
             int no_zero_hop, i64* steps) {
    Graph g(n, m, rp, ci);
    i64 st = 0;
    i32 t = walk(g, start, restart, j, walk_key(seed, tag), stop_threshold(alpha), no_zero_hop != 0, &st);
    if (steps) *steps = st;
    return t;
}

// ---- mode A
void orc_a_forward_push(i64 n, i64 m, const i64* rp, const i32* ci, i32 s, double alpha, double rmax, double* reserve,
                        double* residue, double* rsum, i64* stats /*pushes, edges*/) {
    Graph g(n, m, rp, ci);
    PushStats st = {0, 0};
    double rs = 1.0;
    a_forward_push(g, s, alpha, rmax, reserve, residue, &rs, &st);
    if (rsum) *rsum = rs;
    if (stats) { stats[0] = st.pushes; stats[1] = st.edges; }
}
i64 orc_a_monte_carlo(i64 n, i64 m, const i64* rp, const i32* ci, i32 s, double alpha, double eps, u64 seed, double* ppr,
                      i64* steps) {
    Graph g(n, m, rp, ci);
    i64 st = 0;
    i64 w = a_monte_carlo(g, s, alpha, eps, seed, ppr, &st);
    if (steps) *steps = st;
    return w;
}
// stats: pushes, edges, rounds, walks, steps ; dstats: rmax_final, rsum
void orc_a_fora(i64 n, i64 m, const i64* rp, const i32* ci, i32 s, double alpha, double eps, u64 seed, double walk_cost,
                int max_rounds, int restart_at_source, double* reserve, i64* stats, double* dstats) {
    Graph g(n, m, rp, ci);
    ForaStats fs; memset(&fs, 0, sizeof fs);
    a_fora(g, s, alpha, eps, seed, walk_cost, max_rounds, restart_at_source != 0, reserve, &fs);
    if (stats) { stats[0] = fs.pushes; stats[1] = fs.edges; stats[2] = fs.rounds; stats[3] = fs.walks; stats[4] = fs.steps; }
    if (dstats) { dstats[0] = fs.rmax_final; dstats[1] = fs.rsum; }
}
void orc_a_fora_topk(i64 n, i64 m, const i64* rp, const i32* ci, i32 s, double alpha, double eps, i64 k, u64 seed,
                     int restart_at_source, double* reserve, i64* stats) {
    Graph g(n, m, rp, ci);
    ForaStats fs; memset(&fs, 0, sizeof fs);
    a_fora_topk(g, s, alpha, eps, k, seed, restart_at_source != 0, reserve, &fs);
    if (stats) { stats[0] = fs.pushes; stats[1] = fs.edges; stats[2] = fs.rounds; stats[3] = fs.walks; stats[4] = fs.steps; }
}
void orc_a_backward_search(i64 n, i64 m, const i64* rp, const i32* ci, i32 t, double alpha, double rmax, double* reserve,
                           i64* stats) {
    Graph g(n, m, rp, ci);
    std::vector<double> residue(n);
    PushStats st = {0, 0};
    a_backward_search(g, t, alpha, rmax, reserve, residue.data(), &st);
    if (stats) { stats[0] = st.pushes; stats[1] = st.edges; }
}

// All-pair backward search (J/Base_Whole_Graph.java:58-164), compute part.  Two-call protocol:
// call with cap = 0 to get the pair count, then with buffers.  Output is the inverted index grouped
// by source v: src_ptr[n+1], tgt[], val[]; inside a source, targets ascend (k < 0) or values descend
// with ties in ascending target (k >= 0, keeps everything >= the k-th largest).
// target range [t_lo, t_hi) lets a caller replay a multi-GPU block partition.
i64 orc_a_base_all_pairs(i64 n, i64 m, const i64* rp, const i32* ci, double alpha, double threshold, i64 k, i32 t_lo,
                         i32 t_hi, i64* src_ptr, i32* tgt, double* val, i64 cap, i64* stats) {
    Graph g(n, m, rp, ci);
    g.build_csc();
    std::vector<std::vector<std::pair<i32, double>>> inv(n);
    std::vector<double> reserve(n), residue(n);
    PushStats st = {0, 0};
    for (i32 t = t_lo; t < t_hi; t++) {
        a_backward_search(g, t, alpha, threshold, reserve.data(), residue.data(), &st);
        for (i64 v = 0; v < n; v++)
            if (reserve[v] > 0 && reserve[v] >= threshold) inv[v].push_back(std::make_pair(t, reserve[v]));
    }
    i64 total = 0;
    for (i64 v = 0; v < n; v++) {
        auto& L = inv[v];
        if (k >= 0) {
            std::vector<double> vals; for (auto& p : L) vals.push_back(p.second);
            double kth; bool have = kth_largest(vals, k, &kth);
            std::vector<std::pair<i32, double>> keep;
            for (auto& p : L) if (!have || p.second >= kth) keep.push_back(p);
            std::stable_sort(keep.begin(), keep.end(), [](const std::pair<i32, double>& a, const std::pair<i32, double>& b) { return a.second > b.second; });
            L.swap(keep);
        }
        total += (i64)L.size();
    }
    if (stats) { stats[0] = st.pushes; stats[1] = st.edges; }
    if (cap < total || !src_ptr) return total;
    i64 pos = 0;
    for (i64 v = 0; v < n; v++) {
        src_ptr[v] = pos;
        for (auto& p : inv[v]) { tgt[pos] = p.first; val[pos] = p.second; pos++; }
    }
    src_ptr[n] = pos;
    return total;
}

// ---- selection and metrics
int orc_kth_ppr(const double* vals, i64 cnt, i64 k, double* out) {
    std::vector<double> v(vals, vals + cnt);
    double x = 0; bool ok = kth_largest(v, k, &x);
    *out = x;
    return ok ? 1 : 0;
}
i64 orc_topk_select_f64(const double* vals, i64 n, i64 k, i32* ids_out, i64 cap) { return topk_select<double>(vals, n, k, ids_out, cap); }
i64 orc_topk_select_u64(const u64* vals, i64 n, i64 k, i32* ids_out, i64 cap) { return topk_select<u64>(vals, n, k, ids_out, cap); }

// J/Gen_Util.java:259-326.  out = {precision, NDCG}
void orc_topk_metrics(const double* gnd /*n, exact ppr*/, i64 n, i64 k, const i32* algo_ids, i64 algo_cnt, double* out) {
    std::vector<i32> gids(n);
    i64 gc = topk_select<double>(gnd, n, k, gids.data(), n);
    gids.resize(gc);
    std::vector<char> in_g(n, 0);
    for (i32 v : gids) in_g[v] = 1;
    double hit = 0;
    for (i64 i = 0; i < algo_cnt; i++) if (in_g[algo_ids[i]]) hit += 1;
    out[0] = gc ? hit / (double)gc : 0.0;
    double Zk = 0, DCG = 0;
    for (i64 i = 1; i <= gc; i++) Zk += (std::pow(2.0, gnd[gids[i - 1]]) - 1.0) / std::log(i + 1.0) / std::log(2.0);
    for (i64 i = 1; i <= algo_cnt; i++) {
        double p = in_g[algo_ids[i - 1]] ? gnd[algo_ids[i - 1]] : 0.0;
        DCG += (std::pow(2.0, p) - 1.0) / std::log(i + 1.0) / std::log(2.0);
    }
    out[1] = DCG / Zk;
}
double orc_max_err(const double* gnd, const double* est, i64 n) {
    double mx = 0;
    for (i64 v = 0; v < n; v++) if (gnd[v] > 0) mx = std::max(mx, std::fabs(est[v] - gnd[v]));  // loops over ground-truth entries
    return mx;
}

// ---- mode B
// Trace protocol: iter_ptr (cap_iters+1) and verts (cap_verts) are filled as far as they fit;
// trace_sizes = {n_iters, n_verts} needed.
void orc_b_forward_push(i64 n, i64 m, const i64* rp, const i32* ci, i32 s, double alpha, double rmax, u64* reserve,
                        u64* residue, i64* stats /*pushes, edges, iters*/, i64* iter_ptr, i64 cap_iters, i32* verts,
                        i64 cap_verts, i64* trace_sizes) {
    Graph g(n, m, rp, ci);
    BState S(n);
    Trace tr;
    if (g.deg(s) == 0) { S.reserve[s] = FX_ONE; }
    else { S.r[s] = FX_ONE; b_push_round(g, s, alpha_q64_of(alpha), rmax_to_fx(rmax, alpha_q64_of(alpha)), S, &tr); }
    std::copy(S.reserve.begin(), S.reserve.end(), reserve);
    if (residue) std::copy(S.r.begin(), S.r.end(), residue);
    if (stats) { stats[0] = S.pushes; stats[1] = S.edges; stats[2] = S.iters; }
    i64 ni = (i64)tr.iter_ptr.size() - 1, nv = (i64)tr.verts.size();
    if (trace_sizes) { trace_sizes[0] = ni; trace_sizes[1] = nv; }
    if (iter_ptr) for (i64 i = 0; i <= ni && i <= cap_iters; i++) iter_ptr[i] = tr.iter_ptr[i];
    if (verts) for (i64 i = 0; i < nv && i < cap_verts; i++) verts[i] = tr.verts[i];
}

i64 orc_b_monte_carlo(i64 n, i64 m, const i64* rp, const i32* ci, i32 s, double alpha, double eps, u64 seed, i64 shard,
                      i64 nshards, u64* counts, double* omega_out, i64* steps) {
    Graph g(n, m, rp, ci);
    Conf c = conf_whole(n);
    double omega = 3 * std::log(2 / c.pfail) / eps / eps / c.delta;
    i64 W = (i64)std::floor(omega);
    if (omega_out) *omega_out = omega;
    std::fill(counts, counts + n, 0);
    u64 key = walk_key(seed, 0), tstop = stop_threshold(alpha);
    i64 lo = (i64)(((__int128)W * shard) / nshards), hi = (i64)(((__int128)W * (shard + 1)) / nshards);
    i64 st = 0;
    for (i64 j = lo; j < hi; j++) counts[walk(g, s, s, (u64)j, key, tstop, false, &st)]++;
    if (steps) *steps = st;
    return W;
}

// stats: pushes, edges, iters, rounds, walks, steps, walk_sources, rmax_fx_final, rsum_fx
void orc_b_fora(i64 n, i64 m, const i64* rp, const i32* ci, i32 s, double alpha, double eps, u64 seed, double walk_cost,
                int max_rounds, int restart_at_source, i64 shard, i64 nshards, u64* reserve, i64* stats) {
    Graph g(n, m, rp, ci);
    BForaStats fs;
    b_fora(g, s, alpha, eps, seed, walk_cost, max_rounds, restart_at_source != 0, reserve, &fs, nullptr, shard, nshards);
    if (stats) { stats[0] = fs.pushes; stats[1] = fs.edges; stats[2] = fs.iters; stats[3] = fs.rounds; stats[4] = fs.walks;
                 stats[5] = fs.steps; stats[6] = fs.walk_sources; stats[7] = (i64)fs.rmax_fx_final; stats[8] = (i64)fs.rsum_fx; }
}
void orc_b_fora_topk(i64 n, i64 m, const i64* rp, const i32* ci, i32 s, double alpha, double eps, i64 k, u64 seed,
                     int restart_at_source, u64* est, i64* stats) {
    Graph g(n, m, rp, ci);
    BForaStats fs;
    b_fora_topk(g, s, alpha, eps, k, seed, restart_at_source != 0, est, &fs);
    if (stats) { stats[0] = fs.pushes; stats[1] = fs.edges; stats[2] = fs.iters; stats[3] = fs.rounds; stats[4] = fs.walks;
                 stats[5] = fs.steps; stats[6] = fs.walk_sources; stats[7] = (i64)fs.rmax_fx_final; stats[8] = (i64)fs.rsum_fx; }
}
void orc_b_fora_topk_shard(i64 n, i64 m, const i64* rp, const i32* ci, i32 s, double alpha, double eps, i64 k, u64 seed,
                           int restart_at_source, i64 shard, i64 nshards, orc_reduce_fn reduce, void* user, u64* est, i64* stats) {
    Graph g(n, m, rp, ci);
    BForaStats fs;
    b_fora_topk(g, s, alpha, eps, k, seed, restart_at_source != 0, est, &fs, shard, nshards, reduce, user);
    if (stats) { stats[0] = fs.pushes; stats[1] = fs.edges; stats[2] = fs.iters; stats[3] = fs.rounds; stats[4] = fs.walks;
                 stats[5] = fs.steps; stats[6] = fs.walk_sources; stats[7] = (i64)fs.rmax_fx_final; stats[8] = (i64)fs.rsum_fx; }
}
void orc_b_backward_search(i64 n, i64 m, const i64* rp, const i32* ci, i32 t, double alpha, double rmax, u64* reserve,
                           u64* residue, i64* stats /*pushes, edges, iters*/) {
    Graph g(n, m, rp, ci);
    g.build_csc();
    std::vector<u64> invdeg(n, 0);
    for (i64 v = 0; v < n; v++) if (g.deg(v) > 0) invdeg[v] = ~0ull / (u64)g.deg(v);
    std::vector<u64> r(n);
    PushStats st = {0, 0}; i64 it = 0;
    b_backward(g, invdeg, t, alpha_q64_of(alpha), rmax_abs_fx(rmax, alpha_q64_of(alpha)), reserve, r.data(), &st, &it);
    if (residue) std::copy(r.begin(), r.end(), residue);
    if (stats) { stats[0] = st.pushes; stats[1] = st.edges; stats[2] = it; }
}
// All pairs, fixed point.  Same two-call protocol / ordering as orc_a_base_all_pairs with k < 0
// (ascending target inside each source); val is the FX reserve.
i64 orc_b_base_all_pairs(i64 n, i64 m, const i64* rp, const i32* ci, double alpha, double threshold, i32 t_lo, i32 t_hi,
                         i64* src_ptr, i32* tgt, u64* val, i64 cap, i64* stats) {
    Graph g(n, m, rp, ci);
    g.build_csc();
    std::vector<u64> invdeg(n, 0);
    for (i64 v = 0; v < n; v++) if (g.deg(v) > 0) invdeg[v] = ~0ull / (u64)g.deg(v);
    u64 aq = alpha_q64_of(alpha), rfx = rmax_abs_fx(threshold, aq), tfx = thr_ge_fx(threshold);
    std::vector<std::vector<std::pair<i32, u64>>> inv(n);
    std::vector<u64> reserve(n), r(n);
    PushStats st = {0, 0}; i64 it = 0;
    i64 total = 0;
    for (i32 t = t_lo; t < t_hi; t++) {
        b_backward(g, invdeg, t, aq, rfx, reserve.data(), r.data(), &st, &it);
        for (i64 v = 0; v < n; v++)
            if (reserve[v] > 0 && reserve[v] >= tfx) { inv[v].push_back(std::make_pair(t, reserve[v])); total++; }
    }
    if (stats) { stats[0] = st.pushes; stats[1] = st.edges; stats[2] = it; }
    if (cap < total || !src_ptr) return total;
    i64 pos = 0;
    for (i64 v = 0; v < n; v++) {
        src_ptr[v] = pos;
        for (auto& p : inv[v]) { tgt[pos] = p.first; val[pos] = p.second; pos++; }
    }
    src_ptr[n] = pos;
    return total;
}

}  // extern "C"
