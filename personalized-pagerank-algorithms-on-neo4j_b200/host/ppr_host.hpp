// ppr_host.hpp -- C++ host side above the C ABI: the reference's driver, harness and algorithm classes
// (same names, same public methods, same option names and result-file format) re-hosted over
// libpprb200.so.  The reference's host language is Java (J/ = src/main/java/joezie/fora_neo4j/ of
// joezie/Personalized-PageRank-Algorithms-on-Neo4j); this image has no JDK, so the compiled host layer
// is C++ (this file) and the Java binding is shipped as source (java/).  Every class
// below is a thin shell: all arithmetic happens in the CUDA library, nothing here can fall back to
// the CPU.  `graphDb` arguments of the Java constructors have no counterpart (Neo4j itself is out
// of scope); node ids are the mapped ids (for GOT identical to the Neo4j ids).
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/pprb200.h"

namespace fora_neo4j {

typedef std::unordered_map<int64_t, double> PprMap;  // HashMap<Long, Double> of the reference

inline void check(int32_t rc, ppr_ctx* ctx) {
    if (rc != PPR_OK) throw std::runtime_error(std::string("pprb200 error ") + std::to_string(rc) + ": " + ppr_last_error(ctx));
}

// Graph adjM (org.neo4j.graphalgo.api.Graph in the reference): the device-resident topology + the
// accessors the harness needs (nodeCount, degree).
struct Graph {
    ppr_ctx* ctx = nullptr;
    ppr_graph* g = nullptr;
    int64_t n = 0, m = 0;
    std::vector<int64_t> row_ptr;
    std::vector<std::string> names;
    int64_t nodeCount() const { return n; }
    int64_t degree(int64_t v) const { return row_ptr[v + 1] - row_ptr[v]; }
    ~Graph() { if (g) ppr_graph_destroy(g); }
};

inline PprMap to_map(const std::vector<double>& dense) {
    PprMap out;
    for (size_t v = 0; v < dense.size(); v++) if (dense[v] > 0) out.emplace((int64_t)v, dense[v]);
    return out;
}

// Double.toString-compatible formatting (shortest digits that round-trip, Java's E notation)
inline std::string java_double(double x) {
    if (x == 0) return "0.0";
    char buf[64];
    int prec = 1;
    for (; prec <= 17; prec++) { snprintf(buf, sizeof buf, "%.*e", prec - 1, x); if (std::strtod(buf, nullptr) == x) break; }
    std::string s(buf);
    size_t e = s.find('e');
    std::string mant = s.substr(0, e);
    int ex = std::atoi(s.c_str() + e + 1);
    bool neg = mant[0] == '-';
    if (neg) mant = mant.substr(1);
    std::string digits;
    for (char c : mant) if (c != '.') digits.push_back(c);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    std::string out = neg ? "-" : "";
    double ax = std::fabs(x);
    if (ax >= 1e-3 && ax < 1e7) {
        if (ex >= 0) {
            std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)ex + 1));
            while ((int)ip.size() < ex + 1) ip.push_back('0');
            std::string fp = digits.size() > (size_t)ex + 1 ? digits.substr(ex + 1) : "0";
            out += ip + "." + fp;
        } else {
            out += "0." + std::string((size_t)(-ex - 1), '0') + digits;
        }
    } else {
        out += digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(ex);
    }
    return out;
}

// ------------------------------------------------------------------ interfaces (J/*_Interface.java)
struct Whole_Graph_Util_Interface {
    virtual ~Whole_Graph_Util_Interface() {}
    virtual void computeWholeGraphPPR(int64_t nodeId_start, double param) = 0;
    virtual const PprMap& getWholeGraphPPR() = 0;
    virtual const std::vector<double>& dense() = 0;  // dense image of the same result (error metrics)
};
struct Topk_Util_Interface {
    virtual ~Topk_Util_Interface() {}
    virtual void computeTopKPPR(int64_t nodeId_start, int k, double param) = 0;
    virtual const std::vector<int64_t>& getTopKNodeIds(int k) = 0;
};

// J/Algo_Util.java: kth_ppr + retrieveTopK/getTopKNodeIds over a dense vector (device radix select)
struct Algo_Util {
    std::string node_property;
    explicit Algo_Util(const std::string& p = "name") : node_property(p) {}
    static void topk(Graph& G, const std::vector<double>& vec, int k, std::vector<int64_t>& ids, std::vector<double>& vals, double* kth) {
        std::vector<int32_t> i32(vec.size());
        std::vector<double> v(vec.size());
        int32_t cnt = 0;
        double kk = 0;
        check(ppr_topk(G.ctx, vec.data(), (int64_t)vec.size(), k, i32.data(), v.data(), (int32_t)vec.size(), &cnt, &kk), G.ctx);
        ids.assign(i32.begin(), i32.begin() + cnt);
        vals.assign(v.begin(), v.begin() + cnt);
        if (kth) *kth = kk;
    }
    // J/Algo_Util.java:20-30: the node-property column of the export stands for graphDb; an id without a name prints as itself
    static std::string getNodeName(int64_t nodeId, const Graph& G) {
        return (nodeId >= 0 && (size_t)nodeId < G.names.size()) ? G.names[(size_t)nodeId] : std::to_string(nodeId);
    }
    // The print loops of the reference classes (e.g. J/Forward_Push.java:268-281, :366-391): value-descending (stable, ties keep
    // ascending id), "@name \t Double.toString(value)", at most `limit` rows; `extra` adds the residue column of Forward_Push.
    typedef std::vector<std::pair<int64_t, double>> Rows;
    static Rows rows_of(const std::vector<double>& dense) {
        Rows r;
        for (size_t v = 0; v < dense.size(); v++) if (dense[v] > 0) r.emplace_back((int64_t)v, dense[v]);
        return r;
    }
    static Rows rows_of(const std::vector<int64_t>& ids, const std::vector<double>& dense) {
        Rows r;
        for (int64_t v : ids) r.emplace_back(v, dense[(size_t)v]);
        return r;
    }
    static void print_rows(const Graph& G, const std::string& header, Rows rows, size_t limit, bool sorted = true, const std::vector<double>* extra = nullptr) {
        if (sorted) std::stable_sort(rows.begin(), rows.end(), [](const std::pair<int64_t, double>& a, const std::pair<int64_t, double>& b) { return a.second > b.second; });
        std::printf("%s\n", header.c_str());
        size_t count = 0;
        for (const auto& r : rows) {
            if (count >= limit) break;
            count++;
            std::printf("@%s\t%s", getNodeName(r.first, G).c_str(), java_double(r.second).c_str());
            if (extra) std::printf("\t%s", java_double((*extra)[(size_t)r.first]).c_str());
            std::printf("\n");
        }
    }
    // J/Algo_Util.java:32-50: the k-th largest value; false (Java: null) when k is outside 1..size
    static bool kth_ppr(Graph& G, const std::vector<double>& pprs, int k, double* out) {
        if (k < 1 || (size_t)k > pprs.size()) return false;
        std::vector<int64_t> ids; std::vector<double> vals;
        topk(G, pprs, k, ids, vals, out);
        return true;
    }
};

// Preprocessing_Interface (J/Preprocessing_Interface.java) for the single-source algorithms: one text file per source,
// "<ALGO>_ppr_results/<db>/<param>/<nodeId>.txt", lines "id \t Double.toString(value)" (J/Forward_Push.java:289-363, :432-443)
struct Prep_Store {
    std::string preprocessing_dirName;
    explicit Prep_Store(const std::string& root) : preprocessing_dirName(root) {}
    virtual ~Prep_Store() {}
    virtual void prep_query(int64_t s, double param) = 0;            // computeWholeGraphPPR
    virtual const std::vector<double>& prep_result() = 0;            // its dense result
    virtual void prep_load(const std::vector<double>& dense) = 0;    // readPreprocessedPPR target
    void preprocessing(double /*threshold, unused as in the reference*/, double param, int64_t node_amount) {
        namespace fs = std::filesystem;
        preprocessing_dirName += "/" + java_double(param);
        std::error_code ec;
        fs::remove_all(preprocessing_dirName, ec);
        fs::create_directories(preprocessing_dirName, ec);
        for (int64_t v = 0; v < node_amount; v++) {
            prep_query(v, param);
            std::ofstream f(preprocessing_dirName + "/" + std::to_string(v) + ".txt");
            const std::vector<double>& r = prep_result();
            for (size_t t = 0; t < r.size(); t++) if (r[t] > 0) f << t << '\t' << java_double(r[t]) << '\n';
        }
    }
    void readPreprocessedPPR(int64_t s, int64_t node_amount) {
        std::vector<double> dense((size_t)node_amount, 0.0);
        std::ifstream f(preprocessing_dirName + "/" + std::to_string(s) + ".txt");
        std::string line;
        while (std::getline(f, line)) {
            size_t tab = line.find('\t');
            if (tab == std::string::npos) continue;
            dense[(size_t)std::stoll(line.substr(0, tab))] = std::stod(line.substr(tab + 1));
        }
        prep_load(dense);
    }
    int64_t getPrepSize() const {
        namespace fs = std::filesystem;
        int64_t bytes = 0;
        std::error_code ec;
        for (auto it = fs::recursive_directory_iterator(preprocessing_dirName, ec); !ec && it != fs::recursive_directory_iterator(); it.increment(ec))
            if (it->is_regular_file(ec)) bytes += (int64_t)it->file_size(ec);
        return bytes;
    }
    void deletePrepDir() { std::error_code ec; std::filesystem::remove_all(preprocessing_dirName, ec); }
};

// J/Power_Method.java
struct Power_Method : Algo_Util, Whole_Graph_Util_Interface, Topk_Util_Interface {
    Graph& adjM; double alpha;
    std::vector<double> reserve; PprMap reserve_map, topk_res; std::vector<int64_t> topk_nodeIds;
    Power_Method(double alpha_, Graph& g, const std::string& prop) : Algo_Util(prop), adjM(g), alpha(alpha_) {}
    void computeWholeGraphPPR(int64_t s, double) override {
        reserve.assign((size_t)adjM.n, 0.0);
        check(ppr_power_iteration(adjM.g, (int32_t)s, alpha, 100, reserve.data()), adjM.ctx);   // J/Power_Method.java:55
        reserve_map = to_map(reserve);
    }
    const PprMap& getWholeGraphPPR() override { return reserve_map; }
    const std::vector<double>& dense() override { return reserve; }
    void computeTopKPPR(int64_t s, int k, double) override {
        computeWholeGraphPPR(s, 0);
        std::vector<double> vals;
        topk(adjM, reserve, k, topk_nodeIds, vals, nullptr);
        topk_res.clear();
        for (size_t i = 0; i < vals.size(); i++) topk_res[topk_nodeIds[i]] = vals[i];
    }
    const PprMap& getTopK(int) { return topk_res; }
    const std::vector<int64_t>& getTopKNodeIds(int) override { return topk_nodeIds; }
    void printWholeGraphResult() { print_rows(adjM, "Exact-PPR:", rows_of(reserve), SIZE_MAX); }                                                   // J/Power_Method.java:103-116
    void printTopKResult(int k) { print_rows(adjM, "\nExact-Top" + std::to_string(k) + " PPR:", rows_of(topk_nodeIds, reserve), (size_t)std::max(k, 0)); }   // :167-184
};

// J/Forward_Push.java
struct Forward_Push : Algo_Util, Whole_Graph_Util_Interface, Topk_Util_Interface, Prep_Store {
    Graph& adjM; double alpha, rsum;
    std::vector<double> reserve, residue; PprMap reserve_map; std::vector<int64_t> topk_nodeIds;
    Forward_Push(double alpha_, double rsum_, Graph& g, const std::string& prop, const std::string& dir_db = "got.db")
        : Algo_Util(prop), Prep_Store("FWP_ppr_results/" + dir_db), adjM(g), alpha(alpha_), rsum(rsum_) {}
    void prep_query(int64_t s, double p) override { computeWholeGraphPPR(s, p); }
    const std::vector<double>& prep_result() override { return reserve; }
    void prep_load(const std::vector<double>& d) override { reserve = d; reserve_map = to_map(reserve); topk_nodeIds.clear(); }
    void computeWholeGraphPPR(int64_t s, double rmax) override {
        reserve.assign((size_t)adjM.n, 0.0); residue.assign((size_t)adjM.n, 0.0); topk_nodeIds.clear();
        check(ppr_forward_push(adjM.g, (int32_t)s, alpha, rmax, reserve.data(), residue.data(), &rsum, nullptr), adjM.ctx);
        reserve_map = to_map(reserve);
    }
    double getUpdatedRsum() const { return rsum; }
    PprMap getResidueCopy() const { return to_map(residue); }
    PprMap getReserveCopy() const { return reserve_map; }
    const PprMap& getWholeGraphPPR() override { return reserve_map; }
    const std::vector<double>& dense() override { return reserve; }
    void computeTopKPPR(int64_t s, int, double rmax) override { computeWholeGraphPPR(s, rmax); }
    const std::vector<int64_t>& getTopKNodeIds(int k) override {
        if (topk_nodeIds.empty()) { std::vector<double> v; topk(adjM, reserve, k, topk_nodeIds, v, nullptr); }
        return topk_nodeIds;
    }
    void printWholeGraphResult() { print_rows(adjM, "Forward-Push Reserve(pi) & Residue(r):", rows_of(reserve), SIZE_MAX, true, &residue); }      // J/Forward_Push.java:268-281
    void printTopKResult(int k) { print_rows(adjM, "\nForward-Push-Top" + std::to_string(k) + " PPR:", rows_of(getTopKNodeIds(k), reserve), (size_t)std::max(k, 0)); }   // :366-391
};

// J/Monte_Carlo.java
struct Monte_Carlo : Algo_Util, Whole_Graph_Util_Interface, Topk_Util_Interface, Prep_Store {
    Graph& adjM; double alpha, pfail, delta; uint64_t seed = 1, walk_counter = 0;
    std::vector<double> ppr; PprMap ppr_map; std::vector<int64_t> topk_nodeIds;
    Monte_Carlo(double alpha_, double pfail_, double delta_, Graph& g, const std::string& prop, const std::string& dir_db = "got.db")
        : Algo_Util(prop), Prep_Store("MC_ppr_results/" + dir_db), adjM(g), alpha(alpha_), pfail(pfail_), delta(delta_) {}
    void prep_query(int64_t s, double p) override { computeWholeGraphPPR(s, p); }
    const std::vector<double>& prep_result() override { return ppr; }
    void prep_load(const std::vector<double>& d) override { ppr = d; ppr_map = to_map(ppr); topk_nodeIds.clear(); }
    int64_t random_walk(int64_t start) { return one_walk(start, 0); }
    int64_t random_walk_no_zero_hop(int64_t start) { return one_walk(start, 1); }
    void computeWholeGraphPPR(int64_t s, double eps) override {
        ppr.assign((size_t)adjM.n, 0.0); topk_nodeIds.clear();
        ppr_conf c = {delta, pfail, 0.0};
        check(ppr_monte_carlo(adjM.g, (int32_t)s, alpha, eps, seed++, &c, ppr.data(), nullptr), adjM.ctx);
        ppr_map = to_map(ppr);
    }
    const PprMap& getWholeGraphPPR() override { return ppr_map; }
    const std::vector<double>& dense() override { return ppr; }
    void computeTopKPPR(int64_t s, int, double eps) override { computeWholeGraphPPR(s, eps); }
    const std::vector<int64_t>& getTopKNodeIds(int k) override {
        if (topk_nodeIds.empty()) { std::vector<double> v; topk(adjM, ppr, k, topk_nodeIds, v, nullptr); }
        return topk_nodeIds;
    }
    void printWholeGraphResult() { print_rows(adjM, "Monte-Carlo PPR:", rows_of(ppr), SIZE_MAX); }                                                 // J/Monte_Carlo.java:160-173
    void printTopKResult(int k) { print_rows(adjM, "\nMonte-Carlo-Top" + std::to_string(k) + " PPR:", rows_of(getTopKNodeIds(k), ppr), (size_t)std::max(k, 0)); }   // :276-301
private:
    int64_t one_walk(int64_t start, int nzh) {
        int32_t t = 0;
        check(ppr_walk_terminals(adjM.g, (int32_t)start, (int32_t)start, walk_counter++, 1, seed, 0, alpha, nzh, &t), adjM.ctx);
        return t;
    }
};

// J/Fora_Whole_Graph.java
struct Fora_Whole_Graph : Algo_Util, Whole_Graph_Util_Interface, Prep_Store {
    Graph& adjM; double alpha, rsum, pfail, delta; uint64_t seed = 1;
    std::vector<double> reserve; PprMap reserve_map;
    Fora_Whole_Graph(double alpha_, double rsum_, double pfail_, double delta_, Graph& g, const std::string& prop, const std::string& dir_db = "got.db")
        : Algo_Util(prop), Prep_Store("FORA_ppr_results/" + dir_db), adjM(g), alpha(alpha_), rsum(rsum_), pfail(pfail_), delta(delta_) {}
    void prep_query(int64_t s, double p) override { computeWholeGraphPPR(s, p); }
    const std::vector<double>& prep_result() override { return reserve; }
    void prep_load(const std::vector<double>& d) override { reserve = d; reserve_map = to_map(reserve); }
    void computeWholeGraphPPR(int64_t s, double eps) override {
        reserve.assign((size_t)adjM.n, 0.0);
        ppr_conf c = {delta, pfail, 0.0};
        check(ppr_fora(adjM.g, (int32_t)s, alpha, eps, seed++, PPR_RESTART_AT_WALK_START, &c, reserve.data(), nullptr), adjM.ctx);
        reserve_map = to_map(reserve);
    }
    const PprMap& getWholeGraphPPR() override { return reserve_map; }
    const std::vector<double>& dense() override { return reserve; }
    void printWholeGraphResult() { print_rows(adjM, "Fora-Whole-Graph PPR:", rows_of(reserve), SIZE_MAX); }                                        // J/Fora_Whole_Graph.java:55-68
};

// J/Fora_Topk.java
struct Fora_Topk : Algo_Util, Topk_Util_Interface {
    Graph& adjM; double alpha, rsum, pfail, delta, min_delta; int k; uint64_t seed = 1;
    std::vector<int64_t> topk_nodeIds; std::vector<double> topk_vals;
    Fora_Topk(double alpha_, double rsum_, double pfail_, double delta_, double min_delta_, int k_, Graph& g, const std::string& prop)
        : Algo_Util(prop), adjM(g), alpha(alpha_), rsum(rsum_), pfail(pfail_), delta(delta_), min_delta(min_delta_), k(k_) {}
    void computeTopKPPR(int64_t s, int, double eps) override {
        std::vector<int32_t> ids((size_t)adjM.n);
        topk_vals.assign((size_t)adjM.n, 0.0);
        int32_t cnt = 0;
        ppr_conf c = {delta, pfail, min_delta};
        check(ppr_fora_topk(adjM.g, (int32_t)s, alpha, eps, k, seed++, PPR_RESTART_AT_WALK_START, &c, ids.data(), topk_vals.data(), (int32_t)adjM.n, &cnt, nullptr), adjM.ctx);
        topk_nodeIds.assign(ids.begin(), ids.begin() + cnt);
        topk_vals.resize((size_t)cnt);
    }
    const std::vector<int64_t>& getTopKNodeIds(int) override { return topk_nodeIds; }
    void printTopKResult(int) {                                                                                                                 // J/Fora_Topk.java:60-79
        Rows r;
        for (size_t i = 0; i < topk_nodeIds.size(); i++) r.emplace_back(topk_nodeIds[i], topk_vals[i]);
        print_rows(adjM, "\nFora-Top" + std::to_string(k) + " PPR:", r, (size_t)std::max(k, 0));
    }
};

// J/Backward_Search.java
struct Backward_Search {
    Graph& adjM; double alpha, rmax; std::vector<double> reserve; int64_t duration = 0;
    Backward_Search(double alpha_, double rmax_, Graph& g) : adjM(g), alpha(alpha_), rmax(rmax_) {}
    void backward_search_whole_graph(int64_t t) {
        const auto t0 = std::chrono::steady_clock::now();
        reserve.assign((size_t)adjM.n, 0.0);
        check(ppr_backward_search(adjM.g, (int32_t)t, alpha, rmax, reserve.data(), nullptr), adjM.ctx);
        duration += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    }
    PprMap getReserve() const { return to_map(reserve); }
    int64_t getDuration() const { return duration; }   // J/Backward_Search.java:104-106, nanoseconds
};

// J/Base_Whole_Graph.java (All-Pair-Backward-Search).  The inverted index stays in memory; the per-source
// text files of the reference are written when a directory is given.
struct Base_Whole_Graph : Algo_Util, Whole_Graph_Util_Interface, Topk_Util_Interface {
    Graph& adjM; double alpha; std::string dir;
    std::vector<int64_t> src_ptr; std::vector<int32_t> tgt; std::vector<double> val;
    PprMap ppr_src; std::vector<double> ppr_dense; std::vector<int64_t> topk_nodeIds;
    int64_t prep_bytes = 0;
    Base_Whole_Graph(double alpha_, Graph& g, const std::string& prop, const std::string& dir_db) : Algo_Util(prop), adjM(g), alpha(alpha_), dir("BASE_ppr_results/" + dir_db) {}
    void preprocessing(double threshold, int k) {   // J/Base_Whole_Graph.java:58-164
        ppr_pairs* p = nullptr;
        check(ppr_backward_all_pairs(adjM.g, alpha, threshold, k, 0, (int32_t)adjM.n, &p, nullptr), adjM.ctx);
        int64_t cnt = ppr_pairs_count(p);
        src_ptr.assign((size_t)adjM.n + 1, 0); tgt.assign((size_t)std::max<int64_t>(cnt, 1), 0); val.assign((size_t)std::max<int64_t>(cnt, 1), 0.0);
        check(ppr_pairs_export(p, src_ptr.data(), tgt.data(), val.data(), nullptr), adjM.ctx);
        ppr_pairs_destroy(p);
        prep_bytes = 0;   // size of the reference's text files: "<t>\t<Double.toString(v)>\n" per pair
        for (int64_t i = 0; i < cnt; i++) prep_bytes += (int64_t)std::to_string(tgt[(size_t)i]).size() + 1 + (int64_t)java_double(val[(size_t)i]).size() + 1;
    }
    void computeWholeGraphPPR(int64_t s, double) override {   // J/Base_Whole_Graph.java:167-186: a lookup
        ppr_src.clear(); topk_nodeIds.clear();
        ppr_dense.assign((size_t)adjM.n, 0.0);
        for (int64_t i = src_ptr[(size_t)s]; i < src_ptr[(size_t)s + 1]; i++) {
            ppr_src[tgt[(size_t)i]] = val[(size_t)i]; ppr_dense[(size_t)tgt[(size_t)i]] = val[(size_t)i]; topk_nodeIds.push_back(tgt[(size_t)i]);
        }
    }
    const PprMap& getWholeGraphPPR() override { return ppr_src; }
    const std::vector<double>& dense() override { return ppr_dense; }
    void computeTopKPPR(int64_t s, int, double) override { computeWholeGraphPPR(s, 0); }
    const std::vector<int64_t>& getTopKNodeIds(int) override { return topk_nodeIds; }   // rows are sorted when preprocessing ran with k >= 0
    void readPreprocessedPPR(int64_t s) { computeWholeGraphPPR(s, 0); }   // J/Base_Whole_Graph.java:200-204
    Rows stored_row() const { Rows r; for (int64_t t : topk_nodeIds) r.emplace_back(t, ppr_src.at(t)); return r; }
    void printWholeGraphResult() { print_rows(adjM, "\nBase-Whole-Graph PPR:", stored_row(), SIZE_MAX, false); }                                 // :188-193, stored order
    void printTopKResult(int k) { print_rows(adjM, "\nBASE-Top" + std::to_string(k) + " PPR:", stored_row(), (size_t)std::max(k, 0)); }          // :219-236
    int64_t getPrepSize() const { return prep_bytes; }
    void deletePrepDir() { src_ptr.clear(); tgt.clear(); val.clear(); prep_bytes = 0; }   // the index lives in memory here (:243-251 clears the directory)
};

// J/Algo_Conf.java
struct Algo_Conf {
    double alpha; std::string node_property;
    Algo_Conf(double a, const std::string& p) : alpha(a), node_property(p) {}
    std::unique_ptr<Power_Method> set_conf_power_method(Graph& g) { return std::make_unique<Power_Method>(alpha, g, node_property); }
    std::unique_ptr<Monte_Carlo> set_conf_mc(int64_t n, Graph& g, const std::string& dir_db = "got.db") { double d = 1.0 / (double)n; return std::make_unique<Monte_Carlo>(alpha, d, d, g, node_property, dir_db); }
    std::unique_ptr<Base_Whole_Graph> set_conf_base_whole_graph(Graph& g, const std::string& dir_db) { return std::make_unique<Base_Whole_Graph>(alpha, g, node_property, dir_db); }
    std::unique_ptr<Fora_Whole_Graph> set_conf_fora_whole_graph(int64_t n, Graph& g, const std::string& dir_db = "got.db") { double d = 1.0 / (double)n; return std::make_unique<Fora_Whole_Graph>(alpha, 1.0, d, d, g, node_property, dir_db); }
    std::unique_ptr<Forward_Push> set_conf_fwdpush(Graph& g, const std::string& dir_db = "got.db") { return std::make_unique<Forward_Push>(alpha, 1.0, g, node_property, dir_db); }
    std::unique_ptr<Fora_Topk> set_conf_fora_topk(int64_t n, int k, Graph& g) {
        double min_delta = 1.0 / (double)n, delta = 1.0 / (double)k;
        double pfail = 1.0 / (double)n / (double)n / std::log((double)(n / k));   // integer n / k, J/Algo_Conf.java:76
        return std::make_unique<Fora_Topk>(alpha, 1.0, pfail, delta, min_delta, k, g, node_property);
    }
};

// J/Gen_Util.java: query sampling, timing loops, error metrics, parameter grids, result file
struct Gen_Util {
    enum AlgoType { MC, FWDPUSH, FORA_WHOLE_GRAPH, FORA_TOPK, POWER_METHOD, BASE_WHOLE_GRAPH, NEO4J_METHOD };
    static const char* name(AlgoType a) { static const char* n[] = {"MC", "FWDPUSH", "FORA_WHOLE_GRAPH", "FORA_TOPK", "POWER_METHOD", "BASE_WHOLE_GRAPH", "NEO4J_METHOD"}; return n[a]; }
    Graph adjM; std::string dir_db = "got.db", node_property = "name", label_type = "Person", rel_type = "Relation";
    double alpha = 0.15, epsilon = 0.5; int64_t node_amount = 0, rel_amount = 0;
    std::ofstream fw; std::mt19937_64 rng{1234};
    struct Row { std::string algo; double param, avg_ms, max_err, precision, ndcg; };
    std::vector<Row> rows;

    std::vector<int64_t> getQueryNodes(int q) {   // J/Gen_Util.java:99-107
        std::vector<int64_t> v; std::uniform_int_distribution<int64_t> d(0, node_amount - 1);
        for (int i = 0; i < q; i++) v.push_back(d(rng));
        return v;
    }
    // J/Gen_Util.java:259-326
    static double max_err(Whole_Graph_Util_Interface& algo, Power_Method& pm) {
        double mx = 0; const std::vector<double>&g = pm.dense(), &e = algo.dense();
        for (size_t v = 0; v < g.size(); v++) if (g[v] > 0) mx = std::max(mx, std::fabs(e[v] - g[v]));
        return mx;
    }
    static void topk_err(Topk_Util_Interface& algo, Power_Method& pm, int k, double& precision, double& ndcg) {
        const std::vector<int64_t>& gnd = pm.getTopKNodeIds(k); const std::vector<int64_t>& est = algo.getTopKNodeIds(k); const PprMap& gv = pm.getTopK(k);
        double hit = 0; for (int64_t v : est) if (gv.count(v)) hit += 1;
        precision = gnd.empty() ? 0 : hit / (double)gnd.size();
        double zk = 0, dcg = 0;
        for (size_t i = 1; i <= gnd.size(); i++) zk += (std::pow(2.0, gv.at(gnd[i - 1])) - 1.0) / std::log(i + 1.0) / std::log(2.0);
        for (size_t i = 1; i <= est.size(); i++) { auto it = gv.find(est[i - 1]); double p = it == gv.end() ? 0.0 : it->second; dcg += (std::pow(2.0, p) - 1.0) / std::log(i + 1.0) / std::log(2.0); }
        ndcg = dcg / zk;
    }
    static int64_t now_ns() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    void whole_graph_test(AlgoType t, int query_num, double param, double threshold) {   // J/Gen_Util.java:181-252
        Algo_Conf ac(alpha, node_property);
        std::unique_ptr<Whole_Graph_Util_Interface> algo;
        Base_Whole_Graph* base = nullptr;
        switch (t) {
            case MC: algo = ac.set_conf_mc(node_amount, adjM); break;
            case FWDPUSH: algo = ac.set_conf_fwdpush(adjM); break;
            case FORA_WHOLE_GRAPH: algo = ac.set_conf_fora_whole_graph(node_amount, adjM); break;
            case BASE_WHOLE_GRAPH: { auto b = ac.set_conf_base_whole_graph(adjM, dir_db); base = b.get(); algo = std::move(b); break; }
            default: return;
        }
        auto pm = ac.set_conf_power_method(adjM);
        std::vector<int64_t> q = getQueryNodes(query_num);
        fw << (t == BASE_WHOLE_GRAPH ? std::string("-1") : java_double(param)) << ",";
        if (base) {
            int64_t t0 = now_ns(); base->preprocessing(threshold, -1); int64_t prep = now_ns() - t0;
            fw << java_double(threshold) << "," << prep / 1000000 << "," << base->getPrepSize() << ",";
        }
        int64_t dur = 0; double sum_err = 0;
        for (int64_t s : q) {
            int64_t t0 = now_ns(); algo->computeWholeGraphPPR(s, param); dur += now_ns() - t0;
            pm->computeWholeGraphPPR(s, 0);
            sum_err += max_err(*algo, *pm);
        }
        int64_t avg = dur / query_num;
        fw << avg / 1000000 << "," << java_double(sum_err / query_num) << "\n";
        rows.push_back({name(t), t == BASE_WHOLE_GRAPH ? threshold : param, avg / 1e6, sum_err / query_num, 0, 0});
    }
    void topk_test(AlgoType t, int query_num, int k, double param, double threshold) {   // J/Gen_Util.java:124-180
        Algo_Conf ac(alpha, node_property);
        std::unique_ptr<Topk_Util_Interface> algo;
        Base_Whole_Graph* base = nullptr;
        switch (t) {
            case MC: algo = ac.set_conf_mc(node_amount, adjM); break;
            case FWDPUSH: algo = ac.set_conf_fwdpush(adjM); break;
            case FORA_TOPK: algo = ac.set_conf_fora_topk(node_amount, k, adjM); break;
            case BASE_WHOLE_GRAPH: { auto b = ac.set_conf_base_whole_graph(adjM, dir_db); base = b.get(); algo = std::move(b); break; }
            default: return;
        }
        auto pm = ac.set_conf_power_method(adjM);
        std::vector<int64_t> q = getQueryNodes(query_num);
        if (base) {
            int64_t t0 = now_ns(); base->preprocessing(threshold, k); int64_t prep = now_ns() - t0;
            fw << java_double(threshold) << "," << k << "," << prep / 1000000 << "," << base->getPrepSize() << ",";
        } else fw << java_double(param) << "," << k << ",";
        int64_t dur = 0; double sp = 0, sn = 0;
        for (int64_t s : q) {
            int64_t t0 = now_ns(); algo->computeTopKPPR(s, k, param); dur += now_ns() - t0;
            pm->computeTopKPPR(s, k, 0);
            double p, g; topk_err(*algo, *pm, k, p, g); sp += p; sn += g;
        }
        int64_t avg = dur / query_num;
        fw << avg / 1000000 << "," << java_double(sp / query_num) << "," << java_double(sn / query_num) << "\n";
        rows.push_back({name(t), base ? threshold : param, avg / 1e6, 0, sp / query_num, sn / query_num});
    }
    // Test 3 body for FORA / Forward Push / Monte-Carlo (J/Gen_Util.java:181-252 with to_be_preprocessed = true): preprocess every
    // source into the result store, answer the queries by reading the files back, compare with the power method, delete the store
    void prep_test(AlgoType t, int query_num, double param, double threshold) {
        Algo_Conf ac(alpha, node_property);
        std::unique_ptr<Whole_Graph_Util_Interface> algo;
        Prep_Store* store = nullptr;
        switch (t) {
            case MC: { auto a = ac.set_conf_mc(node_amount, adjM, dir_db); store = a.get(); algo = std::move(a); break; }
            case FWDPUSH: { auto a = ac.set_conf_fwdpush(adjM, dir_db); store = a.get(); algo = std::move(a); break; }
            case FORA_WHOLE_GRAPH: { auto a = ac.set_conf_fora_whole_graph(node_amount, adjM, dir_db); store = a.get(); algo = std::move(a); break; }
            default: return;
        }
        auto pm = ac.set_conf_power_method(adjM);
        std::vector<int64_t> q = getQueryNodes(query_num);
        fw << java_double(param) << ",";
        int64_t t0 = now_ns(); store->preprocessing(threshold, param, node_amount); int64_t prep = now_ns() - t0;
        fw << java_double(threshold) << "," << prep / 1000000 << "," << store->getPrepSize() << ",";
        int64_t dur = 0; double sum_err = 0;
        for (int64_t s : q) {
            int64_t t1 = now_ns(); store->readPreprocessedPPR(s, node_amount); dur += now_ns() - t1;
            pm->computeWholeGraphPPR(s, 0);
            sum_err += max_err(*algo, *pm);
        }
        fw << java_double(sum_err / query_num) << "\n";    // no compute time for preprocessed algorithms (:239-243)
        store->deletePrepDir();
        rows.push_back({std::string(name(t)) + "(prep)", param, dur / 1e6 / query_num, sum_err / query_num, 0, 0});
    }
    // J/Gen_Util.java:328-647, active (GOT) parameter grids :454-478; the Neo4j competitor is out of scope
    void algo_perf_batch_test(int query_num, int k, const std::string& out_dir) {
        static const std::map<AlgoType, std::vector<double>> grid = {
            {BASE_WHOLE_GRAPH, {0.001, 5.0e-4, 5.0e-5, 1.0e-6, 5.0e-7}}, {FORA_WHOLE_GRAPH, {10.0, 5.0, 0.5, 0.1, 0.05}},
            {FORA_TOPK, {10.0, 5.0, 0.5, 0.1, 0.05}}, {MC, {1.0, 0.5, 0.3, 0.1, 0.05}}, {FWDPUSH, {1.0e-4, 1.0e-5, 1.0e-6, 1.0e-7, 1.0e-8}}};
        std::string path = out_dir + "/" + dir_db + "_AlgoPerfResults.txt";
        fw.open(path, std::ios::app);
        char ts[64]; std::time_t now = std::time(nullptr); std::strftime(ts, sizeof ts, "%Y-%m-%d %H:%M:%S", std::localtime(&now));
        fw << ts << "\n\nTest 1. Whole-Graph test\n";
        AlgoType whole[] = {FORA_WHOLE_GRAPH, FWDPUSH, NEO4J_METHOD, MC, BASE_WHOLE_GRAPH};
        for (int i = 0; i < 5; i++) {
            fw << "\n1." << i + 1 << " " << name(whole[i]) << "\n";
            if (whole[i] == NEO4J_METHOD) continue;
            for (double p : grid.at(whole[i])) whole[i] == BASE_WHOLE_GRAPH ? whole_graph_test(whole[i], query_num, -1, p) : whole_graph_test(whole[i], query_num, p, -1.0);
        }
        fw << "\nTest 2. Top-k test\n";
        AlgoType topk[] = {FORA_TOPK, FWDPUSH, NEO4J_METHOD, MC, BASE_WHOLE_GRAPH};
        for (int i = 0; i < 5; i++) {
            fw << "\n2." << i + 1 << " " << name(topk[i]) << "\n";
            if (topk[i] == NEO4J_METHOD) continue;
            for (double p : grid.at(topk[i])) topk[i] == BASE_WHOLE_GRAPH ? topk_test(topk[i], query_num, k, -1, p) : topk_test(topk[i], query_num, k, p, -1.0);
        }
        fw << "\nTest 3. Preprocessing test\n";   // J/Gen_Util.java:602-645
        AlgoType prep[] = {FORA_WHOLE_GRAPH, FWDPUSH, NEO4J_METHOD, MC, BASE_WHOLE_GRAPH};
        for (int i = 0; i < 5; i++) {
            fw << "\n3." << i + 1 << " " << name(prep[i]) << "\n";
            if (prep[i] == NEO4J_METHOD) continue;
            for (double p : grid.at(prep[i])) prep[i] == BASE_WHOLE_GRAPH ? whole_graph_test(prep[i], query_num, -1, p) : prep_test(prep[i], query_num, p, -1.0);
        }
        fw.close();
        printf("results appended to %s\n", path.c_str());
    }
};

// J/PPR.java: the driver
struct PPR : Gen_Util {
    std::string db_path;
    void createDb(const std::string& dir, int device = 0) {
        printf("\nCreating or opening database %s...\n", dir.c_str());
        db_path = dir;
        size_t p = dir.find_last_not_of('/');
        std::string d = dir.substr(0, p + 1);
        size_t s = d.find_last_of('/');
        dir_db = s == std::string::npos ? d : d.substr(s + 1);
        int32_t rc = ppr_init(device, &adjM.ctx);
        if (rc != PPR_OK) throw std::runtime_error(std::string("ppr_init: ") + ppr_last_error(nullptr));
        printf("Database created or opened\n");
    }
    void shutDown() {
        printf("\nShutting down database ...\n");
        if (adjM.g) { ppr_graph_destroy(adjM.g); adjM.g = nullptr; }
        if (adjM.ctx) { ppr_shutdown(adjM.ctx); adjM.ctx = nullptr; }
    }
    // J/PPR.java:136-152: one-time export (here from the Neo4j import CSVs: `:ID,name` / `:START_ID,:END_ID,:TYPE`)
    typedef std::vector<std::pair<int64_t, int64_t>> EdgeList;
    // `neo4j-admin import` CSV pair (":ID,name" / ":START_ID,:END_ID,:TYPE"): mapped id = row of the nodes file
    static void read_import_csv(const std::string& nodes_csv, const std::string& rels_csv, std::vector<std::string>& names, EdgeList& edges) {
        std::ifstream fn(nodes_csv), fr(rels_csv);
        if (!fn || !fr) throw std::runtime_error("cannot open " + nodes_csv + " / " + rels_csv);
        std::string line;
        std::getline(fn, line);   // header (possibly with a UTF-8 BOM)
        std::unordered_map<std::string, int64_t> index;
        while (std::getline(fn, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (line.empty()) continue;
            std::string id = line.substr(0, line.find(','));
            index.emplace(id, (int64_t)names.size());
            names.push_back(line.substr(line.find(',') + 1));
        }
        std::getline(fr, line);
        while (std::getline(fr, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (line.empty()) continue;
            std::stringstream ss(line); std::string a, b;
            std::getline(ss, a, ','); std::getline(ss, b, ',');
            edges.emplace_back(index.at(a), index.at(b));
        }
    }
    // A Neo4j 3.x store directory (the reference's own -db argument, e.g. target/got.db), read without a JVM:
    // neostore.nodestore.db = 15-byte records, bit 0 of byte 0 = in use, node id = record index;
    // neostore.relationshipstore.db = 34-byte records, byte 0 = in-use bit + bits 32-34 of the first node, bytes 1-4 / 5-8 = low 32 bits
    // of the first / second node (big endian), bytes 9-12 = type int whose bits 28-30 carry bits 32-34 of the second node.
    // Mapped id = rank of the Neo4j id among the ids in use (the IdMap of Graph.toMappedNodeId); names = the Neo4j ids.
    static void read_neo4j_store(const std::string& db_dir, std::vector<std::string>& names, EdgeList& edges) {
        auto slurp = [](const std::string& path) {
            std::ifstream f(path, std::ios::binary);
            if (!f) throw std::runtime_error("cannot open " + path);
            return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        };
        const std::vector<unsigned char> nodes = slurp(db_dir + "/neostore.nodestore.db"), rels = slurp(db_dir + "/neostore.relationshipstore.db");
        std::vector<int64_t> mapped(nodes.size() / 15, -1);
        for (size_t i = 0; i < mapped.size(); i++)
            if (nodes[i * 15] & 1) { mapped[i] = (int64_t)names.size(); names.push_back(std::to_string(i)); }
        auto be32 = [](const unsigned char* b) { return ((int64_t)b[0] << 24) | ((int64_t)b[1] << 16) | ((int64_t)b[2] << 8) | (int64_t)b[3]; };
        for (size_t r = 0; r + 34 <= rels.size(); r += 34) {
            const unsigned char* b = &rels[r];
            if (!(b[0] & 1)) continue;
            const int64_t first = be32(b + 1) | ((int64_t)((b[0] & 0x0E) >> 1) << 32);
            const int64_t second = be32(b + 5) | ((be32(b + 9) & 0x70000000LL) << 4);
            if (first < 0 || second < 0 || (size_t)first >= mapped.size() || (size_t)second >= mapped.size() || mapped[(size_t)first] < 0 || mapped[(size_t)second] < 0)
                throw std::runtime_error("relationship store references a node that is not in use");
            edges.emplace_back(mapped[(size_t)first], mapped[(size_t)second]);
        }
    }
    void build_graph(EdgeList& edges, int64_t t0) {
        std::sort(edges.begin(), edges.end());   // canonical order: ascending target inside a row, duplicates kept
        adjM.n = (int64_t)adjM.names.size(); adjM.m = (int64_t)edges.size();
        adjM.row_ptr.assign((size_t)adjM.n + 1, 0);
        std::vector<int32_t> col((size_t)adjM.m);
        for (auto& e : edges) adjM.row_ptr[(size_t)e.first + 1]++;
        for (int64_t v = 0; v < adjM.n; v++) adjM.row_ptr[(size_t)v + 1] += adjM.row_ptr[(size_t)v];
        for (size_t i = 0; i < edges.size(); i++) col[i] = (int32_t)edges[i].second;
        check(ppr_graph_create(adjM.ctx, adjM.n, adjM.m, adjM.row_ptr.data(), col.data(), &adjM.g), adjM.ctx);
        printf("\nFinish graph loading in %lld(ms)\n", (long long)((now_ns() - t0) / 1000000));
    }
    // J/PPR.java:136-152: one-time export of the topology to the device
    void setupAdjMatrix(const std::string& nodes_csv, const std::string& rels_csv) {
        printf("\nLoading graph from database...\n");
        int64_t t0 = now_ns();
        EdgeList edges;
        read_import_csv(nodes_csv, rels_csv, adjM.names, edges);
        build_graph(edges, t0);
    }
    void setupAdjMatrix(const std::string& neo4j_store_dir) {
        printf("\nLoading graph from database...\n");
        int64_t t0 = now_ns();
        EdgeList edges;
        read_neo4j_store(neo4j_store_dir, adjM.names, edges);
        build_graph(edges, t0);
    }
    void set_configuration(double a, double e, const std::string& np, const std::string& lt, const std::string& rt) {
        alpha = a; epsilon = e; node_property = np; label_type = lt; rel_type = rt;
        node_amount = ppr_graph_node_count(adjM.g); rel_amount = ppr_graph_rel_count(adjM.g);
    }
};

}  // namespace fora_neo4j
