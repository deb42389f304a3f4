// rmat.cpp -- seeded R-MAT generator + canonical CSR build (host only; synthetic inputs for bench/tests).
// Not part of the query path.  Deterministic for a given argument list, independent of thread count:
// every edge draws from its own counter-based stream keyed by (seed, edge index, attempt).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/pprgen.h"

// status codes of this small library (same values as pprb200.h)
#define PPR_OK 0
#define PPR_E_INVALID (-1)
#define PPR_E_NOMEM (-3)

typedef uint64_t u64;
typedef int64_t i64;
typedef int32_t i32;

static inline u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Stream {
    u64 s;
    explicit Stream(u64 seed) : s(seed) {}
    inline u64 next() { s += 0x9E3779B97F4A7C15ull; return mix64(s); }
};

template <typename F>
static void parallel_for(i64 n, F f) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::max(1u, std::min(hw ? hw : 1u, 64u));
    if (n < (1 << 16)) nt = 1;
    std::vector<std::thread> th;
    i64 per = (n + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        i64 lo = std::min<i64>(n, t * per), hi = std::min<i64>(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=]() { f(lo, hi); });
    }
    for (auto& x : th) x.join();
}

extern "C" int32_t ppr_rmat_generate(int64_t n, int64_t m, double a, double b, double c, uint64_t seed, int32_t permute,
                                     int32_t symmetric, int64_t* row_ptr, int32_t* col_idx) {
    if (n < 1 || n > 0x7fffffffll || m < 0 || !row_ptr || (m > 0 && !col_idx)) return PPR_E_INVALID;
    if (!(a > 0 && b >= 0 && c >= 0 && a + b + c < 1.0)) return PPR_E_INVALID;
    if (symmetric && (m & 1)) return PPR_E_INVALID;
    int scale = 0;
    while (((i64)1 << scale) < n) scale++;
    const i64 gen = symmetric ? m / 2 : m;
    std::vector<i32> src, dst, perm;
    try {
        src.resize((size_t)m); dst.resize((size_t)m);
        if (permute) perm.resize((size_t)n);
    } catch (...) { return PPR_E_NOMEM; }
    const u64 ta = (u64)(a * 9007199254740992.0), tb = (u64)((a + b) * 9007199254740992.0),
              tc = (u64)((a + b + c) * 9007199254740992.0);
    parallel_for(gen, [&](i64 lo, i64 hi) {
        for (i64 e = lo; e < hi; e++) {
            for (u64 attempt = 0;; attempt++) {
                Stream st(mix64(seed ^ 0xA5A5A5A5DEADBEEFull) ^ mix64((u64)e * 2 + 1) ^ mix64(attempt * 0x632BE59BD9B4E019ull));
                i64 s = 0, d = 0;
                for (int l = 0; l < scale; l++) {
                    u64 u = st.next() >> 11;
                    int sb, db;
                    if (u < ta) { sb = 0; db = 0; }
                    else if (u < tb) { sb = 0; db = 1; }
                    else if (u < tc) { sb = 1; db = 0; }
                    else { sb = 1; db = 1; }
                    s = (s << 1) | sb; d = (d << 1) | db;
                }
                if (s < n && d < n) { src[(size_t)e] = (i32)s; dst[(size_t)e] = (i32)d; break; }
            }
        }
    });
    if (permute) {
        for (i64 v = 0; v < n; v++) perm[(size_t)v] = (i32)v;
        Stream st(mix64(seed + 0x1234567ull));
        for (i64 v = n - 1; v > 0; v--) {
            u64 j = (u64)(((unsigned __int128)st.next() * (u64)(v + 1)) >> 64);
            std::swap(perm[(size_t)v], perm[(size_t)j]);
        }
        parallel_for(gen, [&](i64 lo, i64 hi) {
            for (i64 e = lo; e < hi; e++) { src[(size_t)e] = perm[(size_t)src[(size_t)e]]; dst[(size_t)e] = perm[(size_t)dst[(size_t)e]]; }
        });
    }
    if (symmetric) {
        parallel_for(gen, [&](i64 lo, i64 hi) {
            for (i64 e = lo; e < hi; e++) { src[(size_t)(gen + e)] = dst[(size_t)e]; dst[(size_t)(gen + e)] = src[(size_t)e]; }
        });
    }
    // counting sort by source, then ascending targets inside each row
    std::memset(row_ptr, 0, ((size_t)n + 1) * sizeof(int64_t));
    for (i64 e = 0; e < m; e++) row_ptr[(size_t)src[(size_t)e] + 1]++;
    for (i64 v = 0; v < n; v++) row_ptr[v + 1] += row_ptr[v];
    {
        std::vector<i64> pos(row_ptr, row_ptr + n);
        for (i64 e = 0; e < m; e++) col_idx[(size_t)pos[(size_t)src[(size_t)e]]++] = dst[(size_t)e];
    }
    parallel_for(n, [&](i64 lo, i64 hi) {
        for (i64 v = lo; v < hi; v++) std::sort(col_idx + row_ptr[v], col_idx + row_ptr[v + 1]);
    });
    return PPR_OK;
}
