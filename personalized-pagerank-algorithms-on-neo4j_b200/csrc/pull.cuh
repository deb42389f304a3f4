// pull.cuh -- pull-direction iterations of the forward push (included by push.cuh; sm_100a).
//
// Same synchronous schedule, same bits as the push-direction iteration of push.cuh (J/Forward_Push.java:85-141), for the
// fattest iterations of a query: when the frontier owns a large share of all edges, scattering E_p atomics (181 G RED/s
// is the L2's ceiling) and then sweeping r[] / reserve[] / staging with scattered writes (60-75 us at R-MAT scale 22) costs
// more than reading EVERY in-edge once:
//   iteration i, pull form:   r[u] += sum over in-neighbours v of q_i[v]      (q_i[v] = per-edge amount of v if v is in
//                             frontier i, else 0; dense array, written by the previous pass)
//   fused with phase A of iteration i+1 for the same u: threshold test on the new r[u], reserve update, q_{i+1}[u].
// No atomics (a vertex is owned by one lane; integer adds commute, so the sums equal the scatter's), no staging, no
// sweep, every write coalesced.  Cost per pass ~ m gathers of q (8 bytes, L2-resident) + one stream over the CSC,
// independent of the frontier -- it wins when E_p is above roughly a fifth of m (direction-optimising traversal).
//   * q of the PULL_SLOTS highest out-degree vertices is staged in shared memory per CTA (pull_idx encodes those sources
//     as HOT_BIT | slot), which removes the hub share of the gathers from the L2;
//   * rows longer than PULL_LONG in-edges are cut into chunks summed by separate warps (red.global.add into r[u]) and
//     their threshold test runs after a grid barrier (pull_long); all other rows are owned by a lane of the warp that
//     holds their 32-vertex group: medium rows (>= 32 in-edges) are summed by the whole warp, short rows by their lane;
//   * groups are dealt to warps by a static split balanced on in-edges (pull_task, built at export).
// Entry: a dense sweep that writes q instead of staging (pull_sweep_q); exit: q -> staging (pull_unstage).
#pragma once

#ifndef PULL_SLOTS
#define PULL_SLOTS 8192      // hub sources whose q is staged in shared memory (64 KB per CTA)
#endif
#define PULL_MASK ((unsigned)PULL_SLOTS - 1u)
#define PULL_LONG 2048       // rows with more in-edges are summed chunk-wise by several warps
#define PULL_LCHUNK 2048

__device__ __forceinline__ u64 ldcg_u64(const u64* p) { return __ldcg(reinterpret_cast<const ull*>(p)); }

// value of one in-edge: q of its source (hub sources from shared memory)
__device__ __forceinline__ u64 pull_val(u32 idx, const u64* __restrict__ qcur, const u64* sm_hq) {
    return (idx & HOT_BIT) ? sm_hq[idx & PULL_MASK] : ldcg_u64(&qcur[idx]);
}

// phase A of the next iteration for one vertex whose new residue is x (J/Forward_Push.java:91-97, :101-113 for d == 0)
__device__ __forceinline__ void pull_capture(const PushArgs& A, u64 rmax_fx, i64 u, u64 x, u32 dout, bool dirty, u64* __restrict__ qnext, PhaseATotals& tot) {
    u64 q = 0;
    if (x != 0 && x >= thr_fwd(rmax_fx, dout)) {
        const u64 a = mulhi64(x, A.alpha_q64);
        const u64 y = x - a;
        u64 left;
        if (dout == 0) { left = 0; tot.dang += y; }
        else { q = y / (u64)dout; left = y - q * (u64)dout; }
        A.r[u] = left;
        A.reserve[u] += a;
        tot.resv += a; tot.edges += dout; tot.pushed += 1;
    } else if (dirty) {
        A.r[u] = x;
    }
    qnext[u] = q;
}

// Entry into pull mode: dense sweep = phase A of iteration `par` for every vertex, amounts written to q (0 for vertices
// that do not push) instead of the staging lists.
__device__ __noinline__ void pull_sweep_q(const PushArgs& A, int par, u64 rmax_fx, u64* __restrict__ qout, u32* sm_cnt) {
    PushCtl* ctl = A.ctl;
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    if (gtid == 0) phase_a_recycle(ctl, par);
    PhaseATotals tot = {0, 0, 0, 0};
    const i64 n = A.n;
    const u32 lane = lane_id();
    const i64 gwarp = gtid >> 5, nwarps = gsize >> 5;
    for (i64 tile = gwarp * 128; tile < n; tile += nwarps * 128) {   // 128 consecutive vertices per warp, four loads in flight per lane
        u64 x[4]; u32 d[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const i64 v = tile + k * 32 + lane; x[k] = v < n ? ldcg_u64(&A.r[v]) : 0; d[k] = v < n ? __ldg(&A.out_deg[v]) : 0; }
#pragma unroll
        for (int k = 0; k < 4; k++) { const i64 v = tile + k * 32 + lane; if (v < n) pull_capture(A, rmax_fx, v, x[k], d[k], false, qout, tot); }
    }
    phase_a_publish(ctl, par, tot, sm_cnt);
}

// One pull pass, part 1: iteration `par` (amounts in qcur, its totals already published) + phase A of iteration par^1 for
// every vertex that is not a long row.  sm_hq: PULL_SLOTS u64 of shared memory, zeroed again on exit.
__device__ __noinline__ void pull_main(const PushArgs& A, int par, u64 rmax_fx, const u64* __restrict__ qcur, u64* __restrict__ qnext, u64* sm_hq, u32* sm_cnt, u64 dang_cur) {
    PushCtl* ctl = A.ctl;
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    const i64 gwarp = gtid >> 5, nwarps = gsize >> 5;
    const u32 lane = lane_id();
    const i64 n = A.n;
    // dang_cur: dangling mass of iteration `par`, goes to the source in its phase B (read by the caller before any barrier)
    for (u32 sl = threadIdx.x; sl < A.n_pull_hot; sl += blockDim.x) sm_hq[sl] = ldcg_u64(&qcur[__ldg(&A.pull_hot_id[sl])]);
    __syncthreads();
    PhaseATotals tot = {0, 0, 0, 0};

    // (a) long rows: one warp per chunk, partial sums added to r[u]; the owner's capture waits for the grid barrier (pull_long)
    for (i64 c = gwarp; c < (i64)A.n_long_chunk; c += nwarps) {
        const uint2 desc = A.long_chunk[c];
        const u64 pk = __ldg(&A.in_vtx[desc.x]);
        const u32 off = desc.y * PULL_LCHUNK, len = min((u32)PULL_LCHUNK, vtx_deg(pk) - off);
        const u32* row = A.pull_idx + vtx_begin(pk) + off;
        u64 s = 0;
        for (u32 eb = 0; eb < len; eb += 256) {
            u32 idx[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { const u32 e = eb + 32 * k + lane; idx[k] = e < len ? __ldg(&row[e]) : 0xffffffffu; }
#pragma unroll
            for (int k = 0; k < 8; k++) if (idx[k] != 0xffffffffu) s += pull_val(idx[k], qcur, sm_hq);
        }
        s = warp_sum_u64(s);
        if (lane == 0 && s) red_add_u64(&A.r[desc.x], s);
    }

    // (b) all other rows, by 32-vertex groups; tasks = runs of groups with equal in-edge counts, a contiguous run per warp
    const i64 t_lo = (i64)A.n_pull_task * gwarp / nwarps, t_hi = (i64)A.n_pull_task * (gwarp + 1) / nwarps;
    const i64 g_lo = t_lo < t_hi ? (i64)__ldg(&A.pull_task[t_lo]) : 0, g_hi = t_lo < t_hi ? (i64)__ldg(&A.pull_task[t_hi]) : 0;
    for (i64 g = g_lo; g < g_hi; g++) {
        const i64 u = g * 32 + lane;
        const bool valid = u < n;
        const u64 pk = valid ? __ldg(&A.in_vtx[u]) : 0;
        const u32 din = vtx_deg(pk);
        const u64 x = valid ? ldcg_u64(&A.r[u]) : 0;
        const u32 dout = valid ? __ldg(&A.out_deg[u]) : 0;
        const bool is_long = din > PULL_LONG;
        u64 sum = 0;
        // medium rows: the whole warp sums one row at a time (coalesced 128-byte index loads)
        u32 mm = __ballot_sync(0xffffffffu, din >= 32 && !is_long);
        while (mm) {
            const int l = __ffs(mm) - 1;
            mm &= mm - 1;
            const u64 rpk = __shfl_sync(0xffffffffu, pk, l);
            const u32 d = vtx_deg(rpk);
            const u32* row = A.pull_idx + vtx_begin(rpk);
            u64 s = 0;
            for (u32 eb = 0; eb < d; eb += 128) {
                u32 idx[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { const u32 e = eb + 32 * k + lane; idx[k] = e < d ? __ldg(&row[e]) : 0xffffffffu; }
#pragma unroll
                for (int k = 0; k < 4; k++) if (idx[k] != 0xffffffffu) s += pull_val(idx[k], qcur, sm_hq);
            }
            s = warp_sum_u64(s);
            if ((int)lane == l) sum = s;
        }
        // short rows: each lane walks its own row, four edges in flight
        if (din < 32) {
            const u32* row = A.pull_idx + vtx_begin(pk);
            for (u32 eb = 0; eb < din; eb += 4) {
                u32 idx[4];
#pragma unroll
                for (int k = 0; k < 4; k++) idx[k] = eb + k < din ? __ldg(&row[eb + k]) : 0xffffffffu;
#pragma unroll
                for (int k = 0; k < 4; k++) if (idx[k] != 0xffffffffu) sum += pull_val(idx[k], qcur, sm_hq);
            }
        }
        if (valid && !is_long) {
            u64 xn = x + sum;
            bool dirty = sum != 0;
            if (u == (i64)A.src && dang_cur) { xn += dang_cur; dirty = true; }
            pull_capture(A, rmax_fx, u, xn, dout, dirty, qnext, tot);
        }
    }
    phase_a_publish(ctl, par ^ 1, tot, sm_cnt);   // (ends with a CTA barrier: sm_hq is no longer read)
    for (u32 sl = threadIdx.x; sl < PULL_SLOTS; sl += blockDim.x) sm_hq[sl] = 0;   // the push-direction phase B expects zeroed hub accumulators
    __syncthreads();
}

// One pull pass, part 2 (after a grid barrier): phase A of iteration par^1 for the long rows, then the iteration's bookkeeping.
__device__ __noinline__ void pull_long(const PushArgs& A, int par, u64 rmax_fx, u64* __restrict__ qnext, u32* sm_cnt, u64 dang_cur, u32 fcnt, u64 fedges) {
    PushCtl* ctl = A.ctl;
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    PhaseATotals tot = {0, 0, 0, 0};
    for (i64 i = gtid; i < (i64)A.n_long; i += gsize) {
        const i64 u = (i64)__ldg(&A.long_v[i]);
        u64 x = ldcg_u64(&A.r[u]);
        bool dirty = false;
        if (u == (i64)A.src && dang_cur) { x += dang_cur; dirty = true; }
        pull_capture(A, rmax_fx, u, x, __ldg(&A.out_deg[u]), dirty, qnext, tot);
    }
    phase_a_publish(ctl, par ^ 1, tot, sm_cnt);
    if (gtid == 0) {   // every CTA read this iteration's counters (dang_cur, fcnt, fedges) before the grid barrier that precedes this function
        ctl->iters += 1; ctl->pushes += fcnt; ctl->edges += fedges;
        ctl->cnt[par] = 0; ctl->stage[par] = 0; ctl->dang[par] = 0; ctl->fcnt[par] = 0; ctl->fedges[par] = 0; ctl->nedge[par] = 0;
        ctl->cnt[par ^ 1] = 0; ctl->stage[par ^ 1] = 0; ctl->nedge[par ^ 1] = 0;
    }
}

// Exit from pull mode: the frontier of iteration `par` exists as dense amounts; build the staging lists of the push-direction phase B.
__device__ __noinline__ void pull_unstage(const PushArgs& A, int par, const u64* __restrict__ qcur) {
    PushCtl* ctl = A.ctl;
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    const u32 lane = lane_id();
    const i64 n = A.n, gwarp = gtid >> 5, nwarps = gsize >> 5;
    for (i64 tile = gwarp * 128; tile < n; tile += nwarps * 128) {
        u64 q[4], pk[4];
        u32 cls[4], my_nch = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { const i64 v = tile + k * 32 + lane; q[k] = v < n ? ldcg_u64(&qcur[v]) : 0; }
#pragma unroll
        for (int k = 0; k < 4; k++) { const i64 v = tile + k * 32 + lane; pk[k] = q[k] ? __ldg(&A.out_vtx[v]) : 0; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const u32 d = vtx_deg(pk[k]);
            cls[k] = q[k] == 0 ? 0u : (d <= PUSH_SMALL_MAX ? 1u : 2u);
            if (cls[k] == 2) my_nch += (d + PUSH_CHUNK - 1) / PUSH_CHUNK;
        }
        u32 ms[4], small_total = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { ms[k] = __ballot_sync(0xffffffffu, cls[k] == 1); small_total += __popc(ms[k]); }
        u32 chunk_incl = 0, chunk_total = 0;
        if (__any_sync(0xffffffffu, my_nch != 0)) {
            chunk_incl = my_nch;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, chunk_incl, o); if (lane >= o) chunk_incl += t; }
            chunk_total = __shfl_sync(0xffffffffu, chunk_incl, 31);
        }
        if (small_total | chunk_total) {
            u64 base = 0;
            if (lane == 0) base = atomicAdd((ull*)&ctl->stage[par], ((ull)chunk_total << 32) | (ull)small_total);
            base = __shfl_sync(0xffffffffu, base, 0);
            u32 off = (u32)base;
            u32 cb = (u32)(base >> 32) + chunk_incl - my_nch;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (cls[k] == 1) { const u32 slot = off + __popc(ms[k] & lanemask_lt()); A.s_pack[slot] = pk[k]; A.s_q[slot] = q[k]; }
                off += __popc(ms[k]);
                if (cls[k] == 2) {
                    const u32 nch = (vtx_deg(pk[k]) + PUSH_CHUNK - 1) / PUSH_CHUNK;
                    A.l_pack[cb] = pk[k]; A.l_q[cb] = q[k];
                    for (u32 c = 0; c < nch; c++) A.chunks[cb + c] = make_uint2(cb, c);
                    cb += nch;
                }
            }
        }
    }
}
