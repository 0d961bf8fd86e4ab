// ugs_preproc.cu -- segmentation of the PyG batch and per-graph preprocessing on sm_100a.
//
// Replaces, for the whole batch at once:
//   slice_and_renumber_edge_index_with_map   ugs_sampler_batch_extension.cpp:41-75   (A.1)
//   build_csr                                preproc.cpp:32-86                       (A.2)
//   compute_1dd_ordering                     preproc.cpp:97-166                      (A.3)
//   compute_suffix_and_buckets               preproc.cpp:176-256                     (A.4)
// Two code paths with identical results:
//   * bitmask path  (n <= 64, <= 384 columns, k <= 16): one WARP per graph, everything in shared memory;
//   * general path  (anything else, k <= 32): one BLOCK per graph, tables in the global workspace (L2).
#include <stdio.h>
#include <mutex>

#include "ugs_common.cuh"

// =====================================================================================================
// A.1  column -> graph segmentation
// =====================================================================================================
__global__ void k_seg_keys(const i64* __restrict__ row0, i64 E, const i64* __restrict__ ptr, i64 G,
                           const i64* __restrict__ perm, i32* __restrict__ colkey) {
    i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= E) return;
    i64 col = perm ? perm[j] : j;
    i64 u = row0[col];
    i32 key = (i32)G;
    if (G > 0 && u >= ptr[0] && u < ptr[G]) {
        i64 lo = 0, hi = G - 1;  // last g with ptr[g] <= u
        while (lo < hi) {
            i64 mid = (lo + hi + 1) >> 1;
            if (ptr[mid] <= u) lo = mid; else hi = mid - 1;
        }
        key = (i32)lo;
    }
    colkey[j] = key;
}

// cstart[g] = first column position whose key is >= g. One thread per boundary j in [0, E].
__global__ void k_seg_bounds(const i32* __restrict__ colkey, i64 E, i64 G, i64* __restrict__ cstart,
                             unsigned long long* counters) {
    i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > E) return;
    i64 prev = (j == 0) ? -1 : (i64)colkey[j - 1];
    i64 cur = (j == E) ? G + 1 : (i64)colkey[j];
    if (cur < prev) { atomicOr(&counters[UGS_CTR_UNSORTED], 1ull); return; }
    for (i64 g = prev + 1; g <= cur; ++g) cstart[g] = j;
}

// Fast path for graph-contiguous batches (every PyG batch): cstart[g] = first column whose source node is
// >= ptr[g], by one binary search per GRAPH over the columns. The result is only right if the columns are grouped
// by graph; that is verified for free later (k_classify: monotone bounds covering [0,E); preproc kernels: every
// column of a graph's range has its source inside the graph) and the caller falls back to the per-column path
// (k_seg_keys + stable sort) when a check trips.
__global__ void k_seg_fast(const i64* __restrict__ row0, i64 E, const i64* __restrict__ ptr, i64 G,
                           i64* __restrict__ cstart) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > G + 1) return;
    if (g == G + 1) { cstart[g] = E; return; }
    const i64 target = ptr[g];
    i64 lo = 0, hi = E;  // first j with row0[j] >= target
    while (lo < hi) {
        const i64 mid = (lo + hi) >> 1;
        if (row0[mid] >= target) hi = mid; else lo = mid + 1;
    }
    cstart[g] = lo;
}

// SEG: the fast segmentation (k_seg_fast) done by the classifying thread itself -- the column range of graph g by two
// interleaved binary searches over the columns -- so that a batch costs one launch instead of two here.
template <bool SEG>
__global__ void k_classify(const i64* __restrict__ ptr, i64 G, int k, int skip_lt_k, const i64* __restrict__ fast_row0,
                           i64 E, WS w) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    i64 n = ptr[g + 1] - ptr[g];
    i64 c_lo, c_hi;
    if (SEG) {
        const i64 ta = ptr[g], tb = ptr[g + 1];
        i64 lo_a = 0, hi_a = E, lo_b = 0, hi_b = E;  // first column whose source is >= ta / >= tb
        while (lo_a < hi_a || lo_b < hi_b) {
            if (lo_a < hi_a) { const i64 mid = (lo_a + hi_a) >> 1; if (fast_row0[mid] >= ta) hi_a = mid; else lo_a = mid + 1; }
            if (lo_b < hi_b) { const i64 mid = (lo_b + hi_b) >> 1; if (fast_row0[mid] >= tb) hi_b = mid; else lo_b = mid + 1; }
        }
        c_lo = lo_a; c_hi = lo_b;
        w.cstart[g] = c_lo;
        if (g == G - 1) { w.cstart[G] = c_hi; w.cstart[G + 1] = E; }
    } else {
        c_lo = w.cstart[g]; c_hi = w.cstart[g + 1];
    }
    i64 nc = c_hi - c_lo;
    ugs_ginfo gi;
    gi.Z = 0.0; gi.nnz = 0; gi.level = -1; gi.n_eligible = 0; gi.last_viable = -1; gi.mult = 0;
    if (n < 0) atomicOr(&w.counters[UGS_CTR_UNSORTED], 2ull);  // ptr not non-decreasing
    if (nc < 0 || (g == 0 && c_lo != 0) || (g == G - 1 && c_hi != (SEG ? E : w.cstart[G + 1])))
        atomicOr(&w.counters[UGS_CTR_UNSORTED], 4ull);  // bounds do not tile [0,E): not graph-contiguous
    if (n <= 0 || (skip_lt_k && n < k)) {  // ugs_sampler_batch_extension.cpp:132 (batch driver only)
        gi.cls = 0;
        atomicAdd(&w.counters[UGS_CTR_N_SKIPPED], 1ull);
        // No preprocessing kernel looks at a skipped graph's column range, so the fast segmentation is verified here:
        // a column of ANOTHER graph inside this range (batch not graph-contiguous) would otherwise be lost silently.
        if (fast_row0) {
            const i64 lo_ = ptr[g], hi_ = ptr[g + 1];
            for (i64 c = c_lo; c < c_hi; ++c) {
                const i64 u = fast_row0[c];
                if (u < lo_ || u >= hi_) { atomicOr(&w.counters[UGS_CTR_UNSORTED], 8ull); break; }
            }
        }
    } else if (n <= UGS_SMALL_MAX_N && nc <= UGS_SMALL_MAX_COLS && k <= 16) {
        gi.cls = 1;
    } else {
        gi.cls = 2;
        unsigned long long slot = atomicAdd(&w.counters[UGS_CTR_N_GENERAL], 1ull);
        w.glist[slot] = (i32)g;
        atomicMax(&w.counters[UGS_CTR_MAX_N_GENERAL], (unsigned long long)n);
        if (n > UGS_LARGE_N) w.llist[atomicAdd(&w.counters[UGS_CTR_N_LARGE], 1ull)] = (i32)g;
        if (!ugs_is_medium(n, nc)) atomicAdd(&w.counters[UGS_CTR_N_NOT_MEDIUM], 1ull);
    }
    w.ginfo[g] = gi;
}

// =====================================================================================================
// Bitmask path: one warp per graph
// =====================================================================================================
struct __align__(16) SmallPre {
    u64 adj[64];
    u64 suf[64];
    double b[64];
    u32 deg[64];
    unsigned short rowptr[66];
    unsigned short sd[64];
    unsigned char cu[UGS_SMALL_MAX_COLS];
    unsigned char cv[UGS_SMALL_MAX_COLS];
    unsigned char idx[2 * UGS_SMALL_MAX_COLS];
    unsigned char order[64];
    unsigned char index_of[64];
};

__device__ __forceinline__ int popc_t(u32 m) { return __popc(m); }
__device__ __forceinline__ int popc_t(u64 m) { return __popcll(m); }
__device__ __forceinline__ int ffs_t(u32 m) { return __ffs((int)m); }
__device__ __forceinline__ int ffs_t(u64 m) { return __ffsll((long long)m); }

// MaskT = u32: n <= 32, vertex / position `lane` only; MaskT = u64: vertices / positions lane and lane + 32.
template <int KT, typename MaskT>
__device__ __forceinline__ void preproc_small_tail(SmallPre& s, int n, int k, int nnz, i64 nb, i64 g, i64 csr_base,
                                                   const WS& w) {
    constexpr bool NARROW = sizeof(MaskT) == 4;
    constexpr int H = NARROW ? 1 : 2;
    const int lane = threadIdx.x & 31;
    const u32 lt_mask = (1u << lane) - 1u;
    // ---- A.3 order = stable sort by (degree, id): a counting sort over the DISTINCT degree values (a molecule has four
    //      or five of them): per value two ballots give every vertex of that degree its rank, ids ascending because
    //      lane order is id order (vertices lane and lane + 32).
    {
        const u32 none = 0xFFFFFFFFu;
        const u32 d0 = (lane < n) ? s.deg[lane] : none;
        const u32 d1 = (!NARROW && lane + 32 < n) ? s.deg[lane + 32] : none;
        int base = 0;
        u32 cur = 0;
        bool first = true;
        while (base < n) {
            const u32 c0 = (d0 != none && (first || d0 > cur)) ? d0 : none;
            const u32 c1 = (!NARROW && d1 != none && (first || d1 > cur)) ? d1 : none;
            const u32 mn = __reduce_min_sync(FULL_MASK, NARROW ? c0 : min(c0, c1));
            const u32 b0 = __ballot_sync(FULL_MASK, d0 == mn);
            const u32 b1 = NARROW ? 0u : __ballot_sync(FULL_MASK, d1 == mn);
            if (d0 == mn) {
                const int rank = base + __popc(b0 & lt_mask);
                s.order[rank] = (unsigned char)lane;
                s.index_of[lane] = (unsigned char)rank;
            }
            if (!NARROW && d1 == mn) {
                const int rank = base + __popc(b0) + __popc(b1 & lt_mask);
                s.order[rank] = (unsigned char)(lane + 32);
                s.index_of[lane + 32] = (unsigned char)rank;
            }
            base += __popc(b0) + __popc(b1);
            cur = mn;
            first = false;
        }
    }
    __syncwarp();
    // ---- suffix masks: suf[vi] = { order[vj] : vj >= vi } ----------------------------------------
    if constexpr (NARROW) {
        u32 x = (lane < n) ? (1u << s.order[lane]) : 0u;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u32 y = __shfl_down_sync(FULL_MASK, x, off);
            if (lane + off < 32) x |= y;
        }
        s.suf[lane] = (u64)x;
    } else {
        u64 xh = (lane + 32 < n) ? (1ull << s.order[lane + 32]) : 0ull;
        u64 xl = (lane < n) ? (1ull << s.order[lane]) : 0ull;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            u64 yh = __shfl_down_sync(FULL_MASK, xh, off);
            u64 yl = __shfl_down_sync(FULL_MASK, xl, off);
            if (lane + off < 32) { xh |= yh; xl |= yl; }
        }
        xl |= __shfl_sync(FULL_MASK, xh, 0);
        s.suf[lane] = xl;
        s.suf[lane + 32] = xh;
    }
    __syncwarp();
    // ---- A.4 suffix degree, reachability (bounded BFS as a bitmask closure), weight; the same walk over row(v) also
    //      classifies the row's multiplicities (every neighbour once / twice / both patterns, no self loop) ----------
    int n0 = 0, n1 = 0, lastv = -1;
    bool self_or_bad = false;
    int mu_min = 99, mu_max = 0;
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const int vi = lane + 32 * h;
        double b = 0.0;
        int sd = 0;
        if (vi < n) {
            const int v = s.order[vi];
            const int p0 = s.rowptr[v], p1 = s.rowptr[v + 1];
            MaskT ones = 0, twos = 0, fours = 0;
            for (int p = p0; p < p1; ++p) {
                const int x = s.idx[p];
                sd += (s.index_of[x] >= vi) ? 1 : 0;
                if (x == v) self_or_bad = true;
                const MaskT bit = (MaskT)1 << x;
                const MaskT cy0 = ones & bit; ones ^= bit;
                const MaskT cy1 = twos & cy0; twos ^= cy0;
                fours |= cy1;
            }
            const MaskT A = (MaskT)s.adj[v];
            if (A) {
                int mu = 0;
                if (fours) mu = 0;
                else if (ones == A && twos == 0) mu = 1;
                else if (twos == A && ones == 0) mu = 2;
                else if (ones == A && twos == A) mu = 3;
                if (mu == 0) self_or_bad = true;
                else { mu_min = min(mu_min, mu); mu_max = max(mu_max, mu); }
            }
            const MaskT sufm = (MaskT)s.suf[vi];
            MaskT vis = (MaskT)1 << v, expd = 0;
            bool reach = false;
#pragma unroll
            for (int it = 0; it < (KT ? KT : 16); ++it) {
                if (it >= k) break;
                if (popc_t(vis) >= k) { reach = true; break; }
                const MaskT fr = vis & ~expd;
                if (!fr) break;
                const int u = ffs_t(fr) - 1;
                expd |= (MaskT)1 << u;
                vis |= (MaskT)s.adj[u] & sufm;
            }
            if (reach) {
                const double d = (double)(sd > 1 ? sd : 1);
                b = 1.0;
#pragma unroll
                for (int t = 0; t < (KT ? KT - 1 : 15); ++t) {  // preproc.cpp:228-233: k - 1 sequential multiplies
                    if (t >= k - 1) break;
                    b *= d;
                }
            }
            if (b > 0.0) { n0++; lastv = vi; }
            if (sd > 0) n1++;
        }
        s.b[vi] = b;
        s.sd[vi] = (unsigned short)sd;
    }
    __syncwarp();
    n0 = __reduce_add_sync(FULL_MASK, n0);
    n1 = __reduce_add_sync(FULL_MASK, n1);
    lastv = __reduce_max_sync(FULL_MASK, lastv);
    mu_min = __reduce_min_sync(FULL_MASK, mu_min);
    mu_max = __reduce_max_sync(FULL_MASK, mu_max);
    self_or_bad = __any_sync(FULL_MASK, self_or_bad);
    int mult = 0;
    if (!self_or_bad) mult = (mu_max == 0) ? 1 : (mu_min == mu_max ? mu_min : 0);
    const int level = (n0 > 0) ? 0 : ((n1 > 0) ? 1 : 2);

    // ---- Z and the root CDF. Every weight is an integer-valued double; if the total is <= 2^53 every
    //      partial sum is exact, so the parallel prefix equals the reference's left-to-right sum (:234).
    //      NARROW: lane holds position lane (cum0); else positions 2*lane and 2*lane+1 (cum0, cum1).
    double cum0, cum1 = 0.0, Z;
    {
        const double p0 = NARROW ? s.b[lane] : s.b[2 * lane], p1 = NARROW ? 0.0 : s.b[2 * lane + 1];
        double incl = p0 + p1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            double y = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= off) incl += y;
        }
        const double total = __shfl_sync(FULL_MASK, incl, 31);
        if (total <= 9007199254740992.0) {
            const double excl = incl - (p0 + p1);
            cum0 = excl + p0;
            cum1 = cum0 + p1;
            Z = total;
        } else {
            double run = 0.0;
            cum0 = cum1 = 0.0;
            for (int vi = 0; vi < 32 * H; ++vi) {  // sequential, identical on every lane
                run += s.b[vi];
                if (vi == (NARROW ? lane : 2 * lane)) cum0 = run;
                if (!NARROW && vi == 2 * lane + 1) cum1 = run;
            }
            Z = run;
        }
    }
    if (level == 1) {
        const u32 blo = __ballot_sync(FULL_MASK, s.sd[lane] > 0);
        if constexpr (NARROW) {
            cum0 = (double)__popc(blo & (0xffffffffu >> (31 - lane)));
        } else {
            const u32 bhi = __ballot_sync(FULL_MASK, s.sd[lane + 32] > 0);
            // positions 2*lane and 2*lane+1
            const int a = 2 * lane, bb = 2 * lane + 1;
            auto pref = [&](int vi) -> double {
                int c = (vi < 32) ? __popc(blo & (0xffffffffu >> (31 - vi)))
                                  : __popc(blo) + __popc(bhi & (0xffffffffu >> (63 - vi)));
                return (double)c;
            };
            cum0 = pref(a); cum1 = pref(bb);
        }
    } else if (level == 2) {
        if constexpr (NARROW) cum0 = (double)(lane + 1);
        else { cum0 = (double)(2 * lane + 1); cum1 = (double)(2 * lane + 2); }
    }

    // ---- write the tables -----------------------------------------------------------------------
    u32* rp = w.rowptr + nb + g;
    for (int v = lane; v <= n; v += 32) rp[v] = s.rowptr[v];
    for (int p = lane; p < nnz; p += 32) w.csr_idx[csr_base + p] = (i32)s.idx[p];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const int v = lane + 32 * h;
        if (v < n) {
            w.order[nb + v] = s.order[v];
            w.index_of[nb + v] = s.index_of[v];
            w.suffix_deg[nb + v] = s.sd[v];
            w.bucket_b[nb + v] = s.b[v];
            w.adj[nb + v] = s.adj[v];
            w.suf[nb + v] = s.suf[v];
        }
    }
    if constexpr (NARROW) {
        if (lane < n) w.cum[nb + lane] = cum0;
    } else {
        if (2 * lane < n) w.cum[nb + 2 * lane] = cum0;
        if (2 * lane + 1 < n) w.cum[nb + 2 * lane + 1] = cum1;
    }
    if (lane == 0) {
        ugs_ginfo gi;
        gi.Z = Z; gi.nnz = nnz; gi.level = level;
        gi.n_eligible = (level == 0) ? n0 : (level == 1 ? n1 : n);
        gi.last_viable = lastv; gi.mult = mult; gi.cls = 1;
        w.ginfo[g] = gi;
    }
}

#ifndef PRE_MINB
#define PRE_MINB 6  // min resident CTAs per SM asked of the compiler: 40 registers, 48 warps/SM (B200 A/B: 5 -> 0.389 ms,
#endif              // 6 -> 0.373 ms, 7 -> 0.405 ms with spills)
// KT > 0 fixes k at compile time: the reachability closure and the weight product unroll.
template <int WARPS, int KT>
__global__ void __launch_bounds__(WARPS * 32, PRE_MINB)
k_preproc_small(const i64* __restrict__ row0, const i64* __restrict__ row1, const i64* __restrict__ ptr,
                i64 G, i64 node_base, int k, const i64* __restrict__ perm, WS w) {
    __shared__ SmallPre sm_all[WARPS];
    if (KT) k = KT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 g = (i64)blockIdx.x * WARPS + warp;
    if (g >= G) return;
    if (w.ginfo[g].cls != 1) return;
    SmallPre& s = sm_all[warp];
    const i64 lo = ptr[g];
    const int n = (int)(ptr[g + 1] - lo);
    const i64 c0 = w.cstart[g];
    const int nc = (int)(w.cstart[g + 1] - c0);
    const i64 csr_base = 2 * c0;
    const i64 nb = lo - node_base;
    const u32 lt_mask = (1u << lane) - 1u;

    s.adj[lane] = 0; s.adj[lane + 32] = 0;
    s.deg[lane] = 0; s.deg[lane + 32] = 0;
    __syncwarp();

    // ---- A.2 degrees + adjacency bitmask --------------------------------------------------------
    for (int c = lane; c < nc; c += 32) {
        const i64 j = perm ? perm[c0 + c] : c0 + c;
        const i64 u = row0[j] - lo, v = row1[j] - lo;
        const bool u_ok = (u >= 0) && (u < n);  // always true unless the fast segmentation was fooled
        if (!u_ok) atomicOr(&w.counters[UGS_CTR_UNSORTED], 8ull);
        const bool valid = u_ok && (v >= 0) && (v < n);  // cross-graph / out-of-range columns are dropped (:53)
        s.cu[c] = (unsigned char)u;
        s.cv[c] = valid ? (unsigned char)v : (unsigned char)0xFF;
        if (valid) {
            atomicAdd(&s.deg[u], 1u);
            atomicAdd(&s.deg[v], 1u);
            // 32-bit halves: a native shared-memory atomic each (the 64-bit OR is a compare-and-swap loop)
            atomicOr(reinterpret_cast<u32*>(&s.adj[u]) + (v >> 5), 1u << (v & 31));
            atomicOr(reinterpret_cast<u32*>(&s.adj[v]) + (u >> 5), 1u << (u & 31));
        }
    }
    __syncwarp();
    int nnz;
    {
        const u32 d0 = s.deg[2 * lane], d1 = s.deg[2 * lane + 1];
        u32 incl = d0 + d1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            u32 y = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= off) incl += y;
        }
        const u32 excl = incl - (d0 + d1);
        s.rowptr[2 * lane] = (unsigned short)excl;
        s.rowptr[2 * lane + 1] = (unsigned short)(excl + d0);
        if (lane == 31) s.rowptr[64] = (unsigned short)incl;
        nnz = (int)__shfl_sync(FULL_MASK, incl, 31);
    }
    __syncwarp();
    s.deg[lane] = 0; s.deg[lane + 32] = 0;  // becomes the per-row cursor
    __syncwarp();
    // ---- stable fill: events (column, side) in column order; rank among same-row events of the chunk
    for (int base = 0; base < 2 * nc; base += 32) {
        const int e = base + lane;
        const int c = e >> 1, side = e & 1;
        bool valid = e < 2 * nc;
        int u = 0, v = 0xFF;
        if (valid) { u = s.cu[c]; v = s.cv[c]; valid = (v != 0xFF); }
        const int r = side ? v : u;
        const int val = side ? u : v;
        const u32 peers = __match_any_sync(FULL_MASK, valid ? r : (0x100 + lane));
        const int rank = __popc(peers & lt_mask);
        const u32 cur = valid ? s.deg[r] : 0u;
        __syncwarp();
        if (valid) {
            const int pos = (int)s.rowptr[r] + (int)cur + rank;
            s.idx[pos] = (unsigned char)val;
            w.csr_col[csr_base + pos] = (i32)(perm ? perm[c0 + c] : c0 + c);
            if (rank == __popc(peers) - 1) s.deg[r] = cur + (u32)__popc(peers);
        }
        __syncwarp();
    }
    // s.deg[v] is the multigraph degree again.

    // everything after the CSR: order, suffix masks, suffix degrees, reachability, weights, CDF, table stores. Graphs of
    // at most 32 vertices (all of ZINC / QM9 but a handful) take the 32-bit instantiation: one vertex per lane, every mask
    // operation a single instruction, half the shuffles.
    if (n <= 32) preproc_small_tail<KT, u32>(s, n, k, nnz, nb, g, csr_base, w);
    else preproc_small_tail<KT, u64>(s, n, k, nnz, nb, g, csr_base, w);
}

// =====================================================================================================
// General path: one block per graph, tables in global memory
// =====================================================================================================
#define GEN_THREADS 1024
#define GEN_SORT_SMEM 4096  // u64 keys sorted in shared memory when they fit

// Normalised bitonic network (all comparators ascending) on `a[0..n)`, any n: indices >= n act as +inf.
__device__ void block_bitonic_u64(u64* a, int n) {
    for (int k2 = 2; (k2 >> 1) < n; k2 <<= 1) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int p = i ^ (k2 - 1);
            if (p > i && p < n) { u64 x = a[i], y = a[p]; if (y < x) { a[i] = y; a[p] = x; } }
        }
        __syncthreads();
        for (int j = k2 >> 2; j >= 1; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i && p < n) { u64 x = a[i], y = a[p]; if (y < x) { a[i] = y; a[p] = x; } }
            }
            __syncthreads();
        }
    }
}

// Sort u64 keys living in global memory, staging through shared memory when they fit.
__device__ void block_sort_u64(u64* gkeys, int n, u64* smem_buf) {
    if (n <= GEN_SORT_SMEM) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) smem_buf[i] = gkeys[i];
        __syncthreads();
        block_bitonic_u64(smem_buf, n);
        for (int i = threadIdx.x; i < n; i += blockDim.x) gkeys[i] = smem_buf[i];
        __syncthreads();
    } else {
        block_bitonic_u64(gkeys, n);
    }
}

// Block-wide exclusive scan over n values produced by f(i), written through out(i, excl); returns total.
// Each thread owns a contiguous segment; the per-thread sums are scanned with warp shuffles (two levels: inside every
// warp, then over the warp totals), not by one thread walking all of them. For doubles the callers only rely on the
// result when every partial sum is an exactly representable integer, so the association order does not matter.
template <typename T, typename F, typename O>
__device__ T block_exclusive_scan(int n, F f, O out, T* sm /* blockDim.x + 1 */) {
    const int T_ = blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = T_ >> 5;
    const int seg = (n + T_ - 1) / T_;
    const int b = threadIdx.x * seg, e = min(n, b + seg);
    T local = 0;
    for (int i = b; i < e; ++i) local += f(i);
    T incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T y = __shfl_up_sync(FULL_MASK, incl, off);
        if (lane >= off) incl += y;
    }
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const T v = (lane < nw) ? sm[lane] : (T)0;
        T wi = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const T y = __shfl_up_sync(FULL_MASK, wi, off);
            if (lane >= off) wi += y;
        }
        if (lane < nw) sm[32 + lane] = wi - v;  // exclusive prefix of the warp totals
        if (lane == 31) sm[64] = wi;            // grand total (nw <= 32)
    }
    __syncthreads();
    T run = sm[32 + warp] + (incl - local);
    for (int i = b; i < e; ++i) { T v = f(i); out(i, run); run += v; }
    const T total = sm[64];
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(GEN_THREADS)
k_preproc_general(const i64* __restrict__ row0, const i64* __restrict__ row1, const i64* __restrict__ ptr,
                  i64 node_base, int k, const i64* __restrict__ perm, i64 n_general, WS w) {
    __shared__ u64 sbuf[GEN_SORT_SMEM];
    __shared__ double sscan[GEN_THREADS + 1];
    __shared__ int s_cnt[5];  // [4]: the graph has a self-loop column
    const int tid = threadIdx.x, T_ = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nwarps = T_ >> 5;
    const u32 lt_mask = (1u << lane) - 1u;

    if (n_general < 0) n_general = (i64)w.counters[UGS_CTR_N_GENERAL];  // fused small-batch path: count stays on the device
    for (i64 item = blockIdx.x; item < n_general; item += gridDim.x) {
        const i64 g = w.glist[item];
        const i64 lo = ptr[g];
        const int n = (int)(ptr[g + 1] - lo);
        if (n > UGS_LARGE_N) continue;  // large graphs: k_preproc_large (whole GPU per graph)
        if (w.ginfo[g].level >= 0) continue;  // already preprocessed by k_preproc_medium (shared-memory kernel)
        const i64 c0 = w.cstart[g];
        const i64 nc = w.cstart[g + 1] - c0;
        const i64 nb = lo - node_base;
        u32* rp = w.rowptr + nb + g;
        i32* ord = w.order + nb;
        i32* iof = w.index_of + nb;
        i32* sdg = w.suffix_deg + nb;
        double* bb = w.bucket_b + nb;
        double* cm = w.cum + nb;
        i32* cidx = w.csr_idx + 2 * c0;
        i32* ccol = w.csr_col + 2 * c0;

        // ---- A.2 degrees (iof = degree, ord = cursor) -------------------------------------------
        for (int v = tid; v < n; v += T_) { iof[v] = 0; ord[v] = 0; }
        if (tid < 5) s_cnt[tid] = 0;
        __syncthreads();
        for (i64 c = tid; c < nc; c += T_) {
            const i64 j = perm ? perm[c0 + c] : c0 + c;
            const i64 u = row0[j] - lo, v = row1[j] - lo;
            if (u < 0 || u >= n) { atomicOr(&w.counters[UGS_CTR_UNSORTED], 8ull); continue; }
            if (v >= 0 && v < n) { atomicAdd(&iof[u], 1); atomicAdd(&iof[v], 1); if (u == v) s_cnt[4] = 1; }
        }
        __syncthreads();
        {
            unsigned long long* su = (unsigned long long*)sscan;
            const unsigned long long tot = block_exclusive_scan<unsigned long long>(
                n, [&](int i) { return (unsigned long long)iof[i]; },
                [&](int i, unsigned long long ex) { rp[i] = (u32)ex; }, su);
            if (tid == 0) { rp[n] = (u32)tot; s_cnt[3] = (int)tot; }
        }
        __syncthreads();
        const int nnz = s_cnt[3];
        // ---- fill with atomic cursors (unordered), then sort every row by source column ---------
        for (i64 c = tid; c < nc; c += T_) {
            const i64 j = perm ? perm[c0 + c] : c0 + c;
            const i64 u = row0[j] - lo, v = row1[j] - lo;
            if (u >= 0 && u < n && v >= 0 && v < n) {
                const u32 pu = rp[u] + (u32)atomicAdd(&ord[u], 1);
                cidx[pu] = (i32)v; ccol[pu] = (i32)j;
                const u32 pv = rp[v] + (u32)atomicAdd(&ord[v], 1);
                cidx[pv] = (i32)u; ccol[pv] = (i32)j;
            }
        }
        __syncthreads();
        // rows of <= 32 entries: one warp, bitonic in registers; longer rows are queued in sdg[]
        for (int v = warp; v < n; v += nwarps) {
            const u32 p0 = rp[v];
            const int d = (int)(rp[v + 1] - p0);
            if (d <= 1) continue;
            if (d <= 32) {
                u64 key = (lane < d) ? ((((u64)(u32)ccol[p0 + lane]) << 32) | (u32)cidx[p0 + lane]) : ~0ull;
#pragma unroll
                for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
                    for (int j = k2 >> 1; j >= 1; j >>= 1) {
                        const u64 other = __shfl_xor_sync(FULL_MASK, key, j);
                        const bool up = ((lane & k2) == 0);
                        const bool lower = ((lane & j) == 0);
                        const bool take_min = (up == lower);
                        key = take_min ? (other < key ? other : key) : (other > key ? other : key);
                    }
                }
                if (lane < d) { ccol[p0 + lane] = (i32)(key >> 32); cidx[p0 + lane] = (i32)(u32)key; }
            } else if (d <= 128) {
                // medium row: this warp sorts it alone in its 1 KB slice of the sort buffer (4 keys per lane)
                u64* wb = sbuf + warp * 128;
                for (int i = lane; i < 128; i += 32)
                    wb[i] = (i < d) ? ((((u64)(u32)ccol[p0 + i]) << 32) | (u32)cidx[p0 + i]) : ~0ull;
                __syncwarp();
                for (int k2 = 2; k2 <= 128; k2 <<= 1) {
                    for (int j = k2 >> 1; j >= 1; j >>= 1) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int t = lane + 32 * q;  // 64 compare-exchanges per stage
                            const int lo_ = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                            const int hi_ = lo_ | j;
                            const bool up = ((lo_ & k2) == 0);
                            const u64 a = wb[lo_], b = wb[hi_];
                            if ((a > b) == up) { wb[lo_] = b; wb[hi_] = a; }
                        }
                        __syncwarp();
                    }
                }
                for (int i = lane; i < d; i += 32) { ccol[p0 + i] = (i32)(wb[i] >> 32); cidx[p0 + i] = (i32)(u32)wb[i]; }
                __syncwarp();
            } else if (lane == 0) {
                sdg[atomicAdd(&s_cnt[0], 1)] = v;
            }
        }
        __syncthreads();
        const int nlong = s_cnt[0];
        for (int t = 0; t < nlong; ++t) {
            const int v = sdg[t];
            const u32 p0 = rp[v];
            const int d = (int)(rp[v + 1] - p0);
            if (d <= GEN_SORT_SMEM) {
                for (int i = tid; i < d; i += T_) sbuf[i] = (((u64)(u32)ccol[p0 + i]) << 32) | (u32)cidx[p0 + i];
                __syncthreads();
                block_bitonic_u64(sbuf, d);
                for (int i = tid; i < d; i += T_) { ccol[p0 + i] = (i32)(sbuf[i] >> 32); cidx[p0 + i] = (i32)(u32)sbuf[i]; }
                __syncthreads();
            } else {
                // very long row: sort packed keys in the (still unused) cum[] scratch, chunk-free because
                // a row never has more entries than 2*nc; cum[] holds n doubles, so fall back to an
                // in-place pair network when the row is longer than n.
                if (d <= n) {
                    u64* scratch = (u64*)cm;
                    for (int i = tid; i < d; i += T_) scratch[i] = (((u64)(u32)ccol[p0 + i]) << 32) | (u32)cidx[p0 + i];
                    __syncthreads();
                    block_bitonic_u64(scratch, d);
                    for (int i = tid; i < d; i += T_) { ccol[p0 + i] = (i32)(scratch[i] >> 32); cidx[p0 + i] = (i32)(u32)scratch[i]; }
                    __syncthreads();
                } else {
                    for (int k2 = 2; (k2 >> 1) < d; k2 <<= 1) {
                        for (int j = -1; j != 0; ) {
                            const int jj = (j < 0) ? 0 : j;
                            for (int i = tid; i < d; i += T_) {
                                const int p = (j < 0) ? (i ^ (k2 - 1)) : (i ^ jj);
                                if (p > i && p < d) {
                                    const i32 ci = ccol[p0 + i], cp = ccol[p0 + p];
                                    if (cp < ci) {
                                        ccol[p0 + i] = cp; ccol[p0 + p] = ci;
                                        const i32 xi = cidx[p0 + i]; cidx[p0 + i] = cidx[p0 + p]; cidx[p0 + p] = xi;
                                    }
                                }
                            }
                            __syncthreads();
                            j = (j < 0) ? (k2 >> 2) : (j >> 1);
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- A.3 order = stable sort by (degree, id). Up to 2048 vertices (every PROTEINS-sized graph): rank by counting
        //      -- position of v = number of vertices with a smaller (degree, id) key, degrees staged in shared memory,
        //      two barriers; larger graphs: bitonic network over (degree << 32 | id) (55+ barriers) ------------------
        if (n <= 2048) {
            u32* sdeg = reinterpret_cast<u32*>(sbuf);
            for (int v = tid; v < n; v += T_) sdeg[v] = (u32)iof[v];
            __syncthreads();
            for (int v = tid; v < n; v += T_) {
                const u32 dv = sdeg[v];
                int rank = 0;
                for (int u = 0; u < v; ++u) rank += (sdeg[u] <= dv) ? 1 : 0;      // smaller id: ties come first
                for (int u = v + 1; u < n; ++u) rank += (sdeg[u] < dv) ? 1 : 0;
                ord[rank] = v;
                iof[v] = rank;
            }
        } else {
            u64* keys = (u64*)cm;
            for (int v = tid; v < n; v += T_) keys[v] = (((u64)(u32)iof[v]) << 32) | (u32)v;
            __syncthreads();
            block_sort_u64(keys, n, sbuf);
            for (int vi = tid; vi < n; vi += T_) {
                const int v = (int)(u32)keys[vi];
                ord[vi] = v;
                iof[v] = vi;
            }
        }
        __syncthreads();
        // ---- A.4 suffix degree + reachability. Both walk rows and look up the position of every neighbour: when the
        //      graph fits (PROTEINS-sized graphs always do), neighbour ids and positions are staged in shared memory as
        //      16-bit values first, so that a step of the bounded BFS costs shared-memory instead of L2 latency.
        unsigned short* s_idx = reinterpret_cast<unsigned short*>(sbuf);   // [nnz]  (sort buffer, idle by now)
        unsigned short* s_iof = reinterpret_cast<unsigned short*>(sscan);  // [n]
        const bool staged = (nnz <= GEN_SORT_SMEM * 4) && (n <= GEN_THREADS * 4);
        if (staged) {
            for (int p = tid; p < nnz; p += T_) s_idx[p] = (unsigned short)cidx[p];
            for (int v = tid; v < n; v += T_) s_iof[v] = (unsigned short)iof[v];
        }
        __syncthreads();
        for (int vi = warp; vi < n; vi += nwarps) {
            const int v = ord[vi];
            const u32 p0 = rp[v], p1 = rp[v + 1];
            int cnt = 0;
            if (staged) { for (u32 p = p0 + lane; p < p1; p += 32) cnt += ((int)s_iof[s_idx[p]] >= vi) ? 1 : 0; }
            else { for (u32 p = p0 + lane; p < p1; p += 32) cnt += (iof[cidx[p]] >= vi) ? 1 : 0; }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(FULL_MASK, cnt, off);
            if (lane == 0) sdg[vi] = cnt;
            // ---- reachability (component of v in the suffix graph has >= k vertices) + weight -------
            const int sd = cnt;
            bool reach = (k <= 1);
            if (!reach && sd > 0) {
                int q = (lane == 0) ? v : -1;  // lane j holds queue[j]
                int qn = 1;
                for (int h = 0; h < qn && !reach; ++h) {
                    const int u = __shfl_sync(FULL_MASK, q, h);
                    const u32 q0 = rp[u], q1 = rp[u + 1];
                    for (u32 pb = q0; pb < q1 && !reach; pb += 32) {
                        const u32 p = pb + lane;
                        int x = -1;
                        bool ok = false;
                        if (p < q1) {
                            if (staged) { x = s_idx[p]; ok = ((int)s_iof[x] >= vi); }
                            else { x = cidx[p]; ok = (iof[x] >= vi); }
                        }
                        for (int j = 0; j < qn; ++j) {
                            const int qj = __shfl_sync(FULL_MASK, q, j);  // every lane must take part
                            ok = ok && (x != qj);
                        }
                        const u32 peers = __match_any_sync(FULL_MASK, ok ? x : (-1 - lane));
                        ok = ok && ((peers & lt_mask) == 0);
                        const u32 ball = __ballot_sync(FULL_MASK, ok);
                        const int c = __popc(ball);
                        if (qn + c >= k) { reach = true; break; }
                        // append: destination lane j in [qn, qn+c) takes the (j-qn)-th newcomer
                        const int t = lane - qn;
                        const int src = (t >= 0 && t < c) ? select_bit32(ball, t) : 0;
                        const int xs = __shfl_sync(FULL_MASK, x, src);
                        if (t >= 0 && t < c) q = xs;
                        qn += c;
                    }
                }
            }
            if (lane == 0) {
                double b = 0.0;
                if (reach) {
                    const double d = (double)(sd > 1 ? sd : 1);
                    b = 1.0;
                    for (int t = 0; t < k - 1; ++t) b *= d;
                }
                bb[vi] = b;
            }
        }
        __syncthreads();
        // ---- eligibility counts -----------------------------------------------------------------
        {
            int c0n = 0, c1n = 0, lv = -1;
            for (int vi = tid; vi < n; vi += T_) {
                if (bb[vi] > 0.0) { c0n++; lv = max(lv, vi); }
                if (sdg[vi] > 0) c1n++;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                c0n += __shfl_xor_sync(FULL_MASK, c0n, off);
                c1n += __shfl_xor_sync(FULL_MASK, c1n, off);
                lv = max(lv, __shfl_xor_sync(FULL_MASK, lv, off));
            }
            if (lane == 0) {
                if (c0n) atomicAdd(&s_cnt[0 + 1], c0n);
                if (c1n) atomicAdd(&s_cnt[2], c1n);
            }
            // s_cnt[0] (long-row count) is dead now; reuse as last_viable via atomicMax after a reset
            __syncthreads();
            if (tid == 0) s_cnt[0] = -1;
            __syncthreads();
            if (lane == 0 && lv >= 0) atomicMax(&s_cnt[0], lv);
            __syncthreads();
        }
        const int n0 = s_cnt[1], n1 = s_cnt[2], lastv = s_cnt[0];
        const int level = (n0 > 0) ? 0 : ((n1 > 0) ? 1 : 2);
        // ---- root CDF ---------------------------------------------------------------------------
        double Z = 0.0;
        if (level == 0) {
            const double tot = block_exclusive_scan<double>(
                n, [&](int i) { return bb[i]; }, [&](int i, double ex) { cm[i] = ex + bb[i]; }, sscan);
            Z = tot;
            if (!(tot <= 9007199254740992.0)) {
                // partial sums may round: redo exactly as the reference does, left to right (:234)
                if (tid == 0) {
                    double run = 0.0;
                    int i = 0;
                    for (; i + 8 <= n; i += 8) {
                        double t0 = bb[i], t1 = bb[i + 1], t2 = bb[i + 2], t3 = bb[i + 3];
                        double t4 = bb[i + 4], t5 = bb[i + 5], t6 = bb[i + 6], t7 = bb[i + 7];
                        run += t0; cm[i] = run; run += t1; cm[i + 1] = run; run += t2; cm[i + 2] = run;
                        run += t3; cm[i + 3] = run; run += t4; cm[i + 4] = run; run += t5; cm[i + 5] = run;
                        run += t6; cm[i + 6] = run; run += t7; cm[i + 7] = run;
                    }
                    for (; i < n; ++i) { run += bb[i]; cm[i] = run; }
                    sscan[0] = run;
                }
                __syncthreads();
                Z = sscan[0];
                __syncthreads();
            }
        } else if (level == 1) {
            block_exclusive_scan<double>(
                n, [&](int i) { return sdg[i] > 0 ? 1.0 : 0.0; },
                [&](int i, double ex) { cm[i] = ex + (sdg[i] > 0 ? 1.0 : 0.0); }, sscan);
        } else {
            for (int vi = tid; vi < n; vi += T_) cm[vi] = (double)(vi + 1);
        }
        if (tid == 0) {
            ugs_ginfo gi;
            gi.Z = Z; gi.nnz = nnz; gi.level = level;
            gi.n_eligible = (level == 0) ? n0 : (level == 1 ? n1 : n);
            gi.last_viable = lastv; gi.cls = 2;
            gi.mult = s_cnt[4] ? 0 : -1;  // general path: -1 = no self-loop column (symmetric multigraph rows)
            w.ginfo[g] = gi;
        }
        __syncthreads();
    }
}

// =====================================================================================================
// Host launchers
// =====================================================================================================
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" int ugs_workspace_layout(int64_t E, int64_t G, int64_t NT, ugs_layout* L) {
    if (!L) return ugs_set_error(UGS_EINVAL, "layout pointer is NULL");
    if (E < 0 || G < 0 || NT < 0) return ugs_set_error(UGS_EINVAL, "negative size");
    if (E >= (1ll << 31) - 2 || NT >= (1ll << 31) - 2 || G >= (1ll << 31) - 2)
        return ugs_set_error(UGS_ELIMIT, "batch too large: E, G and total nodes must be < 2^31");
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L->cstart = take(sizeof(i64) * (size_t)(G + 2));
    L->ginfo = take(sizeof(ugs_ginfo) * (size_t)(G + 1));
    L->glist = take(sizeof(i32) * (size_t)(G + 1));
    L->counters = take(sizeof(i64) * UGS_CTR_COUNT);
    L->colkey = take(sizeof(i32) * (size_t)(E + 1));
    L->rowptr = take(sizeof(u32) * (size_t)(NT + G + 1));
    L->order = take(sizeof(i32) * (size_t)(NT + 1));
    L->index_of = take(sizeof(i32) * (size_t)(NT + 1));
    L->suffix_deg = take(sizeof(i32) * (size_t)(NT + 1));
    L->bucket_b = take(sizeof(double) * (size_t)(NT + 1));
    L->cum = take(sizeof(double) * (size_t)(NT + 1));
    L->adj = take(sizeof(u64) * (size_t)(NT + 1));
    L->suf = take(sizeof(u64) * (size_t)(NT + 1));
    L->csr_idx = take(sizeof(i32) * (size_t)(2 * E + 1));
    L->csr_col = take(sizeof(i32) * (size_t)(2 * E + 1));
    L->gaux = take(sizeof(double) * (size_t)(G + 1));
    L->llist = take(sizeof(i32) * (size_t)(G + 1));
    L->lscratch = take(sizeof(double) * 4096);
    L->total_bytes = off;
    return UGS_OK;
}

extern "C" int ugs_segment_batch(const int64_t* edge_index, int64_t E, const int64_t* ptr, int64_t G, int k,
                                 int skip_lt_k, int fast, const int64_t* perm, void* workspace, const ugs_layout* L, void* stream_) {
    if (!workspace || !L || (!ptr && G >= 0)) return ugs_set_error(UGS_EINVAL, "NULL workspace/layout/ptr");
    if (E > 0 && !edge_index) return ugs_set_error(UGS_EINVAL, "edge_index is NULL");
    if (k < 1) return ugs_set_error(UGS_EINVAL, "k must be >= 1");
    if (k > UGS_MAX_K) return ugs_set_error(UGS_ELIMIT, "k must be <= %d", UGS_MAX_K);
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(workspace, L);
    UGS_CUDA_CHECK(cudaMemsetAsync(w.counters, 0, sizeof(i64) * UGS_CTR_COUNT, st));
    const int T = 256;
    if (fast && !perm && G > 0) {
        // graph-contiguous columns: the classifying kernel finds the column ranges itself (one launch)
        k_classify<true><<<(unsigned)((G + T - 1) / T), T, 0, st>>>(ptr, G, k, skip_lt_k, edge_index, E, w);
        UGS_LAUNCH_CHECK("k_classify");
        return UGS_OK;
    }
    if (fast && !perm) {
        k_seg_fast<<<(unsigned)((G + 2 + T - 1) / T), T, 0, st>>>(edge_index, E, ptr, G, w.cstart);
        UGS_LAUNCH_CHECK("k_seg_fast");
    } else {
        if (E > 0) {
            k_seg_keys<<<(unsigned)((E + T - 1) / T), T, 0, st>>>(edge_index, E, ptr, G, perm, w.colkey);
            UGS_LAUNCH_CHECK("k_seg_keys");
        }
        k_seg_bounds<<<(unsigned)((E + 1 + T - 1) / T), T, 0, st>>>(w.colkey, E, G, w.cstart, w.counters);
        UGS_LAUNCH_CHECK("k_seg_bounds");
    }
    if (G > 0) {
        k_classify<false><<<(unsigned)((G + T - 1) / T), T, 0, st>>>(ptr, G, k, skip_lt_k, nullptr, E, w);
        UGS_LAUNCH_CHECK("k_classify");
    }
    return UGS_OK;
}

extern "C" int ugs_batch_info(const void* workspace, const ugs_layout* L, int64_t* h_out, void* stream_) {
    if (!workspace || !L || !h_out) return ugs_set_error(UGS_EINVAL, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream_;
    UGS_CUDA_CHECK(cudaMemcpyAsync(h_out, (const char*)workspace + L->counters, sizeof(i64) * 16,
                                   cudaMemcpyDeviceToHost, st));
    UGS_CUDA_CHECK(cudaStreamSynchronize(st));
    return UGS_OK;
}

// Counters without stalling the stream: the copy runs on a side stream that only waits for the work queued on `stream`
// BEFORE ugs_batch_info_begin, so kernels launched between begin and end (the speculative bitmask-path preprocessing)
// keep the GPU busy while the host reads the segmentation flags.
static cudaStream_t g_side_stream[64];
static cudaEvent_t g_side_event[64];
static bool g_side_ready[64];
static std::mutex g_side_mutex;  // ctypes releases the GIL: two Python threads may be inside the library at once

extern "C" int ugs_batch_info_begin(const void* workspace, const ugs_layout* L, int64_t* h_out, void* stream_) {
    if (!workspace || !L || !h_out) return ugs_set_error(UGS_EINVAL, "NULL argument");
    int dev = 0;
    UGS_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return ugs_set_error(UGS_ELIMIT, "device index %d", dev);
    std::lock_guard<std::mutex> lock(g_side_mutex);  // record + wait + copy must not interleave with another caller's
    if (!g_side_ready[dev]) {
        UGS_CUDA_CHECK(cudaStreamCreateWithFlags(&g_side_stream[dev], cudaStreamNonBlocking));
        UGS_CUDA_CHECK(cudaEventCreateWithFlags(&g_side_event[dev], cudaEventDisableTiming));
        g_side_ready[dev] = true;
    }
    UGS_CUDA_CHECK(cudaEventRecord(g_side_event[dev], (cudaStream_t)stream_));
    UGS_CUDA_CHECK(cudaStreamWaitEvent(g_side_stream[dev], g_side_event[dev], 0));
    UGS_CUDA_CHECK(cudaMemcpyAsync(h_out, (const char*)workspace + L->counters, sizeof(i64) * UGS_CTR_COUNT,
                                   cudaMemcpyDeviceToHost, g_side_stream[dev]));
    return UGS_OK;
}

extern "C" int ugs_batch_info_end(void) {
    int dev = 0;
    UGS_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !g_side_ready[dev]) return ugs_set_error(UGS_EINVAL, "ugs_batch_info_end without begin");
    UGS_CUDA_CHECK(cudaStreamSynchronize(g_side_stream[dev]));
    return UGS_OK;
}

int ugs_launch_preproc_large(const i64* row0, const i64* row1, const i64* ptr, i64 node_base, int k, const i64* perm,
                             i64 n_large, WS& w, cudaStream_t st);  // ugs_preproc_large.cu
int ugs_launch_preproc_medium(const i64* row0, const i64* row1, const i64* ptr, i64 G, i64 node_base, int k,
                              const i64* perm, i64 n_general, WS& w, cudaStream_t st);  // ugs_preproc_medium.cu

extern "C" int ugs_preproc_batch(const int64_t* edge_index, int64_t E, const int64_t* ptr, int64_t G,
                                 int64_t node_base, int k, const int64_t* perm, int64_t n_general, int64_t n_large,
                                 void* workspace, const ugs_layout* L, void* stream_) {
    return ugs_preproc_batch_parts(edge_index, E, ptr, G, node_base, k, perm, n_general, n_large, UGS_PREPROC_ALL,
                                   workspace, L, stream_);
}

extern "C" int ugs_preproc_batch_parts(const int64_t* edge_index, int64_t E, const int64_t* ptr, int64_t G,
                                       int64_t node_base, int k, const int64_t* perm, int64_t n_general,
                                       int64_t n_large, int parts, void* workspace, const ugs_layout* L,
                                       void* stream_) {
    if (!workspace || !L) return ugs_set_error(UGS_EINVAL, "NULL workspace/layout");
    if (k < 1 || k > UGS_MAX_K) return ugs_set_error(UGS_ELIMIT, "k must be in [1, %d]", UGS_MAX_K);
    if (G <= 0) return UGS_OK;
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(workspace, L);
    const int64_t* row0 = edge_index;
    const int64_t* row1 = edge_index ? edge_index + E : nullptr;
    // n_general / n_large < 0: the counts are not known on the host (fused small-batch path, ugs_sample_batch_fused): the
    // kernels are launched with grids sized for the worst case and read the counts from counters[] themselves.
    if ((parts & UGS_PREPROC_SMALL) && n_general < G) {
        constexpr int WARPS = 8;
        const unsigned grid = (unsigned)((G + WARPS - 1) / WARPS);
#define UGS_LAUNCH_PRE(KT_) k_preproc_small<WARPS, KT_><<<grid, WARPS * 32, 0, st>>>(row0, row1, ptr, G, node_base, k, perm, w)
        switch (k) {
            case 3: UGS_LAUNCH_PRE(3); break;
            case 4: UGS_LAUNCH_PRE(4); break;
            case 5: UGS_LAUNCH_PRE(5); break;
            case 6: UGS_LAUNCH_PRE(6); break;
            case 8: UGS_LAUNCH_PRE(8); break;
            default: UGS_LAUNCH_PRE(0); break;
        }
#undef UGS_LAUNCH_PRE
        UGS_LAUNCH_CHECK("k_preproc_small");
    }
    if (!(parts & UGS_PREPROC_GENERAL)) return UGS_OK;
    if (n_large != 0) {
        const int rc = ugs_launch_preproc_large(row0, row1, ptr, node_base, k, perm, n_large, w, st);
        if (rc != UGS_OK) return rc;
    }
    if (n_general < 0 || n_general > n_large) {
        // PROTEINS-sized graphs first, in shared memory; k_preproc_general then takes whatever is left (it skips graphs
        // whose ginfo.level is already set)
        const int rc = ugs_launch_preproc_medium(row0, row1, ptr, G, node_base, k, perm, n_general, w, st);
        if (rc != UGS_OK) return rc;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t bound = n_general < 0 ? G : n_general;
        const unsigned grid = (unsigned)(bound < (int64_t)sms ? bound : sms);
        k_preproc_general<<<grid, GEN_THREADS, 0, st>>>(row0, row1, ptr, node_base, k, perm, n_general, w);
        UGS_LAUNCH_CHECK("k_preproc_general");
    }
    return UGS_OK;
}
