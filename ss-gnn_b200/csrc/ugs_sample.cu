// ugs_sample.cu -- root draw, k-graphlet growth, induced-edge counting, scan and emission on sm_100a.
//
// Replaces, for the whole batch at once:
//   sample()            sampler.cpp:91-181    roots (A.5) + rand_grow (sampler.cpp:36-85, A.6)
//   edge extraction     sampler.cpp:205-280   (A.7) and the batch driver's concatenation / offsets
//                       (ugs_sampler_batch_extension.cpp:188-231,257-290)
// Bitmask path: one THREAD per sample, the graph's adjacency bitmasks / CSR staged in shared memory per warp.
// General path: one WARP per sample, CSR read through L2, the cut kept as a per-warp bitmap in shared memory.
// Both paths draw the same Philox words and pick the r-th smallest vertex id of the distinct cut, so they
// (and oracle/ugs_oracle.c) produce identical samples.
#include "ugs_common.cuh"
#include "ugs_hostio.h"

#define SAMP_WARPS 8
// Measured on B200 (profiles/README.md, A/B of the ZINC-shaped step): 4 resident CTAs (64 registers, 32 bytes of
// spill) 0.444 -> 0.412 ms, plus per-thread 16-byte row stores -> 0.406 ms; 5 CTAs (48 registers) is slower again.
// (Selecting the r-th cut bit by clearing r low bits instead of the popc search was slower: 0.61 ms.)
#ifndef SAMP_MINB
#define SAMP_MINB 4      // min resident CTAs per SM asked of the compiler (register budget)
#endif
#ifndef SAMP_PIPE
#define SAMP_PIPE 1      // 1: large batches take the persistent kernel with cp.async fetch-ahead (k_sample_small_pipe)
#endif
#ifndef SAMP_VECSTORE
#define SAMP_VECSTORE 1  // 1: every thread stores its own node row with 16-byte stores (even k <= 8)
#endif


// =====================================================================================================
// Bitmask path: sample + count
// =====================================================================================================
// r-th set bit of a mask: three byte-prefix popcounts locate the byte, a 256 x 8 table gives the position inside it
// (20 instructions against 33 for the five-level popcount search; A/B on the ZINC-shaped shard: 0.393 -> 0.368 ms).
// The table is copied into shared memory once per CTA (2 KB): random byte reads there cost a few bank-conflict
// replays, through L1 they cost one tag look-up per distinct line.
struct SelLut { unsigned char p[256][8]; };
static constexpr SelLut make_sel_lut() {
    SelLut t{};
    for (int b = 0; b < 256; ++b) {
        int r = 0;
        for (int j = 0; j < 8; ++j) t.p[b][j] = 0;
        for (int j = 0; j < 8; ++j)
            if ((b >> j) & 1) t.p[b][r++] = (unsigned char)j;
    }
    return t;
}
__device__ const SelLut g_sel_lut = make_sel_lut();

// the threads of the CTA copy one 8-byte row each; the caller synchronises the CTA
__device__ __forceinline__ void sel_lut_stage(unsigned char* lut) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x)
        reinterpret_cast<u64*>(lut)[b] = reinterpret_cast<const u64*>(&g_sel_lut)[b];
}

__device__ __forceinline__ int select_bit_lut(const unsigned char* __restrict__ lut, u32 m, int r) {
    const int c0 = __popc(m & 0xffu), c1 = __popc(m & 0xffffu), c2 = __popc(m & 0xffffffu);
    const bool p0 = r >= c0, p1 = r >= c1, p2 = r >= c2;
    const int bi = p2 ? 3 : p1 ? 2 : p0 ? 1 : 0;
    const int base = p2 ? c2 : p1 ? c1 : p0 ? c0 : 0;
    const u32 byte = __byte_perm(m, 0u, 0x4440u + (u32)bi);
    return 8 * bi + (int)lut[byte * 8 + (u32)(r - base)];
}
__device__ __forceinline__ int select_bit_lut(const unsigned char* __restrict__ lut, u64 m, int r) {
    const u32 lo = (u32)m;
    const int c = __popc(lo);
    const bool hi = r >= c;
    return (hi ? 32 : 0) + select_bit_lut(lut, hi ? (u32)(m >> 32) : lo, hi ? r - c : r);
}
struct __align__(16) SmallSamp {
    ugs_ginfo gi;  // 32 bytes, 16-byte aligned: read where needed instead of living in eight registers across the tile loop
    u64 adj[64];
    u64 suf[64];
    double cum[64];
    unsigned char tile[32][16];
    unsigned short rowptr[66];
    unsigned char order[64];
    unsigned char idx[2 * UGS_SMALL_MAX_COLS];
};

__device__ __forceinline__ int popc_m(u32 m) { return __popc(m); }
__device__ __forceinline__ int popc_m(u64 m) { return __popcll(m); }

// One sample: root draw (A.5), growth (A.6) and induced-edge count (A.7). MaskT = u32 when the graph has at
// most 32 vertices (every mask operation is then a single 32-bit instruction), u64 otherwise. KT > 0 fixes k at
// compile time so that the growth loop unrolls and the Philox refills land at fixed steps.
// the tables one sample needs when the graph's multiplicities are uniform (no CSR rows): 1600 bytes per graph
struct __align__(16) SlotSamp {
    u64 adj[64];
    u64 suf[64];
    double cum[64];
    unsigned char order[64];
};

// rowptr / idx: the graph's CSR rows in shared memory (only read when gi.mult == 0), or nullptr.
template <typename MaskT, int KT, typename TabT>
__device__ __forceinline__ void sample_one(const TabT& s, const unsigned short* rowptr, const unsigned char* idx,
                                           const unsigned char* __restrict__ lut, const ugs_ginfo& gi, int n, int k,
                                           u32 i, u32 c2, u32 c3, u32 k0, u32 k1, u64& nlo, u64& nhi, i64& cnt,
                                           int& n_incomplete) {
    const int kk = KT ? KT : k;
    Philox4 r = philox4x32_10(i, 0u, c2, c3, k0, k1);
    const double x = (gi.level == 0) ? root_x(r.x, r.y, gi.Z) : (double)__umulhi(r.x, (u32)gi.n_eligible);
    int a = 0, b = n - 1;
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (s.cum[mid] > x) b = mid; else a = mid + 1;
    }
    if (gi.level == 0 && !(s.cum[a] > x)) a = gi.last_viable;
    const int root_vi = a;
    const int v0 = s.order[root_vi];
    const MaskT sufm = (MaskT)s.suf[root_vi];
    MaskT S = (MaskT)1 << v0;
    MaskT nbr = (MaskT)s.adj[v0];
    nlo = (nlo & ~0xFFull) | (u64)v0;
    int size = 1;
    int pairs = 0;  // member pairs joined by an edge; uniform multiplicity, no self loop: entries = 2 * pairs * mult
#pragma unroll
    for (int step = 1; step < (KT ? KT : 16); ++step) {
        if (step >= kk) break;
        const MaskT cut = nbr & sufm & ~S;
        const int c = popc_m(cut);
        if (c == 0) break;
        const int wsel = (step + 1) & 3;
        if (wsel == 0) r = philox4x32_10(i, (u32)((step + 1) >> 2), c2, c3, k0, k1);
        const u32 wv = (wsel == 0) ? r.x : (wsel == 1) ? r.y : (wsel == 2) ? r.z : r.w;
        const int xnew = select_bit_lut(lut, cut, (int)__umulhi(wv, (u32)c));
        const MaskT ax = (MaskT)s.adj[xnew];
        pairs += popc_m(ax & S);
        S |= (MaskT)1 << xnew;
        nbr |= ax;
        if (step < 8) nlo = (nlo & ~(0xFFull << (8 * step))) | ((u64)xnew << (8 * step));
        else nhi = (nhi & ~(0xFFull << (8 * (step - 8)))) | ((u64)xnew << (8 * (step - 8)));
        size++;
    }
    if (size == kk) {
        int c = 0;
        if (gi.mult > 0) {
            c = 2 * pairs * gi.mult;
        } else if (rowptr) {
            for (int j = 0; j < kk; ++j) {
                const int v = (int)(((j < 8) ? (nlo >> (8 * j)) : (nhi >> (8 * (j - 8)))) & 0xFF);
                for (int p = rowptr[v]; p < rowptr[v + 1]; ++p) c += (int)((S >> idx[p]) & (MaskT)1);
            }
        }
        cnt = c;
    } else {
        n_incomplete++;
    }
}

// Rows of a skipped graph (n <= 0 or n < k): m rows of -1, no edges (ugs_sampler_batch_extension.cpp:132-143).
__device__ __forceinline__ void sample_rows_skipped(i64 g, int t0, int t1, i64 m, int k, i64* __restrict__ nodes_out,
                                                    i64* __restrict__ edge_ptr, u64* __restrict__ packed) {
    const int lane = threadIdx.x & 31;
    for (int t = t0; t < t1; ++t) {
        const i64 s0 = g * m + (i64)t * 32;
        const int nv = (int)min((i64)32, m - (i64)t * 32);
        for (int e = lane; e < nv * k; e += 32) nodes_out[s0 * k + e] = -1;
        if (lane < nv) edge_ptr[1 + s0 + lane] = 0;
        if (packed && lane < nv) packed[s0 + lane] = ~0ull;
    }
}

// All samples of tiles [t0, t1) of graph g (its tables already in shared memory: `s`), one thread per sample.
template <int KT, typename TabT>
__device__ __forceinline__ void sample_tiles_staged(const TabT& s, const unsigned short* rowptr, const unsigned char* idx,
                                                    const unsigned char* __restrict__ lut,
                                                    unsigned char (*tile)[16], const ugs_ginfo& gi, i64 g,
                                                    i64 lo, int n, int t0, int t1, bool first_chunk, i64 m, int k,
                                                    u64 seed, u64 gbase, u32 ibase, int seed_mode, const WS& w,
                                                    i64* __restrict__ nodes_out, i64* __restrict__ edge_ptr,
                                                    u64* __restrict__ packed) {
    const int lane = threadIdx.x & 31;
    const int r32 = 32 % k, q32 = 32 / k;
    const u64 gidx = gbase + (u64)g;
    // seed_mode 1: key = seed + graph index, graph counter 0 -- what a loop of single-graph calls with seed + i draws
    const u64 kseed = seed_mode ? seed + gidx : seed;
    const u64 cg = seed_mode ? 0ull : gidx;
    const u32 c2 = (u32)cg, c3 = (u32)(cg >> 32), k0 = (u32)kseed, k1 = (u32)(kseed >> 32);
    int n_incomplete = 0;
    int cnt_sum = 0;  // this lane's emitted-entry count over the item's tiles (early total, see UGS_CTR_EARLY_TOTAL)

    for (int t = t0; t < t1; ++t) {
        const i64 i = (i64)t * 32 + lane;  // sample index inside the graph
        const bool active = i < m;
        u64 nlo = ~0ull, nhi = ~0ull;      // packed node bytes, 0xFF = -1
        i64 cnt = 0;
        if (active) {
            if (n <= 32) sample_one<u32, KT, TabT>(s, rowptr, idx, lut, gi, n, k, (u32)i + ibase, c2, c3, k0, k1, nlo, nhi, cnt, n_incomplete);
            else sample_one<u64, KT, TabT>(s, rowptr, idx, lut, gi, n, k, (u32)i + ibase, c2, c3, k0, k1, nlo, nhi, cnt, n_incomplete);
        }
        const i64 s0 = g * m + (i64)t * 32;
        const int nv = (int)min((i64)32, m - (i64)t * 32);
        if (packed && active) packed[s0 + lane] = nlo;  // the emission pass reads 8 bytes per sample instead of 8k
        cnt_sum += (int)cnt;
#if SAMP_VECSTORE
        if (KT > 0 && KT <= 8 && (KT & 1) == 0) {
            // the row of lane i starts at byte (s0 + i) * 8k: 16-byte aligned for even k; a warp's rows are contiguous
            if (active) {
                longlong2* out = reinterpret_cast<longlong2*>(nodes_out + (s0 + lane) * KT);
                if (((nlo >> (8 * (KT > 0 ? KT - 1 : 0))) & 0xFFull) != 0xFFull) {  // complete row: the last entry is set, so all are
#pragma unroll
                    for (int q = 0; q < KT / 2; ++q)
                        out[q] = make_longlong2(lo + (i64)((u32)(nlo >> (16 * q)) & 0xFFu),
                                                lo + (i64)((u32)(nlo >> (16 * q + 8)) & 0xFFu));
                } else {
#pragma unroll
                    for (int q = 0; q < KT / 2; ++q) {
                        const u32 b0 = (u32)(nlo >> (16 * q)) & 0xFFu, b1 = (u32)(nlo >> (16 * q + 8)) & 0xFFu;
                        out[q] = make_longlong2((b0 == 0xFFu) ? (i64)-1 : lo + (i64)b0, (b1 == 0xFFu) ? (i64)-1 : lo + (i64)b1);
                    }
                }
            }
        } else
#endif
        {
            // ---- coalesced output through the shared tile ------------------------------------------------
            *reinterpret_cast<ulonglong2*>(&tile[lane][0]) = make_ulonglong2(nlo, nhi);
            __syncwarp();
            int sidx = lane / k, j = lane - sidx * k;
            i64* out = nodes_out + s0 * k;
            for (int e = lane; e < nv * k; e += 32) {
                const unsigned char bv = tile[sidx][j];
                out[e] = (bv == 0xFF) ? (i64)-1 : lo + (i64)bv;  // batch ext :194 adds lo
                j += r32; sidx += q32;
                if (j >= k) { j -= k; sidx++; }
            }
        }
        if (active) edge_ptr[1 + s0 + lane] = cnt;
        __syncwarp();
    }
    // statistics (the only cross-rank exchange later all-gathers these)
    n_incomplete = __reduce_add_sync(FULL_MASK, n_incomplete);
    cnt_sum = __reduce_add_sync(FULL_MASK, cnt_sum);
    if (lane == 0 && cnt_sum)
        atomicAdd(&w.counters[UGS_CTR_EARLY_TOTAL + (blockIdx.x & (UGS_CTR_EARLY_SLOTS - 1))], (unsigned long long)cnt_sum);
    if (lane == 0) {
        if (n_incomplete) atomicAdd(&w.counters[UGS_CTR_INCOMPLETE], (unsigned long long)n_incomplete);
        if (first_chunk && gi.level > 0) atomicAdd(&w.counters[UGS_CTR_RELAXED], 1ull);
    }
}

// Tiles [t0, t1) of graph g by one warp that stages the graph's tables into its shared-memory slice `s` first.
template <int KT>
__device__ __forceinline__ void sample_graph_tiles(SmallSamp& s, i64 g, int chunk, int tiles_per_item,
                                                   const i64* __restrict__ ptr, i64 node_base, i64 m, int k, u64 seed,
                                                   u64 gbase, u32 ibase, int seed_mode, const WS& w,
                                                   i64* __restrict__ nodes_out, i64* __restrict__ edge_ptr,
                                                   i64* __restrict__ sample_ptr, u64* __restrict__ packed,
                                                   const unsigned char* __restrict__ lut) {
    const int lane = threadIdx.x & 31;
    if (chunk == 0 && lane == 0) {
        sample_ptr[g + 1] = (g + 1) * m;  // ugs_sampler_batch_extension.cpp:130
        if (g == 0) { sample_ptr[0] = 0; edge_ptr[0] = 0; }
    }
    __syncwarp();  // a previous graph of this warp (k_sample_small_packed) is done with the slice
    if (lane < 2) reinterpret_cast<ulonglong2*>(&s.gi)[lane] = reinterpret_cast<const ulonglong2*>(&w.ginfo[g])[lane];
    __syncwarp();
    const ugs_ginfo& gi = s.gi;
    if (gi.cls == 2) return;
    if (gi.cls == 1 && k > 16) {  // the workspace was preprocessed for a smaller k: tile[][16] cannot hold the row
        if (lane == 0) atomicOr(&w.counters[UGS_CTR_UNSORTED], 16ull);
        return;
    }
    const int ntiles = (int)((m + 31) >> 5);
    const int t0 = chunk * tiles_per_item, t1 = min(ntiles, t0 + tiles_per_item);
    if (gi.cls == 0) {
        sample_rows_skipped(g, t0, t1, m, k, nodes_out, edge_ptr, packed);
        return;
    }
    const i64 lo = ptr[g];
    const int n = (int)(ptr[g + 1] - lo);
    const i64 nb = lo - node_base;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int v = lane + 32 * h;
        if (v < n) {
            s.adj[v] = w.adj[nb + v];
            s.suf[v] = w.suf[nb + v];
            s.cum[v] = w.cum[nb + v];
            s.order[v] = (unsigned char)w.order[nb + v];
        }
    }
    if (gi.mult == 0) {  // irregular multiplicities / self loops: counts need the CSR rows
        const u32* rp = w.rowptr + nb + g;
        const i32* ci = w.csr_idx + 2 * w.cstart[g];
        for (int v = lane; v <= n; v += 32) s.rowptr[v] = (unsigned short)rp[v];
        for (int p = lane; p < gi.nnz; p += 32) s.idx[p] = (unsigned char)ci[p];
    }
    __syncwarp();
    sample_tiles_staged<KT, SmallSamp>(s, s.rowptr, s.idx, lut, s.tile, gi, g, lo, n, t0, t1, chunk == 0, m, k, seed, gbase, ibase, seed_mode, w,
                                       nodes_out, edge_ptr, packed);
}

template <int KT>
__global__ void __launch_bounds__(SAMP_WARPS * 32, SAMP_MINB)
k_sample_small(const i64* __restrict__ ptr, i64 G, i64 node_base, i64 m, int k, u64 seed, u64 gbase, u32 ibase,
               int seed_mode, int tiles_per_item, i64 items_per_graph, WS w, i64* __restrict__ nodes_out,
               i64* __restrict__ edge_ptr, i64* __restrict__ sample_ptr, u64* __restrict__ packed) {
    __shared__ SmallSamp sm_all[SAMP_WARPS];
    __shared__ __align__(16) unsigned char lut[2048];
    if (KT) k = KT;
    sel_lut_stage(lut);
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const i64 item = (i64)blockIdx.x * SAMP_WARPS + warp;
    const i64 g = (items_per_graph == 1) ? item : item / items_per_graph;
    if (g >= G) return;
    const int chunk = (int)(item - g * items_per_graph);
    sample_graph_tiles<KT>(sm_all[warp], g, chunk, tiles_per_item, ptr, node_base, m, k, seed, gbase, ibase, seed_mode, w,
                           nodes_out, edge_ptr, sample_ptr, packed, lut);
}

// Large batches (at least two graphs per resident warp, every tile of a graph in one item): PERSISTENT warps that walk
// the graphs round-robin and fetch ahead with cp.async. Staging is three dependent global round trips (ginfo / ptr ->
// table addresses -> tables) and a quarter of the warp-time of k_sample_small (long-scoreboard stalls; one tile per
// staging instead of two costs +0.07 ms on the ZINC-shaped shard). Here, while graph j is sampled, the tables of graph
// j+1 (into the other buffer) and the ginfo / ptr pair of graph j+2 (a three-slot ring) are in flight; nothing is held
// in registers for them.
struct __align__(16) PipeTab {
    u64 adj[64];
    u64 suf[64];
    double cum[64];
    i32 order[64];
};
struct __align__(16) PipeWarp {
    PipeTab tab[2];
    ugs_ginfo gi[3];
    i64 pp[3][2];
    unsigned char tile[32][16];
    unsigned short rowptr[66];  // rowptr / idx: graphs with irregular multiplicities only, staged when the graph's turn comes
    unsigned char idx[2 * UGS_SMALL_MAX_COLS];
};

template <int KT>
__global__ void __launch_bounds__(SAMP_WARPS * 32, SAMP_MINB)
k_sample_small_pipe(const i64* __restrict__ ptr, int G, i64 node_base, i64 m, int k, u64 seed, u64 gbase, u32 ibase,
                    int seed_mode, WS w, i64* __restrict__ nodes_out, i64* __restrict__ edge_ptr,
                    i64* __restrict__ sample_ptr, u64* __restrict__ packed) {
    __shared__ PipeWarp sm_all[SAMP_WARPS];
    __shared__ __align__(16) unsigned char lut[2048];
    if (KT) k = KT;
    sel_lut_stage(lut);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    PipeWarp& pw = sm_all[warp];
    const int W = (int)gridDim.x * SAMP_WARPS;  // graphs are dealt round-robin to the grid's warps
    int g = (int)blockIdx.x * SAMP_WARPS + warp;
    if (g >= G) return;
    const int ntiles = (int)((m + 31) >> 5);

    // ginfo (32 bytes) and the ptr pair of graph gg -> ring slot
    auto fetch_meta = [&](int gg, int slot) {
        if (gg < G) {
            if (lane < 2) cp_async_8(&pw.pp[slot][lane], ptr + gg + lane);
            else if (lane < 4) cp_async_16(reinterpret_cast<char*>(&pw.gi[slot]) + 16 * (lane - 2),
                                           reinterpret_cast<const char*>(&w.ginfo[gg]) + 16 * (lane - 2));
        }
    };
    // tables of the graph whose meta data sit in `slot` -> buffer `buf` (bitmask-path graphs only: n <= 64)
    auto fetch_tabs = [&](int slot, int buf) {
        if (pw.gi[slot].cls != 1) return;
        const i64 lo = pw.pp[slot][0];
        const int n = min((int)(pw.pp[slot][1] - lo), UGS_SMALL_MAX_N);
        const i64 nb = lo - node_base;
        PipeTab& tb = pw.tab[buf];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int v = lane + 32 * h;
            if (v < n) {
                cp_async_8(&tb.adj[v], w.adj + nb + v);
                cp_async_8(&tb.suf[v], w.suf + nb + v);
                cp_async_8(&tb.cum[v], w.cum + nb + v);
                cp_async_4(&tb.order[v], w.order + nb + v);
            }
        }
    };

    fetch_meta(g, 0);
    cp_async_wait_all();
    __syncwarp();
    fetch_tabs(0, 0);
    fetch_meta(g + W, 1);
    int slot = 0, buf = 0;
    for (; g < G; g += W) {
        cp_async_wait_all();  // tables of this graph, meta data of the next
        __syncwarp();         // ... visible to every lane; and every lane is done with the previous graph's buffers
        const int slot1 = slot == 2 ? 0 : slot + 1, slot2 = slot1 == 2 ? 0 : slot1 + 1;
        if (g + W < G) fetch_tabs(slot1, buf ^ 1);
        fetch_meta(g + 2 * W, slot2);

        if (lane == 0) {
            sample_ptr[g + 1] = ((i64)g + 1) * m;  // ugs_sampler_batch_extension.cpp:130
            if (g == 0) { sample_ptr[0] = 0; edge_ptr[0] = 0; }
        }
        const ugs_ginfo& gi = pw.gi[slot];  // read from shared memory where needed, not held in registers
        if (gi.cls == 1 && k > 16) {
            if (lane == 0) atomicOr(&w.counters[UGS_CTR_UNSORTED], 16ull);
        } else if (gi.cls == 0) {
            sample_rows_skipped(g, 0, ntiles, m, k, nodes_out, edge_ptr, packed);
        } else if (gi.cls == 1) {
            const i64 lo = pw.pp[slot][0];
            const int n = (int)(pw.pp[slot][1] - lo);
            PipeTab& tb = pw.tab[buf];
            if (gi.mult == 0) {
                const u32* rp = w.rowptr + (lo - node_base) + g;
                const i32* ci = w.csr_idx + 2 * w.cstart[g];
                for (int v = lane; v <= n; v += 32) pw.rowptr[v] = (unsigned short)rp[v];
                for (int p = lane; p < gi.nnz; p += 32) pw.idx[p] = (unsigned char)ci[p];
                __syncwarp();
            }
            sample_tiles_staged<KT, PipeTab>(tb, pw.rowptr, pw.idx, lut, pw.tile, gi, g, lo, n, 0, ntiles, true, m, k, seed, gbase, ibase, seed_mode,
                                             w, nodes_out, edge_ptr, packed);
        }
        slot = slot1;
        buf ^= 1;
    }
}

// m <= 16 samples per graph (QM9-style completeness runs: m = 10) leave most lanes of a (graph, 32-sample tile) warp
// idle, so a warp takes 32 / S graphs at once, S = 8, 10 or 16 lanes each (m = 10, the QM9 configuration: three graphs
// per warp, lanes 30 and 31 idle): every lane group stages the tables of its own
// graph and samples it. Graphs that need their CSR rows for the entry count (irregular multiplicities, self loops) do
// not fit this layout: a warp that holds one handles its graphs one after the other with the code above.
template <int KT, int S>
__global__ void __launch_bounds__(SAMP_WARPS * 32, SAMP_MINB)
k_sample_small_packed(const i64* __restrict__ ptr, i64 G, i64 node_base, i64 m, int k, u64 seed, u64 gbase, u32 ibase,
                      int seed_mode, WS w, i64* __restrict__ nodes_out, i64* __restrict__ edge_ptr,
                      i64* __restrict__ sample_ptr, u64* __restrict__ packed) {
    constexpr int GPW = 32 / S;
    constexpr size_t SLICE = (sizeof(SlotSamp) * GPW > sizeof(SmallSamp)) ? sizeof(SlotSamp) * GPW : sizeof(SmallSamp);
    extern __shared__ __align__(16) unsigned char packed_smem_raw[];  // SAMP_WARPS x SLICE (51 KB for S = 8)
    unsigned char (*raw)[SLICE] = reinterpret_cast<unsigned char (*)[SLICE]>(packed_smem_raw);
    unsigned char* lut = packed_smem_raw + SAMP_WARPS * SLICE;  // 2 KB behind the slices
    if (KT) k = KT;
    sel_lut_stage(lut);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = lane / S, i = lane - slot * S;
    const i64 g_first = ((i64)blockIdx.x * SAMP_WARPS + warp) * GPW;
    if (g_first >= G) return;
    const i64 g = g_first + slot;
    const bool valid = slot < GPW && g < G;  // S = 10: lanes 30 and 31 belong to no graph
    ugs_ginfo gi;
    gi.cls = -1; gi.mult = 0; gi.level = 0; gi.Z = 0.0; gi.nnz = 0; gi.n_eligible = 0; gi.last_viable = 0;
    if (valid) gi = w.ginfo[g];
    if (__any_sync(FULL_MASK, valid && gi.cls == 1 && (gi.mult <= 0 || k > 16))) {
        for (int sl = 0; sl < GPW && g_first + sl < G; ++sl) {
            sample_graph_tiles<KT>(*reinterpret_cast<SmallSamp*>(raw[warp]), g_first + sl, 0, 1, ptr, node_base, m, k, seed,
                                   gbase, ibase, seed_mode, w, nodes_out, edge_ptr, sample_ptr, packed, lut);
            __syncwarp();
        }
        return;
    }
    if (valid && i == 0) {
        sample_ptr[g + 1] = (g + 1) * m;  // ugs_sampler_batch_extension.cpp:130
        if (g == 0) { sample_ptr[0] = 0; edge_ptr[0] = 0; }
    }
    SlotSamp& s = reinterpret_cast<SlotSamp*>(raw[warp])[valid ? slot : 0];
    const bool live = valid && gi.cls == 1;
    const i64 lo = valid ? ptr[g] : 0;
    const int n = live ? (int)(ptr[g + 1] - lo) : 0;
    const i64 nb = lo - node_base;
    for (int v = i; v < n; v += S) {
        s.adj[v] = w.adj[nb + v];
        s.suf[v] = w.suf[nb + v];
        s.cum[v] = w.cum[nb + v];
        s.order[v] = (unsigned char)w.order[nb + v];
    }
    __syncwarp();
    const bool narrow = __all_sync(FULL_MASK, n <= 32);
    const bool active = valid && (i64)i < m;
    const u64 gidx = gbase + (u64)g;
    const u64 kseed = seed_mode ? seed + gidx : seed;
    const u64 cg = seed_mode ? 0ull : gidx;
    const u32 c2 = (u32)cg, c3 = (u32)(cg >> 32), k0 = (u32)kseed, k1 = (u32)(kseed >> 32);
    int n_incomplete = 0;
    u64 nlo = ~0ull, nhi = ~0ull;
    i64 cnt = 0;
    if (live && active) {
        if (narrow) sample_one<u32, KT, SlotSamp>(s, nullptr, nullptr, lut, gi, n, k, (u32)i + ibase, c2, c3, k0, k1, nlo, nhi, cnt, n_incomplete);
        else sample_one<u64, KT, SlotSamp>(s, nullptr, nullptr, lut, gi, n, k, (u32)i + ibase, c2, c3, k0, k1, nlo, nhi, cnt, n_incomplete);
    }
    if (active && (gi.cls == 0 || gi.cls == 1)) {  // cls 0: n < k, rows of -1 (:132-143); cls 2: the general kernel's
        const i64 row = g * m + i;
        if (packed) packed[row] = nlo;
        edge_ptr[1 + row] = cnt;
        i64* out = nodes_out + row * k;
        if (KT > 0 && (KT & 1) == 0 && KT <= 8) {
            longlong2* o2 = reinterpret_cast<longlong2*>(out);
#pragma unroll
            for (int q = 0; q < KT / 2; ++q) {
                const u32 b0 = (u32)(nlo >> (16 * q)) & 0xFFu, b1 = (u32)(nlo >> (16 * q + 8)) & 0xFFu;
                o2[q] = make_longlong2((b0 == 0xFFu) ? (i64)-1 : lo + (i64)b0, (b1 == 0xFFu) ? (i64)-1 : lo + (i64)b1);
            }
        } else {
            for (int j = 0; j < k; ++j) {  // k > 16 only reaches here for skipped graphs: every entry is -1
                const u32 bv = (j >= 16) ? 0xFFu : (u32)(((j < 8) ? (nlo >> (8 * j)) : (nhi >> (8 * (j - 8)))) & 0xFFull);
                out[j] = (bv == 0xFFu) ? (i64)-1 : lo + (i64)bv;
            }
        }
    }
    n_incomplete = __reduce_add_sync(FULL_MASK, n_incomplete);
    const int n_relaxed = __popc(__ballot_sync(FULL_MASK, live && i == 0 && gi.level > 0));
    i64 cnt_sum = cnt;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) cnt_sum += __shfl_xor_sync(FULL_MASK, cnt_sum, off);
    if (lane == 0) {
        if (cnt_sum)
            atomicAdd(&w.counters[UGS_CTR_EARLY_TOTAL + (blockIdx.x & (UGS_CTR_EARLY_SLOTS - 1))], (unsigned long long)cnt_sum);
        if (n_incomplete) atomicAdd(&w.counters[UGS_CTR_INCOMPLETE], (unsigned long long)n_incomplete);
        if (n_relaxed) atomicAdd(&w.counters[UGS_CTR_RELAXED], (unsigned long long)n_relaxed);
    }
}

// =====================================================================================================
// General path: one warp per sample
// =====================================================================================================
// The cut of a sample is a bitmap over the graph's vertices in the warp's slice of shared memory, with a second level
// on top: bc[b] = number of set bits in words [32b, 32b+32), maintained incrementally. Picking the r-th smallest vertex
// of the cut is then one warp pass over the block counts (32 blocks = 32768 vertices per pass) plus one pass over the
// 32 words of the chosen block, instead of a scan of the whole n-bit bitmap at every growth step (a 10^5-node graph:
// 4 + 1 passes instead of 98). Rows are added with one atomicOr per qualifying entry; members of S that a row re-adds
// are taken out again by their owning lanes (k atomics per step instead of k shuffles per 32 row entries). After the
// growth the slice is wiped with vector stores and reused as the membership bitmap of S for the entry count.
#define GEN_WORK_CHUNK 4  // samples a warp claims per atomic on the work counter

// One sample of a general-path graph by one warp: root draw, growth on the two-level cut bitmap, induced-entry count.
// Returns the sample size; lane j < size holds S[j] in `mine`; cntE = induced CSR entries (valid when size == k).
__device__ __forceinline__ int sample_general_one(const u32* __restrict__ rp, const i32* __restrict__ cidx,
                                                  const i32* __restrict__ iof, const i32* __restrict__ ord,
                                                  const double* __restrict__ cm, const ugs_ginfo& gi, int n, int k, u32 ictr,
                                                  u32 c2, u32 c3, u32 k0, u32 k1, u32* bm, u32* bc, int lane, int& mine,
                                                  int& cntE_out) {
    {
                Philox4 r = philox4x32_10(ictr, 0u, c2, c3, k0, k1);
        const double x = (gi.level == 0) ? root_x(r.x, r.y, gi.Z) : (double)__umulhi(r.x, (u32)gi.n_eligible);
        int a = warp_cdf_search(cm, n, x, lane);
        if (gi.level == 0 && !(cm[a] > x)) a = gi.last_viable;
        const int root_vi = a;

        mine = -1;                 // lane j holds S[j] ...
        u32 my_p0 = 0, my_p1 = 0;  // ... and the bounds of its CSR row
        int size = 0;
        int cutn = 0;
        int xnew = ord[root_vi];
        const int words = (n + 31) >> 5;
        const int nblk = (words + 31) >> 5;
        for (int step = 0; step < k; ++step) {
            if (step > 0) {
                if (cutn == 0) break;
                const int wsel = (step + 1) & 3;
                if (wsel == 0) r = philox4x32_10(ictr, (u32)((step + 1) >> 2), c2, c3, k0, k1);
                const u32 wv = (wsel == 0) ? r.x : (wsel == 1) ? r.y : (wsel == 2) ? r.z : r.w;
                int rsel = (int)__umulhi(wv, (u32)cutn);
                // level 1: the block of 32 words that holds the r-th set bit (ascending vertex id)
                int blk = 0;
                if (nblk > 1) {
                    blk = -1;
                    for (int bb = 0; bb < nblk; bb += 32) {
                        const int c = (bb + lane < nblk) ? (int)bc[bb + lane] : 0;
                        const int incl = warp_incl_scan(c, lane);
                        const int tot = __shfl_sync(FULL_MASK, incl, 31);
                        if (rsel < tot) {
                            const int src = __ffs(__ballot_sync(FULL_MASK, incl > rsel)) - 1;
                            rsel -= __shfl_sync(FULL_MASK, incl - c, src);
                            blk = bb + src;
                            break;
                        }
                        rsel -= tot;
                    }
                    if (blk < 0) break;  // cannot happen: cutn is the sum of the block counts
                }
                // level 2: the word inside the block, then the bit inside the word
                {
                    const u32 word = bm[blk * 32 + lane];  // words beyond the graph are zero (bm_words is padded)
                    const int c = __popc(word);
                    const int incl = warp_incl_scan(c, lane);
                    const u32 ball = __ballot_sync(FULL_MASK, incl > rsel);
                    if (ball == 0u) break;  // cannot happen
                    const int src = __ffs(ball) - 1;
                    const int excl_src = __shfl_sync(FULL_MASK, incl - c, src);
                    const u32 word_src = __shfl_sync(FULL_MASK, word, src);
                    xnew = (blk * 32 + src) * 32 + select_bit32(word_src, rsel - excl_src);
                }
                if (lane == 0) { bm[xnew >> 5] &= ~(1u << (xnew & 31)); if (nblk > 1) bc[xnew >> 10] -= 1u; }
                cutn--;
                __syncwarp();
            }
            const u32 p0 = rp[xnew], p1 = rp[xnew + 1];
            if (lane == size) { mine = xnew; my_p0 = p0; my_p1 = p1; }
            size++;
            if (step == k - 1) break;  // the last member's row cannot extend the sample any more
            // add row(xnew) to the cut: suffix vertices (index_of >= root position) not yet in the cut
            int delta = 0;
            for (u32 p = p0 + lane; p < p1; p += 32) {
                const int y = cidx[p];
                if (iof[y] >= root_vi) {
                    const u32 bit = 1u << (y & 31);
                    const u32 old = atomicOr(&bm[y >> 5], bit);
                    if (!(old & bit)) { if (nblk > 1) atomicAdd(&bc[y >> 10], 1u); delta++; }
                }
            }
            __syncwarp();
            // members of S are not part of the cut: their lanes take them out again
            if (lane < size) {
                const u32 bit = 1u << (mine & 31);
                const u32 old = atomicAnd(&bm[mine >> 5], ~bit);
                if (old & bit) { if (nblk > 1) atomicSub(&bc[mine >> 10], 1u); delta--; }
            }
            cutn += __reduce_add_sync(FULL_MASK, delta);
            __syncwarp();
        }
        // ---- wipe the slice (what is left of the cut), then count the induced entries with S as the bitmap -------
        {
            uint4* b4 = reinterpret_cast<uint4*>(bm);
            const int n4 = (nblk * 32) >> 2;
            for (int q = lane; q < n4; q += 32) b4[q] = make_uint4(0u, 0u, 0u, 0u);
            for (int q = lane; q < nblk; q += 32) bc[q] = 0u;
        }
        __syncwarp();
        int cntE = 0;
        if (size == k) {
            if (lane < size) atomicOr(&bm[mine >> 5], 1u << (mine & 31));
            // A graph without self-loop columns (mult < 0) is a symmetric multigraph: v appears in row(u) as often as u in
            // row(v). The entries of the LONGEST member row that point into S are therefore counted from the other
            // members' rows (entries equal to that member) and the long row -- typically a hub -- is not walked at all.
            int skip = -1, hub = -1;
            if (gi.mult < 0) {
                const u32 key = (lane < size) ? (((my_p1 - my_p0) << 5) | (u32)lane) : 0u;
                const u32 best = __reduce_max_sync(FULL_MASK, key);
                if ((best >> 5) > 64u) { skip = (int)(best & 31u); hub = __shfl_sync(FULL_MASK, mine, skip); }
            }
            __syncwarp();
            for (int j = 0; j < size; ++j) {
                if (j == skip) continue;
                const u32 p0 = __shfl_sync(FULL_MASK, my_p0, j), p1 = __shfl_sync(FULL_MASK, my_p1, j);
                for (u32 p = p0 + lane; p < p1; p += 32) {
                    const int y = cidx[p];
                    cntE += (int)((bm[y >> 5] >> (y & 31)) & 1u) + (y == hub ? 1 : 0);
                }
            }
            cntE = __reduce_add_sync(FULL_MASK, cntE);
            __syncwarp();
            if (lane < size) atomicAnd(&bm[mine >> 5], ~(1u << (mine & 31)));
            __syncwarp();
        }
        cntE_out = cntE;
        return size;
    }
}

#ifndef MEDS_THREADS
#define MEDS_THREADS 128  // k_sample_medium: samples (threads) per CTA
#endif
#ifndef MEDS_MIN_SAMPLES
#define MEDS_MIN_SAMPLES 10000  // general-path samples of a batch from which k_sample_medium takes the PROTEINS-sized graphs
#endif
#ifndef GEN_MINB
#define GEN_MINB 5  // resident CTAs per SM the kernel is compiled for (51 registers -> 40 warps per SM)
#endif
// PROTEINS-sized general-path graphs (ugs_is_medium: <= 1024 vertices, <= 3072 columns) are sampled by k_sample_medium
// with their tables in shared memory; this kernel takes the rest.

__global__ void __launch_bounds__(256, GEN_MINB)
k_sample_general(const i64* __restrict__ ptr, i64 node_base, i64 m, int k, u64 seed, u64 gbase, u32 ibase,
                 int seed_mode, i64 n_general, int bm_words, int bm_stride, WS w,
                 i64* __restrict__ nodes_out, i64* __restrict__ edge_ptr) {
    extern __shared__ __align__(16) u32 bm_all[];
    if (n_general < 0) n_general = (i64)w.counters[UGS_CTR_N_GENERAL];  // fused small-batch path
    if (n_general == 0) return;
    // a batch with many general-path samples: PROTEINS-sized graphs belong to k_sample_medium (see there)
    const bool skip_medium = n_general * m >= MEDS_MIN_SAMPLES;
    if (skip_medium && w.counters[UGS_CTR_N_NOT_MEDIUM] == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    u32* bm = bm_all + (size_t)warp * bm_stride;  // [bm_words] cut bitmap (bm_words is a multiple of 32)
    u32* bc = bm + bm_words;                      // [bm_words / 32] set bits per block of 32 words
    for (int i = lane; i < bm_stride; i += 32) bm[i] = 0;
    __syncwarp();
    const i64 total = n_general * m;
    unsigned long long n_incomplete = 0, n_relaxed = 0, n_entries = 0;

    // a claim covers GEN_WORK_CHUNK samples when there is plenty of work, one sample when every warp gets only a few
    const int chunk = (total >= (i64)gridDim.x * (blockDim.x >> 5) * (2 * GEN_WORK_CHUNK)) ? GEN_WORK_CHUNK : 1;
    while (true) {
        i64 tbase = 0;
        if (lane == 0) tbase = (i64)atomicAdd(&w.counters[UGS_CTR_WORK], (unsigned long long)chunk);
        tbase = __shfl_sync(FULL_MASK, tbase, 0);
        if (tbase >= total) break;
        const i64 tend = min(total, tbase + chunk);
        for (i64 t = tbase; t < tend; ++t) {
            const i64 gi_ = t / m;
            const i64 i = t - gi_ * m;
            const i64 g = w.glist[gi_];
            const i64 lo = ptr[g];
            const int n = (int)(ptr[g + 1] - lo);
            if (skip_medium && ugs_is_medium(n, w.cstart[g + 1] - w.cstart[g])) continue;  // k_sample_medium
            const ugs_ginfo gi = w.ginfo[g];
            const i64 nb = lo - node_base;
            const u64 gidx = gbase + (u64)g;
            const u64 kseed = seed_mode ? seed + gidx : seed;
            const u64 cg = seed_mode ? 0ull : gidx;
            if (i == 0 && gi.level > 0) n_relaxed++;
            int mine = -1, cntE = 0;
            const int size = sample_general_one(w.rowptr + nb + g, w.csr_idx + 2 * w.cstart[g], w.index_of + nb, w.order + nb,
                                                w.cum + nb, gi, n, k, (u32)i + ibase, (u32)cg, (u32)(cg >> 32), (u32)kseed,
                                                (u32)(kseed >> 32), bm, bc, lane, mine, cntE);
            const i64 sidx = g * m + i;
            if (lane < k) nodes_out[sidx * k + lane] = (lane < size) ? lo + (i64)mine : (i64)-1;
            if (lane == 0) {
                edge_ptr[1 + sidx] = (size == k) ? (i64)cntE : 0;
                if (size < k) n_incomplete++;
                else n_entries += (unsigned long long)cntE;
            }
        }
    }
    if (lane == 0) {
        if (n_entries) atomicAdd(&w.counters[UGS_CTR_EARLY_TOTAL + (blockIdx.x & (UGS_CTR_EARLY_SLOTS - 1))], n_entries);
        if (n_incomplete) atomicAdd(&w.counters[UGS_CTR_INCOMPLETE], n_incomplete);
        if (n_relaxed) atomicAdd(&w.counters[UGS_CTR_RELAXED], n_relaxed);
    }
}

// PROTEINS-sized general-path graphs (ugs_is_medium) of a batch with at least MEDS_MIN_SAMPLES general-path samples: one
// THREAD per sample, one CTA per (graph, blockDim.x samples). The CTA stages the graph's row pointers, neighbour ids,
// order positions, order and root CDF in shared memory once (<= 44 KB, by asynchronous copies that are all in flight
// together); every thread keeps its cut and its member set as bitmaps of ceil(n / 32) words in shared memory, word w of
// thread t at [w * blockDim.x + t] (always bank t % 32: conflict-free whatever words the threads of a warp touch). A
// growth step is a popcount walk over the words to the r-th set bit and one walk over the new member's row, 32 samples
// per instruction, where the warp-per-sample kernel spends ~300 instructions per sample and step on scans, votes and
// shuffles: 34 M warp instructions for the 17 000 general-path samples of a 1113-graph PROTEINS batch, issue-bound at
// 54 us. The price is latency -- one thread walks the whole chain of a sample, ~25 us whatever the batch -- which is why
// small batches stay with k_sample_general.
struct __align__(16) MedSample {
    double cum[1024];
    i32 idx[6144];
    u32 rowptr[1028];  // 1025 used
    i32 iof[1024];
    i32 ord[1024];
};

__global__ void __launch_bounds__(MEDS_THREADS)
k_sample_medium(const i64* __restrict__ ptr, i64 G, i64 node_base, i64 m, int k, u64 seed, u64 gbase, u32 ibase,
                int seed_mode, i64 n_general, i64 chunks_per_graph, WS w, i64* __restrict__ nodes_out,
                i64* __restrict__ edge_ptr) {
    extern __shared__ __align__(16) unsigned char meds_raw[];
    MedSample& s = *reinterpret_cast<MedSample*>(meds_raw);
    const int T = blockDim.x;
    u32* bm = reinterpret_cast<u32*>(meds_raw + sizeof(MedSample)) + threadIdx.x;  // cut, word w of this thread: bm[w * T]
    u32* ms = bm + 32 * T;                                                         // members of S, same layout
    if (n_general < 0) n_general = (i64)w.counters[UGS_CTR_N_GENERAL];  // fused small-batch path
    if (n_general * m < MEDS_MIN_SAMPLES) return;  // few samples: k_sample_general, a warp per sample, is the faster one
    int n_incomplete = 0, n_relaxed = 0;
    unsigned long long n_entries = 0;
    for (i64 item = blockIdx.x; item < n_general * chunks_per_graph; item += gridDim.x) {
        const i64 gi_ = item / chunks_per_graph;
        const i64 c = item - gi_ * chunks_per_graph;
        const i64 g = w.glist[gi_];
        const i64 lo = ptr[g];
        const int n = (int)(ptr[g + 1] - lo);
        const i64 c0 = w.cstart[g];
        if (!ugs_is_medium(n, w.cstart[g + 1] - c0)) continue;  // k_sample_general
        const ugs_ginfo gi = w.ginfo[g];
        const i64 nb = lo - node_base;
        __syncthreads();  // the previous item's samples are done with the tables
        {
            const u32* rp = w.rowptr + nb + g;
            const i32* ci = w.csr_idx + 2 * c0;  // 8-byte aligned, nnz = 2 * columns
            for (int v = threadIdx.x; v <= n; v += T) cp_async_4(&s.rowptr[v], rp + v);
            for (int v = threadIdx.x; v < n; v += T) {
                cp_async_8(&s.cum[v], w.cum + nb + v);
                cp_async_4(&s.iof[v], w.index_of + nb + v);
                cp_async_4(&s.ord[v], w.order + nb + v);
            }
            for (int p = threadIdx.x; 2 * p < gi.nnz; p += T) cp_async_8(&s.idx[2 * p], ci + 2 * p);
            cp_async_wait_all();
        }
        __syncthreads();
        if (c == 0 && threadIdx.x == 0 && gi.level > 0) n_relaxed++;
        const i64 i = c * T + threadIdx.x;
        if (i >= m) continue;
        const u64 gidx = gbase + (u64)g;
        const u64 kseed = seed_mode ? seed + gidx : seed;
        const u64 cg = seed_mode ? 0ull : gidx;
        const u32 ictr = (u32)i + ibase, c2 = (u32)cg, c3 = (u32)(cg >> 32), k0 = (u32)kseed, k1 = (u32)(kseed >> 32);
        const int words = (n + 31) >> 5;
        for (int q = 0; q < words; ++q) { bm[q * T] = 0u; ms[q * T] = 0u; }

        Philox4 r = philox4x32_10(ictr, 0u, c2, c3, k0, k1);
        const double x = (gi.level == 0) ? root_x(r.x, r.y, gi.Z) : (double)__umulhi(r.x, (u32)gi.n_eligible);
        int root_vi;
        {   // first position with cum > x (n - 1 when there is none), as warp_cdf_search
            int a = 0, b = n;
            while (a < b) { const int mid = (a + b) >> 1; if (s.cum[mid] > x) b = mid; else a = mid + 1; }
            if (a >= n) a = n - 1;
            if (gi.level == 0 && !(s.cum[a] > x)) a = gi.last_viable;
            root_vi = a;
        }
        unsigned short S[UGS_MAX_K];
        int size = 0, cutn = 0;
        int pairs = 0, selfs = 0;  // row entries that point at an EARLIER member / at the member itself
        int xnew = s.ord[root_vi];
        for (int step = 0; step < k; ++step) {
            if (step > 0) {
                if (cutn == 0) break;
                const int wsel = (step + 1) & 3;
                if (wsel == 0) r = philox4x32_10(ictr, (u32)((step + 1) >> 2), c2, c3, k0, k1);
                const u32 wv = (wsel == 0) ? r.x : (wsel == 1) ? r.y : (wsel == 2) ? r.z : r.w;
                int rsel = (int)__umulhi(wv, (u32)cutn);
                int q = 0;
                u32 word = 0u;
                for (; q < words; ++q) {  // the r-th set bit in ascending vertex id
                    word = bm[q * T];
                    const int pc = __popc(word);
                    if (rsel < pc) break;
                    rsel -= pc;
                }
                if (q >= words) break;  // cannot happen: cutn is the number of set bits
                const int bit = select_bit32(word, rsel);
                bm[q * T] = word & ~(1u << bit);
                xnew = q * 32 + bit;
                cutn--;
            }
            S[size++] = (unsigned short)xnew;
            ms[(xnew >> 5) * T] |= 1u << (xnew & 31);
            // One walk over row(xnew) does both jobs: entries that point into S are counted (see below), the other
            // suffix vertices (index_of >= root position) join the cut -- except after the last member, whose row
            // cannot extend the sample any more.
            const bool grow = step < k - 1;
            for (u32 p = s.rowptr[xnew], p1 = s.rowptr[xnew + 1]; p < p1; ++p) {
                const int y = s.idx[p];
                const u32 bit = 1u << (y & 31);
                if (ms[(y >> 5) * T] & bit) {
                    if (y == xnew) selfs++; else pairs++;
                } else if (grow && s.iof[y] >= root_vi) {
                    const u32 old = bm[(y >> 5) * T];
                    if (!(old & bit)) { bm[(y >> 5) * T] = old | bit; cutn++; }
                }
            }
        }
        // Induced CSR entries. Every column is stored in both directions, so u appears in row(v) as often as v in row(u):
        // the entries between two members are twice those seen from the LATER member's row at the moment it joined.
        int cntE = 0;
        if (size == k) {
            cntE = 2 * pairs + selfs;
            n_entries += (unsigned long long)cntE;
        } else {
            n_incomplete++;
        }
        const i64 sidx = g * m + i;
        for (int j = 0; j < k; ++j) nodes_out[sidx * k + j] = (j < size) ? lo + (i64)S[j] : (i64)-1;
        edge_ptr[1 + sidx] = (i64)cntE;
    }
    // per-warp totals (a thread's entries fit 32 bits: at most 2 * 3072 * k per sample, few samples per thread)
    const u32 e32 = __reduce_add_sync(FULL_MASK, (u32)n_entries);
    const int inc = __reduce_add_sync(FULL_MASK, n_incomplete);
    const int rel = __reduce_add_sync(FULL_MASK, n_relaxed);
    if ((threadIdx.x & 31) == 0) {
        if (e32) atomicAdd(&w.counters[UGS_CTR_EARLY_TOTAL + (blockIdx.x & (UGS_CTR_EARLY_SLOTS - 1))], (unsigned long long)e32);
        if (inc) atomicAdd(&w.counters[UGS_CTR_INCOMPLETE], (unsigned long long)inc);
        if (rel) atomicAdd(&w.counters[UGS_CTR_RELAXED], (unsigned long long)rel);
    }
}

// =====================================================================================================
// In-place inclusive scan of the per-sample counts (single pass, decoupled look-back)
// (Fusing this scan into k_sample_small -- CTAs in ticket order publishing their totals and looking back -- was
// measured slower on B200: 470 us against 321 + 46 us. A sampling CTA lives ~12 us and ~600 of them finish together,
// so every CTA walks a long look-back window while it still holds its SM slot.)
// =====================================================================================================
#ifndef SCAN_THREADS
#define SCAN_THREADS 256
#endif
#ifndef SCAN_ITEMS
#define SCAN_ITEMS 16
#endif
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_counts(i64* __restrict__ data, i64 B, unsigned long long* __restrict__ scratch,
              unsigned long long* __restrict__ counters, volatile i64* snap, i64 snap_seq) {
    // Coalesced global access (a warp reads 32 consecutive counts per instruction), transposed through a padded
    // shared tile so that every thread scans SCAN_ITEMS consecutive counts in registers.
    __shared__ i64 s_tile_[SCAN_THREADS / 32][32 * (SCAN_ITEMS + 1)];
    __shared__ i64 s_warp[SCAN_THREADS / 32];
    __shared__ i64 s_prefix;
    __shared__ unsigned int s_tile;
    unsigned long long* ticket = scratch;
    volatile unsigned long long* status = scratch + 1;  // value << 2 | flag (1 = aggregate, 2 = inclusive prefix)
    if (threadIdx.x == 0) s_tile = (unsigned int)atomicAdd(ticket, 1ull);
    __syncthreads();
    const i64 tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    i64* st = s_tile_[warp];
    const i64 wbase = tile * SCAN_TILE + (i64)warp * (32 * SCAN_ITEMS);  // this warp's 512 consecutive counts
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        const int it = q * 32 + lane;  // item index inside the warp chunk
        const i64 gi_ = wbase + it;
        st[(it / SCAN_ITEMS) * (SCAN_ITEMS + 1) + (it % SCAN_ITEMS)] = (gi_ < B) ? data[gi_] : 0;
    }
    __syncwarp();
    i64 v[SCAN_ITEMS];
    i64 local = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        v[q] = st[lane * (SCAN_ITEMS + 1) + q];
        local += v[q];
    }
    i64 incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const i64 y = __shfl_up_sync(FULL_MASK, incl, off);
        if (lane >= off) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    i64 warp_excl = 0, block_total = 0;
#pragma unroll
    for (int q = 0; q < SCAN_THREADS / 32; ++q) {
        const i64 x = s_warp[q];
        if (q < warp) warp_excl += x;
        block_total += x;
    }
    if (warp == 0) {
        // decoupled look-back, one warp: 32 predecessors per step
        i64 prefix = 0;
        if (tile == 0) {
            if (lane == 0) status[0] = ((unsigned long long)block_total << 2) | 2ull;
        } else {
            if (lane == 0) status[tile] = ((unsigned long long)block_total << 2) | 1ull;
            i64 j = tile - 1;
            while (true) {
                const i64 idx = j - lane;
                unsigned long long sv = (idx >= 0) ? status[idx] : 2ull;  // before tile 0: prefix 0, inclusive
                while (__any_sync(FULL_MASK, (sv & 3ull) == 0ull)) sv = (idx >= 0) ? status[idx] : 2ull;
                const u32 incl_mask = __ballot_sync(FULL_MASK, (sv & 3ull) == 2ull);
                const int first = incl_mask ? (__ffs(incl_mask) - 1) : 32;  // nearest predecessor with a full prefix
                i64 contrib = (lane <= first) ? (i64)(sv >> 2) : 0;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) contrib += __shfl_xor_sync(FULL_MASK, contrib, off);
                prefix += contrib;
                if (incl_mask) break;
                j -= 32;
            }
            if (lane == 0) status[tile] = ((unsigned long long)(prefix + block_total) << 2) | 2ull;
        }
        if (lane == 0) {
            s_prefix = prefix;
            if ((tile + 1) * SCAN_TILE >= B) counters[UGS_CTR_TOTAL_EDGES] = (unsigned long long)(prefix + block_total);
        }
        // Fused small-batch path: everything the host waits for (total, class counts, flags, statistics) is final here,
        // so the last tile writes counters[] straight into mapped host memory, sequence word last: the host polls that
        // word and never synchronises with the emission kernels behind this one.
        if (snap && (tile + 1) * SCAN_TILE >= B) {
            for (int q = lane; q < UGS_CTR_COUNT - 1; q += 32)
                snap[q] = (q == UGS_CTR_TOTAL_EDGES) ? (prefix + block_total) : (i64)counters[q];
            __threadfence_system();
            __syncwarp();
            if (lane == 0) {
                __threadfence_system();
                snap[UGS_CTR_COUNT - 1] = snap_seq;
            }
        }
    }
    __syncthreads();
    i64 run = s_prefix + warp_excl + (incl - local);
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        run += v[q];
        st[lane * (SCAN_ITEMS + 1) + q] = run;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        const int it = q * 32 + lane;
        const i64 gi_ = wbase + it;
        if (gi_ < B) data[gi_] = st[(it / SCAN_ITEMS) * (SCAN_ITEMS + 1) + (it % SCAN_ITEMS)];
    }
}

// =====================================================================================================
// Bitmask path: emission. One thread walks the CSR rows of its sample (a single flattened loop over
// (j, p), so lanes only wait for the longest sample, not the longest row), packs every hit into a staging
// word in shared memory at the sample's final offset, and the warp then flushes the staged run with fully
// coalesced 8-byte stores (32 consecutive output entries per store instruction).
// The kernel is bound by shared-memory latency per resident warp, and shared memory per warp bounds the
// resident warps: for k <= 8 in mode "sample" a staged entry is 16 bits (payload:10 | l:3 | j:3) and the
// local-index map holds 4-bit positions, which brings a warp to 5.5 KB and 40 warps per SM (48 registers);
// other modes stage 32 bits with a byte map (9.6 KB, 24 warps per SM).
// (Tried and measured slower, see profiles/README.md: 8 lanes per CSR row with ballot compaction; membership
// and local index from registers instead of lmap[]; two entries per walk step; per-lane node loads.)
// =====================================================================================================
#ifndef EMIT_WARPS
#define EMIT_WARPS 8
#endif
// (Prefetching the tile's node rows / offsets to L2 at kernel entry changed nothing: 1.272 vs 1.274 ms.)
#define EMIT_STAGE 1280  // staged entries per flush; a sample never has more than 768 (= 2 * max columns)
#define LMAP_STRIDE 68   // 17 words: the 32 per-sample rows start in 32 different banks

#ifndef EMIT_MINB
#define EMIT_MINB 5    // resident CTAs per SM the NARROW kernel is compiled for (48 registers)
#endif

template <typename StageT>
struct __align__(16) SmallEmit {
    StageT stage[EMIT_STAGE];
    u32 rowrange[64];                       // byte offsets into ent[]: 2*rowptr[v] | 2*rowptr[v+1] << 16
    unsigned char tile[32][16];             // node rows as local ids, 0xFF = -1
    unsigned char lmap[32][LMAP_STRIDE];    // local vertex id -> position inside its sample, 0xFF = not a member
    unsigned short ent[2 * UGS_SMALL_MAX_COLS];  // CSR entry: neighbour id | payload << 6
};
// B200 A/B of the ZINC-shaped step (profiles/README.md): byte lmap, 4 CTAs/SM 1.27-1.29 ms; nibble lmap, 4 CTAs 1.27 ms;
// nibble lmap + 5 CTAs (40 warps/SM) 1.22 ms. Membership / local index from registers (vertex mask + packed
// rank->position map, no lmap at all, 4.6 KB per warp) was slower at equal occupancy: 1.35 ms (4 CTAs), 1.24-1.26 (5).
#ifndef EMIT_NIBBLE
#define EMIT_NIBBLE 1  // NARROW path: lmap[] holds 4-bit positions (36-byte rows), staging trimmed so that 5 CTAs fit an SM
#endif
#define NIB_STRIDE 36  // 9 words: the 32 per-sample rows start in 32 different banks
#define NIB_STAGE 1232
#if EMIT_NIBBLE
template <>
struct __align__(16) SmallEmit<unsigned short> {  // 5664 bytes per warp: 5 CTAs of 8 warps per SM
    unsigned short stage[NIB_STAGE];
    u32 rowrange[64];
    unsigned char tile[32][8];
    unsigned char lmap[32][NIB_STRIDE];  // two vertices per byte: position (0..7) or 0xF = not a member
    unsigned short ent[2 * UGS_SMALL_MAX_COLS];
};
#endif
// payload = source column relative to the graph's first column (graph-contiguous batches), or the CSR position
// when the columns are read through `perm` (the source column is then gathered from csr_col at flush time).

// NARROW: k <= 8 and mode "sample" / encoder: staged entry = payload << 6 | l << 3 | j  (16 bits)
// else  :                            staged entry = payload << 15 | lane << 10 | l << 5 | j  (32 bits)
// Entries of the samples of tiles [t0, t1) of graph g, by one warp whose shared-memory slice is `s`.
// WIRE (host delivery, ugs_emit_edges_wire): the flush writes the transfer format of csrc/ugs_hostio.h directly instead
// of three int64 arrays -- one byte (j | l << 4) per entry and the source column as an 8-bit delta against bases[] of its
// 1024-entry block (k_wire_src_bases) -- 2 instead of 24 bytes per entry, and no int64 copy of the edges in HBM at all.
struct WireOut {
    unsigned char* ei;
    unsigned char* src;
    const i64* bases;
    unsigned long long* bad;  // entries whose delta does not fit 8 bits (the caller falls back to the int64 path)
};

template <bool NARROW, bool WIRE = false>
__device__ __forceinline__ void emit_graph_tiles(
    SmallEmit<typename std::conditional<NARROW, unsigned short, u32>::type>& s, i64 g, int t0, int t1,
    const i64* __restrict__ ptr, i64 node_base, i64 m, int k, int mode, int use_perm, const WS& w,
    const i64* __restrict__ nodes, const i64* __restrict__ edge_ptr, i64* __restrict__ out_u, i64* __restrict__ out_v,
    i64* __restrict__ out_src, const u64* __restrict__ packed, WireOut wo = WireOut()) {
    typedef typename std::conditional<NARROW, unsigned short, u32>::type stage_t;
    constexpr bool NIBBLE = NARROW && (EMIT_NIBBLE != 0);  // k <= 8: two vertices per lmap byte
    constexpr int STAGE_N = (int)(sizeof(SmallEmit<stage_t>::stage) / sizeof(stage_t));
    constexpr int LROW_BYTES = (int)sizeof(SmallEmit<stage_t>::lmap) / 32;
    const int lane = threadIdx.x & 31;
    if (t0 >= t1) return;
    const ugs_ginfo gi = w.ginfo[g];
    if (gi.cls != 1 || gi.nnz == 0) return;
    const i64 lo = ptr[g];
    const int n = (int)(ptr[g + 1] - lo);
    const i64 nb = lo - node_base;
    {   // nothing to do if this item's samples emit no entries
        const i64 sA = g * m + (i64)t0 * 32, sB = min(g * m + m, g * m + (i64)t1 * 32);
        if (edge_ptr[sA] == edge_ptr[sB]) return;
    }
    const i64 c0 = w.cstart[g];
    const i32* __restrict__ gcol = w.csr_col + 2 * c0;
    {
        const u32* rp = w.rowptr + nb + g;
        const i32* ci = w.csr_idx + 2 * c0;
        for (int v = lane; v < n; v += 32) s.rowrange[v] = (rp[v] << 1) | (rp[v + 1] << 17);
        for (int p = lane; p < gi.nnz; p += 32) {
            const u32 payload = use_perm ? (u32)p : (u32)(gcol[p] - (i32)c0);
            s.ent[p] = (unsigned short)((u32)ci[p] | (payload << 6));
        }
        u32* row = reinterpret_cast<u32*>(&s.lmap[lane][0]);
#pragma unroll
        for (int q = 0; q < LROW_BYTES / 4; ++q) row[q] = 0xFFFFFFFFu;
    }
    unsigned char* lrow = &s.lmap[lane][0];
    const unsigned char* entb = reinterpret_cast<const unsigned char*>(s.ent);
    const int r32 = 32 % k, q32 = 32 / k;

    for (int t = t0; t < t1; ++t) {
        const i64 s0 = g * m + (i64)t * 32;
        const int nv = (int)min((i64)32, m - (i64)t * 32);
        __syncwarp();
        const i64 my_lo = (lane < nv) ? edge_ptr[s0 + lane] : 0;
        const i64 tile_end = edge_ptr[s0 + nv];
        const i64 my_hi = __shfl_down_sync(FULL_MASK, my_lo, 1);
        const i64 my_end = (lane == nv - 1 || lane == 31) ? tile_end : my_hi;
        const i64 tile_lo = __shfl_sync(FULL_MASK, my_lo, 0);
        if (tile_end == tile_lo) continue;
        // NARROW: the sampling kernel left the row of every sample as eight local-id bytes (8 B instead of 8k B)
        const bool have_packed = NARROW && (packed != nullptr);
        const u64 prow = (have_packed && lane < nv) ? packed[s0 + lane] : ~0ull;
        if (!have_packed) {   // node rows -> shared tile (coalesced read)
            int sidx = lane / k, j = lane - sidx * k;
            const i64* in = nodes + s0 * k;
            for (int e = lane; e < nv * k; e += 32) {
                const i64 val = in[e];
                s.tile[sidx][j] = (val < 0) ? (unsigned char)0xFF : (unsigned char)(val - lo);
                j += r32; sidx += q32;
                if (j >= k) { j -= k; sidx++; }
            }
        }
        __syncwarp();
        // per-sample state of the owning lane: node bytes in registers + local-index map (= membership test).
        // Rows with a -1 have no entries (sampler.cpp:220-223), so only rows with entries are walked.
        u32 nb4[4] = {~0u, ~0u, ~0u, ~0u};
        const bool walk = (lane < nv) && (my_end > my_lo);
        if (walk) {
            if constexpr (NIBBLE) {
                if (have_packed) { nb4[0] = (u32)prow; nb4[1] = (u32)(prow >> 32); }
                else {
                    const uint2 nn = *reinterpret_cast<const uint2*>(&s.tile[lane][0]);
                    nb4[0] = nn.x; nb4[1] = nn.y;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < k) {
                        const u32 v = (nb4[j >> 2] >> (8 * (j & 3))) & 63u;
                        const u32 sh = (v & 1u) * 4u;
                        lrow[v >> 1] = (unsigned char)((lrow[v >> 1] & ~(0xFu << sh)) | ((u32)j << sh));
                    }
            } else {
                const uint4 nn = *reinterpret_cast<const uint4*>(&s.tile[lane][0]);
                nb4[0] = nn.x; nb4[1] = nn.y; nb4[2] = nn.z; nb4[3] = nn.w;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < k) lrow[(nb4[j >> 2] >> (8 * (j & 3))) & 0xFFu] = (unsigned char)j;
            }
        }
        // balanced sub-batches: when the tile does not fit the staging buffer, aim for equal shares
        const int tile_total = (int)min(tile_end - tile_lo, (i64)0x7fffffff);
        const int nsplit = (tile_total + STAGE_N - 1) / STAGE_N;
        const int cap_eff = (nsplit <= 1) ? STAGE_N : min(STAGE_N, tile_total / nsplit + 48);
        int sa = 0;
        while (sa < nv) {
            const i64 base = __shfl_sync(FULL_MASK, my_lo, sa);
            const u32 fits = __ballot_sync(FULL_MASK, lane >= sa && lane < nv && (my_end - base) <= cap_eff);
            const u32 nofit = ~fits & ~((1u << sa) - 1u);
            int sb = nofit ? min(nv, __ffs(nofit) - 1) : nv;
            if (sb == sa) sb = sa + 1;  // a single sample always fits EMIT_STAGE (<= 768 entries)
            const int run = (int)(((sb == nv) ? tile_end : __shfl_sync(FULL_MASK, my_lo, sb & 31)) - base);
            if (walk && lane >= sa && lane < sb) {
                stage_t* st = s.stage + (int)(my_lo - base);
                u32 pj = NARROW ? 0u : ((u32)lane << 10);  // | j
                int j = 0;
                u32 rr = s.rowrange[nb4[0] & 0xFFu];
                u32 p = rr & 0xFFFFu, pe = rr >> 16;  // byte offsets into ent[]
                // One loop, two guarded blocks per trip (switch row / take one entry): lanes re-converge every trip,
                // so a warp runs max_lane(entries of its sample) trips, not sum_j max_lane(row length).
#pragma nounroll
                while (true) {
                    if (p == pe) {
                        if (++j == k) break;
                        const u32 wsel = (j < 4) ? nb4[0] : (j < 8) ? nb4[1] : (j < 12) ? nb4[2] : nb4[3];
                        rr = s.rowrange[(wsel >> (8 * (j & 3))) & 0xFFu];
                        p = rr & 0xFFFFu; pe = rr >> 16;
                        pj++;
                    }
                    asm volatile("" ::: "memory");  // keep the two blocks in one loop (no loop nesting by the compiler)
                    if (p != pe) {
                        const u32 e = *reinterpret_cast<const unsigned short*>(entb + p);
                        if constexpr (NIBBLE) {
                            const u32 x = e & 63u;
                            const u32 l = ((u32)lrow[x >> 1] >> ((x & 1u) * 4u)) & 0xFu;
                            if (l != 0xFu) *st++ = (stage_t)((e & 0xFFC0u) + (l << 3) + pj);
                        } else {
                            const u32 l = lrow[e & 63u];
                            if (l != 0xFFu) {
                                if (NARROW) *st++ = (stage_t)((e & 0xFFC0u) + (l << 3) + pj);
                                else *st++ = (stage_t)(((e & 0xFFC0u) << 9) + (l << 5) + pj);
                            }
                        }
                        p += 2;
                    }
                }
            }
            __syncwarp();
            // ---- flush: entry e of the run goes to global position base + e ------------------------------
            if (NARROW && mode == UGS_MODE_ENCODER) {
                // batch-level ids (s*k + j, s*k + l): the owning sample of staged entry e is the last lane of the
                // sub-batch whose first entry is <= e -- a 5-step search over the lanes' start offsets
                const int rel_lo = (lane >= sa && lane < sb) ? (int)(my_lo - base) : 0x7fffffff;
                i64* pu = out_u + base + lane;
                i64* pv = out_v + base + lane;
                i64* ps = out_src + base + lane;
                const stage_t* sg = s.stage + lane;
                for (int e0 = 0; e0 < run; e0 += 32) {
                    const int e = e0 + lane;
                    int a = sa;
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        const int cand = a + step;
                        const int off = __shfl_sync(FULL_MASK, rel_lo, cand & 31);
                        if (cand < sb && off <= e) a = cand;
                    }
                    if (e < run) {
                        const u32 pk = sg[0];
                        const i64 ik = (s0 + a) * (i64)k;
                        pu[0] = ik + (i64)(pk & 7u);
                        pv[0] = ik + (i64)((pk >> 3) & 7u);
                        ps[0] = use_perm ? (i64)__ldg(gcol + (pk >> 6)) : c0 + (i64)(pk >> 6);
                    }
                    pu += 32; pv += 32; ps += 32; sg += 32;
                }
            } else if (NARROW && WIRE) {
                unsigned char* pe8 = wo.ei + base;
                unsigned char* ps8 = wo.src + base;
                int nbad = 0;
                for (int e = lane; e < run; e += 32) {  // 32 consecutive entries: one or two blocks, i.e. a broadcast load
                    const u32 pk = s.stage[e];
                    pe8[e] = (unsigned char)((pk & 7u) | (((pk >> 3) & 7u) << 4));
                    const i64 d = c0 - wo.bases[(base + e) >> 10] + (i64)(pk >> 6);
                    if (d < 0 || d >= 0xFF) nbad++;
                    ps8[e] = (unsigned char)d;
                }
                if (nbad) atomicAdd(wo.bad, (unsigned long long)nbad);
            } else if (NARROW) {
                i64* pu = out_u + base + lane;
                i64* pv = out_v + base + lane;
                i64* ps = out_src + base + lane;
                const stage_t* sg = s.stage + lane;
                int e = lane;
                if (!use_perm) {
                    for (; e + 96 < run; e += 128) {  // four full warp-rows per trip, no bounds checks
                        const u32 k0_ = sg[0], k1_ = sg[32], k2_ = sg[64], k3_ = sg[96];
                        pu[0] = (i64)(k0_ & 7u); pu[32] = (i64)(k1_ & 7u); pu[64] = (i64)(k2_ & 7u); pu[96] = (i64)(k3_ & 7u);
                        pv[0] = (i64)((k0_ >> 3) & 7u); pv[32] = (i64)((k1_ >> 3) & 7u);
                        pv[64] = (i64)((k2_ >> 3) & 7u); pv[96] = (i64)((k3_ >> 3) & 7u);
                        ps[0] = c0 + (i64)(k0_ >> 6); ps[32] = c0 + (i64)(k1_ >> 6);
                        ps[64] = c0 + (i64)(k2_ >> 6); ps[96] = c0 + (i64)(k3_ >> 6);
                        pu += 128; pv += 128; ps += 128; sg += 128;
                    }
                }
                for (; e < run; e += 32) {
                    const u32 pk = sg[0];
                    pu[0] = (i64)(pk & 7u);
                    pv[0] = (i64)((pk >> 3) & 7u);
                    ps[0] = use_perm ? (i64)__ldg(gcol + (pk >> 6)) : c0 + (i64)(pk >> 6);
                    pu += 32; pv += 32; ps += 32; sg += 32;
                }
            } else {
#pragma unroll 2
                for (int e = lane; e < run; e += 32) {
                    const u32 pk = s.stage[e];
                    const i64 o = base + e;
                    const int j = pk & 31, l = (pk >> 5) & 31, sl = (pk >> 10) & 31;
                    i64 av, bv;
                    if (mode == UGS_MODE_SAMPLE) { av = j; bv = l; }
                    else if (mode == UGS_MODE_GRAPH) {
                        const i64 ik = ((i64)t * 32 + sl) * k;
                        av = ik + j; bv = ik + l;
                    } else if (mode == UGS_MODE_ENCODER) {
                        const i64 ik = (s0 + sl) * k;
                        av = ik + j; bv = ik + l;
                    } else {
                        av = lo + s.tile[sl][j]; bv = lo + s.tile[sl][l];
                    }
                    out_u[o] = av;
                    out_v[o] = bv;
                    out_src[o] = use_perm ? (i64)__ldg(gcol + (pk >> 15)) : c0 + (i64)(pk >> 15);
                }
            }
            __syncwarp();
            sa = sb;
        }
        // un-set this tile's members so the map is all-0xFF again
        if (walk) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < k) lrow[((nb4[j >> 2] >> (8 * (j & 3))) & 0xFFu) >> (NIBBLE ? 1 : 0)] = (unsigned char)0xFF;
        }
    }
}

template <bool NARROW, bool WIRE = false>
__global__ void __launch_bounds__(EMIT_WARPS * 32, NARROW ? EMIT_MINB : 3)
k_emit_small(const i64* __restrict__ ptr, i64 G, i64 node_base, i64 m, int k, int mode, int tiles_per_item,
             i64 items_per_graph, int use_perm, WS w, const i64* __restrict__ nodes,
             const i64* __restrict__ edge_ptr, i64* __restrict__ out_u, i64* __restrict__ out_v,
             i64* __restrict__ out_src, const u64* __restrict__ packed, int flags, WireOut wo = WireOut()) {
    // fused path queued on the guess "no general-path graph": a wrong guess leaves garbage offsets, write nothing then
    if ((flags & UGS_EMIT_GUARD_GENERAL) && w.counters[UGS_CTR_N_GENERAL] != 0) return;
    if (flags & UGS_EMIT_COMPACT) {  // rows packed behind one another at the total the scan kernel left on the device
        const i64 t = (i64)w.counters[UGS_CTR_TOTAL_EDGES];
        out_v = out_u + t;
        out_src = out_u + 2 * t;
    }
    typedef typename std::conditional<NARROW, unsigned short, u32>::type stage_t;
    extern __shared__ __align__(16) unsigned char emit_smem_raw[];
    SmallEmit<stage_t>* sm_all = reinterpret_cast<SmallEmit<stage_t>*>(emit_smem_raw);
    const int warp = threadIdx.x >> 5;
    const i64 item = (i64)blockIdx.x * EMIT_WARPS + warp;
    const i64 g = (items_per_graph == 1) ? item : item / items_per_graph;
    if (g >= G) return;
    const int chunk = (int)(item - g * items_per_graph);
    const int ntiles = (int)((m + 31) >> 5);
    const int t0 = chunk * tiles_per_item, t1 = min(ntiles, t0 + tiles_per_item);
    emit_graph_tiles<NARROW, WIRE>(sm_all[warp], g, t0, t1, ptr, node_base, m, k, mode, use_perm, w, nodes, edge_ptr, out_u,
                                   out_v, out_src, packed, wo);
}

// bases[b] for the 8-bit source-column deltas of block b (entries [1024 b, 1024 b + 1024) of the output): the first
// column of the graph that owns the block's first entry. Every later entry of the block belongs to the same or a later
// graph, so its source column is >= that base; k_emit_small<true, true> counts the entries whose delta exceeds 8 bits.
__global__ void k_wire_src_bases(const i64* __restrict__ edge_ptr, i64 B, i64 m, i64 total, const i64* __restrict__ cstart,
                                 i64* __restrict__ bases) {
    const i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 pos = b << 10;  // UGS_DELTA8_BLOCK
    if (pos >= total) return;
    i64 lo = 0, hi = B;  // the last sample s with edge_ptr[s] <= pos (edge_ptr[0] = 0, edge_ptr[B] = total > pos)
    while (hi - lo > 1) {
        const i64 mid = (lo + hi) >> 1;
        if (edge_ptr[mid] <= pos) lo = mid; else hi = mid;
    }
    bases[b] = cstart[lo / m];
}

// m <= 16 (QM9-style runs): 32 / S graphs per warp, S = 8, 10 or 16 lanes each, like k_sample_small_packed. Every lane group
// stages the CSR of its own graph (at most PK_ENT entries; payload = source column relative to the FIRST graph of the
// warp, 10 bits); the output of the warp's graphs is one contiguous run, so staging and the coalesced flush work as in
// the one-graph kernel. k <= 8, mode "sample" / encoder, graph-contiguous columns, rows packed by the sampling kernel.
// A warp holding a general-path graph or a graph with more than PK_ENT entries handles its graphs one after the other.
#define PK_ENT 256
template <int GPW>
struct __align__(16) PackedEmit {
    unsigned short stage[NIB_STAGE];
    unsigned char lmap[32][NIB_STRIDE];
    u32 rowrange[GPW][64];
    unsigned short ent[GPW][PK_ENT];
};

template <int S>
__global__ void __launch_bounds__(EMIT_WARPS * 32, 4)
k_emit_small_packed(const i64* __restrict__ ptr, i64 G, i64 node_base, i64 m, int k, int mode, WS w,
                    const i64* __restrict__ nodes, const i64* __restrict__ edge_ptr, i64* __restrict__ out_u,
                    i64* __restrict__ out_v, i64* __restrict__ out_src, const u64* __restrict__ packed, int flags) {
    if ((flags & UGS_EMIT_GUARD_GENERAL) && w.counters[UGS_CTR_N_GENERAL] != 0) return;  // see k_emit_small
    if (flags & UGS_EMIT_COMPACT) {
        const i64 t = (i64)w.counters[UGS_CTR_TOTAL_EDGES];
        out_v = out_u + t;
        out_src = out_u + 2 * t;
    }
    constexpr int GPW = 32 / S;
    constexpr size_t SLICE = sizeof(PackedEmit<GPW>) > sizeof(SmallEmit<unsigned short>) ? sizeof(PackedEmit<GPW>)
                                                                                            : sizeof(SmallEmit<unsigned short>);
    extern __shared__ __align__(16) unsigned char emit_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* slice = emit_smem_raw + (size_t)warp * SLICE;
    const int slot = lane / S, i = lane - slot * S;
    const i64 g_first = ((i64)blockIdx.x * EMIT_WARPS + warp) * GPW;
    if (g_first >= G) return;
    const i64 g_end = min(G, g_first + GPW);
    const i64 g = g_first + slot;
    const bool valid = slot < GPW && g < G;  // S = 10: lanes 30 and 31 belong to no graph
    const int vslot = valid ? slot : 0;
    ugs_ginfo gi;
    gi.cls = -1; gi.nnz = 0; gi.mult = 0; gi.level = 0; gi.Z = 0.0; gi.n_eligible = 0; gi.last_viable = 0;
    if (valid) gi = w.ginfo[g];
    const i64 tile_lo = edge_ptr[g_first * m], tile_end = edge_ptr[g_end * m];
    if (tile_end == tile_lo) return;
    const i64 c0 = w.cstart[g_first];
    if (__any_sync(FULL_MASK, valid && (gi.cls == 2 || (gi.cls == 1 && gi.nnz > PK_ENT))) || w.cstart[g_end] - c0 > 1023) {
        for (i64 gg = g_first; gg < g_end; ++gg) {
            emit_graph_tiles<true>(*reinterpret_cast<SmallEmit<unsigned short>*>(slice), gg, 0, 1, ptr, node_base, m, k, mode,
                                   0, w, nodes, edge_ptr, out_u, out_v, out_src, packed);
            __syncwarp();
        }
        return;
    }
    PackedEmit<GPW>& s = *reinterpret_cast<PackedEmit<GPW>*>(slice);
    const bool live = valid && gi.cls == 1 && gi.nnz > 0;
    {
        const i64 lo = valid ? ptr[g] : 0;
        const int n = live ? (int)(ptr[g + 1] - lo) : 0;
        const i64 cg = valid ? w.cstart[g] : 0;
        const u32* rp = w.rowptr + (lo - node_base) + g;
        const i32* ci = w.csr_idx + 2 * cg;
        const i32* gc = w.csr_col + 2 * cg;
        for (int v = i; v < n; v += S) s.rowrange[vslot][v] = (rp[v] << 1) | (rp[v + 1] << 17);
        const int nnz = live ? gi.nnz : 0;
        for (int p = i; p < nnz; p += S) s.ent[vslot][p] = (unsigned short)((u32)ci[p] | ((u32)(gc[p] - (i32)c0) << 6));
        u32* row = reinterpret_cast<u32*>(&s.lmap[lane][0]);
#pragma unroll
        for (int q = 0; q < NIB_STRIDE / 4; ++q) row[q] = 0xFFFFFFFFu;
    }
    __syncwarp();
    // this lane's sample: lanes without one (i >= m, graph beyond the batch) get the empty range that keeps the per-lane
    // start offsets monotone (= the start of the next graph's first sample)
    const bool has = valid && (i64)i < m;
    const i64 rowid = has ? g * m + i : -1;
    const i64 my_lo = has ? edge_ptr[rowid] : (valid ? edge_ptr[(g + 1) * m] : tile_end);
    const i64 my_end = has ? edge_ptr[rowid + 1] : my_lo;
    const bool walk = has && live && (my_end > my_lo);
    unsigned char* lrow = &s.lmap[lane][0];
    const unsigned char* entb = reinterpret_cast<const unsigned char*>(s.ent[vslot]);
    const u32* rr_tab = s.rowrange[vslot];
    u32 nb2[2] = {~0u, ~0u};
    if (walk) {
        const u64 prow = packed[rowid];
        nb2[0] = (u32)prow; nb2[1] = (u32)(prow >> 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < k) {
                const u32 v = (nb2[j >> 2] >> (8 * (j & 3))) & 63u;
                const u32 sh = (v & 1u) * 4u;
                lrow[v >> 1] = (unsigned char)((lrow[v >> 1] & ~(0xFu << sh)) | ((u32)j << sh));
            }
    }
    const int tile_total = (int)min(tile_end - tile_lo, (i64)0x7fffffff);
    const int nsplit = (tile_total + NIB_STAGE - 1) / NIB_STAGE;
    const int cap_eff = (nsplit <= 1) ? NIB_STAGE : min(NIB_STAGE, tile_total / nsplit + 48);
    int sa = 0;
    while (sa < 32) {
        const i64 base = __shfl_sync(FULL_MASK, my_lo, sa);
        const u32 fits = __ballot_sync(FULL_MASK, lane >= sa && (my_end - base) <= cap_eff);
        const u32 nofit = ~fits & ~((1u << sa) - 1u);
        int sb = nofit ? (__ffs(nofit) - 1) : 32;
        if (sb == sa) sb = sa + 1;  // a single sample always fits the staging buffer (<= PK_ENT entries)
        const int run = (int)(((sb == 32) ? tile_end : __shfl_sync(FULL_MASK, my_lo, sb & 31)) - base);
        if (walk && lane >= sa && lane < sb) {
            unsigned short* st = s.stage + (int)(my_lo - base);
            u32 pj = 0u;
            int j = 0;
            u32 rr = rr_tab[nb2[0] & 0xFFu];
            u32 p = rr & 0xFFFFu, pe = rr >> 16;
#pragma nounroll
            while (true) {
                if (p == pe) {
                    if (++j == k) break;
                    const u32 wsel = (j < 4) ? nb2[0] : nb2[1];
                    rr = rr_tab[(wsel >> (8 * (j & 3))) & 0xFFu];
                    p = rr & 0xFFFFu; pe = rr >> 16;
                    pj++;
                }
                asm volatile("" ::: "memory");
                if (p != pe) {
                    const u32 e = *reinterpret_cast<const unsigned short*>(entb + p);
                    const u32 x = e & 63u;
                    const u32 l = ((u32)lrow[x >> 1] >> ((x & 1u) * 4u)) & 0xFu;
                    if (l != 0xFu) *st++ = (unsigned short)((e & 0xFFC0u) + (l << 3) + pj);
                    p += 2;
                }
            }
        }
        __syncwarp();
        if (mode == UGS_MODE_ENCODER) {
            const int rel_lo = (lane >= sa && lane < sb) ? (int)(my_lo - base) : 0x7fffffff;
            for (int e0 = 0; e0 < run; e0 += 32) {
                const int e = e0 + lane;
                int a = sa;
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const int cand = a + step;
                    const int off = __shfl_sync(FULL_MASK, rel_lo, cand & 31);
                    if (cand < sb && off <= e) a = cand;
                }
                if (e < run) {
                    const u32 pk = s.stage[e];
                    const i64 ik = ((g_first + a / S) * m + (a % S)) * (i64)k;  // the owning sample's row
                    out_u[base + e] = ik + (i64)(pk & 7u);
                    out_v[base + e] = ik + (i64)((pk >> 3) & 7u);
                    out_src[base + e] = c0 + (i64)(pk >> 6);
                }
            }
        } else {
            for (int e = lane; e < run; e += 32) {
                const u32 pk = s.stage[e];
                out_u[base + e] = (i64)(pk & 7u);
                out_v[base + e] = (i64)((pk >> 3) & 7u);
                out_src[base + e] = c0 + (i64)(pk >> 6);
            }
        }
        __syncwarp();
        sa = sb;
    }
}

// =====================================================================================================
// General path: emission, one warp per sample (also the generic counter for caller-provided rows)
// =====================================================================================================
// Rows are walked 32 entries at a time. Membership and local index of an entry come from a 1024-slot table in shared
// memory (slot y & 1023 holds the position of member y inside the sample, 0xFF = empty) followed by ONE shuffle that
// fetches the candidate member for an exact comparison: a chunk without a possible member costs a load, a table
// look-up and a vote; a hub row of thousands of entries holds a handful of members. Two members that share a slot
// (rare: ~k^2/2048 of the samples) make the sample fall back to comparing against all k members.
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(256)
k_emit_general(const i64* __restrict__ ptr, i64 node_base, i64 m, int k, int mode, i64 n_items, bool use_glist,
               WS w, const i64* __restrict__ nodes, i64* __restrict__ edge_ptr, i64 total,
               i64* __restrict__ out_u, i64* __restrict__ out_v, i64* __restrict__ out_src, int flags) {
    __shared__ __align__(16) unsigned char pos_all[8][1024];
    if (n_items < 0) n_items = (i64)w.counters[UGS_CTR_N_GENERAL];  // fused small-batch path
    if (n_items == 0) return;
    if (!COUNT_ONLY && (flags & UGS_EMIT_COMPACT)) {  // see k_emit_small
        total = (i64)w.counters[UGS_CTR_TOTAL_EDGES];
        out_v = out_u + total;
        out_src = out_u + 2 * total;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    unsigned char* pos = pos_all[warp];
    {
        uint4* p4 = reinterpret_cast<uint4*>(pos);
        p4[lane] = make_uint4(~0u, ~0u, ~0u, ~0u);
        p4[lane + 32] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    __syncwarp();
    const u32 lt_mask = (1u << lane) - 1u;
    const i64 nwarps = (i64)gridDim.x * wpb;
    const i64 total_items = n_items * m;
    for (i64 t = (i64)blockIdx.x * wpb + warp; t < total_items; t += nwarps) {
        const i64 gi_ = t / m;
        const i64 i = t - gi_ * m;
        const i64 g = use_glist ? (i64)w.glist[gi_] : gi_;
        const i64 sidx = g * m + i;
        const ugs_ginfo gi = w.ginfo[g];
        if (COUNT_ONLY && gi.cls == 0) { if (lane == 0) edge_ptr[1 + sidx] = 0; continue; }
        i64 base = 0;
        if (!COUNT_ONLY) {
            base = edge_ptr[sidx];
            if (edge_ptr[sidx + 1] == base) continue;
        }
        const i64 lo = ptr[g];
        const i64 nb = lo - node_base;
        const u32* __restrict__ rp = w.rowptr + nb + g;
        const i32* __restrict__ cidx = w.csr_idx + 2 * w.cstart[g];
        const i32* __restrict__ ccol = w.csr_col + 2 * w.cstart[g];
        const i64 nv = (lane < k) ? nodes[sidx * k + lane] : 0;
        const bool complete = !__any_sync(FULL_MASK, (lane < k) && nv < 0);
        const int mine = (lane < k && nv >= 0) ? (int)(nv - lo) : -1;
        i64 cnt = 0;
        if (complete) {
            // members sharing a table slot (or equal members in caller-provided rows): compare against all k instead
            const u32 peers = __match_any_sync(FULL_MASK, (lane < k) ? (mine & 1023) : (0x10000 + lane));
            const bool clash = __any_sync(FULL_MASK, (lane < k) && (peers != (1u << lane)));
            if (lane < k && !clash) pos[mine & 1023] = (unsigned char)lane;
            const u32 my_p0 = (lane < k) ? rp[mine] : 0u, my_p1 = (lane < k) ? rp[mine + 1] : 0u;
            __syncwarp();
            for (int j = 0; j < k; ++j) {
                const int u = __shfl_sync(FULL_MASK, mine, j);
                const u32 p0 = __shfl_sync(FULL_MASK, my_p0, j), p1 = __shfl_sync(FULL_MASK, my_p1, j);
                for (u32 pb = p0; pb < p1; pb += 32) {
                    const u32 p = pb + lane;
                    int y = -2, l = -1;
                    if (p < p1) y = cidx[p];
                    if (!clash) {
                        const int c = (p < p1) ? (int)pos[y & 1023] : 0xFF;
                        if (!__any_sync(FULL_MASK, c != 0xFF)) continue;
                        const int cand = __shfl_sync(FULL_MASK, mine, c & 31);
                        if (c != 0xFF && cand == y) l = c;
                    } else {
                        for (int q = 0; q < k; ++q) if (y == __shfl_sync(FULL_MASK, mine, q)) l = q;
                    }
                    const bool hit = (l >= 0) && (p < p1);
                    const u32 ball = __ballot_sync(FULL_MASK, hit);
                    if (!COUNT_ONLY && hit) {
                        const i64 o = base + cnt + __popc(ball & lt_mask);
                        i64 a, b;
                        if (mode == UGS_MODE_SAMPLE) { a = j; b = l; }
                        else if (mode == UGS_MODE_GRAPH) { a = i * k + j; b = i * k + l; }
                        else if (mode == UGS_MODE_ENCODER) { a = sidx * k + j; b = sidx * k + l; }
                        else { a = lo + u; b = lo + y; }
                        if (o < total) { out_u[o] = a; out_v[o] = b; out_src[o] = (i64)ccol[p]; }
                    }
                    cnt += __popc(ball);
                }
            }
            __syncwarp();
            if (lane < k && !clash) pos[mine & 1023] = (unsigned char)0xFF;
            __syncwarp();
        }
        if (COUNT_ONLY && lane == 0) edge_ptr[1 + sidx] = cnt;
    }
}

// =====================================================================================================
// Host launchers
// =====================================================================================================
int ugs_launch_sample_uniform(const i64* ptr, i64 G, i64 node_base, i64 m, int k, u64 seed, u64 gbase, i64 n_general,
                              WS& w, i64* nodes, i64* edge_ptr, i64* sample_ptr, cudaStream_t st);  // ugs_variants.cu

static int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

// scratch = [scan status words | packed node rows: one u64 (eight local ids, 0xFF = -1) per sample, k <= 8]
static size_t scan_scratch_bytes(int64_t B) {
    const int64_t tiles = ((B > 0 ? B : 0) + SCAN_TILE - 1) / SCAN_TILE + 1;
    return (size_t)(tiles + 1) * sizeof(unsigned long long);
}
static size_t packed_rows_offset(int64_t B) { return (scan_scratch_bytes(B) + 255) & ~(size_t)255; }

extern "C" size_t ugs_sample_scratch_bytes(int64_t G, int64_t m) {
    const int64_t B = (G > 0 && m > 0) ? G * m : 0;
    return packed_rows_offset(B) + (size_t)B * sizeof(u64);
}

static int run_scan(i64* edge_ptr, i64 B, void* scratch, size_t scratch_bytes, WS& w, cudaStream_t st,
                    i64* snap = nullptr, i64 snap_seq = 0) {
    const size_t need = scan_scratch_bytes(B);
    if (scratch_bytes < need || !scratch) return ugs_set_error(UGS_EINVAL, "scratch too small: %zu < %zu", scratch_bytes, need);
    UGS_CUDA_CHECK(cudaMemsetAsync(scratch, 0, need, st));
    if (B > 0) {
        const i64 tiles = (B + SCAN_TILE - 1) / SCAN_TILE;
        k_scan_counts<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(edge_ptr + 1, B, (unsigned long long*)scratch, w.counters, snap,
                                                                snap_seq);
        UGS_LAUNCH_CHECK("k_scan_counts");
    }
    return UGS_OK;
}

int ugs_run_scan(i64* edge_ptr, i64 B, void* scratch, size_t scratch_bytes, WS& w, cudaStream_t st) {
    return run_scan(edge_ptr, B, scratch, scratch_bytes, w, st);
}

// A warp handles up to 4 tiles of 32 samples of one graph (the graph's tables are staged once per warp). Small batches
// do not have enough (graph, 4-tile) items to fill the GPU -- MUTAG-shaped G=188, m=64: 188 warps on 148 SMs -- so
// they get fewer tiles per warp until there are at least two 8-warp CTAs per SM.
// lanes per graph of the packed kernels (m <= 16): 4, 3 or 2 graphs per warp
static inline int packed_lanes(int64_t m) { return m <= 8 ? 8 : (m <= 10 ? 10 : 16); }

static inline void small_items(int64_t m, int64_t G, int* tiles_per_item, int64_t* items_per_graph) {
    const int64_t ntiles = (m + 31) / 32;
    int tpi = (int)(ntiles < 4 ? (ntiles < 1 ? 1 : ntiles) : 4);
    const int64_t want = (int64_t)sm_count() * 16;
    while (tpi > 1 && G * ((ntiles + tpi - 1) / tpi) < want) tpi >>= 1;
    *tiles_per_item = tpi;
    *items_per_graph = (ntiles + tpi - 1) / tpi;
    if (*items_per_graph < 1) *items_per_graph = 1;
}

extern "C" int ugs_sample_batch(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, uint64_t seed,
                                uint64_t graph_index_base, uint64_t sample_index_base, int seed_mode, int variant,
                                int64_t n_general, int64_t max_n_general, void* workspace, const ugs_layout* L, void* scratch,
                                size_t scratch_bytes, int64_t* nodes, int64_t* edge_ptr, int64_t* sample_ptr,
                                void* stream_) {
    return ugs_sample_batch_early(ptr, G, node_base, m, k, seed, graph_index_base, sample_index_base, seed_mode, variant,
                                  n_general, max_n_general, workspace, L, scratch, scratch_bytes, nodes, edge_ptr,
                                  sample_ptr, nullptr, stream_);
}

extern "C" int ugs_sample_batch_early(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, uint64_t seed,
                                      uint64_t graph_index_base, uint64_t sample_index_base, int seed_mode, int variant,
                                      int64_t n_general, int64_t max_n_general, void* workspace, const ugs_layout* L,
                                      void* scratch, size_t scratch_bytes, int64_t* nodes, int64_t* edge_ptr,
                                      int64_t* sample_ptr, int64_t* h_info, void* stream_) {
    return ugs_sample_batch_impl(ptr, G, node_base, m, k, seed, graph_index_base, sample_index_base, seed_mode, variant,
                                 n_general, max_n_general, workspace, L, scratch, scratch_bytes, nodes, edge_ptr,
                                 sample_ptr, h_info, stream_, false);
}

// counters_fresh: ugs_segment_batch has just zeroed all of counters[] on this stream and nothing has counted since
// (the fused small-batch path): the two counter memsets are skipped.
int ugs_sample_batch_impl(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, uint64_t seed,
                          uint64_t graph_index_base, uint64_t sample_index_base, int seed_mode, int variant,
                          int64_t n_general, int64_t max_n_general, void* workspace, const ugs_layout* L,
                          void* scratch, size_t scratch_bytes, int64_t* nodes, int64_t* edge_ptr,
                          int64_t* sample_ptr, int64_t* h_info, void* stream_, bool counters_fresh, int64_t* snap,
                          int64_t snap_seq) {
    if (!workspace || !L || !edge_ptr || !sample_ptr) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (h_info && variant != UGS_VARIANT_REF) return ugs_set_error(UGS_EINVAL, "the early read-back needs UGS_VARIANT_REF");
    if (k < 1 || k > UGS_MAX_K) return ugs_set_error(UGS_ELIMIT, "k must be in [1, %d]", UGS_MAX_K);
    if (m < 0) return ugs_set_error(UGS_EINVAL, "m_per_graph must be >= 0");
    if (variant != UGS_VARIANT_REF && variant != UGS_VARIANT_UNIFORM) return ugs_set_error(UGS_EINVAL, "unknown variant %d", variant);
    if (m >= (1ll << 32) || sample_index_base + (uint64_t)m > (1ull << 32))
        return ugs_set_error(UGS_ELIMIT, "sample_index_base + m_per_graph must be <= 2^32");
    if (sample_index_base && variant != UGS_VARIANT_REF)
        return ugs_set_error(UGS_EINVAL, "sample_index_base is only supported by UGS_VARIANT_REF");
    if (seed_mode != 0 && seed_mode != 1) return ugs_set_error(UGS_EINVAL, "seed_mode must be 0 or 1");
    if (seed_mode && variant != UGS_VARIANT_REF)
        return ugs_set_error(UGS_EINVAL, "seed_mode 1 is only supported by UGS_VARIANT_REF");
    const u32 ibase = (u32)sample_index_base;
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(workspace, L);
    if (!counters_fresh) {
        UGS_CUDA_CHECK(cudaMemsetAsync(&w.counters[UGS_CTR_TOTAL_EDGES], 0, sizeof(i64) * 4, st));
        UGS_CUDA_CHECK(cudaMemsetAsync(&w.counters[UGS_CTR_EARLY_TOTAL], 0, sizeof(i64) * UGS_CTR_EARLY_SLOTS, st));
    }
    if (G <= 0 || m == 0) {
        if (h_info) {  // nothing is sampled: the snapshot is all the caller needs
            const int rc = ugs_batch_info_begin(workspace, L, h_info, stream_);
            if (rc != UGS_OK) return rc;
        }
        // sample_ptr = arange * m; edge_ptr = [0]
        UGS_CUDA_CHECK(cudaMemsetAsync(edge_ptr, 0, sizeof(i64), st));
        UGS_CUDA_CHECK(cudaMemsetAsync(sample_ptr, 0, sizeof(i64) * (size_t)((G > 0 ? G : 0) + 1), st));
        UGS_CUDA_CHECK(cudaMemsetAsync(&w.counters[UGS_CTR_TOTAL_EDGES], 0, sizeof(i64), st));
        return UGS_OK;
    }
    if (!nodes) return ugs_set_error(UGS_EINVAL, "nodes is NULL");
    if (variant == UGS_VARIANT_UNIFORM) {
        const int rc = ugs_launch_sample_uniform(ptr, G, node_base, m, k, seed, graph_index_base, n_general, w, nodes,
                                                 edge_ptr, sample_ptr, st);
        if (rc != UGS_OK) return rc;
        return run_scan(edge_ptr, G * m, scratch, scratch_bytes, w, st);
    }
    int tpi; int64_t ipg;
    small_items(m, G, &tpi, &ipg);
    // packed node rows for the emission pass (k <= 8), kept behind the scan words of the caller's scratch buffer
    u64* packed = nullptr;
    if (k <= 8 && scratch && scratch_bytes >= ugs_sample_scratch_bytes(G, m))
        packed = reinterpret_cast<u64*>((char*)scratch + packed_rows_offset(G * m));
    if (m <= 16) {
        // several graphs per warp: 8 lanes per graph up to m = 8, 10 lanes up to m = 10 (three graphs), 16 lanes up to m = 16
        const int S = packed_lanes(m);
        const int64_t warps = (G + 32 / S - 1) / (32 / S);
        const unsigned grid = (unsigned)((warps + SAMP_WARPS - 1) / SAMP_WARPS);
        const size_t slice = (sizeof(SlotSamp) * (32 / S) > sizeof(SmallSamp)) ? sizeof(SlotSamp) * (32 / S) : sizeof(SmallSamp);
        const size_t psmem = ((slice + 15) & ~(size_t)15) * SAMP_WARPS + 2048;  // + the select table
#define UGS_LAUNCH_PACKED(KT_)                                                                                          \
    do {                                                                                                                \
        if (S == 8) {                                                                                                   \
            UGS_CUDA_CHECK(cudaFuncSetAttribute(k_sample_small_packed<KT_, 8>,                                          \
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));             \
            k_sample_small_packed<KT_, 8><<<grid, SAMP_WARPS * 32, psmem, st>>>(ptr, G, node_base, m, k, seed,          \
                                                                                 graph_index_base, ibase, seed_mode, w, \
                                                                                 nodes, edge_ptr, sample_ptr, packed); \
        } else if (S == 10) {                                                                                           \
            UGS_CUDA_CHECK(cudaFuncSetAttribute(k_sample_small_packed<KT_, 10>,                                         \
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));             \
            k_sample_small_packed<KT_, 10><<<grid, SAMP_WARPS * 32, psmem, st>>>(ptr, G, node_base, m, k, seed,         \
                                                                                  graph_index_base, ibase, seed_mode, w, \
                                                                                  nodes, edge_ptr, sample_ptr, packed); \
        } else {                                                                                                        \
            UGS_CUDA_CHECK(cudaFuncSetAttribute(k_sample_small_packed<KT_, 16>,                                         \
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));             \
            k_sample_small_packed<KT_, 16><<<grid, SAMP_WARPS * 32, psmem, st>>>(ptr, G, node_base, m, k, seed,         \
                                                                                  graph_index_base, ibase, seed_mode, w, \
                                                                                  nodes, edge_ptr, sample_ptr, packed); \
        }                                                                                                               \
    } while (0)
        switch (k) {
            case 3: UGS_LAUNCH_PACKED(3); break;
            case 4: UGS_LAUNCH_PACKED(4); break;
            case 5: UGS_LAUNCH_PACKED(5); break;
            case 6: UGS_LAUNCH_PACKED(6); break;
            case 8: UGS_LAUNCH_PACKED(8); break;
            default: UGS_LAUNCH_PACKED(0); break;
        }
#undef UGS_LAUNCH_PACKED
        UGS_LAUNCH_CHECK("k_sample_small_packed");
    } else if (SAMP_PIPE && ipg == 1 && G < (1ll << 31) && G >= 2ll * sm_count() * SAMP_MINB * SAMP_WARPS) {
        // persistent warps with cp.async fetch-ahead: at least two graphs per resident warp
        const unsigned grid = (unsigned)(sm_count() * SAMP_MINB);
#define UGS_LAUNCH_PIPE(KT_)                                                                                          \
    k_sample_small_pipe<KT_><<<grid, SAMP_WARPS * 32, 0, st>>>(ptr, (int)G, node_base, m, k, seed, graph_index_base,   \
                                                                 ibase, seed_mode, w, nodes, edge_ptr, sample_ptr, packed)
        switch (k) {
            case 3: UGS_LAUNCH_PIPE(3); break;
            case 4: UGS_LAUNCH_PIPE(4); break;
            case 5: UGS_LAUNCH_PIPE(5); break;
            case 6: UGS_LAUNCH_PIPE(6); break;
            case 8: UGS_LAUNCH_PIPE(8); break;
            default: UGS_LAUNCH_PIPE(0); break;
        }
#undef UGS_LAUNCH_PIPE
        UGS_LAUNCH_CHECK("k_sample_small_pipe");
    } else {
        const int64_t items = G * ipg;
        const unsigned grid = (unsigned)((items + SAMP_WARPS - 1) / SAMP_WARPS);
#define UGS_LAUNCH_SAMPLE(KT_)                                                                                   \
    k_sample_small<KT_><<<grid, SAMP_WARPS * 32, 0, st>>>(ptr, G, node_base, m, k, seed, graph_index_base, ibase, \
                                                            seed_mode, tpi, ipg, w, nodes, edge_ptr, sample_ptr, packed)
        switch (k) {
            case 3: UGS_LAUNCH_SAMPLE(3); break;
            case 4: UGS_LAUNCH_SAMPLE(4); break;
            case 5: UGS_LAUNCH_SAMPLE(5); break;
            case 6: UGS_LAUNCH_SAMPLE(6); break;
            case 8: UGS_LAUNCH_SAMPLE(8); break;
            default: UGS_LAUNCH_SAMPLE(0); break;
        }
#undef UGS_LAUNCH_SAMPLE
        UGS_LAUNCH_CHECK("k_sample_small");
    }
    if (n_general != 0) {  // < 0: count on the device, max_n_general = the caller's bound (total nodes of the batch)
        // per-warp slice: cut bitmap padded to whole blocks of 32 words + one count per block (16-byte aligned)
        const int bm_words = (int)(((max_n_general + 1023) / 1024) * 32);
        const int bm_stride = bm_words + ((bm_words / 32 + 3) & ~3);
        const size_t smem_cap = 200 * 1024;
        int wpb = 8;
        while (wpb > 1 && (size_t)wpb * bm_stride * 4 > smem_cap) wpb >>= 1;
        const size_t smem = (size_t)wpb * bm_stride * 4;
        if (smem > smem_cap)
            return ugs_set_error(UGS_ELIMIT, "graph with %lld nodes exceeds the per-warp cut bitmap (max %zu nodes)",
                                 (long long)max_n_general, smem_cap * 8);
        UGS_CUDA_CHECK(cudaFuncSetAttribute(k_sample_general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        UGS_CUDA_CHECK(cudaFuncSetAttribute(k_sample_general, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            cudaSharedmemCarveoutMaxShared));
        UGS_CUDA_CHECK(cudaMemsetAsync(&w.counters[UGS_CTR_WORK], 0, sizeof(i64), st));
        const int64_t ngb = n_general < 0 ? G : n_general;  // (bound on) the number of general-path graphs
        // PROTEINS-sized graphs of a batch with many general-path samples: tables in shared memory, one thread per sample
        if (ngb * m >= MEDS_MIN_SAMPLES) {
            const int mthreads = (int)(m >= MEDS_THREADS ? MEDS_THREADS : ((m + 31) / 32) * 32);
            const int64_t cpg = (m + mthreads - 1) / mthreads;
            const size_t msmem = sizeof(MedSample) + (size_t)64 * mthreads * sizeof(u32);
            UGS_CUDA_CHECK(cudaFuncSetAttribute(k_sample_medium, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
            int64_t mblocks = ngb * cpg;
            const int64_t mcap = (int64_t)sm_count() * 16;
            if (mblocks > mcap) mblocks = mcap;
            k_sample_medium<<<(unsigned)mblocks, mthreads, msmem, st>>>(ptr, G, node_base, m, k, seed, graph_index_base, ibase,
                                                                       seed_mode, n_general, cpg, w, nodes, edge_ptr);
            UGS_LAUNCH_CHECK("k_sample_medium");
        }
        const int64_t total = ngb * m;
        int64_t blocks = (total + wpb - 1) / wpb;
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        k_sample_general<<<(unsigned)blocks, wpb * 32, smem, st>>>(ptr, node_base, m, k, seed, graph_index_base, ibase,
                                                                  seed_mode, n_general, bm_words, bm_stride, w, nodes,
                                                                  edge_ptr);
        UGS_LAUNCH_CHECK("k_sample_general");
    }
    if (h_info) {  // counters as of the sampling kernels travel to the host while the scan runs
        const int rc = ugs_batch_info_begin(workspace, L, h_info, stream_);
        if (rc != UGS_OK) return rc;
    }
    return run_scan(edge_ptr, G * m, scratch, scratch_bytes, w, st, snap, snap_seq);
}

extern "C" int ugs_count_edges(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int64_t n_general,
                               void* workspace, const ugs_layout* L, void* scratch, size_t scratch_bytes,
                               const int64_t* nodes, int64_t* edge_ptr, void* stream_) {
    if (!workspace || !L || !edge_ptr) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (k < 1 || k > UGS_MAX_K) return ugs_set_error(UGS_ELIMIT, "k must be in [1, %d]", UGS_MAX_K);
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(workspace, L);
    UGS_CUDA_CHECK(cudaMemsetAsync(edge_ptr, 0, sizeof(i64), st));
    if (G <= 0 || m <= 0) {
        UGS_CUDA_CHECK(cudaMemsetAsync(&w.counters[UGS_CTR_TOTAL_EDGES], 0, sizeof(i64), st));
        return UGS_OK;
    }
    int64_t blocks = (G * m + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_emit_general<true><<<(unsigned)blocks, 256, 0, st>>>(ptr, node_base, m, k, 0, G, false, w, nodes, edge_ptr, 0,
                                                          nullptr, nullptr, nullptr, 0);
    UGS_LAUNCH_CHECK("k_count_general");
    (void)n_general;
    return run_scan(edge_ptr, G * m, scratch, scratch_bytes, w, st);
}

extern "C" int ugs_emit_edges(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int mode,
                              int use_perm, int64_t n_general, const void* workspace, const ugs_layout* L, const int64_t* nodes,
                              const int64_t* edge_ptr, int64_t total, int64_t* edge_index_out, int64_t* edge_src_out,
                              void* stream_) {
    return ugs_emit_edges_packed(ptr, G, node_base, m, k, mode, use_perm, n_general, workspace, L, nodes, edge_ptr, total,
                                 edge_index_out, edge_src_out, nullptr, 0, stream_);
}

extern "C" int ugs_emit_edges_packed(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int mode,
                                     int use_perm, int64_t n_general, const void* workspace, const ugs_layout* L,
                                     const int64_t* nodes, const int64_t* edge_ptr, int64_t total,
                                     int64_t* edge_index_out, int64_t* edge_src_out, const void* sample_scratch,
                                     size_t sample_scratch_bytes, void* stream_) {
    return ugs_emit_edges_impl(ptr, G, node_base, m, k, mode, use_perm, n_general, workspace, L, nodes, edge_ptr, total,
                               edge_index_out, edge_src_out, sample_scratch, sample_scratch_bytes, 0, stream_);
}

extern "C" int ugs_emit_edges_wire(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int64_t n_general,
                                   const void* workspace, const ugs_layout* L, const int64_t* nodes,
                                   const int64_t* edge_ptr, int64_t total, void* d_wire, size_t wire_bytes,
                                   const void* sample_scratch, size_t sample_scratch_bytes, void* stream_) {
    if (!workspace || !L || !edge_ptr || !nodes || !d_wire) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (k < 1 || k > 8) return ugs_set_error(UGS_ELIMIT, "the wire emission needs k <= 8");
    if (n_general != 0) return ugs_set_error(UGS_ELIMIT, "the wire emission covers bitmask-path graphs only (n_general = %lld)",
                                             (long long)n_general);
    if (G <= 0 || m <= 0 || total <= 0) return UGS_OK;
    const UgsWire wl = ugs_wire_layout_emit((size_t)(G * m) * (size_t)k, (size_t)total);
    if (wire_bytes < wl.total) return ugs_set_error(UGS_EINVAL, "wire buffer too small");
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(const_cast<void*>(workspace), L);
    char* base = (char*)d_wire;
    WireOut wo;
    wo.ei = (unsigned char*)(base + wl.ei);
    wo.src = (unsigned char*)(base + wl.src);
    wo.bases = (const i64*)(base + wl.src_base);
    wo.bad = (unsigned long long*)(base + wl.bad) + 2;  // behind the two counters of ugs_pack_parts
    UGS_CUDA_CHECK(cudaMemsetAsync(wo.bad, 0, 8, st));
    const u64* packed = nullptr;
    if (sample_scratch && sample_scratch_bytes >= ugs_sample_scratch_bytes(G, m))
        packed = reinterpret_cast<const u64*>((const char*)sample_scratch + packed_rows_offset(G * m));
    const int64_t nblk = (int64_t)ugs_delta8_blocks((size_t)total);
    k_wire_src_bases<<<(unsigned)((nblk + 255) / 256), 256, 0, st>>>(edge_ptr, G * m, m, total, w.cstart, (i64*)(base + wl.src_base));
    UGS_LAUNCH_CHECK("k_wire_src_bases");
    int tpi; int64_t ipg;
    small_items(m, G, &tpi, &ipg);
    const unsigned grid = (unsigned)((G * ipg + EMIT_WARPS - 1) / EMIT_WARPS);
    const size_t smem = sizeof(SmallEmit<unsigned short>) * EMIT_WARPS;
    UGS_CUDA_CHECK(cudaFuncSetAttribute(k_emit_small<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_emit_small<true, true><<<grid, EMIT_WARPS * 32, smem, st>>>(ptr, G, node_base, m, k, UGS_MODE_SAMPLE, tpi, ipg, 0, w, nodes,
                                                                  edge_ptr, nullptr, nullptr, nullptr, packed, 0, wo);
    UGS_LAUNCH_CHECK("k_emit_small_wire");
    return UGS_OK;
}

// flags: UGS_EMIT_GUARD_GENERAL / UGS_EMIT_COMPACT (ugs_common.cuh), set by the fused small-batch path only
int ugs_emit_edges_impl(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int mode, int use_perm,
                        int64_t n_general, const void* workspace, const ugs_layout* L, const int64_t* nodes,
                        const int64_t* edge_ptr, int64_t total, int64_t* edge_index_out, int64_t* edge_src_out,
                        const void* sample_scratch, size_t sample_scratch_bytes, int flags, void* stream_) {
    if (!workspace || !L || !edge_ptr) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (mode < 0 || mode > UGS_MODE_ENCODER) return ugs_set_error(UGS_EINVAL, "mode must be one of: 'sample', 'graph', 'global'");
    if (k < 1 || k > UGS_MAX_K) return ugs_set_error(UGS_ELIMIT, "k must be in [1, %d]", UGS_MAX_K);
    if (G <= 0 || m <= 0 || total <= 0) return UGS_OK;
    if (!nodes || !edge_index_out || !edge_src_out) return ugs_set_error(UGS_EINVAL, "NULL output");
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(const_cast<void*>(workspace), L);
    i64* out_u = edge_index_out;
    i64* out_v = edge_index_out + total;
    // UGS_EMIT_GUARD_GENERAL (ugs_sample_batch_fused with assume_no_general): no general-path launch, and the
    // bitmask-path kernels write nothing when the device count says there was such a graph after all
    if (flags & UGS_EMIT_GUARD_GENERAL) n_general = 0;
    const u64* packed = nullptr;  // rows as ugs_sample_batch packed them for exactly these (G, m, k <= 8)
    if (sample_scratch && k <= 8 && sample_scratch_bytes >= ugs_sample_scratch_bytes(G, m))
        packed = reinterpret_cast<const u64*>((const char*)sample_scratch + packed_rows_offset(G * m));
    if (n_general < G) {
        int tpi; int64_t ipg;
        small_items(m, G, &tpi, &ipg);
        const int64_t items = G * ipg;
        const unsigned grid = (unsigned)((items + EMIT_WARPS - 1) / EMIT_WARPS);
        if (k <= 8 && (mode == UGS_MODE_SAMPLE || mode == UGS_MODE_ENCODER) && m <= 16 && packed && !use_perm) {
            // several graphs per warp (the rows must be the packed ones ugs_sample_batch left behind)
            const int S = packed_lanes(m);
            const int64_t warps = (G + 32 / S - 1) / (32 / S);
            const unsigned pgrid = (unsigned)((warps + EMIT_WARPS - 1) / EMIT_WARPS);
            if (S == 10) {
                const size_t sl = sizeof(PackedEmit<3>) > sizeof(SmallEmit<unsigned short>) ? sizeof(PackedEmit<3>) : sizeof(SmallEmit<unsigned short>);
                UGS_CUDA_CHECK(cudaFuncSetAttribute(k_emit_small_packed<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sl * EMIT_WARPS)));
                k_emit_small_packed<10><<<pgrid, EMIT_WARPS * 32, sl * EMIT_WARPS, st>>>(ptr, G, node_base, m, k, mode, w, nodes, edge_ptr,
                                                                                        out_u, out_v, edge_src_out, packed, flags);
            } else if (S == 8) {
                const size_t sl = sizeof(PackedEmit<4>) > sizeof(SmallEmit<unsigned short>) ? sizeof(PackedEmit<4>) : sizeof(SmallEmit<unsigned short>);
                UGS_CUDA_CHECK(cudaFuncSetAttribute(k_emit_small_packed<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sl * EMIT_WARPS)));
                k_emit_small_packed<8><<<pgrid, EMIT_WARPS * 32, sl * EMIT_WARPS, st>>>(ptr, G, node_base, m, k, mode, w, nodes, edge_ptr,
                                                                                       out_u, out_v, edge_src_out, packed, flags);
            } else {
                const size_t sl = sizeof(PackedEmit<2>) > sizeof(SmallEmit<unsigned short>) ? sizeof(PackedEmit<2>) : sizeof(SmallEmit<unsigned short>);
                UGS_CUDA_CHECK(cudaFuncSetAttribute(k_emit_small_packed<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sl * EMIT_WARPS)));
                k_emit_small_packed<16><<<pgrid, EMIT_WARPS * 32, sl * EMIT_WARPS, st>>>(ptr, G, node_base, m, k, mode, w, nodes, edge_ptr,
                                                                                        out_u, out_v, edge_src_out, packed, flags);
            }
        } else if (k <= 8 && (mode == UGS_MODE_SAMPLE || mode == UGS_MODE_ENCODER)) {
            const size_t smem = sizeof(SmallEmit<unsigned short>) * EMIT_WARPS;
            UGS_CUDA_CHECK(cudaFuncSetAttribute(k_emit_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_emit_small<true><<<grid, EMIT_WARPS * 32, smem, st>>>(ptr, G, node_base, m, k, mode, tpi, ipg, use_perm, w,
                                                                   nodes, edge_ptr, out_u, out_v, edge_src_out, packed, flags);
        } else {
            const size_t smem = sizeof(SmallEmit<u32>) * EMIT_WARPS;
            UGS_CUDA_CHECK(cudaFuncSetAttribute(k_emit_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_emit_small<false><<<grid, EMIT_WARPS * 32, smem, st>>>(ptr, G, node_base, m, k, mode, tpi, ipg, use_perm, w,
                                                                    nodes, edge_ptr, out_u, out_v, edge_src_out, nullptr, flags);
        }
        UGS_LAUNCH_CHECK("k_emit_small");
    }
    if (n_general != 0) {
        const int64_t ngb = n_general < 0 ? G : n_general;
        int64_t blocks = (ngb * m + 7) / 8;
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        k_emit_general<false><<<(unsigned)blocks, 256, 0, st>>>(ptr, node_base, m, k, mode, n_general, true, w, nodes,
                                                               const_cast<i64*>(edge_ptr), total, out_u, out_v,
                                                               edge_src_out, flags);
        UGS_LAUNCH_CHECK("k_emit_general");
    }
    return UGS_OK;
}
