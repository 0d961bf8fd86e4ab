// ugs_variants.cu -- the epsilon-uniform variant on sm_100a.
//
// Replaces src/samplers/epsilon_uniform_sampler/src/epsilon_uniform_sampler.cpp:
//   sample_connected_subgraph_rw :18-84    random frontier growth, path weight w
//   rejection loop               :225-257  accept with min(1, eps/(w+eps)), <= max(10, 10/eps) attempts
//   edge extraction              :259-291  every batch column with both endpoints in the sample, column order
// It runs on the same workspace as the UGS path (only the multigraph CSR is used: the reference's adjacency
// lists keep duplicates exactly like build_csr, :170-174).
// Draw rule shared with oracle/eps_oracle.c: Philox4x32-10, counter (sample, word block, graph, 0x80000000|attempt),
// words consumed in order (start; frontier pick [+ candidate pick] per iteration; two words for the acceptance u53).
#include "ugs_common.cuh"

int ugs_run_scan(i64* edge_ptr, i64 B, void* scratch, size_t scratch_bytes, WS& w, cudaStream_t st);  // ugs_sample.cu

struct WordStream {
    u32 c0, c2, c3, k0, k1, block, idx;
    Philox4 r;
    __device__ __forceinline__ void init(u32 sample, u64 graph, u32 attempt, u64 seed) {
        c0 = sample; c2 = (u32)graph; c3 = 0x80000000u | attempt;
        k0 = (u32)seed; k1 = (u32)(seed >> 32); block = 0; idx = 4;
    }
    __device__ __forceinline__ u32 next() {
        if (idx == 4) { r = philox4x32_10(c0, block++, c2, c3, k0, k1); idx = 0; }
        const u32 v = (idx == 0) ? r.x : (idx == 1) ? r.y : (idx == 2) ? r.z : r.w;
        idx++;
        return v;
    }
};

__device__ __forceinline__ int max_attempts_for(double eps) {
    const int a = (int)(10.0 / eps);
    return a < 10 ? 10 : a;  // epsilon_uniform_sampler.cpp:207
}

// 16 packed bytes (lo: positions 0-7, hi: 8-15)
__device__ __forceinline__ u32 pk_get(u64 lo, u64 hi, int i) {
    return (u32)(((i < 8) ? (lo >> (8 * i)) : (hi >> (8 * (i - 8)))) & 0xFFull);
}
__device__ __forceinline__ void pk_set(u64& lo, u64& hi, int i, u32 v) {
    if (i < 8) lo = (lo & ~(0xFFull << (8 * i))) | ((u64)v << (8 * i));
    else hi = (hi & ~(0xFFull << (8 * (i - 8)))) | ((u64)v << (8 * (i - 8)));
}
__device__ __forceinline__ void pk_erase(u64& lo, u64& hi, int i) {  // remove byte i, shift the rest down
    if (i < 8) {
        const u64 keep = (i == 0) ? 0ull : (~0ull >> (64 - 8 * i));
        lo = (lo & keep) | ((lo >> 8) & ~keep);
        lo = (lo & 0x00FFFFFFFFFFFFFFull) | (hi << 56);
        hi >>= 8;
    } else {
        const int q = i - 8;
        const u64 keep = (q == 0) ? 0ull : (~0ull >> (64 - 8 * q));
        hi = (hi & keep) | ((hi >> 8) & ~keep);
    }
}

// =====================================================================================================
// Bitmask path: one thread per sample
// =====================================================================================================
#define EPS_WARPS 8
struct __align__(16) EpsSamp {
    u64 adj[64];
    unsigned char tile[32][16];
    unsigned short rowptr[66];
    unsigned char idx[2 * UGS_SMALL_MAX_COLS];
};

__global__ void __launch_bounds__(EPS_WARPS * 32)
k_eps_sample_small(const i64* __restrict__ ptr, i64 G, i64 node_base, i64 m, int k, u64 seed, u64 gbase, double eps,
                   int tiles_per_item, i64 items_per_graph, WS w, i64* __restrict__ nodes_out,
                   i64* __restrict__ edge_ptr, i64* __restrict__ sample_ptr) {
    __shared__ EpsSamp sm_all[EPS_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 item = (i64)blockIdx.x * EPS_WARPS + warp;
    const i64 g = item / items_per_graph;
    if (g >= G) return;
    const int chunk = (int)(item - g * items_per_graph);
    if (chunk == 0 && lane == 0) {
        sample_ptr[g + 1] = (g + 1) * m;  // :318, always m rows per graph
        if (g == 0) { sample_ptr[0] = 0; edge_ptr[0] = 0; }
    }
    const ugs_ginfo gi = w.ginfo[g];
    if (gi.cls == 2) return;
    const int ntiles = (int)((m + 31) >> 5);
    const int t0 = chunk * tiles_per_item, t1 = min(ntiles, t0 + tiles_per_item);
    const i64 lo = ptr[g];
    const int r32 = 32 % k, q32 = 32 / k;
    if (gi.cls == 0) {  // n < k: every attempt fails (:26) -> rows of -1
        for (int t = t0; t < t1; ++t) {
            const i64 s0 = g * m + (i64)t * 32;
            const int nv = (int)min((i64)32, m - (i64)t * 32);
            for (int e = lane; e < nv * k; e += 32) nodes_out[s0 * k + e] = -1;
            if (lane < nv) edge_ptr[1 + s0 + lane] = 0;
        }
        return;
    }
    EpsSamp& s = sm_all[warp];
    const int n = (int)(ptr[g + 1] - lo);
    const i64 nb = lo - node_base;
    {
        const u32* rp = w.rowptr + nb + g;
        const i32* ci = w.csr_idx + 2 * w.cstart[g];
        for (int v = lane; v < n; v += 32) s.adj[v] = w.adj[nb + v];
        for (int v = lane; v <= n; v += 32) s.rowptr[v] = (unsigned short)rp[v];
        for (int p = lane; p < gi.nnz; p += 32) s.idx[p] = (unsigned char)ci[p];
    }
    __syncwarp();
    const u64 gidx = gbase + (u64)g;
    const int max_att = max_attempts_for(eps);
    unsigned long long trials = 0, n_fail = 0;

    for (int t = t0; t < t1; ++t) {
        const i64 i = (i64)t * 32 + lane;
        const bool active = i < m;
        u64 nlo = ~0ull, nhi = ~0ull;
        i64 cnt = 0;
        if (active) {
            bool ok = false;
            u64 S = 0;
            for (int a = 0; a < max_att && !ok; ++a) {
                WordStream rs; rs.init((u32)i, gidx, (u32)a, seed);
                trials++;
                const int start = (int)__umulhi(rs.next(), (u32)n);
                S = 1ull << start;
                u64 flo = (u64)start, fhi = 0;
                int size = 1, fn = 1;
                double wt = 1.0;
                wt *= (1.0 / n);
                int iters = 0;
                while (size < k && iters < 100 * k) {
                    iters++;
                    if (fn == 0) break;
                    const int fi = (int)__umulhi(rs.next(), (u32)fn);
                    const int u = (int)pk_get(flo, fhi, fi);
                    int ncand;
                    if (gi.mult > 0) ncand = gi.mult * __popcll(s.adj[u] & ~S);
                    else {
                        ncand = 0;
                        for (int p = s.rowptr[u]; p < s.rowptr[u + 1]; ++p) ncand += (int)(((~S) >> s.idx[p]) & 1ull);
                    }
                    if (ncand == 0) { pk_erase(flo, fhi, fi); fn--; continue; }
                    int ci_ = (int)__umulhi(rs.next(), (u32)ncand);
                    int chosen = 0;
                    for (int p = s.rowptr[u]; p < s.rowptr[u + 1]; ++p) {
                        const int v = s.idx[p];
                        if ((S >> v) & 1ull) continue;
                        if (ci_ == 0) { chosen = v; break; }
                        ci_--;
                    }
                    S |= 1ull << chosen;
                    pk_set(flo, fhi, fn, (u32)chosen);
                    fn++; size++;
                    wt *= (1.0 / (double)fn) * (1.0 / (double)ncand);
                }
                if (size == k) {
                    double pacc = eps / (wt + eps);
                    if (pacc > 1.0) pacc = 1.0;
                    const u32 w0 = rs.next(), w1 = rs.next();
                    const u64 u53 = ((((u64)w0) << 32) | (u64)w1) >> 11;
                    const double uu = (double)u53 * (1.0 / 9007199254740992.0);
                    ok = (uu <= pacc);
                }
            }
            if (ok) {
                // sorted node row (:256) = set bits of S in ascending order; columns with both ends in S = half
                // of the CSR entries between members
                u64 rest = S;
                int c = 0;
                for (int j = 0; j < k; ++j) {
                    const int v = __ffsll((long long)rest) - 1;
                    rest &= rest - 1;
                    pk_set(nlo, nhi, j, (u32)v);
                    if (gi.mult > 0) c += gi.mult * __popcll(s.adj[v] & S);
                    else for (int p = s.rowptr[v]; p < s.rowptr[v + 1]; ++p) c += (int)((S >> s.idx[p]) & 1ull);
                }
                cnt = c >> 1;
            } else {
                n_fail++;
            }
        }
        *reinterpret_cast<ulonglong2*>(&s.tile[lane][0]) = make_ulonglong2(nlo, nhi);
        __syncwarp();
        const i64 s0 = g * m + (i64)t * 32;
        const int nv = (int)min((i64)32, m - (i64)t * 32);
        {
            int sidx = lane / k, j = lane - sidx * k;
            i64* out = nodes_out + s0 * k;
            for (int e = lane; e < nv * k; e += 32) {
                const unsigned char bv = s.tile[sidx][j];
                out[e] = (bv == 0xFF) ? (i64)-1 : lo + (i64)bv;
                j += r32; sidx += q32;
                if (j >= k) { j -= k; sidx++; }
            }
        }
        if (active) edge_ptr[1 + s0 + lane] = cnt;
        __syncwarp();
    }
    trials = __reduce_add_sync(FULL_MASK, (unsigned)trials);
    n_fail = __reduce_add_sync(FULL_MASK, (unsigned)n_fail);
    if (lane == 0) {
        if (trials) atomicAdd(&w.counters[UGS_CTR_TRIALS], trials);
        if (n_fail) atomicAdd(&w.counters[UGS_CTR_INCOMPLETE], n_fail);
    }
}

// =====================================================================================================
// General path: one warp per sample
// =====================================================================================================
__device__ __forceinline__ bool eps_member(int x, int mine, int size) {
    bool hit = false;
    for (int j = 0; j < size; ++j) {
        const int mj = __shfl_sync(FULL_MASK, mine, j);
        hit = hit || (x == mj);
    }
    return hit;
}

__global__ void __launch_bounds__(256)
k_eps_sample_general(const i64* __restrict__ ptr, i64 node_base, i64 m, int k, u64 seed, u64 gbase, double eps,
                     i64 n_general, WS w, i64* __restrict__ nodes_out, i64* __restrict__ edge_ptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    const i64 total = n_general * m;
    const i64 nwarps = (i64)gridDim.x * wpb;
    const int max_att = max_attempts_for(eps);
    unsigned long long trials = 0, n_fail = 0;
    for (i64 t = (i64)blockIdx.x * wpb + warp; t < total; t += nwarps) {
        const i64 gi_ = t / m;
        const i64 i = t - gi_ * m;
        const i64 g = w.glist[gi_];
        const i64 lo = ptr[g];
        const int n = (int)(ptr[g + 1] - lo);
        const i64 nb = lo - node_base;
        const u32* rp = w.rowptr + nb + g;
        const i32* cidx = w.csr_idx + 2 * w.cstart[g];
        const u64 gidx = gbase + (u64)g;
        bool ok = false;
        int mine = -1;
        for (int a = 0; a < max_att && !ok; ++a) {
            WordStream rs; rs.init((u32)i, gidx, (u32)a, seed);
            trials++;
            const int start = (int)__umulhi(rs.next(), (u32)n);
            mine = (lane == 0) ? start : -1;       // lane j holds S[j]
            int fr = (lane == 0) ? start : -1;     // lane j holds frontier[j]
            int size = 1, fn = 1;
            double wt = 1.0;
            wt *= (1.0 / n);
            int iters = 0;
            while (size < k && iters < 100 * k) {
                iters++;
                if (fn == 0) break;
                const int fi = (int)__umulhi(rs.next(), (u32)fn);
                const int u = __shfl_sync(FULL_MASK, fr, fi);
                const u32 p0 = rp[u], p1 = rp[u + 1];
                int ncand = 0;
                for (u32 pb = p0; pb < p1; pb += 32) {
                    const u32 p = pb + lane;
                    const int y = (p < p1) ? cidx[p] : -1;
                    const bool mem = eps_member(y, mine, size);
                    ncand += __popc(__ballot_sync(FULL_MASK, (p < p1) && !mem));
                }
                if (ncand == 0) {  // erase frontier[fi], order kept (:69)
                    const int nxt = __shfl_down_sync(FULL_MASK, fr, 1);
                    if (lane >= fi) fr = (lane == 31) ? -1 : nxt;
                    fn--;
                    continue;
                }
                int ci_ = (int)__umulhi(rs.next(), (u32)ncand);
                int chosen = -1;
                for (u32 pb = p0; pb < p1; pb += 32) {
                    const u32 p = pb + lane;
                    const int y = (p < p1) ? cidx[p] : -1;
                    const bool mem = eps_member(y, mine, size);
                    const u32 ball = __ballot_sync(FULL_MASK, (p < p1) && !mem);
                    const int c = __popc(ball);
                    if (ci_ < c) {
                        const int src = select_bit32(ball, ci_);
                        chosen = __shfl_sync(FULL_MASK, y, src);
                        break;
                    }
                    ci_ -= c;
                }
                if (lane == size) mine = chosen;
                if (lane == fn) fr = chosen;
                size++; fn++;
                wt *= (1.0 / (double)fn) * (1.0 / (double)ncand);
            }
            if (size == k) {
                double pacc = eps / (wt + eps);
                if (pacc > 1.0) pacc = 1.0;
                const u32 w0 = rs.next(), w1 = rs.next();
                const u64 u53 = ((((u64)w0) << 32) | (u64)w1) >> 11;
                const double uu = (double)u53 * (1.0 / 9007199254740992.0);
                ok = (uu <= pacc);
            }
        }
        const i64 sidx = g * m + i;
        if (ok) {
            // rank of each member among the k members = its position in the sorted row
            int rank = 0;
            for (int j = 0; j < k; ++j) {
                const int mj = __shfl_sync(FULL_MASK, mine, j);
                rank += (lane < k && mj < mine) ? 1 : 0;
            }
            if (lane < k) nodes_out[sidx * k + rank] = lo + (i64)mine;
            int cntE = 0;
            for (int j = 0; j < k; ++j) {
                const int u = __shfl_sync(FULL_MASK, mine, j);
                const u32 p0 = rp[u], p1 = rp[u + 1];
                for (u32 pb = p0; pb < p1; pb += 32) {
                    const u32 p = pb + lane;
                    const int y = (p < p1) ? cidx[p] : -1;
                    const bool mem = eps_member(y, mine, k);
                    cntE += __popc(__ballot_sync(FULL_MASK, (p < p1) && mem));
                }
            }
            if (lane == 0) edge_ptr[1 + sidx] = cntE >> 1;
        } else {
            if (lane < k) nodes_out[sidx * k + lane] = -1;
            if (lane == 0) { edge_ptr[1 + sidx] = 0; n_fail++; }
        }
        (void)lt_mask;
    }
    if (lane == 0) {
        if (trials) atomicAdd(&w.counters[UGS_CTR_TRIALS], trials);
        if (n_fail) atomicAdd(&w.counters[UGS_CTR_INCOMPLETE], n_fail);
    }
}

// =====================================================================================================
// Emission, bitmask path: thread per sample scans the graph's columns in order
// =====================================================================================================
#define EPS_STAGE 1024
struct __align__(16) EpsEmit {
    u32 stage[EPS_STAGE];                       // column offset << 8 | b << 4 | a
    unsigned char cu[UGS_SMALL_MAX_COLS];
    unsigned char cv[UGS_SMALL_MAX_COLS];       // 0xFF = column dropped (endpoint outside the graph)
    unsigned char tile[32][16];
};

__global__ void __launch_bounds__(EPS_WARPS * 32)
k_eps_emit_small(const i64* __restrict__ row0, const i64* __restrict__ row1, const i64* __restrict__ ptr, i64 G,
                 i64 m, int k, int mode, const i64* __restrict__ perm, int tiles_per_item, i64 items_per_graph, WS w,
                 const i64* __restrict__ nodes, const i64* __restrict__ edge_ptr, i64* __restrict__ out_u,
                 i64* __restrict__ out_v, i64* __restrict__ out_src) {
    extern __shared__ __align__(16) unsigned char eps_smem_raw[];
    EpsEmit* sm_all = reinterpret_cast<EpsEmit*>(eps_smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 item = (i64)blockIdx.x * EPS_WARPS + warp;
    const i64 g = item / items_per_graph;
    if (g >= G) return;
    const ugs_ginfo gi = w.ginfo[g];
    if (gi.cls != 1 || gi.nnz == 0) return;
    const int chunk = (int)(item - g * items_per_graph);
    const int ntiles = (int)((m + 31) >> 5);
    const int t0 = chunk * tiles_per_item, t1 = min(ntiles, t0 + tiles_per_item);
    if (t0 >= t1) return;
    {
        const i64 sA = g * m + (i64)t0 * 32, sB = min(g * m + m, g * m + (i64)t1 * 32);
        if (edge_ptr[sA] == edge_ptr[sB]) return;
    }
    const i64 lo = ptr[g];
    const int n = (int)(ptr[g + 1] - lo);
    const i64 c0 = w.cstart[g];
    const int nc = (int)(w.cstart[g + 1] - c0);
    EpsEmit& s = sm_all[warp];
    for (int c = lane; c < nc; c += 32) {
        const i64 j = perm ? perm[c0 + c] : c0 + c;
        const i64 u = row0[j] - lo, v = row1[j] - lo;
        s.cu[c] = (unsigned char)u;
        s.cv[c] = (v >= 0 && v < n) ? (unsigned char)v : (unsigned char)0xFF;
    }
    const int r32 = 32 % k, q32 = 32 / k;
    for (int t = t0; t < t1; ++t) {
        const i64 s0 = g * m + (i64)t * 32;
        const int nv = (int)min((i64)32, m - (i64)t * 32);
        __syncwarp();
        {
            int sidx = lane / k, j = lane - sidx * k;
            const i64* in = nodes + s0 * k;
            for (int e = lane; e < nv * k; e += 32) {
                const i64 val = in[e];
                s.tile[sidx][j] = (val < 0) ? (unsigned char)0xFF : (unsigned char)(val - lo);
                j += r32; sidx += q32;
                if (j >= k) { j -= k; sidx++; }
            }
        }
        const i64 my_lo = (lane < nv) ? edge_ptr[s0 + lane] : 0;
        const i64 tile_end = edge_ptr[s0 + nv];
        const i64 my_hi = __shfl_down_sync(FULL_MASK, my_lo, 1);
        const i64 my_end = (lane == nv - 1 || lane == 31) ? tile_end : my_hi;
        __syncwarp();
        if (tile_end == __shfl_sync(FULL_MASK, my_lo, 0)) continue;
        u64 S = 0;
        const bool walk = (lane < nv) && (my_end > my_lo);
        if (walk) {
            const ulonglong2 nn = *reinterpret_cast<const ulonglong2*>(&s.tile[lane][0]);
            for (int j = 0; j < k; ++j) S |= 1ull << pk_get(nn.x, nn.y, j);
        }
        int sa = 0;
        while (sa < nv) {
            const i64 base = __shfl_sync(FULL_MASK, my_lo, sa);
            const u32 fits = __ballot_sync(FULL_MASK, lane >= sa && lane < nv && (my_end - base) <= EPS_STAGE);
            const u32 nofit = ~fits & ~((1u << sa) - 1u);
            int sb = nofit ? min(nv, __ffs(nofit) - 1) : nv;
            if (sb == sa) sb = sa + 1;  // a sample has at most nc <= 384 entries
            const int run = (int)(((sb == nv) ? tile_end : __shfl_sync(FULL_MASK, my_lo, sb & 31)) - base);
            if (walk && lane >= sa && lane < sb) {
                u32* st = s.stage + (int)(my_lo - base);
                for (int c = 0; c < nc; ++c) {
                    const u32 u = s.cu[c], v = s.cv[c];
                    if (v != 0xFFu && ((S >> u) & 1ull) && ((S >> v) & 1ull)) {
                        const u32 a = (u32)__popcll(S & ((1ull << u) - 1ull));   // position in the sorted row
                        const u32 b = (u32)__popcll(S & ((1ull << v) - 1ull));
                        *st++ = ((u32)c << 8) | (b << 4) | a;
                    }
                }
            }
            __syncwarp();
            for (int e = lane; e < run; e += 32) {
                const u32 pk = s.stage[e];
                const i64 o = base + e;
                const int c = (int)(pk >> 8);
                i64 av, bv;
                if (mode == UGS_MODE_SAMPLE) { av = pk & 15u; bv = (pk >> 4) & 15u; }
                else { av = lo + s.cu[c]; bv = lo + s.cv[c]; }  // :282-285
                out_u[o] = av;
                out_v[o] = bv;
                out_src[o] = perm ? perm[c0 + c] : c0 + c;
            }
            __syncwarp();
            sa = sb;
        }
    }
}

// =====================================================================================================
// Emission, general path: warp per sample; hits are gathered from the CSR rows of the members (source side
// only), ordered by column with a rank count, and stored at their final position.
// =====================================================================================================
#define EPSG_CAP 1024
__global__ void __launch_bounds__(128)
k_eps_emit_general(const i64* __restrict__ row0, const i64* __restrict__ ptr, i64 node_base, i64 m, int k, int mode,
                   i64 n_general, WS w, const i64* __restrict__ nodes, const i64* __restrict__ edge_ptr,
                   i64* __restrict__ out_u, i64* __restrict__ out_v, i64* __restrict__ out_src) {
    __shared__ i32 h_col[4][EPSG_CAP];
    __shared__ unsigned short h_ab[4][EPSG_CAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    const i64 nwarps = (i64)gridDim.x * wpb;
    const i64 total_items = n_general * m;
    for (i64 t = (i64)blockIdx.x * wpb + warp; t < total_items; t += nwarps) {
        const i64 gi_ = t / m;
        const i64 i = t - gi_ * m;
        const i64 g = w.glist[gi_];
        const i64 sidx = g * m + i;
        const i64 base = edge_ptr[sidx];
        const int want = (int)(edge_ptr[sidx + 1] - base);
        if (want == 0) continue;
        const i64 lo = ptr[g];
        const i64 nb = lo - node_base;
        const u32* rp = w.rowptr + nb + g;
        const i32* cidx = w.csr_idx + 2 * w.cstart[g];
        const i32* ccol = w.csr_col + 2 * w.cstart[g];
        const i64 nv = (lane < k) ? nodes[sidx * k + lane] : 0;
        const int mine = (lane < k) ? (int)(nv - lo) : -1;  // sorted ascending
        int h = 0;
        for (int j = 0; j < k; ++j) {
            const int u = __shfl_sync(FULL_MASK, mine, j);
            const u32 p0 = rp[u], p1 = rp[u + 1];
            for (u32 pb = p0; pb < p1; pb += 32) {
                const u32 p = pb + lane;
                int y = -2, l = -1, col = -1;
                if (p < p1) { y = cidx[p]; col = ccol[p]; }
                for (int q = 0; q < k; ++q) {
                    const int mq = __shfl_sync(FULL_MASK, mine, q);
                    if (y == mq) l = q;
                }
                bool hit = (l >= 0) && (p < p1);
                if (hit) {
                    // keep the entry that sits in the row of the column's SOURCE endpoint; a self-loop column
                    // appears twice in the same row: keep the first
                    if (row0[col] - lo != u) hit = false;
                    else if (y == u && p > p0 && ccol[p - 1] == col) hit = false;
                }
                const u32 ball = __ballot_sync(FULL_MASK, hit);
                if (hit) {
                    const int o = h + __popc(ball & lt_mask);
                    if (o < EPSG_CAP) { h_col[warp][o] = col; h_ab[warp][o] = (unsigned short)((l << 8) | j); }
                }
                h += __popc(ball);
            }
        }
        __syncwarp();
        if (h > EPSG_CAP || h != want) {  // cannot happen for consistent inputs; never write out of range
            if (lane == 0) atomicAdd(&w.counters[UGS_CTR_EMIT_OVERFLOW], 1ull);
            continue;
        }
        for (int e = lane; e < h; e += 32) {
            const i32 col = h_col[warp][e];
            int rank = 0;
            for (int q = 0; q < h; ++q) rank += (h_col[warp][q] < col) ? 1 : 0;
            const int a = h_ab[warp][e] & 0xFF, b = h_ab[warp][e] >> 8;
            i64 av = a, bv = b;
            if (mode != UGS_MODE_SAMPLE) { av = nodes[sidx * k + a]; bv = nodes[sidx * k + b]; }  // :282-285
            const i64 o = base + rank;
            out_u[o] = av;
            out_v[o] = bv;
            out_src[o] = (i64)col;
        }
        __syncwarp();
    }
}

// =====================================================================================================
// Host launchers
// =====================================================================================================
static int eps_sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}
static inline void eps_items(int64_t m, int* tiles_per_item, int64_t* items_per_graph) {
    const int64_t ntiles = (m + 31) / 32;
    int tpi = (int)(ntiles < 4 ? (ntiles < 1 ? 1 : ntiles) : 4);
    *tiles_per_item = tpi;
    *items_per_graph = (ntiles + tpi - 1) / tpi;
    if (*items_per_graph < 1) *items_per_graph = 1;
}

extern "C" int ugs_eps_sample_batch(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, uint64_t seed,
                                    uint64_t graph_index_base, double epsilon, int64_t n_general, void* workspace,
                                    const ugs_layout* L, void* scratch, size_t scratch_bytes, int64_t* nodes,
                                    int64_t* edge_ptr, int64_t* sample_ptr, void* stream_) {
    if (!workspace || !L || !edge_ptr || !sample_ptr) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (!(epsilon > 0.0 && epsilon <= 1.0)) return ugs_set_error(UGS_EINVAL, "epsilon must be in (0, 1]");
    if (k < 1 || k > UGS_MAX_K) return ugs_set_error(UGS_ELIMIT, "k must be in [1, %d]", UGS_MAX_K);
    if (m < 0 || m >= (1ll << 32)) return ugs_set_error(UGS_EINVAL, "m_per_graph must be in [0, 2^32)");
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(workspace, L);
    UGS_CUDA_CHECK(cudaMemsetAsync(&w.counters[UGS_CTR_TOTAL_EDGES], 0, sizeof(i64) * 4, st));
    if (G <= 0 || m == 0) {
        UGS_CUDA_CHECK(cudaMemsetAsync(edge_ptr, 0, sizeof(i64), st));
        UGS_CUDA_CHECK(cudaMemsetAsync(sample_ptr, 0, sizeof(i64) * (size_t)((G > 0 ? G : 0) + 1), st));
        return UGS_OK;
    }
    if (!nodes) return ugs_set_error(UGS_EINVAL, "nodes is NULL");
    int tpi; int64_t ipg;
    eps_items(m, &tpi, &ipg);
    const int64_t items = G * ipg;
    k_eps_sample_small<<<(unsigned)((items + EPS_WARPS - 1) / EPS_WARPS), EPS_WARPS * 32, 0, st>>>(
        ptr, G, node_base, m, k, seed, graph_index_base, epsilon, tpi, ipg, w, nodes, edge_ptr, sample_ptr);
    UGS_LAUNCH_CHECK("k_eps_sample_small");
    if (n_general > 0) {
        int64_t blocks = (n_general * m + 7) / 8;
        const int64_t cap = (int64_t)eps_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        k_eps_sample_general<<<(unsigned)blocks, 256, 0, st>>>(ptr, node_base, m, k, seed, graph_index_base, epsilon,
                                                              n_general, w, nodes, edge_ptr);
        UGS_LAUNCH_CHECK("k_eps_sample_general");
    }
    return ugs_run_scan(edge_ptr, G * m, scratch, scratch_bytes, w, st);
}

extern "C" int ugs_eps_emit_edges(const int64_t* edge_index, int64_t E, const int64_t* ptr, int64_t G,
                                  int64_t node_base, int64_t m, int k, int mode, const int64_t* perm,
                                  int64_t n_general, const void* workspace, const ugs_layout* L, const int64_t* nodes,
                                  const int64_t* edge_ptr, int64_t total, int64_t* edge_index_out,
                                  int64_t* edge_src_out, void* stream_) {
    if (!workspace || !L || !edge_ptr) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (k < 1 || k > UGS_MAX_K) return ugs_set_error(UGS_ELIMIT, "k must be in [1, %d]", UGS_MAX_K);
    if (G <= 0 || m <= 0 || total <= 0) return UGS_OK;
    if (!nodes || !edge_index_out || !edge_src_out || !edge_index) return ugs_set_error(UGS_EINVAL, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(const_cast<void*>(workspace), L);
    i64* out_u = edge_index_out;
    i64* out_v = edge_index_out + total;
    if (n_general < G) {
        int tpi; int64_t ipg;
        eps_items(m, &tpi, &ipg);
        const int64_t items = G * ipg;
        const size_t smem = sizeof(EpsEmit) * EPS_WARPS;
        UGS_CUDA_CHECK(cudaFuncSetAttribute(k_eps_emit_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_eps_emit_small<<<(unsigned)((items + EPS_WARPS - 1) / EPS_WARPS), EPS_WARPS * 32, smem, st>>>(
            edge_index, edge_index + E, ptr, G, m, k, mode, perm, tpi, ipg, w, nodes, edge_ptr, out_u, out_v,
            edge_src_out);
        UGS_LAUNCH_CHECK("k_eps_emit_small");
    }
    if (n_general > 0) {
        UGS_CUDA_CHECK(cudaMemsetAsync(&w.counters[UGS_CTR_EMIT_OVERFLOW], 0, sizeof(i64), st));
        int64_t blocks = (n_general * m + 3) / 4;
        const int64_t cap = (int64_t)eps_sm_count() * 16;
        if (blocks > cap) blocks = cap;
        k_eps_emit_general<<<(unsigned)blocks, 128, 0, st>>>(edge_index, ptr, node_base, m, k, mode, n_general, w, nodes,
                                                            edge_ptr, out_u, out_v, edge_src_out);
        UGS_LAUNCH_CHECK("k_eps_emit_general");
    }
    return UGS_OK;
}

// =====================================================================================================
// Exact-uniform accept/reject variant (rule stated in oracle/uni_oracle.c): bitmask-path graphs here,
// general-path graphs in ugs_uniform_general.cu.
// =====================================================================================================
#define UNI_WARPS 8

// One warp per graph: degree-dominating order and its tables, overwriting the reference-order tables.
__global__ void __launch_bounds__(UNI_WARPS * 32)
k_uniform_tables_small(const i64* __restrict__ ptr, i64 G, i64 node_base, int k, WS w) {
    __shared__ u64 s_adj[UNI_WARPS][64];
    __shared__ u64 s_suf[UNI_WARPS][64];
    __shared__ double s_b[UNI_WARPS][64];
    __shared__ unsigned char s_order[UNI_WARPS][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 g = (i64)blockIdx.x * UNI_WARPS + warp;
    if (g >= G) return;
    ugs_ginfo gi = w.ginfo[g];
    if (gi.cls != 1) return;
    const i64 lo = ptr[g];
    const int n = (int)(ptr[g + 1] - lo);
    const i64 nb = lo - node_base;
    u64 a0 = 0, a1 = 0;  // simple adjacency (no self bit) of vertices lane and lane+32
    if (lane < n) a0 = w.adj[nb + lane] & ~(1ull << lane);
    if (lane + 32 < n) a1 = w.adj[nb + lane + 32] & ~(1ull << (lane + 32));
    s_adj[warp][lane] = a0; s_adj[warp][lane + 32] = a1;
    u64 rem = (n == 64) ? ~0ull : ((1ull << n) - 1ull);
    int my_d0 = 0, my_d1 = 0;  // dsuf of positions lane / lane+32 (filled as they are removed)
    for (int vi = 0; vi < n; ++vi) {
        int key0 = ((rem >> lane) & 1ull) ? (__popcll(a0 & rem) * 64 + (63 - lane)) : -1;
        int key1 = ((rem >> (lane + 32)) & 1ull) ? (__popcll(a1 & rem) * 64 + (63 - (lane + 32))) : -1;
        int key = max(key0, key1);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) key = max(key, __shfl_xor_sync(FULL_MASK, key, off));
        const int v = 63 - (key & 63), d = key >> 6;
        if (lane == (vi & 31)) {
            s_order[warp][vi] = (unsigned char)v;
            s_suf[warp][vi] = rem;
            if (vi < 32) my_d0 = d; else my_d1 = d;
        }
        rem &= ~(1ull << v);
    }
    __syncwarp();
    // reachability, weights, acceptance bound per position
    double qmin = 1e300;
    int lastv = -1, n0 = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int vi = lane + 32 * h;
        double b = 0.0;
        if (vi < n) {
            const int v = s_order[warp][vi];
            const u64 sufm = s_suf[warp][vi];
            u64 vis = 1ull << v, expd = 0;
            bool reach = false;
            for (int it = 0; it < k; ++it) {
                if (__popcll(vis) >= k) { reach = true; break; }
                const u64 fr = vis & ~expd;
                if (!fr) break;
                const int u = __ffsll((long long)fr) - 1;
                expd |= 1ull << u;
                vis |= s_adj[warp][u] & sufm;
            }
            if (reach) {
                const int dd = h ? my_d1 : my_d0;
                const int d = dd > 1 ? dd : 1;
                b = 1.0;
                for (int t = 0; t < k - 1; ++t) b *= (double)d;
                double den = 1.0;
                const int nsuf = n - vi;
                for (int i = 1; i <= k - 1; ++i) { const int a = i * d, c = nsuf - i; den *= (double)(a < c ? a : c); }
                const double q = b / den;
                qmin = fmin(qmin, q);
                lastv = vi; n0++;
            }
            w.order[nb + vi] = v;
            w.index_of[nb + v] = vi;
            w.suffix_deg[nb + vi] = h ? my_d1 : my_d0;
            w.suf[nb + vi] = sufm;
            w.bucket_b[nb + vi] = b;
        }
        s_b[warp][vi] = b;
    }
    __syncwarp();
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        qmin = fmin(qmin, __shfl_xor_sync(FULL_MASK, qmin, off));
        lastv = max(lastv, __shfl_xor_sync(FULL_MASK, lastv, off));
        n0 += __shfl_xor_sync(FULL_MASK, n0, off);
    }
    // running sums (exact while the total is <= 2^53, see k_preproc_small)
    const double p0 = s_b[warp][2 * lane], p1 = s_b[warp][2 * lane + 1];
    double incl = p0 + p1;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double y = __shfl_up_sync(FULL_MASK, incl, off);
        if (lane >= off) incl += y;
    }
    const double total = __shfl_sync(FULL_MASK, incl, 31);
    double cum0, cum1, Z;
    if (total <= 9007199254740992.0) {
        const double excl = incl - (p0 + p1);
        cum0 = excl + p0; cum1 = cum0 + p1; Z = total;
    } else {
        double run = 0.0; cum0 = cum1 = 0.0;
        for (int vi = 0; vi < 64; ++vi) {
            run += s_b[warp][vi];
            if (vi == 2 * lane) cum0 = run;
            if (vi == 2 * lane + 1) cum1 = run;
        }
        Z = run;
    }
    if (2 * lane < n) w.cum[nb + 2 * lane] = cum0;
    if (2 * lane + 1 < n) w.cum[nb + 2 * lane + 1] = cum1;
    if (lane == 0) {
        gi.Z = Z; gi.level = 0; gi.n_eligible = n0; gi.last_viable = lastv;
        w.ginfo[g] = gi;
        w.gaux[g] = (n0 > 0) ? qmin : 0.0;
    }
}

struct __align__(16) UniSamp {
    u64 adj[64];
    u64 suf[64];
    double cum[64];
    double b[64];
    unsigned char tile[32][16];
    unsigned char order[64];
    unsigned short rowptr[66];
    unsigned char idx[2 * UGS_SMALL_MAX_COLS];
};

__global__ void __launch_bounds__(UNI_WARPS * 32)
k_sample_uniform_small(const i64* __restrict__ ptr, i64 G, i64 node_base, i64 m, int k, u64 seed, u64 gbase,
                       int tiles_per_item, i64 items_per_graph, WS w, i64* __restrict__ nodes_out,
                       i64* __restrict__ edge_ptr, i64* __restrict__ sample_ptr) {
    extern __shared__ __align__(16) unsigned char uni_smem_raw[];
    UniSamp* sm_all = reinterpret_cast<UniSamp*>(uni_smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 item = (i64)blockIdx.x * UNI_WARPS + warp;
    const i64 g = item / items_per_graph;
    if (g >= G) return;
    const int chunk = (int)(item - g * items_per_graph);
    if (chunk == 0 && lane == 0) {
        sample_ptr[g + 1] = (g + 1) * m;
        if (g == 0) { sample_ptr[0] = 0; edge_ptr[0] = 0; }
    }
    const ugs_ginfo gi = w.ginfo[g];
    const int ntiles = (int)((m + 31) >> 5);
    const int t0 = chunk * tiles_per_item, t1 = min(ntiles, t0 + tiles_per_item);
    const i64 lo = ptr[g];
    const int r32 = 32 % k, q32 = 32 / k;
    if (gi.cls != 1) {  // skipped graphs (n < k); general-path graphs are rejected on the host
        for (int t = t0; t < t1; ++t) {
            const i64 s0 = g * m + (i64)t * 32;
            const int nv = (int)min((i64)32, m - (i64)t * 32);
            for (int e = lane; e < nv * k; e += 32) nodes_out[s0 * k + e] = -1;
            if (lane < nv) edge_ptr[1 + s0 + lane] = 0;
        }
        return;
    }
    UniSamp& s = sm_all[warp];
    const int n = (int)(ptr[g + 1] - lo);
    const i64 nb = lo - node_base;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int v = lane + 32 * h;
        if (v < n) {
            s.adj[v] = w.adj[nb + v] & ~(1ull << v);
            s.suf[v] = w.suf[nb + v];
            s.cum[v] = w.cum[nb + v];
            s.b[v] = w.bucket_b[nb + v];
            s.order[v] = (unsigned char)w.order[nb + v];
        }
    }
    {
        const u32* rp = w.rowptr + nb + g;
        const i32* ci = w.csr_idx + 2 * w.cstart[g];
        for (int v = lane; v <= n; v += 32) s.rowptr[v] = (unsigned short)rp[v];
        for (int p = lane; p < gi.nnz; p += 32) s.idx[p] = (unsigned char)ci[p];
    }
    __syncwarp();
    const double C = w.gaux[g];
    const u64 gidx = gbase + (u64)g;
    const int full = (1 << (k - 1)) - 1;
    unsigned long long trials = 0, n_fail = 0;

    for (int t = t0; t < t1; ++t) {
        const i64 i = (i64)t * 32 + lane;
        const bool active = i < m;
        u64 nlo = ~0ull, nhi = ~0ull;
        i64 cnt = 0;
        if (active) {
            bool ok = false;
            int mem[8];
            u64 S = 0;
            if (gi.Z > 0.0) {
                for (int a = 0; a < UGS_UNIFORM_MAX_TRIALS && !ok; ++a) {
                    WordStream rs; rs.init((u32)i, gidx, (u32)a, seed);
                    rs.c3 = 0x40000000u | (u32)a;
                    trials++;
                    const u32 w0 = rs.next(), w1 = rs.next();
                    const double x = root_x(w0, w1, gi.Z);
                    int lo_ = 0, hi_ = n - 1;
                    while (lo_ < hi_) { const int mid = (lo_ + hi_) >> 1; if (s.cum[mid] > x) hi_ = mid; else lo_ = mid + 1; }
                    if (!(s.cum[lo_] > x)) lo_ = gi.last_viable;
                    const int root_vi = lo_;
                    const u64 sufm = s.suf[root_vi];
                    mem[0] = s.order[root_vi];
                    S = 1ull << mem[0];
                    u64 nbr = s.adj[mem[0]];
                    int size = 1;
#pragma unroll
                    for (int step = 1; step < 8; ++step) {
                        if (step < k && size == step) {
                            const u64 cut = nbr & sufm & ~S;
                            const int c = __popcll(cut);
                            if (c > 0) {
                                const int x2 = select_bit64(cut, (int)__umulhi(rs.next(), (u32)c));
                                S |= 1ull << x2; nbr |= s.adj[x2]; mem[step] = x2; size++;
                            }
                        }
                    }
                    if (size < k) continue;
                    // p(S|r): push DP over subsets of S \ {r} (float32, same operation order as the oracle)
                    float f[128];
                    for (int q = 0; q <= full; ++q) f[q] = 0.0f;
                    f[0] = 1.0f;
                    for (int q = 0; q < full; ++q) {
                        const float fq = f[q];
                        if (fq == 0.0f) continue;
                        u64 nb_ = s.adj[mem[0]], vm = 1ull << mem[0];
#pragma unroll
                        for (int j = 1; j < 8; ++j)
                            if (j < k && ((q >> (j - 1)) & 1)) { nb_ |= s.adj[mem[j]]; vm |= 1ull << mem[j]; }
                        const u64 cut = nb_ & sufm & ~vm;
                        const int c = __popcll(cut);
                        if (c == 0) continue;
                        const float qv = __fdiv_rn(fq, (float)c);
#pragma unroll
                        for (int j = 1; j < 8; ++j)
                            if (j < k && !((q >> (j - 1)) & 1) && ((cut >> mem[j]) & 1ull))
                                f[q | (1 << (j - 1))] = __fadd_rn(f[q | (1 << (j - 1))], qv);
                    }
                    const double den = __dmul_rn(s.b[root_vi], (double)f[full]);
                    const double pa = C / den;
                    const u32 a0 = rs.next(), a1 = rs.next();
                    const u64 v53 = ((((u64)a0) << 32) | (u64)a1) >> 11;
                    const double uu = (double)v53 * (1.0 / 9007199254740992.0);
                    ok = (uu < pa);
                }
            }
            if (ok) {
                int c = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < k) {
                        const int v = mem[j];
                        pk_set(nlo, nhi, j, (u32)v);
                        for (int p = s.rowptr[v]; p < s.rowptr[v + 1]; ++p) c += (int)((S >> s.idx[p]) & 1ull);
                    }
                }
                cnt = c;
            } else {
                n_fail++;
            }
        }
        *reinterpret_cast<ulonglong2*>(&s.tile[lane][0]) = make_ulonglong2(nlo, nhi);
        __syncwarp();
        const i64 s0 = g * m + (i64)t * 32;
        const int nv = (int)min((i64)32, m - (i64)t * 32);
        {
            int sidx = lane / k, j = lane - sidx * k;
            i64* out = nodes_out + s0 * k;
            for (int e = lane; e < nv * k; e += 32) {
                const unsigned char bv = s.tile[sidx][j];
                out[e] = (bv == 0xFF) ? (i64)-1 : lo + (i64)bv;
                j += r32; sidx += q32;
                if (j >= k) { j -= k; sidx++; }
            }
        }
        if (active) edge_ptr[1 + s0 + lane] = cnt;
        __syncwarp();
    }
    trials = __reduce_add_sync(FULL_MASK, (unsigned)trials);
    n_fail = __reduce_add_sync(FULL_MASK, (unsigned)n_fail);
    if (lane == 0) {
        if (trials) atomicAdd(&w.counters[UGS_CTR_TRIALS], trials);
        if (n_fail) atomicAdd(&w.counters[UGS_CTR_INCOMPLETE], n_fail);
    }
}

int ugs_launch_uniform_tables_general(const i64* ptr, i64 node_base, int k, i64 n_general, i64 max_n_general, WS& w,
                                      cudaStream_t st);  // ugs_uniform_general.cu
int ugs_launch_sample_uniform_general(const i64* ptr, i64 node_base, i64 m, int k, u64 seed, u64 gbase, i64 n_general,
                                      WS& w, i64* nodes, i64* edge_ptr, cudaStream_t st);

extern "C" int ugs_preproc_uniform(const int64_t* ptr, int64_t G, int64_t node_base, int k, int64_t n_general,
                                   int64_t max_n_general, void* workspace, const ugs_layout* L, void* stream_) {
    if (!workspace || !L) return ugs_set_error(UGS_EINVAL, "NULL workspace/layout");
    if (k < 1 || k > 8) return ugs_set_error(UGS_ELIMIT, "the uniform variant supports k <= 8");
    if (G <= 0) return UGS_OK;
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(workspace, L);
    if (n_general < G) {
        k_uniform_tables_small<<<(unsigned)((G + UNI_WARPS - 1) / UNI_WARPS), UNI_WARPS * 32, 0, st>>>(ptr, G, node_base, k, w);
        UGS_LAUNCH_CHECK("k_uniform_tables_small");
    }
    return ugs_launch_uniform_tables_general(ptr, node_base, k, n_general, max_n_general, w, st);
}

// called from ugs_sample_batch (ugs_sample.cu) for UGS_VARIANT_UNIFORM
int ugs_launch_sample_uniform(const i64* ptr, i64 G, i64 node_base, i64 m, int k, u64 seed, u64 gbase, i64 n_general,
                              WS& w, i64* nodes, i64* edge_ptr, i64* sample_ptr, cudaStream_t st) {
    if (k > 8) return ugs_set_error(UGS_ELIMIT, "the uniform variant supports k <= 8");
    int tpi; int64_t ipg;
    eps_items(m, &tpi, &ipg);
    const int64_t items = G * ipg;
    const size_t smem = sizeof(UniSamp) * UNI_WARPS;
    UGS_CUDA_CHECK(cudaFuncSetAttribute(k_sample_uniform_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sample_uniform_small<<<(unsigned)((items + UNI_WARPS - 1) / UNI_WARPS), UNI_WARPS * 32, smem, st>>>(
        ptr, G, node_base, m, k, seed, gbase, tpi, ipg, w, nodes, edge_ptr, sample_ptr);
    UGS_LAUNCH_CHECK("k_sample_uniform_small");
    // general-path graphs (the kernel above wrote -1 rows for them): one warp per sample, after it on the same stream
    return ugs_launch_sample_uniform_general(ptr, node_base, m, k, seed, gbase, n_general, w, nodes, edge_ptr, st);
}
