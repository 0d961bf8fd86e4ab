// ugs_apx.cu -- APX-UGS (approximate degree-dominating order + estimated cuts + rejection) on sm_100a.
//
// Replaces src/samplers/apx_ugs_sampler/src/apx_ugs_sampler.cpp (Bressan Alg. 5,7,8,9,10):
//   Graph::Graph        :10-41    columns [ptr[0],ptr[1]) only, n = max id + 1, sorted + deduplicated adjacency
//   apx_dd              :52-169   sampled order, h = ceil(10/eta^2 ln n) neighbour draws per vertex, re-sorts,
//                                 small-degree fix-up (exact estimates deg_in_G(v)^5)
//   estimate_cuts       :175-237  h in [10,100] draws per member, threshold ell
//   apx_rand_grow       :243-312
//   apx_prob            :318-382  <= 720 permutations
//   apx_ugs_sample_one  :388-455  <= 10^6 trials, accept min(1, (beta/Z) k^-2 / (est * p_hat))
//   sample_batch        :461-521  output [k, S], S <= m (failed samples dropped), ptr = arange(S+1)
// One graph per call (as in the reference), n <= 1024, k <= 8. The trials of a sample are independent, so they run
// in parallel -- one WARP per (sample, trial), the 100 neighbour draws of a cut estimate spread over the lanes --
// and the sample's result is its FIRST accepted trial (atomicMin), which is what the reference's sequential loop
// returns in law. Draw rule shared with oracle/apx_oracle.c (random-access Philox words).
#include <math.h>
#include <string.h>
#include "ugs_common.cuh"

#define APX_MAX_N 1024
#define APX_MAX_K 8
#define APX_NO_TRIAL 0x7fffffff

struct ApxConst {
    int n, k, h_dd;
    double eta, thr_dd, thr_small;          // apx_dd: threshold 2*eta*h, k/eta
    int h_cut;  double ell_cut;             // estimate_cuts as called from apx_rand_grow
    int h_prob; double ell_prob;            // estimate_cuts as called from apx_prob
    double beta, inv_k2;                    // acceptance: (beta / Z) * k^-2 / (est * p_hat)
    u32 k0, k1;                             // Philox key
};

struct ApxWS {
    u32* bits;       // [n][32] adjacency bitmap (n <= 1024)
    i32* rowptr;     // [n+1]
    i32* adj;        // sorted, deduplicated neighbours
    i32* deg;        // [n]
    i32* position;   // [n]
    i32* order;      // [n]
    double* est;     // [n]
    double* cum;     // [n] running sum of est in vertex order (root draw)
    double* Z;       // [1]
    i32* nn;         // [2]: n, nnz
    i32* best;       // [m] index of the first accepted trial
};

__device__ __forceinline__ u32 apx_word(const ApxConst& C, u32 c0, u32 c1, u32 c2, u32 c3, int i) {
    const Philox4 r = philox4x32_10(c0, c1, c2, c3, C.k0, C.k1);
    return (i == 0) ? r.x : (i == 1) ? r.y : (i == 2) ? r.z : r.w;
}
__device__ __forceinline__ double apx_u53(u32 a, u32 b) {
    return (double)(((((u64)a) << 32) | (u64)b) >> 11) * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------------------------------------------------
// Graph::Graph -- one CTA. The n x n bitmap makes "sort + unique" of every adjacency list an enumeration of bits.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_apx_build(const i64* __restrict__ row0, const i64* __restrict__ row1, i64 e0, i64 e1, ApxWS ws) {
    __shared__ int s_n;
    __shared__ int s_scan[APX_MAX_N + 1];
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) s_n = 0;
    __syncthreads();
    int mx = 0;
    for (i64 i = e0 + tid; i < e1; i += T) mx = max(mx, max((int)row0[i], (int)row1[i]) + 1);
    atomicMax(&s_n, mx);
    __syncthreads();
    const int n = s_n;
    if (tid == 0) ws.nn[0] = n;
    if (n > APX_MAX_N) { if (tid == 0) ws.nn[1] = -1; return; }
    for (int i = tid; i < n * 32; i += T) ws.bits[i] = 0;
    __syncthreads();
    for (i64 i = e0 + tid; i < e1; i += T) {
        const int u = (int)row0[i], v = (int)row1[i];
        atomicOr(&ws.bits[u * 32 + (v >> 5)], 1u << (v & 31));
        atomicOr(&ws.bits[v * 32 + (u >> 5)], 1u << (u & 31));
    }
    __syncthreads();
    for (int v = tid; v < n; v += T) {
        int d = 0;
        for (int q = 0; q < 32; ++q) d += __popc(ws.bits[v * 32 + q]);
        ws.deg[v] = d;
        s_scan[v] = d;
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int v = 0; v < n; ++v) { const int d = s_scan[v]; s_scan[v] = run; run += d; }
        s_scan[n] = run;
        ws.nn[1] = run;
    }
    __syncthreads();
    for (int v = tid; v <= n; v += T) ws.rowptr[v] = s_scan[v];
    for (int v = tid; v < n; v += T) {
        int o = s_scan[v];
        for (int q = 0; q < 32; ++q) {
            u32 wbits = ws.bits[v * 32 + q];
            while (wbits) { const int b = __ffs(wbits) - 1; wbits &= wbits - 1; ws.adj[o++] = q * 32 + b; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// apx_dd -- one CTA of 1024 threads, order / position / scores / estimates in shared memory.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool apx_before(const double* scores, int a, int b) {  // comparator :68-71
    if (scores[a] != scores[b]) return scores[a] > scores[b];
    return a > b;
}
// normalised bitonic network on ord[lo..n) with the comparator above (indices >= n act as "last")
__device__ void apx_sort_tail(int* ord, const double* scores, int lo, int n) {
    const int len = n - lo;
    for (int k2 = 2; (k2 >> 1) < len; k2 <<= 1) {
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
            const int p = i ^ (k2 - 1);
            if (p > i && p < len) { const int x = ord[lo + i], y = ord[lo + p]; if (apx_before(scores, y, x)) { ord[lo + i] = y; ord[lo + p] = x; } }
        }
        __syncthreads();
        for (int j = k2 >> 2; j >= 1; j >>= 1) {
            for (int i = threadIdx.x; i < len; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i && p < len) { const int x = ord[lo + i], y = ord[lo + p]; if (apx_before(scores, y, x)) { ord[lo + i] = y; ord[lo + p] = x; } }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(1024)
k_apx_dd(ApxConst C, ApxWS ws) {
    __shared__ int s_order[APX_MAX_N];
    __shared__ int s_pos[APX_MAX_N];
    __shared__ double s_scores[APX_MAX_N];
    __shared__ double s_est[APX_MAX_N];
    __shared__ int s_count;
    const int tid = threadIdx.x, T = blockDim.x, n = C.n, k = C.k;
    for (int v = tid; v < n; v += T) { s_scores[v] = (double)ws.deg[v]; s_order[v] = v; s_est[v] = 0.0; }
    __syncthreads();
    apx_sort_tail(s_order, s_scores, 0, n);
    for (int i = tid; i < n; i += T) s_pos[s_order[i]] = i;
    __syncthreads();
    for (int idx = 0; idx < n; ++idx) {
        const int v = s_order[idx];
        const int d = ws.deg[v];
        if (d == 0) { if (tid == 0) s_est[v] = 0.0; __syncthreads(); continue; }
        if (tid == 0) s_count = 0;
        __syncthreads();
        const i32* nb = ws.adj + ws.rowptr[v];
        const int pv = s_pos[v];
        int cnt = 0;
        for (int i4 = tid; i4 * 4 < C.h_dd; i4 += T) {  // 4 draws per Philox call
            const Philox4 r = philox4x32_10((u32)i4, (u32)idx, 0xA0000001u, 0u, C.k0, C.k1);
            const int base = i4 * 4;
            if (base + 0 < C.h_dd) cnt += (pv < s_pos[nb[__umulhi(r.x, (u32)d)]]) ? 1 : 0;
            if (base + 1 < C.h_dd) cnt += (pv < s_pos[nb[__umulhi(r.y, (u32)d)]]) ? 1 : 0;
            if (base + 2 < C.h_dd) cnt += (pv < s_pos[nb[__umulhi(r.z, (u32)d)]]) ? 1 : 0;
            if (base + 3 < C.h_dd) cnt += (pv < s_pos[nb[__umulhi(r.w, (u32)d)]]) ? 1 : 0;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(FULL_MASK, cnt, off);
        if ((tid & 31) == 0 && cnt) atomicAdd(&s_count, cnt);
        __syncthreads();
        const int count = s_count;
        __syncthreads();
        if ((double)count >= C.thr_dd) {
            if (tid == 0) { const double dd = (double)d; s_est[v] = dd * dd * dd * dd * dd; }
            __syncthreads();
        } else {
            if (tid == 0) { s_est[v] = 0.0; s_scores[v] = 3.0 * C.eta * d; }
            __syncthreads();
            apx_sort_tail(s_order, s_scores, idx, n);
            for (int i = idx + tid; i < n; i += T) s_pos[s_order[i]] = i;
            __syncthreads();
        }
    }
    // small-degree fix-up (:124-166)
    for (int v = tid; v < n; v += T) {
        if (!((double)ws.deg[v] <= C.thr_small)) continue;
        int q[APX_MAX_K];
        int qn = 0;
        bool found = false;
        q[qn++] = v;
        const int pv = s_pos[v];
        for (int head = 0; head < qn && qn < k && !found; ++head) {
            const int u = q[head];
            for (int p = ws.rowptr[u]; p < ws.rowptr[u + 1]; ++p) {
                const int wv = ws.adj[p];
                bool seen = false;
                for (int a = 0; a < qn; ++a) seen = seen || (q[a] == wv);
                if (seen || !(pv < s_pos[wv])) continue;
                q[qn++] = wv;
                if (qn >= k) { found = true; break; }
            }
        }
        double e = 0.0;
        if (found) {
            int dg = 0;
            for (int p = ws.rowptr[v]; p < ws.rowptr[v + 1]; ++p) dg += (pv < s_pos[ws.adj[p]]) ? 1 : 0;
            const double dd = (double)dg;
            e = dd * dd * dd * dd * dd;
        }
        s_est[v] = e;
    }
    __syncthreads();
    for (int v = tid; v < n; v += T) { ws.order[v] = s_order[v]; ws.position[v] = s_pos[v]; ws.est[v] = s_est[v]; }
    if (tid == 0) {
        double run = 0.0;
        for (int v = 0; v < n; ++v) { run += s_est[v]; ws.cum[v] = run; }
        ws.Z[0] = run;
    }
}

// ---------------------------------------------------------------------------------------------------------
// estimate_cuts (:175-237), one warp: U[] is warp-uniform; lane-parallel over the h <= 100 draws of a member.
// Returns the sum of the cuts (added in member order) and the per-member values in cuts[].
// ---------------------------------------------------------------------------------------------------------
__device__ double apx_estimate_cuts(const ApxConst& C, const ApxWS& ws, int v, const int* U, int sz, int h, double ell,
                                    u32 base0, u32 t, u32 s, u32 stream, double* cuts) {
    const int lane = threadIdx.x & 31;
    const int pv = ws.position[v];
    double total = 0.0;
    for (int a = 0; a < sz; ++a) {
        const int u = U[a];
        const int d = ws.deg[u];
        double c = 0.0;
        if (d != 0) {
            const int rp = ws.rowptr[u];
            int count = 0;
            for (int jb = 0; jb < h; jb += 32) {
                const int j = jb + lane;
                bool ok = false;
                if (j < h) {
                    const u32 wd = apx_word(C, base0 + (u32)(j >> 2) + 32u * (u32)a, t, s, stream, j & 3);
                    const int wv = ws.adj[rp + (int)__umulhi(wd, (u32)d)];
                    bool inU = false;
                    for (int b = 0; b < sz; ++b) inU = inU || (U[b] == wv);
                    ok = (pv < ws.position[wv]) && !inU;
                }
                count += __popc(__ballot_sync(FULL_MASK, ok));
            }
            if ((double)count >= ell) c = (double)(d * count) / (double)h;
        }
        cuts[a] = c;
        total += c;
    }
    return total;
}

__device__ __forceinline__ bool apx_has_edge(const ApxWS& ws, int u, int v) {
    return (ws.bits[u * 32 + (v >> 5)] >> (v & 31)) & 1u;
}

// One trial of apx_ugs_sample_one (:412-451), whole warp, every value warp-uniform. true = accepted, S filled.
__device__ bool apx_trial(const ApxConst& C, const ApxWS& ws, u32 t, u32 s, int* S) {
    const int k = C.k, n = C.n;
    const double Z = ws.Z[0];
    const Philox4 rr0 = philox4x32_10(t, s, 0xA0000002u, 0u, C.k0, C.k1);
    const double r = apx_u53(rr0.x, rr0.y) * Z;
    int lo = 0, hi = n;  // first u with r <= cum[u]; none -> v stays 0 (:416-424)
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (r <= ws.cum[mid]) hi = mid; else lo = mid + 1; }
    const int v = (lo < n) ? lo : 0;
    const double ev = ws.est[v];
    if (ev <= 0.0) return false;
    double cuts[APX_MAX_K];
    int sz = 1;
    S[0] = v;
    const int pv = ws.position[v];
    for (int i = 1; i < k; ++i) {
        const double total = apx_estimate_cuts(C, ws, v, S, sz, C.h_cut, C.ell_cut, 512u * (u32)i, t, s, 0xA0000003u, cuts);
        if (total <= 0.0) return false;
        const Philox4 g = philox4x32_10((u32)i, t, s, 0xA0000004u, C.k0, C.k1);
        const double rsel = apx_u53(g.x, g.y) * total;
        double cs = 0.0;
        int sel = S[0];
        for (int j = 0; j < sz; ++j) { cs += cuts[j]; if (rsel <= cs) { sel = S[j]; break; } }
        int nvalid = 0;
        for (int p = ws.rowptr[sel]; p < ws.rowptr[sel + 1]; ++p) {
            const int wv = ws.adj[p];
            bool inS = false;
            for (int b = 0; b < sz; ++b) inS = inS || (S[b] == wv);
            nvalid += ((pv < ws.position[wv]) && !inS) ? 1 : 0;
        }
        if (nvalid == 0) return false;
        int pick = (int)__umulhi(g.z, (u32)nvalid);
        int chosen = -1;
        for (int p = ws.rowptr[sel]; p < ws.rowptr[sel + 1]; ++p) {
            const int wv = ws.adj[p];
            bool inS = false;
            for (int b = 0; b < sz; ++b) inS = inS || (S[b] == wv);
            if (!(pv < ws.position[wv]) || inS) continue;
            if (pick == 0) { chosen = wv; break; }
            pick--;
        }
        S[sz++] = chosen;
    }
    // APX-PROB (:318-382)
    int perm[APX_MAX_K], rem[APX_MAX_K];
    for (int j = 1; j < k; ++j) rem[j - 1] = S[j];
    for (int a = 1; a < k - 1; ++a) {  // insertion sort of k-1 values
        const int x = rem[a];
        int b = a - 1;
        while (b >= 0 && rem[b] > x) { rem[b + 1] = rem[b]; --b; }
        rem[b + 1] = x;
    }
    double p_hat = 0.0;
    for (int q = 0; q < 720; ++q) {
        perm[0] = v;
        for (int j = 0; j < k - 1; ++j) perm[j + 1] = rem[j];
        double pp = 1.0;
        for (int i = 0; i < k - 1; ++i) {
            int ni = 0;
            for (int a = 0; a <= i; ++a) ni += apx_has_edge(ws, perm[a], perm[i + 1]) ? 1 : 0;
            const double ci = apx_estimate_cuts(C, ws, v, perm, i + 1, C.h_prob, C.ell_prob, 512u * (u32)i, t, s,
                                                0xA0000005u | ((u32)q << 8), cuts);
            if (ci > 0.0) pp *= (double)ni / ci; else { pp = 0.0; break; }
        }
        p_hat += pp;
        int a = k - 3;  // std::next_permutation
        while (a >= 0 && rem[a] >= rem[a + 1]) a--;
        if (a < 0) break;
        int b = k - 2;
        while (rem[b] <= rem[a]) b--;
        { const int tt = rem[a]; rem[a] = rem[b]; rem[b] = tt; }
        for (int x = a + 1, y = k - 2; x < y; ++x, --y) { const int tt = rem[x]; rem[x] = rem[y]; rem[y] = tt; }
    }
    if (p_hat <= 0.0) return false;
    double acc = (C.beta / Z) * C.inv_k2 / (ev * p_hat);
    if (acc > 1.0) acc = 1.0;
    return apx_u53(rr0.z, rr0.w) < acc;
}

// Round of trials: warp w handles sample s = w / per_sample, trial tbase + w % per_sample.
__global__ void __launch_bounds__(256)
k_apx_trials(ApxConst C, ApxWS ws, int m, int tbase, int per_sample, int max_trials) {
    const i64 wid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int s = (int)(wid / per_sample);
    if (s >= m) return;
    const int t = tbase + (int)(wid - (i64)s * per_sample);
    if (t >= max_trials || ws.best[s] <= t) return;  // an earlier trial of this sample was already accepted
    int S[APX_MAX_K];
    if (apx_trial(C, ws, (u32)t, (u32)s, S) && (threadIdx.x & 31) == 0) atomicMin(&ws.best[s], t);
}

// Regenerate the accepted trial of every sample and write the [k, S] output (failed samples dropped, :490-496).
__global__ void __launch_bounds__(256)
k_apx_write(ApxConst C, ApxWS ws, int m, const i32* __restrict__ slot, int n_ok, i64* __restrict__ samples_out) {
    const i64 wid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int s = (int)wid;
    if (s >= m || slot[s] < 0) return;
    int S[APX_MAX_K];
    apx_trial(C, ws, (u32)ws.best[s], (u32)s, S);
    const int lane = threadIdx.x & 31;
    if (lane < C.k) samples_out[(i64)lane * n_ok + slot[s]] = (i64)S[lane];
}

// =====================================================================================================
// C ABI
// =====================================================================================================
static inline size_t apx_align(size_t x) { return (x + 255) / 256 * 256; }

extern "C" size_t ugs_apx_workspace_bytes(int64_t n_columns, int64_t m) {
    size_t b = 0;
    b += apx_align(sizeof(u32) * APX_MAX_N * 32);
    b += apx_align(sizeof(i32) * (APX_MAX_N + 1));
    const size_t nnz_cap = (size_t)(2 * n_columns < (int64_t)APX_MAX_N * APX_MAX_N ? 2 * n_columns : (int64_t)APX_MAX_N * APX_MAX_N);
    b += apx_align(sizeof(i32) * (nnz_cap + 1));
    b += 3 * apx_align(sizeof(i32) * APX_MAX_N);
    b += 2 * apx_align(sizeof(double) * APX_MAX_N);
    b += apx_align(sizeof(double)) + apx_align(sizeof(i32) * 2);
    b += apx_align(sizeof(i32) * (size_t)(m + 1));
    return b;
}

static ApxWS apx_views(void* base, int64_t n_columns, int64_t m) {
    char* p = (char*)base;
    ApxWS w;
    auto take = [&](size_t bytes) { char* r = p; p += apx_align(bytes); return r; };
    w.bits = (u32*)take(sizeof(u32) * APX_MAX_N * 32);
    w.rowptr = (i32*)take(sizeof(i32) * (APX_MAX_N + 1));
    const size_t nnz_cap = (size_t)(2 * n_columns < (int64_t)APX_MAX_N * APX_MAX_N ? 2 * n_columns : (int64_t)APX_MAX_N * APX_MAX_N);
    w.adj = (i32*)take(sizeof(i32) * (nnz_cap + 1));
    w.deg = (i32*)take(sizeof(i32) * APX_MAX_N);
    w.position = (i32*)take(sizeof(i32) * APX_MAX_N);
    w.order = (i32*)take(sizeof(i32) * APX_MAX_N);
    w.est = (double*)take(sizeof(double) * APX_MAX_N);
    w.cum = (double*)take(sizeof(double) * APX_MAX_N);
    w.Z = (double*)take(sizeof(double));
    w.nn = (i32*)take(sizeof(i32) * 2);
    w.best = (i32*)take(sizeof(i32) * (size_t)(m + 1));
    (void)m;
    return w;
}

static void apx_cut_params(int k, double delta, double alpha, double beta, int* h_out, double* ell_out) {
    const double ell_raw = 1.0 / (k * delta * alpha * alpha);
    const double h_double = ell_raw * ell_raw * log(k / beta);
    int h;
    if (isinf(h_double) || h_double > 100) h = 100; else if (h_double < 10.0) h = 10; else h = (int)ceil(h_double);
    *h_out = h;
    *ell_out = ell_raw < h * 0.5 ? ell_raw : h * 0.5;
}

// Step 1: build the graph and the approximate order. Synchronises the stream once (needs n on the host).
// h_info[0] = n, h_info[1] = number of adjacency entries, h_info[2] = 1 if Z > 0.
extern "C" int ugs_apx_prepare(const int64_t* edge_index, int64_t E, int64_t e0, int64_t e1, int64_t m, int k,
                               uint64_t seed, double epsilon, void* workspace, size_t ws_bytes, int64_t* h_info,
                               void* stream_) {
    if (!workspace || !h_info) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (k < 2 || k > APX_MAX_K) return ugs_set_error(UGS_ELIMIT, "apx_ugs: k must be in [2, %d]", APX_MAX_K);
    if (!(epsilon > 0.0)) return ugs_set_error(UGS_EINVAL, "epsilon must be > 0");
    if (e0 < 0 || e1 > E || e0 > e1) return ugs_set_error(UGS_EINVAL, "ptr must describe a column range inside edge_index");
    if (ws_bytes < ugs_apx_workspace_bytes(e1 - e0, m)) return ugs_set_error(UGS_EINVAL, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream_;
    ApxWS ws = apx_views(workspace, e1 - e0, m);
    k_apx_build<<<1, 1024, 0, st>>>(edge_index, edge_index + E, e0, e1, ws);
    UGS_LAUNCH_CHECK("k_apx_build");
    i32 nn[2];
    UGS_CUDA_CHECK(cudaMemcpyAsync(nn, ws.nn, sizeof(nn), cudaMemcpyDeviceToHost, st));
    UGS_CUDA_CHECK(cudaStreamSynchronize(st));
    h_info[0] = nn[0]; h_info[1] = nn[1]; h_info[2] = 0;
    if (nn[0] > APX_MAX_N) return ugs_set_error(UGS_ELIMIT, "apx_ugs: graphs above %d vertices are not supported", APX_MAX_N);
    if (nn[0] < k) return UGS_OK;  // :473-478 empty result
    ApxConst C;
    memset(&C, 0, sizeof(C));
    C.n = nn[0]; C.k = k; C.k0 = (u32)seed; C.k1 = (u32)(seed >> 32);
    const double beta = epsilon / 2.0;
    C.eta = pow(beta, 1.0 / (double)(k - 1)) / (6.0 * k * k);
    C.h_dd = (int)ceil(10.0 / (C.eta * C.eta) * log((double)C.n));
    C.thr_dd = 2.0 * C.eta * C.h_dd;
    C.thr_small = (double)k / C.eta;
    k_apx_dd<<<1, 1024, 0, st>>>(C, ws);
    UGS_LAUNCH_CHECK("k_apx_dd");
    double Z = 0.0;
    UGS_CUDA_CHECK(cudaMemcpyAsync(&Z, ws.Z, sizeof(double), cudaMemcpyDeviceToHost, st));
    UGS_CUDA_CHECK(cudaMemsetAsync(ws.best, 0x7f, sizeof(i32) * (size_t)(m + 1), st));  // 0x7f7f7f7f > any trial index
    UGS_CUDA_CHECK(cudaStreamSynchronize(st));
    h_info[2] = Z > 0.0 ? 1 : 0;
    return UGS_OK;
}

static ApxConst apx_sample_consts(int n, int k, uint64_t seed, double epsilon) {
    ApxConst C;
    memset(&C, 0, sizeof(C));
    C.n = n; C.k = k; C.k0 = (u32)seed; C.k1 = (u32)(seed >> 32);
    const double beta = epsilon / 2.0;
    const double alpha = pow(beta, 1.0 / (double)(k - 1)) / (6.0 * k * k * k);
    const double gamma = epsilon * pow(3.0, -k) * pow((double)k, -2.0);
    apx_cut_params(k, gamma / pow((double)k, 4.0), alpha, beta, &C.h_cut, &C.ell_cut);                  // apx_rand_grow :254,258
    apx_cut_params(k, gamma / (k * k), alpha, beta / pow((double)k, 6.0), &C.h_prob, &C.ell_prob);      // apx_prob :358-359
    C.beta = beta;
    C.inv_k2 = pow((double)k, -2.0);
    return C;
}

// Step 2: one round of `per_sample` trials for each of the m samples starting at trial index tbase. h_best (host,
// m ints) receives the first accepted trial per sample so far (>= APX_NO_TRIAL/2 = none). Synchronises.
extern "C" int ugs_apx_trials(int n, int64_t n_columns, int64_t m, int k, uint64_t seed, double epsilon, int tbase,
                              int per_sample, int max_trials, void* workspace, int32_t* h_best, void* stream_) {
    if (!workspace || !h_best) return ugs_set_error(UGS_EINVAL, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream_;
    ApxWS ws = apx_views(workspace, n_columns, m);
    const ApxConst C = apx_sample_consts(n, k, seed, epsilon);
    const i64 warps = (i64)m * per_sample;
    if (warps > 0) {
        k_apx_trials<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(C, ws, (int)m, tbase, per_sample, max_trials);
        UGS_LAUNCH_CHECK("k_apx_trials");
    }
    UGS_CUDA_CHECK(cudaMemcpyAsync(h_best, ws.best, sizeof(i32) * (size_t)m, cudaMemcpyDeviceToHost, st));
    UGS_CUDA_CHECK(cudaStreamSynchronize(st));
    return UGS_OK;
}

// Step 3: write samples[k, n_ok]; slot[s] (device, m ints) = output column of sample s or -1.
extern "C" int ugs_apx_write(int n, int64_t n_columns, int64_t m, int k, uint64_t seed, double epsilon,
                             const int32_t* slot, int n_ok, void* workspace, int64_t* samples_out, void* stream_) {
    if (!workspace || !slot || (n_ok > 0 && !samples_out)) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (n_ok <= 0) return UGS_OK;
    cudaStream_t st = (cudaStream_t)stream_;
    ApxWS ws = apx_views(workspace, n_columns, m);
    const ApxConst C = apx_sample_consts(n, k, seed, epsilon);
    k_apx_write<<<(unsigned)((m + 7) / 8), 256, 0, st>>>(C, ws, (int)m, slot, n_ok, samples_out);
    UGS_LAUNCH_CHECK("k_apx_write");
    return UGS_OK;
}

// Test hook: copy the graph tables to the host (rowptr[n+1], adj[nnz], order[n], position[n], est[n]).
extern "C" int ugs_apx_dump(int n, int64_t nnz, int64_t n_columns, int64_t m, void* workspace, int32_t* rowptr,
                            int32_t* adj, int32_t* order, int32_t* position, double* est, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    ApxWS ws = apx_views(workspace, n_columns, m);
    UGS_CUDA_CHECK(cudaMemcpyAsync(rowptr, ws.rowptr, sizeof(i32) * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
    UGS_CUDA_CHECK(cudaMemcpyAsync(adj, ws.adj, sizeof(i32) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    UGS_CUDA_CHECK(cudaMemcpyAsync(order, ws.order, sizeof(i32) * (size_t)n, cudaMemcpyDeviceToHost, st));
    UGS_CUDA_CHECK(cudaMemcpyAsync(position, ws.position, sizeof(i32) * (size_t)n, cudaMemcpyDeviceToHost, st));
    UGS_CUDA_CHECK(cudaMemcpyAsync(est, ws.est, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    UGS_CUDA_CHECK(cudaStreamSynchronize(st));
    return UGS_OK;
}
