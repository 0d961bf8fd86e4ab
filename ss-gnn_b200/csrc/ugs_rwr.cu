// ugs_rwr.cu -- Random-Walk-with-Restart connected k-subgraph sampler on sm_100a.
//
// Replaces the reference's src/samplers/rwr_sampler/src/rwr_sampler.cpp:
//   build_adj_per_graph :31-69   adjacency with duplicates, both directions per column, column order
//                                == the multigraph CSR ugs_preproc_batch already builds (same rows, same order)
//   sample_batch        :72-295  per sample: seed vertex uniform in [0,n); repeat at most n*k*10 times: draw r in
//                                [0,1); if r < p_restart or the current vertex has no neighbours go back to the seed,
//                                else move to a uniform entry of the current row (duplicates kept); a vertex seen for
//                                the first time joins the sample; stop at k vertices (:167-193). Fewer than k at the
//                                cap, n == 0 or n < k: an all -1 row without edges (:139-162,196-204).
//   edges               :217-256 for u in sample order, for every entry v of row(u) in row order with v in the sample:
//                                (pos(u), pos(v)) in mode "sample", batch-global ids otherwise; edge_src = -1.
//                                This is exactly the UGS extraction (A.7) with the source column dropped, so the count /
//                                scan / emission kernels of ugs_sample.cu are reused and edge_src is filled with -1.
// Randomness (the reference seeds one SplitMix64 per graph with seed + g + 104729*omp_thread, i.e. its values depend
// on OpenMP scheduling): Philox4x32-10, key = seed, counter = (sample i, word block, graph index, 0xC0000000); words
// are consumed in order: seed vertex = mulhi32(word, n); per iteration one word for r = word * 2^-32 and, when the
// walk moves, one word for the row entry = mulhi32(word, row length). oracle/rwr_oracle.c follows the same rule.
// Bitmask-path graphs: k_rwr_sample_small (rows in shared memory, lanes pull samples from a shared cursor);
// general-path graphs: one thread per sample, rows read through L1/L2.
#include "ugs_common.cuh"

int ugs_run_scan(i64* edge_ptr, i64 B, void* scratch, size_t scratch_bytes, WS& w, cudaStream_t st);  // ugs_sample.cu

struct RwrStream {
    u32 c0, c2, k0, k1, block, idx;
    Philox4 r;
    __device__ __forceinline__ void init(u32 sample, u64 graph, u64 seed) {
        c0 = sample; c2 = (u32)graph; k0 = (u32)seed; k1 = (u32)(seed >> 32); block = 0; idx = 4;
    }
    __device__ __forceinline__ u32 next() {
        if (idx == 4) { r = philox4x32_10(c0, block++, c2, 0xC0000000u, k0, k1); idx = 0; }
        const u32 v = (idx == 0) ? r.x : (idx == 1) ? r.y : (idx == 2) ? r.z : r.w;
        idx++;
        return v;
    }
};

// ---- bitmask-path graphs (<= 64 vertices, <= 384 columns, k <= 16): the graph's rows are staged in shared memory once
//      per warp, a sample's vertex set is a 64-bit mask, and the lanes of a warp pull samples of their graph from a shared
//      cursor: a walk takes 36 iterations on average but several hundred in the tail, so a fixed sample-per-lane mapping
//      leaves most lanes waiting for the slowest (7 of 32 lanes busy per instruction in the thread-per-sample kernel
//      below). Finished lanes park until RWR_PARK of them can write their rows together, then take new samples.
#define RWR_WARPS 8
#define RWR_ITEM 256   // samples of one graph handled by one warp
#define RWR_PARK 8
struct __align__(16) RwrSmall {
    u64 adj[64];
    unsigned short rowptr[66];
    unsigned char idx[2 * UGS_SMALL_MAX_COLS];
    int next;  // cursor: next sample of the item
};

__global__ void __launch_bounds__(RWR_WARPS * 32)
k_rwr_sample_small(const i64* __restrict__ ptr, i64 G, i64 node_base, i64 m, int k, u64 seed, u64 gbase, double p_restart,
                   i64 items_per_graph, WS w, i64* __restrict__ nodes_out, i64* __restrict__ edge_ptr,
                   i64* __restrict__ sample_ptr) {
    __shared__ RwrSmall sm_all[RWR_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 item = (i64)blockIdx.x * RWR_WARPS + warp;
    const i64 g = item / items_per_graph;
    if (g >= G) return;
    const int chunk = (int)(item - g * items_per_graph);
    if (chunk == 0 && lane == 0) {
        sample_ptr[g + 1] = (g + 1) * m;  // rwr_sampler.cpp:268-271 (every graph contributes m rows)
        if (g == 0) { sample_ptr[0] = 0; edge_ptr[0] = 0; }
    }
    const ugs_ginfo gi = w.ginfo[g];
    if (gi.cls == 2) return;  // general-path graphs: k_rwr_sample
    const i64 i0 = (i64)chunk * RWR_ITEM, i1 = min(m, i0 + RWR_ITEM);
    if (gi.cls == 0) {  // n == 0 or n < k: rows of -1 without edges (:139-162)
        for (i64 e = i0 * k + lane; e < i1 * k; e += 32) nodes_out[g * m * k + e] = -1;
        for (i64 i = i0 + lane; i < i1; i += 32) edge_ptr[1 + g * m + i] = 0;
        return;
    }
    RwrSmall& s = sm_all[warp];
    const i64 lo = ptr[g];
    const int n = (int)(ptr[g + 1] - lo);
    const i64 nb = lo - node_base;
    {
        const u32* rp = w.rowptr + nb + g;
        const i32* ci = w.csr_idx + 2 * w.cstart[g];
        for (int v = lane; v <= n; v += 32) s.rowptr[v] = (unsigned short)rp[v];
        for (int v = lane; v < n; v += 32) s.adj[v] = w.adj[nb + v];
        for (int p = lane; p < gi.nnz; p += 32) s.idx[p] = (unsigned char)ci[p];
        if (lane == 0) s.next = 0;
    }
    __syncwarp();
    const int count_i = (int)(i1 - i0);
    const int max_iter = n * k * 10;  // :180
    const u64 gidx = gbase + (u64)g;
    unsigned long long n_iter = 0;
    unsigned n_fail = 0;
    // lane state: 0 = needs a sample, 1 = walking, 2 = parked (finished, row not written yet), 3 = no samples left
    int state = 0, size = 0, it = 0, cur = 0, seed_node = 0, my_i = 0;
    u64 S = 0, clo = 0, chi = 0;  // vertex set; chosen vertices as packed bytes (visiting order)
    RwrStream rs;
    rs.init(0u, gidx, seed);
    while (true) {
        // ---- refill: every lane in state 0 takes the next sample of the item ----------------------------------------
        const u32 need = __ballot_sync(FULL_MASK, state == 0);
        if (need) {
            int base = 0;
            if (lane == 0) { base = s.next; s.next = base + __popc(need); }
            base = __shfl_sync(FULL_MASK, base, 0);
            if (state == 0) {
                my_i = base + __popc(need & ((1u << lane) - 1u));
                if (my_i < count_i) {
                    rs.init((u32)(i0 + my_i), gidx, seed);
                    seed_node = (int)__umulhi(rs.next(), (u32)n);
                    cur = seed_node;
                    S = 1ull << cur; clo = (u64)cur; chi = 0; size = 1; it = 0;
                    state = (size < k && max_iter > 0) ? 1 : 2;
                } else {
                    state = 3;
                }
            }
        }
        if (!__any_sync(FULL_MASK, state == 1 || state == 2)) break;
        // ---- walk until enough lanes are parked (or nobody walks any more) ------------------------------------------
        while (true) {
            if (state == 1) {
                ++it;
                const double r = (double)rs.next() * (1.0 / 4294967296.0);
                const u32 p0 = s.rowptr[cur], p1 = s.rowptr[cur + 1];
                if (r < p_restart || p0 == p1) cur = seed_node;
                else cur = s.idx[p0 + __umulhi(rs.next(), p1 - p0)];
                if (!((S >> cur) & 1ull)) {
                    S |= 1ull << cur;
                    if (size < 8) clo |= (u64)cur << (8 * size); else chi |= (u64)cur << (8 * (size - 8));
                    size++;
                }
                if (size == k || it >= max_iter) state = 2;
            }
            const u32 walking = __ballot_sync(FULL_MASK, state == 1);
            const u32 parked = __ballot_sync(FULL_MASK, state == 2);
            if (walking == 0u || __popc(parked) >= RWR_PARK) break;
        }
        // ---- parked lanes write their rows and entry counts together --------------------------------------------------
        if (state == 2) {
            const i64 srow = g * m + i0 + my_i;
            i64* row = nodes_out + srow * k;
            n_iter += (unsigned long long)it;
            i64 cnt = 0;
            if (size == k) {
                for (int j = 0; j < k; ++j) {
                    const int u = (int)(((j < 8) ? (clo >> (8 * j)) : (chi >> (8 * (j - 8)))) & 0xFF);
                    row[j] = lo + u;
                    if (gi.mult > 0) cnt += (i64)gi.mult * __popcll(s.adj[u] & S);
                    else for (int p = s.rowptr[u]; p < s.rowptr[u + 1]; ++p) cnt += (i64)((S >> s.idx[p]) & 1ull);
                }
            } else {
                n_fail++;
                for (int j = 0; j < k; ++j) row[j] = -1;
            }
            edge_ptr[1 + srow] = cnt;
            state = 0;
        }
    }
    n_fail = __reduce_add_sync(FULL_MASK, n_fail);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) n_iter += __shfl_xor_sync(FULL_MASK, n_iter, off);
    if (lane == 0) {
        if (n_iter) atomicAdd(&w.counters[UGS_CTR_TRIALS], n_iter);  // walk iterations
        if (n_fail) atomicAdd(&w.counters[UGS_CTR_INCOMPLETE], (unsigned long long)n_fail);
    }
}

// ---- general-path graphs: one thread per sample, rows read through L1/L2 (bitmask-path graphs are skipped) -------------
__global__ void __launch_bounds__(128)
k_rwr_sample(const i64* __restrict__ ptr, i64 G, i64 node_base, i64 m, int k, u64 seed, u64 gbase, double p_restart,
             WS w, i64* __restrict__ nodes_out, i64* __restrict__ edge_ptr, i64* __restrict__ sample_ptr) {
    const i64 total = G * m;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    unsigned long long n_fail = 0, n_iter = 0;
    for (i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += stride) {
        const i64 g = s / m;
        const i64 i = s - g * m;
        const ugs_ginfo gi = w.ginfo[g];
        if (gi.cls != 2) continue;  // k_rwr_sample_small
        i64* row = nodes_out + s * k;
        int size = 0;
        i64 cnt = 0;
        i32 chosen[UGS_MAX_K];
        {
            const i64 lo = ptr[g];
            const int n = (int)(ptr[g + 1] - lo);
            const i64 nb = lo - node_base;
            const u32* __restrict__ rp = w.rowptr + nb + g;
            const i32* __restrict__ cidx = w.csr_idx + 2 * w.cstart[g];
            RwrStream rs;
            rs.init((u32)i, gbase + (u64)g, seed);
            const int seed_node = (int)__umulhi(rs.next(), (u32)n);
            int cur = seed_node;
            chosen[size++] = cur;
            const long long max_iter = (long long)n * k * 10;  // :180
            long long it = 0;
            while (size < k && it < max_iter) {
                ++it;
                const double r = (double)rs.next() * (1.0 / 4294967296.0);
                const u32 p0 = rp[cur], p1 = rp[cur + 1];
                if (r < p_restart || p0 == p1) cur = seed_node;
                else cur = cidx[p0 + __umulhi(rs.next(), p1 - p0)];
                bool seen = false;
                for (int j = 0; j < size; ++j) seen = seen || (chosen[j] == cur);
                if (!seen) chosen[size++] = cur;
            }
            n_iter += (unsigned long long)it;
            if (size == k) {
                for (int j = 0; j < k; ++j) {
                    const int u = chosen[j];
                    for (u32 p = rp[u]; p < rp[u + 1]; ++p) {
                        const int y = cidx[p];
                        bool hit = false;
                        for (int q = 0; q < k; ++q) hit = hit || (chosen[q] == y);
                        cnt += hit ? 1 : 0;
                    }
                }
                for (int j = 0; j < k; ++j) row[j] = lo + chosen[j];
            } else {
                n_fail++;
            }
        }
        if (size != k)
            for (int j = 0; j < k; ++j) row[j] = -1;
        edge_ptr[1 + s] = cnt;
    }
    __syncwarp();
    const unsigned fails = __reduce_add_sync(FULL_MASK, (unsigned)n_fail);
    if (n_iter) atomicAdd(&w.counters[UGS_CTR_TRIALS], n_iter);  // walk iterations
    if ((threadIdx.x & 31) == 0 && fails) atomicAdd(&w.counters[UGS_CTR_INCOMPLETE], (unsigned long long)fails);
}

extern "C" int ugs_rwr_sample_batch(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, uint64_t seed,
                                    uint64_t graph_index_base, double p_restart, void* workspace, const ugs_layout* L,
                                    void* scratch, size_t scratch_bytes, int64_t* nodes, int64_t* edge_ptr,
                                    int64_t* sample_ptr, void* stream_) {
    if (!workspace || !L || !edge_ptr || !sample_ptr) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (k < 1) return ugs_set_error(UGS_EINVAL, "k must be >= 1");
    if (k > UGS_MAX_K) return ugs_set_error(UGS_ELIMIT, "k must be in [1, %d]", UGS_MAX_K);
    if (!(p_restart >= 0.0 && p_restart <= 1.0)) return ugs_set_error(UGS_EINVAL, "p_restart in [0,1]");
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
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    {   // bitmask-path and skipped graphs (also writes sample_ptr for every graph)
        const int64_t ipg = (m + RWR_ITEM - 1) / RWR_ITEM;
        const int64_t items = G * ipg;
        k_rwr_sample_small<<<(unsigned)((items + RWR_WARPS - 1) / RWR_WARPS), RWR_WARPS * 32, 0, st>>>(
            ptr, G, node_base, m, k, seed, graph_index_base, p_restart, ipg, w, nodes, edge_ptr, sample_ptr);
        UGS_LAUNCH_CHECK("k_rwr_sample_small");
    }
    int64_t blocks = (G * m + 127) / 128;
    const int64_t cap = (int64_t)sms * 16;
    if (blocks > cap) blocks = cap;
    k_rwr_sample<<<(unsigned)blocks, 128, 0, st>>>(ptr, G, node_base, m, k, seed, graph_index_base, p_restart, w, nodes,
                                                  edge_ptr, sample_ptr);
    UGS_LAUNCH_CHECK("k_rwr_sample");
    return ugs_run_scan(edge_ptr, G * m, scratch, scratch_bytes, w, st);
}

extern "C" int ugs_rwr_emit_edges(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int mode,
                                  int use_perm, int64_t n_general, const void* workspace, const ugs_layout* L,
                                  const int64_t* nodes, const int64_t* edge_ptr, int64_t total,
                                  int64_t* edge_index_out, int64_t* edge_src_out, void* stream_) {
    // any mode other than "sample" emits batch-global node ids (rwr_sampler.cpp:224-241)
    const int md = (mode == UGS_MODE_SAMPLE) ? UGS_MODE_SAMPLE : UGS_MODE_GLOBAL;
    const int rc = ugs_emit_edges(ptr, G, node_base, m, k, md, use_perm, n_general, workspace, L, nodes, edge_ptr, total,
                                  edge_index_out, edge_src_out, stream_);
    if (rc != UGS_OK) return rc;
    if (total > 0 && edge_src_out)  // src_edge = -1: "not tracked" (:243)
        UGS_CUDA_CHECK(cudaMemsetAsync(edge_src_out, 0xFF, sizeof(i64) * (size_t)total, (cudaStream_t)stream_));
    return UGS_OK;
}
