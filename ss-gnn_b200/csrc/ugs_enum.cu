// ugs_enum.cu -- exactly uniform connected k-subgraph sampler by enumeration on sm_100a.
//
// Replaces the reference's src/samplers/uniform_sampler/src/uniform_sampler.cpp:
//   enumerate_subgraphs :47-81    every k-combination of the n vertices, kept if it induces a connected subgraph
//                                 (is_connected :17-42): O(C(n,k)) per graph
//   sample_batch        :86-282   per sample one enumerated subgraph picked uniformly (:196-199), node row in
//                                 ascending vertex order; graphs without a connected k-subgraph (n < k, ...) give all -1
//                                 rows (:183-191); edges = every batch column with both endpoints in the sample, in
//                                 column order, src = the column (:216-245)  == ugs_eps_emit_edges
// Here the connected k-subsets are enumerated directly, each exactly once, by the ESU recursion (extend a set only by
// vertices larger than its minimum that are exclusive neighbours of the vertex added last) on 64-bit adjacency masks:
// work is proportional to the number of connected subsets, not to C(n,k). One warp per graph, one root (= minimum
// vertex) per lane, two passes (count, then fill after a scan of the counts), so the list of graph g is ordered by
// root and, per root, by the depth-first order of the recursion -- the same order oracle/walk_oracle.c produces.
// A sample is list[graph][mulhi64(word64, count)] with word64 from Philox4x32-10 keyed (seed; sample i, graph,
// 0xE0000000). Bitmask-path graphs only (<= 64 vertices, <= 384 columns): enumeration is exponential in k anyway.
#include "ugs_common.cuh"

int ugs_run_scan(i64* edge_ptr, i64 B, void* scratch, size_t scratch_bytes, WS& w, cudaStream_t st);  // ugs_sample.cu

#define ENUM_MAX_K 16
#define ENUM_WARPS 4

// All connected k-subsets whose minimum vertex is v. Returns their number; FILL writes them to out[0..).
template <bool FILL>
__device__ __forceinline__ u64 esu_root(const u64* __restrict__ adj, int v, int k, u64* __restrict__ out) {
    const u64 bv = 1ull << v;
    if (k == 1) { if (FILL) out[0] = bv; return 1; }
    const u64 gt = (v >= 63) ? 0ull : (~0ull << (v + 1));
    u64 S[ENUM_MAX_K], EXT[ENUM_MAX_K], NS[ENUM_MAX_K];  // level d: a connected set of d+1 vertices
    int d = 0;
    S[0] = bv; NS[0] = adj[v] | bv; EXT[0] = adj[v] & gt;
    u64 cnt = 0;
    while (d >= 0) {
        u64 e = EXT[d];
        if (d == k - 2) {  // every extension vertex completes one subset
            if (FILL) {
                while (e) { const u64 low = e & (~e + 1ull); e ^= low; out[cnt++] = S[d] | low; }
            } else {
                cnt += (u64)__popcll(e);
            }
            --d;
            continue;
        }
        if (e == 0) { --d; continue; }
        const int wv = __ffsll((long long)e) - 1;
        e &= e - 1;
        EXT[d] = e;
        const u64 aw = adj[wv];
        S[d + 1] = S[d] | (1ull << wv);
        EXT[d + 1] = e | (aw & gt & ~NS[d]);  // exclusive neighbours of wv above the root
        NS[d + 1] = NS[d] | aw;
        ++d;
    }
    return cnt;
}

// scratch layout: root_cnt[NT] (i64) | graph_off[G+1] (i64; [0] = 0, [1+g] = count of graph g, scanned in place)
template <bool FILL>
__global__ void __launch_bounds__(ENUM_WARPS * 32)
k_enum_subsets(const i64* __restrict__ ptr, i64 G, i64 node_base, int k, WS w, i64* __restrict__ root_cnt,
               i64* __restrict__ graph_off, u64* __restrict__ list) {
    __shared__ u64 adj_all[ENUM_WARPS][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    u64* adj = adj_all[warp];
    const i64 nwarps = (i64)gridDim.x * ENUM_WARPS;
    for (i64 g = (i64)blockIdx.x * ENUM_WARPS + warp; g < G; g += nwarps) {
        const ugs_ginfo gi = w.ginfo[g];
        if (!FILL && lane == 0 && g == 0) graph_off[0] = 0;
        if (gi.cls != 1) {
            if (!FILL && lane == 0) graph_off[1 + g] = 0;
            continue;
        }
        const i64 lo = ptr[g];
        const int n = (int)(ptr[g + 1] - lo);
        const i64 nb = lo - node_base;
        __syncwarp();
        for (int v = lane; v < 64; v += 32) adj[v] = (v < n) ? (w.adj[nb + v] & ~(1ull << v)) : 0ull;
        __syncwarp();
        if (!FILL) {
            u64 tot = 0;
            for (int v = lane; v < n; v += 32) {
                const u64 c = esu_root<false>(adj, v, k, nullptr);
                root_cnt[nb + v] = (i64)c;
                tot += c;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL_MASK, tot, o);
            if (lane == 0) graph_off[1 + g] = (i64)tot;
        } else {
            // exclusive offsets of the roots inside the graph's list: roots 0..31 on the lanes, then 32..63
            i64 base = graph_off[g];
            for (int h = 0; h < 2; ++h) {
                const int v = lane + 32 * h;
                const i64 c = (v < n) ? root_cnt[nb + v] : 0;
                i64 incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const i64 t = __shfl_up_sync(FULL_MASK, incl, o);
                    if (lane >= o) incl += t;
                }
                if (c > 0) esu_root<true>(adj, v, k, list + base + (incl - c));
                base += __shfl_sync(FULL_MASK, incl, 31);
            }
        }
    }
}

__global__ void __launch_bounds__(128)
k_enum_sample(const i64* __restrict__ ptr, i64 G, i64 node_base, i64 m, int k, u64 seed, u64 gbase, WS w,
              const i64* __restrict__ graph_off, const u64* __restrict__ list, i64* __restrict__ nodes_out,
              i64* __restrict__ edge_ptr, i64* __restrict__ sample_ptr) {
    const i64 total = G * m;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    unsigned n_fail = 0;
    for (i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += stride) {
        const i64 g = s / m;
        const i64 i = s - g * m;
        if (i == 0) {
            sample_ptr[g + 1] = (g + 1) * m;  // uniform_sampler.cpp:190,251 (every graph contributes m rows)
            if (g == 0) { sample_ptr[0] = 0; edge_ptr[0] = 0; }
        }
        const i64 base = graph_off[g];
        const u64 c = (u64)(graph_off[g + 1] - base);
        i64* row = nodes_out + s * k;
        i64 cnt = 0;
        if (c == 0) {
            for (int j = 0; j < k; ++j) row[j] = -1;
            n_fail++;
        } else {
            const u64 gidx = gbase + (u64)g;
            const Philox4 r = philox4x32_10((u32)i, 0u, (u32)gidx, 0xE0000000u, (u32)seed, (u32)(seed >> 32));
            const u64 idx = __umul64hi((((u64)r.x) << 32) | (u64)r.y, c);
            u64 S = list[base + idx];
            const i64 lo = ptr[g];
            const u32* __restrict__ rp = w.rowptr + (lo - node_base) + g;
            const i32* __restrict__ cidx = w.csr_idx + 2 * w.cstart[g];
            const u64 mask = S;
            for (int j = 0; j < k; ++j) {
                const int v = __ffsll((long long)S) - 1;
                S &= S - 1;
                row[j] = lo + v;
                for (u32 p = rp[v]; p < rp[v + 1]; ++p) cnt += (i64)((mask >> cidx[p]) & 1ull);
            }
            cnt >>= 1;  // two CSR entries per column inside the sample
        }
        edge_ptr[1 + s] = cnt;
    }
    __syncwarp();
    n_fail = __reduce_add_sync(FULL_MASK, n_fail);
    if ((threadIdx.x & 31) == 0 && n_fail) atomicAdd(&w.counters[UGS_CTR_INCOMPLETE], (unsigned long long)n_fail);
}

static int enum_sms() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

extern "C" size_t ugs_enum_scratch_bytes(int64_t G, int64_t NT) {
    const int64_t g = G > 0 ? G : 0, n = NT > 0 ? NT : 0;
    return (size_t)(n + g + 1) * sizeof(i64) + ugs_sample_scratch_bytes(g, 1) + 64;
}

static inline void enum_views(void* scratch, int64_t G, int64_t NT, i64** root_cnt, i64** graph_off, void** scan) {
    i64* b = (i64*)scratch;
    *root_cnt = b;
    *graph_off = b + (NT > 0 ? NT : 0);
    *scan = (void*)(*graph_off + (G > 0 ? G : 0) + 1);
}

extern "C" int ugs_enum_count(const int64_t* ptr, int64_t G, int64_t NT, int64_t node_base, int k, int64_t n_general,
                              void* workspace, const ugs_layout* L, void* enum_scratch, size_t enum_scratch_bytes,
                              void* stream_) {
    if (!workspace || !L || !enum_scratch) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (k < 1 || k > ENUM_MAX_K) return ugs_set_error(UGS_ELIMIT, "uniform enumeration supports k in [1, %d]", ENUM_MAX_K);
    if (n_general > 0)
        return ugs_set_error(UGS_ELIMIT, "uniform enumeration supports graphs with <= %d nodes and <= %d edge columns",
                             UGS_SMALL_MAX_N, UGS_SMALL_MAX_COLS);
    if (enum_scratch_bytes < ugs_enum_scratch_bytes(G, NT)) return ugs_set_error(UGS_EINVAL, "enum scratch too small");
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(workspace, L);
    i64 *root_cnt, *graph_off; void* scan;
    enum_views(enum_scratch, G, NT, &root_cnt, &graph_off, &scan);
    UGS_CUDA_CHECK(cudaMemsetAsync(&w.counters[UGS_CTR_TOTAL_EDGES], 0, sizeof(i64), st));
    UGS_CUDA_CHECK(cudaMemsetAsync(graph_off, 0, sizeof(i64), st));
    if (G <= 0) return UGS_OK;
    int64_t blocks = (G + ENUM_WARPS - 1) / ENUM_WARPS;
    const int64_t cap = (int64_t)enum_sms() * 16;
    if (blocks > cap) blocks = cap;
    k_enum_subsets<false><<<(unsigned)blocks, ENUM_WARPS * 32, 0, st>>>(ptr, G, node_base, k, w, root_cnt, graph_off, nullptr);
    UGS_LAUNCH_CHECK("k_enum_subsets<count>");
    return ugs_run_scan(graph_off, G, scan, ugs_sample_scratch_bytes(G, 1), w, st);  // total -> counters[TOTAL_EDGES]
}

extern "C" int ugs_enum_fill(const int64_t* ptr, int64_t G, int64_t NT, int64_t node_base, int k, void* workspace,
                             const ugs_layout* L, void* enum_scratch, uint64_t* list, int64_t total, void* stream_) {
    if (!workspace || !L || !enum_scratch) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (k < 1 || k > ENUM_MAX_K) return ugs_set_error(UGS_ELIMIT, "uniform enumeration supports k in [1, %d]", ENUM_MAX_K);
    if (G <= 0 || total <= 0) return UGS_OK;
    if (!list) return ugs_set_error(UGS_EINVAL, "list is NULL");
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(workspace, L);
    i64 *root_cnt, *graph_off; void* scan;
    enum_views(enum_scratch, G, NT, &root_cnt, &graph_off, &scan);
    int64_t blocks = (G + ENUM_WARPS - 1) / ENUM_WARPS;
    const int64_t cap = (int64_t)enum_sms() * 16;
    if (blocks > cap) blocks = cap;
    k_enum_subsets<true><<<(unsigned)blocks, ENUM_WARPS * 32, 0, st>>>(ptr, G, node_base, k, w, root_cnt, graph_off,
                                                                       (u64*)list);
    UGS_LAUNCH_CHECK("k_enum_subsets<fill>");
    return UGS_OK;
}

extern "C" int ugs_enum_sample_batch(const int64_t* ptr, int64_t G, int64_t NT, int64_t node_base, int64_t m, int k,
                                     uint64_t seed, uint64_t graph_index_base, void* workspace, const ugs_layout* L,
                                     void* enum_scratch, const uint64_t* list, void* scratch, size_t scratch_bytes,
                                     int64_t* nodes, int64_t* edge_ptr, int64_t* sample_ptr, void* stream_) {
    if (!workspace || !L || !enum_scratch || !edge_ptr || !sample_ptr) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (k < 1 || k > ENUM_MAX_K) return ugs_set_error(UGS_ELIMIT, "uniform enumeration supports k in [1, %d]", ENUM_MAX_K);
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
    i64 *root_cnt, *graph_off; void* scan;
    enum_views(enum_scratch, G, NT, &root_cnt, &graph_off, &scan);
    int64_t blocks = (G * m + 127) / 128;
    const int64_t cap = (int64_t)enum_sms() * 16;
    if (blocks > cap) blocks = cap;
    k_enum_sample<<<(unsigned)blocks, 128, 0, st>>>(ptr, G, node_base, m, k, seed, graph_index_base, w, graph_off,
                                                   (const u64*)list, nodes, edge_ptr, sample_ptr);
    UGS_LAUNCH_CHECK("k_enum_sample");
    return ugs_run_scan(edge_ptr, G * m, scratch, scratch_bytes, w, st);
}
