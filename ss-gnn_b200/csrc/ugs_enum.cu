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
// 0xE0000000). Graphs with at most 64 vertices use 64-bit masks (one warp per graph, one root per lane; k <= 16);
// larger graphs (up to 2048 vertices, k <= 8) use multi-word bitsets, one thread per root, and store a subset as k
// 16-bit vertex ids. Beyond that the call fails: the reference's own C(n,k) scan is hopeless there as well.
#include "ugs_common.cuh"

int ugs_run_scan(i64* edge_ptr, i64 B, void* scratch, size_t scratch_bytes, WS& w, cudaStream_t st);  // ugs_sample.cu

#define ENUM_MAX_K 16
#define ENUM_WARPS 4

// All connected k-subsets whose minimum vertex is v. Returns their number; FILL writes them to out[0..).
template <bool FILL>
__device__ __forceinline__ u64 esu_root(const u64* __restrict__ adj, int v, int k, u64* __restrict__ out, int rw) {
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
                while (e) { const u64 low = e & (~e + 1ull); e ^= low; out[(cnt++) * rw] = S[d] | low; }
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
               i64* __restrict__ graph_off, u64* __restrict__ list, int rw) {
    __shared__ u64 adj_all[ENUM_WARPS][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    u64* adj = adj_all[warp];
    const i64 nwarps = (i64)gridDim.x * ENUM_WARPS;
    for (i64 g = (i64)blockIdx.x * ENUM_WARPS + warp; g < G; g += nwarps) {
        const ugs_ginfo gi = w.ginfo[g];
        if (!FILL && lane == 0 && g == 0) graph_off[0] = 0;
        const i64 lo = ptr[g];
        const int n = (int)(ptr[g + 1] - lo);
        if (gi.cls == 0 || n > 64) {  // skipped graphs; graphs above 64 vertices belong to k_enum_big
            if (!FILL && lane == 0) graph_off[1 + g] = 0;
            continue;
        }
        const i64 nb = lo - node_base;
        __syncwarp();
        if (gi.cls == 1) {
            for (int v = lane; v < 64; v += 32) adj[v] = (v < n) ? (w.adj[nb + v] & ~(1ull << v)) : 0ull;
        } else {  // <= 64 vertices but too many columns for the bitmask path: masks from the CSR rows
            const u32* rp = w.rowptr + nb + g;
            const i32* cidx = w.csr_idx + 2 * w.cstart[g];
            for (int v = lane; v < 64; v += 32) {
                u64 a = 0;
                if (v < n)
                    for (u32 p = rp[v]; p < rp[v + 1]; ++p) a |= 1ull << cidx[p];
                adj[v] = a & ~(1ull << v);
            }
        }
        __syncwarp();
        if (!FILL) {
            u64 tot = 0;
            for (int v = lane; v < n; v += 32) {
                const u64 c = esu_root<false>(adj, v, k, nullptr, 1);
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
                if (c > 0) esu_root<true>(adj, v, k, list + (base + (incl - c)) * rw, rw);
                base += __shfl_sync(FULL_MASK, incl, 31);
            }
        }
    }
}

// ---- graphs above 64 vertices: multi-word bitsets ------------------------------------------------------------------
#define BIG_MAX_W 64  // 32-bit words per bitset: up to 2048 vertices
#define BIG_MAX_K 8

// adjacency bitsets of the big graphs: graph slot gi (position in glist) at adjw + gi * max_n * W, row v at + v * W
__global__ void __launch_bounds__(256)
k_enum_adj_big(const i64* __restrict__ ptr, i64 node_base, i64 n_general, int max_n, int W, WS w, u32* __restrict__ adjw) {
    const i64 total = n_general * (i64)max_n;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const i64 gi = t / max_n;
        const int v = (int)(t - gi * max_n);
        const i64 g = w.glist[gi];
        const i64 lo = ptr[g];
        const int n = (int)(ptr[g + 1] - lo);
        if (n <= 64 || v >= n) continue;
        const u32* rp = w.rowptr + (lo - node_base) + g;
        const i32* cidx = w.csr_idx + 2 * w.cstart[g];
        u32* row = adjw + (gi * max_n + v) * (i64)W;
        for (u32 p = rp[v]; p < rp[v + 1]; ++p) {
            const int u = cidx[p];
            if (u != v) row[u >> 5] |= 1u << (u & 31);
        }
    }
}

// One thread per (big graph, root): the ESU recursion on W-word bitsets kept in local memory. FILL writes every subset
// as k 16-bit ids (in the order they joined the set; the sampler sorts them) into the 16-byte record list.
template <bool FILL>
__global__ void __launch_bounds__(128)
k_enum_big(const i64* __restrict__ ptr, i64 node_base, i64 n_general, int max_n, int W, int k, WS w,
           const u32* __restrict__ adjw, i64* __restrict__ root_cnt, i64* __restrict__ graph_off, u64* __restrict__ list) {
    const i64 total = n_general * (i64)max_n;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const i64 gi = t / max_n;
        const int v = (int)(t - gi * max_n);
        const i64 g = w.glist[gi];
        const i64 lo = ptr[g];
        const int n = (int)(ptr[g + 1] - lo);
        if (n <= 64 || v >= n) continue;
        const i64 nb = lo - node_base;
        const u32* adj = adjw + gi * (i64)max_n * W;
        const int nw = (n + 31) >> 5;
        unsigned short* out = nullptr;
        if (FILL) {
            // root_cnt holds exclusive offsets inside the graph's list by now (k_enum_big_offsets)
            out = reinterpret_cast<unsigned short*>(list + 2 * (graph_off[g] + root_cnt[nb + v]));
        }
        u64 cnt = 0;
        if (k == 1) {
            if (FILL) out[0] = (unsigned short)v;
            cnt = 1;
        } else {
            u32 EXT[BIG_MAX_K - 1][BIG_MAX_W], NS[BIG_MAX_K - 1][BIG_MAX_W];
            unsigned short S[BIG_MAX_K];
            const int vw = v >> 5;
            auto gt = [&](int wi) -> u32 { return wi < vw ? 0u : (wi > vw ? ~0u : (((v & 31) == 31) ? 0u : (~0u << ((v & 31) + 1)))); };
            int d = 0;
            S[0] = (unsigned short)v;
            for (int wi = 0; wi < nw; ++wi) {
                const u32 a = adj[(i64)v * W + wi];
                NS[0][wi] = a | (wi == vw ? (1u << (v & 31)) : 0u);
                EXT[0][wi] = a & gt(wi);
            }
            while (d >= 0) {
                if (d == k - 2) {  // every extension vertex completes one subset
                    for (int wi = 0; wi < nw; ++wi) {
                        u32 e = EXT[d][wi];
                        if constexpr (!FILL) {
                            cnt += (u64)__popc(e);
                        } else {
                            while (e) {
                                const int b = __ffs((int)e) - 1;
                                e &= e - 1;
                                unsigned short* rec = out + cnt * 8;
                                for (int j = 0; j <= d; ++j) rec[j] = S[j];
                                rec[d + 1] = (unsigned short)(wi * 32 + b);
                                ++cnt;
                            }
                        }
                    }
                    --d;
                    continue;
                }
                int wv = -1;
                for (int wi = 0; wi < nw; ++wi) {
                    const u32 e = EXT[d][wi];
                    if (e) { const int b = __ffs((int)e) - 1; EXT[d][wi] = e & (e - 1); wv = wi * 32 + b; break; }
                }
                if (wv < 0) { --d; continue; }
                S[d + 1] = (unsigned short)wv;
                for (int wi = 0; wi < nw; ++wi) {
                    const u32 a = adj[(i64)wv * W + wi];
                    EXT[d + 1][wi] = EXT[d][wi] | (a & gt(wi) & ~NS[d][wi]);
                    NS[d + 1][wi] = NS[d][wi] | a;
                }
                ++d;
            }
        }
        if (!FILL) {
            root_cnt[nb + v] = (i64)cnt;
            if (cnt) atomicAdd((unsigned long long*)&graph_off[1 + g], (unsigned long long)cnt);
        }
    }
}

// per big graph: root_cnt[nb .. nb+n) -> exclusive offsets inside the graph's list (one CTA per graph)
__global__ void __launch_bounds__(1024)
k_enum_big_offsets(const i64* __restrict__ ptr, i64 node_base, WS w, i64* __restrict__ root_cnt) {
    __shared__ i64 sm[1024];
    __shared__ i64 carry;
    const i64 g = w.glist[blockIdx.x];
    const i64 lo = ptr[g];
    const int n = (int)(ptr[g + 1] - lo);
    if (n <= 64) return;
    i64* rc = root_cnt + (lo - node_base);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const i64 c = (i < n) ? rc[i] : 0;
        sm[threadIdx.x] = c;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const i64 tv = (threadIdx.x >= o) ? sm[threadIdx.x - o] : 0;
            __syncthreads();
            sm[threadIdx.x] += tv;
            __syncthreads();
        }
        if (i < n) rc[i] = carry + sm[threadIdx.x] - c;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sm[1023];
        __syncthreads();
    }
}

// rw = 64-bit words per list record: 1 (every graph has <= 64 vertices: a vertex mask) or 2 (a batch with bigger
// graphs: their records hold k 16-bit ids, small graphs keep the mask in word 0)
__global__ void __launch_bounds__(128)
k_enum_sample(const i64* __restrict__ ptr, i64 G, i64 node_base, i64 m, int k, u64 seed, u64 gbase, WS w,
              const i64* __restrict__ graph_off, const u64* __restrict__ list, int rw, i64* __restrict__ nodes_out,
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
            const i64 lo = ptr[g];
            const int n = (int)(ptr[g + 1] - lo);
            const u32* __restrict__ rp = w.rowptr + (lo - node_base) + g;
            const i32* __restrict__ cidx = w.csr_idx + 2 * w.cstart[g];
            if (n <= 64) {
                u64 S = list[(base + idx) * rw];
                const u64 mask = S;
                for (int j = 0; j < k; ++j) {
                    const int v = __ffsll((long long)S) - 1;
                    S &= S - 1;
                    row[j] = lo + v;
                    for (u32 p = rp[v]; p < rp[v + 1]; ++p) cnt += (i64)((mask >> cidx[p]) & 1ull);
                }
            } else {
                const unsigned short* rec = reinterpret_cast<const unsigned short*>(list + (base + idx) * rw);
                int ids[BIG_MAX_K];
                for (int j = 0; j < k; ++j) {  // ascending order like the reference's combinations
                    const int v = rec[j];
                    int p = j;
                    while (p > 0 && ids[p - 1] > v) { ids[p] = ids[p - 1]; --p; }
                    ids[p] = v;
                }
                for (int j = 0; j < k; ++j) {
                    const int v = ids[j];
                    row[j] = lo + v;
                    for (u32 p = rp[v]; p < rp[v + 1]; ++p) {
                        const int y = cidx[p];
                        bool hit = false;
                        for (int q = 0; q < k; ++q) hit = hit || (ids[q] == y);
                        cnt += hit ? 1 : 0;
                    }
                }
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

static inline bool enum_has_big(int64_t n_general, int64_t max_n_general) { return n_general > 0 && max_n_general > 64; }
static inline int enum_words(int64_t max_n_general) { return (int)((max_n_general + 31) / 32); }
static inline size_t enum_adj_bytes(int64_t n_general, int64_t max_n_general) {
    if (!enum_has_big(n_general, max_n_general)) return 0;
    return (size_t)n_general * (size_t)max_n_general * (size_t)enum_words(max_n_general) * sizeof(u32);
}
static inline size_t enum_fixed_bytes(int64_t G, int64_t NT) {
    const int64_t g = G > 0 ? G : 0, n = NT > 0 ? NT : 0;
    return (((size_t)(n + g + 2) * sizeof(i64) + ugs_sample_scratch_bytes(g, 1)) + 255) & ~(size_t)255;
}

extern "C" int ugs_enum_record_words(int64_t n_general, int64_t max_n_general) {
    return enum_has_big(n_general, max_n_general) ? 2 : 1;
}

extern "C" size_t ugs_enum_scratch_bytes(int64_t G, int64_t NT, int64_t n_general, int64_t max_n_general) {
    return enum_fixed_bytes(G, NT) + enum_adj_bytes(n_general, max_n_general) + 64;
}

static inline void enum_views(void* scratch, int64_t G, int64_t NT, i64** root_cnt, i64** graph_off, void** scan,
                              u32** adjw) {
    i64* b = (i64*)scratch;
    *root_cnt = b;                                        // NT + 1 entries
    *graph_off = b + (NT > 0 ? NT : 0) + 1;               // G + 1 entries
    *scan = (void*)(*graph_off + (G > 0 ? G : 0) + 1);
    *adjw = (u32*)((char*)scratch + enum_fixed_bytes(G, NT));
}

static int enum_check(int k, int64_t n_general, int64_t max_n_general) {
    if (k < 1 || k > ENUM_MAX_K) return ugs_set_error(UGS_ELIMIT, "uniform enumeration supports k in [1, %d]", ENUM_MAX_K);
    if (enum_has_big(n_general, max_n_general)) {
        if (k > BIG_MAX_K)
            return ugs_set_error(UGS_ELIMIT, "uniform enumeration supports k <= %d for graphs with more than 64 nodes", BIG_MAX_K);
        if (max_n_general > 32 * BIG_MAX_W)
            return ugs_set_error(UGS_ELIMIT, "uniform enumeration supports graphs with <= %d nodes", 32 * BIG_MAX_W);
    }
    return UGS_OK;
}

extern "C" int ugs_enum_count(const int64_t* ptr, int64_t G, int64_t NT, int64_t node_base, int k, int64_t n_general,
                              int64_t max_n_general, void* workspace, const ugs_layout* L, void* enum_scratch,
                              size_t enum_scratch_bytes, void* stream_) {
    if (!workspace || !L || !enum_scratch) return ugs_set_error(UGS_EINVAL, "NULL argument");
    { const int rc = enum_check(k, n_general, max_n_general); if (rc != UGS_OK) return rc; }
    if (enum_scratch_bytes < ugs_enum_scratch_bytes(G, NT, n_general, max_n_general))
        return ugs_set_error(UGS_EINVAL, "enum scratch too small");
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(workspace, L);
    i64 *root_cnt, *graph_off; void* scan; u32* adjw;
    enum_views(enum_scratch, G, NT, &root_cnt, &graph_off, &scan, &adjw);
    UGS_CUDA_CHECK(cudaMemsetAsync(&w.counters[UGS_CTR_TOTAL_EDGES], 0, sizeof(i64), st));
    UGS_CUDA_CHECK(cudaMemsetAsync(graph_off, 0, sizeof(i64), st));
    if (G <= 0) return UGS_OK;
    int64_t blocks = (G + ENUM_WARPS - 1) / ENUM_WARPS;
    const int64_t cap = (int64_t)enum_sms() * 16;
    if (blocks > cap) blocks = cap;
    k_enum_subsets<false><<<(unsigned)blocks, ENUM_WARPS * 32, 0, st>>>(ptr, G, node_base, k, w, root_cnt, graph_off, nullptr, 1);
    UGS_LAUNCH_CHECK("k_enum_subsets<count>");
    if (enum_has_big(n_general, max_n_general)) {
        const int W = enum_words(max_n_general);
        UGS_CUDA_CHECK(cudaMemsetAsync(adjw, 0, enum_adj_bytes(n_general, max_n_general), st));
        int64_t tb = (n_general * max_n_general + 255) / 256;
        if (tb > cap * 4) tb = cap * 4;
        k_enum_adj_big<<<(unsigned)tb, 256, 0, st>>>(ptr, node_base, n_general, (int)max_n_general, W, w, adjw);
        UGS_LAUNCH_CHECK("k_enum_adj_big");
        int64_t eb = (n_general * max_n_general + 127) / 128;
        if (eb > cap * 4) eb = cap * 4;
        k_enum_big<false><<<(unsigned)eb, 128, 0, st>>>(ptr, node_base, n_general, (int)max_n_general, W, k, w, adjw,
                                                       root_cnt, graph_off, nullptr);
        UGS_LAUNCH_CHECK("k_enum_big<count>");
        k_enum_big_offsets<<<(unsigned)n_general, 1024, 0, st>>>(ptr, node_base, w, root_cnt);
        UGS_LAUNCH_CHECK("k_enum_big_offsets");
    }
    return ugs_run_scan(graph_off, G, scan, ugs_sample_scratch_bytes(G, 1), w, st);  // total -> counters[TOTAL_EDGES]
}

extern "C" int ugs_enum_fill(const int64_t* ptr, int64_t G, int64_t NT, int64_t node_base, int k, int64_t n_general,
                             int64_t max_n_general, void* workspace, const ugs_layout* L, void* enum_scratch,
                             uint64_t* list, int64_t total, void* stream_) {
    if (!workspace || !L || !enum_scratch) return ugs_set_error(UGS_EINVAL, "NULL argument");
    { const int rc = enum_check(k, n_general, max_n_general); if (rc != UGS_OK) return rc; }
    if (G <= 0 || total <= 0) return UGS_OK;
    if (!list) return ugs_set_error(UGS_EINVAL, "list is NULL");
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(workspace, L);
    i64 *root_cnt, *graph_off; void* scan; u32* adjw;
    enum_views(enum_scratch, G, NT, &root_cnt, &graph_off, &scan, &adjw);
    const int rw = ugs_enum_record_words(n_general, max_n_general);
    if (rw == 2) UGS_CUDA_CHECK(cudaMemsetAsync(list, 0, (size_t)total * 16, st));
    int64_t blocks = (G + ENUM_WARPS - 1) / ENUM_WARPS;
    const int64_t cap = (int64_t)enum_sms() * 16;
    if (blocks > cap) blocks = cap;
    k_enum_subsets<true><<<(unsigned)blocks, ENUM_WARPS * 32, 0, st>>>(ptr, G, node_base, k, w, root_cnt, graph_off,
                                                                       (u64*)list, rw);
    UGS_LAUNCH_CHECK("k_enum_subsets<fill>");
    if (rw == 2) {
        const int W = enum_words(max_n_general);
        int64_t eb = (n_general * max_n_general + 127) / 128;
        if (eb > cap * 4) eb = cap * 4;
        k_enum_big<true><<<(unsigned)eb, 128, 0, st>>>(ptr, node_base, n_general, (int)max_n_general, W, k, w, adjw,
                                                      root_cnt, graph_off, (u64*)list);
        UGS_LAUNCH_CHECK("k_enum_big<fill>");
    }
    return UGS_OK;
}

extern "C" int ugs_enum_sample_batch(const int64_t* ptr, int64_t G, int64_t NT, int64_t node_base, int64_t m, int k,
                                     uint64_t seed, uint64_t graph_index_base, int64_t n_general, int64_t max_n_general,
                                     void* workspace, const ugs_layout* L, void* enum_scratch, const uint64_t* list,
                                     void* scratch, size_t scratch_bytes, int64_t* nodes, int64_t* edge_ptr,
                                     int64_t* sample_ptr, void* stream_) {
    if (!workspace || !L || !enum_scratch || !edge_ptr || !sample_ptr) return ugs_set_error(UGS_EINVAL, "NULL argument");
    { const int rc = enum_check(k, n_general, max_n_general); if (rc != UGS_OK) return rc; }
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
    i64 *root_cnt, *graph_off; void* scan; u32* adjw;
    enum_views(enum_scratch, G, NT, &root_cnt, &graph_off, &scan, &adjw);
    int64_t blocks = (G * m + 127) / 128;
    const int64_t cap = (int64_t)enum_sms() * 16;
    if (blocks > cap) blocks = cap;
    k_enum_sample<<<(unsigned)blocks, 128, 0, st>>>(ptr, G, node_base, m, k, seed, graph_index_base, w, graph_off,
                                                   (const u64*)list, ugs_enum_record_words(n_general, max_n_general),
                                                   nodes, edge_ptr, sample_ptr);
    UGS_LAUNCH_CHECK("k_enum_sample");
    return ugs_run_scan(edge_ptr, G * m, scratch, scratch_bytes, w, st);
}
