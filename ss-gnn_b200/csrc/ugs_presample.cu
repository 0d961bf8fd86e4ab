// ugs_presample.cu -- device-resident presample store: re-stitching cached samples into a mini-batch.
//
// Replaces the reference's Python presample cache:
//   Experiment._setup_presampling          gps/experiment.py:379-440   one sampler call per dataset graph
//                                                                      (ptr = [0, n], seed = cfg.seed + index), results
//                                                                      kept in a dict of CPU tensors
//   Experiment._load_from_presample_cache  gps/experiment.py:895-995   per mini-batch: a Python loop over the graphs
//                                                                      with .item() calls, torch.cat of the pieces
// The store is what ONE ugs_sample_batch(seed_mode = 1) call over the whole split returns (kept in HBM); a mini-batch
// is the list of dataset indices `graph_idx` plus the mini-batch's own `ptr` / `edge_index`. Stitching rules
// (experiment.py:935-990), with d = graph_idx[b]:
//   nodes      row (b*m + i) = cached local ids + ptr[b]               (:957-958; the reference also adds ptr[b] to
//                                                                      the -1 padding -- keep_padding = 0 reproduces
//                                                                      that bit for bit, keep_padding = 1 keeps -1)
//   edge_index unchanged local (j, l) entries                          (:960-963)
//   edge_src   cached column (relative to the graph) + orig_off[b], orig_off[b] = number of mini-batch columns whose
//              source node lies in graphs 0..b-1                       (:926-932,966)
//   edge_ptr   running totals, sample_ptr = b*m                        (:973-981)
// All HBM-bound copies with per-entry offset arithmetic: one thread per output element, consecutive threads on
// consecutive elements; the owning slot of an edge entry is found by binary search over the Gb+1 slot offsets.
#include "ugs_common.cuh"

// counts[0..Gb)   = cached edge entries of slot b;  counts[Gb+1 .. 2Gb+1) = mini-batch columns with source in graph b
__global__ void __launch_bounds__(256)
k_presample_count(const i64* __restrict__ s_edge_ptr, i64 Gd, i64 m, const i64* __restrict__ graph_idx, i64 Gb,
                  const i64* __restrict__ ptr, const i64* __restrict__ row0, i64 E, i64* __restrict__ slot_off,
                  unsigned long long* __restrict__ col_cnt, int* __restrict__ bad) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < Gb) {
        const i64 d = graph_idx[t];
        if (d < 0 || d >= Gd) { *bad = 1; slot_off[1 + t] = 0; }
        else slot_off[1 + t] = s_edge_ptr[(d + 1) * m] - s_edge_ptr[d * m];
    }
    if (t < E && Gb > 0) {
        const i64 u = row0[t];
        i64 a = 0, b = Gb;  // largest g with ptr[g] <= u
        while (b - a > 1) { const i64 mid = (a + b) >> 1; if (ptr[mid] <= u) a = mid; else b = mid; }
        if (u >= ptr[a] && u < ptr[a + 1]) atomicAdd(&col_cnt[a], 1ull);
    }
}

// one CTA: in-place inclusive scans of slot_off[1..Gb] and exclusive scan of col_cnt[0..Gb) -> orig_off
__global__ void __launch_bounds__(1024)
k_presample_scan(i64 Gb, i64* __restrict__ slot_off, i64* __restrict__ orig_off, i64* __restrict__ total_out) {
    __shared__ i64 sa[1024], sb[1024];
    __shared__ i64 carry_a, carry_b;
    if (threadIdx.x == 0) { carry_a = 0; carry_b = 0; slot_off[0] = 0; }
    __syncthreads();
    for (i64 base = 0; base < Gb; base += 1024) {
        const i64 i = base + threadIdx.x;
        const i64 va = (i < Gb) ? slot_off[1 + i] : 0, vb = (i < Gb) ? orig_off[i] : 0;
        sa[threadIdx.x] = va; sb[threadIdx.x] = vb;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const i64 ta = (threadIdx.x >= o) ? sa[threadIdx.x - o] : 0, tb = (threadIdx.x >= o) ? sb[threadIdx.x - o] : 0;
            __syncthreads();
            sa[threadIdx.x] += ta; sb[threadIdx.x] += tb;
            __syncthreads();
        }
        if (i < Gb) { slot_off[1 + i] = carry_a + sa[threadIdx.x]; orig_off[i] = carry_b + sb[threadIdx.x] - vb; }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_a += sa[1023]; carry_b += sb[1023]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_a;
}

__global__ void __launch_bounds__(256)
k_presample_rows(const i64* __restrict__ s_nodes, const i64* __restrict__ s_edge_ptr, const i64* __restrict__ s_ptr,
                 i64 m, int k, const i64* __restrict__ graph_idx, i64 Gb, const i64* __restrict__ ptr,
                 const i64* __restrict__ slot_off, int keep_padding, i64* __restrict__ nodes_out,
                 i64* __restrict__ edge_ptr_out, i64* __restrict__ sample_ptr_out) {
    const i64 mk = m * k;
    const i64 total = Gb * mk;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const i64 b = t / mk, r = t - b * mk;
        const i64 d = graph_idx[b];
        const i64 v = s_nodes[d * mk + r];
        const i64 local = (v < 0) ? v : v - s_ptr[d];
        nodes_out[t] = (v < 0 && keep_padding) ? (i64)-1 : local + ptr[b];  // experiment.py:957-958
    }
    const i64 rows = Gb * m;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < rows; t += stride) {
        const i64 b = t / m, i = t - b * m;
        const i64 d = graph_idx[b];
        edge_ptr_out[1 + t] = slot_off[b] + (s_edge_ptr[d * m + i + 1] - s_edge_ptr[d * m]);  // :973-975
        if (i == 0) sample_ptr_out[b + 1] = (b + 1) * m;                                       // :980-981
        if (t == 0) { edge_ptr_out[0] = 0; sample_ptr_out[0] = 0; }
    }
}

__global__ void __launch_bounds__(256)
k_presample_edges(const i64* __restrict__ s_eu, const i64* __restrict__ s_ev, const i64* __restrict__ s_esrc,
                  const i64* __restrict__ s_edge_ptr, const i64* __restrict__ s_cstart, i64 m,
                  const i64* __restrict__ graph_idx, i64 Gb, const i64* __restrict__ slot_off,
                  const i64* __restrict__ orig_off, i64 total, i64* __restrict__ out_u, i64* __restrict__ out_v,
                  i64* __restrict__ out_src) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        i64 a = 0, b = Gb;  // largest slot with slot_off[slot] <= e
        while (b - a > 1) { const i64 mid = (a + b) >> 1; if (slot_off[mid] <= e) a = mid; else b = mid; }
        const i64 d = graph_idx[a];
        const i64 pos = s_edge_ptr[d * m] + (e - slot_off[a]);
        out_u[e] = s_eu[pos];
        out_v[e] = s_ev[pos];
        out_src[e] = s_esrc[pos] - s_cstart[d] + orig_off[a];  // :966
    }
}

static int ps_grid(i64 threads) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const i64 want = (threads + 255) / 256, cap = (i64)sms * 32;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

extern "C" size_t ugs_presample_scratch_bytes(int64_t Gb) {
    return (size_t)(2 * (Gb > 0 ? Gb : 0) + 4) * sizeof(i64);
}

// scratch: slot_off[Gb+1] | orig_off[Gb] | total | bad
extern "C" int ugs_presample_plan(const int64_t* s_edge_ptr, int64_t Gd, int64_t m, const int64_t* graph_idx, int64_t Gb,
                                  const int64_t* ptr, const int64_t* edge_index_row0, int64_t E, void* scratch,
                                  size_t scratch_bytes, int64_t* h_total, void* stream_) {
    if (Gb < 0 || Gd < 0 || m < 0 || E < 0) return ugs_set_error(UGS_EINVAL, "negative size");
    if (!h_total) return ugs_set_error(UGS_EINVAL, "h_total is NULL");
    *h_total = 0;
    if (Gb == 0) return UGS_OK;
    if (!s_edge_ptr || !graph_idx || !ptr || !scratch) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (scratch_bytes < ugs_presample_scratch_bytes(Gb)) return ugs_set_error(UGS_EINVAL, "scratch too small");
    cudaStream_t st = (cudaStream_t)stream_;
    i64* slot_off = (i64*)scratch;
    i64* orig_off = slot_off + Gb + 1;
    i64* total = orig_off + Gb;
    int* bad = (int*)(total + 1);
    UGS_CUDA_CHECK(cudaMemsetAsync(scratch, 0, ugs_presample_scratch_bytes(Gb), st));
    const i64 threads = Gb > E ? Gb : E;
    k_presample_count<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        s_edge_ptr, Gd, m, graph_idx, Gb, ptr, edge_index_row0, E, slot_off, (unsigned long long*)orig_off, bad);
    UGS_LAUNCH_CHECK("k_presample_count");
    k_presample_scan<<<1, 1024, 0, st>>>(Gb, slot_off, orig_off, total);
    UGS_LAUNCH_CHECK("k_presample_scan");
    i64 h[2] = {0, 0};
    UGS_CUDA_CHECK(cudaMemcpyAsync(h, total, 2 * sizeof(i64), cudaMemcpyDeviceToHost, st));
    UGS_CUDA_CHECK(cudaStreamSynchronize(st));
    if ((int)(h[1] & 0xffffffff) != 0) return ugs_set_error(UGS_EINVAL, "graph_idx out of range [0, %lld)", (long long)Gd);
    *h_total = h[0];
    return UGS_OK;
}

extern "C" int ugs_presample_load(const int64_t* s_nodes, const int64_t* s_edge_index, int64_t s_total,
                                  const int64_t* s_edge_ptr, const int64_t* s_edge_src, const int64_t* s_ptr,
                                  const int64_t* s_cstart, int64_t Gd, int64_t m, int k, const int64_t* graph_idx,
                                  int64_t Gb, const int64_t* ptr, const void* scratch, int64_t total, int keep_padding,
                                  int64_t* nodes_out, int64_t* edge_index_out, int64_t* edge_ptr_out,
                                  int64_t* sample_ptr_out, int64_t* edge_src_out, void* stream_) {
    if (k < 1) return ugs_set_error(UGS_EINVAL, "k must be >= 1");
    if (!edge_ptr_out || !sample_ptr_out) return ugs_set_error(UGS_EINVAL, "NULL output");
    cudaStream_t st = (cudaStream_t)stream_;
    if (Gb <= 0 || m <= 0) {
        UGS_CUDA_CHECK(cudaMemsetAsync(edge_ptr_out, 0, sizeof(i64), st));
        UGS_CUDA_CHECK(cudaMemsetAsync(sample_ptr_out, 0, sizeof(i64) * (size_t)((Gb > 0 ? Gb : 0) + 1), st));
        return UGS_OK;
    }
    if (!s_nodes || !s_edge_ptr || !s_ptr || !s_cstart || !graph_idx || !ptr || !scratch || !nodes_out)
        return ugs_set_error(UGS_EINVAL, "NULL argument");
    const i64* slot_off = (const i64*)scratch;
    const i64* orig_off = slot_off + Gb + 1;
    k_presample_rows<<<ps_grid(Gb * m * k), 256, 0, st>>>(s_nodes, s_edge_ptr, s_ptr, m, k, graph_idx, Gb, ptr, slot_off,
                                                         keep_padding, nodes_out, edge_ptr_out, sample_ptr_out);
    UGS_LAUNCH_CHECK("k_presample_rows");
    if (total > 0) {
        if (!s_edge_index || !s_edge_src || !edge_index_out || !edge_src_out) return ugs_set_error(UGS_EINVAL, "NULL edge array");
        k_presample_edges<<<ps_grid(total), 256, 0, st>>>(s_edge_index, s_edge_index + s_total, s_edge_src, s_edge_ptr,
                                                         s_cstart, m, graph_idx, Gb, slot_off, orig_off, total,
                                                         edge_index_out, edge_index_out + total, edge_src_out);
        UGS_LAUNCH_CHECK("k_presample_edges");
    }
    (void)Gd;
    return UGS_OK;
}
