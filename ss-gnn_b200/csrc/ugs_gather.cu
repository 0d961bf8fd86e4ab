// ugs_gather.cu -- the gather/offset step that consumes the sampler's output.
//
// Replaces, for the whole batch at once, the tensor ops at the top of
//   SubgraphSamplingGNNClassifier.encode_subgraphs   gps/models/ss_gnn.py:427-471
//   SubgraphSamplingGNNWithWL.encode_subgraphs_with_gnn   gps/models/ss_gnn_wl.py:249-306
//     global_x          = x[nodes.clamp(min=0)] * (nodes >= 0)          (:443-453)
//     global_edge_attr  = edge_attr[edge_src_global]                    (:455-459)
//     global_edge_index = edge_index_t + repeat_interleave(arange(B), counts) * k   (:461-463)
//     global_batch_ptr  = repeat_interleave(arange(B), k)               (:465-466)
// All of it is HBM-bound byte movement: rows are copied bit for bit with the widest vector the row size and the
// pointer alignment allow, one vector per thread, consecutive threads on consecutive vectors of the output.
#include "ugs_common.cuh"

template <typename V>
__device__ __forceinline__ V zero_vec();
template <> __device__ __forceinline__ uint4 zero_vec<uint4>() { return make_uint4(0u, 0u, 0u, 0u); }
template <> __device__ __forceinline__ uint2 zero_vec<uint2>() { return make_uint2(0u, 0u); }
template <> __device__ __forceinline__ u32 zero_vec<u32>() { return 0u; }
template <> __device__ __forceinline__ unsigned short zero_vec<unsigned short>() { return 0; }
template <> __device__ __forceinline__ unsigned char zero_vec<unsigned char>() { return 0; }

// out[i, :] = index[i] >= 0 ? src[index[i], :] : 0, rows of R vectors of type V.
// IDX32: the flattened vector index fits 32 bits (one 32-bit division per thread instead of a 64-bit one).
template <typename V, bool IDX32>
__global__ void __launch_bounds__(256)
k_gather_rows(const V* __restrict__ src, i64 n_src_rows, u32 R, const i64* __restrict__ index, i64 n_index,
              V* __restrict__ out, unsigned long long* __restrict__ n_bad) {
    const i64 total = n_index * (i64)R;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        i64 row;
        u32 c;
        if (R == 1) { row = t; c = 0; }
        else if (IDX32) { const u32 q = (u32)t / R; row = q; c = (u32)t - q * R; }
        else { row = t / R; c = (u32)(t - row * R); }
        const i64 ix = __ldg(index + row);
        V v = zero_vec<V>();
        if (ix >= 0) {
            if (ix < n_src_rows) v = __ldg(src + ix * (i64)R + c);
            else if (c == 0 && n_bad) atomicAdd(n_bad, 1ull);
        }
        out[t] = v;
    }
}

// global_edge_index for materialised mode-"sample" output: one warp per tile of 32 samples; the tile's entries are
// swept 32 at a time (coalesced) and every lane finds the sample of its entry by a 5-step search over the tile's
// 33 offsets held one per lane. batch_out[i] = i / k is written by the same warps (k values per sample).
__global__ void __launch_bounds__(256)
k_encoder_offsets(const i64* __restrict__ ei_u, const i64* __restrict__ ei_v, const i64* __restrict__ edge_ptr, i64 B,
                  int k, i64 total, i64* __restrict__ out_u, i64* __restrict__ out_v, i64* __restrict__ batch_out) {
    const int lane = threadIdx.x & 31;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const i64 ntiles = (B + 31) >> 5;
    for (i64 tile = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < ntiles; tile += nwarps) {
        const i64 s0 = tile << 5;
        const int nv = (int)min((i64)32, B - s0);
        if (batch_out) {
            for (int e = lane; e < nv * k; e += 32) batch_out[s0 * k + e] = s0 + e / k;
        }
        if (!ei_u) continue;
        const i64 my_lo = edge_ptr[s0 + min(lane, nv)];   // lanes >= nv hold the tile end
        const i64 tile_lo = __shfl_sync(FULL_MASK, my_lo, 0);
        const i64 tile_end = edge_ptr[s0 + nv];
        for (i64 e0 = tile_lo; e0 < tile_end; e0 += 32) {
            const i64 e = e0 + lane;
            // largest lane index s < nv with my_lo[s] <= e  (offsets are non-decreasing)
            int a = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int cand = a + step;
                const i64 off = __shfl_sync(FULL_MASK, my_lo, cand & 31);
                if (cand < nv && off <= e) a = cand;
            }
            if (e < tile_end) {
                const i64 add = (s0 + a) * (i64)k;
                out_u[e] = ei_u[e] + add;
                out_v[e] = ei_v[e] + add;
            }
        }
    }
    (void)total;
}

static int gather_grid(i64 total_threads) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const i64 want = (total_threads + 255) / 256;
    const i64 cap = (i64)sms * 8 * 4;  // 8 resident CTAs of 256 threads per SM, 4 waves of grid-stride work
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

extern "C" int ugs_gather_rows(const void* src, int64_t n_src_rows, int64_t row_bytes, const int64_t* index,
                               int64_t n_index, void* out, int64_t* n_bad, void* stream_) {
    if (row_bytes < 0 || n_index < 0 || n_src_rows < 0) return ugs_set_error(UGS_EINVAL, "negative size");
    if (n_index == 0 || row_bytes == 0) return UGS_OK;
    if (!out || !index) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (!src && n_src_rows > 0) return ugs_set_error(UGS_EINVAL, "src is NULL");
    cudaStream_t st = (cudaStream_t)stream_;
    const uintptr_t al = (uintptr_t)src | (uintptr_t)out | (uintptr_t)row_bytes;
    unsigned long long* nb = (unsigned long long*)n_bad;
#define UGS_GATHER(V_)                                                                                          \
    do {                                                                                                        \
        const i64 R = row_bytes / (i64)sizeof(V_);                                                              \
        if (R >= (1ll << 31)) return ugs_set_error(UGS_ELIMIT, "row too long");                                 \
        const i64 total = n_index * R;                                                                          \
        const int grid = gather_grid(total);                                                                    \
        if (total < (1ll << 32))                                                                                \
            k_gather_rows<V_, true><<<grid, 256, 0, st>>>((const V_*)src, n_src_rows, (u32)R, index, n_index,  \
                                                          (V_*)out, nb);                                        \
        else                                                                                                    \
            k_gather_rows<V_, false><<<grid, 256, 0, st>>>((const V_*)src, n_src_rows, (u32)R, index, n_index, \
                                                           (V_*)out, nb);                                       \
    } while (0)
    if ((al & 15) == 0) UGS_GATHER(uint4);
    else if ((al & 7) == 0) UGS_GATHER(uint2);
    else if ((al & 3) == 0) UGS_GATHER(u32);
    else if ((al & 1) == 0) UGS_GATHER(unsigned short);
    else UGS_GATHER(unsigned char);
#undef UGS_GATHER
    UGS_LAUNCH_CHECK("k_gather_rows");
    return UGS_OK;
}

extern "C" int ugs_encoder_offsets(const int64_t* edge_index_s, const int64_t* edge_ptr, int64_t B, int k,
                                   int64_t total, int64_t* global_edge_index_out, int64_t* batch_out, void* stream_) {
    if (B < 0 || total < 0) return ugs_set_error(UGS_EINVAL, "negative size");
    if (k < 1) return ugs_set_error(UGS_EINVAL, "k must be >= 1");
    if (B == 0) return UGS_OK;
    const bool edges = edge_index_s != nullptr && total > 0;
    if (edges && (!edge_ptr || !global_edge_index_out)) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (!edges && !batch_out) return UGS_OK;
    cudaStream_t st = (cudaStream_t)stream_;
    const i64 ntiles = (B + 31) / 32;
    const int grid = gather_grid(ntiles * 32);
    k_encoder_offsets<<<grid, 256, 0, st>>>(edges ? edge_index_s : nullptr, edges ? edge_index_s + total : nullptr,
                                            edge_ptr, B, k, total, global_edge_index_out,
                                            edges ? global_edge_index_out + total : nullptr, batch_out);
    UGS_LAUNCH_CHECK("k_encoder_offsets");
    return UGS_OK;
}
