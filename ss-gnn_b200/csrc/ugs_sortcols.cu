// ugs_sortcols.cu -- stable ordering of the batch's edge columns by graph id (LSD radix sort, 8-bit digits).
//
// Only used on the slow path of the segmentation (A.1) when the columns are not grouped by graph -- never for a PyG
// batch. It replaces the reference's per-graph rescan of all E columns (ugs_sampler_batch_extension.cpp:41-75):
// the columns are ordered once, stably (original column order kept inside every graph), and read through `perm`.
// Per pass: per-CTA digit histograms -> one exclusive scan of the (digit, CTA) table (k_scan_counts) -> stable
// scatter (rank of a key among the equal digits before it in its CTA tile, by warp match_any + per-warp tables).
#include "ugs_common.cuh"

int ugs_run_scan(i64* edge_ptr, i64 B, void* scratch, size_t scratch_bytes, WS& w, cudaStream_t st);  // ugs_sample.cu

#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)

__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const i32* __restrict__ keys, i64 E, int shift, i64 nblocks, i64* __restrict__ table /* [256][nblocks] + 1 */) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const i64 base = (i64)blockIdx.x * RS_TILE;
    for (int q = 0; q < RS_ITEMS; ++q) {
        const i64 i = base + (i64)q * RS_THREADS + threadIdx.x;
        if (i < E) atomicAdd(&h[((u32)keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    table[1 + (i64)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // table[0] = 0: inclusive scan -> exclusive
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const i32* __restrict__ keys_in, const i64* __restrict__ vals_in, i32* __restrict__ keys_out,
             i64* __restrict__ vals_out, i64 E, int shift, i64 nblocks, const i64* __restrict__ table, int first_pass) {
    __shared__ u32 base_d[256];             // running count of each digit inside this tile
    __shared__ u32 wh[RS_THREADS / 32][256];  // per-warp digit counts of the current step
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u32 lt = (1u << lane) - 1u;
    base_d[tid] = 0;
    __syncthreads();
    const i64 tile0 = (i64)blockIdx.x * RS_TILE;
    for (int q = 0; q < RS_ITEMS; ++q) {  // step q: the 256 threads hold 256 CONSECUTIVE keys
        const i64 i = tile0 + (i64)q * RS_THREADS + tid;
        const bool valid = i < E;
        const i32 key = valid ? keys_in[i] : 0;
        const u32 d = valid ? (((u32)key >> shift) & 255u) : 0x100u + (u32)lane;
        for (int z = lane; z < 256; z += 32) wh[warp][z] = 0;
        __syncwarp();
        const u32 peers = __match_any_sync(FULL_MASK, d);
        const u32 rank_w = __popc(peers & lt);
        if (valid && rank_w == 0) wh[warp][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            u32 before = base_d[d];
            for (int w2 = 0; w2 < warp; ++w2) before += wh[w2][d];
            const i64 dst = table[(i64)d * nblocks + blockIdx.x] + before + rank_w;  // exclusive offsets after the scan
            keys_out[dst] = key;
            vals_out[dst] = first_pass ? i : vals_in[i];
        }
        __syncthreads();
        {   // advance the running counts by this step's totals
            u32 add = 0;
#pragma unroll
            for (int w2 = 0; w2 < RS_THREADS / 32; ++w2) add += wh[w2][tid];
            base_d[tid] += add;
        }
        __syncthreads();
    }
}

// Bytes of scratch needed to sort E columns.
extern "C" size_t ugs_sort_scratch_bytes(int64_t E) {
    const int64_t nblocks = (E + RS_TILE - 1) / RS_TILE;
    size_t b = 0;
    b += ((sizeof(i32) * (size_t)(E + 1) + 255) / 256) * 256 * 2;                  // key ping-pong
    b += ((sizeof(i64) * (size_t)(E + 1) + 255) / 256) * 256;                      // value pong
    b += ((sizeof(i64) * (size_t)(256 * nblocks + 2) + 255) / 256) * 256;          // digit table
    b += ((ugs_sample_scratch_bytes(256 * nblocks + 1, 1) + 255) / 256) * 256;     // scan status
    return b + 1024;
}

// perm_out[E] (int64) = stable ordering of the columns by workspace.colkey (written by ugs_segment_batch(fast=0)).
extern "C" int ugs_sort_columns(int64_t E, int64_t G, void* workspace, const ugs_layout* L, void* scratch,
                                size_t scratch_bytes, int64_t* perm_out, void* stream_) {
    if (!workspace || !L || (E > 0 && (!scratch || !perm_out))) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (scratch_bytes < ugs_sort_scratch_bytes(E)) return ugs_set_error(UGS_EINVAL, "sort scratch too small");
    if (E <= 0) return UGS_OK;
    cudaStream_t st = (cudaStream_t)stream_;
    WS w = ws_views(workspace, L);
    const i64 nblocks = (E + RS_TILE - 1) / RS_TILE;
    char* p = (char*)scratch;
    auto take = [&](size_t bytes) { char* r = p; p += ((bytes + 255) / 256) * 256; return r; };
    i32* kA = (i32*)take(sizeof(i32) * (size_t)(E + 1));
    i32* kB = (i32*)take(sizeof(i32) * (size_t)(E + 1));
    i64* vB = (i64*)take(sizeof(i64) * (size_t)(E + 1));
    i64* table = (i64*)take(sizeof(i64) * (size_t)(256 * nblocks + 2));
    const size_t scan_bytes = ugs_sample_scratch_bytes(256 * nblocks + 1, 1);
    void* scan_scratch = take(scan_bytes);
    int bits = 1;
    while ((1ll << bits) <= G) ++bits;  // keys are in [0, G]
    const int passes = (bits + 7) / 8;
    // ping-pong so that the LAST pass writes perm_out: values alternate vB <-> perm_out
    const i32* kin = w.colkey;
    const i64* vin = nullptr;
    for (int ps = 0; ps < passes; ++ps) {
        const bool last_to_perm = ((passes - 1 - ps) % 2) == 0;
        i32* kout = (ps % 2 == 0) ? kA : kB;
        i64* vout = last_to_perm ? perm_out : vB;
        UGS_CUDA_CHECK(cudaMemsetAsync(table, 0, sizeof(i64), st));
        k_rs_hist<<<(unsigned)nblocks, RS_THREADS, 0, st>>>(kin, E, 8 * ps, nblocks, table);
        UGS_LAUNCH_CHECK("k_rs_hist");
        // inclusive scan of table[1..] in place == exclusive offsets when read at index-1 -> table[d*nblocks+b]
        const int rc = ugs_run_scan(table, 256 * nblocks, scan_scratch, scan_bytes, w, st);
        if (rc != UGS_OK) return rc;
        k_rs_scatter<<<(unsigned)nblocks, RS_THREADS, 0, st>>>(kin, vin, kout, vout, E, 8 * ps, nblocks, table, ps == 0);
        UGS_LAUNCH_CHECK("k_rs_scatter");
        kin = kout;
        vin = vout;
    }
    return UGS_OK;
}
