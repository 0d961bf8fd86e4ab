// ugs_hostio.cu -- device half of the host delivery (the host half is ugs_hostdeliver.cpp): getting the sampler's output
// into the reference's CPU int64 ABI faster than PCIe can move int64.
//
// The reference returns five CPU pinned int64 tensors (sampler.cpp:190, ugs_sampler_batch_extension.cpp:95). For the
// ZINC-shaped batch that is 6.9 GB per call and a plain device->host copy of it runs at the PCIe rate (57 GB/s
// measured: 120 ms, sixty times the 2 ms the kernels take). Every value fits far fewer bits, so the part of every array
// that the host cores can widen in time (ugs_deliver_outputs decides the split) is packed here first:
//   nodes, edge_src -> 16-bit deltas against the minimum of their 4096-element block (int32 when a block spreads too
//   far), edge_index -> one byte (j | l << 4) per entry in mode "sample" with k <= 16, else two int32 per entry
//   (HBM-bound, ~1 ms for the whole ZINC-shaped result); the packed "wire" buffer is 1/7 of the int64 size.
// The widening on the host is pure format conversion of finished results (sign extension / nibble split); nothing is
// computed there. edge_ptr and sample_ptr are small and copied as they are.
#include "ugs_common.cuh"
#include "ugs_hostio.h"

__global__ void __launch_bounds__(256)
k_pack_i32(const i64* __restrict__ src, i32* __restrict__ dst, i64 n, unsigned long long* __restrict__ bad) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const i64 v = src[i];
        if (v < -2147483648ll || v > 2147483647ll) atomicAdd(bad, 1ull);
        dst[i] = (i32)v;
    }
}

__global__ void __launch_bounds__(256)
k_pack_pairs_u8(const i64* __restrict__ u, const i64* __restrict__ v, unsigned char* __restrict__ dst, i64 n,
                unsigned long long* __restrict__ bad) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const i64 a = u[i], b = v[i];
        if ((a | b) & ~15ll) atomicAdd(bad, 1ull);
        dst[i] = (unsigned char)((a & 15) | ((b & 15) << 4));
    }
}

// Frame-of-reference packing: 4096 consecutive values share one int64 base (their minimum) and travel as 16-bit deltas;
// 0xFFFF = the value -1 (padding of incomplete rows / "no source column"). Consecutive entries of the output belong to
// consecutive samples and graphs, so their node ids / source columns are close together; a block whose values spread
// over more than 65534 (or hold another negative value) is counted in *bad and the caller repacks as int32.
#define DELTA_BLOCK UGS_DELTA_BLOCK
__global__ void __launch_bounds__(256)
k_pack_delta16(const i64* __restrict__ src, i64 n, i64* __restrict__ bases, unsigned short* __restrict__ dst,
               unsigned long long* __restrict__ bad) {
    __shared__ i64 s_min[8];
    const i64 b0 = (i64)blockIdx.x * DELTA_BLOCK;
    i64 v[16];
    i64 mn = 0x7fffffffffffffffll;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const i64 i = b0 + q * 256 + threadIdx.x;
        v[q] = (i < n) ? src[i] : -1;
        if (v[q] >= 0 && v[q] < mn) mn = v[q];
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { const i64 y = __shfl_xor_sync(FULL_MASK, mn, o); mn = y < mn ? y : mn; }
    if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = mn;
    __syncthreads();
    mn = s_min[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) mn = s_min[q] < mn ? s_min[q] : mn;
    if (mn == 0x7fffffffffffffffll) mn = 0;  // no non-negative value in the block
    if (threadIdx.x == 0) bases[blockIdx.x] = mn;
    unsigned nbad = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const i64 i = b0 + q * 256 + threadIdx.x;
        if (i >= n) continue;
        const i64 d = v[q] - mn;
        unsigned short o = 0xFFFFu;
        if (v[q] >= 0) { if (d >= 0xFFFF) nbad++; else o = (unsigned short)d; }
        else if (v[q] != -1) nbad++;
        dst[i] = o;
    }
    if (nbad) atomicAdd(bad, (unsigned long long)nbad);
}

// bad[0]: values that do not fit a 16-bit delta (the caller repacks as int32); bad[1]: values that fit neither one nibble
// per endpoint (pairs_u8) nor int32 -- the packed formats cannot carry them at all.
int ugs_pack_parts(const int64_t* nodes, size_t pn, const int64_t* ei_row0, const int64_t* ei_row1, size_t pi,
                   const int64_t* edge_src, size_t pe, int pairs_u8, int delta16, void* d_wire, const UgsWire* w,
                   void* stream_) {
    if (!d_wire || !w) return ugs_set_error(UGS_EINVAL, "d_wire is NULL");
    cudaStream_t st = (cudaStream_t)stream_;
    char* base = (char*)d_wire;
    unsigned long long* bad = (unsigned long long*)(base + w->bad);
    UGS_CUDA_CHECK(cudaMemsetAsync(bad, 0, 16, st));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 16;
    if (pn > 0) {
        if (delta16)
            k_pack_delta16<<<(unsigned)ugs_delta_blocks(pn), 256, 0, st>>>(nodes, (i64)pn, (i64*)(base + w->nodes_base),
                                                                          (unsigned short*)(base + w->nodes), bad);
        else k_pack_i32<<<grid, 256, 0, st>>>(nodes, (i32*)(base + w->nodes), (i64)pn, bad + 1);
        UGS_LAUNCH_CHECK("k_pack_nodes");
    }
    if (pe > 0) {
        if (delta16)
            k_pack_delta16<<<(unsigned)ugs_delta_blocks(pe), 256, 0, st>>>(edge_src, (i64)pe, (i64*)(base + w->src_base),
                                                                          (unsigned short*)(base + w->src), bad);
        else k_pack_i32<<<grid, 256, 0, st>>>(edge_src, (i32*)(base + w->src), (i64)pe, bad + 1);
        UGS_LAUNCH_CHECK("k_pack_edge_src");
    }
    if (pi > 0) {
        if (pairs_u8) k_pack_pairs_u8<<<grid, 256, 0, st>>>(ei_row0, ei_row1, (unsigned char*)(base + w->ei), (i64)pi, bad + 1);
        else {
            k_pack_i32<<<grid, 256, 0, st>>>(ei_row0, (i32*)(base + w->ei), (i64)pi, bad + 1);
            k_pack_i32<<<grid, 256, 0, st>>>(ei_row1, (i32*)(base + w->ei) + pi, (i64)pi, bad + 1);
        }
        UGS_LAUNCH_CHECK("k_pack_edge_index");
    }
    return UGS_OK;
}
