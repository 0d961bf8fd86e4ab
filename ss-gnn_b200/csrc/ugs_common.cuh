// ugs_common.cuh -- shared device helpers for libugs_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ugs_b200.h"

typedef int64_t i64;
typedef int32_t i32;
typedef uint32_t u32;
typedef uint64_t u64;

#define FULL_MASK 0xffffffffu

// Host-side error plumbing (defined in ugs_capi.cu)
int ugs_set_error(int code, const char* fmt, ...);
#define UGS_CUDA_CHECK(expr)                                                             \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return ugs_set_error(UGS_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)
extern unsigned long long g_ugs_launches;  // kernels launched by this library (bench.py reports it)
#define UGS_LAUNCH_CHECK(name)                                                           \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess)                                                           \
            return ugs_set_error(UGS_ECUDA, "launch %s failed: %s", name, cudaGetErrorString(_e)); \
        ++g_ugs_launches;                                                                \
    } while (0)

// Typed views into the workspace.
struct WS {
    i64* cstart;
    ugs_ginfo* ginfo;
    i32* glist;
    unsigned long long* counters;
    i32* colkey;
    u32* rowptr;
    i32* order;
    i32* index_of;
    i32* suffix_deg;
    double* bucket_b;
    double* cum;
    u64* adj;
    u64* suf;
    i32* csr_idx;
    i32* csr_col;
    double* gaux;
    i32* llist;        // ids of the LARGE general-path graphs (preprocessed by the cooperative kernel)
    double* lscratch;  // [4096] scratch of the cooperative kernel
};

static inline WS ws_views(void* base, const ugs_layout* L) {
    char* b = (char*)base;
    WS w;
    w.cstart = (i64*)(b + L->cstart);
    w.ginfo = (ugs_ginfo*)(b + L->ginfo);
    w.glist = (i32*)(b + L->glist);
    w.counters = (unsigned long long*)(b + L->counters);
    w.colkey = (i32*)(b + L->colkey);
    w.rowptr = (u32*)(b + L->rowptr);
    w.order = (i32*)(b + L->order);
    w.index_of = (i32*)(b + L->index_of);
    w.suffix_deg = (i32*)(b + L->suffix_deg);
    w.bucket_b = (double*)(b + L->bucket_b);
    w.cum = (double*)(b + L->cum);
    w.adj = (u64*)(b + L->adj);
    w.suf = (u64*)(b + L->suf);
    w.csr_idx = (i32*)(b + L->csr_idx);
    w.csr_col = (i32*)(b + L->csr_col);
    w.gaux = (double*)(b + L->gaux);
    w.llist = (i32*)(b + L->llist);
    w.lscratch = (double*)(b + L->lscratch);
    return w;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, counter = (sample i, draw block, graph lo, graph hi), key = (seed lo, seed hi).
// Replaces the reference's xorshift ThreadRNG (sampler.hpp:26-36) and its graph-independent
// seeding (sampler.cpp:160-161). Bit-identical to oracle/ugs_oracle.c::philox4x32_10.
// ---------------------------------------------------------------------------------------------
struct Philox4 {
    u32 x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        u32 hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        u32 hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        u32 n0 = hi1 ^ c1 ^ k0;
        u32 n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }

// position (0-based) of the r-th (0-based) set bit of a 32-bit word; requires r < popc(m)
__device__ __forceinline__ int select_bit32(u32 m, int r) {
    int pos = 0;
    int c = __popc(m & 0xffffu);
    if (r >= c) { r -= c; pos += 16; m >>= 16; }
    c = __popc(m & 0xffu);
    if (r >= c) { r -= c; pos += 8; m >>= 8; }
    c = __popc(m & 0xfu);
    if (r >= c) { r -= c; pos += 4; m >>= 4; }
    c = __popc(m & 0x3u);
    if (r >= c) { r -= c; pos += 2; m >>= 2; }
    c = m & 1u;
    if (r >= c) { pos += 1; }
    return pos;
}

__device__ __forceinline__ int select_bit64(u64 m, int r) {
    u32 lo = (u32)m;
    int c = __popc(lo);
    if (r < c) return select_bit32(lo, r);
    return 32 + select_bit32((u32)(m >> 32), r - c);
}

// root-draw uniform in [0,1) * Z from two 32-bit words (53-bit mantissa), same arithmetic as the oracle
__device__ __forceinline__ double root_x(u32 w0, u32 w1, double Z) {
    u64 u53 = ((((u64)w0) << 32) | (u64)w1) >> 11;
    return ((double)u53 * (1.0 / 9007199254740992.0)) * Z;
}

// inclusive warp prefix sum of a small non-negative int
__device__ __forceinline__ int warp_incl_scan(int c, int lane) {
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(FULL_MASK, incl, off);
        if (lane >= off) incl += y;
    }
    return incl;
}

// First index a in [0, n) with cm[a] > x, or n - 1 when there is none: what the scalar binary search of the bitmask
// path returns, found with 32 probes per round (a 10^5-entry CDF: 4 dependent rounds of L2 latency instead of 17).
__device__ __forceinline__ int warp_cdf_search(const double* __restrict__ cm, int n, double x, int lane) {
    int lo = 0, len = n;
    while (len > 32) {
        const int stride = (len + 31) >> 5;
        const int idx = min(lo + (lane + 1) * stride - 1, lo + len - 1);
        const u32 ball = __ballot_sync(FULL_MASK, cm[idx] > x);
        if (ball == 0u) return n - 1;  // the last probe is the last element of the range
        const int first = __ffs(ball) - 1;
        const int hi = min(lo + (first + 1) * stride - 1, lo + len - 1);  // probe `first`: true
        lo += first * stride;                                            // probe `first - 1` (at lo - 1): false
        len = hi - lo + 1;
    }
    const u32 ball = __ballot_sync(FULL_MASK, lane < len && cm[lo + lane] > x);
    return ball ? lo + __ffs(ball) - 1 : n - 1;
}

// Asynchronous global -> shared copies (LDGSTS): a thread issues any number of them without waiting, so a CTA staging
// several tables has all their loads in flight at once; cp_async_wait_all + __syncthreads() makes them visible.
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// PROTEINS-sized general-path graphs: their tables fit a CTA's shared memory (k_preproc_medium, k_sample_medium)
__host__ __device__ __forceinline__ bool ugs_is_medium(long long n, long long nc) { return n <= 1024 && nc <= 3072; }

// internal entry points shared between translation units
#define UGS_EMIT_GUARD_GENERAL 1  // write nothing when counters[UGS_CTR_N_GENERAL] != 0 (fused path queued on a guess)
#define UGS_EMIT_COMPACT 2        // out_v / out_src follow out_u at counters[UGS_CTR_TOTAL_EDGES] (read on the device)
int ugs_emit_edges_impl(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int mode, int use_perm,
                        int64_t n_general, const void* workspace, const ugs_layout* L, const int64_t* nodes,
                        const int64_t* edge_ptr, int64_t total, int64_t* edge_index_out, int64_t* edge_src_out,
                        const void* sample_scratch, size_t sample_scratch_bytes, int flags, void* stream_);
int ugs_sample_batch_impl(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, uint64_t seed,
                          uint64_t graph_index_base, uint64_t sample_index_base, int seed_mode, int variant,
                          int64_t n_general, int64_t max_n_general, void* workspace, const ugs_layout* L,
                          void* scratch, size_t scratch_bytes, int64_t* nodes, int64_t* edge_ptr,
                          int64_t* sample_ptr, int64_t* h_info, void* stream_, bool counters_fresh,
                          int64_t* snap = nullptr, int64_t snap_seq = 0);  // snap: see k_scan_counts
