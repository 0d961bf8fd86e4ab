// ugs_hostio.cu -- getting the sampler's output into the reference's CPU int64 ABI faster than PCIe can move int64.
//
// The reference returns five CPU pinned int64 tensors (sampler.cpp:190, ugs_sampler_batch_extension.cpp:95). For the
// ZINC-shaped batch that is 6.9 GB per call and a plain device->host copy of it runs at the PCIe rate (57 GB/s
// measured: 120 ms, sixty times the 2 ms the kernels take). Every value fits far fewer bits, so:
//   device   ugs_pack_outputs   nodes, edge_src -> 16-bit deltas against the minimum of their 4096-element block (int32
//                               when a block spreads too far), edge_index -> one byte (j | l << 4) per entry in
//                               mode "sample" with k <= 16, else two int32 per entry            (HBM-bound, ~1 ms)
//   PCIe     the packed "wire" buffer (1.0 GB instead of 6.8 GB) in chunks, one event per chunk
//   host     ugs_fetch_outputs  worker threads widen every chunk into the final pinned int64 tensors as soon as its
//                               event has fired, with non-temporal stores (no read-for-ownership of the destination;
//                               measured on the pool's 16-vCPU hosts: 145 GB/s of int64 written from bytes, 90 GB/s
//                               from int32 -- profiles/microbench/host_widen.cpp)
// The widening is pure format conversion of finished results (sign extension / nibble split); nothing is computed on
// the host. edge_ptr and sample_ptr are small and copied as they are.
#include <emmintrin.h>

#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#include "ugs_common.cuh"

// Chunk of the wire buffer per event (UGS_B200_WIRE_CHUNK_MB overrides). Measured on B200 hosts: 2 ... 64 MB chunks all
// give 1.23-1.28e8 graphlets/s end to end -- the path is bound by host memory bandwidth, not by the chunking.
static long long wire_chunk_bytes() {
    static long long v = 0;
    if (!v) {
        const char* e = getenv("UGS_B200_WIRE_CHUNK_MB");
        long long mb = e ? atoll(e) : 0;
        if (mb < 1 || mb > 1024) mb = 32;
        v = mb << 20;
    }
    return v;
}

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
#define DELTA_BLOCK 4096
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

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct WireLayout {
    // byte offsets; `bad` = one device counter of values that did not fit; *_base = int64 block minima (delta16 only)
    size_t nodes, src, ei, nodes_base, src_base, bad, total;
};
static inline size_t delta_blocks(size_t n) { return (n + DELTA_BLOCK - 1) / DELTA_BLOCK; }
static WireLayout wire_layout(int64_t B, int k, int64_t total, int pairs_u8, int delta16) {
    WireLayout w;
    const size_t nn = (size_t)B * k, ne = (size_t)total, eb = delta16 ? 2 : 4;
    w.nodes = 0;
    w.src = align256(nn * eb);
    w.ei = w.src + align256(ne * eb);
    w.nodes_base = w.ei + align256(pairs_u8 ? ne : ne * 8);
    w.src_base = w.nodes_base + align256(delta16 ? delta_blocks(nn) * 8 : 0);
    w.bad = w.src_base + align256(delta16 ? delta_blocks(ne) * 8 : 0);
    w.total = w.bad + 256;
    return w;
}

extern "C" size_t ugs_wire_bytes(int64_t B, int k, int64_t total, int pairs_u8, int delta16) {
    if (B < 0 || total < 0 || k < 1) return 0;
    return wire_layout(B, k, total, pairs_u8, delta16).total;
}

// Number of values the last ugs_pack_outputs could not represent (synchronises the stream; 8 bytes device -> host).
extern "C" int ugs_pack_status(const void* d_wire, int64_t B, int k, int64_t total, int pairs_u8, int delta16,
                               int64_t* h_bad, void* stream_) {
    if (!d_wire || !h_bad) return ugs_set_error(UGS_EINVAL, "NULL argument");
    const WireLayout w = wire_layout(B, k, total, pairs_u8, delta16);
    UGS_CUDA_CHECK(cudaMemcpyAsync(h_bad, (const char*)d_wire + w.bad, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
    UGS_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream_));
    return UGS_OK;
}

extern "C" int ugs_pack_outputs(const int64_t* nodes, int64_t B, int k, const int64_t* edge_index,
                                const int64_t* edge_src, int64_t total, int pairs_u8, int delta16, void* d_wire,
                                void* stream_) {
    if (B < 0 || total < 0 || k < 1) return ugs_set_error(UGS_EINVAL, "bad size");
    if (!d_wire) return ugs_set_error(UGS_EINVAL, "d_wire is NULL");
    cudaStream_t st = (cudaStream_t)stream_;
    const WireLayout w = wire_layout(B, k, total, pairs_u8, delta16);
    char* base = (char*)d_wire;
    unsigned long long* bad = (unsigned long long*)(base + w.bad);
    UGS_CUDA_CHECK(cudaMemsetAsync(bad, 0, 8, st));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 16;
    if (B * k > 0) {
        if (!nodes) return ugs_set_error(UGS_EINVAL, "nodes is NULL");
        if (delta16)
            k_pack_delta16<<<(unsigned)delta_blocks((size_t)B * k), 256, 0, st>>>(nodes, B * k, (i64*)(base + w.nodes_base),
                                                                                 (unsigned short*)(base + w.nodes), bad);
        else k_pack_i32<<<grid, 256, 0, st>>>(nodes, (i32*)(base + w.nodes), B * k, bad);
        UGS_LAUNCH_CHECK("k_pack_nodes");
    }
    if (total > 0) {
        if (!edge_index || !edge_src) return ugs_set_error(UGS_EINVAL, "NULL edge array");
        if (delta16)
            k_pack_delta16<<<(unsigned)delta_blocks((size_t)total), 256, 0, st>>>(edge_src, total, (i64*)(base + w.src_base),
                                                                              (unsigned short*)(base + w.src), bad);
        else k_pack_i32<<<grid, 256, 0, st>>>(edge_src, (i32*)(base + w.src), total, bad);
        UGS_LAUNCH_CHECK("k_pack_edge_src");
        if (pairs_u8) k_pack_pairs_u8<<<grid, 256, 0, st>>>(edge_index, edge_index + total, (unsigned char*)(base + w.ei), total, bad);
        else k_pack_i32<<<grid, 256, 0, st>>>(edge_index, (i32*)(base + w.ei), 2 * total, bad);
        UGS_LAUNCH_CHECK("k_pack_edge_index");
    }
    return UGS_OK;
}

// ---- host side -------------------------------------------------------------------------------------------------
static void widen_i32(const i32* src, i64* dst, size_t n) {
    for (size_t i = 0; i < n; ++i) _mm_stream_si64((long long*)(dst + i), (long long)src[i]);
}
static void widen_pairs(const unsigned char* src, i64* du, i64* dv, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const unsigned b = src[i];
        _mm_stream_si64((long long*)(du + i), (long long)(b & 15u));
        _mm_stream_si64((long long*)(dv + i), (long long)(b >> 4));
    }
}

// elements [first, first + n) of an array packed as 16-bit deltas; bases[] holds one int64 per DELTA_BLOCK elements
static void widen_delta16(const unsigned short* src, const i64* bases, size_t first, i64* dst, size_t n) {
    size_t i = 0;
    while (i < n) {
        const size_t g = first + i, blk = g / DELTA_BLOCK;
        size_t end = (blk + 1) * DELTA_BLOCK - first;
        if (end > n) end = n;
        const i64 b = bases[blk];
        for (; i < end; ++i) {
            const unsigned d = src[i];
            _mm_stream_si64((long long*)(dst + i), d == 0xFFFFu ? -1ll : (long long)(b + (i64)d));
        }
    }
}

struct WireItem {
    int kind;        // 0: int32 -> int64, 1: byte -> (u, v), 2: 16-bit deltas -> int64
    const i64* bases;  // kind 2
    size_t first;      // kind 2: index of the chunk's element 0 inside its array
    size_t off;      // byte offset of the chunk inside the wire buffer
    size_t n;        // elements in the chunk
    i64* dst;        // destination of element 0 of the chunk
    i64* dst2;       // kind 1: second row
    cudaEvent_t ev;
};

static std::vector<cudaEvent_t> g_event_pool;
static std::mutex g_fetch_mutex;  // one delivery at a time per process: the event pool is shared and the host threads
                                  // of two concurrent deliveries would only fight for the same memory bandwidth

extern "C" int ugs_fetch_outputs(const void* d_wire, void* h_wire, int64_t B, int k, int64_t total, int pairs_u8,
                                 int delta16, int64_t* h_nodes, int64_t* h_edge_index, int64_t* h_edge_src,
                                 int n_threads, void* stream_) {
    if (B < 0 || total < 0 || k < 1) return ugs_set_error(UGS_EINVAL, "bad size");
    if (!d_wire || !h_wire) return ugs_set_error(UGS_EINVAL, "NULL wire buffer");
    if ((B * k > 0 && !h_nodes) || (total > 0 && (!h_edge_index || !h_edge_src))) return ugs_set_error(UGS_EINVAL, "NULL destination");
    cudaStream_t st = (cudaStream_t)stream_;
    std::lock_guard<std::mutex> lock(g_fetch_mutex);
    const WireLayout w = wire_layout(B, k, total, pairs_u8, delta16);
    const char* dsrc = (const char*)d_wire;
    char* hsrc = (char*)h_wire;
    if (delta16) {  // the block minima first: every later chunk event implies they have arrived
        const size_t bytes = w.bad - w.nodes_base;
        if (bytes) UGS_CUDA_CHECK(cudaMemcpyAsync(hsrc + w.nodes_base, dsrc + w.nodes_base, bytes, cudaMemcpyDeviceToHost, st));
    }
    std::vector<WireItem> items;
    auto add_region = [&](int kind, size_t off, size_t n, size_t elem, i64* dst, i64* dst2, size_t bases_off) {
        const size_t per = (size_t)wire_chunk_bytes() / elem;
        for (size_t a = 0; a < n; a += per) {
            WireItem it;
            it.kind = kind; it.off = off + a * elem; it.n = (n - a < per) ? n - a : per;
            it.dst = dst + a; it.dst2 = dst2 ? dst2 + a : nullptr; it.ev = nullptr;
            it.bases = reinterpret_cast<const i64*>(hsrc + bases_off); it.first = a;
            items.push_back(it);
        }
    };
    // biggest consumers of host write bandwidth first, so that the tail of the transfer overlaps the least work
    if (pairs_u8) add_region(1, w.ei, (size_t)total, 1, h_edge_index, h_edge_index + total, 0);
    else add_region(0, w.ei, (size_t)total * 2, 4, h_edge_index, nullptr, 0);
    if (delta16) {
        add_region(2, w.src, (size_t)total, 2, h_edge_src, nullptr, w.src_base);
        add_region(2, w.nodes, (size_t)B * k, 2, h_nodes, nullptr, w.nodes_base);
    } else {
        add_region(0, w.src, (size_t)total, 4, h_edge_src, nullptr, 0);
        add_region(0, w.nodes, (size_t)B * k, 4, h_nodes, nullptr, 0);
    }
    while (g_event_pool.size() < items.size() + 1) {
        cudaEvent_t e;
        UGS_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        g_event_pool.push_back(e);
    }
    for (size_t i = 0; i < items.size(); ++i) {
        const size_t bytes = items[i].n * (items[i].kind == 1 ? 1 : items[i].kind == 2 ? 2 : 4);
        UGS_CUDA_CHECK(cudaMemcpyAsync(hsrc + items[i].off, dsrc + items[i].off, bytes, cudaMemcpyDeviceToHost, st));
        items[i].ev = g_event_pool[i];
        UGS_CUDA_CHECK(cudaEventRecord(items[i].ev, st));
    }
    UGS_CUDA_CHECK(cudaMemcpyAsync(hsrc + w.bad, dsrc + w.bad, 8, cudaMemcpyDeviceToHost, st));
    cudaEvent_t ev_last = g_event_pool[items.size()];
    UGS_CUDA_CHECK(cudaEventRecord(ev_last, st));

    int T = n_threads;
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    int dev = 0;
    cudaGetDevice(&dev);
    auto worker = [&](int t) {
        cudaSetDevice(dev);
        for (size_t i = 0; i < items.size(); ++i) {
            const WireItem& it = items[i];
            cudaEventSynchronize(it.ev);
            // 8-element granularity keeps every thread's slice on whole 64-byte destination lines
            const size_t per = ((it.n + T - 1) / T + 7) & ~(size_t)7;
            const size_t a = (size_t)t * per, b = (a + per < it.n) ? a + per : it.n;
            if (a >= b) continue;
            if (it.kind == 1) widen_pairs((const unsigned char*)(hsrc + it.off) + a, it.dst + a, it.dst2 + a, b - a);
            else if (it.kind == 2) widen_delta16((const unsigned short*)(hsrc + it.off) + a, it.bases, it.first + a, it.dst + a, b - a);
            else widen_i32((const i32*)(hsrc + it.off) + a, it.dst + a, b - a);
        }
        _mm_sfence();
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto& x : th) x.join();
    UGS_CUDA_CHECK(cudaEventSynchronize(ev_last));
    const unsigned long long bad = *(const unsigned long long*)(hsrc + w.bad);
    if (bad) return ugs_set_error(UGS_ELIMIT, "%llu output values do not fit the packed wire format", bad);
    return UGS_OK;
}
