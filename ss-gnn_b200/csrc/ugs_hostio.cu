// ugs_hostio.cu -- getting the sampler's output into the reference's CPU int64 ABI faster than PCIe can move int64.
//
// The reference returns five CPU pinned int64 tensors (sampler.cpp:190, ugs_sampler_batch_extension.cpp:95). For the
// ZINC-shaped batch that is 6.9 GB per call and a plain device->host copy of it runs at the PCIe rate (57 GB/s
// measured: 120 ms, sixty times the 2 ms the kernels take). Every value fits far fewer bits, so:
//   device   ugs_pack_outputs   nodes -> int32, edge_src -> int32, edge_index -> one byte (j | l << 4) per entry in
//                               mode "sample" with k <= 16, else two int32 per entry            (HBM-bound, ~1 ms)
//   PCIe     the packed "wire" buffer (1.6 GB instead of 6.8 GB) in chunks, one event per chunk
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

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct WireLayout {
    size_t nodes, src, ei, bad, total;  // byte offsets; `bad` = one device counter of values that did not fit
};
static WireLayout wire_layout(int64_t B, int k, int64_t total, int pairs_u8) {
    WireLayout w;
    w.nodes = 0;
    w.src = align256((size_t)B * k * 4);
    w.ei = w.src + align256((size_t)total * 4);
    w.bad = w.ei + align256(pairs_u8 ? (size_t)total : (size_t)total * 8);
    w.total = w.bad + 256;
    return w;
}

extern "C" size_t ugs_wire_bytes(int64_t B, int k, int64_t total, int pairs_u8) {
    if (B < 0 || total < 0 || k < 1) return 0;
    return wire_layout(B, k, total, pairs_u8).total;
}

extern "C" int ugs_pack_outputs(const int64_t* nodes, int64_t B, int k, const int64_t* edge_index,
                                const int64_t* edge_src, int64_t total, int pairs_u8, void* d_wire, void* stream_) {
    if (B < 0 || total < 0 || k < 1) return ugs_set_error(UGS_EINVAL, "bad size");
    if (!d_wire) return ugs_set_error(UGS_EINVAL, "d_wire is NULL");
    cudaStream_t st = (cudaStream_t)stream_;
    const WireLayout w = wire_layout(B, k, total, pairs_u8);
    char* base = (char*)d_wire;
    unsigned long long* bad = (unsigned long long*)(base + w.bad);
    UGS_CUDA_CHECK(cudaMemsetAsync(bad, 0, 8, st));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 16;
    if (B * k > 0) {
        if (!nodes) return ugs_set_error(UGS_EINVAL, "nodes is NULL");
        k_pack_i32<<<grid, 256, 0, st>>>(nodes, (i32*)(base + w.nodes), B * k, bad);
        UGS_LAUNCH_CHECK("k_pack_i32");
    }
    if (total > 0) {
        if (!edge_index || !edge_src) return ugs_set_error(UGS_EINVAL, "NULL edge array");
        k_pack_i32<<<grid, 256, 0, st>>>(edge_src, (i32*)(base + w.src), total, bad);
        UGS_LAUNCH_CHECK("k_pack_i32");
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

struct WireItem {
    int kind;        // 0: int32 -> int64, 1: byte -> (u, v)
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
                                 int64_t* h_nodes, int64_t* h_edge_index, int64_t* h_edge_src, int n_threads,
                                 void* stream_) {
    if (B < 0 || total < 0 || k < 1) return ugs_set_error(UGS_EINVAL, "bad size");
    if (!d_wire || !h_wire) return ugs_set_error(UGS_EINVAL, "NULL wire buffer");
    if ((B * k > 0 && !h_nodes) || (total > 0 && (!h_edge_index || !h_edge_src))) return ugs_set_error(UGS_EINVAL, "NULL destination");
    cudaStream_t st = (cudaStream_t)stream_;
    std::lock_guard<std::mutex> lock(g_fetch_mutex);
    const WireLayout w = wire_layout(B, k, total, pairs_u8);
    std::vector<WireItem> items;
    auto add_region = [&](int kind, size_t off, size_t n, size_t elem, i64* dst, i64* dst2) {
        const size_t per = (size_t)wire_chunk_bytes() / elem;
        for (size_t a = 0; a < n; a += per) {
            WireItem it;
            it.kind = kind; it.off = off + a * elem; it.n = (n - a < per) ? n - a : per;
            it.dst = dst + a; it.dst2 = dst2 ? dst2 + a : nullptr; it.ev = nullptr;
            items.push_back(it);
        }
    };
    // biggest consumers of host write bandwidth first, so that the tail of the transfer overlaps the least work
    if (pairs_u8) add_region(1, w.ei, (size_t)total, 1, h_edge_index, h_edge_index + total);
    else add_region(0, w.ei, (size_t)total * 2, 4, h_edge_index, nullptr);
    add_region(0, w.src, (size_t)total, 4, h_edge_src, nullptr);
    add_region(0, w.nodes, (size_t)B * k, 4, h_nodes, nullptr);
    while (g_event_pool.size() < items.size() + 1) {
        cudaEvent_t e;
        UGS_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        g_event_pool.push_back(e);
    }
    const char* dsrc = (const char*)d_wire;
    char* hsrc = (char*)h_wire;
    for (size_t i = 0; i < items.size(); ++i) {
        const size_t bytes = items[i].n * (items[i].kind == 1 ? 1 : 4);
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
