// ugs_hostdeliver.cpp -- host half of the delivery of the sampler's output as the reference's CPU int64 tensors
// (reference sampler.cpp:190, ugs_sampler_batch_extension.cpp:95 return pinned CPU int64 tensors). Compiled by g++.
//
// The ZINC-shaped result is 6.9 GB of int64 per call. Two pipelines run at the same time and each array is split between
// them (ugs_deliver_outputs, `dma_fraction`):
//   * DMA     a prefix of every array is copied as it is, device int64 -> final pinned host tensor, in slices issued
//             between the wire chunks: 8 bytes over PCIe and 8 bytes of host DRAM traffic per value, no CPU work;
//   * widen   the rest is packed on the device (csrc/ugs_hostio.cu: 16-bit deltas per 4096-value block for node ids /
//             source columns, one byte per edge entry), crosses PCIe as ~1/7 of its final size in 32 MB chunks, and
//             worker threads widen every chunk into the final tensors with full-line non-temporal vector stores
//             (AVX-512 / AVX2, chosen at run time) as soon as its event has fired.
// PCIe bounds the first, host cores / memory bandwidth the second; the caller picks the split from the cores it owns
// (one rank with 16 cores: mostly widen; eight ranks sharing 32 cores: mostly DMA). Pure format conversion of finished
// results -- nothing is computed on the host.
// The worker threads live in a process-wide pool (no thread creation per call); concurrent calls share it and the calling
// thread always takes part, so a call never waits for a busy pool.
#include <cuda_runtime_api.h>
#include <immintrin.h>
#include <stdint.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "ugs_b200.h"
#include "ugs_hostio.h"

typedef int64_t i64;
typedef int32_t i32;

// ---- widening kernels ---------------------------------------------------------------------------------------------
// kind 0: int32 -> int64; kind 1: low nibble of a byte -> int64; kind 2: high nibble -> int64;
// kind 3: 16-bit delta + base -> int64 (0xFFFF -> -1), one base per call (the caller splits at block boundaries);
// kind 4: 8-bit delta + base -> int64 (0xFF -> -1), likewise
static inline void st64(i64* p, i64 v) { _mm_stream_si64((long long*)p, (long long)v); }

static void widen_scalar(int kind, const void* src, i64* dst, size_t n, i64 base) {
    if (kind == 0) { const i32* s = (const i32*)src; for (size_t i = 0; i < n; ++i) st64(dst + i, (i64)s[i]); }
    else if (kind == 1) { const uint8_t* s = (const uint8_t*)src; for (size_t i = 0; i < n; ++i) st64(dst + i, (i64)(s[i] & 15u)); }
    else if (kind == 2) { const uint8_t* s = (const uint8_t*)src; for (size_t i = 0; i < n; ++i) st64(dst + i, (i64)(s[i] >> 4)); }
    else if (kind == 3) { const uint16_t* s = (const uint16_t*)src; for (size_t i = 0; i < n; ++i) st64(dst + i, s[i] == 0xFFFFu ? (i64)-1 : base + (i64)s[i]); }
    else { const uint8_t* s = (const uint8_t*)src; for (size_t i = 0; i < n; ++i) st64(dst + i, s[i] == 0xFFu ? (i64)-1 : base + (i64)s[i]); }
}

__attribute__((target("avx2"))) static void widen_avx2(int kind, const void* src, i64* dst, size_t n, i64 base) {
    size_t i = 0;
    const size_t head = ((32 - ((uintptr_t)dst & 31)) & 31) / 8;  // scalar until the destination is 32-byte aligned
    const size_t h = head < n ? head : n;
    widen_scalar(kind, src, dst, h, base);
    i = h;
    const __m256i vbase = _mm256_set1_epi64x(base), vm1 = _mm256_set1_epi64x(-1), vff = _mm256_set1_epi64x(0xFFFF);
    const __m256i vf8 = _mm256_set1_epi64x(0xFF);
    const __m128i nib = _mm_set1_epi8(15);
    for (; i + 4 <= n; i += 4) {
        __m256i v;
        if (kind == 0) v = _mm256_cvtepi32_epi64(_mm_loadu_si128((const __m128i*)((const i32*)src + i)));
        else if (kind == 4) {
            uint32_t w;
            memcpy(&w, (const uint8_t*)src + i, 4);
            const __m256i d = _mm256_cvtepu8_epi64(_mm_cvtsi32_si128((int)w));
            v = _mm256_blendv_epi8(_mm256_add_epi64(d, vbase), vm1, _mm256_cmpeq_epi64(d, vf8));
        } else if (kind == 1 || kind == 2) {
            uint32_t w;
            memcpy(&w, (const uint8_t*)src + i, 4);
            __m128i b = _mm_cvtsi32_si128((int)w);
            if (kind == 2) b = _mm_srli_epi16(b, 4);
            v = _mm256_cvtepu8_epi64(_mm_and_si128(b, nib));
        } else {
            uint64_t w;
            memcpy(&w, (const uint16_t*)src + i, 8);
            const __m256i d = _mm256_cvtepu16_epi64(_mm_cvtsi64_si128((long long)w));
            v = _mm256_blendv_epi8(_mm256_add_epi64(d, vbase), vm1, _mm256_cmpeq_epi64(d, vff));
        }
        _mm256_stream_si256((__m256i*)(dst + i), v);
    }
    const size_t esz = kind == 0 ? 4 : kind == 3 ? 2 : 1;
    widen_scalar(kind, (const char*)src + i * esz, dst + i, n - i, base);
}

__attribute__((target("avx512f"))) static void widen_avx512(int kind, const void* src, i64* dst, size_t n, i64 base) {
    size_t i = 0;
    const size_t head = ((64 - ((uintptr_t)dst & 63)) & 63) / 8;  // scalar until the destination starts a 64-byte line
    const size_t h = head < n ? head : n;
    widen_scalar(kind, src, dst, h, base);
    i = h;
    const __m512i vbase = _mm512_set1_epi64(base), vm1 = _mm512_set1_epi64(-1), vff = _mm512_set1_epi64(0xFFFF);
    const __m512i vf8 = _mm512_set1_epi64(0xFF);
    const __m128i nib = _mm_set1_epi8(15);
    for (; i + 8 <= n; i += 8) {  // one full cache line per store: no read-for-ownership, no partial-line merging
        __m512i v;
        if (kind == 0) v = _mm512_cvtepi32_epi64(_mm256_loadu_si256((const __m256i*)((const i32*)src + i)));
        else if (kind == 4) {
            uint64_t w;
            memcpy(&w, (const uint8_t*)src + i, 8);
            const __m512i d = _mm512_cvtepu8_epi64(_mm_cvtsi64_si128((long long)w));
            v = _mm512_mask_mov_epi64(_mm512_add_epi64(d, vbase), _mm512_cmpeq_epi64_mask(d, vf8), vm1);
        } else if (kind == 1 || kind == 2) {
            uint64_t w;
            memcpy(&w, (const uint8_t*)src + i, 8);
            __m128i b = _mm_cvtsi64_si128((long long)w);
            if (kind == 2) b = _mm_srli_epi16(b, 4);
            v = _mm512_cvtepu8_epi64(_mm_and_si128(b, nib));
        } else {
            const __m512i d = _mm512_cvtepu16_epi64(_mm_loadu_si128((const __m128i*)((const uint16_t*)src + i)));
            v = _mm512_mask_mov_epi64(_mm512_add_epi64(d, vbase), _mm512_cmpeq_epi64_mask(d, vff), vm1);
        }
        _mm512_stream_si512((__m512i*)(dst + i), v);
    }
    const size_t esz = kind == 0 ? 4 : kind == 3 ? 2 : 1;
    widen_scalar(kind, (const char*)src + i * esz, dst + i, n - i, base);
}

typedef void (*widen_fn)(int, const void*, i64*, size_t, i64);
static widen_fn pick_widen() {
    const char* e = getenv("UGS_B200_HOST_ISA");  // "scalar" | "avx2" | "avx512": tests force every variant
    __builtin_cpu_init();
    const bool has512 = __builtin_cpu_supports("avx512f"), has2 = __builtin_cpu_supports("avx2");
    if (e && !strcmp(e, "scalar")) return widen_scalar;
    if (e && !strcmp(e, "avx2") && has2) return widen_avx2;
    if (has512) return widen_avx512;
    if (has2) return widen_avx2;
    return widen_scalar;
}
extern "C" const char* ugs_host_isa(void) {
    widen_fn f = pick_widen();
    return f == widen_avx512 ? "avx512" : f == widen_avx2 ? "avx2" : "scalar";
}

// Host-only self-test hook (tests/test_capi_and_host.py checks every ISA variant against numpy without a GPU):
// widens n elements of `src` into dst with the named variant; returns -1 when this CPU lacks it.
extern "C" int ugs_host_widen(int kind, const void* src, int64_t* dst, size_t n, int64_t base, const char* isa) {
    __builtin_cpu_init();
    widen_fn f = widen_scalar;
    if (isa && !strcmp(isa, "avx2")) { if (!__builtin_cpu_supports("avx2")) return -1; f = widen_avx2; }
    else if (isa && !strcmp(isa, "avx512")) { if (!__builtin_cpu_supports("avx512f")) return -1; f = widen_avx512; }
    if (kind < 0 || kind > 4) return -2;
    f(kind, src, dst, n, base);
    _mm_sfence();
    return 0;
}

// ---- worker pool ----------------------------------------------------------------------------------------------------
struct Job {
    int T = 1;                                  // slices; slice t handles the t-th part of every item
    int next = 0, done = 0;                     // guarded by the pool mutex
    void (*run)(void* ctx, int t, int T) = nullptr;
    void* ctx = nullptr;
};

class Pool {
  public:
    void ensure(int n) {
        std::lock_guard<std::mutex> l(mu_);
        while ((int)th_.size() < n) th_.emplace_back([this] { worker(); });
    }
    // runs job->run(ctx, t, T) for t in [0, T) on the pool's threads and the calling thread; returns when all are done
    void run(Job* j) {
        {
            std::lock_guard<std::mutex> l(mu_);
            q_.push_back(j);
        }
        cv_.notify_all();
        while (true) {
            int t = -1;
            {
                std::lock_guard<std::mutex> l(mu_);
                if (j->next < j->T) { t = j->next++; if (j->next == j->T) drop(j); }
            }
            if (t < 0) break;
            j->run(j->ctx, t, j->T);
            std::lock_guard<std::mutex> l(mu_);
            j->done++;
        }
        std::unique_lock<std::mutex> l(mu_);
        done_cv_.wait(l, [j] { return j->done == j->T; });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.detach();  // process exit: the workers are parked in cv_.wait or gone
    }

  private:
    void drop(Job* j) {
        for (auto it = q_.begin(); it != q_.end(); ++it) if (*it == j) { q_.erase(it); break; }
    }
    void worker() {
        std::unique_lock<std::mutex> l(mu_);
        while (true) {
            cv_.wait(l, [this] { return stop_ || !q_.empty(); });
            if (stop_) return;
            Job* j = q_.front();
            const int t = j->next++;
            if (j->next == j->T) drop(j);
            l.unlock();
            j->run(j->ctx, t, j->T);
            l.lock();
            j->done++;
            if (j->done == j->T) done_cv_.notify_all();  // the job may be destroyed by its owner from here on
        }
    }
    std::vector<std::thread> th_;
    std::deque<Job*> q_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    bool stop_ = false;
};
static Pool& pool() {
    static Pool* p = new Pool();  // intentionally leaked: worker threads may outlive static destruction
    return *p;
}

// ---- events (per device) --------------------------------------------------------------------------------
static std::mutex g_dev_mu;
static std::map<int, std::vector<cudaEvent_t>> g_free_events;

static int take_events(int dev, size_t n, std::vector<cudaEvent_t>& out) {
    std::lock_guard<std::mutex> l(g_dev_mu);
    auto& fl = g_free_events[dev];
    while (out.size() < n) {
        cudaEvent_t e;
        if (!fl.empty()) { e = fl.back(); fl.pop_back(); }
        else if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return -1;
        out.push_back(e);
    }
    return 0;
}
static void give_events(int dev, std::vector<cudaEvent_t>& ev) {
    std::lock_guard<std::mutex> l(g_dev_mu);
    auto& fl = g_free_events[dev];
    fl.insert(fl.end(), ev.begin(), ev.end());
    ev.clear();
}
static long long wire_chunk_bytes() {
    static long long v = 0;
    if (!v) {
        const char* e = getenv("UGS_B200_WIRE_CHUNK_MB");
        long long mb = e ? atoll(e) : 0;
        if (mb < 1 || mb > 1024) mb = 8;  // measured (profiles/r02_chunk_sweep.txt): 1 MB 54.4 ms, 4-8 MB 50.6, 32 MB 52.2, 64 MB 53.3
        v = mb << 20;
    }
    return v;
}

#define CK(expr)                                                                                  \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) { rc = ugs_set_error(UGS_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); goto out; } \
    } while (0)

struct Item {
    int kind;          // widen kind (0..3)
    const char* src;   // host wire address of element 0 of the item
    i64* dst;
    size_t n;          // elements
    size_t first;      // kind 3 / 4: index of element 0 inside the packed part (block = first / blk)
    const i64* bases;  // kind 3 / 4
    size_t blk;        // kind 3: UGS_DELTA_BLOCK, kind 4: UGS_DELTA8_BLOCK
    cudaEvent_t ev;
};
struct DeliverCtx {
    std::vector<Item>* items;
    widen_fn fn;
    int dev;
    std::atomic<int> cuda_error{0};
};

static void deliver_slice(void* ctx_, int t, int T) {
    DeliverCtx* c = (DeliverCtx*)ctx_;
    cudaSetDevice(c->dev);
    for (const Item& it : *c->items) {
        const cudaError_t e = cudaEventSynchronize(it.ev);
        if (e != cudaSuccess) { c->cuda_error.store((int)e); return; }
        // 8-element granularity: every slice starts on a 64-byte line of its destination array (when the array does)
        const size_t per = ((it.n + T - 1) / T + 7) & ~(size_t)7;
        const size_t a = (size_t)t * per, b = (a + per < it.n) ? a + per : it.n;
        if (a >= b) continue;
        if (it.kind != 3 && it.kind != 4) {
            const size_t esz = it.kind == 0 ? 4 : 1;
            c->fn(it.kind, it.src + a * esz, it.dst + a, b - a, 0);
        } else {  // one base per block of it.blk values
            const size_t esz = it.kind == 3 ? 2 : 1;
            size_t i = a;
            while (i < b) {
                const size_t blk = (it.first + i) / it.blk;
                size_t end = (blk + 1) * it.blk - it.first;
                if (end > b) end = b;
                c->fn(it.kind, it.src + i * esz, it.dst + i, end - i, it.bases[blk]);
                i = end;
            }
        }
    }
    _mm_sfence();
}

extern "C" size_t ugs_deliver_wire_bytes(int64_t B, int k, int64_t total, int pairs_u8) {
    if (B < 0 || total < 0 || k < 1) return 0;
    return ugs_wire_layout((size_t)B * k, (size_t)total, (size_t)total, pairs_u8, 0).total;
}

// prepacked: the edge part of d_wire (one byte per entry, 16-bit source-column deltas and their block bases) was written
// by ugs_emit_edges_wire; only the node ids are packed here. No DMA share, no int32 fallback: a value that does not fit
// is reported as UGS_ELIMIT and the caller takes the int64 path.
static int deliver_impl(const int64_t* nodes, int64_t B, int k, const int64_t* edge_index, const int64_t* edge_src,
                        int64_t total, int pairs_u8, double dma_fraction, void* d_wire, void* h_wire, size_t wire_bytes,
                        int64_t* h_nodes, int64_t* h_edge_index, int64_t* h_edge_src, int n_threads, int64_t* h_stats,
                        void* stream_, int prepacked);

extern "C" int ugs_deliver_outputs(const int64_t* nodes, int64_t B, int k, const int64_t* edge_index,
                                   const int64_t* edge_src, int64_t total, int pairs_u8, double dma_fraction,
                                   void* d_wire, void* h_wire, size_t wire_bytes, int64_t* h_nodes, int64_t* h_edge_index,
                                   int64_t* h_edge_src, int n_threads, int64_t* h_stats, void* stream_) {
    return deliver_impl(nodes, B, k, edge_index, edge_src, total, pairs_u8, dma_fraction, d_wire, h_wire, wire_bytes, h_nodes,
                        h_edge_index, h_edge_src, n_threads, h_stats, stream_, 0);
}

extern "C" int ugs_deliver_outputs_wire(const int64_t* nodes, int64_t B, int k, int64_t total, void* d_wire, void* h_wire,
                                        size_t wire_bytes, int64_t* h_nodes, int64_t* h_edge_index, int64_t* h_edge_src,
                                        int n_threads, int64_t* h_stats, void* stream_) {
    return deliver_impl(nodes, B, k, nullptr, nullptr, total, 1, 0.0, d_wire, h_wire, wire_bytes, h_nodes, h_edge_index,
                        h_edge_src, n_threads, h_stats, stream_, 1);
}

static int deliver_impl(const int64_t* nodes, int64_t B, int k, const int64_t* edge_index, const int64_t* edge_src,
                        int64_t total, int pairs_u8, double dma_fraction, void* d_wire, void* h_wire, size_t wire_bytes,
                        int64_t* h_nodes, int64_t* h_edge_index, int64_t* h_edge_src, int n_threads, int64_t* h_stats,
                        void* stream_, int prepacked) {
    if (B < 0 || total < 0 || k < 1) return ugs_set_error(UGS_EINVAL, "bad size");
    if ((B * k > 0 && (!nodes || !h_nodes)) ||
        (total > 0 && ((!prepacked && (!edge_index || !edge_src)) || !h_edge_index || !h_edge_src)))
        return ugs_set_error(UGS_EINVAL, "NULL array");
    cudaStream_t st = (cudaStream_t)stream_;
    int rc = UGS_OK, dev = 0;
    cudaGetDevice(&dev);
    const size_t NN = (size_t)B * k, NE = (size_t)total;
    double f = dma_fraction;
    if (!(f >= 0.0)) f = 0.0;
    if (f > 1.0) f = 1.0;
    // prefix [0, s) of every array goes by DMA (whole 4096-value blocks, so that the packed part starts on a block)
    auto split = [&](size_t n) { size_t s = (size_t)((double)n * f); s &= ~(size_t)(UGS_DELTA_BLOCK - 1); return f >= 1.0 ? n : s; };
    const size_t sn = split(NN), se = split(NE), si = split(NE);
    const size_t pn = NN - sn, pe = NE - se, pi = NE - si;  // packed elements
    std::vector<cudaEvent_t> events;
    std::vector<Item> items;
    int delta16 = 1;
    UgsWire w = prepacked ? ugs_wire_layout_emit(pn, pe) : ugs_wire_layout(pn, pe, pi, pairs_u8, 1);
    if (pn + pe + pi > 0) {  // the int32 layout is the larger one: the buffers must hold it in case the deltas overflow
        if (!d_wire || !h_wire) return ugs_set_error(UGS_EINVAL, "NULL wire buffer");
        if (wire_bytes < (prepacked ? w.total : ugs_wire_layout(pn, pe, pi, pairs_u8, 0).total))
            return ugs_set_error(UGS_EINVAL, "wire buffer too small");
    }
    char* hw = (char*)h_wire;
    char* dw = (char*)d_wire;
    size_t wire_copied = 0;

    // ---- DMA pipeline: the int64 prefixes as a queue of byte ranges. A copy engine serves one copy at a time, so big
    //      DMA copies in front of the wire chunks would starve the widening threads (measured: the two pipelines then run
    //      one after the other); the ranges are therefore issued on the caller's stream in slices BETWEEN the wire
    //      chunks, each slice sized so that DMA and wire traffic finish together.
    struct Seg { const char* src; char* dst; size_t bytes; };
    std::vector<Seg> segs;
    if (si) {
        segs.push_back({(const char*)edge_index, (char*)h_edge_index, si * 8});
        segs.push_back({(const char*)(edge_index + NE), (char*)(h_edge_index + NE), si * 8});
    }
    if (se) segs.push_back({(const char*)edge_src, (char*)h_edge_src, se * 8});
    if (sn) segs.push_back({(const char*)nodes, (char*)h_nodes, sn * 8});
    size_t seg_i = 0, dma_left = 0;
    for (const Seg& sg : segs) dma_left += sg.bytes;
    const size_t dma_total = dma_left;
    auto issue_dma = [&](size_t quota) -> cudaError_t {  // copies the next `quota` bytes of the queue
        while (quota > 0 && seg_i < segs.size()) {
            Seg& sg = segs[seg_i];
            const size_t nb = sg.bytes < quota ? sg.bytes : quota;
            const cudaError_t e = cudaMemcpyAsync(sg.dst, sg.src, nb, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return e;
            sg.src += nb; sg.dst += nb; sg.bytes -= nb; quota -= nb; dma_left -= nb;
            if (sg.bytes == 0) ++seg_i;
        }
        return cudaSuccess;
    };
    // ---- widen pipeline -----------------------------------------------------------------------------------------------
    if (pn + pe + pi > 0) {
        if (prepacked) rc = ugs_pack_parts(nodes, pn, nullptr, nullptr, 0, nullptr, 0, 1, 1, dw, &w, st);
        else rc = ugs_pack_parts(nodes + sn, pn, edge_index + si, edge_index + NE + si, pi, edge_src + se, pe, pairs_u8, 1, dw, &w, st);
        if (rc != UGS_OK) goto out;
        // block minima + the overflow counters first: they decide the format and every widening thread needs the minima
        CK(cudaMemcpyAsync(hw + w.nodes_base, dw + w.nodes_base, w.total - w.nodes_base, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        {
            const unsigned long long* bad = (const unsigned long long*)(hw + w.bad);
            if (bad[1]) { rc = ugs_set_error(UGS_ELIMIT, "%llu edge entries do not fit the packed wire format", bad[1]); goto out; }
            if (prepacked && (bad[0] | bad[2])) {
                rc = ugs_set_error(UGS_ELIMIT, "%llu node ids do not fit a 16-bit / %llu source columns an 8-bit delta", bad[0], bad[2]);
                goto out;
            }
            if (bad[0]) {  // some 4096-value block spreads over more than 65534: int32 for node ids / source columns
                delta16 = 0;
                w = ugs_wire_layout(pn, pe, pi, pairs_u8, 0);
                rc = ugs_pack_parts(nodes + sn, pn, edge_index + si, edge_index + NE + si, pi, edge_src + se, pe, pairs_u8, 0, dw, &w, st);
                if (rc != UGS_OK) goto out;
                CK(cudaMemcpyAsync(hw + w.bad, dw + w.bad, 16, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                const unsigned long long* bad32 = (const unsigned long long*)(hw + w.bad);  // the int32 layout's counters
                if (bad32[0] | bad32[1]) { rc = ugs_set_error(UGS_ELIMIT, "output values do not fit int32"); goto out; }
            }
        }
        // twin_dst: the chunk is widened a second time (high nibbles -> row 1 of edge_index) right after the first
        auto add_region = [&](int kind, size_t off, size_t n, size_t esz, i64* dst, size_t bases_off, i64* twin_dst = nullptr) {
            const size_t per = (size_t)wire_chunk_bytes() / esz;
            for (size_t a = 0; a < n; a += per) {
                Item it;
                it.kind = kind; it.src = hw + off + a * esz; it.dst = dst + a; it.n = (n - a < per) ? n - a : per;
                it.first = a; it.bases = (const i64*)(hw + bases_off); it.ev = nullptr;
                it.blk = kind == 4 ? UGS_DELTA8_BLOCK : UGS_DELTA_BLOCK;
                items.push_back(it);
                if (twin_dst) { it.kind = 2; it.dst = twin_dst + a; items.push_back(it); }
            }
        };
        // biggest consumers of host write bandwidth first: the tail of the transfer then overlaps the least work
        if (pairs_u8) {
            add_region(1, w.ei, pi, 1, h_edge_index + si, 0, h_edge_index + NE + si);
        } else {
            add_region(0, w.ei, pi, 4, h_edge_index + si, 0);
            add_region(0, w.ei + pi * 4, pi, 4, h_edge_index + NE + si, 0);
        }
        if (prepacked) add_region(4, w.src, pe, 1, h_edge_src + se, w.src_base);
        else add_region(delta16 ? 3 : 0, w.src, pe, delta16 ? 2 : 4, h_edge_src + se, w.src_base);
        add_region(delta16 ? 3 : 0, w.nodes, pn, delta16 ? 2 : 4, h_nodes + sn, w.nodes_base);
        const size_t ev0 = events.size();
        if (take_events(dev, ev0 + items.size(), events) != 0) { rc = ugs_set_error(UGS_ECUDA, "cudaEventCreate failed"); goto out; }
        size_t n_chunks = 0;
        for (const Item& it : items) n_chunks += it.kind != 2;
        const size_t quota = (dma_total + n_chunks - 1) / (n_chunks ? n_chunks : 1);
        for (size_t i = 0; i < items.size(); ++i) {
            if (items[i].kind == 2) { items[i].ev = items[i - 1].ev; continue; }  // twin: same chunk, same event
            const size_t esz = items[i].kind == 0 ? 4 : items[i].kind == 3 ? 2 : 1;
            const size_t off = (size_t)(items[i].src - hw);
            CK(cudaMemcpyAsync(hw + off, dw + off, items[i].n * esz, cudaMemcpyDeviceToHost, st));
            wire_copied += items[i].n * esz;
            items[i].ev = events[ev0 + i];
            CK(cudaEventRecord(items[i].ev, st));
            CK(issue_dma(quota));  // the DMA pipeline's share of the link until the next chunk
        }
        CK(issue_dma(dma_left));
        {
            int T = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
            static widen_fn fn = pick_widen();
            DeliverCtx ctx;
            ctx.items = &items; ctx.fn = fn; ctx.dev = dev;
            Job job;
            job.T = T; job.run = deliver_slice; job.ctx = &ctx;
            pool().ensure(T - 1);
            pool().run(&job);
            if (ctx.cuda_error.load()) {
                rc = ugs_set_error(UGS_ECUDA, "cudaEventSynchronize failed in a delivery thread: %s",
                                   cudaGetErrorString((cudaError_t)ctx.cuda_error.load()));
                goto out;
            }
        }
    }
    CK(issue_dma(dma_left));  // everything, when nothing was packed
    CK(cudaStreamSynchronize(st));
    if (h_stats) {
        h_stats[0] = (i64)(wire_copied + (pn + pe + pi > 0 ? w.total - w.nodes_base : 0));  // packed bytes over PCIe
        h_stats[1] = (i64)dma_total;                                                          // int64 bytes by DMA
        h_stats[2] = delta16;
        h_stats[3] = (i64)items.size();
    }
out:
    if (rc != UGS_OK) cudaStreamSynchronize(st);  // leave nothing in flight that still writes into the caller's buffers
    give_events(dev, events);
    return rc;
}

// ---- pageable inputs: parallel copy into a pinned staging buffer -----------------------------------------------------
// The trainer hands over pageable `.cpu()` tensors (experiment.py:877-878). A cudaMemcpy from pageable memory is staged
// by the driver through its own pinned buffer on one thread (~12 GB/s: 8 ms for the 106 MB of the ZINC-shaped shard);
// n_threads threads of the delivery pool fill a pinned buffer of the caller several times faster, and the H2D copy from
// there runs at the link rate.
struct CopyCtx { char* dst; const char* src; size_t bytes; };
static void copy_slice(void* ctx_, int t, int T) {
    CopyCtx* c = (CopyCtx*)ctx_;
    const size_t per = ((c->bytes + T - 1) / T + 4095) & ~(size_t)4095;
    const size_t a = (size_t)t * per, b = (a + per < c->bytes) ? a + per : c->bytes;
    if (a < b) memcpy(c->dst + a, c->src + a, b - a);
}
extern "C" int ugs_host_copy(void* dst, const void* src, size_t bytes, int n_threads) {
    if (!dst || !src) return ugs_set_error(UGS_EINVAL, "NULL buffer");
    int T = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    if (bytes < ((size_t)1 << 20)) T = 1;
    CopyCtx ctx{(char*)dst, (const char*)src, bytes};
    Job job;
    job.T = T; job.run = copy_slice; job.ctx = &ctx;
    pool().ensure(T - 1);
    pool().run(&job);
    return UGS_OK;
}

// ---- what the box's memory system takes: non-temporal stores from n_threads threads -----------------------------------
struct BwCtx { char* dst; size_t bytes; widen_fn fn; };
static void bw_slice(void* ctx_, int t, int T) {
    BwCtx* c = (BwCtx*)ctx_;
    const size_t n = c->bytes / 8;
    const size_t per = ((n + T - 1) / T + 7) & ~(size_t)7;
    const size_t a = (size_t)t * per, b = (a + per < n) ? a + per : n;
    if (a >= b) return;
    // the delivery's dominant conversion: one source byte -> one int64 (source: a 4 KB block re-read from L1)
    static const uint8_t srcbuf[4096] = {1};
    for (size_t i = a; i < b; i += 4096) c->fn(1, srcbuf, (i64*)c->dst + i, (b - i < 4096) ? b - i : 4096, 0);
    _mm_sfence();
}
extern "C" double ugs_host_store_bandwidth(void* dst, size_t bytes, int n_threads, int reps) {
    if (!dst || bytes < 4096) return 0.0;
    const int T = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    static widen_fn fn = pick_widen();
    BwCtx ctx{(char*)dst, bytes, fn};
    pool().ensure(T - 1);
    double best = 0.0;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
        Job job;
        job.T = T; job.run = bw_slice; job.ctx = &ctx;
        const auto t0 = std::chrono::steady_clock::now();
        pool().run(&job);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (s > 0 && bytes / s / 1e9 > best) best = bytes / s / 1e9;
    }
    return best;
}
