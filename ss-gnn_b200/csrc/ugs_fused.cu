// ugs_fused.cu -- the whole sample_batch pipeline as ONE library call with ONE host round trip.
//
// The trainer calls ugs_sampler.sample_batch once per DataLoader mini-batch (gps/experiment.py:882-883; 32..1024 graphs,
// reference timing script tests/test_rwr_profiler.py:45-57 uses 32): a few hundred microseconds of work for the
// reference's CPU loop, a few microseconds per kernel here -- such calls are bound by launch latency and host round
// trips, not by any kernel. The staged sequence (ugs_segment_batch ... ugs_emit_edges, include/ugs_b200.h) needs the
// host twice: to read the class counts before the general-path kernels can be sized, and to read the number of emitted
// entries before the outputs can be allocated. For batches whose worst-case output is small this entry point removes
// both: the general-path kernels are launched with worst-case grids and read their item counts from counters[] on the
// device, and the emission pass writes into caller-provided buffers sized for the upper bound
//     cap = m * (number of CSR entries of the batch) = m * 2E        (a sample emits each CSR entry at most once),
// so that the only host synchronisation is the final read of counters[] (total entries, flags, statistics).
// ugs_fused_finish then narrows the three bound-sized arrays into the exact [2,total] / [total] tensors of the
// reference ABI. Results are bit-identical to the staged sequence (tests/test_gpu_fused.py).
#include "ugs_common.cuh"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

// UGS_B200_FUSED_TRACE=1: events between the stages of one fused call, printed to stderr after its synchronisation
// (where the ~0.1 ms of a DataLoader-sized call goes: stage times INCLUDE the launch gaps in front of their kernels).
struct FusedTrace {
    bool on = false;
    cudaEvent_t ev[6];
    std::chrono::steady_clock::time_point t0;
    FusedTrace() {
        const char* e = getenv("UGS_B200_FUSED_TRACE");
        on = e && e[0] && e[0] != '0';
    }
    void begin(cudaStream_t st) {
        if (!on) return;
        static thread_local bool made = false;
        static thread_local cudaEvent_t evs[6];
        if (!made) { for (auto& x : evs) cudaEventCreate(&x); made = true; }
        for (int i = 0; i < 6; ++i) ev[i] = evs[i];
        t0 = std::chrono::steady_clock::now();
        cudaEventRecord(ev[0], st);
    }
    void mark(int i, cudaStream_t st) { if (on) cudaEventRecord(ev[i], st); }
    void end(double queued_us) {
        if (!on) return;
        const double wall = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        float d[5];
        for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&d[i], ev[i], ev[i + 1]);
        fprintf(stderr, "[ugs fused trace] segment+classify %.1f us | preprocess %.1f | sample+scan %.1f | emit %.1f | "
                        "counters to host %.1f | device total %.1f us; host: queued after %.1f us, returned after %.1f us\n",
                d[0] * 1e3, d[1] * 1e3, d[2] * 1e3, d[3] * 1e3, d[4] * 1e3, (d[0] + d[1] + d[2] + d[3] + d[4]) * 1e3,
                queued_us, wall);
    }
};

extern "C" int64_t ugs_fused_capacity(int64_t E, int64_t m) {
    if (E <= 0 || m <= 0) return 1;
    return 2 * E * m;
}

// h_info: pinned host memory, UGS_CTR_COUNT int64. On return it holds counters[] (the emission may still be running):
// [UGS_CTR_TOTAL_EDGES] = entries emitted, [UGS_CTR_UNSORTED] = segmentation flags -- when any bit is set NOTHING in the
// outputs is valid and the caller must take the staged path (columns not graph-contiguous, ptr decreasing).
// ei_u / ei_v / ei_src: device int64[cap] each, cap >= ugs_fused_capacity(E, m).
extern "C" int ugs_sample_batch_fused(const int64_t* edge_index, int64_t E, const int64_t* ptr, int64_t G,
                                      int64_t node_base, int64_t total_nodes, int64_t m, int k, int mode,
                                      uint64_t seed, uint64_t graph_index_base, uint64_t sample_index_base,
                                      int seed_mode, void* workspace, const ugs_layout* L, void* scratch,
                                      size_t scratch_bytes, int64_t* nodes, int64_t* edge_ptr, int64_t* sample_ptr,
                                      int64_t* ei_u, int64_t* ei_v, int64_t* ei_src, int64_t cap, int64_t* h_info,
                                      int flags, void* stream_) {
    if (!workspace || !L || !h_info) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (G <= 0 || m <= 0) return ugs_set_error(UGS_EINVAL, "the fused path needs G > 0 and m > 0");
    if (cap < ugs_fused_capacity(E, m)) return ugs_set_error(UGS_EINVAL, "output capacity below m * 2E");
    if (ei_v != ei_u + cap) return ugs_set_error(UGS_EINVAL, "ei_v must be ei_u + cap");
    if ((flags & UGS_FUSED_COMPACT) && ei_src != ei_v + cap) return ugs_set_error(UGS_EINVAL, "compact: ei_src must be ei_v + cap");
    const bool assume_no_general = (flags & UGS_FUSED_ASSUME_NO_GENERAL) != 0;
    cudaStream_t st = (cudaStream_t)stream_;
    FusedTrace tr;
    tr.begin(st);
    int rc = ugs_segment_batch(edge_index, E, ptr, G, k, 1, 1, nullptr, workspace, L, stream_);
    if (rc != UGS_OK) return rc;
    tr.mark(1, st);
    // general-path kernels only where a general-path graph can exist at all (host-known bounds)
    // assume_no_general: the caller's guess (the previous batch of this data set had none) drops the five general-path
    // launches -- they return at once when the device count is zero, but cost ~3 us each in a 0.1 ms chain. The guess is
    // checked by the caller against h_info[UGS_CTR_N_GENERAL]: non-zero means the outputs are void and the call must be
    // repeated without the assumption.
    const bool maybe_general = !assume_no_general && ((k > 16) || (total_nodes > UGS_SMALL_MAX_N) || (E > UGS_SMALL_MAX_COLS));
    const int64_t ngen = maybe_general ? -1 : 0;
    const int64_t nlarge = (maybe_general && total_nodes > UGS_LARGE_N) ? -1 : 0;
    rc = ugs_preproc_batch_parts(edge_index, E, ptr, G, node_base, k, nullptr, ngen, nlarge, UGS_PREPROC_ALL, workspace,
                                 L, stream_);
    if (rc != UGS_OK) return rc;
    tr.mark(2, st);
    // The host waits for counters[] only, and those are final once the scan kernel has the total: its last tile writes
    // them into h_info through the mapped address (pinned host memory is device-visible under unified addressing),
    // sequence word last, and the host polls that word instead of synchronising the stream behind the emission kernels.
    static std::atomic<int64_t> next_seq{1};
    const int64_t seq = next_seq.fetch_add(1);
    int64_t* snap = nullptr;
    {
        void* dptr = nullptr;
        if (!tr.on && cudaHostGetDevicePointer(&dptr, h_info, 0) == cudaSuccess) snap = (int64_t*)dptr;
        else (void)cudaGetLastError();  // not mapped: read the counters back with a copy after the emission
    }
    if (snap) reinterpret_cast<volatile int64_t*>(h_info)[UGS_CTR_COUNT - 1] = 0;
    rc = ugs_sample_batch_impl(ptr, G, node_base, m, k, seed, graph_index_base, sample_index_base, seed_mode,
                               UGS_VARIANT_REF, ngen, total_nodes, workspace, L, scratch, scratch_bytes, nodes,
                               edge_ptr, sample_ptr, nullptr, stream_, /*counters_fresh=*/true, snap, seq);
    if (rc != UGS_OK) return rc;
    tr.mark(3, st);
    const int eflags = (assume_no_general ? UGS_EMIT_GUARD_GENERAL : 0) | ((flags & UGS_FUSED_COMPACT) ? UGS_EMIT_COMPACT : 0);
    rc = ugs_emit_edges_impl(ptr, G, node_base, m, k, mode, 0, ngen, workspace, L, nodes, edge_ptr, cap, ei_u, ei_src, scratch,
                             scratch_bytes, eflags, stream_);
    if (rc != UGS_OK) return rc;
    tr.mark(4, st);
    if (snap) {
        volatile int64_t* flag = reinterpret_cast<volatile int64_t*>(h_info) + (UGS_CTR_COUNT - 1);
        for (unsigned spins = 1; *flag != seq; ++spins) {
#if defined(__x86_64__)
            _mm_pause();
#endif
            if ((spins & 0x3FFFu) == 0) {  // a faulted kernel never writes the word: surface the error instead of spinning
                const cudaError_t q = cudaStreamQuery(st);
                if (q == cudaErrorNotReady) continue;
                if (*flag == seq) break;
                if (q != cudaSuccess) return ugs_set_error(UGS_ECUDA, "fused call failed: %s", cudaGetErrorString(q));
                return ugs_set_error(UGS_ECUDA, "fused call finished without delivering its counters");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        return UGS_OK;
    }
    UGS_CUDA_CHECK(cudaMemcpyAsync(h_info, (const char*)workspace + L->counters, sizeof(i64) * UGS_CTR_COUNT,
                                   cudaMemcpyDeviceToHost, st));
    tr.mark(5, st);
    const double queued_us = tr.on ? std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tr.t0).count() : 0.0;
    UGS_CUDA_CHECK(cudaStreamSynchronize(st));
    tr.end(queued_us);
    return UGS_OK;
}

// One launch instead of three cudaMemcpyAsync (12 us of host time each): the three arrays, 16 bytes per thread and trip.
__global__ void __launch_bounds__(256)
k_fused_narrow(const i64* __restrict__ a0, const i64* __restrict__ a1, const i64* __restrict__ a2, i64 total,
               i64* __restrict__ o0, i64* __restrict__ o1, i64* __restrict__ o2) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 pairs = total >> 1;
    // sources start 16-byte aligned (cap is even and the base is 256-byte aligned); o1 = o0 + total is 16-byte aligned
    // only for even totals, so rows 1 and 2 of odd totals take the scalar path
    const bool vec12 = (total & 1) == 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
        reinterpret_cast<longlong2*>(o0)[i] = reinterpret_cast<const longlong2*>(a0)[i];
        if (vec12) {
            reinterpret_cast<longlong2*>(o1)[i] = reinterpret_cast<const longlong2*>(a1)[i];
            reinterpret_cast<longlong2*>(o2)[i] = reinterpret_cast<const longlong2*>(a2)[i];
        } else {
            o1[2 * i] = a1[2 * i]; o1[2 * i + 1] = a1[2 * i + 1];
            o2[2 * i] = a2[2 * i]; o2[2 * i + 1] = a2[2 * i + 1];
        }
    }
    if ((total & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        o0[total - 1] = a0[total - 1]; o1[total - 1] = a1[total - 1]; o2[total - 1] = a2[total - 1];
    }
}

// edge_index_out[2,total] (row 1 at +total) and edge_src_out[total] from the bound-sized arrays of the call above.
// edge_index_out must be 16-byte aligned and edge_src_out = edge_index_out + 2*total or 16-byte aligned for even totals.
extern "C" int ugs_fused_finish(const int64_t* ei_u, const int64_t* ei_v, const int64_t* ei_src, int64_t total,
                                int64_t* edge_index_out, int64_t* edge_src_out, void* stream_) {
    if (total <= 0) return UGS_OK;
    if (!ei_u || !ei_v || !ei_src || !edge_index_out || !edge_src_out) return ugs_set_error(UGS_EINVAL, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream_;
    if ((((uintptr_t)ei_u | (uintptr_t)ei_v | (uintptr_t)ei_src | (uintptr_t)edge_index_out) & 15) ||
        (!(total & 1) && ((uintptr_t)edge_src_out & 15))) {
        const size_t bytes = sizeof(i64) * (size_t)total;
        UGS_CUDA_CHECK(cudaMemcpyAsync(edge_index_out, ei_u, bytes, cudaMemcpyDeviceToDevice, st));
        UGS_CUDA_CHECK(cudaMemcpyAsync(edge_index_out + total, ei_v, bytes, cudaMemcpyDeviceToDevice, st));
        UGS_CUDA_CHECK(cudaMemcpyAsync(edge_src_out, ei_src, bytes, cudaMemcpyDeviceToDevice, st));
        return UGS_OK;
    }
    int64_t blocks = ((total >> 1) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_fused_narrow<<<(unsigned)blocks, 256, 0, st>>>(ei_u, ei_v, ei_src, total, edge_index_out, edge_index_out + total,
                                                    edge_src_out);
    UGS_LAUNCH_CHECK("k_fused_narrow");
    return UGS_OK;
}
