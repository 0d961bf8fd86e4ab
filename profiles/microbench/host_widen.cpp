// Host-side widening bandwidth: how fast can N threads expand compact device output (u8 / i32) into the int64 arrays
// of the reference ABI? Regular stores vs non-temporal (streaming) stores.   g++ -O3 -march=native -pthread
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>
#include <thread>
#include <vector>

template <typename S, bool NT>
static void widen(const S* src, int64_t* dst, size_t n) {
    size_t i = 0;
    if (NT) {
        for (; i < n; ++i) _mm_stream_si64((long long*)(dst + i), (long long)src[i]);
        _mm_sfence();
    } else {
        for (; i < n; ++i) dst[i] = (int64_t)src[i];
    }
}

template <typename S, bool NT>
static double run(const S* src, int64_t* dst, size_t n, int threads) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) {
        size_t a = n * t / threads, b = n * (t + 1) / threads;
        th.emplace_back([=] { widen<S, NT>(src + a, dst + a, b - a); });
    }
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv) {
    const size_t n = (size_t)1 << 28;
    const int threads = argc > 1 ? atoi(argv[1]) : (int)std::thread::hardware_concurrency();
    uint8_t* s8 = (uint8_t*)malloc(n);
    int32_t* s32 = (int32_t*)malloc(n * 4);
    int64_t* dst = (int64_t*)aligned_alloc(64, n * 8);
    memset(s8, 1, n); memset(s32, 1, n * 4); memset(dst, 0, n * 8);
    for (int rep = 0; rep < 2; ++rep) {
        printf("threads %d  u8->i64  regular %.1f GB/s   streaming %.1f GB/s\n", threads,
               n * 8 / run<uint8_t, false>(s8, dst, n, threads) / 1e9, n * 8 / run<uint8_t, true>(s8, dst, n, threads) / 1e9);
        printf("threads %d  i32->i64 regular %.1f GB/s   streaming %.1f GB/s\n", threads,
               n * 8 / run<int32_t, false>(s32, dst, n, threads) / 1e9, n * 8 / run<int32_t, true>(s32, dst, n, threads) / 1e9);
    }
    return 0;
}
