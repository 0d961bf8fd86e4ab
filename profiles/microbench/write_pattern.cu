// Microbenchmark: how fast can B200 absorb the emit kernel's store pattern with no computation at all?
// (a) one streaming array, (b) three arrays written in per-warp runs of RUN 8-byte entries (the emit pattern),
// (c) the same with 16-byte stores. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 write_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef long long i64;

__global__ void stream1(i64* a, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) a[i] = (i64)i;
}
template <int VEC>
__global__ void runs3(i64* u, i64* v, i64* s, size_t n, int run) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t base = warp * (size_t)run;
    if (base >= n) return;
    if (VEC == 1) {
        for (int e = lane; e < run && base + e < n; e += 32) {
            u[base + e] = e & 7; v[base + e] = (e >> 3) & 7; s[base + e] = base + e;
        }
    } else {
        for (int e = 2 * lane; e < run && base + e + 1 < n; e += 64) {
            *reinterpret_cast<longlong2*>(u + base + e) = make_longlong2(e & 7, (e + 1) & 7);
            *reinterpret_cast<longlong2*>(v + base + e) = make_longlong2((e >> 3) & 7, ((e + 1) >> 3) & 7);
            *reinterpret_cast<longlong2*>(s + base + e) = make_longlong2(base + e, base + e + 1);
        }
    }
}
int main() {
    const size_t n = 260389600;  // entries of the bench step
    i64 *u, *v, *s;
    cudaMalloc(&u, n * 8); cudaMalloc(&v, n * 8); cudaMalloc(&s, n * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a); stream1<<<148 * 16, 256>>>(u, n); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("stream1 one array   %7.3f ms  %7.1f GB/s\n", ms, n * 8 / ms / 1e6);
        for (int run : {256, 1024, 2048, 8192}) {
            const size_t warps = (n + run - 1) / run;
            const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
            cudaEventRecord(a); runs3<1><<<blocks, 256>>>(u, v, s, n, run); cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            printf("runs3 8B  run=%5d  %7.3f ms  %7.1f GB/s\n", run, ms, 3.0 * n * 8 / ms / 1e6);
            cudaEventRecord(a); runs3<2><<<blocks, 256>>>(u, v, s, n, run); cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            printf("runs3 16B run=%5d  %7.3f ms  %7.1f GB/s\n", run, ms, 3.0 * n * 8 / ms / 1e6);
        }
    }
    return 0;
}
