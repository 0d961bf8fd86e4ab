// ugs_preproc_large.cu -- preprocessing of LARGE graphs (power-law graphs of 10^4..10^5+ nodes) with the whole GPU.
//
// Same results as k_preproc_general (ugs_preproc.cu), i.e. the reference's build_csr / compute_1dd_ordering /
// compute_suffix_and_buckets (preproc.cpp:32-256), but one graph at a time is spread over a
// cooperative grid (one CTA per SM slot, grid-wide barriers between phases) instead of being confined to one CTA:
// a 10^5-node graph has 6*10^5 columns and 1.2*10^6 CSR entries to move, far too much for 1024 threads.
// Phases: degrees (atomics) | multi-CTA scan | atomic fill | per-row sort by source column (warp: <= 128 entries,
// CTA: longer) | order = global bitonic network over (degree << 32 | id) | suffix degree (warp per vertex) |
// bounded BFS reachability (warp per vertex) | counts | root CDF (exact parallel prefix when the sum fits 2^53, else
// the reference's left-to-right sum by one thread).
#include <cooperative_groups.h>
#include "ugs_common.cuh"

namespace cg = cooperative_groups;

#define LG_THREADS 512
#define LG_SORT_SMEM 2048  // u64 keys: CTA-level sort buffer, also 16 warp slices of 128 keys

__device__ __forceinline__ u64 lg_pack(i32 col, i32 idx) { return (((u64)(u32)col) << 32) | (u32)idx; }

// CTA-wide normalised bitonic network on a[0..n) (shared or global memory), any n
__device__ void lg_block_bitonic(u64* a, int n) {
    for (int k2 = 2; (k2 >> 1) < n; k2 <<= 1) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int p = i ^ (k2 - 1);
            if (p > i && p < n) { const u64 x = a[i], y = a[p]; if (y < x) { a[i] = y; a[p] = x; } }
        }
        __syncthreads();
        for (int j = k2 >> 2; j >= 1; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i && p < n) { const u64 x = a[i], y = a[p]; if (y < x) { a[i] = y; a[p] = x; } }
            }
            __syncthreads();
        }
    }
}

// Grid-wide exclusive scan of f(i), i in [0,n): every CTA owns a contiguous chunk; CTA totals go through `bsum`.
// out(i, exclusive_prefix) is called for every i; returns the grand total (to every thread).
template <typename T, typename F, typename O>
__device__ T lg_grid_scan(cg::grid_group& grid, int n, F f, O out, T* bsum, T* sm /* blockDim.x + 1 */) {
    const int nb = gridDim.x, b = blockIdx.x, T_ = blockDim.x;
    const int chunk = (n + nb - 1) / nb;
    const int c0 = min(n, b * chunk), c1 = min(n, c0 + chunk);
    const int len = c1 - c0;
    const int seg = (len + T_ - 1) / T_;
    const int s0 = c0 + min(len, (int)threadIdx.x * seg), s1 = min(c1, s0 + seg);
    T local = 0;
    for (int i = s0; i < s1; ++i) local += f(i);
    sm[threadIdx.x] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        T run = 0;
        for (int t = 0; t < T_; ++t) { const T v = sm[t]; sm[t] = run; run += v; }
        sm[T_] = run;
        bsum[b] = run;
    }
    __syncthreads();
    const T within = sm[threadIdx.x];
    grid.sync();
    // prefix of the preceding CTAs' totals (every thread sums them: nb is a few hundred)
    T pre = 0, total = 0;
    for (int q = 0; q < nb; ++q) { const T v = bsum[q]; if (q < b) pre += v; total += v; }
    T run = pre + within;
    for (int i = s0; i < s1; ++i) { const T v = f(i); out(i, run); run += v; }
    grid.sync();
    return total;
}

__global__ void __launch_bounds__(LG_THREADS)
k_preproc_large(const i64* __restrict__ row0, const i64* __restrict__ row1, const i64* __restrict__ ptr, i64 node_base,
                int k, const i64* __restrict__ perm, i64 n_large, WS w) {
    cg::grid_group grid = cg::this_grid();
    __shared__ u64 sbuf[LG_SORT_SMEM];
    __shared__ double sscan[LG_THREADS + 1];
    const int tid = threadIdx.x, T_ = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, wpb = T_ >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    const i64 gtid = (i64)blockIdx.x * T_ + tid, gsize = (i64)gridDim.x * T_;
    const i64 gwarp = gtid >> 5, gwarps = gsize >> 5;
    unsigned long long* ctr = w.counters;     // [10] long-row list length, [11] n0, [12] n1, [13] last viable + 1, [14] self-loop flag
    double* lsc = w.lscratch;                 // [0..gridDim) CTA totals, [4095] approximate sum of the weights

    if (n_large < 0) n_large = (i64)ctr[UGS_CTR_N_LARGE];  // fused small-batch path: the count stays on the device
    for (i64 item = 0; item < n_large; ++item) {
        const i64 g = w.llist[item];
        const i64 lo = ptr[g];
        const int n = (int)(ptr[g + 1] - lo);
        const i64 c0 = w.cstart[g];
        const i64 nc = w.cstart[g + 1] - c0;
        const i64 nb = lo - node_base;
        u32* rp = w.rowptr + nb + g;
        i32* ord = w.order + nb;
        i32* iof = w.index_of + nb;
        i32* sdg = w.suffix_deg + nb;
        double* bb = w.bucket_b + nb;
        double* cm = w.cum + nb;
        i32* cidx = w.csr_idx + 2 * c0;
        i32* ccol = w.csr_col + 2 * c0;

        // ---- A.2 degrees (iof = degree, ord = fill cursor) -------------------------------------------
        for (i64 v = gtid; v < n; v += gsize) { iof[v] = 0; ord[v] = 0; }
        if (gtid == 0) { ctr[10] = 0; ctr[11] = 0; ctr[12] = 0; ctr[13] = 0; ctr[14] = 0; lsc[4095] = 0.0; }
        grid.sync();
        for (i64 c = gtid; c < nc; c += gsize) {
            const i64 j = perm ? perm[c0 + c] : c0 + c;
            const i64 u = row0[j] - lo, v = row1[j] - lo;
            if (u < 0 || u >= n) { atomicOr(&ctr[UGS_CTR_UNSORTED], 8ull); continue; }
            if (v >= 0 && v < n) { atomicAdd(&iof[u], 1); atomicAdd(&iof[v], 1); if (u == v) ctr[14] = 1ull; }
        }
        grid.sync();
        int nnz;
        {
            unsigned long long* bs = reinterpret_cast<unsigned long long*>(lsc);
            unsigned long long* su = reinterpret_cast<unsigned long long*>(sscan);
            const unsigned long long tot = lg_grid_scan<unsigned long long>(
                grid, n, [&](int i) { return (unsigned long long)iof[i]; },
                [&](int i, unsigned long long ex) { rp[i] = (u32)ex; }, bs, su);
            if (gtid == 0) rp[n] = (u32)tot;  // read only after the next grid barrier
            nnz = (int)tot;
        }
        // ---- atomic fill, then sort every row by source column (= the reference's column order) --------
        for (i64 c = gtid; c < nc; c += gsize) {
            const i64 j = perm ? perm[c0 + c] : c0 + c;
            const i64 u = row0[j] - lo, v = row1[j] - lo;
            if (u >= 0 && u < n && v >= 0 && v < n) {
                const u32 pu = rp[u] + (u32)atomicAdd(&ord[u], 1);
                cidx[pu] = (i32)v; ccol[pu] = (i32)j;
                const u32 pv = rp[v] + (u32)atomicAdd(&ord[v], 1);
                cidx[pv] = (i32)u; ccol[pv] = (i32)j;
            }
        }
        grid.sync();
        for (i64 v = gwarp; v < n; v += gwarps) {
            const u32 p0 = rp[v];
            const int d = (int)(rp[v + 1] - p0);
            if (d <= 1) continue;
            if (d <= 32) {
                u64 key = (lane < d) ? lg_pack(ccol[p0 + lane], cidx[p0 + lane]) : ~0ull;
#pragma unroll
                for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
                    for (int j = k2 >> 1; j >= 1; j >>= 1) {
                        const u64 other = __shfl_xor_sync(FULL_MASK, key, j);
                        const bool take_min = (((lane & k2) == 0) == ((lane & j) == 0));
                        key = take_min ? (other < key ? other : key) : (other > key ? other : key);
                    }
                }
                if (lane < d) { ccol[p0 + lane] = (i32)(key >> 32); cidx[p0 + lane] = (i32)(u32)key; }
            } else if (d <= 128) {
                u64* wb = sbuf + warp * 128;
                for (int i = lane; i < 128; i += 32) wb[i] = (i < d) ? lg_pack(ccol[p0 + i], cidx[p0 + i]) : ~0ull;
                __syncwarp();
                for (int k2 = 2; k2 <= 128; k2 <<= 1) {
                    for (int j = k2 >> 1; j >= 1; j >>= 1) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int t = lane + 32 * q;
                            const int lo_ = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                            const int hi_ = lo_ | j;
                            const bool up = ((lo_ & k2) == 0);
                            const u64 a = wb[lo_], b = wb[hi_];
                            if ((a > b) == up) { wb[lo_] = b; wb[hi_] = a; }
                        }
                        __syncwarp();
                    }
                }
                for (int i = lane; i < d; i += 32) { ccol[p0 + i] = (i32)(wb[i] >> 32); cidx[p0 + i] = (i32)(u32)wb[i]; }
                __syncwarp();
            } else if (lane == 0) {
                sdg[atomicAdd(&ctr[10], 1ull)] = (i32)v;  // long rows: one CTA each, below
            }
        }
        grid.sync();
        {
            const int nlong = (int)ctr[10];
            for (int t = blockIdx.x; t < nlong; t += gridDim.x) {
                const int v = sdg[t];
                const u32 p0 = rp[v];
                const int d = (int)(rp[v + 1] - p0);
                if (d <= LG_SORT_SMEM) {
                    for (int i = tid; i < d; i += T_) sbuf[i] = lg_pack(ccol[p0 + i], cidx[p0 + i]);
                    __syncthreads();
                    lg_block_bitonic(sbuf, d);
                    for (int i = tid; i < d; i += T_) { ccol[p0 + i] = (i32)(sbuf[i] >> 32); cidx[p0 + i] = (i32)(u32)sbuf[i]; }
                    __syncthreads();
                } else {  // in-place network on the two arrays (columns of a row are distinct except self-loop twins)
                    for (int k2 = 2; (k2 >> 1) < d; k2 <<= 1) {
                        for (int j = -1; j != 0;) {
                            for (int i = tid; i < d; i += T_) {
                                const int p = (j < 0) ? (i ^ (k2 - 1)) : (i ^ j);
                                if (p > i && p < d) {
                                    const i32 ci = ccol[p0 + i], cp = ccol[p0 + p];
                                    if (cp < ci) {
                                        ccol[p0 + i] = cp; ccol[p0 + p] = ci;
                                        const i32 xi = cidx[p0 + i]; cidx[p0 + i] = cidx[p0 + p]; cidx[p0 + p] = xi;
                                    }
                                }
                            }
                            __syncthreads();
                            j = (j < 0) ? (k2 >> 2) : (j >> 1);
                        }
                    }
                }
            }
        }
        grid.sync();
        // ---- A.3 order: bitonic network over (degree << 32 | id). Sub-stages whose partner distance is below
        //      LG_SORT_SMEM stay inside one chunk of the keys and run in shared memory without a grid barrier; only
        //      the flips / strides that cross chunks touch global memory, one grid barrier each (n = 10^5: 28 barriers
        //      instead of 153).
        {
            u64* keys = reinterpret_cast<u64*>(cm);
            constexpr int C = LG_SORT_SMEM;
            const int nchunks = (n + C - 1) / C;
            for (i64 v = gtid; v < n; v += gsize) keys[v] = (((u64)(u32)iof[v]) << 32) | (u32)v;
            grid.sync();
            // chunk-local part of stage k2 (all of it when k2 <= C, else the strides below C), keys padded with +inf
            auto local_stages = [&](int k2_first, int k2_last) {
                for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
                    const int base = c * C;
                    for (int i = tid; i < C; i += T_) sbuf[i] = (base + i < n) ? keys[base + i] : ~0ull;
                    __syncthreads();
                    for (int k2 = k2_first; k2 <= k2_last; k2 <<= 1) {
                        int j = k2 >> 1;
                        if (k2 <= C) {  // flip inside the chunk
                            for (int t = tid; t < C / 2; t += T_) {
                                const int blk = t / j, off = t - blk * j;
                                const int a = blk * k2 + off, b = blk * k2 + k2 - 1 - off;
                                const u64 x = sbuf[a], y = sbuf[b];
                                if (y < x) { sbuf[a] = y; sbuf[b] = x; }
                            }
                            __syncthreads();
                            j >>= 1;
                        } else {
                            j = C >> 1;  // the flip and the strides >= C were done in global memory
                        }
                        for (; j >= 1; j >>= 1) {
                            for (int t = tid; t < C / 2; t += T_) {
                                const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1)), b = a | j;
                                const u64 x = sbuf[a], y = sbuf[b];
                                if (y < x) { sbuf[a] = y; sbuf[b] = x; }
                            }
                            __syncthreads();
                        }
                    }
                    for (int i = tid; i < C; i += T_) if (base + i < n) keys[base + i] = sbuf[i];
                    __syncthreads();
                }
            };
            local_stages(2, C);  // every stage up to k2 = C in one go
            grid.sync();
            for (int k2 = 2 * C; (k2 >> 1) < n; k2 <<= 1) {
                for (i64 i = gtid; i < n; i += gsize) {
                    const i64 p = i ^ (i64)(k2 - 1);
                    if (p > i && p < n) { const u64 x = keys[i], y = keys[p]; if (y < x) { keys[i] = y; keys[p] = x; } }
                }
                grid.sync();
                for (int j = k2 >> 2; j >= C; j >>= 1) {
                    for (i64 i = gtid; i < n; i += gsize) {
                        const i64 p = i ^ (i64)j;
                        if (p > i && p < n) { const u64 x = keys[i], y = keys[p]; if (y < x) { keys[i] = y; keys[p] = x; } }
                    }
                    grid.sync();
                }
                local_stages(k2, k2);
                grid.sync();
            }
            for (i64 vi = gtid; vi < n; vi += gsize) {
                const int v = (int)(u32)keys[vi];
                ord[vi] = v;
                iof[v] = (int)vi;
            }
        }
        grid.sync();
        // ---- A.4 suffix degree ---------------------------------------------------------------------------
        for (i64 vi = gwarp; vi < n; vi += gwarps) {
            const int v = ord[vi];
            const u32 p0 = rp[v], p1 = rp[v + 1];
            int cnt = 0;
            for (u32 p = p0 + lane; p < p1; p += 32) cnt += (iof[cidx[p]] >= vi) ? 1 : 0;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(FULL_MASK, cnt, off);
            if (lane == 0) sdg[vi] = cnt;
        }
        grid.sync();
        // ---- reachability + weight ----------------------------------------------------------------------
        {
            unsigned long long n0 = 0, n1 = 0;
            int lastv = -1;
            double wsum = 0.0;
            for (i64 vi = gwarp; vi < n; vi += gwarps) {
                const int v = ord[vi];
                const int sd = sdg[vi];
                bool reach = (k <= 1);
                if (!reach && sd > 0) {
                    int q = (lane == 0) ? v : -1;
                    int qn = 1;
                    for (int h = 0; h < qn && !reach; ++h) {
                        const int u = __shfl_sync(FULL_MASK, q, h);
                        const u32 p0 = rp[u], p1 = rp[u + 1];
                        for (u32 pb = p0; pb < p1 && !reach; pb += 32) {
                            const u32 p = pb + lane;
                            int x = -1;
                            bool ok = false;
                            if (p < p1) { x = cidx[p]; ok = (iof[x] >= vi); }
                            for (int j = 0; j < qn; ++j) {
                                const int qj = __shfl_sync(FULL_MASK, q, j);
                                ok = ok && (x != qj);
                            }
                            const u32 peers = __match_any_sync(FULL_MASK, ok ? x : (-1 - lane));
                            ok = ok && ((peers & lt_mask) == 0);
                            const u32 ball = __ballot_sync(FULL_MASK, ok);
                            const int cnt = __popc(ball);
                            if (qn + cnt >= k) { reach = true; break; }
                            const int t = lane - qn;
                            const int src = (t >= 0 && t < cnt) ? select_bit32(ball, t) : 0;
                            const int xs = __shfl_sync(FULL_MASK, x, src);
                            if (t >= 0 && t < cnt) q = xs;
                            qn += cnt;
                        }
                    }
                }
                if (lane == 0) {
                    double b = 0.0;
                    if (reach) {
                        const double d = (double)(sd > 1 ? sd : 1);
                        b = 1.0;
                        for (int t = 0; t < k - 1; ++t) b *= d;
                        n0++; lastv = max(lastv, (int)vi); wsum += b;
                    }
                    if (sd > 0) n1++;
                    bb[vi] = b;
                }
            }
            if (lane == 0) {
                if (n0) atomicAdd(&ctr[11], n0);
                if (n1) atomicAdd(&ctr[12], n1);
                if (lastv >= 0) atomicMax(&ctr[13], (unsigned long long)(lastv + 1));
                if (wsum > 0.0) atomicAdd(&lsc[4095], wsum);
            }
        }
        grid.sync();
        const int n0 = (int)ctr[11], n1 = (int)ctr[12], lastv = (int)ctr[13] - 1;
        const int level = (n0 > 0) ? 0 : ((n1 > 0) ? 1 : 2);
        const double approx = lsc[4095];
        grid.sync();  // everybody has read the scratch before the scans below reuse it
        // ---- root CDF -----------------------------------------------------------------------------------
        double Z = 0.0;
        if (level == 0) {
            if (approx <= 4503599627370496.0) {  // 2^52: every partial sum is an exactly representable integer
                Z = lg_grid_scan<double>(grid, n, [&](int i) { return bb[i]; },
                                         [&](int i, double ex) { cm[i] = ex + bb[i]; }, lsc, sscan);
            } else {
                if (gtid == 0) {  // the reference's left-to-right sum (preproc.cpp:234), one thread
                    double run = 0.0;
                    int i = 0;
                    for (; i + 8 <= n; i += 8) {
                        const double t0 = bb[i], t1 = bb[i + 1], t2 = bb[i + 2], t3 = bb[i + 3];
                        const double t4 = bb[i + 4], t5 = bb[i + 5], t6 = bb[i + 6], t7 = bb[i + 7];
                        run += t0; cm[i] = run; run += t1; cm[i + 1] = run; run += t2; cm[i + 2] = run;
                        run += t3; cm[i + 3] = run; run += t4; cm[i + 4] = run; run += t5; cm[i + 5] = run;
                        run += t6; cm[i + 6] = run; run += t7; cm[i + 7] = run;
                    }
                    for (; i < n; ++i) { run += bb[i]; cm[i] = run; }
                    lsc[4094] = run;
                }
                grid.sync();
                Z = lsc[4094];
            }
        } else if (level == 1) {
            lg_grid_scan<double>(grid, n, [&](int i) { return sdg[i] > 0 ? 1.0 : 0.0; },
                                 [&](int i, double ex) { cm[i] = ex + (sdg[i] > 0 ? 1.0 : 0.0); }, lsc, sscan);
        } else {
            for (i64 vi = gtid; vi < n; vi += gsize) cm[vi] = (double)(vi + 1);
        }
        if (gtid == 0) {
            ugs_ginfo gi;
            gi.Z = Z; gi.nnz = nnz; gi.level = level;
            gi.n_eligible = (level == 0) ? n0 : (level == 1 ? n1 : n);
            gi.last_viable = lastv; gi.cls = 2;
            gi.mult = ctr[14] ? 0 : -1;  // general path: -1 = no self-loop column
            w.ginfo[g] = gi;
        }
        grid.sync();
    }
    (void)wpb;
}

// Host launcher (called from ugs_preproc_batch)
int ugs_launch_preproc_large(const i64* row0, const i64* row1, const i64* ptr, i64 node_base, int k, const i64* perm,
                             i64 n_large, WS& w, cudaStream_t st) {
    int dev = 0, sms = 148, per_sm = 1, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return ugs_set_error(UGS_ECUDA, "device does not support cooperative launches");
    UGS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_preproc_large, LG_THREADS, 0));
    if (per_sm < 1) return ugs_set_error(UGS_ECUDA, "k_preproc_large does not fit on an SM");
    if (per_sm > 2) per_sm = 2;
    int blocks = sms * per_sm;
    if (blocks > 4000) blocks = 4000;  // CTA totals live in lscratch[0..4094)
    void* args[] = {(void*)&row0, (void*)&row1, (void*)&ptr, (void*)&node_base, (void*)&k, (void*)&perm,
                    (void*)&n_large, (void*)&w};
    UGS_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_preproc_large, dim3(blocks), dim3(LG_THREADS), args, 0, st));
    ++g_ugs_launches;
    return UGS_OK;
}
