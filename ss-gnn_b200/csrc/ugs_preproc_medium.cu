// ugs_preproc_medium.cu -- preprocessing of MEDIUM general-path graphs (PROTEINS-sized: up to 1024 vertices and 3072
// edge columns) entirely in shared memory, one CTA of 1024 threads per graph.
//
// Same results as k_preproc_general (ugs_preproc.cu), i.e. the reference's build_csr / compute_1dd_ordering /
// compute_suffix_and_buckets (preproc.cpp:32-256). k_preproc_general keeps every table in the global
// workspace: each of its ~15 barrier-separated phases pays global-memory (L2) round trips, and with one CTA per graph
// nothing hides them -- 64 us for a 100-vertex graph, 140 us for a 620-vertex one, which is the critical path of a
// PROTEINS-shaped batch. Here the columns are read once, all phases work on shared memory (degrees, stable-by-column
// rows via an unordered fill + per-row sort, (degree, id) order by rank counting, suffix degree + bounded BFS, weights,
// root CDF), and the tables are written out once at the end.
// A graph that does not fit after all (a row longer than MED_ROW entries) is left untouched: its ugs_ginfo.level stays
// -1 and k_preproc_general, launched afterwards, picks it up (it skips every graph whose level is already >= 0).
#include "ugs_common.cuh"

#define MED_THREADS 1024
#define MED_N 1024     // vertices
#define MED_C 3072     // edge columns  (CSR entries <= 2 * MED_C)
#define MED_ROW 2048   // longest row the CTA-wide row sort handles
#define MED_TMP ((MED_THREADS / 32) * 128 > MED_ROW ? (MED_THREADS / 32) * 128 : MED_ROW)

struct MedSmem {
    double b[MED_N];
    double scan[72];
    u32 deg[MED_N];            // degree, then fill cursor, then degree again
    u32 rowptr[MED_N + 1];
    u32 sd[MED_N];
    u32 tmp[MED_TMP];          // per-warp 128-key sort slices, then the keys of a long row being sorted
    unsigned short cu[MED_C], cv[MED_C];
    unsigned short idx[2 * MED_C], col[2 * MED_C];
    unsigned short order[MED_N], index_of[MED_N];
    u32 vis[MED_THREADS / 32][32];   // bounded BFS: visited bitmap per warp (MED_N bits)
    int queue[MED_THREADS / 32][32]; // ... and its queue (fewer than k <= 32 entries)
    int cnt[8];                // [0] long rows, [1] n0, [2] n1, [3] last viable, [4] self loop, [5] give up
};

__device__ __forceinline__ bool ugs_is_medium(i64 n, i64 nc) { return n <= MED_N && nc <= MED_C; }

// block-wide exclusive scan of f(i), i < n (n <= MED_N): out(i, exclusive prefix); returns the total
template <typename T, typename F, typename O>
__device__ T med_scan(int n, F f, O out, T* sm /* >= 72 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = MED_THREADS >> 5;
    const int seg = (n + MED_THREADS - 1) / MED_THREADS;
    const int b = threadIdx.x * seg, e = min(n, b + seg);
    T local = 0;
    for (int i = b; i < e; ++i) local += f(i);
    T incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T y = __shfl_up_sync(FULL_MASK, incl, off);
        if (lane >= off) incl += y;
    }
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const T v = (lane < nw) ? sm[lane] : (T)0;
        T wi = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const T y = __shfl_up_sync(FULL_MASK, wi, off);
            if (lane >= off) wi += y;
        }
        if (lane < nw) sm[32 + lane] = wi - v;
        if (lane == 31) sm[64] = wi;
    }
    __syncthreads();
    T run = sm[32 + warp] + (incl - local);
    for (int i = b; i < e; ++i) { const T v = f(i); out(i, run); run += v; }
    const T total = sm[64];
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(MED_THREADS)
k_preproc_medium(const i64* __restrict__ row0, const i64* __restrict__ row1, const i64* __restrict__ ptr, i64 node_base,
                 int k, const i64* __restrict__ perm, i64 n_general, WS w) {
    extern __shared__ __align__(16) unsigned char med_raw[];
    MedSmem& s = *reinterpret_cast<MedSmem*>(med_raw);
    const int tid = threadIdx.x, T_ = MED_THREADS;
    const int warp = tid >> 5, lane = tid & 31, nwarps = T_ >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    if (n_general < 0) n_general = (i64)w.counters[UGS_CTR_N_GENERAL];  // fused small-batch path
    for (i64 item = blockIdx.x; item < n_general; item += gridDim.x) {
        const i64 g = w.glist[item];
        const i64 lo = ptr[g];
        const int n = (int)(ptr[g + 1] - lo);
        const i64 c0 = w.cstart[g];
        const int nc = (int)(w.cstart[g + 1] - c0);
        if (!ugs_is_medium(ptr[g + 1] - lo, w.cstart[g + 1] - c0)) continue;
        const i64 nb = lo - node_base;
        __syncthreads();  // the previous graph's write-out is done with the shared tables
        for (int v = tid; v < n; v += T_) s.deg[v] = 0;
        if (tid < 8) s.cnt[tid] = (tid == 3) ? -1 : 0;
        __syncthreads();
        // ---- A.2 columns -> shared memory, degrees ---------------------------------------------------------------------
        for (int c = tid; c < nc; c += T_) {
            const i64 j = perm ? perm[c0 + c] : c0 + c;
            const i64 u = row0[j] - lo, v = row1[j] - lo;
            unsigned short su = 0xFFFF, sv = 0xFFFF;
            if (u < 0 || u >= n) atomicOr(&w.counters[UGS_CTR_UNSORTED], 8ull);  // fast segmentation fooled
            else if (v >= 0 && v < n) {  // cross-graph / out-of-range columns are dropped (preproc.cpp:50,74)
                su = (unsigned short)u; sv = (unsigned short)v;
                atomicAdd(&s.deg[u], 1u);
                atomicAdd(&s.deg[v], 1u);
                if (u == v) s.cnt[4] = 1;
            }
            s.cu[c] = su; s.cv[c] = sv;
        }
        __syncthreads();
        const u32 nnz = med_scan<u32>(n, [&](int i) { return s.deg[i]; }, [&](int i, u32 ex) { s.rowptr[i] = ex; },
                                      reinterpret_cast<u32*>(s.scan));
        if (tid == 0) s.rowptr[n] = nnz;
        for (int v = tid; v < n; v += T_) {
            if (s.deg[v] > MED_ROW) s.cnt[5] = 1;
            s.deg[v] = 0;  // becomes the fill cursor (and equals the degree again after the fill)
        }
        __syncthreads();
        if (s.cnt[5]) continue;  // a row too long for the CTA-wide sort: k_preproc_general takes the graph
        // ---- unordered fill, then every row sorted by (column, neighbour) = the reference's column order ---------------
        for (int c = tid; c < nc; c += T_) {
            const u32 u = s.cu[c], v = s.cv[c];
            if (v == 0xFFFFu) continue;
            const u32 pu = s.rowptr[u] + atomicAdd(&s.deg[u], 1u);
            s.idx[pu] = (unsigned short)v; s.col[pu] = (unsigned short)c;
            const u32 pv = s.rowptr[v] + atomicAdd(&s.deg[v], 1u);
            s.idx[pv] = (unsigned short)u; s.col[pv] = (unsigned short)c;
        }
        __syncthreads();
        for (int v = warp; v < n; v += nwarps) {
            const u32 p0 = s.rowptr[v];
            const int d = (int)(s.rowptr[v + 1] - p0);
            if (d <= 1) continue;
            if (d <= 32) {  // one warp, bitonic network in registers on (column << 16 | neighbour)
                u32 key = (lane < d) ? (((u32)s.col[p0 + lane] << 16) | s.idx[p0 + lane]) : 0xFFFFFFFFu;
#pragma unroll
                for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
                    for (int j = k2 >> 1; j >= 1; j >>= 1) {
                        const u32 other = __shfl_xor_sync(FULL_MASK, key, j);
                        const bool take_min = (((lane & k2) == 0) == ((lane & j) == 0));
                        key = take_min ? min(key, other) : max(key, other);
                    }
                }
                if (lane < d) { s.col[p0 + lane] = (unsigned short)(key >> 16); s.idx[p0 + lane] = (unsigned short)key; }
            } else if (d <= 128) {  // this warp alone, in its 128-key slice of tmp[] (4 keys per lane)
                u32* wb = s.tmp + warp * 128;
                for (int i = lane; i < 128; i += 32)
                    wb[i] = (i < d) ? (((u32)s.col[p0 + i] << 16) | s.idx[p0 + i]) : 0xFFFFFFFFu;
                __syncwarp();
                for (int k2 = 2; k2 <= 128; k2 <<= 1) {
                    for (int j = k2 >> 1; j >= 1; j >>= 1) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int t = lane + 32 * q;  // 64 compare-exchanges per stage
                            const int lo_ = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                            const int hi_ = lo_ | j;
                            const bool up = ((lo_ & k2) == 0);
                            const u32 a = wb[lo_], b = wb[hi_];
                            if ((a > b) == up) { wb[lo_] = b; wb[hi_] = a; }
                        }
                        __syncwarp();
                    }
                }
                for (int i = lane; i < d; i += 32) { s.col[p0 + i] = (unsigned short)(wb[i] >> 16); s.idx[p0 + i] = (unsigned short)wb[i]; }
                __syncwarp();
            } else if (lane == 0) {
                s.sd[atomicAdd(&s.cnt[0], 1)] = (u32)v;  // queue of long rows (sd[] is not in use yet)
            }
        }
        __syncthreads();
        {
            const int nlong = s.cnt[0];
            for (int t = 0; t < nlong; ++t) {  // long rows (129 .. MED_ROW entries): CTA-wide bitonic network in tmp[]
                const u32 v = s.sd[t];
                const u32 p0 = s.rowptr[v];
                const int d = (int)(s.rowptr[v + 1] - p0);
                __syncthreads();
                for (int i = tid; i < d; i += T_) s.tmp[i] = ((u32)s.col[p0 + i] << 16) | s.idx[p0 + i];
                __syncthreads();
                for (int k2 = 2; (k2 >> 1) < d; k2 <<= 1) {
                    for (int i = tid; i < d; i += T_) {
                        const int p = i ^ (k2 - 1);
                        if (p > i && p < d) { const u32 x = s.tmp[i], y = s.tmp[p]; if (y < x) { s.tmp[i] = y; s.tmp[p] = x; } }
                    }
                    __syncthreads();
                    for (int j = k2 >> 2; j >= 1; j >>= 1) {
                        for (int i = tid; i < d; i += T_) {
                            const int p = i ^ j;
                            if (p > i && p < d) { const u32 x = s.tmp[i], y = s.tmp[p]; if (y < x) { s.tmp[i] = y; s.tmp[p] = x; } }
                        }
                        __syncthreads();
                    }
                }
                for (int i = tid; i < d; i += T_) { s.col[p0 + i] = (unsigned short)(s.tmp[i] >> 16); s.idx[p0 + i] = (unsigned short)s.tmp[i]; }
                __syncthreads();
            }
        }
        // deg[] = degree again (the cursors reached it)
        // ---- A.3 order = stable sort by (degree, id): rank by counting ---------------------------------------------------
        for (int v = tid; v < n; v += T_) {
            const u32 dv = s.deg[v];
            int rank = 0;
            for (int u = 0; u < v; ++u) rank += (s.deg[u] <= dv) ? 1 : 0;
            for (int u = v + 1; u < n; ++u) rank += (s.deg[u] < dv) ? 1 : 0;
            s.order[rank] = (unsigned short)v;
            s.index_of[v] = (unsigned short)rank;
        }
        __syncthreads();
        // ---- A.4 suffix degree, bounded BFS inside the suffix, weight: warp per position -------------------------------
        for (int vi = warp; vi < n; vi += nwarps) {
            const int v = s.order[vi];
            const u32 p0 = s.rowptr[v], p1 = s.rowptr[v + 1];
            int c = 0;
            for (u32 p = p0 + lane; p < p1; p += 32) c += ((int)s.index_of[s.idx[p]] >= vi) ? 1 : 0;
            c = __reduce_add_sync(FULL_MASK, c);
            const int sdv = c;
            bool reach = (k <= 1);
            if (!reach && sdv > 0) {
                // bounded BFS inside the suffix: visited set = this warp's 1024-bit map, queue = its 32-entry array (a
                // chunk of 32 row entries costs a load, an atomicOr and a vote; no shuffles over the queue, no match_any)
                u32* vis = s.vis[warp];
                int* qu = s.queue[warp];
                vis[lane] = 0u;
                __syncwarp();
                if (lane == 0) { qu[0] = v; vis[v >> 5] = 1u << (v & 31); }
                __syncwarp();
                int qn = 1;
                for (int h = 0; h < qn && !reach; ++h) {
                    const int u = qu[h];
                    const u32 q0 = s.rowptr[u], q1 = s.rowptr[u + 1];
                    for (u32 pb = q0; pb < q1 && !reach; pb += 32) {
                        const u32 p = pb + lane;
                        int x = -1;
                        bool isnew = false;
                        if (p < q1) {
                            x = s.idx[p];
                            if ((int)s.index_of[x] >= vi) {
                                const u32 bit = 1u << (x & 31);
                                isnew = !(atomicOr(&vis[x >> 5], bit) & bit);
                            }
                        }
                        const u32 ball = __ballot_sync(FULL_MASK, isnew);
                        const int cc = __popc(ball);
                        if (qn + cc >= k) { reach = true; break; }
                        if (isnew) qu[qn + __popc(ball & lt_mask)] = x;
                        qn += cc;
                        __syncwarp();
                    }
                }
                __syncwarp();
            }
            if (lane == 0) {
                double b = 0.0;
                if (reach) {
                    const double d = (double)(sdv > 1 ? sdv : 1);
                    b = 1.0;
                    for (int t = 0; t < k - 1; ++t) b *= d;  // preproc.cpp:228-233
                    atomicAdd(&s.cnt[1], 1);
                    atomicMax(&s.cnt[3], vi);
                }
                if (sdv > 0) atomicAdd(&s.cnt[2], 1);
                s.b[vi] = b;
                s.sd[vi] = (u32)sdv;
            }
        }
        __syncthreads();
        const int n0 = s.cnt[1], n1 = s.cnt[2], lastv = s.cnt[3];
        const int level = (n0 > 0) ? 0 : ((n1 > 0) ? 1 : 2);
        // ---- root CDF (straight into the workspace) -------------------------------------------------------------------------
        double* cm = w.cum + nb;
        double Z = 0.0;
        if (level == 0) {
            const double tot = med_scan<double>(n, [&](int i) { return s.b[i]; }, [&](int i, double ex) { cm[i] = ex + s.b[i]; },
                                                s.scan);
            Z = tot;
            if (!(tot <= 9007199254740992.0)) {  // partial sums may round: the reference's left-to-right sum (:234)
                if (tid == 0) {
                    double run = 0.0;
                    for (int i = 0; i < n; ++i) { run += s.b[i]; cm[i] = run; }
                    s.scan[0] = run;
                }
                __syncthreads();
                Z = s.scan[0];
                __syncthreads();
            }
        } else if (level == 1) {
            med_scan<double>(n, [&](int i) { return s.sd[i] > 0 ? 1.0 : 0.0; },
                             [&](int i, double ex) { cm[i] = ex + (s.sd[i] > 0 ? 1.0 : 0.0); }, s.scan);
        } else {
            for (int vi = tid; vi < n; vi += T_) cm[vi] = (double)(vi + 1);
        }
        // ---- write the tables ---------------------------------------------------------------------------------------------------
        {
            u32* rp = w.rowptr + nb + g;
            i32* cidx = w.csr_idx + 2 * c0;
            i32* ccol = w.csr_col + 2 * c0;
            for (int v = tid; v <= n; v += T_) rp[v] = s.rowptr[v];
            for (u32 p = tid; p < nnz; p += T_) {
                cidx[p] = (i32)s.idx[p];
                const i64 c = c0 + (i64)s.col[p];
                ccol[p] = (i32)(perm ? perm[c] : c);
            }
            for (int v = tid; v < n; v += T_) {
                w.order[nb + v] = (i32)s.order[v];
                w.index_of[nb + v] = (i32)s.index_of[v];
                w.suffix_deg[nb + v] = (i32)s.sd[v];
                w.bucket_b[nb + v] = s.b[v];
            }
            if (tid == 0) {
                ugs_ginfo gi;
                gi.Z = Z; gi.nnz = (i32)nnz; gi.level = level;
                gi.n_eligible = (level == 0) ? n0 : (level == 1 ? n1 : n);
                gi.last_viable = lastv; gi.cls = 2;
                gi.mult = s.cnt[4] ? 0 : -1;  // general path: -1 = no self-loop column
                w.ginfo[g] = gi;
            }
        }
    }
}

int ugs_launch_preproc_medium(const i64* row0, const i64* row1, const i64* ptr, i64 G, i64 node_base, int k,
                              const i64* perm, i64 n_general, WS& w, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = sizeof(MedSmem);
    UGS_CUDA_CHECK(cudaFuncSetAttribute(k_preproc_medium, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 bound = n_general < 0 ? G : n_general;
    i64 grid = bound < (i64)sms * 2 ? bound : (i64)sms * 2;
    if (grid < 1) grid = 1;
    k_preproc_medium<<<(unsigned)grid, MED_THREADS, smem, st>>>(row0, row1, ptr, node_base, k, perm, n_general, w);
    UGS_LAUNCH_CHECK("k_preproc_medium");
    return UGS_OK;
}
