// ugs_uniform_general.cu -- the exact-uniform accept/reject mode for general-path graphs
// (more than 64 vertices or 384 columns; at most UNIG_MAX_N vertices). Not in the reference; the rule is stated in
// oracle/uni_oracle.c ("the same mode for graphs of ANY size") and the output is bit-identical to that statement.
//   k_uniform_tables_general   CTA per graph: degree-dominating order by repeated removal of the vertex of maximum
//                              remaining SIMPLE degree (ties: smallest id), its tables (position, suffix degree, bounded
//                              reachability, weights, acceptance constant C, root CDF), overwriting the reference-order
//                              tables exactly like k_uniform_tables_small does for bitmask-path graphs.
//   k_sample_uniform_general   warp per sample: root draw, growth on the two-level cut bitmap of k_sample_general (the
//                              multigraph CSR gives the distinct cut for free), then p(S|r): every suffix neighbour of the
//                              sample gets a k-bit mask "adjacent to which members" (byte array in shared memory), the
//                              masks of the vertices outside S are histogrammed, a subset-sum transform turns the
//                              histogram into |cut(T)| for all 2^(k-1) member subsets T at once, and one lane runs the
//                              float32 push DP in the oracle's operation order. Accept with C / (b[root] * p).
#include "ugs_common.cuh"

#define UNIG_MAX_N 4096
#define UNIG_THREADS 1024
#define UNIG_MAX_TRIALS UGS_UNIFORM_MAX_TRIALS

__global__ void __launch_bounds__(UNIG_THREADS)
k_uniform_tables_general(const i64* __restrict__ ptr, i64 node_base, int k, i64 n_general, WS w) {
    __shared__ int s_drem[UNIG_MAX_N];
    __shared__ u32 s_gone[UNIG_MAX_N / 32];
    __shared__ u32 s_mark[32][UNIG_MAX_N / 32];  // per-warp dedupe bitmaps (phase 1); row 0 is the block's in phase 2
    __shared__ u32 s_red[32];
    __shared__ double s_q[32];
    __shared__ int s_cnt[4];
    const int tid = threadIdx.x, T_ = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = T_ >> 5;
    const u32 lt_mask = (1u << lane) - 1u;
    for (i64 item = blockIdx.x; item < n_general; item += gridDim.x) {
        const i64 g = w.glist[item];
        const i64 lo = ptr[g];
        const int n = (int)(ptr[g + 1] - lo);
        if (n > UNIG_MAX_N) continue;  // refused on the host
        const i64 nb = lo - node_base;
        const u32* __restrict__ rp = w.rowptr + nb + g;
        const i32* __restrict__ cidx = w.csr_idx + 2 * w.cstart[g];
        i32* ord = w.order + nb;
        i32* pos = w.index_of + nb;
        i32* dsf = w.suffix_deg + nb;
        double* bb = w.bucket_b + nb;
        double* cm = w.cum + nb;
        for (int i = tid; i < UNIG_MAX_N / 32; i += T_) s_gone[i] = 0;
        for (int i = tid; i < 32 * (UNIG_MAX_N / 32); i += T_) (&s_mark[0][0])[i] = 0;
        __syncthreads();
        // ---- simple degrees: distinct neighbours, no self loop (warp per row, per-warp dedupe bitmap) ----------------
        for (int v = warp; v < n; v += nwarps) {
            const u32 p0 = rp[v], p1 = rp[v + 1];
            int c = 0;
            for (u32 p = p0 + lane; p < p1; p += 32) {
                const int y = cidx[p];
                if (y == v) continue;
                const u32 bit = 1u << (y & 31);
                if (!(atomicOr(&s_mark[warp][y >> 5], bit) & bit)) c++;
            }
            c = __reduce_add_sync(FULL_MASK, c);
            __syncwarp();
            for (u32 p = p0 + lane; p < p1; p += 32) s_mark[warp][cidx[p] >> 5] = 0;
            __syncwarp();
            if (lane == 0) s_drem[v] = c;
        }
        __syncthreads();
        // ---- degree-dominating order: n rounds of (argmax over the remaining vertices, remove it, update its distinct
        //      remaining neighbours) ---------------------------------------------------------------------------------
        for (int vi = 0; vi < n; ++vi) {
            u32 key = 0;  // (degree + 1) << 12 | (4095 - v): maximum = highest degree, smallest id; 0 = none
            for (int v = tid; v < n; v += T_)
                if (!((s_gone[v >> 5] >> (v & 31)) & 1u)) key = max(key, ((u32)(s_drem[v] + 1) << 12) | (u32)(4095 - v));
            key = __reduce_max_sync(FULL_MASK, key);
            if (lane == 0) s_red[warp] = key;
            __syncthreads();
            key = s_red[lane < nwarps ? lane : 0];
            key = __reduce_max_sync(FULL_MASK, key);
            const int v = 4095 - (int)(key & 4095u), d = (int)(key >> 12) - 1;
            if (tid == 0) {
                ord[vi] = v; pos[v] = vi; dsf[vi] = d;
                s_gone[v >> 5] |= 1u << (v & 31);
            }
            __syncthreads();
            const u32 p0 = rp[v], p1 = rp[v + 1];
            for (u32 p = p0 + tid; p < p1; p += T_) {
                const int y = cidx[p];
                if (y == v || ((s_gone[y >> 5] >> (y & 31)) & 1u)) continue;
                const u32 bit = 1u << (y & 31);
                if (!(atomicOr(&s_mark[0][y >> 5], bit) & bit)) atomicSub(&s_drem[y], 1);
            }
            __syncthreads();
            for (u32 p = p0 + tid; p < p1; p += T_) s_mark[0][cidx[p] >> 5] = 0;
            // (the next round's first barrier orders these clears before the next use of s_mark[0])
        }
        __syncthreads();
        // ---- reachability inside the suffix, weight, acceptance bound: warp per position ---------------------------
        if (tid < 4) s_cnt[tid] = (tid == 1) ? -1 : 0;  // [0] viable count, [1] last viable
        double qmin = 1e300;
        __syncthreads();
        for (int vi = warp; vi < n; vi += nwarps) {
            const int v = ord[vi];
            bool reach = (k <= 1);
            if (!reach) {
                int q = (lane == 0) ? v : -1;  // lane j holds queue[j]
                int qn = 1;
                for (int h = 0; h < qn && !reach; ++h) {
                    const int u = __shfl_sync(FULL_MASK, q, h);
                    const u32 q0 = rp[u], q1 = rp[u + 1];
                    for (u32 pb = q0; pb < q1 && !reach; pb += 32) {
                        const u32 p = pb + lane;
                        int x = -1;
                        bool ok = false;
                        if (p < q1) { x = cidx[p]; ok = (pos[x] >= vi); }
                        for (int j = 0; j < qn; ++j) {
                            const int qj = __shfl_sync(FULL_MASK, q, j);
                            ok = ok && (x != qj);
                        }
                        const u32 peers = __match_any_sync(FULL_MASK, ok ? x : (-1 - lane));
                        ok = ok && ((peers & lt_mask) == 0);
                        const u32 ball = __ballot_sync(FULL_MASK, ok);
                        const int c = __popc(ball);
                        if (qn + c >= k) { reach = true; break; }
                        const int t = lane - qn;
                        const int src = (t >= 0 && t < c) ? select_bit32(ball, t) : 0;
                        const int xs = __shfl_sync(FULL_MASK, x, src);
                        if (t >= 0 && t < c) q = xs;
                        qn += c;
                    }
                }
            }
            if (lane == 0) {
                double b = 0.0;
                if (reach) {
                    const int d = dsf[vi] > 1 ? dsf[vi] : 1;
                    b = 1.0;
                    for (int t = 0; t < k - 1; ++t) b *= (double)d;
                    double den = 1.0;
                    const int nsuf = n - vi;
                    for (int i = 1; i <= k - 1; ++i) { const int a = i * d, c = nsuf - i; den *= (double)(a < c ? a : c); }
                    qmin = fmin(qmin, b / den);
                    atomicAdd(&s_cnt[0], 1);
                    atomicMax(&s_cnt[1], vi);
                }
                bb[vi] = b;
            }
        }
        if (lane == 0) s_q[warp] = qmin;
        __syncthreads();
        // ---- root CDF: the oracle's left-to-right running sum, one thread (n <= 4096) ---------------------------------
        if (tid == 0) {
            double run = 0.0;
            for (int vi = 0; vi < n; ++vi) { run += bb[vi]; cm[vi] = run; }
            double c = 1e300;
            for (int q = 0; q < nwarps; ++q) c = fmin(c, s_q[q]);
            ugs_ginfo gi = w.ginfo[g];
            gi.Z = run; gi.level = 0; gi.n_eligible = s_cnt[0]; gi.last_viable = s_cnt[1];
            w.ginfo[g] = gi;
            w.gaux[g] = (s_cnt[0] > 0) ? c : 0.0;
        }
        __syncthreads();
    }
}

struct UniStream {  // Philox words in order: counter (sample, block, graph, 0x40000000 | attempt) -- oracle us_init/us_next
    u32 c0, c2, c3, k0, k1, block, idx;
    Philox4 r;
    __device__ __forceinline__ void init(u32 sample, u64 graph, u32 attempt, u64 seed) {
        c0 = sample; c2 = (u32)graph; c3 = 0x40000000u | attempt;
        k0 = (u32)seed; k1 = (u32)(seed >> 32); block = 0; idx = 4;
    }
    __device__ __forceinline__ u32 next() {
        if (idx == 4) { r = philox4x32_10(c0, block++, c2, c3, k0, k1); idx = 0; }
        const u32 v = (idx == 0) ? r.x : (idx == 1) ? r.y : (idx == 2) ? r.z : r.w;
        idx++;
        return v;
    }
};

#define UNIG_WARPS 4
// per-warp shared memory: cut bitmap [128 words + 4 block counts], adjacency-mask bytes [4096], histogram [256],
// cut sizes / transition masks per subset [128], DP values [128]
struct __align__(16) UniGenWarp {
    u32 bm[UNIG_MAX_N / 32];
    u32 bc[4];
    u32 am[UNIG_MAX_N / 4];
    u32 hist[256];
    u32 ctab[128];  // |cut(T)| << 8 | members outside T that are adjacent to T
    float f[128];
};

__global__ void __launch_bounds__(UNIG_WARPS * 32)
k_sample_uniform_general(const i64* __restrict__ ptr, i64 node_base, i64 m, int k, u64 seed, u64 gbase, i64 n_general, WS w,
                         i64* __restrict__ nodes_out, i64* __restrict__ edge_ptr) {
    extern __shared__ __align__(16) unsigned char unig_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    UniGenWarp& s = reinterpret_cast<UniGenWarp*>(unig_raw)[warp];
    for (int i = lane; i < (int)(sizeof(UniGenWarp) / 4); i += 32) reinterpret_cast<u32*>(&s)[i] = 0;
    __syncwarp();
    const i64 total = n_general * m;
    const i64 nwarps = (i64)gridDim.x * wpb;
    const int full = (1 << (k - 1)) - 1, kmask = (1 << k) - 1;
    unsigned long long trials = 0, n_fail = 0, n_entries = 0;
    for (i64 t = (i64)blockIdx.x * wpb + warp; t < total; t += nwarps) {
        const i64 gi_ = t / m;
        const i64 i = t - gi_ * m;
        const i64 g = w.glist[gi_];
        const ugs_ginfo gi = w.ginfo[g];
        const i64 lo = ptr[g];
        const int n = (int)(ptr[g + 1] - lo);
        const i64 nb = lo - node_base;
        const u32* __restrict__ rp = w.rowptr + nb + g;
        const i32* __restrict__ ord = w.order + nb;
        const i32* __restrict__ iof = w.index_of + nb;
        const double* __restrict__ cm = w.cum + nb;
        const double* __restrict__ bb = w.bucket_b + nb;
        const i32* __restrict__ cidx = w.csr_idx + 2 * w.cstart[g];
        const double C = w.gaux[g];
        const u64 gidx = gbase + (u64)g;
        const int words = (n + 31) >> 5;
        const int nblk = (words + 31) >> 5;
        int mine = -1;
        u32 my_p0 = 0, my_p1 = 0;
        bool ok = false;
        if (gi.Z > 0.0 && n <= UNIG_MAX_N) {
            for (int a = 0; a < UNIG_MAX_TRIALS && !ok; ++a) {
                UniStream rs;
                rs.init((u32)i, gidx, (u32)a, seed);
                trials++;
                const u32 w0 = rs.next(), w1 = rs.next();
                const double x = root_x(w0, w1, gi.Z);
                int r_vi = warp_cdf_search(cm, n, x, lane);
                if (!(cm[r_vi] > x)) r_vi = gi.last_viable;
                const int root_vi = r_vi;
                // ---- growth: the r-th smallest vertex of the distinct suffix cut (two-level bitmap) ----------------------
                mine = -1;
                int size = 0, cutn = 0, xnew = ord[root_vi];
                for (int step = 0; step < k; ++step) {
                    if (step > 0) {
                        if (cutn == 0) break;
                        int rsel = (int)__umulhi(rs.next(), (u32)cutn);
                        int blk = 0;
                        if (nblk > 1) {
                            const int c = (lane < nblk) ? (int)s.bc[lane] : 0;  // nblk <= 4
                            const int incl = warp_incl_scan(c, lane);
                            const int src = __ffs(__ballot_sync(FULL_MASK, incl > rsel)) - 1;
                            rsel -= __shfl_sync(FULL_MASK, incl - c, src);
                            blk = src;
                        }
                        const u32 word = s.bm[blk * 32 + lane];
                        const int c = __popc(word);
                        const int incl = warp_incl_scan(c, lane);
                        const int src = __ffs(__ballot_sync(FULL_MASK, incl > rsel)) - 1;
                        const int excl_src = __shfl_sync(FULL_MASK, incl - c, src);
                        const u32 word_src = __shfl_sync(FULL_MASK, word, src);
                        xnew = (blk * 32 + src) * 32 + select_bit32(word_src, rsel - excl_src);
                        if (lane == 0) { s.bm[xnew >> 5] &= ~(1u << (xnew & 31)); s.bc[xnew >> 10] -= 1u; }
                        cutn--;
                        __syncwarp();
                    }
                    const u32 p0 = rp[xnew], p1 = rp[xnew + 1];
                    if (lane == size) { mine = xnew; my_p0 = p0; my_p1 = p1; }
                    size++;
                    if (step == k - 1) break;
                    int delta = 0;
                    for (u32 p = p0 + lane; p < p1; p += 32) {
                        const int y = cidx[p];
                        if (iof[y] >= root_vi) {
                            const u32 bit = 1u << (y & 31);
                            const u32 old = atomicOr(&s.bm[y >> 5], bit);
                            if (!(old & bit)) { atomicAdd(&s.bc[y >> 10], 1u); delta++; }
                        }
                    }
                    __syncwarp();
                    if (lane < size) {
                        const u32 bit = 1u << (mine & 31);
                        const u32 old = atomicAnd(&s.bm[mine >> 5], ~bit);
                        if (old & bit) { atomicSub(&s.bc[mine >> 10], 1u); delta--; }
                    }
                    cutn += __reduce_add_sync(FULL_MASK, delta);
                    __syncwarp();
                }
                for (int q = lane; q < nblk * 32; q += 32) s.bm[q] = 0u;  // what is left of the cut
                if (lane < 4) s.bc[lane] = 0u;
                __syncwarp();
                if (size < k) continue;
                // ---- adjacency masks: byte y of am[] = members (bit j = S[j]) that have y in their row, suffix only --------
                for (int j = 0; j < k; ++j) {
                    const int u = __shfl_sync(FULL_MASK, mine, j);
                    const u32 p0 = __shfl_sync(FULL_MASK, my_p0, j), p1 = __shfl_sync(FULL_MASK, my_p1, j);
                    for (u32 p = p0 + lane; p < p1; p += 32) {
                        const int y = cidx[p];
                        if (y != u && iof[y] >= root_vi) atomicOr(&s.am[y >> 2], (1u << j) << (8 * (y & 3)));
                    }
                }
                for (int q = lane; q < 256; q += 32) s.hist[q] = 0u;
                __syncwarp();
                // the members' own masks (which members they touch), then members are taken out of the table
                u32 adjM = 0;
                if (lane < k) adjM = (s.am[mine >> 2] >> (8 * (mine & 3))) & 0xFFu;
                __syncwarp();
                if (lane < k) atomicAnd(&s.am[mine >> 2], ~(0xFFu << (8 * (mine & 3))));
                __syncwarp();
                // every vertex outside S is claimed once (the byte is read and cleared atomically) and counted by its mask
                for (int j = 0; j < k; ++j) {
                    const u32 p0 = __shfl_sync(FULL_MASK, my_p0, j), p1 = __shfl_sync(FULL_MASK, my_p1, j);
                    for (u32 p = p0 + lane; p < p1; p += 32) {
                        const int y = cidx[p];
                        const int sh = 8 * (y & 3);
                        if ((s.am[y >> 2] >> sh) & 0xFFu) {
                            const u32 old = atomicAnd(&s.am[y >> 2], ~(0xFFu << sh));
                            const u32 mk = (old >> sh) & 0xFFu;
                            if (mk) atomicAdd(&s.hist[mk], 1u);
                        }
                    }
                }
                __syncwarp();
                // subset sums: hist[X] <- sum of hist[Y] over Y subset of X  (X < 2^k)
                for (int b = 0; b < k; ++b) {
                    u32 add[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int X = lane + 32 * q;
                        add[q] = (X <= kmask && ((X >> b) & 1)) ? s.hist[X ^ (1 << b)] : 0u;
                    }
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int X = lane + 32 * q;
                        if (add[q]) s.hist[X] += add[q];
                    }
                    __syncwarp();
                }
                const u32 total_ext = s.hist[kmask];
                u32 adjAll[8];  // every lane gets all members' masks (full-warp shuffles, before the partial-warp loop)
#pragma unroll
                for (int j = 0; j < 8; ++j) adjAll[j] = __shfl_sync(FULL_MASK, adjM, j);
                // |cut(T)| and the members T can grow into, for every T = {root} + subset q of the other members
                for (int q = lane; q < full; q += 32) {
                    const int tmask = 1 | (q << 1);
                    u32 c = total_ext - s.hist[kmask & ~tmask];  // outside vertices adjacent to at least one member of T
                    u32 tr = 0;
#pragma unroll
                    for (int j = 1; j < 8; ++j)
                        if (j < k && !((tmask >> j) & 1) && (adjAll[j] & (u32)tmask)) { c++; tr |= 1u << j; }
                    s.ctab[q] = (c << 8) | tr;
                    s.f[q] = 0.0f;
                }
                if (lane == 0) { s.f[full] = 0.0f; }
                __syncwarp();
                if (lane == 0) {  // float32 push DP: ascending subset index, ascending member index (oracle order)
                    s.f[0] = 1.0f;
                    for (int q = 0; q < full; ++q) {
                        const float fq = s.f[q];
                        if (fq == 0.0f) continue;
                        const u32 ct = s.ctab[q];
                        const u32 c = ct >> 8;
                        if (c == 0) continue;
                        const float qv = __fdiv_rn(fq, (float)c);
                        for (int j = 1; j < k; ++j)
                            if ((ct >> j) & 1u) s.f[q | (1 << (j - 1))] = __fadd_rn(s.f[q | (1 << (j - 1))], qv);
                    }
                }
                __syncwarp();
                const double den = __dmul_rn(bb[root_vi], (double)s.f[full]);
                const double pa = C / den;
                const u32 a0 = rs.next(), a1 = rs.next();
                const u64 v53 = ((((u64)a0) << 32) | (u64)a1) >> 11;
                const double uu = (double)v53 * (1.0 / 9007199254740992.0);
                ok = (uu < pa);
                __syncwarp();
            }
        }
        // ---- outputs: node row in growth order, induced-entry count (multigraph rows, S as the bitmap) -------------------
        int cntE = 0;
        if (ok) {
            if (lane < k) atomicOr(&s.bm[mine >> 5], 1u << (mine & 31));
            __syncwarp();
            for (int j = 0; j < k; ++j) {
                const u32 p0 = __shfl_sync(FULL_MASK, my_p0, j), p1 = __shfl_sync(FULL_MASK, my_p1, j);
                for (u32 p = p0 + lane; p < p1; p += 32) {
                    const int y = cidx[p];
                    cntE += (int)((s.bm[y >> 5] >> (y & 31)) & 1u);
                }
            }
            cntE = __reduce_add_sync(FULL_MASK, cntE);
            __syncwarp();
            if (lane < k) atomicAnd(&s.bm[mine >> 5], ~(1u << (mine & 31)));
            __syncwarp();
        }
        const i64 sidx = g * m + i;
        if (lane < k) nodes_out[sidx * k + lane] = ok ? lo + (i64)mine : (i64)-1;
        if (lane == 0) {
            edge_ptr[1 + sidx] = ok ? (i64)cntE : 0;
            if (!ok) n_fail++;
            else n_entries += (unsigned long long)cntE;
        }
    }
    if (lane == 0) {
        if (trials) atomicAdd(&w.counters[UGS_CTR_TRIALS], trials);
        if (n_fail) atomicAdd(&w.counters[UGS_CTR_INCOMPLETE], n_fail);
        (void)n_entries;
    }
}

static int unig_sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

int ugs_launch_uniform_tables_general(const i64* ptr, i64 node_base, int k, i64 n_general, i64 max_n_general, WS& w,
                                      cudaStream_t st) {
    if (n_general <= 0) return UGS_OK;
    if (max_n_general > UNIG_MAX_N)
        return ugs_set_error(UGS_ELIMIT, "the uniform variant supports graphs with at most %d nodes (largest: %lld)",
                             UNIG_MAX_N, (long long)max_n_general);
    const i64 sms = unig_sm_count();
    k_uniform_tables_general<<<(unsigned)(n_general < sms ? n_general : sms), UNIG_THREADS, 0, st>>>(ptr, node_base, k,
                                                                                                    n_general, w);
    UGS_LAUNCH_CHECK("k_uniform_tables_general");
    return UGS_OK;
}

int ugs_launch_sample_uniform_general(const i64* ptr, i64 node_base, i64 m, int k, u64 seed, u64 gbase, i64 n_general,
                                      WS& w, i64* nodes, i64* edge_ptr, cudaStream_t st) {
    if (n_general <= 0 || m <= 0) return UGS_OK;
    const size_t smem = sizeof(UniGenWarp) * UNIG_WARPS;
    UGS_CUDA_CHECK(cudaFuncSetAttribute(k_sample_uniform_general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    i64 blocks = (n_general * m + UNIG_WARPS - 1) / UNIG_WARPS;
    const i64 cap = (i64)unig_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_sample_uniform_general<<<(unsigned)blocks, UNIG_WARPS * 32, smem, st>>>(ptr, node_base, m, k, seed, gbase, n_general, w,
                                                                             nodes, edge_ptr);
    UGS_LAUNCH_CHECK("k_sample_uniform_general");
    return UGS_OK;
}
