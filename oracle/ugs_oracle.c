/*
 * oracle/ugs_oracle.c -- TEST INFRASTRUCTURE ONLY. Never linked into, imported by, or called from the
 * product path (ss-gnn_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this file's shared object.
 *
 * A plain-C, single-threaded CPU restatement of the reference UGS k-graphlet sampler
 * (AniruddhaMandal/SS-GNN, src/samplers/ugs_sampler), written from the semantics in SURVEY.md
 * Appendix A. Each function cites the reference file:line it follows.
 *
 * Parity status: PINNED. The deterministic stages (CSR, order, index_of, suffix_deg, bucket_b, Z,
 * alias table, induced-edge emission, batch driver layout) are checked bit-for-bit against the real
 * reference built in oracle/_ref (tests/test_oracle_vs_ref.py) and against committed golden vectors
 * generated from it (tests/golden/, tests/golden/make_golden.py).
 *
 * What is deliberately NOT the reference's: sample VALUES. The reference picks `cut[rng % |cut|]`
 * in libstdc++ unordered_set iteration order with an xorshift64* stream keyed on the sample index only
 * (sampler.cpp:55,79-80,160-161) -- implementation-defined and graph-independent. The law
 *  is what is portable. This oracle and the CUDA path share one explicit rule:
 *   - Philox4x32-10, key = (seed lo, seed hi), counter = (sample i, draw block, graph lo, graph hi);
 *   - root: level 0  x = u53 * 2^-53 * Z, root_vi = min{vi : cum[vi] > x}, cum = left-to-right
 *           running sum of bucket_b (so cum[n-1] == Z bit-exactly);
 *           level 1/2: idx = mulhi32(w0, count), the idx-th eligible position;
 *   - growth step t: cut = DISTINCT suffix neighbours of S not in S; choose the r-th SMALLEST vertex id,
 *           r = mulhi32(w, |cut|).
 * so the CUDA sampler can be compared with this oracle bit-exactly, and both are compared with the
 * reference in distribution (exact P_ref enumerator, oracle/exact_law.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;
typedef int32_t i32;
typedef uint32_t u32;
typedef uint64_t u64;

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al. 2011). Replaces the reference's ThreadRNG (sampler.hpp:26-36) and its
 * graph-independent seeding (sampler.cpp:160-161) -- see SURVEY.md section 0 fact 4.
 * ---------------------------------------------------------------------------------------------- */
static void philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32 out[4]) {
    for (int r = 0; r < 10; ++r) {
        u64 p0 = (u64)0xD2511F53u * c0;
        u64 p1 = (u64)0xCD9E8D57u * c2;
        u32 n0 = (u32)(p1 >> 32) ^ c1 ^ k0;
        u32 n1 = (u32)p1;
        u32 n2 = (u32)(p0 >> 32) ^ c3 ^ k1;
        u32 n3 = (u32)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_philox(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32* out) {
    philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}

static inline u32 mulhi32(u32 a, u32 b) { return (u32)(((u64)a * (u64)b) >> 32); }

/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    i64 n, nnz;
    int k;
    i64* indptr;      /* n+1 */
    i32* indices;     /* nnz */
    i32* edge_col;    /* nnz */
    i32* order;       /* n   */
    i32* index_of;    /* n   */
    i32* suffix_deg;  /* n (by position) */
    double* bucket_b; /* n (by position) */
    double* cum;      /* n: running sum used by the shared root-draw rule */
    double Z;
    int level;        /* relaxation level 0/1/2 (sampler.cpp:125-146) */
    i64 n_eligible;   /* #roots eligible at that level */
    i64 last_viable;  /* largest position with bucket_b > 0 */
    double* alias_prob;
    i32* alias_alias;
} OPreproc;

/* build_csr -- preproc.cpp:32-86. Multigraph CSR: every column appends v to row u THEN u to row v,
 * both tagged with the column index; columns with an endpoint outside [0,n) are skipped (:50,:74). */
static void build_csr(const i64* row0, const i64* row1, i64 ncols, OPreproc* P) {
    const i64 n = P->n;
    P->indptr = (i64*)calloc((size_t)n + 1, sizeof(i64));
    for (i64 j = 0; j < ncols; ++j) {
        i64 u = row0[j], v = row1[j];
        if (u < 0 || v < 0 || u >= n || v >= n) continue;
        P->indptr[u + 1]++; P->indptr[v + 1]++;
    }
    for (i64 i = 1; i <= n; ++i) P->indptr[i] += P->indptr[i - 1];
    P->nnz = P->indptr[n];
    P->indices = (i32*)malloc(sizeof(i32) * (size_t)(P->nnz ? P->nnz : 1));
    P->edge_col = (i32*)malloc(sizeof(i32) * (size_t)(P->nnz ? P->nnz : 1));
    i64* cur = (i64*)malloc(sizeof(i64) * (size_t)(n ? n : 1));
    for (i64 i = 0; i < n; ++i) cur[i] = P->indptr[i];
    for (i64 j = 0; j < ncols; ++j) {
        i64 u = row0[j], v = row1[j];
        if (u < 0 || v < 0 || u >= n || v >= n) continue;
        i64 pu = cur[u]++; P->indices[pu] = (i32)v; P->edge_col[pu] = (i32)j;
        i64 pv = cur[v]++; P->indices[pv] = (i32)u; P->edge_col[pv] = (i32)j;
    }
    free(cur);
}

/* compute_1dd_ordering -- preproc.cpp:97-166, restated step by step (bucket peeling with stale
 * entries, cur_max only descending, LIFO pops, final reverse). SURVEY.md section 0 fact 2 proves the result is
 * a stable sort by (multidegree, id); tests assert both. Buckets are singly linked stacks here. */
static void compute_order(OPreproc* P) {
    const i64 n = P->n;
    P->order = (i32*)malloc(sizeof(i32) * (size_t)(n ? n : 1));
    P->index_of = (i32*)malloc(sizeof(i32) * (size_t)(n ? n : 1));
    if (n == 0) return;
    i32* deg = (i32*)malloc(sizeof(i32) * (size_t)n);
    int max_deg = 0;
    for (i64 v = 0; v < n; ++v) {
        deg[v] = (i32)(P->indptr[v + 1] - P->indptr[v]);
        if (deg[v] > max_deg) max_deg = deg[v];
    }
    /* stack nodes: at most n initial + one per CSR entry pushed */
    i64 cap = n + P->nnz + 1, used = 0;
    i32* node_v = (i32*)malloc(sizeof(i32) * (size_t)cap);
    i64* node_next = (i64*)malloc(sizeof(i64) * (size_t)cap);
    i64* head = (i64*)malloc(sizeof(i64) * (size_t)(max_deg + 1));
    for (int d = 0; d <= max_deg; ++d) head[d] = -1;
    for (i64 v = 0; v < n; ++v) { /* push_back in id order => top of stack is the largest id */
        node_v[used] = (i32)v; node_next[used] = head[deg[v]]; head[deg[v]] = used++;
    }
    char* removed = (char*)calloc((size_t)n, 1);
    i64 out = 0;
    int cur_max = max_deg;
    while (out < n) {
        while (cur_max >= 0 && head[cur_max] < 0) cur_max--;
        if (cur_max < 0) break; /* cannot happen: every vertex keeps its original entry */
        i64 t = head[cur_max]; head[cur_max] = node_next[t];
        i32 v = node_v[t];
        if (removed[v]) continue;
        removed[v] = 1;
        P->order[out++] = v;
        for (i64 p = P->indptr[v]; p < P->indptr[v + 1]; ++p) {
            i32 u = P->indices[p];
            if (removed[u]) continue;
            int od = deg[u];
            deg[u] = od - 1;
            if (od - 1 >= 0) { node_v[used] = u; node_next[used] = head[od - 1]; head[od - 1] = used++; }
        }
    }
    for (i64 a = 0, b = n - 1; a < b; ++a, --b) { i32 t = P->order[a]; P->order[a] = P->order[b]; P->order[b] = t; }
    for (i64 i = 0; i < n; ++i) P->index_of[P->order[i]] = (i32)i;
    free(deg); free(node_v); free(node_next); free(head); free(removed);
}

/* compute_suffix_and_buckets -- preproc.cpp:176-243. suffix_deg counts duplicate and self entries;
 * reach = bounded FIFO BFS in the suffix graph stopping at k queued vertices (:207-224);
 * b = max(1,sd)^(k-1) by k-1 sequential double multiplies (:228-233); Z accumulated in vi order. */
static void compute_weights(OPreproc* P, int k) {
    const i64 n = P->n;
    P->suffix_deg = (i32*)calloc((size_t)(n ? n : 1), sizeof(i32));
    P->bucket_b = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
    P->cum = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
    P->Z = 0.0;
    for (i64 vi = 0; vi < n; ++vi) {
        i32 v = P->order[vi];
        i32 c = 0;
        for (i64 p = P->indptr[v]; p < P->indptr[v + 1]; ++p)
            if (P->index_of[P->indices[p]] >= vi) c++;
        P->suffix_deg[vi] = c;
    }
    i32* queue = (i32*)malloc(sizeof(i32) * (size_t)(k > 0 ? k + 1 : 1));
    i64* stamp = (i64*)malloc(sizeof(i64) * (size_t)(n ? n : 1)); /* visited[] without the O(n) re-allocation */
    for (i64 i = 0; i < n; ++i) stamp[i] = -1;
    for (i64 vi = 0; vi < n; ++vi) {
        i32 v = P->order[vi];
        int qn = 0;
        queue[qn++] = v; stamp[v] = vi;
        for (int head = 0; head < qn && qn < k; ++head) {
            i32 u = queue[head];
            for (i64 p = P->indptr[u]; p < P->indptr[u + 1]; ++p) {
                i32 w = P->indices[p];
                if (P->index_of[w] < vi) continue;
                if (stamp[w] == vi) continue;
                stamp[w] = vi; queue[qn++] = w;
                if (qn >= k) break;
            }
        }
        if (qn >= k) {
            int d = P->suffix_deg[vi] > 1 ? P->suffix_deg[vi] : 1;
            double b = 1.0;
            for (int t = 0; t < k - 1; ++t) b *= (double)d;
            P->bucket_b[vi] = b;
            P->Z += b;
        }
        P->cum[vi] = P->Z;
    }
    free(queue); free(stamp);
}

/* AliasTable::build -- sampler.hpp:44-69 (Vose, LIFO small/large stacks, left-to-right sum). Built
 * only when Z > 0 (preproc.cpp:241-243). Not used by the shared root-draw rule; kept so the oracle can be
 * pinned against the reference's alias table. */
static void build_alias(OPreproc* P) {
    const i64 n = P->n;
    P->alias_prob = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
    P->alias_alias = (i32*)calloc((size_t)(n ? n : 1), sizeof(i32));
    if (!(P->Z > 0.0) || n == 0) return;
    double sum = 0.0;
    for (i64 i = 0; i < n; ++i) sum += P->bucket_b[i];
    double* p = (double*)malloc(sizeof(double) * (size_t)n);
    i32* small = (i32*)malloc(sizeof(i32) * (size_t)n); i64 ns = 0;
    i32* large = (i32*)malloc(sizeof(i32) * (size_t)n); i64 nl = 0;
    for (i64 i = 0; i < n; ++i) p[i] = P->bucket_b[i] * (double)(int)n / (sum > 0 ? sum : 1.0);
    for (i64 i = 0; i < n; ++i) { if (p[i] < 1.0) small[ns++] = (i32)i; else large[nl++] = (i32)i; }
    while (ns > 0 && nl > 0) {
        i32 s = small[--ns];
        i32 l = large[nl - 1];
        P->alias_prob[s] = p[s];
        P->alias_alias[s] = l;
        p[l] = (p[l] + p[s]) - 1.0;
        if (p[l] < 1.0) { small[ns++] = l; nl--; }
    }
    for (i64 i = 0; i < nl; ++i) P->alias_prob[large[i]] = 1.0;
    for (i64 i = 0; i < ns; ++i) P->alias_prob[small[i]] = 1.0;
    free(p); free(small); free(large);
}

/* relaxation levels -- sampler.cpp:117-146. For levels 1/2 the root CDF `cum` is re-expressed as the running
 * count of eligible positions (level 1: suffix_deg > 0; level 2: every position), so that one CDF search
 * serves every level (shared root-draw rule, see header). */
static void compute_level(OPreproc* P) {
    i64 c0 = 0, c1 = 0;
    for (i64 vi = 0; vi < P->n; ++vi) { if (P->bucket_b[vi] > 0.0) c0++; if (P->suffix_deg[vi] > 0) c1++; }
    P->last_viable = -1;
    for (i64 vi = 0; vi < P->n; ++vi) if (P->bucket_b[vi] > 0.0) P->last_viable = vi;
    if (c0 > 0) { P->level = 0; P->n_eligible = c0; }
    else if (c1 > 0) {
        P->level = 1; P->n_eligible = c1;
        i64 run = 0;
        for (i64 vi = 0; vi < P->n; ++vi) { if (P->suffix_deg[vi] > 0) run++; P->cum[vi] = (double)run; }
    } else {
        P->level = 2; P->n_eligible = P->n;
        for (i64 vi = 0; vi < P->n; ++vi) P->cum[vi] = (double)(vi + 1);
    }
}

/* create_preproc -- preproc.cpp:262-284 */
OPreproc* oracle_preproc_create(const i64* row0, const i64* row1, i64 ncols, i64 n, int k) {
    OPreproc* P = (OPreproc*)calloc(1, sizeof(OPreproc));
    P->n = n; P->k = k;
    build_csr(row0, row1, ncols, P);
    compute_order(P);
    compute_weights(P, k);
    build_alias(P);
    compute_level(P);
    return P;
}

void oracle_preproc_free(OPreproc* P) {
    if (!P) return;
    free(P->indptr); free(P->indices); free(P->edge_col); free(P->order); free(P->index_of);
    free(P->suffix_deg); free(P->bucket_b); free(P->cum); free(P->alias_prob); free(P->alias_alias);
    free(P);
}

i64 oracle_preproc_n(const OPreproc* P) { return P->n; }
i64 oracle_preproc_nnz(const OPreproc* P) { return P->nnz; }
double oracle_preproc_Z(const OPreproc* P) { return P->Z; }
int oracle_preproc_level(const OPreproc* P) { return P->level; }
i64 oracle_preproc_bucket_count_nonzero(const OPreproc* P) { /* preproc.cpp:311-312 */
    i64 c = 0; for (i64 i = 0; i < P->n; ++i) if (P->bucket_b[i] > 0.0) c++; return c;
}
void oracle_preproc_copy(const OPreproc* P, i64* indptr, i32* indices, i32* edge_col, i32* order,
                         i32* index_of, i32* suffix_deg, double* bucket_b, double* cum,
                         double* alias_prob, i32* alias_alias) {
    size_t n = (size_t)P->n, nnz = (size_t)P->nnz;
    memcpy(indptr, P->indptr, (n + 1) * sizeof(i64));
    memcpy(indices, P->indices, nnz * sizeof(i32));
    memcpy(edge_col, P->edge_col, nnz * sizeof(i32));
    memcpy(order, P->order, n * sizeof(i32));
    memcpy(index_of, P->index_of, n * sizeof(i32));
    memcpy(suffix_deg, P->suffix_deg, n * sizeof(i32));
    memcpy(bucket_b, P->bucket_b, n * sizeof(double));
    memcpy(cum, P->cum, n * sizeof(double));
    memcpy(alias_prob, P->alias_prob, n * sizeof(double));
    memcpy(alias_alias, P->alias_alias, n * sizeof(i32));
}

/* ------------------------------------------------------------------------------------------------
 * Root draw (law A.5: sampler.cpp:163-173 + sampler.hpp:72-77) under the shared rule.
 * ---------------------------------------------------------------------------------------------- */
static i64 draw_root(const OPreproc* P, const u32 w[4]) {
    const i64 n = P->n;
    double x;
    if (P->level == 0) {
        u64 u53 = ((((u64)w[0]) << 32) | (u64)w[1]) >> 11;
        x = ((double)u53 * (1.0 / 9007199254740992.0)) * P->Z;
    } else {
        x = (double)mulhi32(w[0], (u32)P->n_eligible); /* idx-th eligible position, uniform */
    }
    /* smallest vi with cum[vi] > x (binary search; cum is non-decreasing) */
    i64 lo = 0, hi = n - 1;
    while (lo < hi) { i64 mid = (lo + hi) >> 1; if (P->cum[mid] > x) hi = mid; else lo = mid + 1; }
    /* guard: at level 0, x can round up to Z; fall back to the last position with b > 0 */
    if (P->level == 0 && !(P->cum[lo] > x)) lo = P->last_viable;
    return lo;
}

static int cmp_i32(const void* a, const void* b) { i32 x = *(const i32*)a, y = *(const i32*)b; return (x > y) - (x < y); }

/* rand_grow -- sampler.cpp:36-85 (law A.6) under the shared rule. Returns #vertices written (<= k). */
static int grow(const OPreproc* P, int k, i64 root_vi, u64 seed, u64 graph_index, u32 sample_i,
                u32 w[4], i32* verts, i32* scratch /* >= sum of row lengths of k vertices */) {
    int size = 0;
    verts[size++] = P->order[root_vi];
    u32 block = 0; int widx = 2; /* w[0],w[1] consumed by the root draw */
    for (int step = 1; step < k; ++step) {
        i64 nc = 0;
        for (int a = 0; a < size; ++a) {
            i32 u = verts[a];
            for (i64 p = P->indptr[u]; p < P->indptr[u + 1]; ++p) {
                i32 x = P->indices[p];
                if (P->index_of[x] < root_vi) continue;
                int inS = 0; for (int b = 0; b < size; ++b) if (verts[b] == x) { inS = 1; break; }
                if (inS) continue;
                scratch[nc++] = x;
            }
        }
        if (nc == 0) return size;
        qsort(scratch, (size_t)nc, sizeof(i32), cmp_i32);
        i64 nd = 0;
        for (i64 a = 0; a < nc; ++a) if (a == 0 || scratch[a] != scratch[a - 1]) scratch[nd++] = scratch[a];
        if (widx == 4) {
            block++;
            philox4x32_10(sample_i, block, (u32)graph_index, (u32)(graph_index >> 32), (u32)seed, (u32)(seed >> 32), w);
            widx = 0;
        }
        u32 r = mulhi32(w[widx++], (u32)nd);
        verts[size++] = scratch[r];
    }
    return size;
}

/* sample -- sampler.cpp:91-181 (node rows only). nodes_out: m*k local vertex ids, -1 padded. */
static u32 g_sample_index_base = 0; /* Philox counter word 0 = g_sample_index_base + i (sample sharding) */
void oracle_set_sample_index_base(u64 b) { g_sample_index_base = (u32)b; }

int oracle_sample_nodes(const OPreproc* P, int m, int k, u64 seed, u64 graph_index, i32* nodes_out) {
    if (P->n <= 0) return -1; /* "No viable roots available" sampler.cpp:148-150 */
    i64 maxrow = 0;
    for (i64 v = 0; v < P->n; ++v) { i64 d = P->indptr[v + 1] - P->indptr[v]; if (d > maxrow) maxrow = d; }
    i32* scratch = (i32*)malloc(sizeof(i32) * (size_t)(maxrow * (k > 0 ? k : 1) + 1));
    i32* verts = (i32*)malloc(sizeof(i32) * (size_t)(k > 0 ? k : 1));
    for (int i = 0; i < m; ++i) {
        u32 w[4];
        const u32 ictr = (u32)i + g_sample_index_base;
        philox4x32_10(ictr, 0u, (u32)graph_index, (u32)(graph_index >> 32), (u32)seed, (u32)(seed >> 32), w);
        i64 root_vi = draw_root(P, w);
        int size = grow(P, k, root_vi, seed, graph_index, ictr, w, verts, scratch);
        for (int j = 0; j < k; ++j) nodes_out[(i64)i * k + j] = j < size ? verts[j] : -1;
    }
    free(scratch); free(verts);
    return 0;
}

/* Induced-edge emission -- sampler.cpp:205-246 (law A.7). For a COMPLETE row (no -1): for j in growth
 * order, for p over row(S[j]) in CSR order, emit (j, local(indices[p]), edge_col[p]) when indices[p] in S.
 * Returns the number of entries; writes them if the out pointers are non-NULL. */
i64 oracle_extract_one(const OPreproc* P, int k, const i32* verts, i32* out_j, i32* out_l, i32* out_col) {
    for (int j = 0; j < k; ++j) if (verts[j] < 0) return 0; /* incomplete: sampler.cpp:220-223 */
    i64 cnt = 0;
    for (int j = 0; j < k; ++j) {
        i32 u = verts[j];
        for (i64 p = P->indptr[u]; p < P->indptr[u + 1]; ++p) {
            i32 x = P->indices[p];
            int l = -1;
            /* unordered_map built by ascending j with overwrite => the LAST j wins for duplicate ids
             * (sampler.cpp:226-229); samples never contain duplicates, so first == last here. */
            for (int b = 0; b < k; ++b) if (verts[b] == x) l = b;
            if (l < 0) continue;
            if (out_j) { out_j[cnt] = j; out_l[cnt] = l; out_col[cnt] = P->edge_col[p]; }
            cnt++;
        }
    }
    return cnt;
}

/* ------------------------------------------------------------------------------------------------
 * Batch driver -- ugs_sampler_batch_extension.cpp:77-300 (laws A.1, A.7 modes, A.12 post-processing).
 * The reference rescans all E columns per graph (:41-75, O(G*E)); the result is the per-graph column list
 * in original column order, produced here by one bucketing pass (same lists, same order).
 * mode: 0 = "sample" (local j,l), 1 = "graph" (i*k+j, i = sample index inside its graph), 2 = "global".
 * given_nodes: if non-NULL, [G*m*k] node rows (global ids, -1 pad) to extract edges for instead of sampling
 *              (used to pin the extraction stage against the reference's own samples).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    i64 G, m, k, total_edges;
    i64* nodes;       /* G*m*k */
    i64* edge_u;      /* total_edges */
    i64* edge_v;
    i64* edge_ptr;    /* G*m+1 */
    i64* sample_ptr;  /* G+1 */
    i64* edge_src;    /* total_edges */
} OBatch;

static i64 find_graph(const i64* ptr, i64 G, i64 u) {
    /* graph g with ptr[g] <= u < ptr[g+1]; assumes ptr non-decreasing. -1 if none. */
    if (G <= 0 || u < ptr[0] || u >= ptr[G]) return -1;
    i64 lo = 0, hi = G - 1;
    while (lo < hi) { i64 mid = (lo + hi + 1) >> 1; if (ptr[mid] <= u) lo = mid; else hi = mid - 1; }
    /* skip empty graphs sharing the same start: pick the one whose range contains u */
    while (lo < G && !(ptr[lo] <= u && u < ptr[lo + 1])) lo++;
    return lo < G ? lo : -1;
}

OBatch* oracle_sample_batch(const i64* row0, const i64* row1, i64 E, const i64* ptr, i64 G, int m, int k,
                            int mode, u64 seed, u64 graph_index_base, const i64* given_nodes) {
    OBatch* B = (OBatch*)calloc(1, sizeof(OBatch));
    B->G = G; B->m = m; B->k = k;
    const i64 BT = G * (i64)m;
    B->nodes = (i64*)malloc(sizeof(i64) * (size_t)(BT * k + 1));
    for (i64 i = 0; i < BT * k; ++i) B->nodes[i] = -1;
    B->edge_ptr = (i64*)calloc((size_t)BT + 1, sizeof(i64));
    B->sample_ptr = (i64*)calloc((size_t)G + 1, sizeof(i64));
    /* A.1: bucket columns per graph, preserving column order */
    i64* gcount = (i64*)calloc((size_t)G + 1, sizeof(i64));
    i64* gid = (i64*)malloc(sizeof(i64) * (size_t)(E ? E : 1));
    for (i64 j = 0; j < E; ++j) {
        i64 g = find_graph(ptr, G, row0[j]);
        if (g >= 0 && !(row1[j] >= ptr[g] && row1[j] < ptr[g + 1])) g = -1;
        gid[j] = g;
        if (g >= 0) gcount[g + 1]++;
    }
    for (i64 g = 0; g < G; ++g) gcount[g + 1] += gcount[g];
    i64* gcols = (i64*)malloc(sizeof(i64) * (size_t)(E ? E : 1));
    {
        i64* cur = (i64*)malloc(sizeof(i64) * (size_t)(G ? G : 1));
        for (i64 g = 0; g < G; ++g) cur[g] = gcount[g];
        for (i64 j = 0; j < E; ++j) if (gid[j] >= 0) gcols[cur[gid[j]]++] = j;
        free(cur);
    }
    i64 cap = 1024, ne = 0;
    B->edge_u = (i64*)malloc(sizeof(i64) * (size_t)cap);
    B->edge_v = (i64*)malloc(sizeof(i64) * (size_t)cap);
    B->edge_src = (i64*)malloc(sizeof(i64) * (size_t)cap);
    i32* verts = (i32*)malloc(sizeof(i32) * (size_t)(k > 0 ? k : 1));
    for (i64 g = 0; g < G; ++g) {
        const i64 lo = ptr[g], hi = ptr[g + 1], n = hi - lo;
        B->sample_ptr[g + 1] = B->sample_ptr[g] + m; /* :130 unconditional */
        const i64 s0 = g * (i64)m;
        if (n <= 0 || n < k) { /* :132-143 */
            for (int s = 0; s < m; ++s) B->edge_ptr[s0 + s + 1] = ne;
            continue;
        }
        const i64 c0 = gcount[g], nc = gcount[g + 1] - gcount[g];
        i64* l0 = (i64*)malloc(sizeof(i64) * (size_t)(nc ? nc : 1));
        i64* l1 = (i64*)malloc(sizeof(i64) * (size_t)(nc ? nc : 1));
        for (i64 c = 0; c < nc; ++c) { l0[c] = row0[gcols[c0 + c]] - lo; l1[c] = row1[gcols[c0 + c]] - lo; }
        OPreproc* P = oracle_preproc_create(l0, l1, nc, n, k);
        i32* ln = (i32*)malloc(sizeof(i32) * (size_t)((i64)m * k + 1));
        if (given_nodes) {
            for (i64 t = 0; t < (i64)m * k; ++t) { i64 v = given_nodes[s0 * k + t]; ln[t] = v < 0 ? -1 : (i32)(v - lo); }
        } else {
            oracle_sample_nodes(P, m, k, seed, graph_index_base + (u64)g, ln);
        }
        for (int s = 0; s < m; ++s) {
            const i32* row = ln + (i64)s * k;
            for (int j = 0; j < k; ++j) B->nodes[(s0 + s) * k + j] = row[j] < 0 ? -1 : (i64)row[j] + lo; /* :188-196 */
            i64 cnt = oracle_extract_one(P, k, row, NULL, NULL, NULL);
            if (ne + cnt > cap) {
                while (ne + cnt > cap) cap *= 2;
                B->edge_u = (i64*)realloc(B->edge_u, sizeof(i64) * (size_t)cap);
                B->edge_v = (i64*)realloc(B->edge_v, sizeof(i64) * (size_t)cap);
                B->edge_src = (i64*)realloc(B->edge_src, sizeof(i64) * (size_t)cap);
            }
            if (cnt > 0) {
                i32* ej = (i32*)malloc(sizeof(i32) * (size_t)cnt * 3);
                i32* el = ej + cnt; i32* ec = el + cnt;
                oracle_extract_one(P, k, row, ej, el, ec);
                for (i64 t = 0; t < cnt; ++t) {
                    i64 a, b;
                    if (mode == 0) { a = ej[t]; b = el[t]; }                                     /* sampler.cpp:262-264 */
                    else if (mode == 1) { a = (i64)s * k + ej[t]; b = (i64)s * k + el[t]; }     /* :265-267 */
                    else { a = lo + row[ej[t]]; b = lo + row[el[t]]; }                          /* :268-273 */
                    B->edge_u[ne + t] = a; B->edge_v[ne + t] = b;
                    B->edge_src[ne + t] = gcols[c0 + ec[t]];                                    /* batch ext :213-224 */
                }
                free(ej);
            }
            ne += cnt;
            B->edge_ptr[s0 + s + 1] = ne; /* :199-202 */
        }
        free(ln); free(l0); free(l1);
        oracle_preproc_free(P);
    }
    B->total_edges = ne;
    free(verts); free(gcount); free(gid); free(gcols);
    return B;
}

i64 oracle_batch_total_edges(const OBatch* B) { return B->total_edges; }
void oracle_batch_copy(const OBatch* B, i64* nodes, i64* edge_index /* [2,Ebar] */, i64* edge_ptr, i64* sample_ptr, i64* edge_src) {
    memcpy(nodes, B->nodes, sizeof(i64) * (size_t)(B->G * B->m * B->k));
    memcpy(edge_index, B->edge_u, sizeof(i64) * (size_t)B->total_edges);
    memcpy(edge_index + B->total_edges, B->edge_v, sizeof(i64) * (size_t)B->total_edges);
    memcpy(edge_ptr, B->edge_ptr, sizeof(i64) * (size_t)(B->G * B->m + 1));
    memcpy(sample_ptr, B->sample_ptr, sizeof(i64) * (size_t)(B->G + 1));
    memcpy(edge_src, B->edge_src, sizeof(i64) * (size_t)B->total_edges);
}
void oracle_batch_free(OBatch* B) {
    if (!B) return;
    free(B->nodes); free(B->edge_u); free(B->edge_v); free(B->edge_ptr); free(B->sample_ptr); free(B->edge_src); free(B);
}

/* raw CSR accessors (used by eps_oracle.c, which shares the adjacency-with-duplicates rows) */
const i64* oracle_preproc_indptr(const OPreproc* P) { return P->indptr; }
const i32* oracle_preproc_indices(const OPreproc* P) { return P->indices; }
