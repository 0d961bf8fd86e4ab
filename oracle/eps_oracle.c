/*
 * oracle/eps_oracle.c -- TEST INFRASTRUCTURE ONLY (see ugs_oracle.c header for the rules on who may load it).
 *
 * Plain-C restatement of the reference's epsilon_uniform_sampler
 * (src/samplers/epsilon_uniform_sampler/src/epsilon_uniform_sampler.cpp), semantics in SURVEY.md Appendix A.9.
 *
 * Parity status: PINNED for the deterministic part (node sorting + edge extraction + output layout, checked
 * bit-exactly against the real reference's outputs for its own samples, tests/test_eps_oracle.py) and for the
 * LAW (exact enumerator oracle/exact_law.py::p_epsilon + chi-square against the real reference run with one
 * OpenMP thread). Sample VALUES are not the reference's: it draws from per-OpenMP-thread mt19937_64 streams with
 * std::uniform_*_distribution (implementation-defined, scheduling-dependent, :209-213); this oracle and the CUDA
 * path share one explicit rule instead:
 *   Philox4x32-10, key = seed, counter = (sample i, word block, graph index, 0x80000000 | attempt);
 *   words are consumed in order: start vertex; per growth iteration one word for the frontier pick and, when the
 *   picked vertex has candidates, one for the candidate pick (candidates = row entries not in S, in CSR row order,
 *   duplicates kept, :59-65); after a successful growth two words form the 53-bit acceptance uniform.
 *   All integer draws are mulhi32(word, range).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;
typedef int32_t i32;
typedef uint32_t u32;
typedef uint64_t u64;

/* from ugs_oracle.c */
void oracle_philox(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32* out);
typedef struct OPreprocOpaque OPreprocOpaque;
OPreprocOpaque* oracle_preproc_create(const i64* row0, const i64* row1, i64 ncols, i64 n, int k);
void oracle_preproc_free(OPreprocOpaque* P);
const i64* oracle_preproc_indptr(const OPreprocOpaque* P);
const i32* oracle_preproc_indices(const OPreprocOpaque* P);

typedef struct {
    u32 c0, c2, c3, k0, k1, block, idx;
    u32 w[4];
} WordStream;

static void ws_init(WordStream* s, u32 sample, u64 graph, u32 attempt, u64 seed) {
    s->c0 = sample; s->c2 = (u32)graph; s->c3 = 0x80000000u | attempt;
    s->k0 = (u32)seed; s->k1 = (u32)(seed >> 32); s->block = 0; s->idx = 4;
}
static u32 ws_next(WordStream* s) {
    if (s->idx == 4) { oracle_philox(s->c0, s->block++, s->c2, s->c3, s->k0, s->k1, s->w); s->idx = 0; }
    return s->w[s->idx++];
}
static inline u32 mulhi32(u32 a, u32 b) { return (u32)(((u64)a * (u64)b) >> 32); }

/* sample_connected_subgraph_rw -- epsilon_uniform_sampler.cpp:18-84. Returns 1 on success. */
static int grow_rw(const i64* indptr, const i32* indices, int n, int k, WordStream* rs, i32* S, double* weight) {
    if (n < k) return 0;
    int size = 0, fn = 0;
    i32* frontier = (i32*)malloc(sizeof(i32) * (size_t)(k + 1));
    double w = 1.0;
    i32 start = (i32)mulhi32(ws_next(rs), (u32)n);
    S[size++] = start;
    w *= (1.0 / n);
    frontier[fn++] = start;
    int attempts = 0;
    const int max_attempts = k * 100;
    while (size < k && attempts < max_attempts) {
        attempts++;
        if (fn == 0) { free(frontier); return 0; }
        int fi = (int)mulhi32(ws_next(rs), (u32)fn);
        i32 u = frontier[fi];
        i64 ncand = 0;
        for (i64 p = indptr[u]; p < indptr[u + 1]; ++p) {
            i32 v = indices[p];
            int in = 0; for (int a = 0; a < size; ++a) if (S[a] == v) { in = 1; break; }
            if (!in) ncand++;
        }
        if (ncand == 0) {
            for (int a = fi; a + 1 < fn; ++a) frontier[a] = frontier[a + 1]; /* erase, order kept (:69) */
            fn--;
            continue;
        }
        i64 ci = (i64)mulhi32(ws_next(rs), (u32)ncand);
        i32 chosen = -1;
        for (i64 p = indptr[u]; p < indptr[u + 1]; ++p) {
            i32 v = indices[p];
            int in = 0; for (int a = 0; a < size; ++a) if (S[a] == v) { in = 1; break; }
            if (in) continue;
            if (ci == 0) { chosen = v; break; }
            ci--;
        }
        S[size++] = chosen;
        frontier[fn++] = chosen;
        w *= (1.0 / (double)fn) * (1.0 / (double)ncand); /* :80, frontier size AFTER the push */
    }
    free(frontier);
    *weight = w;
    return size == k;
}

static int cmp_i32(const void* a, const void* b) { i32 x = *(const i32*)a, y = *(const i32*)b; return (x > y) - (x < y); }

/* One graph: m rows of k SORTED local ids, or all -1 when every attempt failed (:225-257).
 * trials_out (optional): total growth attempts. */
int eps_oracle_sample_nodes(const i64* indptr, const i32* indices, int n, int m, int k, u64 seed, u64 graph_index,
                            double epsilon, i32* nodes_out, i64* trials_out) {
    int max_att = (int)(10.0 / epsilon);
    if (max_att < 10) max_att = 10; /* :207 */
    i32* S = (i32*)malloc(sizeof(i32) * (size_t)(k + 1));
    i64 trials = 0;
    for (int i = 0; i < m; ++i) {
        int ok = 0;
        for (int a = 0; a < max_att && !ok; ++a) {
            WordStream rs; ws_init(&rs, (u32)i, graph_index, (u32)a, seed);
            double w;
            trials++;
            if (!grow_rw(indptr, indices, n, k, &rs, S, &w)) continue;
            double p = epsilon / (w + epsilon);
            if (p > 1.0) p = 1.0; /* :238 */
            u32 w0 = ws_next(&rs), w1 = ws_next(&rs);
            u64 u53 = ((((u64)w0) << 32) | (u64)w1) >> 11;
            double u = (double)u53 * (1.0 / 9007199254740992.0);
            if (u <= p) ok = 1; /* :240 */
        }
        if (ok) { qsort(S, (size_t)k, sizeof(i32), cmp_i32); for (int j = 0; j < k; ++j) nodes_out[(i64)i * k + j] = S[j]; } /* :256 */
        else for (int j = 0; j < k; ++j) nodes_out[(i64)i * k + j] = -1;
    }
    free(S);
    if (trials_out) *trials_out = trials;
    return 0;
}

/* Batch driver -- epsilon_uniform_sampler.cpp:122-362 with graphs appended in graph order (the reference's
 * order depends on OpenMP scheduling, :295-319). mode 0 = "sample" (positions in the sorted row), anything else
 * = batch-global node ids (:279-285). Edges: every batch column with both endpoints in the sample, in column
 * order, src = the column index (:267-290). given_nodes as in ugs_oracle.c. */
typedef struct {
    i64 G, m, k, total_edges;
    i64 *nodes, *edge_u, *edge_v, *edge_ptr, *sample_ptr, *edge_src;
} EBatch;

EBatch* eps_oracle_sample_batch(const i64* row0, const i64* row1, i64 E, const i64* ptr, i64 G, int m, int k, int mode,
                                u64 seed, u64 graph_index_base, double epsilon, const i64* given_nodes) {
    EBatch* B = (EBatch*)calloc(1, sizeof(EBatch));
    B->G = G; B->m = m; B->k = k;
    const i64 BT = G * (i64)m;
    B->nodes = (i64*)malloc(sizeof(i64) * (size_t)(BT * k + 1));
    for (i64 i = 0; i < BT * k; ++i) B->nodes[i] = -1;
    B->edge_ptr = (i64*)calloc((size_t)BT + 1, sizeof(i64));
    B->sample_ptr = (i64*)calloc((size_t)G + 1, sizeof(i64));
    i64 cap = 1024, ne = 0;
    B->edge_u = (i64*)malloc(sizeof(i64) * (size_t)cap);
    B->edge_v = (i64*)malloc(sizeof(i64) * (size_t)cap);
    B->edge_src = (i64*)malloc(sizeof(i64) * (size_t)cap);
    /* columns per graph (both endpoints inside the graph's node range, :166-168), column order kept */
    i64* gcount = (i64*)calloc((size_t)G + 2, sizeof(i64));
    i64* gid = (i64*)malloc(sizeof(i64) * (size_t)(E ? E : 1));
    for (i64 j = 0; j < E; ++j) {
        i64 g = -1;
        if (G > 0 && row0[j] >= ptr[0] && row0[j] < ptr[G]) {
            i64 lo = 0, hi = G - 1;
            while (lo < hi) { i64 mid = (lo + hi + 1) >> 1; if (ptr[mid] <= row0[j]) lo = mid; else hi = mid - 1; }
            g = lo;
            if (!(row1[j] >= ptr[g] && row1[j] < ptr[g + 1])) g = -1;
        }
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
    for (i64 g = 0; g < G; ++g) {
        const i64 lo = ptr[g], n = ptr[g + 1] - ptr[g];
        const i64 s0 = g * (i64)m;
        B->sample_ptr[g + 1] = B->sample_ptr[g] + m; /* :318, always m rows per graph */
        const i64 c0 = gcount[g], nc = gcount[g + 1] - gcount[g];
        i32* ln = (i32*)malloc(sizeof(i32) * (size_t)((i64)m * k + 1));
        if (given_nodes) {
            for (i64 t = 0; t < (i64)m * k; ++t) { i64 v = given_nodes[s0 * k + t]; ln[t] = v < 0 ? -1 : (i32)(v - lo); }
        } else if (n >= k && n > 0) {
            i64* l0 = (i64*)malloc(sizeof(i64) * (size_t)(nc ? nc : 1));
            i64* l1 = (i64*)malloc(sizeof(i64) * (size_t)(nc ? nc : 1));
            for (i64 c = 0; c < nc; ++c) { l0[c] = row0[gcols[c0 + c]] - lo; l1[c] = row1[gcols[c0 + c]] - lo; }
            OPreprocOpaque* P = oracle_preproc_create(l0, l1, nc, n, k); /* only its CSR (adjacency with duplicates, :170-174) is used */
            eps_oracle_sample_nodes(oracle_preproc_indptr(P), oracle_preproc_indices(P), (int)n, m, k, seed,
                                    graph_index_base + (u64)g, epsilon, ln, NULL);
            oracle_preproc_free(P);
            free(l0); free(l1);
        } else {
            for (i64 t = 0; t < (i64)m * k; ++t) ln[t] = -1; /* n < k: every attempt fails (:26) */
        }
        for (int s = 0; s < m; ++s) {
            const i32* row = ln + (i64)s * k;
            if (row[0] >= 0) {
                for (int j = 0; j < k; ++j) B->nodes[(s0 + s) * k + j] = lo + row[j];
                for (i64 c = 0; c < nc; ++c) {
                    i64 col = gcols[c0 + c];
                    i64 ul = row0[col] - lo, vl = row1[col] - lo;
                    int a = -1, b = -1;
                    for (int j = 0; j < k; ++j) { if (row[j] == ul) a = j; if (row[j] == vl) b = j; }
                    if (a < 0 || b < 0) continue;
                    if (ne + 1 > cap) {
                        cap *= 2;
                        B->edge_u = (i64*)realloc(B->edge_u, sizeof(i64) * (size_t)cap);
                        B->edge_v = (i64*)realloc(B->edge_v, sizeof(i64) * (size_t)cap);
                        B->edge_src = (i64*)realloc(B->edge_src, sizeof(i64) * (size_t)cap);
                    }
                    B->edge_u[ne] = mode == 0 ? a : lo + ul;
                    B->edge_v[ne] = mode == 0 ? b : lo + vl;
                    B->edge_src[ne] = col;
                    ne++;
                }
            }
            B->edge_ptr[s0 + s + 1] = ne;
        }
        free(ln);
    }
    B->total_edges = ne;
    free(gcount); free(gid); free(gcols);
    return B;
}

i64 eps_oracle_batch_total_edges(const EBatch* B) { return B->total_edges; }
void eps_oracle_batch_copy(const EBatch* B, i64* nodes, i64* edge_index, i64* edge_ptr, i64* sample_ptr, i64* edge_src) {
    memcpy(nodes, B->nodes, sizeof(i64) * (size_t)(B->G * B->m * B->k));
    memcpy(edge_index, B->edge_u, sizeof(i64) * (size_t)B->total_edges);
    memcpy(edge_index + B->total_edges, B->edge_v, sizeof(i64) * (size_t)B->total_edges);
    memcpy(edge_ptr, B->edge_ptr, sizeof(i64) * (size_t)(B->G * B->m + 1));
    memcpy(sample_ptr, B->sample_ptr, sizeof(i64) * (size_t)(B->G + 1));
    memcpy(edge_src, B->edge_src, sizeof(i64) * (size_t)B->total_edges);
}
void eps_oracle_batch_free(EBatch* B) {
    if (!B) return;
    free(B->nodes); free(B->edge_u); free(B->edge_v); free(B->edge_ptr); free(B->sample_ptr); free(B->edge_src); free(B);
}
