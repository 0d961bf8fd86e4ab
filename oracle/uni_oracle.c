/*
 * oracle/uni_oracle.c -- TEST INFRASTRUCTURE ONLY (see ugs_oracle.c header).
 *
 * CPU statement of the exact-uniform accept/reject mode (Bressan, "Efficient and
 * near-optimal algorithms for sampling small connected subgraphs", UGS). This mode does NOT exist in the
 * reference (its `sample()` has no rejection step, sampler.cpp:158-181); it is the explicitly
 * named variant that makes a chi-square against the UNIFORM law pass. Parity: "unpinned" against the reference by
 * construction; pinned against the exact uniform law (oracle/exact_law.py::connected_subsets) and bit-exactly
 * against the CUDA path, which follows the same rule:
 *   simple adjacency A[v] (distinct neighbours, no self loops), graphs of <= 64 vertices, k <= 8;
 *   degree-dominating order: repeatedly remove the vertex of maximum remaining degree (ties: smallest id);
 *   position vi: suffix = vertices not yet removed, d = its degree inside the suffix;
 *   b[vi] = max(1,d)^(k-1) if the component of v in the suffix graph has >= k vertices else 0;
 *   C = min over viable vi of b[vi] / prod_{i=1..k-1} min(i*max(1,d), n_suffix - i)   (so C <= b*p(S|r) for all S);
 *   trial a: Philox(sample, block, graph, 0x40000000|a) words in order: 2 root words (53-bit CDF inversion),
 *   one word per growth step (r-th smallest vertex of the distinct cut), 2 acceptance words;
 *   p(S|r) by a push DP over subsets in float32 (ascending subset index, ascending member index);
 *   accept iff u53*2^-53 < C / (b[root] * p); at most UNI_MAX_TRIALS trials, then a row of -1.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;
typedef int32_t i32;
typedef uint32_t u32;
typedef uint64_t u64;

#define UNI_MAX_TRIALS 4096

void oracle_philox(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32* out);

typedef struct {
    int n, k;
    u64 A[64];
    int order[64], dsuf[64];
    u64 suf[64];
    double b[64], cum[64], Z, C;
    int last_viable;
} UniTables;

static inline int popc64(u64 x) { return __builtin_popcountll(x); }
static inline u32 mulhi32(u32 a, u32 b) { return (u32)(((u64)a * (u64)b) >> 32); }

int uni_oracle_tables(const i64* indptr, const i32* indices, int n, int k, UniTables* T) {
    if (n > 64 || k > 8 || k < 1) return -1;
    memset(T, 0, sizeof(*T));
    T->n = n; T->k = k;
    for (int v = 0; v < n; ++v)
        for (i64 p = indptr[v]; p < indptr[v + 1]; ++p) if (indices[p] != v) T->A[v] |= 1ull << indices[p];
    u64 rem = n == 64 ? ~0ull : ((1ull << n) - 1ull);
    for (int vi = 0; vi < n; ++vi) {
        int best = -1, bd = -1;
        for (int v = 0; v < n; ++v) if ((rem >> v) & 1ull) { int d = popc64(T->A[v] & rem); if (d > bd) { bd = d; best = v; } }
        T->order[vi] = best; T->dsuf[vi] = bd; T->suf[vi] = rem;
        rem &= ~(1ull << best);
    }
    T->Z = 0.0; T->C = 0.0; T->last_viable = -1;
    int have = 0;
    for (int vi = 0; vi < n; ++vi) {
        u64 sufm = T->suf[vi], vis = 1ull << T->order[vi], expd = 0;
        int reach = 0;
        for (int it = 0; it < k; ++it) {
            if (popc64(vis) >= k) { reach = 1; break; }
            u64 fr = vis & ~expd;
            if (!fr) break;
            int u = __builtin_ctzll(fr);
            expd |= 1ull << u;
            vis |= T->A[u] & sufm;
        }
        double b = 0.0;
        if (reach) {
            int d = T->dsuf[vi] > 1 ? T->dsuf[vi] : 1;
            b = 1.0;
            for (int t = 0; t < k - 1; ++t) b *= (double)d;
            double den = 1.0;
            int nsuf = n - vi;
            for (int i = 1; i <= k - 1; ++i) { int a = i * d, c = nsuf - i; den *= (double)(a < c ? a : c); }
            double q = b / den;
            if (!have || q < T->C) { T->C = q; have = 1; }
            T->last_viable = vi;
        }
        T->b[vi] = b;
        T->Z += b;
        T->cum[vi] = T->Z;
    }
    return 0;
}

typedef struct { u32 c0, c2, c3, k0, k1, block, idx; u32 w[4]; } UStream;
static void us_init(UStream* s, u32 sample, u64 graph, u32 attempt, u64 seed) {
    s->c0 = sample; s->c2 = (u32)graph; s->c3 = 0x40000000u | attempt;
    s->k0 = (u32)seed; s->k1 = (u32)(seed >> 32); s->block = 0; s->idx = 4;
}
static u32 us_next(UStream* s) {
    if (s->idx == 4) { oracle_philox(s->c0, s->block++, s->c2, s->c3, s->k0, s->k1, s->w); s->idx = 0; }
    return s->w[s->idx++];
}

static int select_bit(u64 m, int r) { for (int i = 0; i < r; ++i) m &= m - 1; return __builtin_ctzll(m); }

/* p(S|r) -- A.8: sum over growth orders of prod 1/|cut|, float32 push DP */
static float p_given_root(const UniTables* T, const int* mem, int k, u64 sufm) {
    float f[128];
    const int full = (1 << (k - 1)) - 1;
    for (int t = 0; t <= full; ++t) f[t] = 0.0f;
    f[0] = 1.0f;
    for (int t = 0; t < full; ++t) {
        if (f[t] == 0.0f) continue;
        u64 nbr = T->A[mem[0]], vm = 1ull << mem[0];
        for (int j = 1; j < k; ++j) if ((t >> (j - 1)) & 1) { nbr |= T->A[mem[j]]; vm |= 1ull << mem[j]; }
        u64 cut = nbr & sufm & ~vm;
        int c = popc64(cut);
        if (c == 0) continue;
        float q = f[t] / (float)c;
        for (int j = 1; j < k; ++j)
            if (!((t >> (j - 1)) & 1) && ((cut >> mem[j]) & 1ull)) f[t | (1 << (j - 1))] = f[t | (1 << (j - 1))] + q;
    }
    return f[full];
}

/* m rows of k local ids in growth order, or -1 rows. trials_out: total trials. */
int uni_oracle_sample_nodes(const UniTables* T, int m, u64 seed, u64 graph_index, i32* nodes_out, i64* trials_out) {
    const int n = T->n, k = T->k;
    i64 trials = 0;
    for (int i = 0; i < m; ++i) {
        int ok = 0, mem[8];
        if (T->Z > 0.0) {
            for (int a = 0; a < UNI_MAX_TRIALS && !ok; ++a) {
                UStream rs; us_init(&rs, (u32)i, graph_index, (u32)a, seed);
                trials++;
                u32 w0 = us_next(&rs), w1 = us_next(&rs);
                u64 u53 = ((((u64)w0) << 32) | (u64)w1) >> 11;
                double x = ((double)u53 * (1.0 / 9007199254740992.0)) * T->Z;
                int lo = 0, hi = n - 1;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (T->cum[mid] > x) hi = mid; else lo = mid + 1; }
                if (!(T->cum[lo] > x)) lo = T->last_viable;
                const int root_vi = lo;
                const u64 sufm = T->suf[root_vi];
                mem[0] = T->order[root_vi];
                u64 S = 1ull << mem[0], nbr = T->A[mem[0]];
                int size = 1;
                for (int step = 1; step < k; ++step) {
                    u64 cut = nbr & sufm & ~S;
                    int c = popc64(cut);
                    if (c == 0) break;
                    int x2 = select_bit(cut, (int)mulhi32(us_next(&rs), (u32)c));
                    S |= 1ull << x2; nbr |= T->A[x2]; mem[size++] = x2;
                }
                if (size < k) continue;
                float p = p_given_root(T, mem, k, sufm);
                double den = T->b[root_vi] * (double)p;
                double pa = T->C / den;
                u32 a0 = us_next(&rs), a1 = us_next(&rs);
                u64 v53 = ((((u64)a0) << 32) | (u64)a1) >> 11;
                double uu = (double)v53 * (1.0 / 9007199254740992.0);
                if (uu < pa) ok = 1;
            }
        }
        for (int j = 0; j < k; ++j) nodes_out[(i64)i * k + j] = ok ? mem[j] : -1;
    }
    if (trials_out) *trials_out = trials;
    return 0;
}

/* convenience for ctypes */
UniTables* uni_oracle_create(const i64* indptr, const i32* indices, int n, int k) {
    UniTables* T = (UniTables*)malloc(sizeof(UniTables));
    if (uni_oracle_tables(indptr, indices, n, k, T) != 0) { free(T); return NULL; }
    return T;
}
void uni_oracle_free(UniTables* T) { free(T); }
void uni_oracle_dump(const UniTables* T, i32* order, i32* dsuf, u64* suf, double* b, double* cum, double* ZC) {
    for (int i = 0; i < T->n; ++i) { order[i] = T->order[i]; dsuf[i] = T->dsuf[i]; suf[i] = T->suf[i]; b[i] = T->b[i]; cum[i] = T->cum[i]; }
    ZC[0] = T->Z; ZC[1] = T->C;
}

/* ---------------------------------------------------------------------------------------------------------------------
 * The same mode for graphs of ANY size (general path of the CUDA implementation: graphs above 64 vertices). Same rule,
 * stated on adjacency lists instead of 64-bit masks; for n <= 64 it reproduces the mask version above bit for bit
 * (tests/test_eps_oracle.py::test_uniform_general_oracle_equals_mask_oracle):
 *   simple adjacency = distinct neighbours without self loops; degree-dominating order by repeated removal of the vertex
 *   of maximum remaining degree (ties: smallest id); pos[] = position in that order; suffix of vi = {w : pos[w] >= vi};
 *   cut of a vertex set T inside the suffix of the root = distinct w adjacent to T with pos[w] >= root_vi, w not in T;
 *   growth picks the r-th smallest vertex id of the cut; p(S|r) by the same float32 push DP, |cut| per subset counted on
 *   the lists.
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct {
    int n, k;
    i64* ap;      /* simple adjacency CSR */
    i32* ai;
    i32 *order, *pos, *dsuf;
    double *b, *cum, Z, C;
    int last_viable;
} UniGen;

static int cmp_i32(const void* a, const void* b) { i32 x = *(const i32*)a, y = *(const i32*)b; return (x > y) - (x < y); }

void uni_gen_free(UniGen* T) {
    if (!T) return;
    free(T->ap); free(T->ai); free(T->order); free(T->pos); free(T->dsuf); free(T->b); free(T->cum); free(T);
}

UniGen* uni_gen_create(const i64* indptr, const i32* indices, int n, int k) {
    if (k < 1 || k > 8 || n < 1) return NULL;
    UniGen* T = (UniGen*)calloc(1, sizeof(UniGen));
    T->n = n; T->k = k;
    T->ap = (i64*)calloc((size_t)n + 1, sizeof(i64));
    T->ai = (i32*)malloc(sizeof(i32) * (size_t)(indptr[n] > 0 ? indptr[n] : 1));
    i64 w = 0;
    for (int v = 0; v < n; ++v) {  /* sorted distinct neighbours, no self loop */
        T->ap[v] = w;
        i64 s = w;
        for (i64 p = indptr[v]; p < indptr[v + 1]; ++p) if (indices[p] != v) T->ai[w++] = indices[p];
        qsort(T->ai + s, (size_t)(w - s), sizeof(i32), cmp_i32);
        i64 o = s;
        for (i64 p = s; p < w; ++p) if (p == s || T->ai[p] != T->ai[p - 1]) T->ai[o++] = T->ai[p];
        w = o;
    }
    T->ap[n] = w;
    T->order = (i32*)malloc(sizeof(i32) * n); T->pos = (i32*)malloc(sizeof(i32) * n); T->dsuf = (i32*)malloc(sizeof(i32) * n);
    T->b = (double*)calloc(n, sizeof(double)); T->cum = (double*)calloc(n, sizeof(double));
    i32* drem = (i32*)malloc(sizeof(i32) * n);
    char* gone = (char*)calloc(n, 1);
    for (int v = 0; v < n; ++v) drem[v] = (i32)(T->ap[v + 1] - T->ap[v]);
    for (int vi = 0; vi < n; ++vi) {
        int best = -1, bd = -1;
        for (int v = 0; v < n; ++v) if (!gone[v] && drem[v] > bd) { bd = drem[v]; best = v; }
        T->order[vi] = best; T->pos[best] = vi; T->dsuf[vi] = bd;
        gone[best] = 1;
        for (i64 p = T->ap[best]; p < T->ap[best + 1]; ++p) if (!gone[T->ai[p]]) drem[T->ai[p]]--;
    }
    free(drem); free(gone);
    T->Z = 0.0; T->C = 0.0; T->last_viable = -1;
    int have = 0;
    i32* q = (i32*)malloc(sizeof(i32) * (size_t)(k + 1));
    for (int vi = 0; vi < n; ++vi) {
        /* bounded BFS inside the suffix: does the component of order[vi] reach k vertices? (any visiting order) */
        int qn = 1, reach = (k <= 1);
        q[0] = T->order[vi];
        for (int h = 0; h < qn && !reach; ++h) {
            const int u = q[h];
            for (i64 p = T->ap[u]; p < T->ap[u + 1] && !reach; ++p) {
                const int x = T->ai[p];
                if (T->pos[x] < vi) continue;
                int seen = 0;
                for (int j = 0; j < qn; ++j) if (q[j] == x) { seen = 1; break; }
                if (seen) continue;
                q[qn++] = x;
                if (qn >= k) reach = 1;
            }
        }
        double b = 0.0;
        if (reach) {
            int d = T->dsuf[vi] > 1 ? T->dsuf[vi] : 1;
            b = 1.0;
            for (int t = 0; t < k - 1; ++t) b *= (double)d;
            double den = 1.0;
            int nsuf = n - vi;
            for (int i = 1; i <= k - 1; ++i) { int a = i * d, c = nsuf - i; den *= (double)(a < c ? a : c); }
            double qq = b / den;
            if (!have || qq < T->C) { T->C = qq; have = 1; }
            T->last_viable = vi;
        }
        T->b[vi] = b;
        T->Z += b;
        T->cum[vi] = T->Z;
    }
    free(q);
    return T;
}

void uni_gen_dump(const UniGen* T, i32* order, i32* dsuf, double* b, double* cum, double* ZC) {
    for (int i = 0; i < T->n; ++i) { order[i] = T->order[i]; dsuf[i] = T->dsuf[i]; b[i] = T->b[i]; cum[i] = T->cum[i]; }
    ZC[0] = T->Z; ZC[1] = T->C;
}

/* distinct cut of the members flagged in `tmask` (bit j = mem[j]) inside the suffix of root_vi, ascending ids into
 * out[] (capacity cap); returns its size */
static int gen_cut(const UniGen* T, const int* mem, int k, int tmask, int root_vi, i32* out, int cap) {
    int c = 0;
    for (int j = 0; j < k; ++j) {
        if (!((tmask >> j) & 1)) continue;
        for (i64 p = T->ap[mem[j]]; p < T->ap[mem[j] + 1]; ++p) {
            const int x = T->ai[p];
            if (T->pos[x] < root_vi) continue;
            int skip = 0;
            for (int l = 0; l < k; ++l) if (((tmask >> l) & 1) && mem[l] == x) { skip = 1; break; }
            if (skip) continue;
            if (c < cap) out[c] = x;
            c++;
        }
    }
    if (c > cap) return -1;
    qsort(out, (size_t)c, sizeof(i32), cmp_i32);
    int o = 0;
    for (int i = 0; i < c; ++i) if (i == 0 || out[i] != out[i - 1]) out[o++] = out[i];
    return o;
}

int uni_gen_sample_nodes(const UniGen* T, int m, u64 seed, u64 graph_index, i32* nodes_out, i64* trials_out) {
    const int n = T->n, k = T->k;
    i64 trials = 0;
    i64 maxdeg = 0;
    for (int v = 0; v < n; ++v) if (T->ap[v + 1] - T->ap[v] > maxdeg) maxdeg = T->ap[v + 1] - T->ap[v];
    const int cap = (int)(maxdeg * k + 8);
    i32* cut = (i32*)malloc(sizeof(i32) * (size_t)cap);
    for (int i = 0; i < m; ++i) {
        int ok = 0, mem[8];
        if (T->Z > 0.0) {
            for (int a = 0; a < UNI_MAX_TRIALS && !ok; ++a) {
                UStream rs; us_init(&rs, (u32)i, graph_index, (u32)a, seed);
                trials++;
                u32 w0 = us_next(&rs), w1 = us_next(&rs);
                u64 u53 = ((((u64)w0) << 32) | (u64)w1) >> 11;
                double x = ((double)u53 * (1.0 / 9007199254740992.0)) * T->Z;
                int lo = 0, hi = n - 1;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (T->cum[mid] > x) hi = mid; else lo = mid + 1; }
                if (!(T->cum[lo] > x)) lo = T->last_viable;
                const int root_vi = lo;
                mem[0] = T->order[root_vi];
                int size = 1;
                for (int step = 1; step < k; ++step) {
                    const int c = gen_cut(T, mem, size, (1 << size) - 1, root_vi, cut, cap);
                    if (c <= 0) break;
                    mem[size++] = cut[mulhi32(us_next(&rs), (u32)c)];
                }
                if (size < k) continue;
                /* p(S|r): float32 push DP, ascending subset index, ascending member index */
                float f[128];
                const int full = (1 << (k - 1)) - 1;
                for (int t = 0; t <= full; ++t) f[t] = 0.0f;
                f[0] = 1.0f;
                for (int t = 0; t < full; ++t) {
                    if (f[t] == 0.0f) continue;
                    const int tmask = 1 | (t << 1);
                    const int c = gen_cut(T, mem, k, tmask, root_vi, cut, cap);
                    if (c <= 0) continue;
                    float qv = f[t] / (float)c;
                    for (int j = 1; j < k; ++j) {
                        if ((t >> (j - 1)) & 1) continue;
                        int in = 0;
                        for (int e = 0; e < c; ++e) if (cut[e] == mem[j]) { in = 1; break; }
                        if (in) f[t | (1 << (j - 1))] = f[t | (1 << (j - 1))] + qv;
                    }
                }
                double den = T->b[root_vi] * (double)f[full];
                double pa = T->C / den;
                u32 a0 = us_next(&rs), a1 = us_next(&rs);
                u64 v53 = ((((u64)a0) << 32) | (u64)a1) >> 11;
                double uu = (double)v53 * (1.0 / 9007199254740992.0);
                if (uu < pa) ok = 1;
            }
        }
        for (int j = 0; j < k; ++j) nodes_out[(i64)i * k + j] = ok ? mem[j] : -1;
    }
    free(cut);
    if (trials_out) *trials_out = trials;
    return 0;
}
