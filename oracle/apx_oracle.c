/*
 * oracle/apx_oracle.c -- TEST INFRASTRUCTURE ONLY (see ugs_oracle.c header for who may load it).
 *
 * Plain-C restatement of the reference's apx_ugs_sampler (Bressan APX-UGS, Algorithms 5,7,8,9,10):
 * src/samplers/apx_ugs_sampler/src/apx_ugs_sampler.cpp -- constants and quirks in SURVEY.md section 8 (a14).
 * Every function cites the reference lines it follows.
 *
 * Parity status: the reference draws from one sequential std::mt19937_64 with std::uniform_*_distribution
 * (implementation-defined streams, apx_ugs_sampler.hpp:12-24), so sample VALUES cannot be pinned; what is pinned is
 * (1) the deterministic pieces (graph construction, n, sorted+deduplicated adjacency, the initial order, the
 * small-degree fix-up that makes every estimate exact) through tests on the real reference's outputs, and (2) the
 * output LAW, by a two-sample chi-square against frequencies recorded from the real reference
 * (tests/golden/apx_ref_counts.json, ~10 s per reference sample). This oracle and the CUDA path share an explicit,
 * random-access draw rule so that they can be compared bit-exactly:
 *   word(c0,c1,c2,c3)[i] = Philox4x32-10(key = seed)(c0,c1,c2,c3)[i]; integer draws are mulhi32(word, range);
 *   reals are ((w_a << 32 | w_b) >> 11) * 2^-53.
 *   APX-DD step idx, draw i:               (i >> 2, idx, 0xA0000001, 0)[i & 3]
 *   trial t of sample s, root / accept:    (t, s, 0xA0000002, 0)[0,1] / [2,3]
 *   growth step i: cuts of member a, draw j: ((j >> 2) + 32*a + 512*i, t, s, 0xA0000003)[j & 3]
 *   growth step i: member / neighbour pick: (i, t, s, 0xA0000004)[0,1] / [2]
 *   APX-PROB permutation q, step i, member a, draw j: ((j >> 2) + 32*a + 512*i, t, s, 0xA0000005 | q << 8)[j & 3]
 * Trials of one sample are independent, so "the first accepted trial" is well defined without a sequential stream.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;
typedef int32_t i32;
typedef uint32_t u32;
typedef uint64_t u64;

void oracle_philox(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32* out);

#define APX_MAX_TRIALS 1000000
#define APX_MAX_PERMS 720

typedef struct {
    int n, k;
    i64* rowptr;   /* n+1 */
    i32* adj;      /* sorted, deduplicated neighbours (a self loop makes a vertex its own neighbour) */
    i32* deg;
    i32* order;    /* DDOrder::order */
    i32* position; /* DDOrder::position */
    double* est;   /* DDOrder::estimates */
    double Z;
    u64 seed;
    double eps;
} ApxGraph;

static u32 word(const ApxGraph* G, u32 c0, u32 c1, u32 c2, u32 c3, int i) {
    u32 w[4];
    oracle_philox(c0, c1, c2, c3, (u32)G->seed, (u32)(G->seed >> 32), w);
    return w[i];
}
static inline u32 mulhi32(u32 a, u32 b) { return (u32)(((u64)a * (u64)b) >> 32); }
static double u53(u32 a, u32 b) { return (double)(((((u64)a) << 32) | (u64)b) >> 11) * (1.0 / 9007199254740992.0); }

static int cmp_i32(const void* a, const void* b) { i32 x = *(const i32*)a, y = *(const i32*)b; return (x > y) - (x < y); }

/* Graph::Graph -- apx_ugs_sampler.cpp:10-41. Columns [e0,e1) only (ptr is read as an EDGE range), n = max id + 1. */
static void build_graph(ApxGraph* G, const i64* row0, const i64* row1, i64 e0, i64 e1) {
    int n = 0;
    for (i64 i = e0; i < e1; ++i) {
        if ((int)row0[i] + 1 > n) n = (int)row0[i] + 1;
        if ((int)row1[i] + 1 > n) n = (int)row1[i] + 1;
    }
    G->n = n;
    i64* cnt = (i64*)calloc((size_t)n + 1, sizeof(i64));
    for (i64 i = e0; i < e1; ++i) { cnt[row0[i] + 1]++; cnt[row1[i] + 1]++; }
    for (int v = 0; v < n; ++v) cnt[v + 1] += cnt[v];
    i32* tmp = (i32*)malloc(sizeof(i32) * (size_t)(cnt[n] + 1));
    i64* cur = (i64*)malloc(sizeof(i64) * (size_t)(n + 1));
    memcpy(cur, cnt, sizeof(i64) * (size_t)(n + 1));
    for (i64 i = e0; i < e1; ++i) { tmp[cur[row0[i]]++] = (i32)row1[i]; tmp[cur[row1[i]]++] = (i32)row0[i]; }
    G->rowptr = (i64*)calloc((size_t)n + 1, sizeof(i64));
    G->adj = (i32*)malloc(sizeof(i32) * (size_t)(cnt[n] + 1));
    G->deg = (i32*)calloc((size_t)(n ? n : 1), sizeof(i32));
    i64 out = 0;
    for (int v = 0; v < n; ++v) {
        i64 a = cnt[v], b = cnt[v + 1];
        qsort(tmp + a, (size_t)(b - a), sizeof(i32), cmp_i32);
        G->rowptr[v] = out;
        for (i64 p = a; p < b; ++p) if (p == a || tmp[p] != tmp[p - 1]) G->adj[out++] = tmp[p];
        G->deg[v] = (i32)(out - G->rowptr[v]);
    }
    G->rowptr[n] = out;
    free(cnt); free(tmp); free(cur);
}

static int has_edge(const ApxGraph* G, int u, int v) { /* :43-46 */
    if (G->deg[u] < G->deg[v]) { int t = u; u = v; v = t; }
    const i32* a = G->adj + G->rowptr[v];
    int lo = 0, hi = G->deg[v] - 1;
    while (lo <= hi) { int mid = (lo + hi) >> 1; if (a[mid] == u) return 1; if (a[mid] < u) lo = mid + 1; else hi = mid - 1; }
    return 0;
}

/* order comparator of apx_dd (:68-71): scores descending, ids descending */
static const double* g_scores;
static int cmp_order(const void* pa, const void* pb) {
    i32 a = *(const i32*)pa, b = *(const i32*)pb;
    if (g_scores[a] != g_scores[b]) return g_scores[a] > g_scores[b] ? -1 : 1;
    return a > b ? -1 : (a < b ? 1 : 0);
}

/* apx_dd -- apx_ugs_sampler.cpp:52-169 */
static void apx_dd(ApxGraph* G, double beta) {
    const int n = G->n, k = G->k;
    const double eta = pow(beta, 1.0 / (double)(k - 1)) / (6.0 * k * k);
    const int h = (int)ceil(10.0 / (eta * eta) * log((double)n));
    double* scores = (double*)malloc(sizeof(double) * (size_t)(n ? n : 1));
    G->order = (i32*)malloc(sizeof(i32) * (size_t)(n ? n : 1));
    G->position = (i32*)malloc(sizeof(i32) * (size_t)(n ? n : 1));
    G->est = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
    for (int v = 0; v < n; ++v) { scores[v] = (double)G->deg[v]; G->order[v] = v; }
    g_scores = scores;
    qsort(G->order, (size_t)n, sizeof(i32), cmp_order);
    for (int i = 0; i < n; ++i) G->position[G->order[i]] = i;
    for (int idx = 0; idx < n; ++idx) {
        const int v = G->order[idx];
        if (G->deg[v] == 0) { G->est[v] = 0.0; continue; }
        int count = 0;
        const i32* nb = G->adj + G->rowptr[v];
        for (int i = 0; i < h; ++i) {
            const int u = nb[mulhi32(word(G, (u32)(i >> 2), (u32)idx, 0xA0000001u, 0u, i & 3), (u32)G->deg[v])];
            if (G->position[v] < G->position[u]) count++;
        }
        const double threshold = 2.0 * eta * h;
        if (count >= threshold) {
            G->est[v] = pow((double)G->deg[v], 5.0);
        } else {
            G->est[v] = 0.0;
            scores[v] = 3.0 * eta * G->deg[v];
            qsort(G->order + idx, (size_t)(n - idx), sizeof(i32), cmp_order); /* re-sort the tail (:111-115) */
            for (int i = idx; i < n; ++i) G->position[G->order[i]] = i;
        }
    }
    /* small-degree fix-up (:124-166): deg <= k/eta (practically every vertex) -> exact estimate */
    const double thr = (double)k / eta;
    i32* q = (i32*)malloc(sizeof(i32) * (size_t)(k + 1));
    for (int v = 0; v < n; ++v) {
        if (!(G->deg[v] <= thr)) continue;
        int qn = 0, found = 0;
        q[qn++] = v;
        for (int head = 0; head < qn && qn < k && !found; ++head) {
            const int u = q[head];
            for (i64 p = G->rowptr[u]; p < G->rowptr[u + 1]; ++p) {
                const int w = G->adj[p];
                int seen = 0; for (int a = 0; a < qn; ++a) if (q[a] == w) { seen = 1; break; }
                if (seen || !(G->position[v] < G->position[w])) continue;
                q[qn++] = w;
                if (qn >= k) { found = 1; break; }
            }
        }
        if (found) {
            int d = 0;
            for (i64 p = G->rowptr[v]; p < G->rowptr[v + 1]; ++p) if (G->position[v] < G->position[G->adj[p]]) d++;
            G->est[v] = pow((double)d, 5.0);
        } else {
            G->est[v] = 0.0;
        }
    }
    free(q); free(scores);
}

/* estimate_cuts -- :175-237. With the reference's constants ell_raw is astronomically large, so h = 100 and
 * ell = 50 always; the formulas are kept so that any change of constants shows up. Returns the SUM of cuts and,
 * through cuts[], the per-member values. stream: c3 of the draws; base0 = 512*i is added by the caller. */
static double estimate_cuts(const ApxGraph* G, int v, const int* U, int sz, double alpha, double beta, double delta,
                            u32 base0, u32 t, u32 s, u32 stream, double* cuts) {
    const int k = G->k;
    const double ell_raw = 1.0 / (k * delta * alpha * alpha);
    const double h_double = ell_raw * ell_raw * log(k / beta);
    int h;
    if (isinf(h_double) || h_double > 100) h = 100; else if (h_double < 10.0) h = 10; else h = (int)ceil(h_double);
    const double ell = ell_raw < h * 0.5 ? ell_raw : h * 0.5;
    double total = 0.0;
    for (int a = 0; a < sz; ++a) {
        const int u = U[a];
        cuts[a] = 0.0;
        if (G->deg[u] == 0) continue;
        int count = 0;
        for (int j = 0; j < h; ++j) {
            const int w = G->adj[G->rowptr[u] + mulhi32(word(G, base0 + (u32)(j >> 2) + 32u * (u32)a, t, s, stream, j & 3), (u32)G->deg[u])];
            int inU = 0; for (int b = 0; b < sz; ++b) if (U[b] == w) { inU = 1; break; }
            if (G->position[v] < G->position[w] && !inU) count++;
        }
        if (count >= ell) cuts[a] = (double)(G->deg[u] * count) / (double)h;
        total += cuts[a];
    }
    return total;
}

/* One trial of apx_ugs_sample_one (:412-451). Returns 1 and fills S on acceptance. */
static int apx_trial(const ApxGraph* G, u32 t, u32 s, double alpha, double beta, int* S) {
    const int k = G->k, n = G->n;
    const double eps = G->eps;
    /* root ~ estimates (:414-426) */
    const double r = u53(word(G, t, s, 0xA0000002u, 0u, 0), word(G, t, s, 0xA0000002u, 0u, 1)) * G->Z;
    double cumsum = 0.0;
    int v = 0;
    for (int u = 0; u < n; ++u) { cumsum += G->est[u]; if (r <= cumsum) { v = u; break; } }
    if (G->est[v] <= 0.0) return 0;
    /* APX-RAND-GROW (:243-312) */
    const double gamma = eps * pow(3.0, -k) * pow((double)k, -2.0);
    const double delta = gamma / pow((double)k, 4.0);
    double cuts[16];
    int sz = 1;
    S[0] = v;
    for (int i = 1; i < k; ++i) {
        const double total = estimate_cuts(G, v, S, sz, alpha, beta, delta, 512u * (u32)i, t, s, 0xA0000003u, cuts);
        if (total <= 0.0) return 0;
        const double rr = u53(word(G, (u32)i, t, s, 0xA0000004u, 0), word(G, (u32)i, t, s, 0xA0000004u, 1)) * total;
        double cs = 0.0;
        int sel = S[0];
        for (int j = 0; j < sz; ++j) { cs += cuts[j]; if (rr <= cs) { sel = S[j]; break; } }
        int nvalid = 0;
        for (i64 p = G->rowptr[sel]; p < G->rowptr[sel + 1]; ++p) {
            const int w = G->adj[p];
            if (!(G->position[v] < G->position[w])) continue;
            int inS = 0; for (int b = 0; b < sz; ++b) if (S[b] == w) { inS = 1; break; }
            if (!inS) nvalid++;
        }
        if (nvalid == 0) return 0;
        int pick = (int)mulhi32(word(G, (u32)i, t, s, 0xA0000004u, 2), (u32)nvalid);
        int chosen = -1;
        for (i64 p = G->rowptr[sel]; p < G->rowptr[sel + 1]; ++p) {
            const int w = G->adj[p];
            if (!(G->position[v] < G->position[w])) continue;
            int inS = 0; for (int b = 0; b < sz; ++b) if (S[b] == w) { inS = 1; break; }
            if (inS) continue;
            if (pick == 0) { chosen = w; break; }
            pick--;
        }
        S[sz++] = chosen;
    }
    /* APX-PROB (:318-382): permutations of the sorted remaining vertices in next_permutation order, <= 720 */
    const double rho = eps * pow(3.0, -k) * pow((double)k, -2.0);
    int perm[16], rem[16];
    for (int j = 1; j < k; ++j) rem[j - 1] = S[j];
    qsort(rem, (size_t)(k - 1), sizeof(int), cmp_i32);
    double p_hat = 0.0;
    for (int q = 0; q < APX_MAX_PERMS; ++q) {
        perm[0] = v;
        for (int j = 0; j < k - 1; ++j) perm[j + 1] = rem[j];
        double pp = 1.0;
        for (int i = 0; i < k - 1; ++i) {
            int ni = 0;
            for (int a = 0; a <= i; ++a) if (has_edge(G, perm[a], perm[i + 1])) ni++;
            const double ci = estimate_cuts(G, v, perm, i + 1, alpha, beta / pow((double)k, 6.0), rho / (k * k),
                                            512u * (u32)i, t, s, 0xA0000005u | ((u32)q << 8), cuts);
            if (ci > 0.0) pp *= (double)ni / ci; else { pp = 0.0; break; }
        }
        p_hat += pp;
        /* std::next_permutation(rem) */
        int a = k - 3;
        while (a >= 0 && rem[a] >= rem[a + 1]) a--;
        if (a < 0) break;
        int b = k - 2;
        while (rem[b] <= rem[a]) b--;
        { int tt = rem[a]; rem[a] = rem[b]; rem[b] = tt; }
        for (int x = a + 1, y = k - 2; x < y; ++x, --y) { int tt = rem[x]; rem[x] = rem[y]; rem[y] = tt; }
    }
    if (p_hat <= 0.0) return 0;
    double acc = (beta / G->Z) * pow((double)k, -2.0) / (G->est[v] * p_hat);
    if (acc > 1.0) acc = 1.0;
    const double uu = u53(word(G, t, s, 0xA0000002u, 0u, 2), word(G, t, s, 0xA0000002u, 0u, 3));
    return uu < acc;
}

ApxGraph* apx_oracle_create(const i64* row0, const i64* row1, i64 e0, i64 e1, int k, u64 seed, double eps) {
    ApxGraph* G = (ApxGraph*)calloc(1, sizeof(ApxGraph));
    G->k = k; G->seed = seed; G->eps = eps;
    build_graph(G, row0, row1, e0, e1);
    if (G->n >= k) {
        apx_dd(G, eps / 2.0);
        G->Z = 0.0;
        for (int v = 0; v < G->n; ++v) G->Z += G->est[v];
    }
    return G;
}
void apx_oracle_free(ApxGraph* G) {
    if (!G) return;
    free(G->rowptr); free(G->adj); free(G->deg); free(G->order); free(G->position); free(G->est); free(G);
}
int apx_oracle_n(const ApxGraph* G) { return G->n; }
i64 apx_oracle_nnz(const ApxGraph* G) { return G->rowptr[G->n]; }
double apx_oracle_Z(const ApxGraph* G) { return G->Z; }
void apx_oracle_dump(const ApxGraph* G, i64* rowptr, i32* adj, i32* order, i32* position, double* est) {
    memcpy(rowptr, G->rowptr, sizeof(i64) * (size_t)(G->n + 1));
    memcpy(adj, G->adj, sizeof(i32) * (size_t)G->rowptr[G->n]);
    if (G->order) {
        memcpy(order, G->order, sizeof(i32) * (size_t)G->n);
        memcpy(position, G->position, sizeof(i32) * (size_t)G->n);
        memcpy(est, G->est, sizeof(double) * (size_t)G->n);
    }
}

/* sample_batch -- :461-521. samples_out: [m][k] (caller transposes to the reference's [k,S]); returns S <= m.
 * max_trials <= 0 means the reference's 10^6. trials_out[s] = index of the accepted trial or -1. */
int apx_oracle_sample(const ApxGraph* G, int m, int max_trials, i32* samples_out, i64* trials_out) {
    if (G->n < G->k || !(G->Z > 0.0)) { for (int s = 0; s < m; ++s) if (trials_out) trials_out[s] = -1; return 0; }
    if (max_trials <= 0) max_trials = APX_MAX_TRIALS;
    const double beta = G->eps / 2.0;
    const double alpha = pow(beta, 1.0 / (double)(G->k - 1)) / (6.0 * G->k * G->k * G->k);
    int S[16], got = 0;
    for (int s = 0; s < m; ++s) {
        int ok = 0;
        i64 tacc = -1;
        for (int t = 0; t < max_trials; ++t) if (apx_trial(G, (u32)t, (u32)s, alpha, beta, S)) { ok = 1; tacc = t; break; }
        if (trials_out) trials_out[s] = tacc;
        if (ok) { for (int j = 0; j < G->k; ++j) samples_out[(i64)got * G->k + j] = S[j]; got++; }
    }
    return got;
}
