/*
 * oracle/walk_oracle.c -- TEST INFRASTRUCTURE ONLY (see ugs_oracle.c header for the rules on who may load it).
 *
 * Plain-C restatements of two more reference samplers, both on the multigraph CSR of ugs_oracle.c (which is the
 * adjacency both of them build: every in-graph column adds v to row u and u to row v, column order):
 *
 *  (1) rwr_sampler  -- src/samplers/rwr_sampler/src/rwr_sampler.cpp:72-295 (walk :167-193).
 *  (2) uniform_sampler -- src/samplers/uniform_sampler/src/uniform_sampler.cpp:47-81,183-210: the set of connected
 *      induced k-subgraphs and a uniform pick among them.
 *
 * Parity status: PINNED for the deterministic parts -- the edge extraction of both samplers against the real
 * reference's outputs for its own node rows (tests/golden/walk_golden.npz), and the enumerated set against the
 * reference's support and a brute-force C(n,k) filter (tests/test_walk_oracle.py) -- and for the LAWS (exact
 * enumerators oracle/exact_law.py::p_rwr / uniform 1/T, chi-square against the real reference's frequencies).
 * Sample VALUES are not the reference's (rwr: SplitMix64 seeded per graph AND per OpenMP thread, :129; uniform: one
 * mt19937_64 + std::uniform_int_distribution, implementation-defined). This oracle and the CUDA path share explicit
 * rules instead:
 *   rwr:     Philox4x32-10, key = seed, counter = (sample i, word block, graph index, 0xC0000000); words in order:
 *            seed vertex = mulhi32(word, n); per walk step one word r = word * 2^-32 and, when the walk moves, one
 *            word for the row entry = mulhi32(word, row length).
 *   uniform: subsets listed by minimum vertex, then in the depth-first order of the ESU recursion that always extends
 *            by the smallest candidate first; pick = mulhi64(w0 << 32 | w1, count) with (w0, w1, ..) =
 *            Philox(counter = (sample i, 0, graph index, 0xE0000000)); the row is the subset in ascending order.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;
typedef int32_t i32;
typedef uint32_t u32;
typedef uint64_t u64;

void oracle_philox(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32* out); /* ugs_oracle.c */

static inline u32 mulhi32(u32 a, u32 b) { return (u32)(((u64)a * (u64)b) >> 32); }

typedef struct { u32 c0, c2, c3, k0, k1, block, idx; u32 w[4]; } Words;
static void words_init(Words* s, u32 sample, u64 graph, u32 tag, u64 seed) {
    s->c0 = sample; s->c2 = (u32)graph; s->c3 = tag; s->k0 = (u32)seed; s->k1 = (u32)(seed >> 32);
    s->block = 0; s->idx = 4;
}
static u32 words_next(Words* s) {
    if (s->idx == 4) { oracle_philox(s->c0, s->block++, s->c2, s->c3, s->k0, s->k1, s->w); s->idx = 0; }
    return s->w[s->idx++];
}

/* rwr_sampler.cpp:165-204 for m samples of one graph. nodes_out: m*k local ids in visiting order, -1 rows on failure. */
int rwr_oracle_sample_nodes(const i64* indptr, const i32* indices, int n, int m, int k, u64 seed, u64 graph_index,
                            double p_restart, i32* nodes_out) {
    i32* chosen = (i32*)malloc(sizeof(i32) * (size_t)(k > 0 ? k : 1));
    for (int i = 0; i < m; ++i) {
        i32* row = nodes_out + (i64)i * k;
        int size = 0;
        if (n > 0 && n >= k) { /* :139-162 */
            Words rs;
            words_init(&rs, (u32)i, graph_index, 0xC0000000u, seed);
            const int seed_node = (int)mulhi32(words_next(&rs), (u32)n);
            int cur = seed_node;
            chosen[size++] = cur;
            const long long max_iter = (long long)n * k * 10; /* :180 */
            long long it = 0;
            while (size < k && it < max_iter) {
                ++it;
                const double r = (double)words_next(&rs) * (1.0 / 4294967296.0);
                const i64 p0 = indptr[cur], p1 = indptr[cur + 1];
                if (r < p_restart || p0 == p1) cur = seed_node;                          /* :185-186 */
                else cur = indices[p0 + (i64)mulhi32(words_next(&rs), (u32)(p1 - p0))];  /* :188-189 */
                int seen = 0;
                for (int j = 0; j < size; ++j) seen |= (chosen[j] == cur);
                if (!seen) chosen[size++] = cur;                                         /* :190-193 */
            }
        }
        for (int j = 0; j < k; ++j) row[j] = (size == k) ? chosen[j] : -1;               /* :196-204 */
    }
    free(chosen);
    return 0;
}

/* ---- uniform_sampler: connected induced k-subsets of a graph with n <= 64 vertices ------------------------------ */
typedef struct { u64* out; i64 cnt, cap; const u64* adj; u64 gt; int k; } Esu;

static void esu_rec(Esu* E, u64 S, u64 ext, u64 NS, int size) {
    if (size == E->k) {
        if (E->out && E->cnt < E->cap) E->out[E->cnt] = S;
        E->cnt++;
        return;
    }
    while (ext) {
        const u64 low = ext & (~ext + 1ull);
        ext ^= low;
        int w = 0;
        while (!((low >> w) & 1ull)) ++w;
        esu_rec(E, S | low, ext | (E->adj[w] & E->gt & ~NS), NS | E->adj[w], size + 1);
    }
}

/* Writes up to cap subsets (64-bit vertex masks) to out (may be NULL) and returns their number; -1 if n > 64. */
i64 enum_oracle_list(const i64* indptr, const i32* indices, int n, int k, u64* out, i64 cap) {
    if (n > 64 || k < 1) return -1;
    u64 adj[64];
    memset(adj, 0, sizeof(adj));
    for (int v = 0; v < n; ++v)
        for (i64 p = indptr[v]; p < indptr[v + 1]; ++p)
            if (indices[p] != v) adj[v] |= 1ull << indices[p];
    Esu E; E.out = out; E.cnt = 0; E.cap = cap; E.adj = adj; E.k = k;
    for (int v = 0; v < n; ++v) {
        E.gt = (v >= 63) ? 0ull : (~0ull << (v + 1));
        esu_rec(&E, 1ull << v, adj[v] & E.gt, adj[v] | (1ull << v), 1);
    }
    return E.cnt;
}

/* uniform_sampler.cpp:183-210 for m samples of one graph given its list. nodes_out: m*k local ids ascending. */
int enum_oracle_sample_nodes(const u64* list, i64 count, int m, int k, u64 seed, u64 graph_index, i32* nodes_out) {
    for (int i = 0; i < m; ++i) {
        i32* row = nodes_out + (i64)i * k;
        if (count <= 0) { for (int j = 0; j < k; ++j) row[j] = -1; continue; }
        u32 w[4];
        oracle_philox((u32)i, 0u, (u32)graph_index, 0xE0000000u, (u32)seed, (u32)(seed >> 32), w);
        const u64 word = (((u64)w[0]) << 32) | (u64)w[1];
        const u64 idx = (u64)(((unsigned __int128)word * (unsigned __int128)(u64)count) >> 64);
        u64 S = list[idx];
        for (int j = 0; j < k; ++j) {
            int v = 0;
            while (!((S >> v) & 1ull)) ++v;
            S &= S - 1;
            row[j] = v;
        }
    }
    return 0;
}
