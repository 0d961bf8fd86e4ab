// ugs_wl.cu -- Weisfeiler-Lehman graphlet classification on sm_100a.
//
// Replaces, for all samples of a batch at once, the per-sample Python/NetworkX path of the reference:
//   compute_wl_hash              gps/utils/wl_vocab.py:21-67     nx.Graph from the sample's local edge list, node
//                                                                attribute = str(degree), then
//                                                                nx.weisfeiler_lehman_graph_hash(node_attr, iterations)
//   extract_subgraph_from_batch  gps/utils/wl_vocab.py:70-108    num_nodes = #(row entries >= 0), edges = the sample's
//                                                                edge_ptr range of edge_index_t
//   _compute_wl_ids / hash_to_id gps/models/ss_gnn_wl.py:210-247, wl_vocab.py:206-217   hash -> vocabulary id
// The output is the SAME hash NetworkX prints (its 16-byte BLAKE2b digest; the host turns it into the hex string), so
// vocabularies pickled by the reference keep working. NetworkX's algorithm (graph_hashing.py, node_attr given):
//   label_0(v) = str(deg v)   (nx.Graph: parallel entries merge, a self loop adds 2 to the degree and makes v its own
//                              neighbour once), or, with use_node_features, md5(feature row bytes).hexdigest()[:8]
//   label_{t+1}(v) = hex(blake2b_16(label_t(v) + "".join(sorted(label_t(u) for u in N(v)))))
//   after every iteration the sorted (label, count) pairs are appended to a list; the hash is
//   hex(blake2b_16(str(tuple(list)))), Python's tuple repr: "(('ab..', 2), ('cd..', 1))", "(x,)" for one element.
// One thread per sample (k <= 16): BLAKE2b's 12 rounds are fully unrolled so the message words stay in registers;
// labels live in local memory as byte strings; the final tuple string is hashed as it is produced (128-byte blocks).
#include "ugs_common.cuh"

#define WL_MAX_K 16
#define WL_LAB 32  // a label is at most 32 bytes (hex of a 16-byte digest)

__device__ __forceinline__ u64 rotr64(u64 x, int r) { return (x >> r) | (x << (64 - r)); }

struct B2b {
    u64 h[8];
    u64 t;
    int n;
    __align__(8) unsigned char buf[128];
};

#define B2B_G(a, b, c, d, x, y)                   \
    do {                                           \
        v[a] = v[a] + v[b] + (x);                  \
        v[d] = rotr64(v[d] ^ v[a], 32);            \
        v[c] = v[c] + v[d];                        \
        v[b] = rotr64(v[b] ^ v[c], 24);            \
        v[a] = v[a] + v[b] + (y);                  \
        v[d] = rotr64(v[d] ^ v[a], 16);            \
        v[c] = v[c] + v[d];                        \
        v[b] = rotr64(v[b] ^ v[c], 63);            \
    } while (0)

#define B2B_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                \
        B2B_G(0, 4, 8, 12, m[s0], m[s1]);                                               \
        B2B_G(1, 5, 9, 13, m[s2], m[s3]);                                               \
        B2B_G(2, 6, 10, 14, m[s4], m[s5]);                                              \
        B2B_G(3, 7, 11, 15, m[s6], m[s7]);                                              \
        B2B_G(0, 5, 10, 15, m[s8], m[s9]);                                              \
        B2B_G(1, 6, 11, 12, m[s10], m[s11]);                                            \
        B2B_G(2, 7, 8, 13, m[s12], m[s13]);                                             \
        B2B_G(3, 4, 9, 14, m[s14], m[s15]);                                             \
    } while (0)

__device__ __noinline__ void b2b_compress(B2b& S, bool last) {
    const u64 IV0 = 0x6a09e667f3bcc908ull, IV1 = 0xbb67ae8584caa73bull, IV2 = 0x3c6ef372fe94f82bull,
              IV3 = 0xa54ff53a5f1d36f1ull, IV4 = 0x510e527fade682d1ull, IV5 = 0x9b05688c2b3e6c1full,
              IV6 = 0x1f83d9abfb41bd6bull, IV7 = 0x5be0cd19137e2179ull;
    u64 m[16], v[16];
    const u64* b = reinterpret_cast<const u64*>(S.buf);
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = b[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = S.h[i];
    v[8] = IV0; v[9] = IV1; v[10] = IV2; v[11] = IV3;
    v[12] = IV4 ^ S.t; v[13] = IV5; v[14] = last ? ~IV6 : IV6; v[15] = IV7;
    B2B_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    B2B_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    B2B_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    B2B_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    B2B_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    B2B_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    B2B_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    B2B_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    B2B_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    B2B_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
    B2B_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    B2B_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
#pragma unroll
    for (int i = 0; i < 8; ++i) S.h[i] ^= v[i] ^ v[i + 8];
}

__device__ __forceinline__ void b2b_init(B2b& S) {  // digest_size = 16, no key, fanout = depth = 1
    S.h[0] = 0x6a09e667f3bcc908ull ^ 0x01010000ull ^ 16ull;
    S.h[1] = 0xbb67ae8584caa73bull; S.h[2] = 0x3c6ef372fe94f82bull; S.h[3] = 0xa54ff53a5f1d36f1ull;
    S.h[4] = 0x510e527fade682d1ull; S.h[5] = 0x9b05688c2b3e6c1full; S.h[6] = 0x1f83d9abfb41bd6bull;
    S.h[7] = 0x5be0cd19137e2179ull;
    S.t = 0; S.n = 0;
}
__device__ __forceinline__ void b2b_put(B2b& S, unsigned char c) {
    if (S.n == 128) { S.t += 128; b2b_compress(S, false); S.n = 0; }  // a full block is only compressed when more follows
    S.buf[S.n++] = c;
}
__device__ __forceinline__ void b2b_put_str(B2b& S, const unsigned char* p, int len) {
    for (int i = 0; i < len; ++i) b2b_put(S, p[i]);
}
__device__ __forceinline__ void b2b_final(B2b& S, u64& d0, u64& d1) {
    S.t += (u64)S.n;
    for (int i = S.n; i < 128; ++i) S.buf[i] = 0;
    b2b_compress(S, true);
    d0 = S.h[0]; d1 = S.h[1];
}
__device__ __forceinline__ void hex16(u64 d0, u64 d1, unsigned char* out) {  // hexdigest(): bytes in memory order
    const char* hx = "0123456789abcdef";
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned b0 = (unsigned)(d0 >> (8 * i)) & 0xFFu, b1 = (unsigned)(d1 >> (8 * i)) & 0xFFu;
        out[2 * i] = hx[b0 >> 4]; out[2 * i + 1] = hx[b0 & 15];
        out[16 + 2 * i] = hx[b1 >> 4]; out[16 + 2 * i + 1] = hx[b1 & 15];
    }
}
__device__ __forceinline__ int put_dec(unsigned char* out, unsigned v) {  // str(int) for v < 1000
    int n = 0;
    if (v >= 100) out[n++] = (unsigned char)('0' + v / 100);
    if (v >= 10) out[n++] = (unsigned char)('0' + (v / 10) % 10);
    out[n++] = (unsigned char)('0' + v % 10);
    return n;
}
// Python str comparison: lexicographic by byte, a proper prefix is smaller
__device__ __forceinline__ int lab_cmp(const unsigned char* a, int la, const unsigned char* b, int lb) {
    const int n = la < lb ? la : lb;
    for (int i = 0; i < n; ++i)
        if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return la == lb ? 0 : (la < lb ? -1 : 1);
}

// ---- MD5 of every feature row: the node attribute when use_node_features = True (wl_vocab.py:45-49:
//      hashlib.md5(node_features[i].numpy().tobytes()).hexdigest()[:8] -> the first four digest bytes) ------------------
__device__ __forceinline__ u32 rotl32(u32 x, int r) { return (x << r) | (x >> (32 - r)); }

__device__ void md5_block(u32* h, const unsigned char* blk) {
    const u32 K[64] = {
        0xd76aa478u, 0xe8c7b756u, 0x242070dbu, 0xc1bdceeeu, 0xf57c0fafu, 0x4787c62au, 0xa8304613u, 0xfd469501u,
        0x698098d8u, 0x8b44f7afu, 0xffff5bb1u, 0x895cd7beu, 0x6b901122u, 0xfd987193u, 0xa679438eu, 0x49b40821u,
        0xf61e2562u, 0xc040b340u, 0x265e5a51u, 0xe9b6c7aau, 0xd62f105du, 0x02441453u, 0xd8a1e681u, 0xe7d3fbc8u,
        0x21e1cde6u, 0xc33707d6u, 0xf4d50d87u, 0x455a14edu, 0xa9e3e905u, 0xfcefa3f8u, 0x676f02d9u, 0x8d2a4c8au,
        0xfffa3942u, 0x8771f681u, 0x6d9d6122u, 0xfde5380cu, 0xa4beea44u, 0x4bdecfa9u, 0xf6bb4b60u, 0xbebfbc70u,
        0x289b7ec6u, 0xeaa127fau, 0xd4ef3085u, 0x04881d05u, 0xd9d4d039u, 0xe6db99e5u, 0x1fa27cf8u, 0xc4ac5665u,
        0xf4292244u, 0x432aff97u, 0xab9423a7u, 0xfc93a039u, 0x655b59c3u, 0x8f0ccc92u, 0xffeff47du, 0x85845dd1u,
        0x6fa87e4fu, 0xfe2ce6e0u, 0xa3014314u, 0x4e0811a1u, 0xf7537e82u, 0xbd3af235u, 0x2ad7d2bbu, 0xeb86d391u};
    const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                       5, 9, 14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21,
                       6, 10, 15, 21, 6, 10, 15, 21};
    u32 M[16];
    for (int i = 0; i < 16; ++i)
        M[i] = (u32)blk[4 * i] | ((u32)blk[4 * i + 1] << 8) | ((u32)blk[4 * i + 2] << 16) | ((u32)blk[4 * i + 3] << 24);
    u32 a = h[0], b = h[1], c = h[2], d = h[3];
    for (int i = 0; i < 64; ++i) {
        u32 f; int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
        else { f = c ^ (b | ~d); g = (7 * i) & 15; }
        const u32 t = d;
        d = c; c = b;
        b = b + rotl32(a + f + K[i] + M[g], S[i]);
        a = t;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
}

// attr32[v] = first four bytes of md5(row v) (memory order: byte 0 in the low bits), one thread per row
__global__ void __launch_bounds__(128)
k_md5_rows(const unsigned char* __restrict__ x, i64 n_rows, i64 row_bytes, u32* __restrict__ attr32) {
    const i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_rows) return;
    const unsigned char* p = x + v * row_bytes;
    u32 h[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    unsigned char blk[64];
    i64 off = 0;
    for (; off + 64 <= row_bytes; off += 64) {
        for (int i = 0; i < 64; ++i) blk[i] = p[off + i];
        md5_block(h, blk);
    }
    const int rem = (int)(row_bytes - off);
    for (int i = 0; i < 64; ++i) blk[i] = (i < rem) ? p[off + i] : (i == rem ? 0x80 : 0);
    if (rem >= 56) {
        md5_block(h, blk);
        for (int i = 0; i < 64; ++i) blk[i] = 0;
    }
    const u64 bits = (u64)row_bytes * 8ull;
    for (int i = 0; i < 8; ++i) blk[56 + i] = (unsigned char)(bits >> (8 * i));
    md5_block(h, blk);
    attr32[v] = h[0];
}

__global__ void __launch_bounds__(128)
k_wl_hash(const i64* __restrict__ nodes, i64 B, int k, const i64* __restrict__ ei_u, const i64* __restrict__ ei_v,
          const i64* __restrict__ edge_ptr, int iterations, const u32* __restrict__ attr32, i64 n_attr,
          u64* __restrict__ digest_out, i32* __restrict__ num_nodes_out) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x; s < B; s += stride) {
        int nn = 0;
        for (int j = 0; j < k; ++j) nn += (nodes[s * k + j] >= 0) ? 1 : 0;  // wl_vocab.py:92-93
        num_nodes_out[s] = nn;
        if (nn == 0) { digest_out[2 * s] = 0; digest_out[2 * s + 1] = 0; continue; }
        u32 adj[WL_MAX_K];
#pragma unroll
        for (int j = 0; j < WL_MAX_K; ++j) adj[j] = 0;
        for (i64 e = edge_ptr[s]; e < edge_ptr[s + 1]; ++e) {  // nx.Graph.add_edges_from: undirected, simple
            const int u = (int)ei_u[e], v = (int)ei_v[e];
            if (u >= 0 && u < WL_MAX_K && v >= 0 && v < WL_MAX_K) { adj[u] |= 1u << v; adj[v] |= 1u << u; }
        }
        unsigned char lab[2][WL_MAX_K][WL_LAB];
        unsigned char len[2][WL_MAX_K];
        if (attr32) {  // attr = md5(feature row of the v-th VALID node).hexdigest()[:8]   (wl_vocab.py:45-49,99-103)
            int v = 0;
            for (int j = 0; j < k && v < nn; ++j) {
                const i64 id = nodes[s * k + j];
                if (id < 0) continue;
                const u32 a = (id < n_attr) ? attr32[id] : 0u;
                const char* hx = "0123456789abcdef";
                for (int b = 0; b < 4; ++b) {
                    const unsigned byte = (a >> (8 * b)) & 0xFFu;
                    lab[0][v][2 * b] = hx[byte >> 4]; lab[0][v][2 * b + 1] = hx[byte & 15];
                }
                len[0][v] = 8;
                ++v;
            }
        } else {
            for (int v = 0; v < nn; ++v) {  // attr = str(G.degree(v)); a self loop counts twice
                const unsigned deg = (unsigned)__popc(adj[v] & ~(1u << v)) + (((adj[v] >> v) & 1u) ? 2u : 0u);
                len[0][v] = (unsigned char)put_dec(lab[0][v], deg);
            }
        }
        B2b F;  // the hash of str(tuple(subgraph_hash_counts)), fed as the string is produced
        b2b_init(F);
        b2b_put(F, '(');
        int entries = 0;
        int cur = 0;
        for (int it = 0; it < iterations; ++it) {
            const int nxt = cur ^ 1;
            for (int v = 0; v < nn; ++v) {
                // neighbours of v sorted by label (insertion sort of indices)
                int ord[WL_MAX_K];
                int cnt = 0;
                u32 mnb = adj[v] & ((nn >= 32) ? ~0u : ((1u << nn) - 1u));
                while (mnb) {
                    const int u = __ffs(mnb) - 1;
                    mnb &= mnb - 1;
                    int p = cnt++;
                    while (p > 0 && lab_cmp(lab[cur][ord[p - 1]], len[cur][ord[p - 1]], lab[cur][u], len[cur][u]) > 0) {
                        ord[p] = ord[p - 1];
                        --p;
                    }
                    ord[p] = u;
                }
                B2b H;
                b2b_init(H);
                b2b_put_str(H, lab[cur][v], len[cur][v]);
                for (int q = 0; q < cnt; ++q) b2b_put_str(H, lab[cur][ord[q]], len[cur][ord[q]]);
                u64 d0, d1;
                b2b_final(H, d0, d1);
                hex16(d0, d1, lab[nxt][v]);
                len[nxt][v] = 32;
            }
            // Counter(node_labels.values()) sorted by label
            int ord[WL_MAX_K];
            for (int v = 0; v < nn; ++v) {
                int p = v;
                while (p > 0 && lab_cmp(lab[nxt][ord[p - 1]], 32, lab[nxt][v], 32) > 0) { ord[p] = ord[p - 1]; --p; }
                ord[p] = v;
            }
            for (int a = 0; a < nn;) {
                int b = a + 1;
                while (b < nn && lab_cmp(lab[nxt][ord[a]], 32, lab[nxt][ord[b]], 32) == 0) ++b;
                if (entries++) { b2b_put(F, ','); b2b_put(F, ' '); }
                b2b_put(F, '('); b2b_put(F, '\'');
                b2b_put_str(F, lab[nxt][ord[a]], 32);
                b2b_put(F, '\''); b2b_put(F, ','); b2b_put(F, ' ');
                unsigned char dec[4];
                const int dl = put_dec(dec, (unsigned)(b - a));
                b2b_put_str(F, dec, dl);
                b2b_put(F, ')');
                a = b;
            }
            cur = nxt;
        }
        if (entries == 1) b2b_put(F, ',');
        b2b_put(F, ')');
        u64 d0, d1;
        b2b_final(F, d0, d1);
        digest_out[2 * s] = d0;
        digest_out[2 * s + 1] = d1;
    }
}

// ids[s] = vocabulary id of digest s (binary search over the vocabulary's digests sorted as (bswap d0, bswap d1),
// i.e. in hex-string order), unknown_id when absent or when the sample has no valid node (ss_gnn_wl.py:230-244).
__global__ void __launch_bounds__(256)
k_wl_lookup(const u64* __restrict__ digest, const i32* __restrict__ num_nodes, i64 B, const u64* __restrict__ vocab,
            const i64* __restrict__ vocab_ids, i64 V, i64 unknown_id, i64* __restrict__ ids_out) {
    const i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B) return;
    i64 id = unknown_id;
    if (num_nodes[s] > 0 && V > 0) {
        const u64 k0 = __byte_perm((u32)(digest[2 * s] >> 32), 0, 0x0123) |
                       ((u64)__byte_perm((u32)digest[2 * s], 0, 0x0123) << 32);
        const u64 k1 = __byte_perm((u32)(digest[2 * s + 1] >> 32), 0, 0x0123) |
                       ((u64)__byte_perm((u32)digest[2 * s + 1], 0, 0x0123) << 32);
        i64 a = 0, b = V;
        while (a < b) {
            const i64 mid = (a + b) >> 1;
            const u64 m0 = vocab[2 * mid], m1 = vocab[2 * mid + 1];
            if (m0 < k0 || (m0 == k0 && m1 < k1)) a = mid + 1; else b = mid;
        }
        if (a < V && vocab[2 * a] == k0 && vocab[2 * a + 1] == k1) id = vocab_ids[a];
    }
    ids_out[s] = id;
}

extern "C" int ugs_wl_node_attr(const void* x, int64_t n_rows, int64_t row_bytes, uint32_t* attr32_out, void* stream_) {
    if (n_rows < 0 || row_bytes < 0) return ugs_set_error(UGS_EINVAL, "negative size");
    if (n_rows == 0) return UGS_OK;
    if (!attr32_out || (!x && row_bytes > 0)) return ugs_set_error(UGS_EINVAL, "NULL argument");
    k_md5_rows<<<(unsigned)((n_rows + 127) / 128), 128, 0, (cudaStream_t)stream_>>>((const unsigned char*)x, n_rows,
                                                                                  row_bytes, attr32_out);
    UGS_LAUNCH_CHECK("k_md5_rows");
    return UGS_OK;
}

extern "C" int ugs_wl_hash(const int64_t* nodes, int64_t B, int k, const int64_t* edge_index_s, int64_t total,
                           const int64_t* edge_ptr, int iterations, const uint32_t* node_attr32, int64_t n_attr,
                           uint64_t* digest_out, int32_t* num_nodes_out, void* stream_) {
    if (B < 0 || total < 0) return ugs_set_error(UGS_EINVAL, "negative size");
    if (k < 1 || k > WL_MAX_K) return ugs_set_error(UGS_ELIMIT, "WL hashing supports k in [1, %d]", WL_MAX_K);
    if (iterations <= 0) return ugs_set_error(UGS_EINVAL, "The WL algorithm requires that `iterations` be positive");
    if (B == 0) return UGS_OK;
    if (!nodes || !edge_ptr || !digest_out || !num_nodes_out) return ugs_set_error(UGS_EINVAL, "NULL argument");
    if (total > 0 && !edge_index_s) return ugs_set_error(UGS_EINVAL, "edge_index is NULL");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (B + 127) / 128;
    const int64_t cap = (int64_t)sms * 8;
    if (blocks > cap) blocks = cap;
    k_wl_hash<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream_>>>(nodes, B, k, edge_index_s, edge_index_s + total, edge_ptr,
                                                                  iterations, node_attr32, n_attr, (u64*)digest_out,
                                                                  num_nodes_out);
    UGS_LAUNCH_CHECK("k_wl_hash");
    return UGS_OK;
}

extern "C" int ugs_wl_lookup(const uint64_t* digest, const int32_t* num_nodes, int64_t B, const uint64_t* vocab_keys,
                             const int64_t* vocab_ids, int64_t V, int64_t unknown_id, int64_t* ids_out, void* stream_) {
    if (B < 0 || V < 0) return ugs_set_error(UGS_EINVAL, "negative size");
    if (B == 0) return UGS_OK;
    if (!digest || !num_nodes || !ids_out || (V > 0 && (!vocab_keys || !vocab_ids))) return ugs_set_error(UGS_EINVAL, "NULL argument");
    k_wl_lookup<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream_>>>((const u64*)digest, num_nodes, B,
                                                                               (const u64*)vocab_keys, vocab_ids, V,
                                                                               unknown_id, ids_out);
    UGS_LAUNCH_CHECK("k_wl_lookup");
    return UGS_OK;
}
