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
// One thread per sample (k <= 16): BLAKE2b's 12 rounds are fully unrolled with message block and state in registers;
// labels live in shared memory as 16-byte digests (not strings); the final tuple string is hashed as it is produced.
#include "ugs_common.cuh"

#define WL_MAX_K 16

__device__ __forceinline__ u64 rotr64(u64 x, int r) { return (x >> r) | (x << (64 - r)); }

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

__constant__ __align__(16) unsigned char b2b_sigma[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};

#define WLT 128  // threads per CTA of the WL kernel; every per-thread array in shared memory is [entries][WLT]

// BLAKE2b compression. The 16 message words of the thread sit in shared memory (column `mcol`, stride WLT: conflict
// free) and are fetched in the round's permuted order, so the twelve rounds are ONE loop body of eight G functions:
// the fully unrolled form (48 KB of code per call site) thrashes the instruction cache and ran 2x slower.
// t = message bytes hashed so far including this block; last = final block; words [nwords, 16) of the block are
// zeroed here first (padding of the last block).
__device__ __noinline__ void b2b_compress(u64* h, u64* mcol, int nwords, u64 t, bool last) {
    for (int i = nwords; i < 16; ++i) mcol[i * WLT] = 0ull;
    const u64 IV0 = 0x6a09e667f3bcc908ull, IV1 = 0xbb67ae8584caa73bull, IV2 = 0x3c6ef372fe94f82bull,
              IV3 = 0xa54ff53a5f1d36f1ull, IV4 = 0x510e527fade682d1ull, IV5 = 0x9b05688c2b3e6c1full,
              IV6 = 0x1f83d9abfb41bd6bull, IV7 = 0x5be0cd19137e2179ull;
    u64 v[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = h[i];
    v[8] = IV0; v[9] = IV1; v[10] = IV2; v[11] = IV3;
    v[12] = IV4 ^ t; v[13] = IV5; v[14] = last ? ~IV6 : IV6; v[15] = IV7;
    // One loop body for the twelve rounds (the fully unrolled form, 48 KB of code, is slower: 82 vs 69 ms -- instruction
    // cache). The round's permutation comes as one 16-byte constant load and its sixteen message words are fetched from
    // shared memory together, before the eight G functions, so that their latency overlaps.
#pragma unroll 1
    for (int r = 0; r < 12; ++r) {
        const uint4 sg = reinterpret_cast<const uint4*>(b2b_sigma)[r];
        u64 mw[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const u32 w = (i < 4) ? sg.x : (i < 8) ? sg.y : (i < 12) ? sg.z : sg.w;
            mw[i] = mcol[((w >> (8 * (i & 3))) & 0xFFu) * WLT];
        }
        B2B_G(0, 4, 8, 12, mw[0], mw[1]);
        B2B_G(1, 5, 9, 13, mw[2], mw[3]);
        B2B_G(2, 6, 10, 14, mw[4], mw[5]);
        B2B_G(3, 7, 11, 15, mw[6], mw[7]);
        B2B_G(0, 5, 10, 15, mw[8], mw[9]);
        B2B_G(1, 6, 11, 12, mw[10], mw[11]);
        B2B_G(2, 7, 8, 13, mw[12], mw[13]);
        B2B_G(3, 4, 9, 14, mw[14], mw[15]);
    }
#define B2B_M(i) 0
#undef B2B_M
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}

__device__ __forceinline__ void b2b_init(u64* h) {  // digest_size = 16, no key, fanout = depth = 1
    h[0] = 0x6a09e667f3bcc908ull ^ 0x01010000ull ^ 16ull;
    h[1] = 0xbb67ae8584caa73bull; h[2] = 0x3c6ef372fe94f82bull; h[3] = 0xa54ff53a5f1d36f1ull;
    h[4] = 0x510e527fade682d1ull; h[5] = 0x9b05688c2b3e6c1full; h[6] = 0x1f83d9abfb41bd6bull;
    h[7] = 0x5be0cd19137e2179ull;
}

__device__ __forceinline__ u64 bswap64(u64 x) {
    return ((u64)__byte_perm((u32)x, 0, 0x0123) << 32) | (u64)__byte_perm((u32)(x >> 32), 0, 0x0123);
}

// Eight hex characters (one little-endian message word, first character in the low byte) of four digest bytes given
// big-endian in w (first byte in the top bits): what hexdigest() prints for them.
__device__ __forceinline__ u64 hex8(u32 w) {
    const u32 hi = (w >> 4) & 0x0F0F0F0Fu, lo = w & 0x0F0F0F0Fu;
    u32 a = __byte_perm(hi, lo, 0x6273);  // characters 0..3: byte0 hi, byte0 lo, byte1 hi, byte1 lo
    u32 b = __byte_perm(hi, lo, 0x4051);  // characters 4..7: byte2 hi, byte2 lo, byte3 hi, byte3 lo
    a = a + 0x30303030u + (((a + 0x06060606u) >> 4) & 0x01010101u) * 0x27u;  // 0-9 -> '0'..'9', 10-15 -> 'a'..'f'
    b = b + 0x30303030u + (((b + 0x06060606u) >> 4) & 0x01010101u) * 0x27u;
    return ((u64)b << 32) | (u64)a;
}

// lexicographic order of two labels held as big-endian keys (equal to Python's str comparison of the hex / decimal
// strings: zero padding sorts a proper prefix first)
__device__ __forceinline__ bool key_lt(u64 a0, u64 a1, u64 b0, u64 b1) { return a0 < b0 || (a0 == b0 && a1 < b1); }

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

// One thread per sample. Labels are kept as 16-byte digests (two big-endian u64 keys per node, in shared memory, one
// column per thread), never as strings: the 32 hex characters a label contributes to a message are produced four message
// words at a time when the block is assembled, neighbour labels are sorted by comparing keys, and the tuple string of the
// final hash is streamed through an 8-byte shift queue into its own 128-byte block. First-iteration labels (decimal
// degree or eight hex characters of a feature hash) are at most 8 characters: key word 0 holds them big-endian and zero
// padded, key word 1 their length.
// Shared memory per CTA: labels [2 buffers][2 words][k][WLT] + the message block being compressed [16][WLT].
__global__ void __launch_bounds__(WLT)
k_wl_hash(const i64* __restrict__ nodes, i64 B, int k, const i64* __restrict__ ei_u, const i64* __restrict__ ei_v,
          const i64* __restrict__ edge_ptr, int iterations, const u32* __restrict__ attr32, i64 n_attr,
          u64* __restrict__ digest_out, i32* __restrict__ num_nodes_out) {
    extern __shared__ u64 wl_sm[];
    const int tid = threadIdx.x;
    u64* labs = wl_sm;                               // ((buf * 2 + word) * k + node) * WLT + tid
    u64* mbuf = wl_sm + (size_t)4 * k * WLT + tid;   // message block being compressed: word i at mbuf[i * WLT]
#define WL_LAB_AT(buf, word, node) labs[(size_t)((((buf) * 2 + (word)) * k) + (node)) * WLT + tid]
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 s = (i64)blockIdx.x * blockDim.x + tid; s < B; s += stride) {
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
        const u32 nmask = (nn >= 32) ? ~0u : ((1u << nn) - 1u);
        // ---- first labels ------------------------------------------------------------------------------------------
        if (attr32) {  // attr = md5(feature row of the v-th VALID node).hexdigest()[:8]   (wl_vocab.py:45-49,99-103)
            int v = 0;
            for (int j = 0; j < k && v < nn; ++j) {
                const i64 id = nodes[s * k + j];
                if (id < 0) continue;
                const u32 a = (id < n_attr) ? attr32[id] : 0u;   // digest bytes 0..3, byte 0 in the low bits
                WL_LAB_AT(0, 0, v) = bswap64(hex8(__byte_perm(a, 0, 0x0123)));  // characters big-endian
                WL_LAB_AT(0, 1, v) = 8;
                ++v;
            }
        } else {
            for (int v = 0; v < nn; ++v) {  // attr = str(G.degree(v)); a self loop counts twice
                const unsigned deg = (unsigned)__popc(adj[v] & ~(1u << v)) + (((adj[v] >> v) & 1u) ? 2u : 0u);
                u64 key; int len;
                if (deg >= 10) { key = ((u64)('0' + deg / 10) << 56) | ((u64)('0' + deg % 10) << 48); len = 2; }
                else { key = (u64)('0' + deg) << 56; len = 1; }
                WL_LAB_AT(0, 0, v) = key;
                WL_LAB_AT(0, 1, v) = (u64)len;
            }
        }
        // ---- the final hash, hash(str(tuple(counts))), fed as the string is produced ---------------------------------
        u64 fh[8];
        b2b_init(fh);
        // The string is queued as 8-byte words in a per-thread buffer and its complete 128-byte blocks are compressed
        // after every iteration, in a loop all lanes of the warp walk together. (Compressing inside f_put, whenever a
        // lane's block fills up, happens at a different moment in every lane: the 3000-instruction compression then runs
        // once per lane instead of once per warp.) One iteration appends at most 16 entries of 42 characters.
        u64 fs[WL_MAX_K * 6 + 20];
        u64 f_acc = 0, f_blocks = 0;  // shift queue; full blocks already compressed
        int f_nacc = 0, f_wi = 0;     // bytes in the queue; words in fs[]
        auto f_put = [&](u64 bytes, int len) {  // 1 <= len <= 8 characters, first character in the low byte
            f_acc |= bytes << (8 * f_nacc);
            f_nacc += len;
            if (f_nacc >= 8) {
                fs[f_wi++] = f_acc;
                f_nacc -= 8;
                f_acc = f_nacc ? (bytes >> (8 * (len - f_nacc))) : 0ull;
            }
        };
        // compress the complete blocks in fs[] (all of them: more data follows) and move the tail to the front
        auto f_drain = [&]() {
            const int nfull = f_wi >> 4;
            for (int b = 0; b < nfull; ++b) {
#pragma unroll
                for (int i = 0; i < 16; ++i) mbuf[i * WLT] = fs[16 * b + i];
                ++f_blocks;
                b2b_compress(fh, mbuf, 16, f_blocks * 128ull, false);
            }
            const int tail = f_wi - 16 * nfull;
            for (int i = 0; i < tail; ++i) fs[i] = fs[16 * nfull + i];
            f_wi = tail;
        };
        f_put('(', 1);
        int entries = 0;
        int cur = 0;
        for (int it = 0; it < iterations; ++it) {
            const int nxt = cur ^ 1;
            for (int v = 0; v < nn; ++v) {
                u32 rem = adj[v] & nmask;
                const int pieces = __popc(rem) + 1;  // own label, then the neighbours' labels in sorted order
                u64 h[8];
                b2b_init(h);
                u64 acc = 0, total = 0, blocks = 0;
                int nacc = 0, wi = 0;
                for (int p = 0; p < pieces; ++p) {
                    u64 a0, a1;
                    if (p == 0) { a0 = WL_LAB_AT(cur, 0, v); a1 = WL_LAB_AT(cur, 1, v); }
                    else {  // the smallest remaining neighbour label
                        int best = -1;
                        a0 = 0; a1 = 0;
                        for (u32 r = rem; r; r &= r - 1) {
                            const int u = __ffs(r) - 1;
                            const u64 k0 = WL_LAB_AT(cur, 0, u), k1 = WL_LAB_AT(cur, 1, u);
                            if (best < 0 || key_lt(k0, k1, a0, a1)) { best = u; a0 = k0; a1 = k1; }
                        }
                        rem &= ~(1u << best);
                    }
                    if (it > 0) {  // 32 hex characters: four whole message words
                        if (wi == 16) { ++blocks; b2b_compress(h, mbuf, 16, blocks * 128ull, false); wi = 0; }
                        mbuf[wi * WLT] = hex8((u32)(a0 >> 32)); mbuf[(wi + 1) * WLT] = hex8((u32)a0);
                        mbuf[(wi + 2) * WLT] = hex8((u32)(a1 >> 32)); mbuf[(wi + 3) * WLT] = hex8((u32)a1);
                        wi += 4;
                        total += 32;
                    } else {       // at most 8 characters, byte granular: through the shift queue
                        const int len = (int)a1;
                        const u64 bytes = bswap64(a0);  // first character in the low byte
                        total += (u64)len;
                        acc |= bytes << (8 * nacc);
                        nacc += len;
                        if (nacc >= 8) {
                            if (wi == 16) { ++blocks; b2b_compress(h, mbuf, 16, blocks * 128ull, false); wi = 0; }
                            mbuf[wi * WLT] = acc;
                            ++wi;
                            nacc -= 8;
                            acc = nacc ? (bytes >> (8 * (len - nacc))) : 0ull;
                        }
                    }
                }
                if (nacc) {
                    if (wi == 16) { ++blocks; b2b_compress(h, mbuf, 16, blocks * 128ull, false); wi = 0; }
                    mbuf[wi * WLT] = acc;
                    ++wi;
                }
                b2b_compress(h, mbuf, wi, total, true);
                WL_LAB_AT(nxt, 0, v) = bswap64(h[0]);  // big-endian keys: hex-string order == numeric order
                WL_LAB_AT(nxt, 1, v) = bswap64(h[1]);
            }
            // ---- Counter(node_labels.values()), sorted by label: repeatedly the smallest label above the previous one
            u64 p0 = 0, p1 = 0;
            bool have_prev = false;
            for (int done = 0; done < nn;) {
                u64 b0 = 0, b1 = 0;
                int c = 0;
                for (int v = 0; v < nn; ++v) {
                    const u64 k0 = WL_LAB_AT(nxt, 0, v), k1 = WL_LAB_AT(nxt, 1, v);
                    if (have_prev && !key_lt(p0, p1, k0, k1)) continue;
                    if (c == 0 || key_lt(k0, k1, b0, b1)) { b0 = k0; b1 = k1; c = 1; }
                    else if (k0 == b0 && k1 == b1) ++c;
                }
                if (entries++) f_put((u64)',' | ((u64)' ' << 8), 2);
                f_put((u64)'(' | ((u64)'\'' << 8), 2);
                f_put(hex8((u32)(b0 >> 32)), 8); f_put(hex8((u32)b0), 8);
                f_put(hex8((u32)(b1 >> 32)), 8); f_put(hex8((u32)b1), 8);
                f_put((u64)'\'' | ((u64)',' << 8) | ((u64)' ' << 16), 3);
                if (c >= 10) f_put((u64)('0' + c / 10) | ((u64)('0' + c % 10) << 8), 2);
                else f_put((u64)('0' + c), 1);
                f_put(')', 1);
                p0 = b0; p1 = b1; have_prev = true;
                done += c;
            }
            f_drain();  // ")" (or ",)") still follows: every complete block can go
            cur = nxt;
        }
        if (entries == 1) f_put(',', 1);
        f_put(')', 1);
        {
            const u64 total = f_blocks * 128ull + (u64)f_wi * 8ull + (u64)f_nacc;
            if (f_nacc) fs[f_wi++] = f_acc;
            if (f_wi > 16) {  // the two closing characters completed a block and spilled over: that block is not the last
#pragma unroll
                for (int i = 0; i < 16; ++i) mbuf[i * WLT] = fs[i];
                ++f_blocks;
                b2b_compress(fh, mbuf, 16, f_blocks * 128ull, false);
                for (int i = 16; i < f_wi; ++i) fs[i - 16] = fs[i];
                f_wi -= 16;
            }
            for (int i = 0; i < f_wi; ++i) mbuf[i * WLT] = fs[i];
            b2b_compress(fh, mbuf, f_wi, total, true);
        }
        digest_out[2 * s] = fh[0];
        digest_out[2 * s + 1] = fh[1];
    }
#undef WL_LAB_AT
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
    int64_t blocks = (B + WLT - 1) / WLT;
    const int64_t cap = (int64_t)sms * 8;
    if (blocks > cap) blocks = cap;
    const size_t smem = sizeof(u64) * WLT * (size_t)(4 * k + 16);  // k = 8: 48 KB, k = 16: 80 KB
    UGS_CUDA_CHECK(cudaFuncSetAttribute(k_wl_hash, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_wl_hash<<<(unsigned)blocks, WLT, smem, (cudaStream_t)stream_>>>(nodes, B, k, edge_index_s, edge_index_s + total, edge_ptr,
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
