/*
 * ugs_b200.h -- C ABI of libugs_b200.so: the B200 (sm_100a) uniform k-graphlet sampler that replaces
 * SS-GNN's `ugs_sampler` hot path (reference: src/samplers/ugs_sampler, v2.2.1).
 *
 * Plain pointers and sizes only: no torch types. All data pointers are DEVICE pointers unless a
 * parameter is named h_*. Every entry point returns 0 on success or a negative UGS_E* code; the
 * message is available from ugs_last_error() (thread-local). Nothing here synchronises the stream
 * except ugs_batch_info(), which copies 64 bytes back to the host.
 *
 * Call sequence for one batch (what ugs_sampler.sample_batch does,
 * reference ugs_sampler_batch_extension.cpp:77-300):
 *
 *   ugs_workspace_layout(E, G, NT, &L)                         host only, sizes the workspace
 *   ugs_segment_batch(...)        A.1  columns -> graphs       replaces slice_and_renumber (:41-75)
 *   ugs_batch_info(...)                host reads {unsorted, n_general, max_n_general, ...}
 *                                      (or ugs_batch_info_begin / ugs_preproc_batch_parts(SMALL) / ugs_batch_info_end:
 *                                      the round trip then overlaps the bitmask-path preprocessing)
 *   [ugs_sort_columns(...) + ugs_segment_batch(perm)]   only when the columns are not graph-contiguous
 *   ugs_preproc_batch(...)        A.2-A.5  CSR, order, suffix degrees, weights
 *                                      replaces create_preproc (preproc.cpp:262-284) per graph
 *   ugs_sample_batch(...)         A.5-A.6  roots + growth + per-sample edge counts + scan
 *                                      replaces sample() (sampler.cpp:91-181) per graph
 *   <host reads *d_total, allocates edge_index[2,Ebar], edge_src[Ebar]>
 *   ugs_emit_edges[_packed](...)  A.7  induced-edge extraction + compaction
 *                                      replaces sampler.cpp:205-280 and the batch driver's
 *                                      post-processing (ugs_sampler_batch_extension.cpp:188-231)
 *
 *   [ugs_deliver_outputs]         CPU int64 tensors of the reference ABI: part by DMA, part packed + widened
 *   ugs_sample_batch_fused        all of the above in one call for DataLoader-sized batches
 *
 * The other reference samplers (ugs_eps_*, ugs_apx_*, ugs_rwr_*, ugs_enum_*) share the workspace built by the first
 * three calls; ugs_gather_rows / ugs_encoder_offsets, ugs_presample_* and ugs_wl_* cover the callers either side of the
 * sampler (SURVEY.md section 8 f). "A.n" refers to SURVEY.md Appendix A (exact semantics).
 */
#ifndef UGS_B200_H
#define UGS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UGS_B200_ABI_VERSION 3

/* error codes */
#define UGS_OK 0
#define UGS_EINVAL (-1)   /* bad argument (message says which) */
#define UGS_ECUDA (-2)    /* CUDA runtime error */
#define UGS_ELIMIT (-3)   /* input exceeds an implementation limit (k > 32, graph too large, ...) */

/* edge modes: sample_batch mode "sample"/"graph"/"global"  ==  sample() edge_mode "local"/"flat"/"global"
 * (reference ugs_sampler_batch_extension.cpp:170-174, sampler.cpp:258-280) */
#define UGS_MODE_SAMPLE 0
#define UGS_MODE_GRAPH 1
#define UGS_MODE_GLOBAL 2
/* not a reference mode: entries (s*k + j, s*k + l) with s = the batch-wide sample row. This is what the consumer of the
 * sampler builds from mode "sample" output: edge_index_t + repeat_interleave(arange(B), counts) * k
 * (gps/models/ss_gnn.py:461-462, ss_gnn_wl.py:283-287), produced here by the emission pass itself. */
#define UGS_MODE_ENCODER 3

/* sampler variants */
#define UGS_VARIANT_REF 0      /* the reference law: root ~ b/Z, one rand_grow, no rejection (sampler.cpp:158-181) */
#define UGS_VARIANT_UNIFORM 1  /* exact-uniform accept/reject on top of the same root/growth laws */

#define UGS_MAX_K 32           /* general path; the bitmask fast path takes k <= 16 */
#define UGS_SMALL_MAX_N 64     /* graphs with <= 64 nodes take the one-thread-per-sample bitmask path */
#define UGS_SMALL_MAX_COLS 384 /* ... if they also have <= 384 edge columns */
#define UGS_LARGE_N 4096       /* general-path graphs above this are preprocessed by the whole GPU, one at a time */

/* Byte offsets of every array inside the caller-allocated preprocessing workspace. The arrays are the
 * device-resident equivalent of the reference's `struct Preproc` (sampler.hpp:81-93), stored once for the
 * whole batch and indexed by batch-global node id (minus node_base) / CSR position. */
typedef struct ugs_layout {
    size_t total_bytes;
    size_t cstart;     /* int64  [G+2]   column range of graph g: [cstart[g], cstart[g+1]) */
    size_t ginfo;      /* ugs_ginfo [G]                                                    */
    size_t glist;      /* int32  [G]     ids of the graphs on the general (warp-per-sample) path */
    size_t counters;   /* int64  [64]    see UGS_CTR_*                                      */
    size_t colkey;     /* int32  [E]     graph id of each column (G = belongs to no graph)  */
    size_t rowptr;     /* uint32 [NT+G]  graph g: n+1 entries at (lo-node_base)+g, relative to 2*cstart[g] */
    size_t order;      /* int32  [NT]    vertex (local id) at order position vi   (Preproc::order)      */
    size_t index_of;   /* int32  [NT]    position of local vertex v                (Preproc::index_of)   */
    size_t suffix_deg; /* int32  [NT]    by position                               (Preproc::suffix_deg) */
    size_t bucket_b;   /* double [NT]    by position                               (Preproc::bucket_b)   */
    size_t cum;        /* double [NT]    root-draw CDF by position (level 0: running sum of bucket_b)    */
    size_t adj;        /* uint64 [NT]    adjacency bitmask (bitmask path only)                           */
    size_t suf;        /* uint64 [NT]    by position: {v : index_of[v] >= vi} (bitmask path only)        */
    size_t csr_idx;    /* int32  [2E]    neighbour local id                        (Preproc::indices)    */
    size_t csr_col;    /* int32  [2E]    batch-global source column                (edge_col_of_csr_pos -> g_edge2global) */
    size_t gaux;       /* double [G]     per-graph acceptance constant C of UGS_VARIANT_UNIFORM                  */
    size_t llist;      /* int32  [G]     ids of the large general-path graphs (> UGS_LARGE_N nodes)              */
    size_t lscratch;   /* double [4096]  scratch of the cooperative large-graph kernel                           */
} ugs_layout;

typedef struct ugs_ginfo {
    double Z;            /* sum of bucket_b in position order (Preproc::Z)                           */
    int32_t nnz;         /* CSR entries stored (get_preproc_info num_edges_stored)                   */
    int32_t level;       /* relaxation level 0/1/2 (sampler.cpp:125-146); -1 if the graph is skipped */
    int32_t n_eligible;  /* #root positions eligible at that level; level 0 = bucket_count_nonzero   */
    int32_t last_viable; /* largest position with bucket_b > 0 (root-draw guard)                     */
    int32_t mult;        /* bitmask path: >0 = every adjacent pair appears exactly `mult` times per row and there are
                            no self loops, 0 = irregular; general path: -1 = no self-loop column, 0 = some */
    int32_t cls;         /* 0 = skipped (n <= 0 or n < k), 1 = bitmask path, 2 = general path        */
} ugs_ginfo;

/* counters[] slots */
#define UGS_CTR_UNSORTED 0      /* bit0: columns not grouped by graph (sort them: perm); bit1: ptr decreasing;
                                   bit2/bit3: the fast segmentation does not apply -> redo with fast = 0;
                                   bit4: ugs_sample_batch was called with k > 16 on a workspace whose bitmask-path
                                   graphs were preprocessed for a smaller k (nothing was sampled for them)   */
#define UGS_CTR_N_GENERAL 1
#define UGS_CTR_MAX_N_GENERAL 2
#define UGS_CTR_N_SKIPPED 3
#define UGS_CTR_TOTAL_EDGES 4   /* written by ugs_sample_batch                                          */
#define UGS_CTR_INCOMPLETE 5    /* rows with a -1                                                        */
#define UGS_CTR_RELAXED 6       /* graphs sampled at relaxation level > 0                                */
#define UGS_CTR_TRIALS 7        /* growth attempts, written by the rejection variants / walk steps (rwr). VARIANT_REF
                                   grows every row exactly once and leaves it 0: trials == sampled rows there */
#define UGS_CTR_SCAN_TICKET 8
#define UGS_CTR_N_LARGE 9        /* general-path graphs with more than UGS_LARGE_N nodes                   */
#define UGS_CTR_WORK 14          /* work-claim cursor of the persistent general-path sampling kernel          */
#define UGS_CTR_EMIT_OVERFLOW 15 /* samples ugs_eps_emit_edges could not emit (more than 1024 induced columns on the
                                    general path, or counts that do not match edge_ptr): their slice of the output is
                                    left untouched -- the caller must treat a non-zero value as an error      */
#define UGS_CTR_N_NOT_MEDIUM 48  /* general-path graphs above the shared-memory kernels' size (more than 1024 nodes or
                                    3072 columns): 0 lets the global-memory general-path kernels return at once */
#define UGS_CTR_COUNT 64         /* int64 words in counters[]                                             */
#define UGS_CTR_EARLY_TOTAL 16   /* [16, 48): partial sums of the per-sample entry counts, added up by the sampling
                                    kernels themselves (32 addresses to spread the atomics); their sum is the total
                                    number of emitted entries, known BEFORE the scan kernel has run          */
#define UGS_CTR_EARLY_SLOTS 32

int ugs_abi_version(void);
/* Number of CUDA kernels this library has launched in this process (monotonic; for measurement). */
uint64_t ugs_launch_count(void);
const char* ugs_last_error(void);

/* Sizes and offsets of the workspace for a batch with E columns, G graphs and NT = ptr[G]-ptr[0] nodes. */
int ugs_workspace_layout(int64_t E, int64_t G, int64_t NT, ugs_layout* out);

/* A.1 -- map every column to its graph, find per-graph column ranges, classify graphs.
 * edge_index: [2,E] int64 row-major (row 1 at edge_index + E). ptr: [G+1] int64, non-decreasing.
 * perm: NULL, or int64[E] = a stable ordering of the columns by graph id produced after a first call
 * reported UGS_CTR_UNSORTED (columns are then read through perm).
 * fast: 1 = assume graph-contiguous columns (every PyG batch): one binary search per graph instead of one per
 *       column. The assumption is verified (counters[UGS_CTR_UNSORTED] bit 2 right away, bit 3 by the
 *       preprocessing kernels); when a bit is set the caller must redo the batch with fast = 0.
 * skip_lt_k: 1 = graphs with fewer than k nodes produce all -1 rows (sample_batch, :132-143);
 *            0 = they are preprocessed and sampled at a relaxation level (handle API, sampler.cpp:125-146). */
int ugs_segment_batch(const int64_t* edge_index, int64_t E, const int64_t* ptr, int64_t G, int k,
                      int skip_lt_k, int fast, const int64_t* perm, void* workspace, const ugs_layout* L, void* stream);

/* Slow path of A.1 (columns not grouped by graph; never a PyG batch): perm_out[E] = stable ordering of the columns
 * by the graph ids ugs_segment_batch(fast = 0) left in the workspace (LSD radix sort). */
size_t ugs_sort_scratch_bytes(int64_t E);
int ugs_sort_columns(int64_t E, int64_t G, void* workspace, const ugs_layout* L, void* scratch, size_t scratch_bytes,
                     int64_t* perm_out, void* stream);

/* Copies counters[0..15] to h_out (16 x int64; synchronises the stream). */
int ugs_batch_info(const void* workspace, const ugs_layout* L, int64_t* h_out, void* stream);
/* The same without stalling `stream`: _begin snapshots the counters as of the work queued on `stream` so far (the copy
 * runs on a library-owned side stream; h_out must be pinned host memory of UGS_CTR_COUNT int64), _end blocks the host until the snapshot has
 * arrived. Work queued on `stream` between the two calls overlaps the round trip: the drop-in launches the
 * bitmask-path preprocessing there (UGS_PREPROC_SMALL) while it waits for the segmentation flags / class counts. */
int ugs_batch_info_begin(const void* workspace, const ugs_layout* L, int64_t* h_out, void* stream);
int ugs_batch_info_end(void);

/* A.2-A.5 -- per-graph CSR (multigraph, column order), (degree,id) order, suffix degrees, reachability,
 * bucket weights, root CDF. n_general / max_n_general come from ugs_batch_info. */
int ugs_preproc_batch(const int64_t* edge_index, int64_t E, const int64_t* ptr, int64_t G, int64_t node_base,
                      int k, const int64_t* perm, int64_t n_general, int64_t n_large, void* workspace,
                      const ugs_layout* L, void* stream);
/* The same in two parts. UGS_PREPROC_SMALL (warp-per-graph kernel over the bitmask-path graphs) needs nothing from the
 * host and may be launched before the counters have been read (pass n_general = n_large = 0); UGS_PREPROC_GENERAL
 * (CTA-per-graph / whole-GPU kernels) needs n_general and n_large. */
#define UGS_PREPROC_SMALL 1
#define UGS_PREPROC_GENERAL 2
#define UGS_PREPROC_ALL 3
int ugs_preproc_batch_parts(const int64_t* edge_index, int64_t E, const int64_t* ptr, int64_t G, int64_t node_base,
                            int k, const int64_t* perm, int64_t n_general, int64_t n_large, int parts,
                            void* workspace, const ugs_layout* L, void* stream);

/* ugs_sample_batch with an early read-back: when h_info (pinned host memory, UGS_CTR_COUNT int64) is not NULL, a
 * snapshot of counters[] taken after the sampling kernels and BEFORE the scan travels to h_info over the library's side
 * stream; ugs_batch_info_end() waits for it. The host thereby learns the number of emitted entries
 * (sum of h_info[UGS_CTR_EARLY_TOTAL .. +UGS_CTR_EARLY_SLOTS)), the incomplete / relaxed counts and the late
 * segmentation flag while the scan kernel is still running, allocates edge_index / edge_src and queues the emission
 * pass behind the scan: the data-dependent allocation no longer costs GPU idle time. UGS_VARIANT_REF only. */
int ugs_sample_batch_early(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, uint64_t seed,
                           uint64_t graph_index_base, uint64_t sample_index_base, int seed_mode, int variant,
                           int64_t n_general, int64_t max_n_general, void* workspace, const ugs_layout* L,
                           void* scratch, size_t scratch_bytes, int64_t* nodes, int64_t* edge_ptr,
                           int64_t* sample_ptr, int64_t* h_info, void* stream);

/* Scratch bytes ugs_sample_batch needs for G*m samples (scan status words + one packed row of 8 bytes per sample). */
size_t ugs_sample_scratch_bytes(int64_t G, int64_t m);

/* A.5-A.6 + counts. Writes nodes[G*m,k] (batch-global ids, -1 pad; lo is added for every mode exactly as the
 * batch driver does, :188-196), edge_ptr[G*m+1] (exclusive running totals), sample_ptr[G+1] (= g*m, :130),
 * counters[UGS_CTR_TOTAL_EDGES..]. Randomness: Philox4x32-10 keyed (seed; graph_index_base+g, sample_index_base+i):
 * a rank that holds graphs [g0,g1) of a batch passes graph_index_base = g0; a rank that draws samples [i0,i1) of
 * every graph (one huge graph replicated on all ranks) passes m = i1-i0 and sample_index_base = i0.
 * Either way the rows are the ones a single call over the whole job produces.
 * seed_mode 0: as above. seed_mode 1: key = seed + graph_index_base + g and graph counter 0, i.e. graph g gets the rows
 * a single-graph call with seed + (graph_index_base + g) produces -- the reference's presampling loop
 * (gps/experiment.py:405-417: one call per dataset graph with seed = cfg.seed + index) as ONE batch call. */
int ugs_sample_batch(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, uint64_t seed,
                     uint64_t graph_index_base, uint64_t sample_index_base, int seed_mode, int variant,
                     int64_t n_general, int64_t max_n_general,
                     void* workspace, const ugs_layout* L, void* scratch, size_t scratch_bytes,
                     int64_t* nodes, int64_t* edge_ptr, int64_t* sample_ptr, void* stream);

/* A.7 -- induced-edge emission for the node rows in `nodes` (normally the ones ugs_sample_batch wrote; any
 * rows whose per-sample counts are in edge_ptr work, which is how the extraction stage is pinned against the
 * reference's own samples). edge_index_out: [2,total] (row 1 at +total), edge_src_out: [total].
 * use_perm: 1 if the workspace was built with a column permutation (ugs_segment_batch perm != NULL). */
int ugs_emit_edges(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int mode,
                   int use_perm, int64_t n_general, const void* workspace, const ugs_layout* L, const int64_t* nodes,
                   const int64_t* edge_ptr, int64_t total, int64_t* edge_index_out, int64_t* edge_src_out,
                   void* stream);

/* The same when the rows are the ones ugs_sample_batch has just written for the same (G, m, k): sample_scratch is the
 * scratch buffer of that call. For k <= 8 it holds every row again as eight local-id bytes, so the emission pass reads
 * 8 instead of 8k bytes per sample (0.47 GB less DRAM traffic per ZINC-shaped step). NULL = read `nodes`. */
int ugs_emit_edges_packed(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int mode, int use_perm,
                          int64_t n_general, const void* workspace, const ugs_layout* L, const int64_t* nodes,
                          const int64_t* edge_ptr, int64_t total, int64_t* edge_index_out, int64_t* edge_src_out,
                          const void* sample_scratch, size_t sample_scratch_bytes, void* stream);

/* Per-sample induced-edge counts for caller-provided node rows (writes edge_ptr like ugs_sample_batch). */
int ugs_count_edges(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int64_t n_general,
                    void* workspace, const ugs_layout* L, void* scratch, size_t scratch_bytes,
                    const int64_t* nodes, int64_t* edge_ptr, void* stream);

/* ---- small batches: the whole pipeline in one call with one host round trip (csrc/ugs_fused.cu) -----------------
 * What ugs_sampler.sample_batch costs for a DataLoader mini-batch (gps/experiment.py:882-883) is launch latency and
 * host round trips, not kernel time. ugs_sample_batch_fused queues segment -> preprocess -> sample -> scan -> emit
 * back to back (UGS_VARIANT_REF, graph-contiguous columns): the general-path kernels get worst-case grids and read
 * their item counts from counters[] on the device, and the emission pass writes into ei_u / ei_v (= ei_u + cap) /
 * ei_src of cap >= ugs_fused_capacity(E, m) = 2*E*m entries (a sample emits every CSR entry at most once). The only
 * thing the host waits for is counters[] in h_info (pinned host memory, UGS_CTR_COUNT int64; word UGS_CTR_COUNT-1 is
 * used as the arrival flag): the scan kernel writes them through the mapped address as soon as the total is known and
 * the call returns when they have arrived -- the emission kernels may still be running, the outputs are valid in STREAM
 * order (the next kernel or copy on `stream` sees them; a host read needs a stream synchronisation first). When h_info
 * is not device-mapped the counters are copied back after the emission and the stream is synchronised instead.
 * h_info[UGS_CTR_TOTAL_EDGES] = Ebar; a non-zero h_info[UGS_CTR_UNSORTED] means the batch needs the staged path and no
 * output is valid. ugs_fused_finish narrows the bound-sized arrays into edge_index_out[2,Ebar] / edge_src_out[Ebar].
 * Outputs are bit-identical to the staged sequence.
 * flags:
 *   UGS_FUSED_ASSUME_NO_GENERAL leaves the general-path kernels out of the queue (five launches that would return at
 *     once -- ~3 us each in a 0.1 ms chain -- when the batch has no graph above 64 nodes / 384 columns, as every MUTAG /
 *     QM9 / ZINC batch): the caller must then check h_info[UGS_CTR_N_GENERAL]; non-zero voids the outputs and the call
 *     has to be repeated without the flag.
 *   UGS_FUSED_COMPACT makes the emission kernels read Ebar on the device and write the three rows behind one another:
 *     ei_u[0, Ebar) | [Ebar, 2 Ebar) | [2 Ebar, 3 Ebar) of ONE array of 3*cap entries (ei_v = ei_u + cap and
 *     ei_src = ei_v + cap on entry). The first 3*Ebar entries then ARE edge_index[2,Ebar] and edge_src[Ebar]; no
 *     ugs_fused_finish pass. For callers who can live with results that are views of the bound-sized allocation. */
#define UGS_FUSED_ASSUME_NO_GENERAL 1
#define UGS_FUSED_COMPACT 2
int64_t ugs_fused_capacity(int64_t E, int64_t m);
int ugs_sample_batch_fused(const int64_t* edge_index, int64_t E, const int64_t* ptr, int64_t G, int64_t node_base,
                           int64_t total_nodes, int64_t m, int k, int mode, uint64_t seed, uint64_t graph_index_base,
                           uint64_t sample_index_base, int seed_mode, void* workspace, const ugs_layout* L,
                           void* scratch, size_t scratch_bytes, int64_t* nodes, int64_t* edge_ptr, int64_t* sample_ptr,
                           int64_t* ei_u, int64_t* ei_v, int64_t* ei_src, int64_t cap, int64_t* h_info, int flags,
                           void* stream);
int ugs_fused_finish(const int64_t* ei_u, const int64_t* ei_v, const int64_t* ei_src, int64_t total,
                     int64_t* edge_index_out, int64_t* edge_src_out, void* stream);

/* ---- exact-uniform accept/reject variant (not in the reference, see oracle/uni_oracle.c) ---------
 * ugs_preproc_uniform: call after ugs_preproc_batch. Replaces order / suffix_deg / suf / bucket_b / cum / Z by
 * the degree-dominating-order tables of the uniform mode and writes the acceptance constant C per graph.
 * k <= 8. Bitmask-path graphs (<= 64 nodes) get their tables from a warp each; general-path graphs (up to 4096 nodes;
 * n_general / max_n_general from ugs_batch_info) from a CTA each, and their samples from one warp per sample whose
 * p(S|r) comes from a subset-sum transform over the members' adjacency masks (csrc/ugs_uniform_general.cu).
 * UGS_ELIMIT for k > 8 or a graph above 4096 nodes.
 * Then ugs_sample_batch(variant = UGS_VARIANT_UNIFORM) and ugs_emit_edges as usual. */
#define UGS_UNIFORM_MAX_TRIALS 4096
int ugs_preproc_uniform(const int64_t* ptr, int64_t G, int64_t node_base, int k, int64_t n_general,
                        int64_t max_n_general, void* workspace, const ugs_layout* L, void* stream);

/* ---- epsilon-uniform variant (reference src/samplers/epsilon_uniform_sampler) --------------------
 * Same workspace as the UGS path (ugs_segment_batch + ugs_preproc_batch); only the multigraph CSR is used.
 * ugs_eps_sample_batch: random frontier growth + rejection min(1, eps/(w+eps)), <= max(10, 10/eps) attempts
 * (epsilon_uniform_sampler.cpp:18-84,225-257); writes SORTED node rows (:256; -1 rows when every attempt failed),
 * per-sample column counts scanned into edge_ptr, sample_ptr = g*m. counters[UGS_CTR_TRIALS] = growth attempts,
 * counters[UGS_CTR_INCOMPLETE] = failed rows.
 * ugs_eps_emit_edges: every batch column with both endpoints in the sample, in column order, src = the column
 * (:267-290); mode "sample" = positions in the sorted row, other modes = batch-global node ids (:279-285). */
int ugs_eps_sample_batch(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, uint64_t seed,
                         uint64_t graph_index_base, double epsilon, int64_t n_general, void* workspace,
                         const ugs_layout* L, void* scratch, size_t scratch_bytes, int64_t* nodes,
                         int64_t* edge_ptr, int64_t* sample_ptr, void* stream);
int ugs_eps_emit_edges(const int64_t* edge_index, int64_t E, const int64_t* ptr, int64_t G, int64_t node_base,
                       int64_t m, int k, int mode, const int64_t* perm, int64_t n_general, const void* workspace,
                       const ugs_layout* L, const int64_t* nodes, const int64_t* edge_ptr, int64_t total,
                       int64_t* edge_index_out, int64_t* edge_src_out, void* stream);

/* ---- APX-UGS (reference src/samplers/apx_ugs_sampler/src/apx_ugs_sampler.cpp) -----------------
 * One graph per call, like the reference: the columns [e0, e1) of edge_index (the reference reads `ptr` as an EDGE
 * range, :15-33), n = max vertex id + 1 <= 1024, 2 <= k <= 8.
 *   ugs_apx_prepare  Graph::Graph + apx_dd (:10-169). Syncs; h_info = {n, adjacency entries, Z > 0}.
 *   ugs_apx_trials   one round: trials [tbase, tbase+per_sample) of every sample, one warp per (sample, trial);
 *                    h_best[s] = first accepted trial so far (>= 10^6: none). Syncs. The caller loops rounds until
 *                    every sample has best[s] < tbase + per_sample or tbase reaches max_trials (10^6, :411).
 *   ugs_apx_write    samples[k, n_ok] (row-major, :508-515); slot[s] = output column of sample s or -1 (dropped). */
size_t ugs_apx_workspace_bytes(int64_t n_columns, int64_t m);
int ugs_apx_prepare(const int64_t* edge_index, int64_t E, int64_t e0, int64_t e1, int64_t m, int k, uint64_t seed,
                    double epsilon, void* workspace, size_t ws_bytes, int64_t* h_info, void* stream);
int ugs_apx_trials(int n, int64_t n_columns, int64_t m, int k, uint64_t seed, double epsilon, int tbase,
                   int per_sample, int max_trials, void* workspace, int32_t* h_best, void* stream);
int ugs_apx_write(int n, int64_t n_columns, int64_t m, int k, uint64_t seed, double epsilon, const int32_t* slot,
                  int n_ok, void* workspace, int64_t* samples_out, void* stream);
int ugs_apx_dump(int n, int64_t nnz, int64_t n_columns, int64_t m, void* workspace, int32_t* rowptr, int32_t* adj,
                 int32_t* order, int32_t* position, double* est, void* stream);

/* ---- Random-Walk-with-Restart sampler (reference src/samplers/rwr_sampler/src/rwr_sampler.cpp:72-295) ------------
 * Same workspace as the UGS path (ugs_segment_batch + ugs_preproc_batch; only the multigraph CSR is used, which is the
 * reference's adjacency with duplicates, :31-69). ugs_rwr_sample_batch: per sample a uniform seed vertex, then at most
 * n*k*10 walk steps (restart to the seed with probability p_restart or from a vertex without neighbours, else a uniform
 * row entry), the first k distinct vertices in visiting order; all -1 rows when the cap is hit or n < k (:139-204).
 * Writes nodes, scanned per-sample entry counts (edge_ptr), sample_ptr = g*m; counters[UGS_CTR_TRIALS] = walk steps,
 * counters[UGS_CTR_INCOMPLETE] = failed rows. ugs_rwr_emit_edges: the UGS extraction (rows of the sample in visiting
 * order, entries in row order) with edge_src = -1 (:217-256); mode "sample" = positions in the row, any other mode =
 * batch-global node ids. */
int ugs_rwr_sample_batch(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, uint64_t seed,
                         uint64_t graph_index_base, double p_restart, void* workspace, const ugs_layout* L,
                         void* scratch, size_t scratch_bytes, int64_t* nodes, int64_t* edge_ptr, int64_t* sample_ptr,
                         void* stream);
int ugs_rwr_emit_edges(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int mode, int use_perm,
                       int64_t n_general, const void* workspace, const ugs_layout* L, const int64_t* nodes,
                       const int64_t* edge_ptr, int64_t total, int64_t* edge_index_out, int64_t* edge_src_out,
                       void* stream);

/* ---- exactly uniform sampler by enumeration (reference src/samplers/uniform_sampler/src/uniform_sampler.cpp:86-282) ---
 * Same workspace as the UGS path (adjacency bitmasks + CSR); n_general / max_n_general come from ugs_batch_info.
 * Graphs with <= 64 vertices are enumerated on 64-bit masks (k <= 16), larger ones (<= 2048 vertices, k <= 8) on
 * multi-word bitsets; the reference itself tests all C(n,k) combinations per graph (:47-81).
 *   ugs_enum_count         every connected k-subset counted once (ESU recursion); per-root and per-graph counts go to
 *                          enum_scratch, their total to counters[UGS_CTR_TOTAL_EDGES] (read it with ugs_batch_info)
 *   ugs_enum_fill          list[total * ugs_enum_record_words()] = the subsets, graph by graph: one 64-bit vertex mask
 *                          per subset, or (batches with graphs above 64 vertices) 16-byte records holding the mask or
 *                          k 16-bit vertex ids
 *   ugs_enum_sample_batch  row = ascending vertices of list[graph][mulhi64(Philox word64, count)] (:196-210); all -1
 *                          rows for graphs without a connected k-subset (:183-191); edge_ptr = scanned per-sample column
 *                          counts; sample_ptr = g*m. Edges: ugs_eps_emit_edges (same column-scan semantics, :216-245). */
size_t ugs_enum_scratch_bytes(int64_t G, int64_t NT, int64_t n_general, int64_t max_n_general);
int ugs_enum_record_words(int64_t n_general, int64_t max_n_general);
int ugs_enum_count(const int64_t* ptr, int64_t G, int64_t NT, int64_t node_base, int k, int64_t n_general,
                   int64_t max_n_general, void* workspace, const ugs_layout* L, void* enum_scratch,
                   size_t enum_scratch_bytes, void* stream);
int ugs_enum_fill(const int64_t* ptr, int64_t G, int64_t NT, int64_t node_base, int k, int64_t n_general,
                  int64_t max_n_general, void* workspace, const ugs_layout* L, void* enum_scratch, uint64_t* list,
                  int64_t total, void* stream);
int ugs_enum_sample_batch(const int64_t* ptr, int64_t G, int64_t NT, int64_t node_base, int64_t m, int k, uint64_t seed,
                          uint64_t graph_index_base, int64_t n_general, int64_t max_n_general, void* workspace,
                          const ugs_layout* L, void* enum_scratch, const uint64_t* list, void* scratch,
                          size_t scratch_bytes, int64_t* nodes, int64_t* edge_ptr, int64_t* sample_ptr, void* stream);

/* ---- presample store: re-stitching cached samples into a mini-batch ---------------------------
 * Replaces Experiment._setup_presampling / _load_from_presample_cache (gps/experiment.py:379-440,895-995). The store
 * (s_*) is the output of one ugs_sample_batch(seed_mode = 1, mode "sample") call over a whole dataset split of Gd
 * graphs, kept in HBM, plus that batch's ptr (s_ptr) and per-graph first columns (s_cstart = workspace cstart).
 *   ugs_presample_plan  per-slot entry counts and the mini-batch's per-graph column offsets (:926-932), scanned into
 *                       scratch; *h_total = entries of the stitched mini-batch (synchronises the stream).
 *                       graph_idx[Gb] = dataset indices; ptr[Gb+1], edge_index_row0[E] = the mini-batch itself.
 *   ugs_presample_load  nodes (cached local ids + ptr[b]; keep_padding = 0 adds ptr[b] to the -1 padding too, exactly
 *                       like :957-958), unchanged local edge_index, edge_src re-based to the mini-batch's columns
 *                       (:966), edge_ptr / sample_ptr running totals (:973-981). */
size_t ugs_presample_scratch_bytes(int64_t Gb);
int ugs_presample_plan(const int64_t* s_edge_ptr, int64_t Gd, int64_t m, const int64_t* graph_idx, int64_t Gb,
                       const int64_t* ptr, const int64_t* edge_index_row0, int64_t E, void* scratch,
                       size_t scratch_bytes, int64_t* h_total, void* stream);
int ugs_presample_load(const int64_t* s_nodes, const int64_t* s_edge_index, int64_t s_total, const int64_t* s_edge_ptr,
                       const int64_t* s_edge_src, const int64_t* s_ptr, const int64_t* s_cstart, int64_t Gd, int64_t m,
                       int k, const int64_t* graph_idx, int64_t Gb, const int64_t* ptr, const void* scratch,
                       int64_t total, int keep_padding, int64_t* nodes_out, int64_t* edge_index_out,
                       int64_t* edge_ptr_out, int64_t* sample_ptr_out, int64_t* edge_src_out, void* stream);

/* ---- Weisfeiler-Lehman graphlet classification ----------------------------------------------
 * Replaces compute_wl_hash + extract_subgraph_from_batch (gps/utils/wl_vocab.py:21-108) and the per-sample loop
 * SubgraphSamplingGNNWithWL._compute_wl_ids (gps/models/ss_gnn_wl.py:210-247).
 *   ugs_wl_node_attr  use_node_features = True only: attr32_out[v] = first four bytes of md5(row v of x) for every node
 *                  of the batch (the reference's attribute is md5(features.tobytes()).hexdigest()[:8], wl_vocab.py:45-49).
 *   ugs_wl_hash    digest_out[B][2] = the 16-byte BLAKE2b digest whose hex string is exactly what
 *                  nx.weisfeiler_lehman_graph_hash(G, node_attr = str(degree), iterations) returns for the sample's
 *                  nx.Graph (memory order = hex order); num_nodes_out[B] = #(row entries >= 0), 0 = "unknown id".
 *                  node_attr32 (NULL = degree attributes): per batch node, from ugs_wl_node_attr; n_attr = its length.
 *                  edge_index_s: the sampler's mode-"sample" output [2,total] (row 1 at +total). k <= 16.
 *   ugs_wl_lookup  ids_out[s] = vocab_ids[i] where vocab_keys[i] (two big-endian uint64 per hash, sorted ascending =
 *                  hex-string order) equals digest s; unknown_id when absent or num_nodes == 0 (hash_to_id, :206-217). */
int ugs_wl_node_attr(const void* x, int64_t n_rows, int64_t row_bytes, uint32_t* attr32_out, void* stream);
int ugs_wl_hash(const int64_t* nodes, int64_t B, int k, const int64_t* edge_index_s, int64_t total,
                const int64_t* edge_ptr, int iterations, const uint32_t* node_attr32, int64_t n_attr,
                uint64_t* digest_out, int32_t* num_nodes_out, void* stream);
int ugs_wl_lookup(const uint64_t* digest, const int32_t* num_nodes, int64_t B, const uint64_t* vocab_keys,
                  const int64_t* vocab_ids, int64_t V, int64_t unknown_id, int64_t* ids_out, void* stream);

/* ---- host delivery of the reference's CPU int64 ABI (reference sampler.cpp:190, ugs_sampler_batch_extension.cpp:95) -
 * The reference returns CPU pinned int64 tensors; copying them as they are is PCIe-bound (6.9 GB per ZINC-shaped call).
 * ugs_deliver_outputs splits every array between two pipelines that run concurrently:
 *   DMA    the first dma_fraction of nodes / edge_index rows / edge_src (whole 4096-value blocks) is copied as int64 from
 *          the device tensors straight into the final host arrays, in slices issued between the wire chunks (no CPU work);
 *   widen  the rest is packed on the device into d_wire -- nodes and edge_src as 16-bit deltas against the minimum of
 *          their 4096-value block (int32 when some block spreads too far), edge_index as one byte per entry (pairs_u8 = 1:
 *          both rows < 16, i.e. mode "sample" with k <= 16) or two int32 --, copied to the pinned staging buffer h_wire
 *          in chunks, and n_threads host threads (a persistent pool + the caller) widen every chunk into the final arrays
 *          with non-temporal vector stores while the next chunks are in flight.
 * Blocks until everything has arrived. dma_fraction = 1 is the plain int64 copy, 0 packs everything (1/7 of the bytes
 * over PCIe, host-memory bound); the caller chooses by the host threads it owns. d_wire / h_wire: wire_bytes >=
 * ugs_deliver_wire_bytes(). h_edge_index is [2,total] (row 1 at +total). h_stats (may be NULL): {packed bytes copied, int64
 * bytes copied by DMA, 1 = 16-bit deltas / 0 = int32, chunks}. UGS_ELIMIT: a value fits neither format (the caller copies
 * the int64 tensors as they are). ugs_host_store_bandwidth: GB/s of non-temporal stores into dst[bytes] from n_threads
 * threads of the same pool (best of reps) -- the host-side roofline of the widen pipeline. ugs_host_isa: "avx512" |
 * "avx2" | "scalar", the widening code chosen for this CPU (UGS_B200_HOST_ISA overrides). */
size_t ugs_deliver_wire_bytes(int64_t B, int k, int64_t total, int pairs_u8);
int ugs_deliver_outputs(const int64_t* nodes, int64_t B, int k, const int64_t* edge_index, const int64_t* edge_src,
                        int64_t total, int pairs_u8, double dma_fraction, void* d_wire, void* h_wire, size_t wire_bytes,
                        int64_t* h_nodes, int64_t* h_edge_index, int64_t* h_edge_src, int n_threads, int64_t* h_stats,
                        void* stream);
/* The same delivery without an int64 copy of the edges in HBM (mode "sample", k <= 8, graph-contiguous columns, no
 * general-path graph -- the ZINC-shaped headline batch): ugs_emit_edges_wire is the emission pass of ugs_emit_edges_packed
 * writing the transfer format straight into d_wire (one byte j | l << 4 per entry; the source column as an 8-bit delta
 * against the first column of the graph that owns its 1024-entry block), 2 instead of 24 bytes per entry;
 * ugs_deliver_outputs_wire then packs only the node ids and runs the copy / widen pipeline (dma_fraction 0). Replaces the
 * same reference lines as ugs_emit_edges + ugs_deliver_outputs (sampler.cpp:205-280, :190;
 * ugs_sampler_batch_extension.cpp:95,257-290). wire_bytes >= ugs_deliver_wire_bytes(G * m, k, total, 1), same value in
 * both calls. n_general (the batch's count of general-path graphs) must be 0 (UGS_ELIMIT otherwise). UGS_ELIMIT from the delivery: a delta does not fit its 8 / 16 bits -- the caller emits the int64 arrays
 * (ugs_emit_edges_packed) and delivers them with ugs_deliver_outputs. */
int ugs_emit_edges_wire(const int64_t* ptr, int64_t G, int64_t node_base, int64_t m, int k, int64_t n_general,
                        const void* workspace, const ugs_layout* L, const int64_t* nodes, const int64_t* edge_ptr,
                        int64_t total, void* d_wire, size_t wire_bytes, const void* sample_scratch,
                        size_t sample_scratch_bytes, void* stream);
int ugs_deliver_outputs_wire(const int64_t* nodes, int64_t B, int k, int64_t total, void* d_wire, void* h_wire,
                             size_t wire_bytes, int64_t* h_nodes, int64_t* h_edge_index, int64_t* h_edge_src,
                             int n_threads, int64_t* h_stats, void* stream);
/* Host-only: copies bytes from src (pageable memory of the caller: the trainer's `.cpu()` tensors, experiment.py:877-878)
 * to dst (a pinned staging buffer) with n_threads threads of the delivery pool; the H2D copy then runs from pinned
 * memory at the link rate instead of being staged by the driver on one thread. */
int ugs_host_copy(void* dst, const void* src, size_t bytes, int n_threads);
double ugs_host_store_bandwidth(void* dst, size_t bytes, int n_threads, int reps);
const char* ugs_host_isa(void);
/* host-only test hook: widen n packed elements with the named code variant ("scalar" | "avx2" | "avx512"; -1 = this CPU
 * lacks it). kind 0: int32 -> int64, 1 / 2: low / high nibble of a byte -> int64, 3: 16-bit delta + base (0xFFFF -> -1) */
int ugs_host_widen(int kind, const void* src, int64_t* dst, size_t n, int64_t base, const char* isa);

/* ---- encoder inputs: the gather/offset step right after the sampler -----------------------
 * Reference: SubgraphSamplingGNNClassifier.encode_subgraphs, gps/models/ss_gnn.py:427-471 (same code in
 * ss_gnn_wl.py:249-306): global_x = x[nodes.clamp(0)] * (nodes >= 0), global_edge_attr = edge_attr[edge_src],
 * global_edge_index = edge_index_t + repeat_interleave(arange(B), counts) * k, batch = repeat_interleave(arange(B), k).
 *
 * ugs_gather_rows     out[i, :] = index[i] >= 0 ? src[index[i], :] : 0   for rows of row_bytes bytes (any dtype; the
 *                     copy is bit-exact). Rows with index >= n_src_rows are zero-filled and counted in *n_bad (device
 *                     int64, may be NULL; the reference raises IndexError for them).
 * ugs_encoder_offsets for already materialised mode-"sample" output: global_edge_index[2,total] (row 1 at +total) and
 *                     batch_out[B*k] = i / k. edge_index_s may be NULL (only batch_out is written) and batch_out may
 *                     be NULL. With ugs_emit_edges(mode = UGS_MODE_ENCODER) the first output needs no extra pass. */
int ugs_gather_rows(const void* src, int64_t n_src_rows, int64_t row_bytes, const int64_t* index, int64_t n_index,
                    void* out, int64_t* n_bad, void* stream);
int ugs_encoder_offsets(const int64_t* edge_index_s, const int64_t* edge_ptr, int64_t B, int k, int64_t total,
                        int64_t* global_edge_index_out, int64_t* batch_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UGS_B200_H */
