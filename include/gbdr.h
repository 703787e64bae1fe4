/*
 * gbdr.h — C ABI of the B200-native search path of gbnns_dim_red.
 *
 * This is the drop-in boundary (SURVEY.md §8 b1'): everything the reference's
 * search/search_function.h, search/final_test.cpp, search/prepare_graph.cpp and
 * wrap/c_support.cpp do on the hot path is reachable through these entry points
 * with plain pointers and sizes.  No C++ types, no torch types, no exceptions
 * cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success and a negative gbdr_status on error;
 *     gbdr_last_error() returns a thread-local message for the last failure.
 *   - "host" entry points take caller-owned HOST buffers and do the H2D / D2H
 *     copies themselves (this is what the reference-facing host code calls).
 *   - "_dev" entry points take DEVICE pointers on the index's device plus a
 *     cudaStream_t (passed as void*) and never synchronise; they are what a
 *     caller that already keeps data in HBM (bench `value` leg, torch tensors,
 *     the multi-GPU merge) uses.
 *   - a handle is thread-compatible: one host thread per handle at a time.
 *   - there is NO CPU fallback: if no sm_100-class device is present every
 *     compute entry point fails with GBDR_E_NO_DEVICE.
 *
 * Reference interface each entry point replaces is cited as file:line relative
 * to the reference repository root.
 */
#ifndef GBDR_H_
#define GBDR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GBDR_VERSION 100

typedef enum gbdr_status {
    GBDR_OK = 0,
    GBDR_E_INVALID = -1,    /* bad argument (null pointer, size mismatch, d%4 …)  */
    GBDR_E_NO_DEVICE = -2,  /* no CUDA device / wrong architecture                */
    GBDR_E_CUDA = -3,       /* a CUDA runtime call failed (message has details)   */
    GBDR_E_STATE = -4,      /* index is missing a component the call needs        */
    GBDR_E_CAPACITY = -5,   /* per-query on-chip state overflowed at every retry  */
    GBDR_E_NOMEM = -6
} gbdr_status;

#define GBDR_PAD_ID 0xFFFFFFFFu /* padding in adjacency rows and short result rows */

typedef struct gbdr_index gbdr_index; /* opaque, one per GPU */

/* ------------------------------------------------------------------ misc */
int gbdr_version(void);
const char *gbdr_last_error(void);
/* number of visible CUDA devices (0 and GBDR_OK when there is none) */
int gbdr_device_count(int *count);

/* ------------------------------------------------------------ index state */
/* One index object per device.  Holds db, db_low, graph, net in HBM. */
int gbdr_index_create(int device, gbdr_index **out);
int gbdr_index_destroy(gbdr_index *h);
/* A second handle on the SAME resident data (db, db_low, graphs, net are shared, not copied) with its
 * own CUDA stream and workspaces, so that two batches can be in flight on one GPU: the tail of one
 * batch's persistent search kernel overlaps the head of the next, and one batch's host<->device
 * copies overlap the other's kernels.  (The reference gets its concurrency from OpenMP threads over
 * the queries of one batch, search_function.h:152; this is the GPU-side equivalent across batches.)
 * Views are read-only (gbdr_index_set_* on a view fails); they follow later changes of the parent the
 * next time they are used, provided nothing is in flight on them then.  Destroy views before the
 * parent (gbdr_index_destroy of a parent with live views fails with GBDR_E_STATE). */
int gbdr_index_create_view(gbdr_index *parent, gbdr_index **out);

/* Original-dimension base vectors, row-major [n x d] float32.
 * Replaces `vector<float> db = loadXvecs<float>(…_base.fvecs, d, n)`
 * (search/final_test.cpp:50) as consumed by getRealNearest
 * (search/search_function.h:105-125).  Like the reference's L2Metric
 * (search/support_func.h:111-112) only the first (d/4)*4 dimensions take
 * part in distances. */
int gbdr_index_set_base(gbdr_index *h, const float *db, uint64_t n, uint32_t d);

/* Low-dimensional (transformed) base vectors [n x d_low]
 * (`db_ar`, search/final_test.cpp:56; `ds_low` of performTest,
 * search/search_function.h:129,160). */
int gbdr_index_set_low(gbdr_index *h, const float *db_low, uint64_t n, uint32_t d_low);

/* Search graph as flattened adjacency: offsets[n+1] (uint64) into edges[].
 * Replaces `vector<vector<uint32_t>> main_graph` (search/search_function.h:44,
 * produced by loadEdges, search/support_func.h:231-249).  Neighbour order is
 * preserved (it is semantically relevant at exact distance ties). */
int gbdr_index_set_graph(gbdr_index *h, const uint64_t *offsets, const uint32_t *edges, uint64_t n);

/* The auxiliary ("long link") graph of getOneSearchResults' use_second_graph mode
 * (search/search_function.h:45,73-80; built by KLgraph for naive_test.cpp:98-105), same flattened
 * form and vertex count as the main graph.  hops_bound and llf are the reference's parameters of the
 * same names (:47; performRealTests fixes hops_bound = 50, :309).  offsets == NULL removes it. */
int gbdr_index_set_aux_graph(gbdr_index *h, const uint64_t *offsets, const uint32_t *edges, uint64_t n,
                             uint32_t hops_bound, int llf);

/* Projection net, three matrices in the reference's on-disk layout
 * `[out][in+1]` row-major with the bias in the last column
 * (search/support_func.h:45-49, 624-633; written by
 * dim_red/support_func.py:517-555; loaded at search/final_test.cpp:73-76).
 * l1: [d_hidden x (d+1)], l2: [d_hidden2 x (d_hidden+1)], l3: [d_low x (d_hidden2+1)]. */
int gbdr_index_set_net(gbdr_index *h, const float *l1, const float *l2, const float *l3,
                       uint32_t d, uint32_t d_hidden, uint32_t d_hidden2, uint32_t d_low);

/* For sharded indexes: value added to every result id (global id = local id +
 * id_offset).  Default 0.  Result ids are 32-bit: id_offset + n must stay below 2^31
 * (offsets >= 2^31 are rejected with GBDR_E_INVALID), and an index holds at most 2^31 - 1
 * vertices (gbdr_index_set_graph rejects more): shard larger sets. */
int gbdr_index_set_id_offset(gbdr_index *h, uint64_t id_offset);

/* Projection arithmetic. 0 = GBDR_PROJ_3XTF32 (default): tcgen05 kind::tf32 with
 * 3-term error compensation (fp32-class accuracy, |rel err| <= 2e-6 on q_low);
 * 1 = GBDR_PROJ_TF32: single-pass tf32 (|rel err| <= 2e-3);
 * 2 = GBDR_PROJ_FP32: CUDA-core fp32 FMA GEMM. */
#define GBDR_PROJ_3XTF32 0
#define GBDR_PROJ_TF32 1
#define GBDR_PROJ_FP32 2
int gbdr_index_set_projection_mode(gbdr_index *h, int mode);

/* ------------------------------------------------------------- hot path */

/* y = normalize(W3 relu(W2 relu(W1 x + b1) + b2) + b3) for n_q queries at once.
 * Replaces the per-query GetLowQueryFromNet (search/support_func.h:645-658)
 * called in the loop at search/search_function.h:354-355. */
int gbdr_project(gbdr_index *h, const float *queries, uint32_t n_q, float *q_low);
int gbdr_project_dev(gbdr_index *h, const float *d_queries, uint32_t n_q, float *d_q_low,
                     void *stream);

/* Flags for gbdr_search */
#define GBDR_SEARCH_RERANK 1u  /* re-rank the ef low-dim survivors by exact L2 in the original
                                  dimension (performTest branch search_function.h:158-164) */
#define GBDR_SEARCH_PLAIN 2u   /* search the ORIGINAL vectors (d == d_low branch,
                                  search_function.h:174-182); q_low is ignored          */
#define GBDR_SEARCH_SECOND_GRAPH 4u /* use_second_graph == true (search_function.h:73-89): every hop
                                  first scans the auxiliary adjacency row while hops < hops_bound,
                                  then the main row unless (llf && the auxiliary row produced a
                                  candidate).  Needs gbdr_index_set_aux_graph.            */

/* Batched getOneSearchResults (+ getRealNearest when GBDR_SEARCH_RERANK).
 *
 *   queries  [n_q x d]      original-dimension queries (needed for re-rank, for
 *                           PLAIN, and for projection when q_low == NULL)
 *   q_low    [n_q x d_low]  low-dimensional queries, or NULL -> computed with
 *                           the index's net (performNetTest, search_function.h:353-361)
 *   ef                      beam width (`ef` / `recheck_size`, search_function.h:160-161)
 *   k                       results per query, 1 <= k <= ef
 *   entry    [n_q]          one entry vertex per query (`inter_points[i][0]`,
 *                           search_function.h:54-64,297-307); every id must be < n
 *                           (GBDR_E_INVALID otherwise, checked before anything runs)
 *   out_ids  [n_q x k]      ascending by (dist, tie rule of the reference);
 *                           GBDR_PAD_ID where fewer than k vertices were reached
 *   out_dists[n_q x k]      squared L2 (original dim if RERANK/PLAIN else low dim); may be NULL
 *   hops     [n_q]          TripleResult.hops        (search_function.h:90,100); may be NULL
 *   dist_calc[n_q]          TripleResult.dist_calc   (search_function.h:29,52,100), plus ef
 *                           when RERANK (search_function.h:164); may be NULL
 *   gpu_seconds             device time of the batch incl. H2D/D2H (CUDA events); may be NULL
 *
 * With k == 1 and GBDR_SEARCH_RERANK, out_ids[i] equals `ans[i]` of performTest
 * (search_function.h:163). */
int gbdr_search(gbdr_index *h, const float *queries, const float *q_low, uint32_t n_q,
                uint32_t ef, uint32_t k, uint32_t flags, const uint32_t *entry,
                uint32_t *out_ids, float *out_dists, int32_t *hops, int32_t *dist_calc,
                double *gpu_seconds);

/* The same call split in two, so that a host thread can keep several batches in flight: submit
 * enqueues the H2D copies, the kernels and the D2H copies on the handle's stream and returns; wait
 * blocks until they finished, checks the status word and reports the device time.  All buffers must
 * stay valid (and unchanged / unread) until wait returns; page-locked buffers (gbdr_host_alloc_pinned)
 * make the copies overlap with other handles' kernels.  One call in flight per handle: use
 * gbdr_index_create_view for the second batch.  gbdr_search == submit + wait.
 * A call that repeats the previous one on this handle — same shape, same (page-locked) buffers, unchanged index —
 * is captured into a CUDA graph the second time and replayed as one launch from then on; the buffers' CONTENTS are
 * read at every call as usual.  Per-kernel times (gbdr_kernel_ms) are those of the last call that was not replayed. */
int gbdr_search_submit(gbdr_index *h, const float *queries, const float *q_low, uint32_t n_q,
                       uint32_t ef, uint32_t k, uint32_t flags, const uint32_t *entry,
                       uint32_t *out_ids, float *out_dists, int32_t *hops, int32_t *dist_calc);
int gbdr_search_wait(gbdr_index *h, double *gpu_seconds);

/* Same, all buffers in device memory, asynchronous on `stream`.
 * d_scanned [n_q] (may be NULL) receives the number of adjacency ids scanned per
 * query (roofline accounting, SURVEY.md §8d `E`). */
int gbdr_search_dev(gbdr_index *h, const float *d_queries, const float *d_q_low, uint32_t n_q,
                    uint32_t ef, uint32_t k, uint32_t flags, const uint32_t *d_entry,
                    uint32_t *d_out_ids, float *d_out_dists, int32_t *d_hops,
                    int32_t *d_dist_calc, int32_t *d_scanned, void *stream);

/* Device time (ms) of the kernels of the last gbdr_search* call on this handle,
 * measured with CUDA events on the launching stream.  Any pointer may be NULL.
 * Calling it synchronises the handle's stream. */
int gbdr_last_kernel_ms(gbdr_index *h, float *project_ms, float *search_ms, float *rerank_ms);
/* Same, averaged over the last `last_n` (<= 256) search calls on this handle: every call records
 * its own CUDA events on the launching stream, so a timed loop needs no per-step synchronisation. */
int gbdr_kernel_ms(gbdr_index *h, uint32_t last_n, float *project_ms, float *search_ms,
                   float *rerank_ms);

/* Failure/information flags of the last search on this handle (synchronises the device).
 * bit0: some query spilled its visited set to HBM (informational); bit1: visited-set capacity
 * exhausted; bit2: boundary-tie slack exhausted; bit3: internal watchdog; bit4: an entry id was not a
 * vertex of the graph (that query's results are GBDR_PAD_ID; nothing is read out of bounds).
 * gbdr_search / gbdr_search_wait check the word themselves: bit4 -> GBDR_E_INVALID (the host entry
 * points also validate `entry` before anything is enqueued), bit1/bit2 -> GBDR_E_CAPACITY.  The HBM
 * overflow tables are sized for the beam width (2 x the expected visited count per query, >= 2048 ids)
 * and grown to 65 536 ids per resident query once a call exhausts them: gbdr_search_wait re-runs such
 * a call itself; a gbdr_search_dev caller that reads bit1 here repeats its call, which then gets the
 * large tables.  Callers of the asynchronous gbdr_search_dev check this when they synchronise. */
int gbdr_index_status(gbdr_index *h, uint32_t *flags);

/* Number of kernels this library has launched since load (bench `gpu_launches`). */
uint64_t gbdr_launch_count(void);

/* ------------------------------------------------------- graph building */

/* Exact k nearest neighbours of every row of Q[n_q x d] among B[n x d]
 * (squared L2, ascending by (dist, id); a row of B identical to the query row
 * is included, so for Q == B the row itself is rank 0).
 * Replaces get_nearestneighbors / get_nearestneighbors_partly
 * (dim_red/support_func.py:20-74, 374-384; call sites dim_red/triplet.py:266-272).
 * Distances are the direct-difference fp32 form used by the C++ side
 * (search/support_func.h:107-128), not the expansion form.
 * out_ids [n_q x k] uint32; out_dists [n_q x k] may be NULL.  Host buffers. */
int gbdr_knn(int device, const float *Q, uint64_t n_q, const float *B, uint64_t n, uint32_t d,
             uint32_t k, uint32_t *out_ids, float *out_dists, double *gpu_seconds);
/* Device-pointer variant; q_begin..q_end selects the row block of Q this call
 * computes (row-block sharding across GPUs); output rows are
 * written at [0, q_end-q_begin). */
int gbdr_knn_dev(int device, const float *d_Q, uint64_t q_begin, uint64_t q_end, const float *d_B,
                 uint64_t n, uint32_t d, uint32_t k, uint32_t *d_out_ids, float *d_out_dists,
                 void *stream);

/* hnswlikeGD(graph, ds, M, N, d, metric, reverse, need_const_degree)
 * (search/support_func.h:521-575, with addReverseEdgesForGD :402-445 and
 * getConstantDegreeForGD :466-485), as called by search/prepare_graph.cpp:70.
 * knn given as flattened adjacency (offsets[n+1], edges).  Output: caller
 * provides out_offsets[n+1] and out_edges with capacity n * 2*M entries.
 * The per-vertex prune runs on the GPU; the order-dependent reverse pass
 * reproduces the reference's sequential i-ascending semantics. */
int gbdr_gd_prune(int device, const uint64_t *knn_offsets, const uint32_t *knn_edges,
                  const float *db_low, uint64_t n, uint32_t d_low, uint32_t M, int reverse,
                  int need_const_degree, uint64_t *out_offsets, uint32_t *out_edges,
                  double *gpu_seconds);

/* cutKNNbyK(knn, ds, knn_size, N, d, metric) (search/support_func.h:309-340): every candidate list is ordered by
 * the distance of its entries to the list's own vertex (ascending (dist, id); the reference's std::sort leaves exact
 * ties unordered) and cut to its knn_size nearest; the vertex itself stays if it is listed (distance 0 is not
 * dropped here, unlike hnswlikeGD).  out_edges: capacity n * knn_size. */
int gbdr_knn_cut(int device, const uint64_t *knn_offsets, const uint32_t *knn_edges, const float *db,
                 uint64_t n, uint32_t d, uint32_t knn_size, uint64_t *out_offsets, uint32_t *out_edges,
                 double *gpu_seconds);

/* The same in pieces on DEVICE buffers, so that a kNN result that already lies in HBM (gbdr_knn_dev) never
 * travels to the host and back, and so that the per-vertex part can be split by rows across GPUs:
 *   gbdr_gd_prune_dev   forward lists (support_func.h:528-565) of rows [row_begin, row_end): d_knn holds the
 *                       candidate lists of THOSE rows ([rows x kstride], k ids each, GBDR_PAD_ID tail allowed;
 *                       every id must be < n: they are not validated), d_db_low all n vectors; writes
 *                       d_fwd [rows x 2M] and d_deg [rows].  Asynchronous on `stream`.
 *   gbdr_gd_finish_dev  everything after that, on the forward lists of ALL n vertices (d_fwd [n x 2M], d_deg [n],
 *                       both modified): reverse edges (:402-445: membership tests on the GPU, the order-dependent
 *                       "row full yet" walk on the host), optional constant-degree fill (:466-485, GPU; needs the
 *                       candidate lists of all vertices in d_knn), flattened graph into out_offsets / out_edges
 *                       (host, capacity n * 2M).  Blocks. */
int gbdr_gd_prune_dev(int device, const uint32_t *d_knn, uint32_t k, uint32_t kstride, uint64_t row_begin,
                      uint64_t row_end, const float *d_db_low, uint64_t n, uint32_t d_low, uint32_t M,
                      uint32_t *d_fwd, uint32_t *d_deg, void *stream);
int gbdr_gd_finish_dev(int device, uint32_t *d_fwd, uint32_t *d_deg, uint64_t n, uint32_t M, int reverse,
                       int need_const_degree, const uint32_t *d_knn, uint32_t k, uint32_t kstride,
                       uint64_t *out_offsets, uint32_t *out_edges, void *stream);

/* The whole graph build in HBM: kNN-`knn_k` self-join of db_low (the Python step, dim_red/triplet.py:266-268)
 * followed by hnswlikeGD (search/prepare_graph.cpp:64-74), one upload of db_low, one download of the graph.
 * knn_out (host, [n x knn_k], may be NULL) receives the kNN lists as well (the `_knn_1k_` file), streamed out
 * behind the computation.  timings (may be NULL): seconds of upload, kNN, forward prune, reverse pass + output. */
int gbdr_build_graph(int device, const float *db_low, uint64_t n, uint32_t d_low, uint32_t knn_k, uint32_t M,
                     int reverse, int need_const_degree, uint64_t *out_offsets, uint32_t *out_edges,
                     uint32_t *knn_out, double timings[4]);

/* ------------------------------------------------------ multi-GPU merge */

/* k-way merge of `parts` sorted (dist,id) result lists per query into the best
 * k_out by (dist, id).  in_ids/in_dists: [parts x n_q x k_in] (e.g. the
 * all-gathered per-shard outputs).  Device pointers, asynchronous. */
int gbdr_merge_topk_dev(int device, const uint32_t *d_in_ids, const float *d_in_dists,
                        uint32_t parts, uint32_t n_q, uint32_t k_in, uint32_t k_out,
                        uint32_t *d_out_ids, float *d_out_dists, void *stream);

/* --------------------------------------------------- several GPUs of one node
 * The reference's only parallelism is a thread team over the queries of one batch
 * (omp_set_num_threads / #pragma omp parallel for, search/search_function.h:147-152).  A group is the
 * multi-GPU equivalent behind the same kind of call: one host process, one gbdr_index per device, one
 * internal host thread per device to enqueue its work.
 *   GBDR_GROUP_REPLICATED  every device holds the whole index; a batch is cut into contiguous slices,
 *                          one per device; no exchange step (results land in the caller's buffers).
 *   GBDR_GROUP_SHARDED     device i holds rows [b_i, e_i) of base / low-dimensional base (contiguous,
 *                          balanced: the first n % N shards get one extra row) and a graph over LOCAL
 *                          ids; every device answers every query on its shard (ids made global by the
 *                          shard's id offset) and the per-shard (dist, id) lists are merged on the
 *                          first device.  `entry` is then [n_devices x n_q]: local entry vertices per shard.
 * Exchange step of the sharded mode: GBDR_EXCHANGE_PEER (default where the devices have peer access) =
 * one kernel on the first device that reads the members' result lists where they lie, over NVLink
 * peer loads, and merges them (gather and merge fused); GBDR_EXCHANGE_NCCL = ncclAllGather of the
 * lists + the merge kernel of gbdr_merge_topk_dev (libnccl.so.2 is loaded at run time).  Same results. */
/* A device may be listed more than once (every entry is its own index with its own stream: a set sharded over
 * fewer devices than shards, tests on one GPU); the NCCL exchange needs every device listed once. */
typedef struct gbdr_group gbdr_group;
#define GBDR_GROUP_REPLICATED 0
#define GBDR_GROUP_SHARDED 1
#define GBDR_EXCHANGE_PEER 0
#define GBDR_EXCHANGE_NCCL 1
int gbdr_group_create(const int *devices, int n_devices, int mode, gbdr_group **out);
int gbdr_group_destroy(gbdr_group *g);
int gbdr_group_size(const gbdr_group *g);
/* the per-device index (borrowed: destroyed with the group), e.g. for gbdr_index_set_aux_graph */
int gbdr_group_member(gbdr_group *g, int i, gbdr_index **out);
int gbdr_group_set_exchange(gbdr_group *g, int exchange);
/* same arguments as the gbdr_index_set_* calls; replicated: uploaded to every device (in parallel);
 * sharded: set_base / set_low split the rows and set each shard's id offset */
int gbdr_group_set_net(gbdr_group *g, const float *l1, const float *l2, const float *l3, uint32_t d,
                       uint32_t d_hidden, uint32_t d_hidden2, uint32_t d_low);
int gbdr_group_set_base(gbdr_group *g, const float *db, uint64_t n, uint32_t d);
int gbdr_group_set_low(gbdr_group *g, const float *db_low, uint64_t n, uint32_t d_low);
int gbdr_group_set_graph(gbdr_group *g, const uint64_t *offsets, const uint32_t *edges, uint64_t n);
/* sharded: rows of shard i (valid once set_base / set_low ran) and its graph over local ids 0 .. e_i - b_i - 1 */
int gbdr_group_shard_rows(const gbdr_group *g, int i, uint64_t *begin, uint64_t *end);
int gbdr_group_set_shard_graph(gbdr_group *g, int i, const uint64_t *offsets, const uint32_t *edges, uint64_t n_i);
/* gbdr_search over the group (blocking; host buffers; arguments as gbdr_search except `entry` in sharded
 * mode, see above).  Sharded: hops / dist_calc are summed over the shards; gpu_seconds = device time on the
 * first device from the call's start to the merged result's download.  Replicated: per-slice values, and
 * gpu_seconds = the slowest device's. */
int gbdr_group_search(gbdr_group *g, const float *queries, const float *q_low, uint32_t n_q, uint32_t ef,
                      uint32_t k, uint32_t flags, const uint32_t *entry, uint32_t *out_ids, float *out_dists,
                      int32_t *hops, int32_t *dist_calc, double *gpu_seconds);

/* gbdr_build_graph row-block sharded over the group's devices (either mode): block i of db_low goes up over device
 * i's own PCIe link and the matrix is all-gathered over NVLink (NCCL in place, or peer copies with
 * GBDR_EXCHANGE_PEER); device i computes the kNN lists and forward lists of its block; the forward lists are
 * gathered on the first device for the reverse pass.  Same graph as one device.  timings: upload + all-gather,
 * kNN, forward prune + gather, reverse pass + output. */
int gbdr_group_build_graph(gbdr_group *g, const float *db_low, uint64_t n, uint32_t d_low, uint32_t knn_k,
                           uint32_t M, int reverse, int need_const_degree, uint64_t *out_offsets,
                           uint32_t *out_edges, uint32_t *knn_out, double timings[4]);

/* ------------------------------------------------------- raw device memory
 * Small helpers so hosts without a CUDA toolchain (ctypes, cgo …) can keep
 * buffers resident.  Thin wrappers over cudaMalloc/cudaMemcpy/cudaHostAlloc. */
int gbdr_dev_malloc(int device, size_t bytes, void **out);
int gbdr_dev_free(int device, void *p);
int gbdr_memcpy_h2d(int device, void *dst, const void *src, size_t bytes);
int gbdr_memcpy_d2h(int device, void *dst, const void *src, size_t bytes);
int gbdr_host_alloc_pinned(size_t bytes, void **out);
int gbdr_host_free_pinned(void *p);
/* page-lock / release a buffer the caller allocated itself (std::vector, numpy); GBDR_E_CUDA if the platform
 * refuses (e.g. already registered) — the buffer then stays pageable and copies from it are staged */
int gbdr_host_register(void *p, size_t bytes);
int gbdr_host_unregister(void *p);
int gbdr_device_synchronize(int device);
/* the index's own stream (cudaStream_t as void*) */
int gbdr_index_stream(gbdr_index *h, void **stream);
/* The launch plan the beam-search kernel would get for (ef, d_low or d, vertices): pure host logic, no device
 * needed; for tools/tests.  out[0..9] = kernel variant (0 shared-memory list, 1 register list, 2 batched-merge),
 * list capacity, warps per CTA, CTAs per SM, shared bytes per warp, visited-table bytes, visited entries,
 * tag bits (0 = 32-bit slots), displacement bits, total shared bytes per SM incl. 1 KB reserved per CTA. */
int gbdr_beam_plan_info(uint32_t ef, uint32_t dim, uint64_t n_vertices, int second_graph, uint32_t out[10]);
/* device pointers of the resident components (NULL if not set); for tools/tests */
int gbdr_index_device_ptrs(gbdr_index *h, const float **d_db, const float **d_db_low,
                           const uint32_t **d_adj, uint32_t *adj_stride);

#ifdef __cplusplus
}
#endif
#endif /* GBDR_H_ */
