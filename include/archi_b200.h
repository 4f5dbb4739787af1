/*
 * archi_b200.h -- C ABI of libarchi_b200.so: B200 (sm_100a) implementation of archi's retrieval
 * hot path (embed-tail pool+normalise, exact top-k search, hybrid BM25+dense fusion).
 *
 * The reference (archi-physics/archi v1.2.4) has no FFI for this path: the boundary is the Python
 * class PostgresVectorStore, whose arithmetic is executed by PostgreSQL/pgvector through SQL
 * strings.  Each entry point below names the reference interface it replaces (paths relative to
 * the reference tree).  The Python mirror of that class (archi_b200/vectorstore.py) binds these
 * symbols with ctypes; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ARCHI_E* code otherwise; the message is
 *     available from archi_last_error() (thread-local).
 *   - plain pointers and sizes only.  `*_loc` arguments say where a buffer lives
 *     (ARCHI_HOST / ARCHI_DEVICE).  Device pointers must belong to the store's device.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls whose
 *     inputs and outputs are all on the device only ENQUEUE work -- no host synchronisation, no
 *     device-to-host copy the host waits for -- so searches pipeline back to back and can be
 *     captured in a CUDA graph once the workspaces are warm; calls with a host output synchronise
 *     the stream once, before returning.
 *   - a store handle may be shared between threads; the host side of calls on one handle
 *     serialises on an internal mutex (held while work is enqueued, not while the GPU runs it),
 *     and a call on another stream than the previous call on the handle first waits, on the
 *     device, for that call's work: rows, tombstones and scratch are shared by the handle.  (The
 *     reference store is stateless per call and therefore re-entrant under Flask threads,
 *     postgres_vectorstore.py:94-103.)
 *   - row ids are dense row indices in insertion order (the reference uses a SERIAL column,
 *     src/cli/templates/init.sql:256-276); `id_offset` is added on output so row-sharded stores
 *     can return global ids.
 */
#ifndef ARCHI_B200_H
#define ARCHI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARCHI_ABI_VERSION 2

/* distance metric: postgres_vectorstore.py:74-78 ("cosine" <=>, "l2" <->, "inner_product" <#>) */
enum { ARCHI_COSINE = 0, ARCHI_L2 = 1, ARCHI_IP = 2 };
/* element types */
enum { ARCHI_F32 = 0, ARCHI_BF16 = 1, ARCHI_I32 = 2, ARCHI_I64 = 3 };
/* buffer location */
enum { ARCHI_HOST = 0, ARCHI_DEVICE = 1 };
/* search path selection (ARCHI_PATH_AUTO picks by batch size) */
enum { ARCHI_PATH_AUTO = 0, ARCHI_PATH_STREAM = 1, ARCHI_PATH_TENSOR = 2 };

enum {
    ARCHI_OK = 0,
    ARCHI_EINVAL = -1,   /* bad argument */
    ARCHI_ECUDA = -2,    /* CUDA runtime / driver error */
    ARCHI_ENOMEM = -3,   /* capacity exceeded / allocation failed */
    ARCHI_EIO = -4,      /* save / load */
    ARCHI_EUNSUPPORTED = -5
};

typedef struct archi_store archi_store_t;

const char *archi_last_error(void);
int archi_abi_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t archi_kernel_launches(void);

/* ---- the store: replaces the document_chunks.embedding column (init.sql:256-276) -------------
 * One store = one row shard resident in the HBM of `device`.  dim and metric are fixed at
 * creation, as they are in the reference (vector(D) column + operator class, init.sql:266,282). */
int archi_store_create(int device, int dim, int metric, int storage_dtype, int64_t capacity_rows,
                       archi_store_t **out);
int archi_store_destroy(archi_store_t *s);
/* COUNT(*) (postgres_vectorstore.py:570-585): rows appended and not deleted. */
int archi_store_count(archi_store_t *s, int64_t *out_live_rows);
/* Rows appended so far including deleted ones (= next row id). */
int archi_store_rows(archi_store_t *s, int64_t *out_rows);
int archi_store_info(archi_store_t *s, int *dim, int *metric, int *storage_dtype, int *device,
                     int64_t *capacity_rows);
/* Grow (never shrinks); rows are preserved. */
int archi_store_reserve(archi_store_t *s, int64_t capacity_rows);
/* TRUNCATE (manager.py:103-153). */
int archi_store_reset(archi_store_t *s);

/* INSERT ... %s::vector (postgres_vectorstore.py:168-180, manager.py:414-422): append n rows of
 * `src_dtype` (ARCHI_F32 | ARCHI_BF16) from host or device memory; converts to the storage dtype
 * and records |row|^2.  *out_first_row receives the id of the first appended row. */
int archi_store_append(archi_store_t *s, const void *rows, int src_dtype, int src_loc, int64_t n,
                       void *stream, int64_t *out_first_row);
/* DELETE FROM document_chunks (postgres_vectorstore.py:493-535) and the is_deleted join
 * (:304-308): tombstones; rows_host are row ids on the host. */
int archi_store_delete_rows(archi_store_t *s, const int64_t *rows_host, int64_t n);
/* Copy stored rows back as fp32 (debug / snapshot / tests). */
int archi_store_read_rows(archi_store_t *s, int64_t first_row, int64_t n, float *out_host);
/* Snapshot / restore of the shard (the reference's persistence is the table itself). */
int archi_store_save(archi_store_t *s, const char *path);
int archi_store_load(const char *path, int device, archi_store_t **out);

/* ---- pool + normalise: replaces the tail of Embeddings.embed_documents / embed_query ----------
 * (manager.py:373, postgres_vectorstore.py:143,245,390; sentence-transformers Pooling(mean) +
 * Normalize): out[b] = normalise( sum_t hidden[b,t,:]*mask[b,t] / max(sum_t mask[b,t], 1e-9) ).
 * hidden [B,L,H] (ARCHI_F32 | ARCHI_BF16) and mask [B,L] (ARCHI_I32 | ARCHI_I64) on the device.
 * Either output may be NULL. */
int archi_pool_normalize(const void *hidden_dev, int hidden_dtype, const void *mask_dev,
                         int mask_dtype, int B, int L, int H, void *out_bf16_dev,
                         float *out_f32_dev, void *stream);
/* Same, writing the B rows straight into the store's tail (H must equal the store's dim). */
int archi_pool_normalize_append(archi_store_t *s, const void *hidden_dev, int hidden_dtype,
                                const void *mask_dev, int mask_dtype, int B, int L,
                                float *out_f32_dev, void *stream, int64_t *out_first_row);

/* ---- exact top-k: replaces similarity_search_by_vector_with_score ----------------------------
 * (postgres_vectorstore.py:272-364: SELECT emb <op> q AS distance ... ORDER BY distance ASC
 * LIMIT k, then score = 1 - distance for cosine, the raw distance otherwise, :361).
 * queries [nq, dim] fp32.  filter_mask_dev: optional device bitmask, bit (i & 31) of word i >> 5
 * set = row i passes the WHERE clause (:296-310); deleted rows are always excluded unless
 * include_deleted != 0.
 * Outputs, best first per query: out_scores [nq, k] fp32 in the reference's score convention
 * (cosine: similarity; l2: distance; inner_product: NEGATIVE inner product), out_ids [nq, k]
 * int64 = row id + id_offset.  When fewer than k rows pass, the tail is id -1 / score NaN.
 * Ties on the score are broken by the lower row id.
 * Batches of >= 2 queries take the tensor-core path: a bf16 coarse pass nominates candidates, every
 * returned row is re-scored exactly in fp32 and each query carries a proof that no rejected row can
 * belong to its top-k; a query whose proof fails is re-scanned by the exact streaming kernel inside
 * the same call, driven from the device (no host round trip). */
int archi_search(archi_store_t *s, const float *queries, int queries_loc, int nq, int k,
                 const uint32_t *filter_mask_dev, int include_deleted, int path,
                 float *out_scores, int64_t *out_ids, int out_loc, int64_t id_offset,
                 void *stream);

/* ---- hybrid: replaces hybrid_search's SQL (postgres_vectorstore.py:435-457) ------------------
 * combined = (1.0 - (emb <op> q)) * w_sem + COALESCE(bm25, 0) * w_bm25, ORDER BY combined DESC
 * LIMIT k.  bm25_dev: [nq, rows] fp32 on the device, 0 for rows without a lexical match (the
 * COALESCE), as produced by archi_bm25_accumulate.  out_scores = combined, best first. */
int archi_hybrid_search(archi_store_t *s, const float *queries, int queries_loc, int nq, int k,
                        float w_sem, float w_bm25, const float *bm25_dev,
                        const uint32_t *filter_mask_dev, int include_deleted,
                        float *out_scores, int64_t *out_ids, int out_loc, int64_t id_offset,
                        void *stream);

/* The same statement without a per-row BM25 vector: the query terms' posting lists go in directly.
 * For rows without a lexical match combined = w_sem * semantic, so (w_sem > 0, bm25 >= 0) the exact fused top-k is
 * the top-k of (dense top-k over all rows  U  exact combined score of the rows matching a term): one ordinary
 * search (tensor-core path for batches) plus work proportional to the postings of the query terms.  Queries whose
 * terms match more than 1/8 of the rows (or sign < 0, w_sem <= 0) take the dense-vector scan of archi_hybrid_search.
 * terms: (query, term) occurrences grouped by query -- term_query[j] ascending in [0, nq) --, each with its posting
 * range [post_start[j], post_end[j]) into doc_ids_dev / tfs_dev and its idf (host arrays); doc_len_dev [rows].
 * BM25(row) = sum_j idf[j] * tf*(k1+1) / (tf + k1*(1 - b + b*doc_len[row]/avgdl)) * sign, rows never touched are
 * SQL NULL (COALESCE -> 0).  *out_path (may be NULL): 1 = posting-list path, 2 = dense-vector path.
 * The posting walks and row gathers run on an internal stream forked from and joined back into `stream`, beside the
 * dense search: for the caller the call is ordered on `stream` like every other entry point. */
typedef struct {
    int n_terms;
    const int32_t *term_query;
    const int64_t *post_start;
    const int64_t *post_end;
    const float *idf;
    const int32_t *doc_ids_dev;
    const int32_t *tfs_dev;
    const float *doc_len_dev;
    float avgdl, k1, b, sign;
} archi_bm25_terms_t;
int archi_hybrid_search_terms(archi_store_t *s, const float *queries, int queries_loc, int nq, int k,
                              float w_sem, float w_bm25, const archi_bm25_terms_t *terms,
                              const uint32_t *filter_mask_dev, int include_deleted, float *out_scores,
                              int64_t *out_ids, int out_loc, int64_t id_offset, void *stream, int *out_path);

/* BM25 over device posting lists (replaces pg_textsearch's `chunk_text <@> to_bm25query(...)`,
 * postgres_vectorstore.py:433).  For each query term t (n_terms of them) with postings
 * doc_ids[post_start[t] .. post_end[t]) / tfs[...]:
 *   out[doc] += idf[t] * tf*(k1+1) / (tf + k1*(1 - b + b*doc_len[doc]/avgdl)) * sign
 * Terms are applied in order on `stream` (deterministic sums).  out_dev [rows] fp32 must be zeroed by
 * the caller (rows never touched stay 0 = COALESCE). */
int archi_bm25_accumulate(const int64_t *post_start_host, const int64_t *post_end_host, int n_terms,
                          const float *idf_host, const int32_t *doc_ids_dev, const int32_t *tfs_dev,
                          const float *doc_len_dev, float avgdl, float k1, float b, float sign,
                          float *out_dev, void *stream);

/* ---- shard merge: the final step after the NCCL allgather of per-shard k-lists ---------------
 * lists: scores [n_lists, nq, k] fp32 and ids [n_lists, nq, k] int64 on the device, each list
 * best first (id -1 = empty slot).  larger_is_better: 1 for cosine / hybrid scores, 0 for l2 /
 * inner_product scores (which are distances).  Outputs [nq, k] on the device. */
int archi_merge_topk(int device, const float *scores_dev, const int64_t *ids_dev, int n_lists,
                     int nq, int k, int larger_is_better, float *out_scores_dev,
                     int64_t *out_ids_dev, void *stream);

/* Same merge over lists that are not densely packed: list l starts scores_list_stride fp32 elements
 * (ids_list_stride int64 elements) after list l-1, each list itself a dense [nq, k] block.  This is the
 * layout of ONE all-gather of per-rank records {ids [nq,k] int64 | scores [nq,k] fp32 | padding}, so
 * the shard exchange needs a single collective and no repacking. */
int archi_merge_topk_strided(int device, const float *scores_dev, const int64_t *ids_dev,
                             int64_t scores_list_stride, int64_t ids_list_stride, int n_lists, int nq,
                             int k, int larger_is_better, float *out_scores_dev, int64_t *out_ids_dev,
                             void *stream);

/* ---- shard exchange + merge in one kernel over NVLink peer memory --------------------------
 * Replaces "all-gather the per-rank k-lists with NCCL, then archi_merge_topk_strided" for the ranks
 * of one box that can map each other's memory (CUDA IPC; one process per GPU).  No counterpart in the
 * reference, which is single-backend.  Set-up, once per process group:
 *   archi_exchange_create        allocate this rank's gather buffer ([2][world][max_record_bytes] + flags)
 *   archi_exchange_local_handle  64-byte IPC handle of that buffer; all-gather the handles over any
 *                                transport (torch.distributed, MPI, a file)
 *   archi_exchange_connect       handles [world][64] indexed by rank -> map every peer's buffer
 *                                (ARCHI_EUNSUPPORTED when peer mapping is impossible: keep using NCCL)
 * Per search, on every rank, in the same order, on one stream:
 *   archi_exchange_merge_topk    record_dev = {ids [nq,k] int64 | scores [nq,k] fp32}, 16-byte aligned,
 *                                on the device; it is pushed into every peer's buffer, the kernel waits
 *                                (bounded, 5 s) for all world records of this call and writes the merged
 *                                [nq,k] lists.  larger_is_better as in archi_merge_topk.
 *                                A call whose wait times out writes id -1 / score NaN into ALL its outputs (never a
 *                                merge of stale slots) and latches a host-visible status: every later
 *                                archi_exchange_merge_topk on the handle returns an error without launching.
 *   archi_exchange_status        synchronises the device; *timed_out = 1 if any call gave up waiting.
 * archi_exchange_destroy must be preceded by a barrier across the ranks. */
#define ARCHI_EXCHANGE_HANDLE_BYTES 64
typedef struct archi_exchange archi_exchange_t;
int archi_exchange_create(int device, int rank, int world, int64_t max_record_bytes, archi_exchange_t **out);
int archi_exchange_local_handle(archi_exchange_t *x, void *handle_out);
int archi_exchange_connect(archi_exchange_t *x, const void *handles);
int archi_exchange_merge_topk(archi_exchange_t *x, const void *record_dev, int nq, int k, int larger_is_better,
                              float *out_scores_dev, int64_t *out_ids_dev, void *stream);
int archi_exchange_status(archi_exchange_t *x, int *timed_out);
int archi_exchange_destroy(archi_exchange_t *x);

/* Statistics of the last archi_search on this handle (path taken, passes, tensor-path fallbacks). */
typedef struct {
    int path;              /* ARCHI_PATH_STREAM | ARCHI_PATH_TENSOR */
    int passes;            /* corpus passes */
    int grid;              /* CTAs of the dominant kernel */
    int unverified_queries;/* tensor path: queries whose exactness proof failed and that were re-scanned exactly */
    double last_kernel_ms; /* device time of the dominant kernel(s) (CUDA events), if requested: the scan
                              kernel of one pass, or the tensor path's coarse launches + threshold kernels */
    int coarse_dtype;      /* tensor path: element type the coarse pass read (ARCHI_BF16: bf16 rows or the
                              bf16 shadow of fp32 rows; ARCHI_F32: fp32 rows as tf32) */
    int coarse_launches;   /* tensor path: coarse-kernel launches per pass (warm-up phases + main) */
    int unproven_queries;  /* running total since the store was created: queries returned as id -1 / score NaN
                              because more proofs failed in one launch than the device-side rescue list holds
                              (256).  Only possible for batches > 256 with DEVICE outputs; with host outputs
                              such queries are re-scanned before the call returns. */
} archi_search_stats_t;
/* Synchronises with the last search on the handle (the proof verdicts travel asynchronously). */
int archi_store_last_stats(archi_store_t *s, archi_search_stats_t *out);
/* When enabled, archi_search brackets its dominant kernel with CUDA events on the caller's
 * stream and synchronises to fill last_kernel_ms (bench.py's roofline leg). */
int archi_store_set_timing(archi_store_t *s, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* ARCHI_B200_H */
