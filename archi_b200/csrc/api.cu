// api.cu -- the C ABI of libarchi_b200.so (see include/archi_b200.h for the contract and for the
// reference interface each entry point replaces).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace archi {

static thread_local char t_error[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

// Makes `device` current for the scope of one entry point and restores the caller's device after
// (the host process also runs PyTorch, which tracks the current device itself).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define ARCHI_DEVICE_GUARD(dev)                                              \
    archi::DeviceGuard _guard(dev);                                          \
    if (!_guard.ok) {                                                        \
        archi::set_error("cudaSetDevice(%d) failed", (int)(dev));            \
        return ARCHI_ECUDA;                                                  \
    }

static int alloc_store_buffers(archi_store *s, int64_t capacity, void **data, float **norm2, uint32_t **alive)
{
    *data = nullptr;
    *norm2 = nullptr;
    *alive = nullptr;
    const size_t words = (size_t)((capacity + 31) / 32) + 1;
    if (capacity > 0) {
        ARCHI_CUDA(cudaMalloc(data, (size_t)capacity * s->ld * elt_size(s->dtype)));
        ARCHI_CUDA(cudaMalloc(norm2, (size_t)capacity * sizeof(float)));
    }
    ARCHI_CUDA(cudaMalloc(alive, words * sizeof(uint32_t)));
    ARCHI_CUDA(cudaMemset(*alive, 0, words * sizeof(uint32_t)));
    return ARCHI_OK;
}

static void free_workspace(Workspace &w)
{
    if (w.part_key) cudaFree(w.part_key);
    if (w.part_id) cudaFree(w.part_id);
    if (w.q_dev) cudaFree(w.q_dev);
    if (w.cursor_key) cudaFree(w.cursor_key);
    if (w.cursor_id) cudaFree(w.cursor_id);
    if (w.out_scores) cudaFree(w.out_scores);
    if (w.out_ids) cudaFree(w.out_ids);
    if (w.ev0) cudaEventDestroy(w.ev0);
    if (w.ev1) cudaEventDestroy(w.ev1);
    if (w.resc_key) cudaFree(w.resc_key);
    if (w.resc_id) cudaFree(w.resc_id);
    w = Workspace();
}

static int ensure_query_staging(archi_store *s, int nq)
{
    Workspace &w = s->ws;
    if (w.q_cap < nq) {
        if (w.q_dev) cudaFree(w.q_dev);
        w.q_dev = nullptr;
        int cap = nq < 64 ? 64 : nq;
        ARCHI_CUDA(cudaMalloc(&w.q_dev, (size_t)cap * s->dim * sizeof(float)));
        w.q_cap = cap;
    }
    if (!w.cursor_key) {
        ARCHI_CUDA(cudaMalloc(&w.cursor_key, kMaxQB * sizeof(float)));
        ARCHI_CUDA(cudaMalloc(&w.cursor_id, kMaxQB * sizeof(int)));
    }
    return ARCHI_OK;
}

static int ensure_out_staging(archi_store *s, int64_t elems)
{
    Workspace &w = s->ws;
    if (w.out_cap < elems) {
        if (w.out_scores) cudaFree(w.out_scores);
        if (w.out_ids) cudaFree(w.out_ids);
        w.out_scores = nullptr;
        w.out_ids = nullptr;
        ARCHI_CUDA(cudaMalloc(&w.out_scores, (size_t)elems * sizeof(float)));
        ARCHI_CUDA(cudaMalloc(&w.out_ids, (size_t)elems * sizeof(int64_t)));
        w.out_cap = elems;
    }
    return ARCHI_OK;
}

// Host-driven exact re-scan of every query of a tensor launch whose proof failed (flags in tws.unverified):
// only needed when more proofs failed than the device-side rescue list holds, and only possible with host
// outputs (the call synchronises anyway).  Synchronises `st`.
static int rescan_flagged(archi_store *s, ScanArgs a, const float *q_dev, int nb, int k, float *o_scores, int64_t *o_ids,
                          int64_t id_offset, cudaStream_t st)
{
    std::vector<int> flags(nb);
    ARCHI_CUDA(cudaMemcpyAsync(flags.data(), s->tws.unverified, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    ARCHI_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < nb; ++i) {
        if (!flags[i]) continue;
        a.nqb = 1;
        a.k = k;
        a.queries = q_dev + (size_t)i * s->dim;
        a.bias = nullptr;
        a.cursor_key = nullptr;
        a.cursor_id = nullptr;
        int g2 = 0;
        int rc = launch_scan(s, a, st, &g2);
        if (rc != ARCHI_OK) return rc;
        rc = launch_scan_finalize(s, a, g2, k, 0, o_scores + (size_t)i * k, o_ids + (size_t)i * k, id_offset, nullptr,
                                  nullptr, st);
        if (rc != ARCHI_OK) return rc;
    }
    return ARCHI_OK;
}

// The store's buffers and scratch are shared by every call on the handle: a call on another stream than the
// previous one first waits (on the device) for that call's work.
static int order_after_previous(archi_store *s, cudaStream_t st)
{
    if (!s->order_ev) ARCHI_CUDA(cudaEventCreateWithFlags(&s->order_ev, cudaEventDisableTiming));
    else if (s->order_stream != st) ARCHI_CUDA(cudaStreamWaitEvent(st, s->order_ev, 0));
    return ARCHI_OK;
}

static int mark_done(archi_store *s, cudaStream_t st)
{
    if (!s->order_ev) ARCHI_CUDA(cudaEventCreateWithFlags(&s->order_ev, cudaEventDisableTiming));
    ARCHI_CUDA(cudaEventRecord(s->order_ev, st));
    s->order_stream = st;
    return ARCHI_OK;
}

// The search proper: queries and outputs on the device, the handle's mutex held by the caller.  Nothing here
// waits for the GPU unless `host_sync_ok` (the caller synchronises anyway) and a batch > 2048 needs it.
static int search_core(archi_store *s, const float *q_dev, int nq, int k, const uint32_t *filter_mask_dev,
                       int include_deleted, int path, int hybrid, float w_sem, float w_bias, const float *bias_dev,
                       float *o_scores, int64_t *o_ids, int64_t id_offset, cudaStream_t st, bool host_sync_ok,
                       bool *used_tensor)
{
    if (s->timing && !s->ws.ev0) {
        ARCHI_CUDA(cudaEventCreate(&s->ws.ev0));
        ARCHI_CUDA(cudaEventCreate(&s->ws.ev1));
    }
    // ---- path selection: the tensor-core path needs a real batch, k <= 128 and no per-row bias ----
    bool use_tensor = false;
    if (path == ARCHI_PATH_TENSOR) {
        if (hybrid || !tensor_path_supported(s, k)) {
            set_error("search: the tensor-core path is not supported for this call (hybrid=%d, k=%d, rows=%lld)",
                      hybrid, k, (long long)s->rows);
            return ARCHI_EUNSUPPORTED;
        }
        use_tensor = true;
    } else if (path == ARCHI_PATH_AUTO) {
        use_tensor = !hybrid && nq >= kTensorMinBatch && tensor_path_supported(s, k);
    }
    if (used_tensor) *used_tensor = use_tensor;

    ScanArgs a;
    a.corpus = s->data;
    a.dtype = s->dtype;
    a.n = s->rows;
    a.dim = s->dim;
    a.ld = s->ld;
    a.metric = s->metric;
    a.norm2 = s->norm2;
    a.alive = include_deleted ? nullptr : s->alive;
    a.filter = filter_mask_dev;
    a.hybrid = hybrid;
    a.w_sem = w_sem;
    a.w_bias = w_bias;
    a.bias_stride = s->rows;

    int passes = 0, grid = 0;
    double kernel_ms = 0.0;
    TensorWorkspace &tw = s->tws;
    tw.verdict_pending = false;
    tw.verdict_launches = 0;
    if (use_tensor) {
        // Everything is enqueued without a host round trip: coarse launches, select + exact rescoring +
        // proof, and the device-driven exact re-scan of the queries whose proof failed (a no-op costing a few
        // microseconds when there are none).  The per-launch counters travel to pinned memory behind an event
        // and are read when the host next synchronises anyway (host outputs, archi_store_last_stats).
        const bool multi = nq > kTensorMaxBatch;
        for (int q0 = 0; q0 < nq; q0 += kTensorMaxBatch) {
            const int nb = nq - q0 < kTensorMaxBatch ? nq - q0 : kTensorMaxBatch;
            const int slot = tw.verdict_launches < 64 ? tw.verdict_launches : 63;
            double ms = 0.0;
            int max_sel = 0;
            int rc = launch_tensor_search(s, q_dev + (size_t)q0 * s->dim, nb, k, filter_mask_dev, include_deleted,
                                          o_scores + (size_t)q0 * k, o_ids + (size_t)q0 * k, id_offset, st, &max_sel, &ms);
            if (rc != ARCHI_OK) return rc;
            kernel_ms += ms;
            ++passes;
            grid = s->stats.grid;
            a.k = k;
            a.queries = q_dev + (size_t)q0 * s->dim;
            a.bias = nullptr;
            a.cursor_key = nullptr;
            a.cursor_id = nullptr;
            rc = launch_rescue(s, a, tw.unv_list, tw.unv_count, max_sel, o_scores + (size_t)q0 * k,
                               o_ids + (size_t)q0 * k, id_offset, st);
            if (rc != ARCHI_OK) return rc;
            ARCHI_CUDA(cudaMemcpyAsync(tw.h_verdict + slot, tw.unv_count, sizeof(int), cudaMemcpyDeviceToHost, st));
            tw.verdict_max_sel[slot] = max_sel;
            tw.verdict_launches = slot + 1;
            if (multi && host_sync_ok && nb > max_sel) {
                // more than one launch shares the flag buffer: settle this one before the next overwrites it
                ARCHI_CUDA(cudaStreamSynchronize(st));
                if (tw.h_verdict[slot] > max_sel) {
                    rc = rescan_flagged(s, a, q_dev + (size_t)q0 * s->dim, nb, k, o_scores + (size_t)q0 * k,
                                        o_ids + (size_t)q0 * k, id_offset, st);
                    if (rc != ARCHI_OK) return rc;
                    tw.sticky_fixed += tw.h_verdict[slot] - max_sel;     // those queries were answered after all
                }
            }
        }
        ARCHI_CUDA(cudaEventRecord(tw.verdict_ev, st));
        tw.verdict_pending = true;
    } else {
        for (int q0 = 0; q0 < nq; q0 += kMaxQB) {
            a.nqb = nq - q0 < kMaxQB ? nq - q0 : kMaxQB;
            a.queries = q_dev + (size_t)q0 * s->dim;
            a.bias = bias_dev ? bias_dev + (size_t)q0 * s->rows : nullptr;
            for (int col0 = 0; col0 < k; col0 += kMaxListK) {
                a.k = k - col0 < kMaxListK ? k - col0 : kMaxListK;
                a.cursor_key = col0 > 0 ? s->ws.cursor_key : nullptr;
                a.cursor_id = col0 > 0 ? s->ws.cursor_id : nullptr;
                if (s->timing) ARCHI_CUDA(cudaEventRecord(s->ws.ev0, st));
                int rc = launch_scan(s, a, st, &grid);
                if (rc != ARCHI_OK) return rc;
                if (s->timing) {
                    ARCHI_CUDA(cudaEventRecord(s->ws.ev1, st));
                    ARCHI_CUDA(cudaEventSynchronize(s->ws.ev1));
                    float ms = 0.f;
                    ARCHI_CUDA(cudaEventElapsedTime(&ms, s->ws.ev0, s->ws.ev1));
                    kernel_ms += ms;
                }
                const bool more = col0 + a.k < k;
                rc = launch_scan_finalize(s, a, grid, k, col0, o_scores + (size_t)q0 * k, o_ids + (size_t)q0 * k,
                                          id_offset, more ? s->ws.cursor_key : nullptr,
                                          more ? s->ws.cursor_id : nullptr, st);
                if (rc != ARCHI_OK) return rc;
                ++passes;
            }
        }
    }
    s->stats.path = use_tensor ? ARCHI_PATH_TENSOR : ARCHI_PATH_STREAM;
    s->stats.passes = passes;
    s->stats.grid = grid;
    s->stats.unverified_queries = 0;      // tensor path: filled in lazily from the pinned verdict (archi_store_last_stats)
    s->stats.last_kernel_ms = passes ? kernel_ms / passes : 0.0;
    return ARCHI_OK;
}

// Query / output staging shared by the search entry points.  With host outputs the device-side staging
// buffers are returned and `finish_outputs` copies them back and synchronises.
struct Staged {
    const float *q_dev;
    float *o_scores;
    int64_t *o_ids;
};

static int stage_io(archi_store *s, const float *queries, int queries_loc, int nq, int k, float *out_scores,
                    int64_t *out_ids, int out_loc, cudaStream_t st, Staged *io)
{
    io->q_dev = queries;
    int rc = ensure_query_staging(s, nq);
    if (rc != ARCHI_OK) return rc;
    if (queries_loc == ARCHI_HOST) {
        ARCHI_CUDA(cudaMemcpyAsync(s->ws.q_dev, queries, (size_t)nq * s->dim * sizeof(float), cudaMemcpyHostToDevice, st));
        io->q_dev = s->ws.q_dev;
    }
    io->o_scores = out_scores;
    io->o_ids = out_ids;
    if (out_loc == ARCHI_HOST) {
        if (ensure_out_staging(s, (int64_t)nq * k) != ARCHI_OK) return ARCHI_ECUDA;
        io->o_scores = s->ws.out_scores;
        io->o_ids = s->ws.out_ids;
    }
    return ARCHI_OK;
}

static int copy_outputs_to_host(const Staged &io, int nq, int k, float *out_scores, int64_t *out_ids, cudaStream_t st)
{
    ARCHI_CUDA(cudaMemcpyAsync(out_scores, io.o_scores, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
    ARCHI_CUDA(cudaMemcpyAsync(out_ids, io.o_ids, (size_t)nq * k * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    ARCHI_CUDA(cudaStreamSynchronize(st));
    return ARCHI_OK;
}

static int check_search_args(archi_store *s, const float *queries, int queries_loc, int nq, int k, float *out_scores,
                             int64_t *out_ids, int out_loc)
{
    ARCHI_REQUIRE(s != nullptr, "search: null store");
    ARCHI_REQUIRE(nq >= 0 && k >= 0, "search: nq=%d k=%d must be non-negative", nq, k);
    ARCHI_REQUIRE(nq == 0 || queries != nullptr, "search: null queries");
    ARCHI_REQUIRE(nq == 0 || k == 0 || (out_scores && out_ids), "search: null outputs");
    ARCHI_REQUIRE(queries_loc == ARCHI_HOST || queries_loc == ARCHI_DEVICE, "search: bad queries_loc");
    ARCHI_REQUIRE(out_loc == ARCHI_HOST || out_loc == ARCHI_DEVICE, "search: bad out_loc");
    return ARCHI_OK;
}

// The common body of archi_search / archi_hybrid_search.
static int search_impl(archi_store *s, const float *queries, int queries_loc, int nq, int k,
                       const uint32_t *filter_mask_dev, int include_deleted, int path, int hybrid,
                       float w_sem, float w_bias, const float *bias_dev, float *out_scores,
                       int64_t *out_ids, int out_loc, int64_t id_offset, void *stream)
{
    int rc = check_search_args(s, queries, queries_loc, nq, k, out_scores, out_ids, out_loc);
    if (rc != ARCHI_OK) return rc;
    ARCHI_REQUIRE(path == ARCHI_PATH_AUTO || path == ARCHI_PATH_STREAM || path == ARCHI_PATH_TENSOR,
                  "search: bad path %d", path);
    if (nq == 0 || k == 0) return ARCHI_OK;

    std::lock_guard<std::mutex> lock(s->mu);
    ARCHI_DEVICE_GUARD(s->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if ((rc = order_after_previous(s, st)) != ARCHI_OK) return rc;
    Staged io;
    if ((rc = stage_io(s, queries, queries_loc, nq, k, out_scores, out_ids, out_loc, st, &io)) != ARCHI_OK) return rc;
    bool used_tensor = false;
    rc = search_core(s, io.q_dev, nq, k, filter_mask_dev, include_deleted, path, hybrid, w_sem, w_bias, bias_dev,
                     io.o_scores, io.o_ids, id_offset, st, out_loc == ARCHI_HOST, &used_tensor);
    if (rc != ARCHI_OK) return rc;
    if (out_loc == ARCHI_HOST) {
        if ((rc = copy_outputs_to_host(io, nq, k, out_scores, out_ids, st)) != ARCHI_OK) return rc;
        TensorWorkspace &tw = s->tws;
        if (used_tensor && nq <= kTensorMaxBatch && tw.h_verdict[0] > tw.verdict_max_sel[0]) {
            // (only batches > 256 can get here) more proofs failed than the device-side rescue takes:
            // re-scan every flagged query now and fetch the rows again
            ScanArgs a;
            a.corpus = s->data;
            a.dtype = s->dtype;
            a.n = s->rows;
            a.dim = s->dim;
            a.ld = s->ld;
            a.metric = s->metric;
            a.norm2 = s->norm2;
            a.alive = include_deleted ? nullptr : s->alive;
            a.filter = filter_mask_dev;
            a.hybrid = 0;
            a.w_sem = 1.f;
            a.w_bias = 0.f;
            a.bias_stride = s->rows;
            if ((rc = rescan_flagged(s, a, io.q_dev, nq, k, io.o_scores, io.o_ids, id_offset, st)) != ARCHI_OK) return rc;
            if ((rc = copy_outputs_to_host(io, nq, k, out_scores, out_ids, st)) != ARCHI_OK) return rc;
            tw.sticky_fixed += tw.h_verdict[0] - tw.verdict_max_sel[0];  // those queries were answered after all
        }
    }
    return mark_done(s, st);
}

// ---- hybrid over posting lists (hybrid.cu) --------------------------------------------------------------------
// Dense-vector fallback for one query whose terms match a large part of the corpus: BM25 accumulated into a
// [rows] vector (zeroed first), then the streaming scan with the fused bias term.
static int hybrid_dense_vector_one(archi_store *s, const float *q_dev, int k, float w_sem, float w_bm25, const HybridTerms &t,
                                   int pair0, int n_pairs, const uint32_t *filter, int include_deleted, float *o_scores,
                                   int64_t *o_ids, int64_t id_offset, cudaStream_t st)
{
    HybridWorkspace &w = s->hws;
    const size_t need = (size_t)(s->capacity > 0 ? s->capacity : 1) * sizeof(float);
    if (!w.bias || w.bias_bytes < need) {
        if (w.bias) cudaFree(w.bias);
        w.bias = nullptr;
        w.bias_bytes = 0;
        ARCHI_CUDA(cudaMalloc(&w.bias, need));
        w.bias_bytes = need;
    }
    ARCHI_CUDA(cudaMemsetAsync(w.bias, 0, (size_t)s->rows * sizeof(float), st));
    for (int j = 0; j < n_pairs; ++j) {
        const int64_t b0 = t.post_start[pair0 + j], b1 = t.post_end[pair0 + j];
        int rc = launch_bm25(t.doc_ids_dev + b0, t.tfs_dev + b0, b1 - b0, t.idf[pair0 + j], t.doc_len_dev, t.avgdl, t.k1, t.b,
                             t.sign, w.bias, st);
        if (rc != ARCHI_OK) return rc;
    }
    return search_core(s, q_dev, 1, k, filter, include_deleted, ARCHI_PATH_STREAM, 1, w_sem, w_bm25, w.bias, o_scores, o_ids,
                       id_offset, st, false, nullptr);
}

static int hybrid_terms_impl(archi_store *s, const float *queries, int queries_loc, int nq, int k, float w_sem, float w_bm25,
                             const archi_bm25_terms_t *terms, const uint32_t *filter, int include_deleted, float *out_scores,
                             int64_t *out_ids, int out_loc, int64_t id_offset, void *stream, int *out_path)
{
    int rc = check_search_args(s, queries, queries_loc, nq, k, out_scores, out_ids, out_loc);
    if (rc != ARCHI_OK) return rc;
    ARCHI_REQUIRE(terms != nullptr && terms->n_terms >= 0, "hybrid_search_terms: null terms");
    ARCHI_REQUIRE(terms->n_terms == 0 || (terms->term_query && terms->post_start && terms->post_end && terms->idf &&
                                          terms->doc_ids_dev && terms->tfs_dev && terms->doc_len_dev),
                  "hybrid_search_terms: null posting arrays");
    ARCHI_REQUIRE(terms->avgdl > 0.f, "hybrid_search_terms: avgdl must be positive");
    if (out_path) *out_path = 0;
    if (nq == 0 || k == 0) return ARCHI_OK;
    HybridTerms t;
    t.post_start = terms->post_start;
    t.post_end = terms->post_end;
    t.idf = terms->idf;
    t.doc_ids_dev = terms->doc_ids_dev;
    t.tfs_dev = terms->tfs_dev;
    t.doc_len_dev = terms->doc_len_dev;
    t.avgdl = terms->avgdl;
    t.k1 = terms->k1;
    t.b = terms->b;
    t.sign = terms->sign;
    // per-query term ranges (term_query ascending) and posting counts
    std::vector<int> first(nq + 1, 0);
    std::vector<long long> postings(nq, 0);
    {
        int prev = 0;
        for (int j = 0; j < terms->n_terms; ++j) {
            const int q = terms->term_query[j];
            ARCHI_REQUIRE(q >= prev && q < nq, "hybrid_search_terms: term_query must be ascending and < nq");
            ARCHI_REQUIRE(t.post_start[j] >= 0 && t.post_end[j] >= t.post_start[j], "hybrid_search_terms: bad posting range");
            prev = q;
            first[q + 1] = j + 1;
            postings[q] += t.post_end[j] - t.post_start[j];
        }
        for (int q = 0; q < nq; ++q)
            if (first[q + 1] < first[q]) first[q + 1] = first[q];
    }

    std::lock_guard<std::mutex> lock(s->mu);
    ARCHI_DEVICE_GUARD(s->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if ((rc = order_after_previous(s, st)) != ARCHI_OK) return rc;
    Staged io;
    if ((rc = stage_io(s, queries, queries_loc, nq, k, out_scores, out_ids, out_loc, st, &io)) != ARCHI_OK) return rc;

    // The sparse decomposition needs bm25 >= 0 (sign > 0), w_sem > 0 and w_bm25 >= 0; it pays while the terms of a
    // query match a small part of the corpus.
    const long long sparse_limit = s->rows / 8 > 4096 ? s->rows / 8 : 4096;
    auto sparse_ok = [&](int q) {
        return t.sign > 0.f && w_sem > 0.f && w_bm25 >= 0.f && k <= kMaxListK && s->rows > 0 && postings[q] <= sparse_limit &&
               first[q + 1] - first[q] <= kHybMaxPairsHost;
    };
    int n_sparse = 0;
    for (int q = 0; q < nq; ++q) n_sparse += sparse_ok(q) ? 1 : 0;
    HybridWorkspace &w = s->hws;
    if (n_sparse == nq) {
        // dense top-k of the whole batch first (<= 256 queries per call: every failed proof is rescued on the device)
        const size_t need_s = (size_t)nq * k * sizeof(float), need_i = (size_t)nq * k * sizeof(int64_t);
        if (!w.dense_scores || w.dense_bytes < need_i) {
            if (w.dense_scores) cudaFree(w.dense_scores);
            if (w.dense_ids) cudaFree(w.dense_ids);
            w.dense_scores = nullptr;
            w.dense_ids = nullptr;
            w.dense_bytes = 0;
            ARCHI_CUDA(cudaMalloc(&w.dense_scores, need_s));
            ARCHI_CUDA(cudaMalloc(&w.dense_ids, need_i));
            w.dense_bytes = need_i;
        }
        // The sparse chain of every round goes to the side stream first (posting walks and row gathers: latency
        // bound, a few CTAs), the dense search then streams the corpus on the caller's stream beside it; the merges
        // follow the join.
        if (!w.side) {
            ARCHI_CUDA(cudaStreamCreateWithFlags(&w.side, cudaStreamNonBlocking));
            ARCHI_CUDA(cudaEventCreateWithFlags(&w.ev_fork, cudaEventDisableTiming));
            ARCHI_CUDA(cudaEventCreateWithFlags(&w.ev_join, cudaEventDisableTiming));
        }
        if ((rc = hybrid_ensure_part_lists(s, nq, st)) != ARCHI_OK) return rc;
        ARCHI_CUDA(cudaEventRecord(w.ev_fork, st));
        ARCHI_CUDA(cudaStreamWaitEvent(w.side, w.ev_fork, 0));
        struct Round { int q0, q1, cps; };
        std::vector<Round> rounds;
        int q0 = 0;
        while (q0 < nq) {          // rounds of <= kHybMaxSlotsHost queries and <= kHybMaxPairsHost terms
            int q1 = q0, pairs = 0;
            while (q1 < nq && q1 - q0 < kHybMaxSlotsHost && pairs + (first[q1 + 1] - first[q1]) <= kHybMaxPairsHost) {
                pairs += first[q1 + 1] - first[q1];
                ++q1;
            }
            std::vector<int> pair_slot(pairs > 0 ? pairs : 1);
            for (int q = q0; q < q1; ++q)
                for (int j = first[q]; j < first[q + 1]; ++j) pair_slot[j - first[q0]] = q - q0;
            int cps = 1;
            rc = launch_hybrid_sparse_round(s, io.q_dev + (size_t)q0 * s->dim, q1 - q0, k, q0, w_sem, w_bm25, t.sign, t, first[q0],
                                            pairs, pair_slot.data(), filter, include_deleted, w.side, &cps);
            if (rc != ARCHI_OK) {
                cudaEventRecord(w.ev_join, w.side);          // never leave the side stream dangling
                cudaStreamWaitEvent(st, w.ev_join, 0);
                return rc;
            }
            rounds.push_back({q0, q1, cps});
            q0 = q1;
        }
        ARCHI_CUDA(cudaEventRecord(w.ev_join, w.side));
        // dense top-k of the whole batch (<= 256 queries per call: every failed proof is rescued on the device)
        for (int d0 = 0; d0 < nq; d0 += 256) {
            const int nb = nq - d0 < 256 ? nq - d0 : 256;
            rc = search_core(s, io.q_dev + (size_t)d0 * s->dim, nb, k, filter, include_deleted, ARCHI_PATH_AUTO, 0, 1.f, 0.f,
                             nullptr, w.dense_scores + (size_t)d0 * k, w.dense_ids + (size_t)d0 * k, id_offset, st, false, nullptr);
            if (rc != ARCHI_OK) {
                cudaStreamWaitEvent(st, w.ev_join, 0);
                return rc;
            }
        }
        ARCHI_CUDA(cudaStreamWaitEvent(st, w.ev_join, 0));
        for (const Round &rd : rounds) {
            rc = launch_hybrid_merge(s, rd.q1 - rd.q0, k, rd.q0, rd.cps, w.dense_scores + (size_t)rd.q0 * k,
                                     w.dense_ids + (size_t)rd.q0 * k, w_sem, io.o_scores + (size_t)rd.q0 * k,
                                     io.o_ids + (size_t)rd.q0 * k, id_offset, st);
            if (rc != ARCHI_OK) return rc;
        }
        if (out_path) *out_path = 1;
    } else {
        for (int q = 0; q < nq; ++q) {
            rc = hybrid_dense_vector_one(s, io.q_dev + (size_t)q * s->dim, k, w_sem, w_bm25, t, first[q], first[q + 1] - first[q],
                                         filter, include_deleted, io.o_scores + (size_t)q * k, io.o_ids + (size_t)q * k, id_offset, st);
            if (rc != ARCHI_OK) return rc;
        }
        if (out_path) *out_path = 2;
    }
    if (out_loc == ARCHI_HOST && (rc = copy_outputs_to_host(io, nq, k, out_scores, out_ids, st)) != ARCHI_OK) return rc;
    return mark_done(s, st);
}

}  // namespace archi

using namespace archi;

extern "C" {

const char *archi_last_error(void) { return t_error; }
int archi_abi_version(void) { return ARCHI_ABI_VERSION; }
int64_t archi_kernel_launches(void) { return g_launches.load(); }

int archi_store_create(int device, int dim, int metric, int storage_dtype, int64_t capacity_rows,
                       archi_store_t **out)
{
    ARCHI_REQUIRE(out != nullptr, "store_create: null out");
    *out = nullptr;
    ARCHI_REQUIRE(dim >= 1 && dim <= 16384, "store_create: dim=%d out of range [1, 16384]", dim);
    ARCHI_REQUIRE(metric == ARCHI_COSINE || metric == ARCHI_L2 || metric == ARCHI_IP,
                  "store_create: distance_metric must be one of cosine, l2, inner_product");
    ARCHI_REQUIRE(storage_dtype == ARCHI_F32 || storage_dtype == ARCHI_BF16,
                  "store_create: storage dtype must be f32 or bf16");
    ARCHI_REQUIRE(capacity_rows >= 0 && capacity_rows < (1ll << 31), "store_create: capacity out of range");
    int ndev = 0;
    ARCHI_CUDA(cudaGetDeviceCount(&ndev));
    ARCHI_REQUIRE(device >= 0 && device < ndev, "store_create: device %d not present (%d visible)", device, ndev);
    ARCHI_DEVICE_GUARD(device);
    cudaDeviceProp prop;
    ARCHI_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("store_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                  prop.minor);
        return ARCHI_EUNSUPPORTED;
    }
    archi_store *s = new archi_store();
    s->device = device;
    s->dim = dim;
    s->ld = round_up(dim, storage_dtype == ARCHI_BF16 ? 8 : 4);
    s->metric = metric;
    s->dtype = storage_dtype;
    s->capacity = capacity_rows;
    s->sm_count = prop.multiProcessorCount;
    int rc = alloc_store_buffers(s, capacity_rows, &s->data, &s->norm2, &s->alive);
    if (rc != ARCHI_OK) {
        delete s;
        return rc == ARCHI_ECUDA ? ARCHI_ENOMEM : rc;
    }
    *out = s;
    return ARCHI_OK;
}

int archi_store_destroy(archi_store_t *s)
{
    if (!s) return ARCHI_OK;
    ARCHI_DEVICE_GUARD(s->device);
    cudaDeviceSynchronize();
    if (s->data) cudaFree(s->data);
    if (s->norm2) cudaFree(s->norm2);
    if (s->alive) cudaFree(s->alive);
    free_workspace(s->ws);
    free_tensor_workspace(s->tws);
    free_hybrid_workspace(s->hws);
    if (s->order_ev) cudaEventDestroy(s->order_ev);
    delete s;
    return ARCHI_OK;
}

int archi_store_count(archi_store_t *s, int64_t *out_live_rows)
{
    ARCHI_REQUIRE(s && out_live_rows, "store_count: null argument");
    std::lock_guard<std::mutex> lock(s->mu);
    *out_live_rows = s->rows - s->deleted;
    return ARCHI_OK;
}

int archi_store_rows(archi_store_t *s, int64_t *out_rows)
{
    ARCHI_REQUIRE(s && out_rows, "store_rows: null argument");
    std::lock_guard<std::mutex> lock(s->mu);
    *out_rows = s->rows;
    return ARCHI_OK;
}

int archi_store_info(archi_store_t *s, int *dim, int *metric, int *storage_dtype, int *device,
                     int64_t *capacity_rows)
{
    ARCHI_REQUIRE(s != nullptr, "store_info: null store");
    if (dim) *dim = s->dim;
    if (metric) *metric = s->metric;
    if (storage_dtype) *storage_dtype = s->dtype;
    if (device) *device = s->device;
    if (capacity_rows) *capacity_rows = s->capacity;
    return ARCHI_OK;
}

static int reserve_locked(archi_store *s, int64_t capacity_rows)
{
    if (capacity_rows <= s->capacity) return ARCHI_OK;
    ARCHI_REQUIRE(capacity_rows < (1ll << 31), "store_reserve: capacity out of range");
    ARCHI_DEVICE_GUARD(s->device);
    ARCHI_CUDA(cudaDeviceSynchronize());
    void *data;
    float *norm2;
    uint32_t *alive;
    int rc = alloc_store_buffers(s, capacity_rows, &data, &norm2, &alive);
    if (rc != ARCHI_OK) return rc == ARCHI_ECUDA ? ARCHI_ENOMEM : rc;
    if (s->rows > 0) {
        ARCHI_CUDA(cudaMemcpy(data, s->data, (size_t)s->rows * s->ld * elt_size(s->dtype), cudaMemcpyDeviceToDevice));
        ARCHI_CUDA(cudaMemcpy(norm2, s->norm2, (size_t)s->rows * sizeof(float), cudaMemcpyDeviceToDevice));
    }
    ARCHI_CUDA(cudaMemcpy(alive, s->alive, (size_t)((s->capacity + 31) / 32 + 1) * sizeof(uint32_t),
                          cudaMemcpyDeviceToDevice));
    if (s->data) cudaFree(s->data);
    if (s->norm2) cudaFree(s->norm2);
    cudaFree(s->alive);
    s->data = data;
    s->norm2 = norm2;
    s->alive = alive;
    s->capacity = capacity_rows;
    return ARCHI_OK;
}

int archi_store_reserve(archi_store_t *s, int64_t capacity_rows)
{
    ARCHI_REQUIRE(s != nullptr, "store_reserve: null store");
    std::lock_guard<std::mutex> lock(s->mu);
    return reserve_locked(s, capacity_rows);
}

int archi_store_reset(archi_store_t *s)
{
    ARCHI_REQUIRE(s != nullptr, "store_reset: null store");
    std::lock_guard<std::mutex> lock(s->mu);
    ARCHI_DEVICE_GUARD(s->device);
    ARCHI_CUDA(cudaDeviceSynchronize());
    ARCHI_CUDA(cudaMemset(s->alive, 0, (size_t)((s->capacity + 31) / 32 + 1) * sizeof(uint32_t)));
    s->rows = 0;
    s->deleted = 0;
    s->epoch++;
    s->reset_epoch++;
    return ARCHI_OK;
}

int archi_store_append(archi_store_t *s, const void *rows, int src_dtype, int src_loc, int64_t n, void *stream,
                       int64_t *out_first_row)
{
    ARCHI_REQUIRE(s != nullptr, "store_append: null store");
    ARCHI_REQUIRE(n >= 0, "store_append: n < 0");
    ARCHI_REQUIRE(n == 0 || rows != nullptr, "store_append: null rows");
    ARCHI_REQUIRE(src_dtype == ARCHI_F32 || src_dtype == ARCHI_BF16, "store_append: source dtype must be f32 or bf16");
    std::lock_guard<std::mutex> lock(s->mu);
    if (out_first_row) *out_first_row = s->rows;
    if (n == 0) return ARCHI_OK;
    ARCHI_DEVICE_GUARD(s->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    {
        const int rc = order_after_previous(s, st);
        if (rc != ARCHI_OK) return rc;
    }
    if (s->rows + n > s->capacity) {
        int64_t want = s->capacity * 2 > s->rows + n ? s->capacity * 2 : s->rows + n;
        if (want < 1024) want = 1024;
        if (want >= (1ll << 31)) want = (1ll << 31) - 1;
        if (want < s->rows + n) {
            set_error("store_append: %lld rows exceed the per-shard limit", (long long)(s->rows + n));
            return ARCHI_ENOMEM;
        }
        int rc = reserve_locked(s, want);
        if (rc != ARCHI_OK) return rc;
    }
    const void *src_dev = rows;
    void *tmp = nullptr;
    if (src_loc == ARCHI_HOST) {
        const size_t bytes = (size_t)n * s->dim * elt_size(src_dtype);
        ARCHI_CUDA(cudaMallocAsync(&tmp, bytes, st));
        ARCHI_CUDA(cudaMemcpyAsync(tmp, rows, bytes, cudaMemcpyHostToDevice, st));
        src_dev = tmp;
    }
    int rc = launch_append(s, src_dev, src_dtype, s->rows, n, st);
    if (tmp) {
        cudaFreeAsync(tmp, st);
        cudaStreamSynchronize(st);  // the host buffer may be reused by the caller
    }
    if (rc != ARCHI_OK) return rc;
    s->rows += n;
    s->epoch++;
    return mark_done(s, st);
}

int archi_store_delete_rows(archi_store_t *s, const int64_t *rows_host, int64_t n)
{
    ARCHI_REQUIRE(s != nullptr, "store_delete_rows: null store");
    ARCHI_REQUIRE(n >= 0 && (n == 0 || rows_host), "store_delete_rows: bad arguments");
    if (n == 0) return ARCHI_OK;
    std::lock_guard<std::mutex> lock(s->mu);
    ARCHI_DEVICE_GUARD(s->device);
    {
        const int rc = order_after_previous(s, nullptr);     // runs on the legacy stream, after the previous call
        if (rc != ARCHI_OK) return rc;
    }
    // one allocation: [n row ids | changed counter]; freed on every path
    long long *rows_dev = nullptr;
    ARCHI_CUDA(cudaMalloc(&rows_dev, (size_t)(n + 1) * sizeof(long long)));
    int *changed_dev = reinterpret_cast<int *>(rows_dev + n);
    int changed = 0;
    int rc = ARCHI_OK;
    cudaError_t e = cudaMemset(changed_dev, 0, sizeof(long long));
    if (e == cudaSuccess) e = cudaMemcpy(rows_dev, rows_host, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = launch_delete_rows(s, rows_dev, n, changed_dev, 0);
        if (rc == ARCHI_OK) e = cudaMemcpy(&changed, changed_dev, sizeof(int), cudaMemcpyDeviceToHost);
    }
    cudaFree(rows_dev);
    if (e != cudaSuccess) {
        set_error("store_delete_rows: %s", cudaGetErrorString(e));
        return ARCHI_ECUDA;
    }
    if (rc != ARCHI_OK) return rc;
    s->deleted += changed;
    s->epoch++;
    return mark_done(s, nullptr);
}

int archi_store_read_rows(archi_store_t *s, int64_t first_row, int64_t n, float *out_host)
{
    ARCHI_REQUIRE(s != nullptr && (n == 0 || out_host), "store_read_rows: null argument");
    std::lock_guard<std::mutex> lock(s->mu);
    ARCHI_REQUIRE(first_row >= 0 && n >= 0 && first_row + n <= s->rows, "store_read_rows: range [%lld, %lld) outside [0, %lld)",
                  (long long)first_row, (long long)(first_row + n), (long long)s->rows);
    if (n == 0) return ARCHI_OK;
    ARCHI_DEVICE_GUARD(s->device);
    {
        const int rc = order_after_previous(s, nullptr);
        if (rc != ARCHI_OK) return rc;
    }
    float *tmp = nullptr;
    ARCHI_CUDA(cudaMalloc(&tmp, (size_t)n * s->dim * sizeof(float)));
    int rc = launch_read_rows(s, first_row, n, tmp, 0);
    if (rc == ARCHI_OK)
        ARCHI_CUDA(cudaMemcpy(out_host, tmp, (size_t)n * s->dim * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(tmp);
    return rc;
}

// ---- snapshot / restore -------------------------------------------------------------------------
struct SnapshotHeader {
    char magic[8];  // "ARCHIB2\0"
    int32_t version, dim, ld, metric, dtype, reserved;
    int64_t rows, deleted;
};

static int copy_dev_to_file(FILE *f, const void *dev, size_t bytes)
{
    const size_t chunk = 64u << 20;
    std::vector<char> buf(bytes < chunk ? bytes : chunk);
    for (size_t off = 0; off < bytes; off += chunk) {
        const size_t nb = bytes - off < chunk ? bytes - off : chunk;
        ARCHI_CUDA(cudaMemcpy(buf.data(), (const char *)dev + off, nb, cudaMemcpyDeviceToHost));
        if (fwrite(buf.data(), 1, nb, f) != nb) {
            set_error("store_save: short write");
            return ARCHI_EIO;
        }
    }
    return ARCHI_OK;
}

static int copy_file_to_dev(FILE *f, void *dev, size_t bytes)
{
    const size_t chunk = 64u << 20;
    std::vector<char> buf(bytes < chunk ? bytes : chunk);
    for (size_t off = 0; off < bytes; off += chunk) {
        const size_t nb = bytes - off < chunk ? bytes - off : chunk;
        if (fread(buf.data(), 1, nb, f) != nb) {
            set_error("store_load: short read");
            return ARCHI_EIO;
        }
        ARCHI_CUDA(cudaMemcpy((char *)dev + off, buf.data(), nb, cudaMemcpyHostToDevice));
    }
    return ARCHI_OK;
}

int archi_store_save(archi_store_t *s, const char *path)
{
    ARCHI_REQUIRE(s && path, "store_save: null argument");
    std::lock_guard<std::mutex> lock(s->mu);
    ARCHI_DEVICE_GUARD(s->device);
    ARCHI_CUDA(cudaDeviceSynchronize());
    FILE *f = fopen(path, "wb");
    if (!f) {
        set_error("store_save: cannot open %s", path);
        return ARCHI_EIO;
    }
    SnapshotHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "ARCHIB2", 8);
    h.version = 1;
    h.dim = s->dim;
    h.ld = s->ld;
    h.metric = s->metric;
    h.dtype = s->dtype;
    h.rows = s->rows;
    h.deleted = s->deleted;
    int rc = fwrite(&h, sizeof(h), 1, f) == 1 ? ARCHI_OK : ARCHI_EIO;
    if (rc == ARCHI_OK && s->rows > 0) {
        rc = copy_dev_to_file(f, s->data, (size_t)s->rows * s->ld * elt_size(s->dtype));
        if (rc == ARCHI_OK) rc = copy_dev_to_file(f, s->norm2, (size_t)s->rows * sizeof(float));
        if (rc == ARCHI_OK) rc = copy_dev_to_file(f, s->alive, (size_t)((s->rows + 31) / 32) * sizeof(uint32_t));
    }
    fclose(f);
    if (rc == ARCHI_EIO && t_error[0] == 0) set_error("store_save: write failed");
    return rc;
}

int archi_store_load(const char *path, int device, archi_store_t **out)
{
    ARCHI_REQUIRE(path && out, "store_load: null argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) {
        set_error("store_load: cannot open %s", path);
        return ARCHI_EIO;
    }
    SnapshotHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "ARCHIB2", 8) != 0 || h.version != 1) {
        fclose(f);
        set_error("store_load: %s is not an archi_b200 snapshot", path);
        return ARCHI_EIO;
    }
    archi_store *s = nullptr;
    int rc = archi_store_create(device, h.dim, h.metric, h.dtype, h.rows, &s);
    if (rc != ARCHI_OK) {
        fclose(f);
        return rc;
    }
    if (h.rows > 0) {
        rc = copy_file_to_dev(f, s->data, (size_t)h.rows * s->ld * elt_size(s->dtype));
        if (rc == ARCHI_OK) rc = copy_file_to_dev(f, s->norm2, (size_t)h.rows * sizeof(float));
        if (rc == ARCHI_OK) rc = copy_file_to_dev(f, s->alive, (size_t)((h.rows + 31) / 32) * sizeof(uint32_t));
    }
    fclose(f);
    if (rc != ARCHI_OK) {
        archi_store_destroy(s);
        return rc;
    }
    s->rows = h.rows;
    s->deleted = h.deleted;
    *out = s;
    return ARCHI_OK;
}

// ---- pool + normalise -----------------------------------------------------------------------------
int archi_pool_normalize(const void *hidden_dev, int hidden_dtype, const void *mask_dev, int mask_dtype, int B, int L,
                         int H, void *out_bf16_dev, float *out_f32_dev, void *stream)
{
    ARCHI_REQUIRE(B == 0 || (hidden_dev && mask_dev), "pool_normalize: null input");
    return launch_pool_normalize(hidden_dev, hidden_dtype, mask_dev, mask_dtype, B, L, H, nullptr, ARCHI_F32, 0,
                                 nullptr, nullptr, 0, out_bf16_dev, out_f32_dev,
                                 reinterpret_cast<cudaStream_t>(stream));
}

int archi_pool_normalize_append(archi_store_t *s, const void *hidden_dev, int hidden_dtype, const void *mask_dev,
                                int mask_dtype, int B, int L, float *out_f32_dev, void *stream,
                                int64_t *out_first_row)
{
    ARCHI_REQUIRE(s != nullptr, "pool_normalize_append: null store");
    ARCHI_REQUIRE(B >= 0 && (B == 0 || (hidden_dev && mask_dev)), "pool_normalize_append: bad input");
    std::lock_guard<std::mutex> lock(s->mu);
    if (out_first_row) *out_first_row = s->rows;
    if (B == 0) return ARCHI_OK;
    ARCHI_DEVICE_GUARD(s->device);
    {
        const int rc = order_after_previous(s, reinterpret_cast<cudaStream_t>(stream));
        if (rc != ARCHI_OK) return rc;
    }
    if (s->rows + B > s->capacity) {
        int64_t want = s->capacity * 2 > s->rows + B ? s->capacity * 2 : s->rows + B;
        if (want < 1024) want = 1024;
        int rc = reserve_locked(s, want);
        if (rc != ARCHI_OK) return rc;
    }
    char *rows = (char *)s->data + (size_t)s->rows * s->ld * elt_size(s->dtype);
    int rc = launch_pool_normalize(hidden_dev, hidden_dtype, mask_dev, mask_dtype, B, L, s->dim, rows, s->dtype,
                                   s->ld, s->norm2 + s->rows, s->alive, s->rows, nullptr, out_f32_dev,
                                   reinterpret_cast<cudaStream_t>(stream));
    if (rc != ARCHI_OK) return rc;
    s->rows += B;
    s->epoch++;
    return mark_done(s, reinterpret_cast<cudaStream_t>(stream));
}

// ---- search -------------------------------------------------------------------------------------------
int archi_search(archi_store_t *s, const float *queries, int queries_loc, int nq, int k,
                 const uint32_t *filter_mask_dev, int include_deleted, int path, float *out_scores,
                 int64_t *out_ids, int out_loc, int64_t id_offset, void *stream)
{
    return search_impl(s, queries, queries_loc, nq, k, filter_mask_dev, include_deleted, path, 0, 1.f, 0.f, nullptr,
                       out_scores, out_ids, out_loc, id_offset, stream);
}

int archi_hybrid_search(archi_store_t *s, const float *queries, int queries_loc, int nq, int k, float w_sem,
                        float w_bm25, const float *bm25_dev, const uint32_t *filter_mask_dev, int include_deleted,
                        float *out_scores, int64_t *out_ids, int out_loc, int64_t id_offset, void *stream)
{
    return search_impl(s, queries, queries_loc, nq, k, filter_mask_dev, include_deleted, ARCHI_PATH_STREAM, 1, w_sem,
                       w_bm25, bm25_dev, out_scores, out_ids, out_loc, id_offset, stream);
}

int archi_hybrid_search_terms(archi_store_t *s, const float *queries, int queries_loc, int nq, int k, float w_sem,
                              float w_bm25, const archi_bm25_terms_t *terms, const uint32_t *filter_mask_dev,
                              int include_deleted, float *out_scores, int64_t *out_ids, int out_loc, int64_t id_offset,
                              void *stream, int *out_path)
{
    return hybrid_terms_impl(s, queries, queries_loc, nq, k, w_sem, w_bm25, terms, filter_mask_dev, include_deleted,
                             out_scores, out_ids, out_loc, id_offset, stream, out_path);
}

int archi_bm25_accumulate(const int64_t *post_start_host, const int64_t *post_end_host, int n_terms,
                          const float *idf_host, const int32_t *doc_ids_dev, const int32_t *tfs_dev,
                          const float *doc_len_dev, float avgdl, float k1, float b, float sign, float *out_dev,
                          void *stream)
{
    ARCHI_REQUIRE(n_terms >= 0, "bm25_accumulate: n_terms < 0");
    ARCHI_REQUIRE(n_terms == 0 || (post_start_host && post_end_host && idf_host && out_dev && doc_len_dev),
                  "bm25_accumulate: null argument");
    ARCHI_REQUIRE(avgdl > 0.f, "bm25_accumulate: avgdl must be positive");
    if (n_terms == 0) return ARCHI_OK;
    // launch on the device that owns the output vector, whatever the caller's current device is
    cudaPointerAttributes attr;
    ARCHI_CUDA(cudaPointerGetAttributes(&attr, out_dev));
    ARCHI_REQUIRE(attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged,
                  "bm25_accumulate: out_dev is not device memory");
    ARCHI_DEVICE_GUARD(attr.device);
    for (int t = 0; t < n_terms; ++t) {
        const int64_t b0 = post_start_host[t], b1 = post_end_host[t];
        ARCHI_REQUIRE(b0 >= 0 && b1 >= b0, "bm25_accumulate: bad posting range [%lld, %lld)", (long long)b0, (long long)b1);
        int rc = launch_bm25(doc_ids_dev + b0, tfs_dev + b0, b1 - b0, idf_host[t], doc_len_dev, avgdl, k1, b, sign,
                             out_dev, reinterpret_cast<cudaStream_t>(stream));
        if (rc != ARCHI_OK) return rc;
    }
    return ARCHI_OK;
}

int archi_merge_topk(int device, const float *scores_dev, const int64_t *ids_dev, int n_lists, int nq, int k,
                     int larger_is_better, float *out_scores_dev, int64_t *out_ids_dev, void *stream)
{
    ARCHI_REQUIRE(nq >= 0 && k >= 0, "merge_topk: negative size");
    ARCHI_REQUIRE(nq == 0 || k == 0 || (scores_dev && ids_dev && out_scores_dev && out_ids_dev),
                  "merge_topk: null argument");
    ARCHI_DEVICE_GUARD(device);
    const size_t dense = (size_t)nq * (size_t)k;
    return launch_merge_lists(scores_dev, ids_dev, dense, dense, n_lists, nq, k, larger_is_better, out_scores_dev,
                              out_ids_dev, reinterpret_cast<cudaStream_t>(stream));
}

int archi_merge_topk_strided(int device, const float *scores_dev, const int64_t *ids_dev, int64_t scores_list_stride,
                             int64_t ids_list_stride, int n_lists, int nq, int k, int larger_is_better,
                             float *out_scores_dev, int64_t *out_ids_dev, void *stream)
{
    ARCHI_REQUIRE(nq >= 0 && k >= 0, "merge_topk_strided: negative size");
    ARCHI_REQUIRE(nq == 0 || k == 0 || (scores_dev && ids_dev && out_scores_dev && out_ids_dev),
                  "merge_topk_strided: null argument");
    ARCHI_REQUIRE(scores_list_stride >= (int64_t)nq * k && ids_list_stride >= (int64_t)nq * k,
                  "merge_topk_strided: list strides (%lld, %lld) are shorter than one list (%lld elements)",
                  (long long)scores_list_stride, (long long)ids_list_stride, (long long)nq * k);
    ARCHI_DEVICE_GUARD(device);
    return launch_merge_lists(scores_dev, ids_dev, (size_t)scores_list_stride, (size_t)ids_list_stride, n_lists, nq, k,
                              larger_is_better, out_scores_dev, out_ids_dev, reinterpret_cast<cudaStream_t>(stream));
}

int archi_store_last_stats(archi_store_t *s, archi_search_stats_t *out)
{
    ARCHI_REQUIRE(s && out, "store_last_stats: null argument");
    std::lock_guard<std::mutex> lock(s->mu);
    TensorWorkspace &tw = s->tws;
    if (tw.verdict_pending) {
        ARCHI_DEVICE_GUARD(s->device);
        ARCHI_CUDA(cudaEventSynchronize(tw.verdict_ev));
        int total = 0;
        for (int i = 0; i < tw.verdict_launches; ++i) total += tw.h_verdict[i];
        s->stats.unverified_queries = total;
        int sticky = 0;
        ARCHI_CUDA(cudaMemcpy(&sticky, tw.sticky_dev, sizeof(int), cudaMemcpyDeviceToHost));
        s->stats.unproven_queries = sticky - tw.sticky_fixed;
        tw.verdict_pending = false;
    }
    *out = s->stats;
    return ARCHI_OK;
}

int archi_store_set_timing(archi_store_t *s, int enabled)
{
    ARCHI_REQUIRE(s != nullptr, "store_set_timing: null store");
    std::lock_guard<std::mutex> lock(s->mu);
    s->timing = enabled != 0;
    return ARCHI_OK;
}

}  // extern "C"
