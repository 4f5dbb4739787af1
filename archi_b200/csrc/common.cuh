// common.cuh -- shared declarations for libarchi_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/archi_b200.h"

namespace archi {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;

#define ARCHI_CUDA(call)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            archi::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                             cudaGetErrorString(_e));                                            \
            return ARCHI_ECUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define ARCHI_CHECK_LAUNCH()                                                                     \
    do {                                                                                         \
        archi::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
        ARCHI_CUDA(cudaGetLastError());                                                          \
    } while (0)

#define ARCHI_REQUIRE(cond, ...)                                                                 \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            archi::set_error(__VA_ARGS__);                                                       \
            return ARCHI_EINVAL;                                                                 \
        }                                                                                        \
    } while (0)

// ---------------------------------------------------------------------------------------------
// the store (one row shard resident in one GPU's HBM)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxListK = 128;     // entries a warp-register list holds (4 per lane)
constexpr int kScanThreads = 256;  // 8 warps per CTA
constexpr int kMaxQB = 8;          // queries per corpus pass on the streaming path
constexpr int kTensorMinBatch = 2; // ARCHI_PATH_AUTO: batches at least this large take the tensor path (measured crossover)
constexpr int kTensorMaxBatch = 2048;  // queries per tensor-path launch (16 query tiles)

struct Workspace {
    // per-CTA partial lists of the streaming scan: [grid][kMaxQB][kMaxListK]
    float *part_key = nullptr;
    int *part_id = nullptr;
    int part_grid = 0;
    // per-pass query staging and multi-pass cursors
    float *q_dev = nullptr;        // [q_cap, dim]
    int q_cap = 0;
    float *cursor_key = nullptr;   // [q_cap]
    int *cursor_id = nullptr;      // [q_cap]
    float *out_scores = nullptr;   // [q_cap * k_cap] device staging for host outputs
    int64_t *out_ids = nullptr;
    int64_t out_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // partial lists of the device-driven rescue scan: [passes][grid][kMaxQB][kMaxListK]
    float *resc_key = nullptr;
    int *resc_id = nullptr;
    long long resc_elems = 0;
};

// scratch of the tensor-core path (tensor.cu)
struct TensorWorkspace {
    void *qstage = nullptr;      size_t qstage_bytes = 0;   // staged queries (bf16 / padded fp32)
    void *qinfo = nullptr;       size_t qinfo_bytes = 0;
    uint32_t *thr_g = nullptr;   size_t thr_bytes = 0;      // shared thresholds
    int *unverified = nullptr;   size_t unv_bytes = 0;      // per-query flag + counter
    void *cand = nullptr;        size_t cand_bytes = 0;     // [grid][128][cap] candidates
    int *cand_cnt = nullptr;     size_t cnt_bytes = 0;
    void *aux = nullptr;         size_t aux_bytes = 0;      // [capacity] (a, b) per row
    float *chunkmax = nullptr;   size_t chunkmax_bytes = 0;  // chunk maxima of the probe launch
    uint32_t *progress = nullptr; size_t progress_bytes = 0; // per-CTA tile counters of a long scan (launch tag | tiles issued)
    uint32_t launch_tag = 0;
    float *max_norm2 = nullptr;  // device [2]: max / min |row|^2 over live rows
    float h_max_norm2 = 0.f, h_min_norm2 = 0.f;   // host copies (refreshed with maxnorm_epoch)
    // bf16 shadow of an fp32 store: the coarse pass reads it (half the bytes, full-rate MMA); the
    // exact rescoring still reads the fp32 rows
    void *shadow = nullptr;      size_t shadow_bytes = 0;
    int64_t shadow_rows = 0;     // rows [0, shadow_rows) are converted
    int64_t shadow_reset_epoch = -1;
    int64_t maxnorm_epoch = -1;
    bool maxnorm_all_rows = false;   // the cached norm bounds include tombstoned rows (include_deleted searches)
    int64_t aux_epoch = -1;
    bool aux_alive = false, aux_had_filter = false;
    // queries whose exactness proof failed: device list [unv_cap] filled by tc_select_kernel, re-scanned by
    // the device-driven rescue launches; the count travels to pinned host memory behind `verdict_ev`
    int *unv_list = nullptr;     size_t unv_list_bytes = 0;
    int *unv_count = nullptr;    // device counter of the current launch (lives behind the flags of `unverified`)
    int *h_verdict = nullptr;    // pinned [64]: unproven queries of each tensor launch of the last search
    int verdict_max_sel[64];     // ... and how many of them the device-side rescue could take
    int verdict_launches = 0;
    cudaEvent_t verdict_ev = nullptr;
    bool verdict_pending = false;
    int *sticky_dev = nullptr;   // device [1]: queries ever written as id -1 / NaN because the rescue list was full
    int sticky_fixed = 0;        // ... of which the host re-scanned before returning (host outputs)
    // TMA descriptors are rebuilt only when what they describe changes
    unsigned char tmap_q[128] __attribute__((aligned(64)));
    unsigned char tmap_c[128] __attribute__((aligned(64)));
    const void *tmq_base = nullptr;  long long tmq_rows = -1;  int tmq_ld = 0, tmq_f32 = -1;
    const void *tmc_base = nullptr;  long long tmc_rows = -1;  int tmc_ld = 0, tmc_f32 = -1, tmc_box = 0;
};

// scratch of the posting-list hybrid search (hybrid.cu)
struct HybridWorkspace {
    unsigned long long *acc = nullptr;  size_t acc_bytes = 0;   // [slots][capacity] BM25 sums, 32.32 fixed point, all zero between calls
    bool acc_dirty = false;
    float *part_key = nullptr;          size_t part_key_bytes = 0;
    int *part_id = nullptr;             size_t part_id_bytes = 0;
    float *bias = nullptr;              size_t bias_bytes = 0;    // dense-vector fallback: [capacity] BM25 per row
    float *dense_scores = nullptr;      // dense top-k of the batch
    int64_t *dense_ids = nullptr;       size_t dense_bytes = 0;
    // the sparse chain (posting walks + gathers, latency bound) runs on `side` while the dense search streams the corpus
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

// the query terms of a hybrid search: host arrays indexed by (query, term) occurrence + the device posting lists
struct HybridTerms {
    const int64_t *post_start, *post_end;
    const float *idf;
    const int32_t *doc_ids_dev, *tfs_dev;
    const float *doc_len_dev;
    float avgdl, k1, b, sign;
};
constexpr int kHybMaxPairsHost = 96;   // = kHybMaxPairs / kHybMaxSlots of hybrid.cu
constexpr int kHybMaxSlotsHost = 16;

}  // namespace archi

struct archi_store {
    int device = 0;
    int dim = 0;
    int ld = 0;               // row stride in elements (dim rounded up to a 16-byte multiple)
    int metric = 0;
    int dtype = 0;            // ARCHI_F32 | ARCHI_BF16
    int64_t capacity = 0;
    int64_t rows = 0;         // appended so far (next row id)
    int64_t deleted = 0;
    void *data = nullptr;     // [capacity, ld] storage dtype
    float *norm2 = nullptr;   // [capacity] |row|^2 of the stored values
    uint32_t *alive = nullptr;// [capacity/32] tombstone bitmask (1 = live)
    int sm_count = 0;
    int timing = 0;
    archi_search_stats_t stats{};
    archi::Workspace ws;
    archi::TensorWorkspace tws;
    archi::HybridWorkspace hws;
    int64_t epoch = 0;        // bumped whenever rows / tombstones change (invalidates cached aux)
    int64_t reset_epoch = 0;  // bumped when existing rows are discarded or moved (reset / load)
    // stream ordering between calls on one handle (the buffers and scratch are shared): recorded at the end of
    // every call that enqueues work, waited on by the next call when it uses another stream
    cudaEvent_t order_ev = nullptr;
    cudaStream_t order_stream = nullptr;
    std::mutex mu;
};

namespace archi {

inline size_t elt_size(int dtype) { return dtype == ARCHI_BF16 ? 2 : 4; }
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// launchers implemented in the .cu files (all enqueue on `st`, no host sync)
// ---------------------------------------------------------------------------------------------
struct ScanArgs {
    const void *corpus;
    int dtype;
    int64_t n;
    int dim, ld;
    int metric;
    const float *queries;     // device [nqb, dim]
    int nqb;                  // 1..kMaxQB
    int k;                    // 1..kMaxListK (entries to produce this pass)
    const float *norm2;       // [n]
    const uint32_t *alive;    // may be null
    const uint32_t *filter;   // may be null
    int hybrid;
    const float *bias;        // [nqb][bias_stride] or null
    int64_t bias_stride;
    float w_sem, w_bias;
    const float *cursor_key;  // [nqb] or null: only rows sorting strictly after the cursor
    const int *cursor_id;
};

// Streaming scan: fills ws.part_* with per-CTA lists; returns grid in *grid_out.
int launch_scan(archi_store *s, const ScanArgs &a, cudaStream_t st, int *grid_out);
// Merges `grid` per-CTA lists into the final rows [col0, col0+k) of the outputs, converts keys to
// the reference's score convention and (optionally) records the cursor for the next pass.
int launch_scan_finalize(archi_store *s, const ScanArgs &a, int grid, int k_total, int col0,
                         float *out_scores, int64_t *out_ids, int64_t id_offset,
                         float *cursor_key_out, int *cursor_id_out, cudaStream_t st);

int launch_append(archi_store *s, const void *src_dev, int src_dtype, int64_t first_row, int64_t n,
                  cudaStream_t st);
int launch_delete_rows(archi_store *s, const long long *rows_dev, int64_t n, int *changed_dev,
                       cudaStream_t st);
int launch_read_rows(archi_store *s, int64_t first_row, int64_t n, float *out_dev, cudaStream_t st);
int launch_pool_normalize(const void *hidden, int hidden_dtype, const void *mask, int mask_dtype,
                          int B, int L, int H, void *store_rows, int store_dtype, int store_ld,
                          float *store_norm2, uint32_t *alive, long long first_row, void *out_bf16,
                          float *out_f32, cudaStream_t st);
int launch_bm25(const int32_t *doc_ids, const int32_t *tfs, int64_t n_post, float idf,
                const float *doc_len, float avgdl, float k1, float b, float sign, float *out,
                cudaStream_t st);
int tensor_path_supported(const archi_store *s, int k);
// Enqueues the whole tensor-path search of one batch (<= kTensorMaxBatch queries) on `st`, no host sync:
// coarse launches, select + exact rescoring + proof, then the device-driven rescue of up to *max_sel_out
// unproven queries.  The number of unproven queries lands in s->tws.h_verdict[0] behind s->tws.verdict_ev;
// queries beyond max_sel (only possible for batches > 256) are returned as id -1 / score NaN and flagged in
// s->tws.unverified.
int launch_tensor_search(archi_store *s, const float *q_dev, int nq, int k, const uint32_t *filter, int include_deleted,
                         float *out_scores, int64_t *out_ids, int64_t id_offset, cudaStream_t st,
                         int *max_sel_out, double *coarse_ms);
int launch_rescue(archi_store *s, const ScanArgs &a, const int *qsel_dev, const int *nsel_dev, int max_sel,
                  float *out_scores, int64_t *out_ids, int64_t id_offset, cudaStream_t st);
void free_tensor_workspace(TensorWorkspace &w);
void free_hybrid_workspace(HybridWorkspace &w);
// One round of the sparse chain of the posting-list hybrid search on stream `st`: n_slots <= 16 queries (slot0 = index
// of the first one in the call), n_pairs <= 96 (query, term) pairs starting at pair0 of `t`, pair_slot[j] = query slot
// of pair j.  Leaves the round's partial lists (exact combined scores of the rows matching a term) in the workspace.
int launch_hybrid_sparse_round(archi_store *s, const float *q_dev, int n_slots, int k, int slot0, float w_sem, float w_bm25,
                               float sign, const HybridTerms &t, int pair0, int n_pairs, const int *pair_slot,
                               const uint32_t *filter, int include_deleted, cudaStream_t st, int *out_cps);
int hybrid_ensure_part_lists(archi_store *s, int nq, cudaStream_t st);
// dense_* = the queries' ordinary top-k (device) + the round's partial lists -> the k best combined scores
int launch_hybrid_merge(archi_store *s, int n_slots, int k, int slot0, int cps, const float *dense_scores,
                        const int64_t *dense_ids, float w_sem, float *out_scores, int64_t *out_ids, int64_t id_offset,
                        cudaStream_t st);
int launch_merge_lists(const float *scores, const int64_t *ids, size_t scores_list_stride, size_t ids_list_stride,
                       int n_lists, int nq, int k, int larger_is_better, float *out_scores, int64_t *out_ids,
                       cudaStream_t st);

}  // namespace archi
