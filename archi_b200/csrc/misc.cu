// misc.cu -- small kernels around the hot path: row append (convert + |row|^2), tombstones,
// row read-back, BM25 posting-list accumulation, and the shard-merge of per-GPU k-lists.
#include "common.cuh"
#include "merge.cuh"
#include "topk.cuh"

namespace archi {

// ---------------------------------------------------------------------------------------------
// append: INSERT ... %s::vector (postgres_vectorstore.py:168-180; manager.py:414-422).
// One warp per row: convert to the storage dtype, zero the row padding, record |stored row|^2,
// set the live bit.
// ---------------------------------------------------------------------------------------------
template <typename SRC>
__device__ __forceinline__ float load_as_f32(const SRC *p, size_t i);
template <>
__device__ __forceinline__ float load_as_f32<float>(const float *p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_f32<__nv_bfloat16>(const __nv_bfloat16 *p, size_t i)
{
    return __bfloat162float(p[i]);
}

template <typename SRC, typename DST>
__global__ void __launch_bounds__(256) append_kernel(const SRC *src, DST *dst, float *norm2,
                                                     uint32_t *alive, long long first_row,
                                                     long long n, int dim, int ld)
{
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    const long long row = first_row + r;
    float ss = 0.f;
    for (int e = lane; e < ld; e += 32) {
        const float v = e < dim ? load_as_f32<SRC>(src, (size_t)r * dim + e) : 0.f;
        float stored;
        if constexpr (sizeof(DST) == 2) {
            const __nv_bfloat16 o = __float2bfloat16_rn(v);
            dst[(size_t)row * ld + e] = o;
            stored = __bfloat162float(o);
        } else {
            dst[(size_t)row * ld + e] = v;
            stored = v;
        }
        ss = fmaf(stored, stored, ss);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) ss += __shfl_xor_sync(kFull, ss, d);
    if (lane == 0) {
        norm2[row] = ss;
        atomicOr(&alive[row >> 5], 1u << (row & 31));
    }
}

int launch_append(archi_store *s, const void *src_dev, int src_dtype, int64_t first_row, int64_t n,
                  cudaStream_t st)
{
    if (n == 0) return ARCHI_OK;
    const unsigned grid = (unsigned)((n + 7) / 8);
    if (src_dtype == ARCHI_F32 && s->dtype == ARCHI_F32)
        append_kernel<float, float><<<grid, 256, 0, st>>>((const float *)src_dev, (float *)s->data, s->norm2,
                                                          s->alive, first_row, n, s->dim, s->ld);
    else if (src_dtype == ARCHI_F32 && s->dtype == ARCHI_BF16)
        append_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float *)src_dev, (__nv_bfloat16 *)s->data,
                                                                  s->norm2, s->alive, first_row, n, s->dim, s->ld);
    else if (src_dtype == ARCHI_BF16 && s->dtype == ARCHI_F32)
        append_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)src_dev, (float *)s->data,
                                                                  s->norm2, s->alive, first_row, n, s->dim, s->ld);
    else if (src_dtype == ARCHI_BF16 && s->dtype == ARCHI_BF16)
        append_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(
            (const __nv_bfloat16 *)src_dev, (__nv_bfloat16 *)s->data, s->norm2, s->alive, first_row, n, s->dim, s->ld);
    else {
        set_error("append: source dtype must be f32 or bf16");
        return ARCHI_EINVAL;
    }
    ARCHI_CHECK_LAUNCH();
    return ARCHI_OK;
}

// ---------------------------------------------------------------------------------------------
// tombstones: DELETE FROM document_chunks ... (postgres_vectorstore.py:493-535)
// ---------------------------------------------------------------------------------------------
__global__ void delete_rows_kernel(uint32_t *alive, const long long *rows, long long n, long long limit,
                                   int *changed)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long row = rows[i];
    if (row < 0 || row >= limit) return;
    const uint32_t bit = 1u << (row & 31);
    const uint32_t old = atomicAnd(&alive[row >> 5], ~bit);
    if (old & bit) atomicAdd(changed, 1);
}

int launch_delete_rows(archi_store *s, const long long *rows_dev, int64_t n, int *changed_dev, cudaStream_t st)
{
    if (n == 0) return ARCHI_OK;
    delete_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->alive, rows_dev, n, s->rows, changed_dev);
    ARCHI_CHECK_LAUNCH();
    return ARCHI_OK;
}

// ---------------------------------------------------------------------------------------------
// read rows back as fp32 (snapshot / tests)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void read_rows_kernel(const T *data, long long first_row, long long n, int dim, int ld, float *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * dim) return;
    const long long r = i / dim;
    const int e = (int)(i - r * dim);
    out[i] = load_as_f32<T>(data, (size_t)(first_row + r) * ld + e);
}

int launch_read_rows(archi_store *s, int64_t first_row, int64_t n, float *out_dev, cudaStream_t st)
{
    if (n == 0) return ARCHI_OK;
    const long long tot = (long long)n * s->dim;
    const unsigned grid = (unsigned)((tot + 255) / 256);
    if (s->dtype == ARCHI_BF16)
        read_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)s->data, first_row, n, s->dim, s->ld, out_dev);
    else
        read_rows_kernel<float><<<grid, 256, 0, st>>>((const float *)s->data, first_row, n, s->dim, s->ld, out_dev);
    ARCHI_CHECK_LAUNCH();
    return ARCHI_OK;
}

// ---------------------------------------------------------------------------------------------
// BM25 over one term's posting list (pg_textsearch `<@>` [external], postgres_vectorstore.py:433).
// Doc ids are unique inside one posting list and terms are launched back to back on one stream,
// so plain read-modify-write is race-free and the summation order is deterministic.
// ---------------------------------------------------------------------------------------------
__global__ void bm25_term_kernel(const int32_t *doc_ids, const int32_t *tfs, long long n_post, float idf,
                                 const float *doc_len, float avgdl, float k1, float b, float sign, float *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_post) return;
    const int doc = doc_ids[i];
    const float tf = (float)tfs[i];
    const float denom = tf + k1 * (1.0f - b + b * doc_len[doc] / avgdl);
    out[doc] += sign * idf * tf * (k1 + 1.0f) / denom;
}

int launch_bm25(const int32_t *doc_ids, const int32_t *tfs, int64_t n_post, float idf, const float *doc_len,
                float avgdl, float k1, float b, float sign, float *out, cudaStream_t st)
{
    if (n_post == 0) return ARCHI_OK;
    bm25_term_kernel<<<(unsigned)((n_post + 255) / 256), 256, 0, st>>>(doc_ids, tfs, n_post, idf, doc_len, avgdl,
                                                                       k1, b, sign, out);
    ARCHI_CHECK_LAUNCH();
    return ARCHI_OK;
}

// ---------------------------------------------------------------------------------------------
// shard merge: n_lists sorted k-lists per query -> one k-list (after the NCCL allgather).
// One warp per query; lane l walks list l (n_lists <= 32); each step a warp arg-best picks the
// next output.  Works for any k.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) merge_lists_kernel(const float *scores, const long long *ids,
                                                          size_t s_stride, size_t i_stride, int n_lists, int nq,
                                                          int k, int larger, float *out_scores, long long *out_ids)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q >= nq) return;
    merge_query_lists(scores, ids, s_stride, i_stride, n_lists, nq, q, k, larger, out_scores, out_ids, lane);
}

int launch_merge_lists(const float *scores, const int64_t *ids, size_t scores_list_stride, size_t ids_list_stride,
                       int n_lists, int nq, int k, int larger_is_better, float *out_scores, int64_t *out_ids,
                       cudaStream_t st)
{
    ARCHI_REQUIRE(n_lists >= 1 && n_lists <= 32, "merge_topk: n_lists=%d must be in [1, 32]", n_lists);
    if (nq == 0 || k == 0) return ARCHI_OK;
    merge_lists_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, st>>>(scores, (const long long *)ids, scores_list_stride,
                                                                 ids_list_stride, n_lists, nq, k, larger_is_better,
                                                                 out_scores, (long long *)out_ids);
    ARCHI_CHECK_LAUNCH();
    return ARCHI_OK;
}

}  // namespace archi
