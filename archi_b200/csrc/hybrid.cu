// hybrid.cu -- hybrid BM25 + dense top-k without a dense per-row BM25 vector.
//
// Replaces hybrid_search's SQL (postgres_vectorstore.py:435-457):
//     combined = (1.0 - (emb <op> q)) * w_sem + COALESCE(bm25, 0) * w_bm25   ORDER BY combined DESC LIMIT k
// Rows without a lexical match have combined = w_sem * semantic, so (w_sem > 0, bm25 >= 0)
//     exact fused top-k  =  top-k of ( dense top-k over ALL rows  U  exact combined score of S_q )
// where S_q = rows that match at least one query term: a row outside S_q that is not in the dense top-k is
// preceded by k rows whose combined score is at least their own dense score, ties included (lower id first).
// Per batch of queries, on two streams (the sparse chain is a few short latency-bound kernels; the dense search
// streams the corpus beside it):
//   side stream
//   1. hyb_scatter_kernel   walks the posting lists of the query terms and accumulates BM25 per (query, row) into
//                           a persistent, all-zero accumulator (64-bit fixed point: the sum does not depend on the
//                           order of the atomics);
//   2. hyb_score_kernel     walks the postings again, 32 per warp: the lane that swaps a row's BM25 sum out of the
//                           accumulator (atomicExch back to zero: the accumulator is clean again, and a row listed
//                           under several terms has one owner) scores it -- coalesced row loads four rows at a time,
//                           exact fp32 dot, combined score, warp-register top-k, block merge -> partial lists;
//   caller's stream
//   3. the dense top-k of every query -- the ordinary search (tensor-core path for >= 2 queries);
//   4. (after the join) hyb_merge_kernel: partial lists + dense list (converted to combined scores, copies of rows
//                           that are listed in S_q dropped by id) -> the k best (combined desc, id asc).
// Work per query is O(postings of its terms) + one dense search instead of O(rows) extra traffic and a memset.
// Queries whose terms match a large part of the corpus (stop-word-like terms) keep the dense-vector scan.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "topk.cuh"

namespace archi {

constexpr int kHybMaxPairs = kHybMaxPairsHost;    // (query, term) pairs per round (kernel-parameter space)
constexpr int kHybMaxSlots = kHybMaxSlotsHost;    // queries per round (accumulator planes)
constexpr int kHybCpsMax = 592;     // most CTAs scoring the candidates of one query (cps is chosen per round)
constexpr int kHybPartStride = 4096;    // partial-list entries per query (cps * k never exceeds it)
constexpr float kFix = 4294967296.0f;           // 2^32: BM25 contributions are accumulated in 32.32 fixed point
constexpr float kUnfix = 2.3283064365386963e-10f;

struct HybPair {
    long long start, end;   // posting range of the term
    float idf;
    int slot;               // query slot of this round
};

struct HybRound {
    int n_pairs, n_slots;
    long long total;                       // postings of all pairs
    long long prefix[kHybMaxPairs + 1];    // exclusive prefix sums of the pairs' posting counts
    HybPair pairs[kHybMaxPairs];
    long long slot_p0[kHybMaxSlots + 1];   // postings [slot_p0[s], slot_p0[s + 1]) of the round belong to query slot s
};

struct HybBm25 {
    const int32_t *doc_ids;
    const int32_t *tfs;
    const float *doc_len;
    float avgdl, k1, b;
};

__device__ __forceinline__ int hyb_pair_of(const long long *prefix, int n_pairs, long long p)
{
    int lo = 0, hi = n_pairs;             // largest j with prefix[j] <= p
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] <= p) lo = mid;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) hyb_scatter_kernel(const HybRound r, const HybBm25 bm, unsigned long long *acc,
                                                          long long acc_stride)
{
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < r.total; p += (long long)gridDim.x * blockDim.x) {
        const int j = hyb_pair_of(r.prefix, r.n_pairs, p);
        const long long off = r.pairs[j].start + (p - r.prefix[j]);
        const int doc = bm.doc_ids[off];
        const float tf = (float)bm.tfs[off];
        const float denom = tf + bm.k1 * (1.0f - bm.b + bm.b * bm.doc_len[doc] / bm.avgdl);
        const float c = r.pairs[j].idf * tf * (bm.k1 + 1.0f) / denom;       // same expression as bm25_term_kernel
        atomicAdd(acc + (size_t)r.pairs[j].slot * acc_stride + doc, __float2ull_rn(c * kFix));
    }
}

struct HybScore {
    const void *corpus;
    int dtype, dim, ld, metric;
    const float *norm2;
    const uint32_t *alive;
    const uint32_t *filter;
    const float *queries;        // [n_slots, dim] of this round
    float w_sem, w_bm25, sign;
    int k;
    const int32_t *doc_ids;      // the posting lists
    unsigned long long *acc;     // [slots][acc_stride] BM25 sums of the round (32.32 fixed point), zeroed row by row here
    long long acc_stride;
    float *part_key;             // [n_slots][kHybPartStride]: cps lists of k entries per slot
    int *part_id;
    int cps;                     // CTAs per query slot
    int pb;                      // postings a warp claims per step (8, 16 or 32: fewer when there are warps to spare, so
                                 // that the rows of a step are gathered in one or two rounds of four)
};

// Merge `nlists` sorted lists of 32*M entries staged in shared memory into `res` (same walk as scan.cu's block merge).
template <int M>
__device__ __forceinline__ void hyb_merge_staged(const float *skey, const int *sid, int nlists, int k, int lane, WarpTopK<M> &res)
{
    float tk = -CUDART_INF_F;
    int ti = INT_MAX;
    for (int l = 0; l < nlists; ++l) {
#pragma unroll 1
        for (int c = 0; c < M; ++c) {
            const float ek = skey[l * 32 * M + c * 32 + lane];
            const int ei = sid[l * 32 * M + c * 32 + lane];
            unsigned cand = __ballot_sync(kFull, better(ek, ei, tk, ti));
            if (cand == 0) break;
            while (cand) {
                const int src = __ffs(cand) - 1;
                cand &= cand - 1;
                const float nk = __shfl_sync(kFull, ek, src);
                const int ni = __shfl_sync(kFull, ei, src);
                if (res.insert(nk, ni, k, lane)) res.threshold(k, tk, ti);
            }
        }
    }
}

template <int M>
__global__ void __launch_bounds__(256) hyb_score_kernel(const HybScore p, const HybRound r)
{
    extern __shared__ __align__(16) unsigned char hsmem[];
    float *sq = reinterpret_cast<float *>(hsmem);            // [ld] the slot's query, zero padded
    __shared__ float s_qrn;
    const int slot = blockIdx.x / p.cps, cta = blockIdx.x % p.cps;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int WARPS = 8;
    for (int e = tid; e < p.ld; e += 256) sq[e] = e < p.dim ? p.queries[(size_t)slot * p.dim + e] : 0.f;
    __syncthreads();
    if (warp == 0) {
        float ss = 0.f;
        for (int e = lane; e < p.ld; e += 32) ss = fmaf(sq[e], sq[e], ss);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) ss += __shfl_xor_sync(kFull, ss, d);
        if (lane == 0) s_qrn = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
    }
    __syncthreads();
    const float qrn = s_qrn;
    const int vec = p.dtype == ARCHI_BF16 ? 8 : 4;
    const int nvec = p.ld / vec;
    const size_t row_bytes = (size_t)p.ld * (p.dtype == ARCHI_BF16 ? 2 : 4);
    unsigned long long *acc_slot = p.acc + (size_t)slot * p.acc_stride;
    WarpTopK<M> list;
    list.init();
    // A warp takes 32 postings of the slot at a time, one per lane.  The lane that swaps a row's BM25 sum out of the
    // accumulator (atomicExch back to zero: the plane is clean for the next call, and a row listed under several
    // terms has exactly one owner) scores the row; the warp then works through the owned rows four at a time with
    // the loads of all four in flight (random 16-byte-vector gathers are latency bound).
    constexpr int RW = 4;
    const long long p_end = r.slot_p0[slot + 1];
    for (long long pb = r.slot_p0[slot] + (long long)(cta * WARPS + warp) * p.pb; pb < p_end; pb += (long long)p.cps * WARPS * p.pb) {
        const long long pp = pb + lane;
        int my_doc = 0;
        float my_bm = 0.f, my_n2 = 1.f;
        bool own = false;
        if (lane < p.pb && pp < p_end) {
            const int j = hyb_pair_of(r.prefix, r.n_pairs, pp);
            my_doc = p.doc_ids[r.pairs[j].start + (pp - r.prefix[j])];
            const unsigned long long sum = atomicExch(acc_slot + my_doc, 0ull);
            own = sum != 0ull;
            if (own && p.alive) own = (p.alive[my_doc >> 5] >> (my_doc & 31)) & 1u;
            if (own && p.filter) own = (p.filter[my_doc >> 5] >> (my_doc & 31)) & 1u;
            if (own) {
                my_bm = __ull2float_rn(sum) * kUnfix;
                if (p.metric == ARCHI_COSINE) my_n2 = p.norm2[my_doc];
            }
        }
        unsigned owners = __ballot_sync(kFull, own);
        while (owners) {
            int doc[RW];
            bool ok[RW];
            const uint4 *rp[RW];
            float acc[RW], bm[RW], n2[RW];
#pragma unroll
            for (int j = 0; j < RW; ++j) {
                ok[j] = owners != 0u;
                const int src = ok[j] ? __ffs(owners) - 1 : 0;
                if (ok[j]) owners &= owners - 1;
                doc[j] = __shfl_sync(kFull, my_doc, src);
                bm[j] = __shfl_sync(kFull, my_bm, src);
                n2[j] = __shfl_sync(kFull, my_n2, src);
                rp[j] = reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(p.corpus) + (size_t)doc[j] * row_bytes);
                acc[j] = 0.f;
            }
            for (int v = lane; v < nvec; v += 32) {
                uint4 d[RW];
#pragma unroll
                for (int j = 0; j < RW; ++j) d[j] = ok[j] ? __ldg(rp[j] + v) : make_uint4(0u, 0u, 0u, 0u);
                const float *qq = sq + v * vec;
#pragma unroll
                for (int j = 0; j < RW; ++j) {
                    float x[8];
                    if (p.dtype == ARCHI_BF16) {
                        x[0] = __uint_as_float(d[j].x << 16); x[1] = __uint_as_float(d[j].x & 0xffff0000u);
                        x[2] = __uint_as_float(d[j].y << 16); x[3] = __uint_as_float(d[j].y & 0xffff0000u);
                        x[4] = __uint_as_float(d[j].z << 16); x[5] = __uint_as_float(d[j].z & 0xffff0000u);
                        x[6] = __uint_as_float(d[j].w << 16); x[7] = __uint_as_float(d[j].w & 0xffff0000u);
                    } else {
                        x[0] = __uint_as_float(d[j].x); x[1] = __uint_as_float(d[j].y);
                        x[2] = __uint_as_float(d[j].z); x[3] = __uint_as_float(d[j].w);
                        x[4] = x[5] = x[6] = x[7] = 0.f;
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        if (e < vec) {
                            if (p.metric == ARCHI_L2) {
                                const float t = x[e] - qq[e];
                                acc[j] = fmaf(t, t, acc[j]);
                            } else {
                                acc[j] = fmaf(x[e], qq[e], acc[j]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < RW; ++j) {
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) acc[j] += __shfl_xor_sync(kFull, acc[j], d);
            }
#pragma unroll
            for (int j = 0; j < RW; ++j) {
                if (!ok[j]) continue;                                // warp-uniform
                float sem;                                           // the scan kernel's hybrid key, term by term
                if (p.metric == ARCHI_L2) {
                    sem = 1.0f - sqrtf(acc[j]);
                } else if (p.metric == ARCHI_COSINE) {
                    sem = fminf(1.0f, fmaxf(-1.0f, acc[j] * qrn * (n2[j] > 0.f ? 1.0f / sqrtf(n2[j]) : 0.f)));
                } else {
                    sem = 1.0f + acc[j];
                }
                const float key = fmaf(sem, p.w_sem, p.sign * bm[j] * p.w_bm25);
                list.insert(key, doc[j], p.k, lane);
            }
        }
    }
    // block merge of the 8 warp lists
    constexpr int LEN = 32 * M;
    __syncthreads();
    float *skey = reinterpret_cast<float *>(hsmem);
    int *sid = reinterpret_cast<int *>(hsmem) + WARPS * LEN;
#pragma unroll
    for (int s = 0; s < M; ++s) {
        skey[warp * LEN + s * 32 + lane] = list.key[s];
        sid[warp * LEN + s * 32 + lane] = list.id[s];
    }
    __syncthreads();
    if (warp == 0) {
        WarpTopK<M> res;
        res.init();
        hyb_merge_staged<M>(skey, sid, WARPS, p.k, lane, res);
        float *ok_ = p.part_key + (size_t)slot * kHybPartStride + (size_t)cta * p.k;
        int *oi_ = p.part_id + (size_t)slot * kHybPartStride + (size_t)cta * p.k;
#pragma unroll
        for (int s = 0; s < M; ++s) {
            const int rank = s * 32 + lane;
            if (rank < p.k) {
                ok_[rank] = res.key[s];
                oi_[rank] = res.id[s];
            }
        }
    }
}

// One CTA per query (all loads at once, then one warp merges): the partial lists of S_q (rows matching a query term, exact combined scores) and the dense
// top-k of the ordinary search -> the k best.  The dense list comes from a search that ran CONCURRENTLY with the
// sparse chain, so it is converted here (combined = w_sem * semantic, :441) and de-duplicated here: a dense row that
// is also in S_q carries a combined score >= its dense-only score.  If its S_q twin is among the k best of S_q, the
// dense copy is dropped by id; if it is not, k rows of S_q precede the twin and therefore the copy, which then
// cannot enter the list (`insert` admits only entries better than the current k-th).
constexpr int kHybMergeThreads = 256;

template <int M>
__global__ void __launch_bounds__(kHybMergeThreads) hyb_merge_kernel(const float *dense_scores, const long long *dense_ids, int metric,
                                                                     float w_sem, const float *part_key, const int *part_id, int cps,
                                                                     int part_stride, int k, long long id_offset, float *out_scores,
                                                                     long long *out_ids)
{
    extern __shared__ __align__(16) unsigned char msm[];
    const int slot = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int n = cps * k;                                  // <= part_stride
    float *skey = reinterpret_cast<float *>(msm);           // [n] then [k] dense
    int *sid = reinterpret_cast<int *>(skey + n + k);
    // every load of the partial lists and of the dense list is issued at once by the whole CTA (a single warp walking
    // them 32 at a time pays one L2 round trip per step); warp 0 then merges from shared memory
    for (int i = tid; i < n; i += kHybMergeThreads) {
        skey[i] = part_key[(size_t)slot * part_stride + i];
        sid[i] = part_id[(size_t)slot * part_stride + i];
    }
    for (int rnk = tid; rnk < k; rnk += kHybMergeThreads) {
        const long long gid = dense_ids[(size_t)slot * k + rnk];
        float dk = -CUDART_INF_F;
        int di = INT_MAX;
        if (gid >= 0) {
            const float sc = dense_scores[(size_t)slot * k + rnk];
            const float sem = metric == ARCHI_COSINE ? sc : 1.0f - sc;     // semantic = 1.0 - (emb <op> q), :441
            dk = sem * w_sem;
            di = (int)(gid - id_offset);
        }
        skey[n + rnk] = dk;
        sid[n + rnk] = di;
    }
    __syncthreads();
    // eight warps reduce a strided eighth of the partial lists each, warp 0 merges their lists
    constexpr int MW = kHybMergeThreads / 32, LEN = 32 * M;
    float *wkey = reinterpret_cast<float *>(sid + n + k);   // [MW][LEN]
    int *wid = reinterpret_cast<int *>(wkey + MW * LEN);
    {
        const int warp = tid >> 5;
        WarpTopK<M> mine;
        mine.init();
        float tk = -CUDART_INF_F;   // the list's k-th entry: candidates that do not beat it are rejected by one ballot
        int ti = INT_MAX;
        for (int i0 = warp * 32; i0 < n; i0 += MW * 32) {
            const int i = i0 + lane;
            const float ek = i < n ? skey[i] : -CUDART_INF_F;
            const int ei = i < n ? sid[i] : INT_MAX;
            unsigned cand = __ballot_sync(kFull, ei != INT_MAX && better(ek, ei, tk, ti));
            while (cand) {
                const int src = __ffs(cand) - 1;
                cand &= cand - 1;
                if (mine.insert(__shfl_sync(kFull, ek, src), __shfl_sync(kFull, ei, src), k, lane)) mine.threshold(k, tk, ti);
            }
        }
#pragma unroll
        for (int s = 0; s < M; ++s) {
            wkey[warp * LEN + s * 32 + lane] = mine.key[s];
            wid[warp * LEN + s * 32 + lane] = mine.id[s];
        }
    }
    __syncthreads();
    if (tid >= 32) return;
    WarpTopK<M> res;
    res.init();
    hyb_merge_staged<M>(wkey, wid, MW, k, lane, res);
    float tk = -CUDART_INF_F;
    int ti = INT_MAX;
    res.threshold(k, tk, ti);
    // the dense list: M entries per lane; copies of listed S_q rows are dropped (against the list as it stands BEFORE
    // any dense entry goes in)
    float dk[M];
    int di[M];
#pragma unroll
    for (int s = 0; s < M; ++s) {
        const int rnk = s * 32 + lane;
        dk[s] = rnk < k ? skey[n + rnk] : -CUDART_INF_F;
        di[s] = rnk < k ? sid[n + rnk] : INT_MAX;
    }
#pragma unroll
    for (int s = 0; s < M; ++s) {
        for (int src = 0; src < 32; ++src) {
            if (s * 32 + src >= k) break;
            const int id = __shfl_sync(kFull, di[s], src);
            bool hit = false;
#pragma unroll
            for (int u = 0; u < M; ++u) hit = hit || (res.id[u] == id);
            const bool dup = id != INT_MAX && __any_sync(kFull, hit);
            if (dup && lane == src) di[s] = INT_MAX;
        }
    }
#pragma unroll
    for (int s = 0; s < M; ++s) {
        unsigned cand = __ballot_sync(kFull, di[s] != INT_MAX && better(dk[s], di[s], tk, ti));
        while (cand) {
            const int src = __ffs(cand) - 1;
            cand &= cand - 1;
            if (res.insert(__shfl_sync(kFull, dk[s], src), __shfl_sync(kFull, di[s], src), k, lane)) res.threshold(k, tk, ti);
        }
    }
#pragma unroll
    for (int s = 0; s < M; ++s) {
        const int rank = s * 32 + lane;
        if (rank < k) {
            const bool empty = res.id[s] == INT_MAX;
            out_scores[(size_t)slot * k + rank] = empty ? CUDART_NAN_F : res.key[s];
            out_ids[(size_t)slot * k + rank] = empty ? -1ll : (long long)res.id[s] + id_offset;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side: one round of <= kHybMaxSlots queries whose terms are all "sparse"
// ---------------------------------------------------------------------------------------------
template <typename T>
static int hyb_ensure(T **ptr, size_t *cap, size_t need, bool zero, cudaStream_t st)
{
    if (*ptr && *cap >= need) return ARCHI_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    ARCHI_CUDA(cudaMalloc(ptr, need));
    if (zero) ARCHI_CUDA(cudaMemsetAsync(*ptr, 0, need, st));
    *cap = need;
    return ARCHI_OK;
}

int launch_hybrid_sparse_round(archi_store *s, const float *q_dev /* [n_slots, dim] */, int n_slots, int k, int slot0,
                               float w_sem, float w_bm25, float sign, const HybridTerms &t, int pair0, int n_pairs,
                               const int *pair_slot, const uint32_t *filter, int include_deleted, cudaStream_t st,
                               int *out_cps)
{
    ARCHI_REQUIRE(n_slots >= 1 && n_slots <= kHybMaxSlots && n_pairs <= kHybMaxPairs, "hybrid: round too large");
    ARCHI_REQUIRE(k >= 1 && k <= kMaxListK, "hybrid: k=%d out of range for the sparse path", k);
    HybridWorkspace &w = s->hws;
    HybRound r;
    r.n_pairs = n_pairs;
    r.n_slots = n_slots;
    long long per_slot[kHybMaxSlots] = {0};
    long long run = 0;
    for (int j = 0; j < n_pairs; ++j) {
        r.prefix[j] = run;
        r.pairs[j].start = t.post_start[pair0 + j];
        r.pairs[j].end = t.post_end[pair0 + j];
        r.pairs[j].idf = t.idf[pair0 + j];
        r.pairs[j].slot = pair_slot[j];
        const long long n = r.pairs[j].end - r.pairs[j].start;
        run += n;
        per_slot[pair_slot[j]] += n;
    }
    r.prefix[n_pairs] = run;
    r.total = run;
    long long off = 0;
    for (int sl = 0; sl < n_slots; ++sl) {
        r.slot_p0[sl] = off;
        off += per_slot[sl];
    }
    r.slot_p0[n_slots] = off;

    int rc;
    // the accumulator planes are all-zero between calls (hyb_collect_kernel restores every entry it read)
    const size_t acc_need = (size_t)kHybMaxSlots * (size_t)s->capacity * sizeof(unsigned long long);
    if (!w.acc || w.acc_bytes < acc_need || w.acc_dirty) {
        if (w.acc && w.acc_bytes >= acc_need) ARCHI_CUDA(cudaMemsetAsync(w.acc, 0, w.acc_bytes, st));
        else if ((rc = hyb_ensure(&w.acc, &w.acc_bytes, acc_need, true, st)) != ARCHI_OK) return rc;
        w.acc_dirty = false;
    }
    // CTAs per query: up to four per SM over the whole round (the kernel is a chain of dependent latencies: posting
    // -> accumulator swap -> row gather), at least one block of 32 postings per warp, and few enough partial lists
    // (cps * k entries) for the merge
    long long max_slot = 1;
    for (int sl = 0; sl < n_slots; ++sl) max_slot = per_slot[sl] > max_slot ? per_slot[sl] : max_slot;
    int cps = (4 * s->sm_count) / n_slots;
    if (cps > kHybCpsMax) cps = kHybCpsMax;
    if ((long long)cps * k > kHybPartStride) cps = kHybPartStride / k;
    if ((long long)cps * 8 * 8 > max_slot) cps = (int)(max_slot / (8 * 8));
    if (cps < 1) cps = 1;
    *out_cps = cps;
    // postings per warp step: 32 when every warp has several steps of work anyway, fewer when warps would idle
    int pb = 32;
    while (pb > 8 && (long long)cps * 8 * pb > max_slot) pb >>= 1;

    HybBm25 bm;
    bm.doc_ids = t.doc_ids_dev;
    bm.tfs = t.tfs_dev;
    bm.doc_len = t.doc_len_dev;
    bm.avgdl = t.avgdl;
    bm.k1 = t.k1;
    bm.b = t.b;
    const long long acc_stride = s->capacity;
    w.acc_dirty = true;          // until hyb_score_kernel (which swaps every touched entry back to zero) is enqueued
    if (r.total > 0) {
        long long blocks = (r.total + 255) / 256;
        if (blocks > (long long)s->sm_count * 8) blocks = (long long)s->sm_count * 8;
        hyb_scatter_kernel<<<(unsigned)blocks, 256, 0, st>>>(r, bm, w.acc, acc_stride);
        ARCHI_CHECK_LAUNCH();
    }
    HybScore sp;
    sp.corpus = s->data;
    sp.dtype = s->dtype;
    sp.dim = s->dim;
    sp.ld = s->ld;
    sp.metric = s->metric;
    sp.norm2 = s->norm2;
    sp.alive = include_deleted ? nullptr : s->alive;
    sp.filter = filter;
    sp.queries = q_dev;
    sp.w_sem = w_sem;
    sp.w_bm25 = w_bm25;
    sp.sign = sign;
    sp.k = k;
    sp.doc_ids = t.doc_ids_dev;
    sp.acc = w.acc;
    sp.acc_stride = acc_stride;
    sp.part_key = w.part_key + (size_t)slot0 * kHybPartStride;
    sp.part_id = w.part_id + (size_t)slot0 * kHybPartStride;
    sp.cps = cps;
    sp.pb = pb;
    const int M = k <= 32 ? 1 : 4;
    const size_t q_bytes = (size_t)s->ld * sizeof(float);
    const size_t m_bytes = (size_t)8 * 32 * M * 8;
    const size_t smem = q_bytes > m_bytes ? q_bytes : m_bytes;
    if (M == 1) {
        if (smem > 40 * 1024) ARCHI_CUDA(cudaFuncSetAttribute((const void *)hyb_score_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hyb_score_kernel<1><<<n_slots * cps, 256, smem, st>>>(sp, r);
    } else {
        if (smem > 40 * 1024) ARCHI_CUDA(cudaFuncSetAttribute((const void *)hyb_score_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hyb_score_kernel<4><<<n_slots * cps, 256, smem, st>>>(sp, r);
    }
    ARCHI_CHECK_LAUNCH();
    w.acc_dirty = false;
    return ARCHI_OK;
}

// The partial lists of a call hold kHybPartStride entries per query (all rounds of the call are in flight at once on
// the side stream, the merges run after the join).
int hybrid_ensure_part_lists(archi_store *s, int nq, cudaStream_t st)
{
    HybridWorkspace &w = s->hws;
    int rc;
    const size_t need = (size_t)(nq > kHybMaxSlots ? nq : kHybMaxSlots) * kHybPartStride;
    if ((rc = hyb_ensure(&w.part_key, &w.part_key_bytes, need * sizeof(float), false, st)) != ARCHI_OK) return rc;
    if ((rc = hyb_ensure(&w.part_id, &w.part_id_bytes, need * sizeof(int), false, st)) != ARCHI_OK) return rc;
    return ARCHI_OK;
}

// Merge of one round: dense top-k (scores / global ids of the ordinary search) + the round's partial lists.
int launch_hybrid_merge(archi_store *s, int n_slots, int k, int slot0, int cps, const float *dense_scores,
                        const int64_t *dense_ids, float w_sem, float *out_scores, int64_t *out_ids, int64_t id_offset,
                        cudaStream_t st)
{
    HybridWorkspace &w = s->hws;
    const float *pk = w.part_key + (size_t)slot0 * kHybPartStride;
    const int *pi = w.part_id + (size_t)slot0 * kHybPartStride;
    const size_t msm = (size_t)(cps * k + k) * 8 + (size_t)(kHybMergeThreads / 32) * 32 * (k <= 32 ? 1 : 4) * 8;   // < 48 KB
    if (k <= 32)
        hyb_merge_kernel<1><<<n_slots, kHybMergeThreads, msm, st>>>(dense_scores, reinterpret_cast<const long long *>(dense_ids), s->metric, w_sem,
                                                    pk, pi, cps, kHybPartStride, k, id_offset, out_scores,
                                                    reinterpret_cast<long long *>(out_ids));
    else
        hyb_merge_kernel<4><<<n_slots, kHybMergeThreads, msm, st>>>(dense_scores, reinterpret_cast<const long long *>(dense_ids), s->metric, w_sem,
                                                    pk, pi, cps, kHybPartStride, k, id_offset, out_scores,
                                                    reinterpret_cast<long long *>(out_ids));
    ARCHI_CHECK_LAUNCH();
    return ARCHI_OK;
}

void free_hybrid_workspace(HybridWorkspace &w)
{
    void *ptrs[] = {w.acc, w.part_key, w.part_id, w.bias, w.dense_scores, w.dense_ids};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (w.side) cudaStreamDestroy(w.side);
    if (w.ev_fork) cudaEventDestroy(w.ev_fork);
    if (w.ev_join) cudaEventDestroy(w.ev_join);
    w = HybridWorkspace();
}

}  // namespace archi
