// tensor.cu -- the large-batch search path: TMA-fed tcgen05/TMEM coarse scorer with a fused
// threshold filter, followed by an exact fp32 rescoring of the surviving candidates.
//
// Replaces the same reference statement as scan.cu (postgres_vectorstore.py:317-332 + :361) for
// query batches where the scan is a real dense contraction.  No N x Q score matrix is written:
//   1. tc_prep_kernel     stages the queries for the tensor pipe (bf16 copy for bf16 stores, padded
//                         fp32 copy read as tf32 for fp32 stores) and computes, per query, a bound
//                         eps on |coarse key - exact key|.
//   2. tc_aux_kernel      per-row epilogue constants (a, b): key = dot * a + b  (cosine: a = 1/|c|;
//                         l2: a = 2, b = -|c|^2; ip: a = 1); masked / deleted rows get b = -inf.
//   3. tc_coarse_kernel   persistent warp-specialised GEMM.  One CTA per SM (or a CTA pair with
//                         cta_group::2, M = 256): warp 0 = TMA producer (128B-swizzled tiles of 128
//                         queries and 256 corpus rows through an mbarrier ring; short rows keep the
//                         query tile resident), warp 1 = tcgen05.mma issuer (fp32 accumulators
//                         double-buffered in TMEM, 2 x 256 columns), warps 2-9 = two epilogue groups:
//                         tcgen05.ld the scores, one query per thread, and keep only keys above that
//                         query's threshold in a per-(CTA, group, query) candidate buffer; when a
//                         buffer fills, the owning warp selects its k' best by a ballot/REDUX
//                         bisection on the key bits and raises the threshold (shared between CTAs
//                         through an atomicMax word per query).  A first PROBE launch of the same
//                         kernel over ~1/12 of the rows only records chunk maxima, from which
//                         tc_maxima_threshold_kernel derives the thresholds the main scan starts with.
//   4. tc_select_kernel   per query: union of the candidate buffers -> k' best coarse keys -> exact
//                         fp32 rescoring against the stored rows -> top-k, plus the proof that no
//                         rejected row can belong to the exact top-k:  e_k > T + eps.  Queries that
//                         fail the proof are flagged and re-run on the exact streaming path.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "topk.cuh"

namespace archi {
namespace tc {

constexpr int BM = 128;            // queries per tile (MMA M, TMEM lanes)
constexpr int BN = 256;            // corpus rows per tile (MMA N, TMEM columns)
constexpr int A_BYTES = BM * 128;  // one 128-byte swizzle row per query per K chunk
// 1-CTA mode: a stage holds the whole 256-row corpus tile (48 KB, 4 stages); CTA-pair mode
// (cta_group::2): each CTA holds its 128 queries and HALF of the corpus tile (32 KB, 6 stages)
template <bool TWO> struct Geo {
    static constexpr int B_ROWS = TWO ? BN / 2 : BN;
    static constexpr int B_BYTES = B_ROWS * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = TWO ? 6 : 4;
};
constexpr int RING_BYTES = 4 * (A_BYTES + BN * 128);   // = 6 * (A_BYTES + BN/2 * 128) = 196608
constexpr int MAX_STAGES = 6;
constexpr int NTHREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2..5 / 6..9 epilogue groups 0 / 1
constexpr int EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;
constexpr int MAX_QT = 16;         // query tiles per launch (2048 queries)
constexpr int AUX_BYTES = EPI_WARPS * BN * 8;       // a private (a, b) tile copy per epilogue warp
constexpr int SMEM_BYTES = 1024 + RING_BYTES + AUX_BYTES + 256;

// shared threshold word: 0 = "no threshold published yet"
__device__ __forceinline__ float thr_from_word(uint32_t u) { return u ? funmap(u) : -CUDART_INF_F; }

struct QInfo {
    float rn_q;   // 1 / |q|
    float qn2;    // |q|^2
    float eps;    // bound on |coarse key - exact key| in coarse-key space
    float pad;
};

// ---------------------------------------------------------------------------------------------
// 1. query staging + error bound
// ---------------------------------------------------------------------------------------------
struct PrepParams {
    const float *queries;  // [nq, dim]
    int nq, dim, ldq;      // ldq: row stride of the staged copy, in elements
    int tf32;              // 1: staged copy is fp32 (read as tf32), 0: bf16
    int metric;
    void *qstage;          // [nq_pad, ldq]
    QInfo *qinfo;          // [nq]
    uint32_t *thr_g;       // [nq] shared thresholds (mapped), reset to 0
    int *unverified;       // [nq]
    int *n_unverified;     // counter of unverified queries, reset here
    const float *max_norm2;// [1] max |row|^2 over the store
    float corpus_rel_err;  // 2^-9 when the coarse pass reads a bf16 shadow of fp32 rows, else 0
    int cos_raw_unnorm;    // cosine scored as a raw dot over rows that are only approximately unit length
    float norm_dev;        // max |1/|c| - 1| over the store (cos_raw_unnorm)
};

__global__ void __launch_bounds__(128) tc_prep_kernel(const PrepParams p)
{
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ float s_a[4], s_b[4];
    const float *src = p.queries + (size_t)q * p.dim;
    float n2 = 0.f, e2 = 0.f;
    for (int e = tid; e < p.ldq; e += 128) {
        const float v = e < p.dim ? src[e] : 0.f;
        n2 = fmaf(v, v, n2);
        if (p.tf32) {
            reinterpret_cast<float *>(p.qstage)[(size_t)q * p.ldq + e] = v;
        } else {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            reinterpret_cast<__nv_bfloat16 *>(p.qstage)[(size_t)q * p.ldq + e] = h;
            const float d = v - __bfloat162float(h);
            e2 = fmaf(d, d, e2);
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        n2 += __shfl_xor_sync(kFull, n2, d);
        e2 += __shfl_xor_sync(kFull, e2, d);
    }
    if (lane == 0) {
        s_a[warp] = n2;
        s_b[warp] = e2;
    }
    __syncthreads();
    if (tid == 0) {
        n2 = s_a[0] + s_a[1] + s_a[2] + s_a[3];
        e2 = s_b[0] + s_b[1] + s_b[2] + s_b[3];
        const float qn = sqrtf(n2);
        // Error of the coarse dot q~.c~ against the exact q.c, term by term, as `unit * |c|` (q~, c~ = the operands
        // the tensor cores read; |c| = the norm of the row the exact key is defined on):
        //   (q~ - q).c~     <= |q - bf16(q)| |c~|, and a row rounded to bf16 is at most (1 + 2^-9) longer than the
        //                      row it rounds -> sqrt(e2) * 1.002 (covers the fp32 rounding of sqrt and of e2 too);
        //                      kind::tf32: both operands truncated to 10 mantissa bits -> 2^-9 (1 + 2^-11) |q|;
        //   q.(c~ - c)      <= |q| 2^-9 |c| when the rows read are a bf16 shadow of fp32 rows (0 for bf16 stores:
        //                      the stored row IS the row), 1.0005 covers the shadow's fp32 normalisation;
        //   accumulation    <= dim * 2^-22 |q| |c| (fp32 accumulate, exact products).
        float unit = p.tf32 ? 1.0005f * 0.001953125f * qn : sqrtf(e2) * 1.002f;
        unit += p.corpus_rel_err * 1.0005f * qn;
        unit += (float)p.dim * 2.4e-7f * qn;
        const float maxn = sqrtf(*p.max_norm2);
        float eps;
        if (p.metric == ARCHI_COSINE && p.cos_raw_unnorm) eps = (unit + qn * p.norm_dev * 1.001f) * maxn;  // key = dot
        else if (p.metric == ARCHI_COSINE) eps = unit;            // key = dot / |c|
        else if (p.metric == ARCHI_IP) eps = unit * maxn;         // key = dot
        else eps = 2.f * unit * maxn + 1e-6f * maxn * maxn;       // key = 2 dot - |c|^2
        QInfo qi;
        qi.rn_q = n2 > 0.f ? 1.0f / qn : 0.f;
        qi.qn2 = n2;
        qi.eps = eps;
        qi.pad = 0.f;
        p.qinfo[q] = qi;
        p.thr_g[q] = 0u;
        p.unverified[q] = 0;
        if (q == 0) *p.n_unverified = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// 2. per-row epilogue constants
// ---------------------------------------------------------------------------------------------
__global__ void tc_aux_kernel(float2 *aux, const float *norm2, const uint32_t *alive, const uint32_t *filter,
                              long long n, int metric)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool ok = true;
    if (alive) ok = ok && ((alive[i >> 5] >> (i & 31)) & 1u);
    if (filter) ok = ok && ((filter[i >> 5] >> (i & 31)) & 1u);
    const float n2 = norm2[i];
    float2 ab;
    if (metric == ARCHI_COSINE) {
        ab = make_float2(n2 > 0.f ? 1.0f / sqrtf(n2) : 0.f, 0.f);
        if (!(n2 > 0.f)) ok = false;
    } else if (metric == ARCHI_IP) {
        ab = make_float2(1.f, 0.f);
    } else {
        ab = make_float2(2.f, -n2);
    }
    if (!ok) ab = make_float2(0.f, -CUDART_INF_F);
    aux[i] = ab;
}

// fp32 rows -> bf16 shadow rows (row strides ld_src / ld_dst elements, padding zeroed)
// With `norm2` (cosine stores) the shadow rows are scaled to unit length, so that the coarse key of
// the cosine metric is the raw dot product.
__global__ void tc_shadow_kernel(const float *src, __nv_bfloat16 *dst, long long first, long long n, int dim,
                                 int ld_src, int ld_dst, const float *norm2)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * ld_dst) return;
    const long long r = first + i / ld_dst;
    const int e = (int)(i % ld_dst);
    float v = e < dim ? src[r * ld_src + e] : 0.f;
    if (norm2) {
        const float n2 = norm2[r];
        v = n2 > 0.f ? v * (1.0f / sqrtf(n2)) : 0.f;
    }
    dst[r * ld_dst + e] = __float2bfloat16_rn(v);
}

// out[0] = max |row|^2, out[1] = min |row|^2 over live rows (non-negative floats order like their bits)
__global__ void tc_maxnorm_kernel(const float *norm2, const uint32_t *alive, long long n, float *out)
{
    float m = 0.f, mn = CUDART_INF_F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (alive && !((alive[i >> 5] >> (i & 31)) & 1u)) continue;
        const float v = norm2[i];
        m = fmaxf(m, v);
        mn = fminf(mn, v);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(kFull, m, d));
        mn = fminf(mn, __shfl_xor_sync(kFull, mn, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(reinterpret_cast<unsigned int *>(out), __float_as_uint(m));
        atomicMin(reinterpret_cast<unsigned int *>(out + 1), __float_as_uint(mn));
    }
}

// ---------------------------------------------------------------------------------------------
// 3. the coarse scorer
// ---------------------------------------------------------------------------------------------
struct CoarseParams {
    long long n;          // corpus rows
    int n_ctiles;         // ceil(n / BN)
    int kchunks;          // ceil(ld / elements per 128 B)
    int kelems;           // elements per 128-byte chunk (64 bf16, 32 fp32)
    int nq;               // queries in this launch
    int qt_count;         // query tiles
    int ngroups;          // gridDim.x / qt_count
    int kprime;           // candidates kept per compaction
    int cap;              // entries per candidate buffer (C)
    int trigger;          // a buffer holding at least this many entries is compacted after the tile
    int tile_begin, tile_end;  // corpus tiles [tile_begin, tile_end) are scanned by this launch
    int resume;           // 1: continue from the candidate counts / thresholds left by a previous launch
    int a_res, a_stages;  // resident query tile: on/off, corpus-chunk stages behind it
    int split;            // 1: both epilogue groups drain every tile, half the columns each
    int maxima_only;      // 1: probe launch -- record the maximum live key of every 32-column chunk, store nothing
    float *chunkmax;      // [grid * 2][cm_slots][BM] chunk maxima of a probe launch
    int cm_slots;         // maxima kept per (CTA, group, query)
    int debug;            // profiling aid (ARCHI_TC_DEBUG): 1 = skip the MMAs, 2 = skip the epilogue math
    int exit_cap;         // buffers larger than this are compacted before the CTA exits
    int throttle_win;     // > 0: a CTA issues tile u only when every CTA of its corpus group has issued tile u - win
    uint32_t launch_tag;  // distinguishes this launch's progress words from stale ones (12 bits)
    uint32_t *progress;   // [grid] (launch_tag << 20) | tiles issued
    const float2 *aux;    // [n] (a, b) per row -- aux mode only
    const uint32_t *alive;   // raw mode: tombstone bitmask (may be null)
    const uint32_t *filter;  // raw mode: per-search filter bitmask (may be null)
    uint2 *cand;          // [grid][BM][cap]  (key bits, row id)
    int *cand_cnt;        // [grid][BM]
    uint32_t *thr_g;      // [nq]
};

__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr)
{
    // K-major, 128-byte swizzle: 8-row x 128 B atoms, 1024 B apart (SBO); LBO unused (1);
    // descriptor version 1 (sm_100); layout type 2 = SWIZZLE_128B.
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Warp-collective: keep the kprime best entries of lane `owner`'s buffer (n entries, n <= 32*SLOTS),
// return the new count and the new threshold (a key T with exactly/at least kprime entries >= T).
template <int SLOTS>
__device__ __forceinline__ void compact_buffer(uint2 *buf, int n, int kprime, int lane, int &new_cnt, float &new_thr)
{
    uint32_t uk[SLOTS];
    uint32_t id[SLOTS];
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) {
        const int i = t * 32 + lane;
        uk[t] = 0u;
        id[t] = 0u;
        if (i < n) {
            const uint2 e = buf[i];
            uk[t] = fmap(__uint_as_float(e.x));
            id[t] = e.y;
        }
    }
    // Bisection on the mapped key bits for a T with kprime <= count(uk >= T) <= kprime + slack.
    // Only the bits below the highest bit in which the largest and smallest key differ are searched.
    uint32_t kmax = 0u, kmin = 0xffffffffu;
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) {
        if ((t * 32 + lane) < n) {
            kmax = max(kmax, uk[t]);
            kmin = min(kmin, uk[t]);
        }
    }
    kmax = __reduce_max_sync(kFull, kmax);
    kmin = __reduce_min_sync(kFull, kmin);
    const uint32_t diff = kmax ^ kmin;
    uint32_t T = kmin;                                   // all keys equal (or n <= kprime): keep everything
    if (diff != 0u && n > kprime) {
        const int hb = 31 - __clz(diff);
        T = hb == 31 ? 0u : (kmax & ~((2u << hb) - 1u));  // common prefix of every key
        const int slack = kprime >> 2;
#pragma unroll 1
        for (int b = hb; b >= 0; --b) {
            const uint32_t trial = T | (1u << b);
            int c = 0;
#pragma unroll
            for (int t = 0; t < SLOTS; ++t) c += (uk[t] >= trial) ? 1 : 0;
            c = __reduce_add_sync(kFull, c);
            if (c >= kprime) {
                T = trial;
                if (c <= kprime + slack) break;           // tight enough: stop early
            }
        }
    }
    // keep everything above T and only as many entries equal to T as are needed to reach kprime
    // (dropping a row whose coarse key equals the new threshold is covered by the proof)
    int c_gt = 0;
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) c_gt += (uk[t] > T) ? 1 : 0;
    c_gt = __reduce_add_sync(kFull, c_gt);
    int need_eq = kprime - c_gt;                          // <= 0: no entry equal to T is needed
    int base = 0;
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) {
        const bool valid = (t * 32 + lane) < n;
        const bool gt = valid && uk[t] > T;
        const bool eq = valid && uk[t] == T;
        const unsigned m_eq = __ballot_sync(kFull, eq);
        const bool keep = gt || (eq && __popc(m_eq & ((1u << lane) - 1u)) < need_eq);
        const unsigned m = __ballot_sync(kFull, keep);
        if (keep) {
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            buf[pos] = make_uint2(__float_as_uint(funmap(uk[t])), id[t]);
        }
        base += __popc(m);
        const int took_eq = __popc(m_eq);
        need_eq -= took_eq < need_eq ? took_eq : need_eq;
    }
    new_cnt = base;
    new_thr = funmap(T);
}

__device__ __forceinline__ void compact_dispatch(uint2 *buf, int n, int kprime, int cap, int lane, int &nc, float &nt)
{
    (void)cap;
    if (n <= 64) compact_buffer<2>(buf, n, kprime, lane, nc, nt);
    else if (n <= 128) compact_buffer<4>(buf, n, kprime, lane, nc, nt);
    else if (n <= 256) compact_buffer<8>(buf, n, kprime, lane, nc, nt);
    else if (n <= 512) compact_buffer<16>(buf, n, kprime, lane, nc, nt);
    else compact_buffer<32>(buf, n, kprime, lane, nc, nt);
}

// TWO = CTA-pair mode: a cluster of two CTAs (an SM pair) scores 256 queries (128 per CTA) against
// one 256-row corpus tile with tcgen05.mma.cta_group::2 (M = 256).  Each CTA loads its own query tile
// and HALF of the corpus tile; the leader CTA (cluster rank 0) issues the MMAs for the pair, and each
// CTA's epilogue drains its own TMEM.  Per SM this halves the corpus bytes written to and read from
// shared memory, which is what bounds the 1-CTA kernel.
// RAW = the coarse key is the raw dot product (inner product; cosine over unit-norm rows): no per-column
// constants, masked rows are cleared with one bitmask word per 32 columns, and a chunk whose maximum
// does not beat the thresholds of any lane is skipped after a 3-input max tree.
template <bool TF32, bool TWO, bool RAW>
__device__ __forceinline__ void coarse_body(const CUtensorMap &tmap_q, const CUtensorMap &tmap_c, const CoarseParams &p)
{
    // Two shared-memory plans.  Streaming (default): a ring of STAGES x (query chunk | corpus chunk).
    // Resident queries (p.a_res, short rows): the CTA's whole query tile is loaded once and stays at the
    // bottom of the ring area; the ring behind it only carries corpus chunks.  That halves the L2 -> SM
    // traffic and the shared-memory writes of a D <= 512 scan.
    constexpr int B_BYTES = Geo<TWO>::B_BYTES;
    const bool ares = p.a_res != 0;
    const int nst = ares ? p.a_stages : Geo<TWO>::STAGES;
    const int stage_bytes = ares ? B_BYTES : Geo<TWO>::STAGE_BYTES;
    const int b_off = ares ? 0 : A_BYTES;                    // corpus chunk inside a stage
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw = ptx::smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;           // 1024-byte aligned (128B swizzle atoms)
    const uint32_t ring = ares ? base + p.kchunks * A_BYTES : base;
    unsigned char *base_ptr = smem_dyn + (base - raw);
    // layout: [STAGES x (A | B)] [aux: EPI_WARPS x BN float2] [barriers, tmem ptr: 256 B]
    float2 *s_aux = reinterpret_cast<float2 *>(base_ptr + RING_BYTES);
    const uint32_t bar0 = base + RING_BYTES + AUX_BYTES;
    const uint32_t full_bar = bar0, empty_bar = bar0 + 8 * MAX_STAGES;
    const uint32_t tfull_bar = bar0 + 16 * MAX_STAGES, tempty_bar = tfull_bar + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(base_ptr + RING_BYTES + AUX_BYTES + 16 * MAX_STAGES + 32);
    const uint32_t afull_bar = bar0 + 16 * MAX_STAGES + 40;  // resident query tile landed

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x % p.qt_count;                 // pair mode: qt_count is even, the pair is (2j, 2j+1)
    const int group = blockIdx.x / p.qt_count;
    const uint32_t cta_rank = TWO ? ptx::cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0u;

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tmap_q);
        ptx::prefetch_tensormap(&tmap_c);
        for (int s = 0; s < MAX_STAGES; ++s) {
            ptx::mbar_init(full_bar + 8 * s, 1);
            ptx::mbar_init(empty_bar + 8 * s, 1);
        }
        ptx::mbar_init(afull_bar, 1);
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(tfull_bar + 8 * a, 1);
            // one arrival per epilogue warp that drains the buffer (pair mode: both CTAs' warps release the leader)
            ptx::mbar_init(tempty_bar + 8 * a, (TWO ? 8 : 4) * (p.split ? 2 : 1));
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        if (TWO) {
            ptx::tmem_alloc_2sm(ptx::smem_u32(tmem_slot), TMEM_COLS);
            ptx::tmem_relinquish_2sm();
        } else {
            ptx::tmem_alloc(ptx::smem_u32(tmem_slot), TMEM_COLS);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    if (TWO) ptx::cluster_sync();      // the peer's barriers must be initialised before anything signals them
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // addresses of the LEADER's barriers as seen from this CTA
    const uint32_t full_bar_leader = TWO ? ptx::mapa(full_bar, 0) : full_bar;
    const uint32_t tempty_bar_leader = TWO ? ptx::mapa(tempty_bar, 0) : tempty_bar;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            if (ares && p.tile_begin + group < p.tile_end) {
                // the query tile, once (pair mode: the leader's barrier collects both CTAs' tiles)
                if (TWO) {
                    if (leader) ptx::mbar_arrive_expect_tx(afull_bar, 2 * p.kchunks * A_BYTES);
                    const uint32_t afull_leader = ptx::mapa(afull_bar, 0);
                    for (int kc = 0; kc < p.kchunks; ++kc)
                        ptx::tma_load_2d_2sm(base + kc * A_BYTES, &tmap_q, afull_leader, kc * p.kelems, qt * BM);
                } else {
                    ptx::mbar_arrive_expect_tx(afull_bar, p.kchunks * A_BYTES);
                    for (int kc = 0; kc < p.kchunks; ++kc)
                        ptx::tma_load_2d(base + kc * A_BYTES, &tmap_q, afull_bar, kc * p.kelems, qt * BM);
                }
            }
            const uint32_t tx_bytes = (uint32_t)((TWO ? 2 : 1) * stage_bytes);
            // Long scans: the CTAs of a corpus group read the same tiles and rely on L2 for all but the first read.
            // They run at the same tensor-bound pace, but a drift of a fraction of a percent over thousands of
            // tiles exceeds what L2 holds per group, and the stragglers then fetch every tile from HBM again (ncu,
            // 10M x 768: 2.2x the algorithmic bytes).  A soft throttle keeps the group inside a window of a few
            // tiles: every CTA publishes the number of tiles it has issued, and every fourth tile it waits -- bounded,
            // it never blocks for good -- until the slowest CTA of its group is within the window.
            volatile uint32_t *prog = p.throttle_win > 0 ? p.progress + (size_t)group * p.qt_count : nullptr;
            const uint32_t tag = p.launch_tag << 20;
            int u_t = 0;
            for (int ct = p.tile_begin + group; ct < p.tile_end; ct += p.ngroups, ++u_t) {
                if (prog) {
                    prog[qt] = tag | (uint32_t)u_t;
                    if (u_t >= p.throttle_win && (u_t & 3) == 0) {
                        for (int spins = 0; spins < 400; ++spins) {
                            uint32_t slowest = 0xfffffu;
                            for (int j = 0; j < p.qt_count; ++j) {
                                const uint32_t v = prog[j];
                                const uint32_t pj = (v >> 20) == p.launch_tag ? (v & 0xfffffu) : 0u;   // not started yet: 0
                                slowest = pj < slowest ? pj : slowest;
                            }
                            if ((uint32_t)u_t <= slowest + (uint32_t)p.throttle_win) break;
                            __nanosleep(200);
                        }
                    }
                }
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    ptx::mbar_wait(empty_bar + 8 * s, ph ^ 1u);
                    const uint32_t sa = ring + s * stage_bytes;
                    if (TWO) {
                        // the leader's barrier collects the bytes of both CTAs
                        if (leader) ptx::mbar_arrive_expect_tx(full_bar + 8 * s, tx_bytes);
                        if (!ares) ptx::tma_load_2d_2sm(sa, &tmap_q, full_bar_leader + 8 * s, kc * p.kelems, qt * BM);
                        ptx::tma_load_2d_2sm(sa + b_off, &tmap_c, full_bar_leader + 8 * s, kc * p.kelems,
                                             ct * BN + (int)cta_rank * (BN / 2));
                    } else {
                        ptx::mbar_arrive_expect_tx(full_bar + 8 * s, tx_bytes);
                        if (!ares) ptx::tma_load_2d(sa, &tmap_q, full_bar + 8 * s, kc * p.kelems, qt * BM);
                        ptx::tma_load_2d(sa + b_off, &tmap_c, full_bar + 8 * s, kc * p.kelems, ct * BN);
                    }
                    if (++s == nst) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
            if (prog) prog[qt] = tag | 0xfffffu;           // done: never hold the others back
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0 && leader) {
            // instruction descriptor: D fp32, A/B bf16 (1) or tf32 (2), both K-major, N = 256,
            // M = 128 (one CTA) or 256 (CTA pair)
            const uint32_t fmt = TF32 ? 2u : 1u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)((TWO ? 2 * BM : BM) >> 4) << 24);
            int s = 0, u = 0;
            uint32_t ph = 0;
            for (int ct = p.tile_begin + group; ct < p.tile_end; ct += p.ngroups, ++u) {
                const int acc = u & 1;
                const uint32_t aph = (uint32_t)(u >> 1) & 1u;
                ptx::mbar_wait(tempty_bar + 8 * acc, aph ^ 1u);
                if (ares && u == 0) ptx::mbar_wait(afull_bar, 0u);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    ptx::mbar_wait(full_bar + 8 * s, ph);
                    ptx::tc_fence_after();
                    const uint32_t sa = ring + s * stage_bytes;
                    const uint64_t adesc = smem_desc_sw128(ares ? base + kc * A_BYTES : sa);
                    const uint64_t bdesc = smem_desc_sw128(sa + b_off);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        if (p.debug & 1) break;
                        // +32 bytes along K inside the swizzle atom = +2 in the (addr >> 4) field
                        const uint32_t accum = (kc | k4) ? 1u : 0u;
                        if (TWO) {
                            if (TF32) ptx::mma_tf32_2sm(tmem_d, adesc + 2 * k4, bdesc + 2 * k4, idesc, accum);
                            else ptx::mma_bf16_2sm(tmem_d, adesc + 2 * k4, bdesc + 2 * k4, idesc, accum);
                        } else {
                            if (TF32) ptx::mma_tf32(tmem_d, adesc + 2 * k4, bdesc + 2 * k4, idesc, accum);
                            else ptx::mma_bf16(tmem_d, adesc + 2 * k4, bdesc + 2 * k4, idesc, accum);
                        }
                    }
                    // smem stage reusable (in both CTAs) once these MMAs retire
                    if (TWO) ptx::tc_commit_2sm(empty_bar + 8 * s, 3);
                    else ptx::tc_commit(empty_bar + 8 * s);
                    if (++s == nst) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                // accumulator ready for the epilogue (of both CTAs in pair mode)
                if (TWO) ptx::tc_commit_2sm(tfull_bar + 8 * acc, 3);
                else ptx::tc_commit(tfull_bar + 8 * acc);
            }
        }
    } else {
        // ================= epilogue: one query per thread =================
        // Two groups of four warps; group g drains accumulator g (tiles u = g, g+2, ...), so each
        // group has two MMA tile periods per tile.  Within a group, warp (warp & 3) owns that TMEM
        // lane quadrant and thread `lane` owns query tq = quad*32 + lane.  Candidate bookkeeping is
        // per (CTA, group, query): the two groups behave like two virtual CTAs.
        const int grp = (warp - 2) >> 2;
        const int quad = warp & 3;
        const int tq = quad * 32 + lane;
        const int q = qt * BM + tq;
        const bool active = q < p.nq;
        const size_t vcta = (size_t)blockIdx.x * 2 + grp;
        uint2 *buf = p.cand + (vcta * BM + tq) * p.cap;
        float2 *aux_w = s_aux + (warp - 2) * BN;         // this warp's private copy of the tile's (a, b)
        float thr = active ? -CUDART_INF_F : CUDART_INF_F;
        int cnt = (p.resume && active) ? p.cand_cnt[vcta * BM + tq] : 0;
        if (p.resume && active) {
            // a resumed launch starts from the shared threshold published by tc_threshold_kernel:
            // drop (lane-parallel, no selection needed) everything the new threshold already excludes
            thr = thr_from_word(__ldcg(p.thr_g + q));
            int w = 0;
            for (int i0 = 0; i0 < cnt; i0 += 8) {
                uint2 e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) e[j] = i0 + j < cnt ? buf[i0 + j] : make_uint2(0xff800000u, 0u);  // -inf
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (i0 + j < cnt && __uint_as_float(e[j].x) >= thr) buf[w++] = e[j];
            }
            cnt = w;
        }
        float *cmx = p.chunkmax + vcta * p.cm_slots * BM + tq;     // [slot][query]: coalesced per warp
        int mslot = 0;
        if (p.maxima_only && active)
            for (int i = 0; i < p.cm_slots; ++i) cmx[i * BM] = -CUDART_INF_F;
        // split == 0: group g drains accumulator g (tiles u = g, g+2, ...), all 8 chunks of the tile.
        // split == 1: both groups drain every tile, group g taking chunks 4g .. 4g+3, so a buffer is
        //             handed back after half an epilogue and the MMA warp tolerates T_epi <= 2 T_mma.
        const int ustep = p.split ? 1 : 2;
        const int c_begin = p.split ? grp * (BN / 64) : 0, c_end = p.split ? c_begin + BN / 64 : BN / 32;
        uint32_t thr_word = active ? __ldcg(p.thr_g + q) : 0u;
        int u = p.split ? 0 : grp;
        for (int ct = p.tile_begin + group + u * p.ngroups; ct < p.tile_end; ct += ustep * p.ngroups, u += ustep) {
            const uint32_t aph = (uint32_t)(u >> 1) & 1u;
            const int acc = u & 1;
            // this tile's per-row constants, fetched while the MMAs run (warp-private: no CTA barrier)
            uint32_t tile_word = 0u;                 // raw mode: lane j < 8 holds the live-row bits of chunk j
            if (RAW) {
                if (lane < BN / 32) {
                    const long long r0 = (long long)ct * BN + lane * 32;
                    if (r0 < p.n) {
                        tile_word = r0 + 32 <= p.n ? 0xffffffffu : ((1u << (int)(p.n - r0)) - 1u);
                        if (p.alive) tile_word &= __ldg(p.alive + (r0 >> 5));
                        if (p.filter) tile_word &= __ldg(p.filter + (r0 >> 5));
                    }
                }
            } else {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < BN / 32; ++i) {
                    const long long r = (long long)ct * BN + i * 32 + lane;
                    aux_w[i * 32 + lane] = r < p.n ? __ldg(p.aux + r) : make_float2(0.f, -CUDART_INF_F);
                }
            }
            // shared threshold: use the word fetched during the previous tile and fetch the next one
            // now, so that the L2 round trip never sits between "accumulator full" and the first compare
            if (active) {
                thr = fmaxf(thr, thr_from_word(thr_word));
                thr_word = __ldcg(p.thr_g + q);
            }
            __syncwarp();

            ptx::mbar_wait(tfull_bar + 8 * acc, aph);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
            const uint32_t row_base = (uint32_t)ct * BN;

            // One 32-column chunk in two passes so that the loads / FMAs of all columns overlap:
            //  (1) keys and a 32-bit admission mask (plain code, no ordering constraints);
            //  (2) the few admitted columns are appended.  cnt <= cap - BN holds at every tile start
            //      (compaction trigger), so a tile cannot overflow the buffer.
            auto process = [&](uint32_t (&r)[32], int c) {
                uint32_t mask = 0u;
                if (p.debug & 4) return;                 // timing experiment: TMEM loads only
                if (p.maxima_only) {
                    // probe launch: the largest key among the live columns of this chunk; the k'-th
                    // largest chunk maximum of a query bounds its final k'-th best key from below
                    // (chunk maxima belong to distinct rows)
                    float mx = -CUDART_INF_F;
                    if (RAW) {
                        const uint32_t live = __shfl_sync(kFull, tile_word, c);
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            mx = fmaxf(mx, ((live >> j) & 1u) ? __uint_as_float(r[j]) : -CUDART_INF_F);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float2 ab = aux_w[c * 32 + j];
                            mx = fmaxf(mx, fmaf(__uint_as_float(r[j]), ab.x, ab.y));
                        }
                    }
                    if (active && mslot < p.cm_slots) cmx[mslot * BM] = mx;
                    ++mslot;
                    return;
                }
                if (RAW) {
                    // chunk maximum first: most chunks of the main phase beat no lane's threshold
                    float m[11];
#pragma unroll
                    for (int j = 0; j < 10; ++j)
                        m[j] = fmaxf(fmaxf(__uint_as_float(r[3 * j]), __uint_as_float(r[3 * j + 1])), __uint_as_float(r[3 * j + 2]));
                    m[10] = fmaxf(__uint_as_float(r[30]), __uint_as_float(r[31]));
                    const float m0 = fmaxf(fmaxf(m[0], m[1]), m[2]), m1 = fmaxf(fmaxf(m[3], m[4]), m[5]);
                    const float m2 = fmaxf(fmaxf(m[6], m[7]), m[8]), m3 = fmaxf(m[9], m[10]);
                    const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    if (p.debug & 8) return;             // timing experiment: chunk maximum only
                    if (!__any_sync(kFull, mx > thr)) return;
                    const uint32_t live = __shfl_sync(kFull, tile_word, c);
#pragma unroll
                    for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(r[j]) > thr) ? (1u << j) : 0u;
                    mask &= live;
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float2 ab = aux_w[c * 32 + j];
                        const float key = fmaf(__uint_as_float(r[j]), ab.x, ab.y);
                        r[j] = __float_as_uint(key);
                        mask |= (key > thr) ? (1u << j) : 0u;
                    }
                }
                // (2) columns admitted by ANY lane of the warp (few): visit them one by one; the column
                //     index is warp-uniform, so picking the register is a uniform jump, and only the
                //     lanes that admitted the column store (key, row id) into their own buffer
                unsigned um = __reduce_or_sync(kFull, mask);
                if (__popc(um) > 3) {
                    // dense phase (loose thresholds): 32 predicated stores beat the column walk
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t bit = (mask >> j) & 1u;
                        asm volatile(
                            "{\n\t"
                            ".reg .pred p;\n\t"
                            "setp.ne.u32 p, %3, 0;\n\t"
                            "@p st.global.v2.b32 [%0], {%1, %2};\n\t"
                            "}"
                            :: "l"(buf + cnt), "r"(r[j]), "r"(row_base + c * 32 + j), "r"(bit));
                        cnt += (int)bit;
                    }
                    um = 0u;
                }
                while (um) {
                    const int j = __ffs(um) - 1;
                    um &= um - 1;
                    uint32_t kb;
                    switch (j) {
                        case 0: kb = r[0]; break;
                        case 1: kb = r[1]; break;
                        case 2: kb = r[2]; break;
                        case 3: kb = r[3]; break;
                        case 4: kb = r[4]; break;
                        case 5: kb = r[5]; break;
                        case 6: kb = r[6]; break;
                        case 7: kb = r[7]; break;
                        case 8: kb = r[8]; break;
                        case 9: kb = r[9]; break;
                        case 10: kb = r[10]; break;
                        case 11: kb = r[11]; break;
                        case 12: kb = r[12]; break;
                        case 13: kb = r[13]; break;
                        case 14: kb = r[14]; break;
                        case 15: kb = r[15]; break;
                        case 16: kb = r[16]; break;
                        case 17: kb = r[17]; break;
                        case 18: kb = r[18]; break;
                        case 19: kb = r[19]; break;
                        case 20: kb = r[20]; break;
                        case 21: kb = r[21]; break;
                        case 22: kb = r[22]; break;
                        case 23: kb = r[23]; break;
                        case 24: kb = r[24]; break;
                        case 25: kb = r[25]; break;
                        case 26: kb = r[26]; break;
                        case 27: kb = r[27]; break;
                        case 28: kb = r[28]; break;
                        case 29: kb = r[29]; break;
                        case 30: kb = r[30]; break;
                        case 31: kb = r[31]; break;
                        default: kb = 0u; break;
                    }
                    if ((mask >> j) & 1u) {
                        buf[cnt] = make_uint2(kb, row_base + c * 32 + j);
                        ++cnt;
                    }
                }
            };
            uint32_t ra[32], rb[32];
            ptx::tmem_ld_32x32(taddr + c_begin * 32, ra);
#pragma unroll 1
            for (int c = c_begin; c < ((p.debug & 2) ? 0 : c_end); c += 2) {
                ptx::tmem_ld_wait();
                ptx::tmem_ld_32x32(taddr + (c + 1) * 32, rb);      // next chunk in flight
                process(ra, c);
                ptx::tmem_ld_wait();
                if (c + 2 < c_end) ptx::tmem_ld_32x32(taddr + (c + 2) * 32, ra);
                process(rb, c + 1);
            }
            // accumulator drained: hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (TWO) ptx::mbar_arrive_cluster(tempty_bar_leader + 8 * acc);
                else ptx::mbar_arrive(tempty_bar + 8 * acc);
            }

            // eager compaction keeps the thresholds tight (warp-collective, one owner lane at a time)
            unsigned need = __ballot_sync(kFull, cnt >= p.trigger);
            while (need) {
                const int owner = __ffs(need) - 1;
                need &= need - 1;
                const int n_o = __shfl_sync(kFull, cnt, owner);
                uint2 *buf_o = p.cand + (vcta * BM + quad * 32 + owner) * p.cap;
                int nc;
                float nt;
                compact_dispatch(buf_o, n_o, p.kprime, p.cap, lane, nc, nt);
                if (lane == owner) {
                    cnt = nc;
                    thr = fmaxf(thr, nt);
                    atomicMax(p.thr_g + q, fmap(thr));
                }
            }
        }
        // exit: bound the number of candidates the select kernel has to look at
        unsigned need = __ballot_sync(kFull, !p.maxima_only && cnt > p.exit_cap);
        while (need) {
            const int owner = __ffs(need) - 1;
            need &= need - 1;
            const int n_o = __shfl_sync(kFull, cnt, owner);
            uint2 *buf_o = p.cand + (vcta * BM + quad * 32 + owner) * p.cap;
            int nc;
            float nt;
            compact_dispatch(buf_o, n_o, p.kprime, p.cap, lane, nc, nt);
            if (lane == owner) {
                cnt = nc;
                thr = fmaxf(thr, nt);
                atomicMax(p.thr_g + q, fmap(thr));
            }
        }
        p.cand_cnt[vcta * BM + tq] = active ? cnt : 0;
    }

    ptx::tc_fence_before();
    if (TWO) ptx::cluster_sync();      // neither CTA may leave while the other can still signal / read it
    else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        if (TWO) ptx::tmem_dealloc_2sm(tmem_base, TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <bool TF32, bool RAW>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_coarse_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                 const CoarseParams p)
{
    coarse_body<TF32, false, RAW>(tmap_q, tmap_c, p);
}

template <bool TF32, bool RAW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
tc_coarse_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                      const CoarseParams p)
{
    coarse_body<TF32, true, RAW>(tmap_q, tmap_c, p);
}

// ---------------------------------------------------------------------------------------------
// 4. select k' best coarse candidates, rescore exactly, prove, write the outputs
// ---------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;
constexpr int KEPT_MAX = 448;
// Margin scheme: thresholds and the final selection are placed `kMarginMult * eps_q` BELOW the k-th best coarse
// key they are derived from.  The k rows with coarse key >= c_k have exact keys >= c_k - eps, every row that was
// rejected or not rescored has coarse key <= c_k - 2.1 eps, hence exact key <= c_k - 1.1 eps < e_k: the proof
// holds by construction unless more than KEPT_MAX rows crowd into that window (duplicates).
constexpr float kMarginMult = 2.1f;

struct SelectParams {
    const void *corpus;
    int dtype, dim, ld, metric;
    const float *norm2;
    const float *queries;     // [nq, dim] fp32 originals
    const QInfo *qinfo;
    const uint2 *cand;
    const int *cand_cnt;
    const uint32_t *thr_g;
    int qt_count, ngroups, cap;
    int k, kprime;            // kprime: rank the selection threshold is taken at (k itself with the margin scheme)
    int margin;               // 1: rescore everything within kMarginMult * eps of the k-th best coarse key
    int stage_cap;            // keys the shared-memory stage holds (more candidates: the lists are re-walked in L2)
    float *out_scores;
    long long *out_ids;
    long long id_offset;
    int *unverified;          // [nq] flag
    int *n_unverified;        // [1] counter
    int *unv_list;            // [max_sel] the unproven queries, in arrival order (input of the rescue scan)
    int max_sel;              // unproven queries beyond this many are returned as id -1 / score NaN
    int *sticky;              // [1] running count of such queries (never reset)
};

__device__ __forceinline__ float row_elem(const void *corpus, int dtype, size_t idx)
{
    if (dtype == ARCHI_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(corpus)[idx]);
    return reinterpret_cast<const float *>(corpus)[idx];
}

// Shared front half of tc_select_kernel / tc_threshold_kernel (one CTA per query): T = the
// kprime-th largest mapped coarse key over every candidate list of query q (0 when there are at most
// kprime candidates).  MSB-first radix select, 8 bits per pass: each pass re-walks the lists in
// global memory (L2-resident, coalesced: warp w takes lists w, w+8, ...; lanes take entries), builds
// a 256-bin histogram of the keys that match the prefix found so far, and one warp picks the digit.
struct SelCommon {
    const uint2 *cand;
    const int *cand_cnt;
    int qt_count, ngroups, cap, kprime;
};

__device__ __forceinline__ size_t sel_list_base(const SelCommon &c, int l, int qt)
{
    return ((size_t)((l >> 1) * c.qt_count + qt)) * 2 + (l & 1);
}

constexpr int SEL_STAGE = 7680;   // keys staged in shared memory for the radix passes when they fit

constexpr int SEL_LISTS_MAX = 320;   // 2 * ngroups <= 296

// Candidate `i` of the query's concatenated lists (s_off = exclusive prefix sums of the list sizes,
// s_off[nlists] = total): which list, and where inside it.
__device__ __forceinline__ const uint2 *sel_locate(const SelCommon &c, const int *s_off, int nlists, int qt, int tq, int i)
{
    int lo = 0, hi = nlists;               // largest l with s_off[l] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= i) lo = mid;
        else hi = mid;
    }
    return c.cand + (sel_list_base(c, lo, qt) * BM + tq) * c.cap + (i - s_off[lo]);
}

// s_cnt[SEL_LISTS_MAX], s_off[SEL_LISTS_MAX + 1], s_hist[256 * warps], s_misc[8], s_stage[SEL_STAGE] are
// shared-memory scratch (s_stage holds stage_cap keys); returns T, the number of candidates through total_out.  When
// total <= stage_cap the mapped keys are left in s_stage in concatenated-list order (see sel_locate).
template <int NT>
__device__ __forceinline__ uint32_t sel_radix_threshold(const SelCommon &c, int q, int *s_cnt, int *s_off, int *s_hist,
                                                       int *s_misc, uint32_t *s_stage, int stage_cap, int &total_out)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qt = q / BM, tq = q % BM;
    const int nlists = 2 * c.ngroups;
    for (int l = tid; l < nlists; l += NT) s_cnt[l] = c.cand_cnt[sel_list_base(c, l, qt) * BM + tq];
    __syncthreads();
    if (warp == 0) {
        // exclusive prefix sums: lane L owns lists [10 L, 10 L + 10)
        int mine = 0;
        for (int j = 0; j < SEL_LISTS_MAX / 32; ++j) {
            const int l = lane * (SEL_LISTS_MAX / 32) + j;
            mine += l < nlists ? s_cnt[l] : 0;
        }
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += o;
        }
        int run = incl - mine;
        for (int j = 0; j < SEL_LISTS_MAX / 32; ++j) {
            const int l = lane * (SEL_LISTS_MAX / 32) + j;
            if (l < nlists) {
                s_off[l] = run;
                run += s_cnt[l];
            }
        }
        if (lane == 31) s_off[nlists] = incl;
    }
    __syncthreads();
    const int total = s_off[nlists];
    total_out = total;
    if (total <= c.kprime) return 0u;
    const bool staged = total <= stage_cap;
    if (staged) {
        // one walk over global memory, every load independent of the others (this kernel is latency bound)
        for (int i0 = tid; i0 < total; i0 += 4 * NT) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * NT;
                if (i < total) v[u] = __ldcg(&sel_locate(c, s_off, nlists, qt, tq, i)->x);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * NT;
                if (i < total) s_stage[i] = fmap(__uint_as_float(v[u]));
            }
        }
        __syncthreads();
        int n_gt;
        return block_radix_kth(s_stage, total, c.kprime, s_hist, s_misc, n_gt);
    }

    // too many candidates to stage (rare): the same selection, re-walking the lists in global memory
    uint32_t prefix = 0u, known = 0u;   // bits of T fixed so far / their mask
    int need = c.kprime;                // rank still to be located inside the current prefix bucket
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < 256; i += NT) s_hist[i] = 0;
        __syncthreads();
        for (int l = warp; l < nlists; l += NT / 32) {
            const int n = s_cnt[l];
            const uint2 *src = c.cand + (sel_list_base(c, l, qt) * BM + tq) * c.cap;
            for (int i = lane; i < n; i += 32) {
                const uint32_t key = fmap(__uint_as_float(src[i].x));
                if ((key & known) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1);
            }
        }
        __syncthreads();
        if (warp == 0) {
            // lane L owns bins [8L, 8L+8); walk from the top bin down until `need` keys are covered
            int mine = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) mine += s_hist[lane * 8 + j];
            // inclusive suffix sum over lanes (bins of higher lanes hold larger keys)
            int suf = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_down_sync(kFull, suf, d);
                if (lane + d < 32) suf += o;
            }
            const int above = suf - mine;           // keys in bins owned by higher lanes
            const bool here = above < need && suf >= need;     // the crossing happens in my 8 bins
            const unsigned who = __ballot_sync(kFull, here);
            const int src_lane = __ffs(who) - 1;
            if (lane == src_lane) {
                int acc = above, digit = lane * 8;
                for (int j = 7; j >= 0; --j) {
                    const int h = s_hist[lane * 8 + j];
                    if (acc + h >= need) {
                        digit = lane * 8 + j;
                        break;
                    }
                    acc += h;
                }
                s_misc[1] = digit;
                s_misc[2] = need - acc;                         // rank inside the chosen bin
            }
        }
        __syncthreads();
        prefix |= (uint32_t)s_misc[1] << shift;
        known |= 255u << shift;
        need = s_misc[2];
        __syncthreads();
    }
    return prefix;
}

// After a warm-up phase: publish, per query, the kprime-th best coarse key over ALL its candidate
// lists as the shared threshold.  It is a valid lower bound of the final kprime-th best (the rows seen
// so far are a subset of the corpus) and far tighter than any single CTA's local threshold.
struct ThresholdParams {
    SelCommon c;
    uint32_t *thr_g;
    const QInfo *qinfo;
    int margin;
};

__global__ void __launch_bounds__(SEL_THREADS) tc_threshold_kernel(const ThresholdParams p)
{
    __shared__ int s_cnt[SEL_LISTS_MAX];
    __shared__ int s_off[SEL_LISTS_MAX + 1];
    __shared__ int s_hist[256 * (SEL_THREADS / 32)];
    __shared__ int s_misc[8];
    __shared__ uint32_t s_stage[SEL_STAGE];
    int total;
    const uint32_t T = sel_radix_threshold<SEL_THREADS>(p.c, blockIdx.x, s_cnt, s_off, s_hist, s_misc, s_stage, SEL_STAGE, total);
    if (threadIdx.x == 0 && total > p.c.kprime) {
        const uint32_t Tm = p.margin ? fmap(funmap(T) - kMarginMult * p.qinfo[blockIdx.x].eps) : T;
        atomicMax(p.thr_g + blockIdx.x, Tm);
    }
}

// After a probe launch: per query, the kprime-th largest chunk maximum over all lists becomes the
// shared threshold (chunk maxima are keys of distinct rows, so at least kprime rows reach it).
struct MaxThrParams {
    const float *chunkmax;
    int qt_count, ngroups, cm_slots, used_slots, kprime, nq;
    uint32_t *thr_g;
    const QInfo *qinfo;
    int margin;
    int one_warp;             // 1 (default): one warp per query; 0 (ARCHI_TC_THR4=1): four warps per query for k <= 32
};

constexpr int MAXTHR_Q = 8;            // queries per CTA of tc_maxima_threshold_kernel: one warp each for the selection
constexpr int MAXTHR_THREADS = 1024;   // ... after all 32 warps gathered the maxima (latency bound: wide is fast)

__global__ void __launch_bounds__(MAXTHR_THREADS) tc_maxima_threshold_kernel(const MaxThrParams p)
{
    extern __shared__ uint32_t s_mkeys[];            // [MAXTHR_Q][stride >= 2 * ngroups * used_slots]
    __shared__ int s_hist[MAXTHR_Q][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * MAXTHR_Q;            // MAXTHR_Q divides BM: all in one query tile
    const int qt = q0 / BM, tq0 = q0 % BM;
    const int total = 2 * p.ngroups * p.used_slots;
    const int stride = ((total + 31) & ~31) + 4;     // bank-conflict-free transposed stores
    // gather: 8 consecutive queries of one (list, slot) are one 32-byte sector; 8 loads in flight per thread
    const int j = tid % MAXTHR_Q, n_el = total * MAXTHR_Q;
    for (int base = tid; base < n_el; base += MAXTHR_THREADS * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * MAXTHR_THREADS;
            if (i < n_el) {
                const int e = i / MAXTHR_Q;
                const int l = e / p.used_slots, sl = e - l * p.used_slots;
                const size_t vcta = ((size_t)((l >> 1) * p.qt_count + qt)) * 2 + (l & 1);
                v[u] = __ldcg(p.chunkmax + (vcta * p.cm_slots + sl) * BM + tq0 + j);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * MAXTHR_THREADS;
            if (i < n_el) s_mkeys[j * stride + i / MAXTHR_Q] = fmap(v[u]);
        }
    }
    __syncthreads();
    if (p.kprime > 32 || p.one_warp) {
        // one warp per query: a sorted register list for k <= 32, radix select over the whole list for larger k
        const int q = q0 + warp;
        if (warp >= MAXTHR_Q || q >= p.nq) return;
        const uint32_t *keys = s_mkeys + warp * stride;
        const uint32_t none = fmap(-CUDART_INF_F);
        int valid = 0;
#pragma unroll 8
        for (int i = lane; i < total; i += 32) valid += keys[i] > none ? 1 : 0;
        valid = __reduce_add_sync(kFull, valid);
        if (valid < p.kprime) return;                    // fewer live rows than kprime seen: no threshold yet
        const uint32_t T = p.kprime <= 32 ? warp_kth_small(keys, total, p.kprime, lane)
                                          : warp_radix_kth(keys, total, p.kprime, s_hist[warp], lane);
        if (lane == 0) atomicMax(p.thr_g + q, p.margin ? fmap(funmap(T) - kMarginMult * p.qinfo[q].eps) : T);
        return;
    }
    // k <= 32: four warps per query, each keeps the 32 largest of a quarter of the maxima (sorted register list);
    // the first warp of the group then takes the k-th largest of the four lists -- the k-th largest of the union.
    // s_hist is reused: [MAXTHR_Q][4][32] keys.  Keys of dead chunks map to fmap(-inf): they never count.
    uint32_t *s_top = reinterpret_cast<uint32_t *>(&s_hist[0][0]);
    const int grp = warp >> 2, sub = warp & 3;
    {
        const uint32_t *keys = s_mkeys + grp * stride;
        const int quarter = (((total + 3) >> 2) + 31) & ~31;
        const int begin = sub * quarter, end = min(total, begin + quarter);
        const uint32_t mine = begin < end ? warp_top32_small(keys, begin, end, p.kprime, lane) : 0u;
        s_top[(grp * 4 + sub) * 32 + lane] = mine;
    }
    __syncthreads();
    const int q = q0 + grp;
    if (sub != 0 || q >= p.nq) return;
    const uint32_t T = warp_kth_small(s_top + grp * 128, 128, p.kprime, lane);
    // fewer than kprime live maxima seen (T is then 0 or the key of a dead chunk): no threshold yet
    if (T <= fmap(-CUDART_INF_F)) return;
    if (lane == 0) atomicMax(p.thr_g + q, p.margin ? fmap(funmap(T) - kMarginMult * p.qinfo[q].eps) : T);
}

// NT = 256 threads per query, or 128 for batches of more CTAs than fit on the machine at 256 (64 registers per thread:
// four 256-thread CTAs per SM, eight 128-thread ones -- 1024 queries are then ONE wave of this latency-bound chain).
template <int NT>
__global__ void __launch_bounds__(NT) tc_select_kernel(const SelectParams p)
{
    __shared__ int s_cnt[SEL_LISTS_MAX];            // per-list sizes and their exclusive prefix sums
    __shared__ int s_off[SEL_LISTS_MAX + 1];
    __shared__ int s_hist[256 * (NT / 32)];
    __shared__ int s_misc[8];
    __shared__ int s_nk;
    __shared__ uint32_t s_kid[KEPT_MAX];
    __shared__ float s_ex[KEPT_MAX];                // exact key (coarse-key space)
    __shared__ float s_sc[KEPT_MAX];                // output score
    extern __shared__ __align__(16) float s_q[];    // [ld] the query, zero padded to the row stride, then the key stage
    uint32_t *s_stage = reinterpret_cast<uint32_t *>(s_q + p.ld);      // [stage_cap]
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qt = q / BM, tq = q % BM;
    const int nlists = 2 * p.ngroups;
    SelCommon sc;
    sc.cand = p.cand;
    sc.cand_cnt = p.cand_cnt;
    sc.qt_count = p.qt_count;
    sc.ngroups = p.ngroups;
    sc.cap = p.cap;
    sc.kprime = p.kprime;
    if (tid == 0) s_nk = 0;
    // the query is needed last: fetch it first, its latency hides behind the selection
    for (int e = tid; e < p.ld; e += NT) s_q[e] = e < p.dim ? p.queries[(size_t)q * p.dim + e] : 0.f;
    const QInfo qi = p.qinfo[q];
    int total;
    uint32_t T = sel_radix_threshold<NT>(sc, q, s_cnt, s_off, s_hist, s_misc, s_stage, p.stage_cap, total);
    // margin scheme: T is the k-th best coarse key c_k; everything down to c_k - kMarginMult * eps is rescored
    if (p.margin && total > p.kprime) T = fmap(funmap(T) - kMarginMult * qi.eps);
    __syncthreads();

    // gather the survivors (key >= T)
    if (total <= p.stage_cap) {
        // keys are staged in list order: only the survivors' row ids come from global memory
        for (int i = tid; i < total; i += NT) {
            if (total <= p.kprime || s_stage[i] >= T) {
                const int slot = atomicAdd(&s_nk, 1);
                if (slot < KEPT_MAX) s_kid[slot] = __ldcg(&sel_locate(sc, s_off, nlists, qt, tq, i)->y);
            }
        }
    } else {
        for (int l = warp; l < nlists; l += NT / 32) {
            const int n = s_cnt[l];
            const uint2 *src = p.cand + (sel_list_base(sc, l, qt) * BM + tq) * p.cap;
            for (int i = lane; i < n; i += 32) {
                const uint2 e = src[i];
                if (fmap(__uint_as_float(e.x)) >= T) {
                    const int slot = atomicAdd(&s_nk, 1);
                    if (slot < KEPT_MAX) s_kid[slot] = e.y;
                }
            }
        }
    }
    __syncthreads();
    const bool overflow = s_nk > KEPT_MAX;          // more ties at T than we can hold: cannot prove
    const int nk = overflow ? KEPT_MAX : s_nk;

    // exact fp32 rescoring: a warp takes four candidates at a time and keeps the loads of all four
    // rows (16-byte loads) and their norms in flight together
    const int vec = p.dtype == ARCHI_BF16 ? 8 : 4;
    const int nvec = p.ld / vec;
    const size_t row_bytes = (size_t)p.ld * (p.dtype == ARCHI_BF16 ? 2 : 4);
    for (int cb = warp * 4; cb < nk; cb += (NT / 32) * 4) {
        size_t row[4];
        const uint4 *rp[4];
        float n2[4], acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            row[j] = s_kid[cb + j < nk ? cb + j : cb];
            rp[j] = reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(p.corpus) + row[j] * row_bytes);
            n2[j] = p.metric == ARCHI_COSINE ? __ldg(p.norm2 + row[j]) : 1.f;
            acc[j] = 0.f;
        }
        for (int v = lane; v < nvec; v += 32) {
            uint4 d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = __ldg(rp[j] + v);
            const float *qq = s_q + v * vec;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
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
                for (int i = 0; i < 8; ++i) {
                    if (i < vec) {
                        if (p.metric == ARCHI_L2) {
                            const float t = x[i] - qq[i];
                            acc[j] = fmaf(t, t, acc[j]);
                        } else {
                            acc[j] = fmaf(x[i], qq[i], acc[j]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) acc[j] += __shfl_xor_sync(kFull, acc[j], d);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (lane == j && cb + j < nk) {
                float ex, scv;
                if (p.metric == ARCHI_COSINE) {
                    const float rn = n2[j] > 0.f ? 1.0f / sqrtf(n2[j]) : 0.f;
                    ex = n2[j] > 0.f ? acc[j] * rn : -CUDART_INF_F;        // coarse-key space: dot / |c|
                    scv = fminf(1.f, fmaxf(-1.f, acc[j] * qi.rn_q * rn));
                } else if (p.metric == ARCHI_IP) {
                    ex = acc[j];
                    scv = -acc[j];
                } else {
                    ex = qi.qn2 - acc[j];                                  // 2 q.c - |c|^2 = |q|^2 - d^2
                    scv = sqrtf(fmaxf(acc[j], 0.f));
                }
                s_ex[cb + j] = ex;
                s_sc[cb + j] = scv;
            }
        }
    }
    __syncthreads();

    // rank by exact key (ties: lower id) and write the k best
    float ek = -CUDART_INF_F;                         // exact key at rank k-1
    for (int i = tid; i < nk; i += NT) {
        const float mine = s_ex[i];
        const int mid = (int)s_kid[i];
        int rank = 0;
        for (int j = 0; j < nk; ++j) rank += better(s_ex[j], (int)s_kid[j], mine, mid) ? 1 : 0;
        if (rank < p.k) {
            const bool dead = mine == -CUDART_INF_F;              // zero-norm row under cosine: not a match
            p.out_scores[(size_t)q * p.k + rank] = dead ? CUDART_NAN_F : s_sc[i];
            p.out_ids[(size_t)q * p.k + rank] = dead ? -1ll : (long long)mid + p.id_offset;
        }
        if (rank == p.k - 1) ek = mine;
    }
    for (int r = nk + tid; r < p.k; r += NT) {
        p.out_scores[(size_t)q * p.k + r] = CUDART_NAN_F;
        p.out_ids[(size_t)q * p.k + r] = -1;
    }
    // proof: every row outside the kept set has coarse key <= Tv, hence exact key <= Tv + eps
    __shared__ float s_ek;
    __shared__ int s_poison;
    if (tid == 0) {
        s_ek = -CUDART_INF_F;
        s_poison = 0;
    }
    __syncthreads();
    if (ek > -CUDART_INF_F) s_ek = ek;               // exactly one thread holds rank k-1
    __syncthreads();
    if (tid == 0) {
        float Tv = total > p.kprime ? funmap(T) : -CUDART_INF_F;
        Tv = fmaxf(Tv, thr_from_word(p.thr_g[q]));
        bool ok = !overflow;
        if (Tv > -CUDART_INF_F) ok = ok && nk >= p.k && s_ek > Tv + qi.eps;
        if (!ok) {
            // unproven: queue the query for the exact re-scan that follows on the same stream
            p.unverified[q] = 1;
            const int slot = atomicAdd(p.n_unverified, 1);
            if (slot < p.max_sel) p.unv_list[slot] = q;
            else {
                s_poison = 1;                        // no room in the rescue list: never return an unproven row
                atomicAdd(p.sticky, 1);
            }
        }
    }
    __syncthreads();
    if (s_poison) {
        for (int r = tid; r < p.k; r += NT) {
            p.out_scores[(size_t)q * p.k + r] = CUDART_NAN_F;
            p.out_ids[(size_t)q * p.k + r] = -1;
        }
    }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_fn()
{
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(ptr);
    }
    return fn;
}

// [rows, ld] row-major matrix of 2- or 4-byte elements; box = 128 bytes x box_rows, 128B swizzle.
static int make_tmap(CUtensorMap *map, const void *base, int is_f32, long long rows, int ld, int box_rows)
{
    encode_tiled_fn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return ARCHI_ECUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)(rows > 0 ? rows : 1)};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * (is_f32 ? 4u : 2u)};
    const cuuint32_t box[2] = {(cuuint32_t)(is_f32 ? 32 : 64), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                    const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld ld=%d)", (int)r, rows, ld);
        return ARCHI_ECUDA;
    }
    return ARCHI_OK;
}

// The descriptor of a matrix is rebuilt only when its base, shape or box changes (cuTensorMapEncodeTiled costs
// microseconds of host time per call, which is all a small shard's search has).
static int cached_tmap(unsigned char *slot, const void **c_base, long long *c_rows, int *c_ld, int *c_f32, int *c_box,
                       const void *base, int is_f32, long long rows, int ld, int box_rows)
{
    if (*c_base == base && *c_rows == rows && *c_ld == ld && *c_f32 == is_f32 && (!c_box || *c_box == box_rows))
        return ARCHI_OK;
    int rc = make_tmap(reinterpret_cast<CUtensorMap *>(slot), base, is_f32, rows, ld, box_rows);
    if (rc != ARCHI_OK) {
        *c_base = nullptr;
        return rc;
    }
    *c_base = base;
    *c_rows = rows;
    *c_ld = ld;
    *c_f32 = is_f32;
    if (c_box) *c_box = box_rows;
    return ARCHI_OK;
}

// cudaFuncSetAttribute once per kernel and size (a driver call per search otherwise)
static int set_dyn_smem_once(const void *fn, int bytes)
{
    static std::mutex mu;
    static const void *fns[32];
    static int sizes[32];
    static int n = 0;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n; ++i)
        if (fns[i] == fn) {
            if (sizes[i] >= bytes) return ARCHI_OK;
            ARCHI_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            sizes[i] = bytes;
            return ARCHI_OK;
        }
    ARCHI_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (n < 32) {
        fns[n] = fn;
        sizes[n] = bytes;
        ++n;
    }
    return ARCHI_OK;
}

template <typename T>
static int ensure_buf(T **ptr, size_t *cap_bytes, size_t need_bytes)
{
    if (*cap_bytes >= need_bytes && *ptr) return ARCHI_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap_bytes = 0;
    ARCHI_CUDA(cudaMalloc(ptr, need_bytes));
    *cap_bytes = need_bytes;
    return ARCHI_OK;
}

int tensor_path_supported(const archi_store *s, int k)
{
    // row stride must be a whole number of 16-byte units (always true) and k' must fit a buffer
    return s->rows > 0 && k >= 1 && k <= kMaxListK && s->rows < (1ll << 31);
}

int launch_tensor_search(archi_store *s, const float *q_dev, int nq, int k, const uint32_t *filter, int include_deleted,
                         float *out_scores, int64_t *out_ids, int64_t id_offset, cudaStream_t st,
                         int *max_sel_out, double *coarse_ms)
{
    using namespace tc;
    TensorWorkspace &w = s->tws;
    // fp32 stores: the coarse pass reads a bf16 shadow copy (ARCHI_NO_SHADOW=1 keeps kind::tf32 on the
    // fp32 rows instead: no extra memory, half the MMA rate and twice the bytes)
    static const bool no_shadow = getenv("ARCHI_NO_SHADOW") && atoi(getenv("ARCHI_NO_SHADOW")) != 0;
    const bool use_shadow = s->dtype == ARCHI_F32 && !no_shadow;
    const bool tf32 = s->dtype == ARCHI_F32 && !use_shadow;
    const int kelems = tf32 ? 32 : 64;
    const int ldq = round_up(s->dim, tf32 ? 4 : 8);
    const int ld_sh = round_up(s->dim, 8);
    // CTA-pair mode (cta_group::2) needs at least two query tiles; the tile count is padded to even
    static const int pair_env = getenv("ARCHI_TC_PAIR") ? atoi(getenv("ARCHI_TC_PAIR")) : 1;
    int qt_count = (nq + BM - 1) / BM;
    const bool pair = pair_env != 0 && qt_count >= 2;
    if (pair) qt_count = (qt_count + 1) & ~1;
    ARCHI_REQUIRE(qt_count <= MAX_QT, "tensor path: at most %d queries per launch", MAX_QT * BM);
    const int nq_pad = qt_count * BM;
    int kprime = k <= 10 ? 32 : round_up(2 * k + 12, 32);
    if (kprime > 256) kprime = 256;
    // ARCHI_TC_MARGIN=0 restores the older scheme (thresholds at the k'-th best key, no margin)
    static const int margin = getenv("ARCHI_TC_MARGIN") ? atoi(getenv("ARCHI_TC_MARGIN")) : 1;
    // buffer capacity: after a compaction (kprime entries) a buffer absorbs cap - BN - kprime more
    // candidates before the next one; a whole tile (BN) always fits
    const int cap = 1024;
    const int n_ctiles = (int)((s->rows + BN - 1) / BN);
    int ngroups = s->sm_count / qt_count;
    if (ngroups > 148) ngroups = 148;
    if (ngroups < 1) ngroups = 1;
    if (ngroups > n_ctiles) ngroups = n_ctiles;
    const int grid = ngroups * qt_count;

    // ---- workspace ----
    int rc;
    if ((rc = ensure_buf(&w.qstage, &w.qstage_bytes, (size_t)nq_pad * ldq * 4)) != ARCHI_OK) return rc;
    if ((rc = ensure_buf(&w.qinfo, &w.qinfo_bytes, (size_t)nq_pad * sizeof(QInfo))) != ARCHI_OK) return rc;
    if ((rc = ensure_buf(&w.thr_g, &w.thr_bytes, (size_t)nq_pad * 4)) != ARCHI_OK) return rc;
    if ((rc = ensure_buf(&w.unverified, &w.unv_bytes, (size_t)(nq_pad + 1) * 4)) != ARCHI_OK) return rc;
    // every query of a batch of <= 256 can be rescued; larger batches rescue up to 256 of theirs
    const int max_sel = nq <= 256 ? round_up(nq, kMaxQB) : 256;
    if ((rc = ensure_buf(&w.unv_list, &w.unv_list_bytes, (size_t)256 * 4)) != ARCHI_OK) return rc;
    if (!w.h_verdict) {
        ARCHI_CUDA(cudaHostAlloc(&w.h_verdict, 64 * sizeof(int), cudaHostAllocDefault));
        ARCHI_CUDA(cudaEventCreateWithFlags(&w.verdict_ev, cudaEventDisableTiming));
        ARCHI_CUDA(cudaMalloc(&w.sticky_dev, sizeof(int)));
        ARCHI_CUDA(cudaMemsetAsync(w.sticky_dev, 0, sizeof(int), st));
    }
    w.unv_count = w.unverified + nq_pad;
    if ((rc = ensure_buf(&w.cand, &w.cand_bytes, (size_t)grid * 2 * BM * cap * sizeof(uint2))) != ARCHI_OK) return rc;
    if ((rc = ensure_buf(&w.cand_cnt, &w.cnt_bytes, (size_t)grid * 2 * BM * 4)) != ARCHI_OK) return rc;
    {
        const size_t need = (size_t)s->capacity * sizeof(float2);
        const bool realloc = !w.aux || w.aux_bytes < need;
        if ((rc = ensure_buf(&w.aux, &w.aux_bytes, need)) != ARCHI_OK) return rc;
        if (realloc) w.aux_epoch = -1;
    }
    if (!w.max_norm2) ARCHI_CUDA(cudaMalloc(&w.max_norm2, 8));
    if (use_shadow) {
        const size_t need = (size_t)s->capacity * ld_sh * 2;
        const bool realloc = !w.shadow || w.shadow_bytes < need;
        if ((rc = ensure_buf(&w.shadow, &w.shadow_bytes, need)) != ARCHI_OK) return rc;
        if (realloc || w.shadow_reset_epoch != s->reset_epoch) w.shadow_rows = 0;
        w.shadow_reset_epoch = s->reset_epoch;
        if (w.shadow_rows < s->rows) {
            const long long n_new = s->rows - w.shadow_rows;
            const long long tot = n_new * ld_sh;
            tc_shadow_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
                (const float *)s->data, (__nv_bfloat16 *)w.shadow, w.shadow_rows, n_new, s->dim, s->ld, ld_sh,
                s->metric == ARCHI_COSINE ? s->norm2 : nullptr);
            ARCHI_CHECK_LAUNCH();
            w.shadow_rows = s->rows;
        }
    }

    // ---- cached per-store data: max |row|^2 and the (a, b) constants ----
    const uint32_t *alive = include_deleted ? nullptr : s->alive;
    if (w.maxnorm_epoch != s->epoch || w.maxnorm_all_rows != (include_deleted != 0)) {
        const uint32_t init[2] = {0u, 0x7f800000u};   // max = 0, min = +inf
        ARCHI_CUDA(cudaMemcpyAsync(w.max_norm2, init, 8, cudaMemcpyHostToDevice, st));
        // the bounds must cover every row the scan can return: tombstoned rows too when include_deleted is set
        tc_maxnorm_kernel<<<s->sm_count * 2, 256, 0, st>>>(s->norm2, alive, s->rows, w.max_norm2);
        ARCHI_CHECK_LAUNCH();
        float h[2];
        ARCHI_CUDA(cudaMemcpyAsync(h, w.max_norm2, 8, cudaMemcpyDeviceToHost, st));
        ARCHI_CUDA(cudaStreamSynchronize(st));
        w.h_max_norm2 = h[0];
        w.h_min_norm2 = h[1];
        w.maxnorm_epoch = s->epoch;
        w.maxnorm_all_rows = include_deleted != 0;
    }
    // Raw-key epilogue (no per-column constants): inner product always; cosine when the rows the
    // coarse pass reads are unit length -- exactly (normalised bf16 shadow of an fp32 store) or within
    // a small, measured deviation that is added to the error bound (bf16 stores of unit-norm rows).
    bool raw = false, cos_raw_unnorm = false;
    float norm_dev = 0.f;
    static const int raw_env = getenv("ARCHI_TC_RAW") ? atoi(getenv("ARCHI_TC_RAW")) : 1;
    if (raw_env) {
        if (s->metric == ARCHI_IP) raw = true;
        else if (s->metric == ARCHI_COSINE && use_shadow) raw = true;
        else if (s->metric == ARCHI_COSINE && w.h_min_norm2 > 0.f) {
            const float d0 = fabsf(1.0f / sqrtf(w.h_min_norm2) - 1.0f), d1 = fabsf(1.0f / sqrtf(w.h_max_norm2) - 1.0f);
            norm_dev = d0 > d1 ? d0 : d1;
            if (norm_dev < 1.0f / 64.0f) raw = cos_raw_unnorm = true;
        }
    }
    const bool aux_cached = w.aux_epoch == s->epoch && w.aux_alive == (alive != nullptr) && filter == nullptr &&
                            !w.aux_had_filter;
    if (!raw && !aux_cached) {
        tc_aux_kernel<<<(unsigned)((s->rows + 255) / 256), 256, 0, st>>>(reinterpret_cast<float2 *>(w.aux), s->norm2, alive,
                                                                        filter, s->rows, s->metric);
        ARCHI_CHECK_LAUNCH();
        w.aux_epoch = s->epoch;
        w.aux_alive = alive != nullptr;
        w.aux_had_filter = filter != nullptr;
    }

    // ---- 1. stage queries ----
    if (nq_pad > nq)  // padded query rows must be finite (zeros) for the MMA
        ARCHI_CUDA(cudaMemsetAsync((char *)w.qstage + (size_t)nq * ldq * (tf32 ? 4 : 2), 0,
                                   (size_t)(nq_pad - nq) * ldq * (tf32 ? 4 : 2), st));
    PrepParams pp;
    pp.queries = q_dev;
    pp.nq = nq;
    pp.dim = s->dim;
    pp.ldq = ldq;
    pp.tf32 = tf32;
    pp.metric = s->metric;
    pp.qstage = w.qstage;
    pp.qinfo = reinterpret_cast<QInfo *>(w.qinfo);
    pp.thr_g = w.thr_g;
    pp.unverified = w.unverified;
    pp.max_norm2 = w.max_norm2;
    pp.corpus_rel_err = use_shadow ? 0.001953125f : 0.f;
    pp.cos_raw_unnorm = cos_raw_unnorm;
    pp.norm_dev = norm_dev;
    pp.n_unverified = w.unverified + nq_pad;
    tc_prep_kernel<<<nq, 128, 0, st>>>(pp);
    ARCHI_CHECK_LAUNCH();

    // ---- 3. coarse scorer ----
    if ((rc = cached_tmap(w.tmap_q, &w.tmq_base, &w.tmq_rows, &w.tmq_ld, &w.tmq_f32, nullptr, w.qstage, tf32, nq_pad, ldq,
                          BM)) != ARCHI_OK)
        return rc;
    if ((rc = cached_tmap(w.tmap_c, &w.tmc_base, &w.tmc_rows, &w.tmc_ld, &w.tmc_f32, &w.tmc_box,
                          use_shadow ? w.shadow : s->data, tf32, s->rows, use_shadow ? ld_sh : s->ld,
                          pair ? BN / 2 : BN)) != ARCHI_OK)
        return rc;
    const CUtensorMap &tmap_q = *reinterpret_cast<const CUtensorMap *>(w.tmap_q);
    const CUtensorMap &tmap_c = *reinterpret_cast<const CUtensorMap *>(w.tmap_c);
    CoarseParams cp;
    cp.n = s->rows;
    cp.n_ctiles = n_ctiles;
    cp.kchunks = ((use_shadow ? ld_sh : s->ld) + kelems - 1) / kelems;
    cp.kelems = kelems;
    cp.nq = nq;
    cp.qt_count = qt_count;
    cp.ngroups = ngroups;
    cp.kprime = kprime;
    cp.cap = cap;
    cp.trigger = cap - BN + 1;   // lazy: appends are cheap, compactions are not; a whole tile always fits
    {
        const char *dbg = getenv("ARCHI_TC_DEBUG");
        cp.debug = dbg ? atoi(dbg) : 0;
        // column-split epilogue: pays when a tile's MMAs are short (measured: D=384 step -8 %, D>=768 +-2 %)
        static const int split = getenv("ARCHI_TC_SPLIT") ? atoi(getenv("ARCHI_TC_SPLIT")) : -1;
        cp.split = split >= 0 ? split : (cp.kchunks * (tf32 ? 2 : 1) <= 8 ? 1 : 0);
        // resident query tile when it leaves room for at least 4 corpus-chunk stages
        static const int ares_env = getenv("ARCHI_TC_ARES") ? atoi(getenv("ARCHI_TC_ARES")) : 1;
        const int b_bytes = (pair ? BN / 2 : BN) * 128;
        int a_stages = (RING_BYTES - cp.kchunks * A_BYTES) / b_bytes;
        if (a_stages > MAX_STAGES) a_stages = MAX_STAGES;
        cp.a_res = ares_env && a_stages >= 4;
        cp.a_stages = a_stages;
    }
    cp.exit_cap = cap;           // final launch: nothing to bound (the select kernel walks global memory)
    {
        // soft throttle of long scans (see the producer loop): window sized so that the tiles in flight of all groups
        // stay within about a third of L2; ARCHI_TC_THROTTLE = 0 off, 1 always on, unset = scans of >= 512 tiles per CTA
        static const int thr_env = getenv("ARCHI_TC_THROTTLE") ? atoi(getenv("ARCHI_TC_THROTTLE")) : -1;
        const long long tiles_per_cta = n_ctiles / ngroups;
        const bool on = thr_env == 1 || (thr_env != 0 && qt_count >= 2 && tiles_per_cta >= 512);
        cp.throttle_win = 0;
        cp.progress = nullptr;
        cp.launch_tag = 0;
        if (on) {
            const double tile_bytes = (double)BN * (use_shadow ? ld_sh : s->ld) * (tf32 ? 4 : 2);
            int win = (int)(40e6 / (tile_bytes * ngroups));
            win = win < 2 ? 2 : (win > 16 ? 16 : win);
            const bool fresh = !w.progress || w.progress_bytes < (size_t)grid * 4;
            if ((rc = ensure_buf(&w.progress, &w.progress_bytes, (size_t)grid * 4)) != ARCHI_OK) return rc;
            if (fresh) ARCHI_CUDA(cudaMemsetAsync(w.progress, 0, w.progress_bytes, st));
            cp.throttle_win = win;
            cp.progress = w.progress;
        }
    }
    cp.aux = reinterpret_cast<const float2 *>(w.aux);
    cp.alive = alive;
    cp.filter = filter;
    cp.cand = reinterpret_cast<uint2 *>(w.cand);
    cp.cand_cnt = w.cand_cnt;
    cp.thr_g = w.thr_g;
    void (*kern)(const CUtensorMap, const CUtensorMap, const CoarseParams);
    if (pair) kern = tf32 ? (raw ? tc_coarse_pair_kernel<true, true> : tc_coarse_pair_kernel<true, false>)
                          : (raw ? tc_coarse_pair_kernel<false, true> : tc_coarse_pair_kernel<false, false>);
    else kern = tf32 ? (raw ? tc_coarse_kernel<true, true> : tc_coarse_kernel<true, false>)
                     : (raw ? tc_coarse_kernel<false, true> : tc_coarse_kernel<false, false>);
    if ((rc = set_dyn_smem_once((const void *)kern, SMEM_BYTES)) != ARCHI_OK) return rc;
    // Thresholds local to one (CTA, epilogue group) only ever see 1/(2*ngroups) of the rows, so most of
    // a plain run would be spent storing candidates that a global view rejects.  The scan therefore
    // starts from ONE shared threshold per query, derived from a small sample of the corpus.
    static const int warm = getenv("ARCHI_TC_WARM") ? atoi(getenv("ARCHI_TC_WARM")) : 1;
    // warm == 1 (default): PROBE scheme.  A first launch scans a small prefix of the corpus (about
    // 1/12, at most 8 tiles per epilogue group) and only records the maximum live key of every
    // 32-column chunk; tc_maxima_threshold_kernel turns the kprime-th largest chunk maximum of each
    // query into its shared threshold; one main launch then scans everything.  No flood of
    // candidates, no resume.  warm == 2: the older flood + resume phases (kept for comparison): every
    // epilogue group scans one tile and keeps everything, tc_threshold_kernel turns the union of all
    // lists into the shared threshold, the same again after ~1/16 of the corpus, then the rest.
    // warm == 0: local thresholds only.
    bool probed = false;
    if (warm == 1 && n_ctiles >= 8 * ngroups) {
        const int nlists = 2 * ngroups;
        int cm_slots = (4096 / nlists) / 8 * 8;       // MAXTHR_Q queries x <= 4096 maxima x 4 B of shared memory
        if (cm_slots > 64) cm_slots = 64;
        if (cm_slots < 8) cm_slots = 8;
        static const int probe_div = getenv("ARCHI_TC_PROBE") ? atoi(getenv("ARCHI_TC_PROBE")) : 12;
        int per_vcta = (n_ctiles / probe_div) / nlists;
        if (per_vcta > cm_slots / 8) per_vcta = cm_slots / 8;
        if (per_vcta < 1) per_vcta = 1;
        if ((rc = ensure_buf(&w.chunkmax, &w.chunkmax_bytes, (size_t)grid * 2 * BM * cm_slots * sizeof(float))) != ARCHI_OK)
            return rc;
        if (s->timing) ARCHI_CUDA(cudaEventRecord(s->ws.ev0, st));
        cp.tile_begin = 0;
        cp.tile_end = per_vcta * nlists;
        cp.resume = 0;
        cp.maxima_only = 1;
        cp.chunkmax = w.chunkmax;
        cp.cm_slots = cm_slots;
        cp.launch_tag = (++w.launch_tag) & 0xfffu;
        kern<<<grid, NTHREADS, SMEM_BYTES, st>>>(tmap_q, tmap_c, cp);
        ARCHI_CHECK_LAUNCH();
        MaxThrParams mp;
        mp.chunkmax = w.chunkmax;
        mp.qt_count = qt_count;
        mp.ngroups = ngroups;
        mp.cm_slots = cm_slots;
        mp.used_slots = per_vcta * 8;
        mp.kprime = margin ? k : kprime;
        mp.nq = nq;
        mp.thr_g = w.thr_g;
        mp.qinfo = reinterpret_cast<const QInfo *>(w.qinfo);
        mp.margin = margin;
        // ARCHI_TC_THR4=1: four warps per query for k <= 32 (measured slower: 36 vs 23 us at config 2, +11 us per step
        // on a 125k-row shard), kept for experiments
        static const int thr4 = getenv("ARCHI_TC_THR4") ? atoi(getenv("ARCHI_TC_THR4")) : 0;
        mp.one_warp = thr4 ? 0 : 1;
        const size_t mt_smem = (size_t)MAXTHR_Q * (round_up(nlists * mp.used_slots, 32) + 4) * 4;
        if ((rc = set_dyn_smem_once((const void *)tc_maxima_threshold_kernel, (int)mt_smem)) != ARCHI_OK) return rc;
        tc_maxima_threshold_kernel<<<(nq + MAXTHR_Q - 1) / MAXTHR_Q, MAXTHR_THREADS, mt_smem, st>>>(mp);
        ARCHI_CHECK_LAUNCH();
        probed = true;
    }
    cp.maxima_only = 0;
    cp.chunkmax = w.chunkmax;
    cp.cm_slots = 8;
    int bounds[4] = {0, 0, 0, 0};
    int n_phases = 1;
    if (warm == 2 && n_ctiles >= 8 * ngroups) {
        bounds[n_phases++] = 2 * ngroups < 32 ? 2 * ngroups : 32;   // <= 32 x 256 keys per query: staged select
        // the main launch wants thresholds drawn from >= ~200 tiles (fewer than ~0.5 admitted columns
        // per warp and 32-column chunk); a phase change costs ~60 us, so only when the run is long enough
        int b2 = n_ctiles / 16 > 216 ? n_ctiles / 16 : 216;
        b2 = round_up(b2, 2 * ngroups);
        if (n_ctiles >= 4 * b2) bounds[n_phases++] = b2;
    }
    if (probed) {
        // Long scans: the probe's threshold (k' of ~probe rows) stays loose for the whole main launch
        // unless buffers fill.  Tighten it once from the candidates of the first 1/16 of the corpus.
        static const int p2div = getenv("ARCHI_TC_P2") ? atoi(getenv("ARCHI_TC_P2")) : 16;
        const int probe_tiles = cp.tile_end;
        if (p2div > 0 && n_ctiles / p2div >= 3 * probe_tiles) bounds[n_phases++] = round_up(n_ctiles / p2div, 2 * ngroups);
    }
    bounds[n_phases] = n_ctiles;
    ThresholdParams tp;
    tp.c.cand = reinterpret_cast<const uint2 *>(w.cand);
    tp.c.cand_cnt = w.cand_cnt;
    tp.c.qt_count = qt_count;
    tp.c.ngroups = ngroups;
    tp.c.cap = cap;
    tp.c.kprime = margin ? k : kprime;
    tp.thr_g = w.thr_g;
    tp.qinfo = reinterpret_cast<const QInfo *>(w.qinfo);
    tp.margin = margin;
    if (s->timing && !probed) ARCHI_CUDA(cudaEventRecord(s->ws.ev0, st));
    for (int ph = 0; ph < n_phases; ++ph) {
        cp.tile_begin = bounds[ph];
        cp.tile_end = bounds[ph + 1];
        cp.resume = ph > 0;
        cp.exit_cap = ph + 1 < n_phases ? cap - BN : cap;   // resumable launches leave room for a whole tile
        cp.launch_tag = (++w.launch_tag) & 0xfffu;
        kern<<<grid, NTHREADS, SMEM_BYTES, st>>>(tmap_q, tmap_c, cp);
        ARCHI_CHECK_LAUNCH();
        if (ph + 1 < n_phases) {
            tc_threshold_kernel<<<nq, SEL_THREADS, 0, st>>>(tp);
            ARCHI_CHECK_LAUNCH();
        }
    }
    if (s->timing) {
        ARCHI_CUDA(cudaEventRecord(s->ws.ev1, st));
        ARCHI_CUDA(cudaEventSynchronize(s->ws.ev1));
        float ms = 0.f;
        ARCHI_CUDA(cudaEventElapsedTime(&ms, s->ws.ev0, s->ws.ev1));
        *coarse_ms = ms;
    }

    // ---- 4. select + rescore + proof ----
    SelectParams sp;
    sp.corpus = s->data;
    sp.dtype = s->dtype;
    sp.dim = s->dim;
    sp.ld = s->ld;
    sp.metric = s->metric;
    sp.norm2 = s->norm2;
    sp.queries = q_dev;
    sp.qinfo = reinterpret_cast<const QInfo *>(w.qinfo);
    sp.cand = reinterpret_cast<const uint2 *>(w.cand);
    sp.cand_cnt = w.cand_cnt;
    sp.thr_g = w.thr_g;
    sp.qt_count = qt_count;
    sp.ngroups = ngroups;
    sp.cap = cap;
    sp.k = k;
    sp.kprime = margin ? k : kprime;
    sp.margin = margin;
    sp.out_scores = out_scores;
    sp.out_ids = reinterpret_cast<long long *>(out_ids);
    sp.id_offset = id_offset;
    sp.unverified = w.unverified;
    sp.n_unverified = w.unverified + nq_pad;
    sp.unv_list = w.unv_list;
    sp.max_sel = max_sel;
    sp.sticky = w.sticky_dev;
    // The key stage decides how many select CTAs fit on an SM (the kernel is a chain of dependent L2 / HBM latencies:
    // residency is its throughput).  A probed single-phase scan leaves a few hundred candidates per query (k times
    // the probe ratio, widened by the margin): 2048 slots -> 8 CTAs per SM, a batch of 1024 queries is one wave.
    // Long multi-phase scans keep thousands: 7680 slots.  A query that exceeds the stage takes the L2 re-walk path.
    sp.stage_cap = (probed && n_phases == 1) ? 2048 : SEL_STAGE;
    const size_t q_smem = (size_t)s->ld * sizeof(float) + (size_t)sp.stage_cap * sizeof(uint32_t);
    if (nq > 4 * s->sm_count) {
        if ((rc = set_dyn_smem_once((const void *)tc_select_kernel<128>, (int)q_smem)) != ARCHI_OK) return rc;
        tc_select_kernel<128><<<nq, 128, q_smem, st>>>(sp);
    } else {
        if ((rc = set_dyn_smem_once((const void *)tc_select_kernel<SEL_THREADS>, (int)q_smem)) != ARCHI_OK) return rc;
        tc_select_kernel<SEL_THREADS><<<nq, SEL_THREADS, q_smem, st>>>(sp);
    }
    ARCHI_CHECK_LAUNCH();
    // No host round trip here: the caller enqueues the device-driven rescue of the listed queries
    // (launch_rescue with qsel = w.unv_list, nsel = the counter) and reads the verdict when it next
    // synchronises anyway.
    *max_sel_out = max_sel;
    s->stats.grid = grid;
    s->stats.coarse_dtype = tf32 ? ARCHI_F32 : ARCHI_BF16;
    s->stats.coarse_launches = n_phases + (probed ? 1 : 0);
    return ARCHI_OK;
}

void free_tensor_workspace(TensorWorkspace &w)
{
    void *ptrs[] = {w.qstage, w.qinfo, w.thr_g, w.unverified, w.cand, w.cand_cnt, w.aux, w.max_norm2, w.shadow, w.chunkmax,
                    w.unv_list, w.sticky_dev, w.progress};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (w.h_verdict) cudaFreeHost(w.h_verdict);
    if (w.verdict_ev) cudaEventDestroy(w.verdict_ev);
    w = TensorWorkspace();
}

}  // namespace archi
