// scan.cu -- streaming exact scorer with fused warp-register top-k (the small-batch path).
//
// Replaces, for up to kMaxQB queries per corpus pass, the reference's
//   SELECT emb <op> q AS distance ... ORDER BY distance ASC LIMIT k      (postgres_vectorstore.py:317-332)
// and hybrid_search's scored CTE + ORDER BY combined DESC LIMIT k           (postgres_vectorstore.py:435-457).
//
// HBM-bound by construction: every corpus byte is read exactly once per pass with 128-bit
// streaming loads (ld.global.nc.L1::no_allocate), R rows (R*512 B per lane-wide load group) in
// flight per warp, queries resident in shared memory, fp32 FMA accumulation, one warp-transposed
// reduction per R rows, and a register top-k per warp guarded by a running threshold, so no score
// ever goes back to memory.  Per-CTA lists are merged in shared memory, the per-CTA results by
// scan_finalize_kernel.
#include "common.cuh"
#include "topk.cuh"

namespace archi {

struct ScanParams {
    const void *corpus;
    long long n;
    int dim, ld;
    int metric;
    const float *queries;
    int nqb;
    int k;
    const float *norm2;
    const uint32_t *alive;
    const uint32_t *filter;
    int hybrid;
    const float *bias;
    long long bias_stride;
    float w_sem, w_bias;
    const float *cursor_key;
    const int *cursor_id;
    float *part_key;
    int *part_id;
    // rescue mode (RESCUE kernels only): the queries to scan are named by a DEVICE list, so that the launch
    // needs no host knowledge of how many there are: slots [0, min(*nsel, max_sel)) of qsel are row indices
    // into `queries`; the kernel takes them QB at a time and writes the partial lists of pass j at
    // part_* + j * pass_stride.  *nsel == 0 makes the launch a no-op.
    const int *qsel;
    const int *nsel;
    int max_sel;
    long long pass_stride;
};

template <typename T>
struct Elt;
template <>
struct Elt<float> {
    static constexpr int VEC = 4;
};
template <>
struct Elt<__nv_bfloat16> {
    static constexpr int VEC = 8;
};

__device__ __forceinline__ uint4 ldg_stream16(const void *p)
{
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
        : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
        : "l"(p));
    return r;
}

// dot / squared-difference of one 16-byte corpus vector against the matching query slice.
template <typename T, bool L2>
__device__ __forceinline__ float accum16(float acc, const uint4 d, const float4 qa, const float4 qb)
{
    if constexpr (Elt<T>::VEC == 4) {
        const float e0 = __uint_as_float(d.x), e1 = __uint_as_float(d.y);
        const float e2 = __uint_as_float(d.z), e3 = __uint_as_float(d.w);
        if constexpr (L2) {
            float t;
            t = e0 - qa.x; acc = fmaf(t, t, acc);
            t = e1 - qa.y; acc = fmaf(t, t, acc);
            t = e2 - qa.z; acc = fmaf(t, t, acc);
            t = e3 - qa.w; acc = fmaf(t, t, acc);
        } else {
            acc = fmaf(e0, qa.x, acc);
            acc = fmaf(e1, qa.y, acc);
            acc = fmaf(e2, qa.z, acc);
            acc = fmaf(e3, qa.w, acc);
        }
    } else {
        // 8 bf16: element 2i in the low half of word i, element 2i+1 in the high half
        const float e0 = __uint_as_float(d.x << 16), e1 = __uint_as_float(d.x & 0xffff0000u);
        const float e2 = __uint_as_float(d.y << 16), e3 = __uint_as_float(d.y & 0xffff0000u);
        const float e4 = __uint_as_float(d.z << 16), e5 = __uint_as_float(d.z & 0xffff0000u);
        const float e6 = __uint_as_float(d.w << 16), e7 = __uint_as_float(d.w & 0xffff0000u);
        if constexpr (L2) {
            float t;
            t = e0 - qa.x; acc = fmaf(t, t, acc);
            t = e1 - qa.y; acc = fmaf(t, t, acc);
            t = e2 - qa.z; acc = fmaf(t, t, acc);
            t = e3 - qa.w; acc = fmaf(t, t, acc);
            t = e4 - qb.x; acc = fmaf(t, t, acc);
            t = e5 - qb.y; acc = fmaf(t, t, acc);
            t = e6 - qb.z; acc = fmaf(t, t, acc);
            t = e7 - qb.w; acc = fmaf(t, t, acc);
        } else {
            acc = fmaf(e0, qa.x, acc);
            acc = fmaf(e1, qa.y, acc);
            acc = fmaf(e2, qa.z, acc);
            acc = fmaf(e3, qa.w, acc);
            acc = fmaf(e4, qb.x, acc);
            acc = fmaf(e5, qb.y, acc);
            acc = fmaf(e6, qb.z, acc);
            acc = fmaf(e7, qb.w, acc);
        }
    }
    return acc;
}

// Merge `nlists` sorted lists of LEN = 32*M entries staged in shared memory into `res`.
template <int M>
__device__ __forceinline__ void merge_staged(const float *skey, const int *sid, int nlists, int k,
                                             int lane, WarpTopK<M> &res)
{
    float tk = -CUDART_INF_F;
    int ti = INT_MAX;
    for (int l = 0; l < nlists; ++l) {
#pragma unroll 1
        for (int c = 0; c < M; ++c) {
            const float ek = skey[l * 32 * M + c * 32 + lane];
            const int ei = sid[l * 32 * M + c * 32 + lane];
            unsigned cand = __ballot_sync(kFull, better(ek, ei, tk, ti));
            if (cand == 0) break;  // lists are sorted: the rest of this list is worse
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

template <typename T, int QB, int R, int M, bool L2, bool RESCUE = false>
__global__ void __launch_bounds__(kScanThreads, (QB >= 8 ? 1 : 2)) scan_topk_kernel(const ScanParams p)
{
    constexpr int VEC = Elt<T>::VEC;
    constexpr int NV = QB * R;
    constexpr int SH = 5 - Log2<NV>::value;
    constexpr int WARPS = kScanThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sq = reinterpret_cast<float *>(smem_raw);
    __shared__ float s_qrn[QB];
    __shared__ float s_ckey[QB];
    __shared__ int s_cid[QB];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nvec = p.ld / VEC;         // 16-byte vectors per row
    const int J = (nvec + 31) >> 5;      // lane-wide load groups per row
    const int qstride = J * 32 * VEC;    // floats per staged query (zero padded)

    int n_sel = p.nqb;
    if constexpr (RESCUE) {
        n_sel = *p.nsel;
        if (n_sel > p.max_sel) n_sel = p.max_sel;
    }
    // one iteration unless RESCUE: the selected queries are taken QB at a time
#pragma unroll 1
    for (int q0 = 0; q0 < n_sel; q0 += QB) {
    const int nqb = n_sel - q0 < QB ? n_sel - q0 : QB;
    float *const part_key = p.part_key + (RESCUE ? (q0 / QB) * p.pass_stride : 0);
    int *const part_id = p.part_id + (RESCUE ? (q0 / QB) * p.pass_stride : 0);

    // ---- stage the queries in shared memory (bf16 rows: split into lo/hi float4 planes so that
    //      both LDS.128 of a lane are conflict-free) -------------------------------------------
    for (int idx = tid; idx < QB * qstride; idx += kScanThreads) {
        const int q = idx / qstride, e = idx - q * qstride;
        const size_t qrow = RESCUE ? (size_t)(q < nqb ? p.qsel[q0 + q] : 0) : (size_t)q;
        const float val = (q < nqb && e < p.dim) ? p.queries[qrow * p.dim + e] : 0.f;
        int pos = e;
        if (VEC == 8) {
            const int v = e >> 3, c = e & 7;
            pos = (c >> 2) * (qstride >> 1) + v * 4 + (c & 3);
        }
        sq[q * qstride + pos] = val;
    }
    __syncthreads();
    for (int q = warp; q < QB; q += WARPS) {
        float ss = 0.f;
        for (int e = lane; e < qstride; e += 32) {
            const float x = sq[q * qstride + e];
            ss = fmaf(x, x, ss);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) ss += __shfl_xor_sync(kFull, ss, d);
        if (lane == 0) {
            s_qrn[q] = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
            const bool cur = p.cursor_key != nullptr && q < nqb;
            s_ckey[q] = cur ? p.cursor_key[q] : CUDART_INF_F;
            s_cid[q] = cur ? p.cursor_id[q] : -1;
        }
    }
    __syncthreads();

    // ---- per-lane role after the transposed reduction: value index = lane >> SH -------------
    const int my_idx = lane >> SH;
    const int my_r = my_idx / QB, my_q = my_idx % QB;
    const bool owner = (lane & ((1 << SH) - 1)) == 0 && my_q < nqb;
    const float my_qrn = s_qrn[my_q];
    const float my_ckey = s_ckey[my_q];
    const int my_cid = s_cid[my_q];
    float my_tk = -CUDART_INF_F;  // admission threshold of my query's list
    int my_ti = INT_MAX;

    WarpTopK<M> lists[QB];
#pragma unroll
    for (int q = 0; q < QB; ++q) lists[q].init();

    const long long ngroups = (p.n + R - 1) / R;
    const long long wstride = (long long)gridDim.x * WARPS;
    const unsigned char *corpus = reinterpret_cast<const unsigned char *>(p.corpus);
    const size_t row_bytes = (size_t)p.ld * sizeof(T);
    const float4 *sq4 = reinterpret_cast<const float4 *>(sq);
    const int q4stride = qstride >> 2;
    const int hi4 = qstride >> 3;  // float4 offset of the hi plane (bf16)

    for (long long g = (long long)blockIdx.x * WARPS + warp; g < ngroups; g += wstride) {
        const long long row0 = g * R;
        // rows of this group that exist and pass the tombstone / filter masks (R divides 32 and
        // row0 is a multiple of R, so the R bits sit in one mask word)
        unsigned ok = (1u << R) - 1u;
        if (row0 + R > p.n) ok = (1u << (int)(p.n - row0)) - 1u;
        if (p.alive) ok &= p.alive[row0 >> 5] >> (row0 & 31);
        if (p.filter) ok &= p.filter[row0 >> 5] >> (row0 & 31);
        if (ok == 0) continue;

        float acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.f;

        const unsigned char *gbase = corpus + (size_t)row0 * row_bytes + (size_t)lane * 16;
#pragma unroll 2
        for (int j = 0; j < J; ++j) {
            const bool inrow = (j * 32 + lane) < nvec;
            uint4 d[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (inrow && ((ok >> r) & 1u))
                    d[r] = ldg_stream16(gbase + (size_t)r * row_bytes + (size_t)j * 512);
                else
                    d[r] = make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int q = 0; q < QB; ++q) {
                float4 qa, qb;
                if (VEC == 4) {
                    qa = sq4[q * q4stride + j * 32 + lane];
                    qb = qa;
                } else {
                    qa = sq4[q * q4stride + j * 32 + lane];
                    qb = sq4[q * q4stride + hi4 + j * 32 + lane];
                }
#pragma unroll
                for (int r = 0; r < R; ++r)
                    acc[r * QB + q] = accum16<T, L2>(acc[r * QB + q], d[r], qa, qb);
            }
        }

        const float total = warp_multi_reduce<NV>(acc, lane);

        // ---- key of (row0 + my_r, my_q): larger is better ---------------------------------
        const long long row = row0 + my_r;
        const bool valid = owner && ((ok >> my_r) & 1u);
        float key = total;
        if (valid) {
            if (p.hybrid) {
                // semantic = 1.0 - (emb <op> q) for every metric (postgres_vectorstore.py:441)
                float sem;
                if (L2) {
                    sem = 1.0f - sqrtf(total);
                } else if (p.metric == ARCHI_COSINE) {
                    const float n2 = p.norm2[row];
                    const float sim = total * my_qrn * (n2 > 0.f ? 1.0f / sqrtf(n2) : 0.f);
                    sem = fminf(1.0f, fmaxf(-1.0f, sim));
                } else {
                    sem = 1.0f + total;
                }
                const float bm = p.bias ? p.bias[(size_t)my_q * p.bias_stride + row] : 0.f;
                key = fmaf(sem, p.w_sem, bm * p.w_bias);
            } else if (L2) {
                key = -total;
            } else if (p.metric == ARCHI_COSINE) {
                const float n2 = p.norm2[row];
                key = n2 > 0.f ? fminf(1.0f, fmaxf(-1.0f, total * my_qrn * (1.0f / sqrtf(n2))))
                               : -CUDART_INF_F;
            }
        }
        const int rid = (int)row;
        bool pass = valid && better(key, rid, my_tk, my_ti) && better(my_ckey, my_cid, key, rid);
        unsigned cand = __ballot_sync(kFull, pass);
        while (cand) {
            const int src = __ffs(cand) - 1;
            cand &= cand - 1;
            const float nk = __shfl_sync(kFull, key, src);
            const int ni = __shfl_sync(kFull, rid, src);
            const int qq = (src >> SH) % QB;
            float tk = 0.f;
            int ti = 0;
            bool changed = false;
#pragma unroll
            for (int q = 0; q < QB; ++q) {
                if (qq == q) {
                    changed = lists[q].insert(nk, ni, p.k, lane);
                    if (changed) lists[q].threshold(p.k, tk, ti);
                }
            }
            if (changed && my_q == qq) {
                my_tk = tk;
                my_ti = ti;
            }
        }
    }

    // ---- block merge: stage the 8 warp lists in shared memory (re-using the query area) ------
    __syncthreads();
    constexpr int LEN = 32 * M;
    float *skey = reinterpret_cast<float *>(smem_raw);
    int *sid = reinterpret_cast<int *>(smem_raw) + WARPS * QB * LEN;
#pragma unroll
    for (int q = 0; q < QB; ++q) {
#pragma unroll
        for (int s = 0; s < M; ++s) {
            skey[(q * WARPS + warp) * LEN + s * 32 + lane] = lists[q].key[s];
            sid[(q * WARPS + warp) * LEN + s * 32 + lane] = lists[q].id[s];
        }
    }
    __syncthreads();
    for (int q = warp; q < nqb; q += WARPS) {
        WarpTopK<M> res;
        res.init();
        merge_staged<M>(skey + q * WARPS * LEN, sid + q * WARPS * LEN, WARPS, p.k, lane, res);
        float *ok_ = part_key + ((size_t)blockIdx.x * kMaxQB + q) * kMaxListK;
        int *oi_ = part_id + ((size_t)blockIdx.x * kMaxQB + q) * kMaxListK;
#pragma unroll
        for (int s = 0; s < M; ++s) {
            const int rank = s * 32 + lane;
            if (rank < p.k) {
                ok_[rank] = res.key[s];
                oi_[rank] = res.id[s];
            }
        }
    }
    __syncthreads();      // the staged lists share the query area of the next iteration
    }
}

// One CTA per query: merge the per-CTA lists, convert keys to the reference's score convention
// (postgres_vectorstore.py:361) and write ranks [col0, col0 + k) of the outputs.
struct FinalizeParams {
    const float *part_key;
    const int *part_id;
    int grid;
    int k;        // entries per partial list / produced this pass
    int k_total;  // row stride of the outputs
    int col0;
    int metric, hybrid;
    float *out_scores;
    long long *out_ids;
    long long id_offset;
    float *cursor_key_out;
    int *cursor_id_out;
    // rescue mode: CTA i finalises slot i of the device list (see ScanParams), writes output row qsel[i]
    const int *qsel;
    const int *nsel;
    int max_sel;
    long long pass_stride;
};

// Radix select instead of list insertion: the grid*k candidate keys are staged in shared memory,
// the k-th largest key T is found in four 8-bit passes, everything above T survives together with
// the lowest-id entries equal to T, and the <= k survivors are ranked by (key desc, id asc).
constexpr int kFinKeep = kMaxListK;      // survivors (k <= 128)
constexpr int kFinEq = 1024;             // entries tying the k-th key that are considered for the tie rule

__global__ void __launch_bounds__(kScanThreads) scan_finalize_kernel(const FinalizeParams p)
{
    extern __shared__ uint32_t s_keys[];             // [grid * k] mapped keys
    __shared__ int s_hist[256 * (kScanThreads / 32)];
    __shared__ int s_misc[8];
    __shared__ int s_nk, s_ne;
    __shared__ float s_skey[kFinKeep];
    __shared__ int s_sid[kFinKeep];
    __shared__ int s_eid[kFinEq];
    const int tid = threadIdx.x;
    int q = blockIdx.x;            // which partial list of each CTA
    int out_row = blockIdx.x;      // which row of the outputs
    size_t pass_off = 0;
    if (p.qsel) {
        int n_sel = *p.nsel;
        if (n_sel > p.max_sel) n_sel = p.max_sel;
        if ((int)blockIdx.x >= n_sel) return;
        out_row = p.qsel[blockIdx.x];
        q = blockIdx.x % kMaxQB;
        pass_off = (size_t)(blockIdx.x / kMaxQB) * (size_t)p.pass_stride;
    }
    const int total = p.grid * p.k;
    auto slot_of = [&](int i) {
        const int b = i / p.k, r = i - b * p.k;
        return pass_off + ((size_t)b * kMaxQB + q) * kMaxListK + r;
    };
    if (tid == 0) {
        s_nk = 0;
        s_ne = 0;
    }
#pragma unroll 4
    for (int i = tid; i < total; i += kScanThreads) s_keys[i] = fmap(p.part_key[slot_of(i)]);
    __syncthreads();
    const int kk = p.k < total ? p.k : total;
    int n_gt = 0;
    const uint32_t T = block_radix_kth(s_keys, total, kk, s_hist, s_misc, n_gt);
    const int need_eq = kk - n_gt;

    // survivors: key > T, and the need_eq lowest ids among key == T
    for (int i = tid; i < total; i += kScanThreads) {
        const uint32_t u = s_keys[i];
        if (u > T) {
            const int s = atomicAdd(&s_nk, 1);
            s_skey[s] = funmap(u);
            s_sid[s] = p.part_id[slot_of(i)];
        } else if (u == T) {
            const int s = atomicAdd(&s_ne, 1);
            if (s < kFinEq) s_eid[s] = p.part_id[slot_of(i)];
        }
    }
    __syncthreads();
    const int ne = s_ne < kFinEq ? s_ne : kFinEq;
    for (int i = tid; i < ne; i += kScanThreads) {
        const int mine = s_eid[i];
        int rank = 0;
        for (int j = 0; j < ne; ++j) rank += (s_eid[j] < mine || (s_eid[j] == mine && j < i)) ? 1 : 0;
        if (rank < need_eq) {
            s_skey[n_gt + rank] = funmap(T);
            s_sid[n_gt + rank] = mine;
        }
    }
    __syncthreads();
    const int nk = n_gt + (need_eq < ne ? need_eq : ne);

    for (int i = tid; i < p.k; i += kScanThreads) {
        if (i < nk) {
            const float key = s_skey[i];
            const int id = s_sid[i];
            int rank = 0;
            for (int j = 0; j < nk; ++j)
                rank += (better(s_skey[j], s_sid[j], key, id) || (s_skey[j] == key && s_sid[j] == id && j < i)) ? 1 : 0;
            const bool empty = id == INT_MAX;
            float score;
            if (empty) score = CUDART_NAN_F;
            else if (p.hybrid || p.metric == ARCHI_COSINE) score = key;
            else if (p.metric == ARCHI_L2) score = sqrtf(fmaxf(-key, 0.f));
            else score = -key;
            const size_t o = (size_t)out_row * p.k_total + p.col0 + rank;
            p.out_scores[o] = score;
            p.out_ids[o] = empty ? -1ll : (long long)id + p.id_offset;
            if (rank == p.k - 1 && p.cursor_key_out) {
                p.cursor_key_out[q] = key;
                p.cursor_id_out[q] = id;
            }
        } else {
            const size_t o = (size_t)out_row * p.k_total + p.col0 + i;
            p.out_scores[o] = CUDART_NAN_F;
            p.out_ids[o] = -1ll;
            if (i == p.k - 1 && p.cursor_key_out) {
                p.cursor_key_out[q] = -CUDART_INF_F;
                p.cursor_id_out[q] = INT_MAX;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef void (*scan_fn_t)(const ScanParams);

template <typename T, int QB, int R>
static scan_fn_t pick_m_l2(int M, bool l2)
{
    if (M == 1) return l2 ? scan_topk_kernel<T, QB, R, 1, true> : scan_topk_kernel<T, QB, R, 1, false>;
    return l2 ? scan_topk_kernel<T, QB, R, 4, true> : scan_topk_kernel<T, QB, R, 4, false>;
}

template <typename T>
static scan_fn_t pick_qb(int QB, int M, bool l2)
{
    switch (QB) {
        case 1: return pick_m_l2<T, 1, 8>(M, l2);
        case 2: return pick_m_l2<T, 2, 4>(M, l2);
        case 4: return pick_m_l2<T, 4, 4>(M, l2);
        default: return pick_m_l2<T, 8, 4>(M, l2);
    }
}

static int qb_for(int nqb) { return nqb <= 1 ? 1 : nqb <= 2 ? 2 : nqb <= 4 ? 4 : 8; }

int launch_scan(archi_store *s, const ScanArgs &a, cudaStream_t st, int *grid_out)
{
    ARCHI_REQUIRE(a.nqb >= 1 && a.nqb <= kMaxQB, "scan: nqb=%d out of range", a.nqb);
    ARCHI_REQUIRE(a.k >= 1 && a.k <= kMaxListK, "scan: k=%d out of range", a.k);
    const int QB = qb_for(a.nqb);
    const int R = QB == 1 ? 8 : 4;
    const int M = a.k <= 32 ? 1 : 4;
    const bool l2 = a.metric == ARCHI_L2;
    const int VEC = a.dtype == ARCHI_BF16 ? 8 : 4;
    const int nvec = a.ld / VEC;
    const int J = (nvec + 31) / 32;
    const size_t q_bytes = (size_t)QB * J * 32 * VEC * sizeof(float);
    const size_t m_bytes = (size_t)(kScanThreads / 32) * QB * 32 * M * 8;
    const size_t smem = q_bytes > m_bytes ? q_bytes : m_bytes;
    ARCHI_REQUIRE(smem <= 200 * 1024, "scan: dim=%d needs %zu B of shared memory per CTA", a.dim, smem);

    scan_fn_t fn = a.dtype == ARCHI_BF16 ? pick_qb<__nv_bfloat16>(QB, M, l2) : pick_qb<float>(QB, M, l2);
    ARCHI_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    ARCHI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fn, kScanThreads, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    const long long ngroups = (a.n + R - 1) / R;
    long long want = (ngroups + (kScanThreads / 32) - 1) / (kScanThreads / 32);
    long long grid = (long long)s->sm_count * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;

    // partial-list workspace
    if (s->ws.part_grid < grid) {
        if (s->ws.part_key) cudaFree(s->ws.part_key);
        if (s->ws.part_id) cudaFree(s->ws.part_id);
        s->ws.part_key = nullptr;
        s->ws.part_id = nullptr;
        const long long cap = (long long)s->sm_count * 4;
        const long long g = cap > grid ? cap : grid;
        ARCHI_CUDA(cudaMalloc(&s->ws.part_key, (size_t)g * kMaxQB * kMaxListK * sizeof(float)));
        ARCHI_CUDA(cudaMalloc(&s->ws.part_id, (size_t)g * kMaxQB * kMaxListK * sizeof(int)));
        s->ws.part_grid = (int)g;
    }

    ScanParams p;
    p.corpus = a.corpus;
    p.n = a.n;
    p.dim = a.dim;
    p.ld = a.ld;
    p.metric = a.metric;
    p.queries = a.queries;
    p.nqb = a.nqb;
    p.k = a.k;
    p.norm2 = a.norm2;
    p.alive = a.alive;
    p.filter = a.filter;
    p.hybrid = a.hybrid;
    p.bias = a.bias;
    p.bias_stride = a.bias_stride;
    p.w_sem = a.w_sem;
    p.w_bias = a.w_bias;
    p.cursor_key = a.cursor_key;
    p.cursor_id = a.cursor_id;
    p.part_key = s->ws.part_key;
    p.part_id = s->ws.part_id;
    p.qsel = nullptr;
    p.nsel = nullptr;
    p.max_sel = 0;
    p.pass_stride = 0;
    fn<<<(unsigned)grid, kScanThreads, smem, st>>>(p);
    ARCHI_CHECK_LAUNCH();
    *grid_out = (int)grid;
    return ARCHI_OK;
}

template <typename T>
static scan_fn_t pick_rescue(int M, bool l2)
{
    if (M == 1) return l2 ? scan_topk_kernel<T, 8, 4, 1, true, true> : scan_topk_kernel<T, 8, 4, 1, false, true>;
    return l2 ? scan_topk_kernel<T, 8, 4, 4, true, true> : scan_topk_kernel<T, 8, 4, 4, false, true>;
}

// Device-driven exact re-scan of the queries named by a device list (the tensor path's unproven queries):
// two launches that cost a few microseconds when the list is empty and need no host round trip.
int launch_rescue(archi_store *s, const ScanArgs &a, const int *qsel_dev, const int *nsel_dev, int max_sel,
                  float *out_scores, int64_t *out_ids, int64_t id_offset, cudaStream_t st)
{
    ARCHI_REQUIRE(a.k >= 1 && a.k <= kMaxListK, "rescue: k=%d out of range", a.k);
    ARCHI_REQUIRE(max_sel >= 1 && max_sel % kMaxQB == 0, "rescue: max_sel=%d must be a positive multiple of %d", max_sel, kMaxQB);
    constexpr int QB = 8, R = 4;
    const int M = a.k <= 32 ? 1 : 4;
    const bool l2 = a.metric == ARCHI_L2;
    const int VEC = a.dtype == ARCHI_BF16 ? 8 : 4;
    const int J = (a.ld / VEC + 31) / 32;
    const size_t q_bytes = (size_t)QB * J * 32 * VEC * sizeof(float);
    const size_t m_bytes = (size_t)(kScanThreads / 32) * QB * 32 * M * 8;
    const size_t smem = q_bytes > m_bytes ? q_bytes : m_bytes;
    ARCHI_REQUIRE(smem <= 200 * 1024, "rescue: dim=%d needs %zu B of shared memory per CTA", a.dim, smem);
    scan_fn_t fn = a.dtype == ARCHI_BF16 ? pick_rescue<__nv_bfloat16>(M, l2) : pick_rescue<float>(M, l2);
    ARCHI_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ngroups = (a.n + R - 1) / R;
    long long grid = s->sm_count;      // one CTA per SM (launch bounds of the QB = 8 kernels)
    const long long want = (ngroups + (kScanThreads / 32) - 1) / (kScanThreads / 32);
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    const int passes = max_sel / kMaxQB;
    const long long pass_stride = grid * kMaxQB * kMaxListK;
    Workspace &w = s->ws;
    if (w.resc_elems < pass_stride * passes) {
        if (w.resc_key) cudaFree(w.resc_key);
        if (w.resc_id) cudaFree(w.resc_id);
        w.resc_key = nullptr;
        w.resc_id = nullptr;
        w.resc_elems = 0;
        ARCHI_CUDA(cudaMalloc(&w.resc_key, (size_t)pass_stride * passes * sizeof(float)));
        ARCHI_CUDA(cudaMalloc(&w.resc_id, (size_t)pass_stride * passes * sizeof(int)));
        w.resc_elems = pass_stride * passes;
    }
    ScanParams p;
    p.corpus = a.corpus;
    p.n = a.n;
    p.dim = a.dim;
    p.ld = a.ld;
    p.metric = a.metric;
    p.queries = a.queries;
    p.nqb = 0;
    p.k = a.k;
    p.norm2 = a.norm2;
    p.alive = a.alive;
    p.filter = a.filter;
    p.hybrid = 0;
    p.bias = nullptr;
    p.bias_stride = 0;
    p.w_sem = 1.f;
    p.w_bias = 0.f;
    p.cursor_key = nullptr;
    p.cursor_id = nullptr;
    p.part_key = w.resc_key;
    p.part_id = w.resc_id;
    p.qsel = qsel_dev;
    p.nsel = nsel_dev;
    p.max_sel = max_sel;
    p.pass_stride = pass_stride;
    fn<<<(unsigned)grid, kScanThreads, smem, st>>>(p);
    ARCHI_CHECK_LAUNCH();

    FinalizeParams f;
    f.part_key = w.resc_key;
    f.part_id = w.resc_id;
    f.grid = (int)grid;
    f.k = a.k;
    f.k_total = a.k;
    f.col0 = 0;
    f.metric = a.metric;
    f.hybrid = 0;
    f.out_scores = out_scores;
    f.out_ids = reinterpret_cast<long long *>(out_ids);
    f.id_offset = id_offset;
    f.cursor_key_out = nullptr;
    f.cursor_id_out = nullptr;
    f.qsel = qsel_dev;
    f.nsel = nsel_dev;
    f.max_sel = max_sel;
    f.pass_stride = pass_stride;
    const size_t fsmem = (size_t)grid * a.k * sizeof(uint32_t);
    ARCHI_REQUIRE(fsmem <= 200 * 1024, "rescue: %zu B of candidate keys do not fit in shared memory", fsmem);
    if (fsmem > 40 * 1024)
        ARCHI_CUDA(cudaFuncSetAttribute((const void *)scan_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
    scan_finalize_kernel<<<max_sel, kScanThreads, fsmem, st>>>(f);
    ARCHI_CHECK_LAUNCH();
    return ARCHI_OK;
}

int launch_scan_finalize(archi_store *s, const ScanArgs &a, int grid, int k_total, int col0,
                         float *out_scores, int64_t *out_ids, int64_t id_offset,
                         float *cursor_key_out, int *cursor_id_out, cudaStream_t st)
{
    FinalizeParams p;
    p.part_key = s->ws.part_key;
    p.part_id = s->ws.part_id;
    p.grid = grid;
    p.k = a.k;
    p.k_total = k_total;
    p.col0 = col0;
    p.metric = a.metric;
    p.hybrid = a.hybrid;
    p.out_scores = out_scores;
    p.out_ids = reinterpret_cast<long long *>(out_ids);
    p.id_offset = id_offset;
    p.cursor_key_out = cursor_key_out;
    p.cursor_id_out = cursor_id_out;
    p.qsel = nullptr;
    p.nsel = nullptr;
    p.max_sel = 0;
    p.pass_stride = 0;
    const size_t smem = (size_t)grid * a.k * sizeof(uint32_t);
    ARCHI_REQUIRE(smem <= 200 * 1024, "scan_finalize: %zu B of candidate keys do not fit in shared memory", smem);
    if (smem > 40 * 1024)
        ARCHI_CUDA(cudaFuncSetAttribute((const void *)scan_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    scan_finalize_kernel<<<a.nqb, kScanThreads, smem, st>>>(p);
    ARCHI_CHECK_LAUNCH();
    return ARCHI_OK;
}

}  // namespace archi
