// topk.cuh -- warp-level register top-k (shuffle/ballot) and the warp multi-value reduction.
//
// A WarpTopK<M> is a sorted list of 32*M (key, id) entries distributed over the 32 lanes of a
// warp: rank r lives in slot r/32 of lane r%32.  Larger key = better; equal keys are ordered by
// the lower id, which is the tie rule the oracle uses.  All member functions are warp-collective
// and must be called with warp-uniform arguments.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <math_constants.h>

namespace archi {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ bool better(float ak, int ai, float bk, int bi)
{
    return ak > bk || (ak == bk && ai < bi);
}

template <int M>
struct WarpTopK {
    float key[M];
    int id[M];

    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int s = 0; s < M; ++s) {
            key[s] = -CUDART_INF_F;
            id[s] = INT_MAX;
        }
    }

    // Insert (nk, nid) keeping the best k entries in ranks [0, k).  Returns true when the list
    // changed.  NaN keys must be filtered by the caller.
    __device__ __forceinline__ bool insert(float nk, int nid, int k, int lane)
    {
        int p = 0;
#pragma unroll
        for (int s = 0; s < M; ++s)
            p += __popc(__ballot_sync(kFull, better(key[s], id[s], nk, nid)));
        if (p >= k) return false;
#pragma unroll
        for (int s = M - 1; s >= 0; --s) {
            float pk = __shfl_up_sync(kFull, key[s], 1);
            int pi = __shfl_up_sync(kFull, id[s], 1);
            if (s > 0) {
                float ck = __shfl_sync(kFull, key[s - 1], 31);
                int ci = __shfl_sync(kFull, id[s - 1], 31);
                if (lane == 0) {
                    pk = ck;
                    pi = ci;
                }
            }
            const int rank = s * 32 + lane;
            if (rank == p) {
                key[s] = nk;
                id[s] = nid;
            } else if (rank > p) {
                key[s] = pk;
                id[s] = pi;
            }
        }
        return true;
    }

    // The entry at rank k-1 (the admission threshold), broadcast to all lanes.
    __device__ __forceinline__ void threshold(int k, float &tk, int &ti) const
    {
        const int ss = (k - 1) >> 5, ll = (k - 1) & 31;
        float a = key[0];
        int b = id[0];
#pragma unroll
        for (int s = 1; s < M; ++s)
            if (ss == s) {
                a = key[s];
                b = id[s];
            }
        tk = __shfl_sync(kFull, a, ll);
        ti = __shfl_sync(kFull, b, ll);
    }
};

// Sum NV per-lane values across the warp with NV-1+log2(32/NV) shuffles (instead of 5*NV): at each
// stage half of the values move to the partner lane.  Lane l returns the total of value
// l >> (5 - log2(NV)).  v[] is clobbered.
template <int NV>
__device__ __forceinline__ float warp_multi_reduce(float (&v)[NV], int lane)
{
    static_assert(NV == 1 || NV == 2 || NV == 4 || NV == 8 || NV == 16 || NV == 32, "NV");
    int d = 16;
#pragma unroll
    for (int cnt = NV; cnt > 1; cnt >>= 1) {
        const bool upper = (lane & d) != 0;
#pragma unroll
        for (int i = 0; i < cnt / 2; ++i) {
            const float send = upper ? v[i] : v[i + cnt / 2];
            const float keep = upper ? v[i + cnt / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, d);
        }
        d >>= 1;
    }
    float x = v[0];
#pragma unroll
    for (; d >= 1; d >>= 1) x += __shfl_xor_sync(kFull, x, d);
    return x;
}

// order-preserving map float -> uint32 (larger float <=> larger uint) and back
__device__ __forceinline__ uint32_t fmap(float f)
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float funmap(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Block-level radix select over `total` uint32 keys in shared memory: returns the k-th largest key
// (1 <= k <= total) and, through n_gt, how many keys are strictly larger.  The bits above the
// highest bit in which the smallest and largest key differ are skipped (scores cluster in a narrow
// band, so the top byte(s) are usually common and would serialise every histogram update on one
// bin); below that, 8 bits per pass, one private histogram per warp.  s_hist needs
// 256 * (blockDim.x / 32) ints, s_misc 8 ints.  Every thread of the block must call it.
__device__ __forceinline__ uint32_t block_radix_kth(const uint32_t *keys, int total, int k, int *s_hist, int *s_misc,
                                                   int &n_gt)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarp = nthr >> 5;
    // block min / max
    uint32_t kmax = 0u, kmin = 0xffffffffu;
    for (int i = tid; i < total; i += nthr) {
        const uint32_t key = keys[i];
        kmax = max(kmax, key);
        kmin = min(kmin, key);
    }
    kmax = __reduce_max_sync(kFull, kmax);
    kmin = __reduce_min_sync(kFull, kmin);
    if (tid == 0) {
        s_misc[4] = 0;
        s_misc[5] = (int)0xffffffffu;
    }
    __syncthreads();
    if (lane == 0) {
        atomicMax(reinterpret_cast<unsigned int *>(&s_misc[4]), kmax);
        atomicMin(reinterpret_cast<unsigned int *>(&s_misc[5]), kmin);
    }
    __syncthreads();
    kmax = (uint32_t)s_misc[4];
    kmin = (uint32_t)s_misc[5];
    n_gt = 0;
    if (kmax == kmin) return kmax;                       // all keys equal
    int top = 32 - __clz(kmax ^ kmin);                   // bits [0, top) still have to be resolved
    uint32_t known = top >= 32 ? 0u : ~((1u << top) - 1u);
    uint32_t prefix = kmax & known;
    int need = k, above_total = 0;
    while (top > 0) {
        const int width = top < 8 ? top : 8;
        const int shift = top - width;
        const uint32_t dmask = (1u << width) - 1u;
        for (int i = tid; i < 256 * nwarp; i += nthr) s_hist[i] = 0;
        __syncthreads();
        int *myhist = s_hist + warp * 256;
        for (int i = tid; i < total; i += nthr) {
            const uint32_t key = keys[i];
            if ((key & known) == prefix) atomicAdd(&myhist[(key >> shift) & dmask], 1);
        }
        __syncthreads();
        for (int i = tid; i < 256; i += nthr) {          // fold the per-warp histograms into the first
            int sum = 0;
            for (int w = 0; w < nwarp; ++w) sum += s_hist[w * 256 + i];
            s_hist[i] = sum;
        }
        __syncthreads();
        if (warp == 0) {
            int mine = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) mine += s_hist[lane * 8 + j];
            int suf = mine;   // inclusive suffix sum over lanes (higher lanes = larger digits)
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_down_sync(kFull, suf, d);
                if (lane + d < 32) suf += o;
            }
            const int above = suf - mine;
            const unsigned who = __ballot_sync(kFull, above < need && suf >= need);
            if (lane == __ffs(who) - 1) {
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
                s_misc[2] = need - acc;   // rank inside the chosen bin
                s_misc[3] = acc;          // keys of this prefix bucket that are above the chosen bin
            }
        }
        __syncthreads();
        prefix |= (uint32_t)s_misc[1] << shift;
        known |= dmask << shift;
        need = s_misc[2];
        above_total += s_misc[3];
        top = shift;
        __syncthreads();
    }
    n_gt = above_total;
    return prefix;
}

// The 32 largest of keys[begin, end) as a sorted register list (rank r in lane r, 0 = no entry; keys are > 0), one warp:
// only keys above the current k-th entry enter -- about k ln(n / k) insertions, no histogram, no shared-memory atomics
// (scores cluster in a few histogram bins, which serialises radix passes).  Four 32-key chunks are tested per trip,
// so the common trip (no key beats the k-th entry) is four independent loads and four ballots.
__device__ __forceinline__ uint32_t warp_top32_small(const uint32_t *keys, int begin, int end, int k, int lane)
{
    uint32_t mine = 0u, thr = 0u;
    for (int i0 = begin; i0 < end; i0 += 128) {
        uint32_t key[4];
        unsigned cand[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32 + lane;
            key[u] = i < end ? keys[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) cand[u] = __ballot_sync(kFull, key[u] > thr);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            while (cand[u]) {
                const int src = __ffs(cand[u]) - 1;
                cand[u] &= cand[u] - 1;
                const uint32_t nk = __shfl_sync(kFull, key[u], src);
                if (nk <= thr) continue;                                  // the threshold rose since the ballot
                const int p = __popc(__ballot_sync(kFull, mine >= nk));   // entries that stay in front of it
                const uint32_t up = __shfl_up_sync(kFull, mine, 1);
                if (lane == p) mine = nk;
                else if (lane > p) mine = up;
                thr = __shfl_sync(kFull, mine, k - 1);
            }
        }
    }
    return mine;
}

// k-th largest of keys[0, total) for k <= 32 (0 when fewer than k keys are > 0), one warp.
__device__ __forceinline__ uint32_t warp_kth_small(const uint32_t *keys, int total, int k, int lane)
{
    const uint32_t mine = warp_top32_small(keys, 0, total, k, lane);
    return __shfl_sync(kFull, mine, k - 1);
}

// Warp-level variant of block_radix_kth: one warp, its own 256-bin histogram in shared memory, no
// block barriers.  keys[] (shared memory) are read lane-strided; every lane of the warp must call it.
__device__ __forceinline__ uint32_t warp_radix_kth(const uint32_t *keys, int total, int k, int *hist, int lane)
{
    uint32_t kmax = 0u, kmin = 0xffffffffu;
#pragma unroll 8
    for (int i = lane; i < total; i += 32) {
        const uint32_t key = keys[i];
        kmax = max(kmax, key);
        kmin = min(kmin, key);
    }
    kmax = __reduce_max_sync(kFull, kmax);
    kmin = __reduce_min_sync(kFull, kmin);
    if (kmax == kmin) return kmax;
    int top = 32 - __clz(kmax ^ kmin);
    uint32_t known = top >= 32 ? 0u : ~((1u << top) - 1u);
    uint32_t prefix = kmax & known;
    int need = k;
    while (top > 0) {
        const int width = top < 8 ? top : 8;
        const int shift = top - width;
        const uint32_t dmask = (1u << width) - 1u;
#pragma unroll
        for (int j = 0; j < 8; ++j) hist[j * 32 + lane] = 0;
        __syncwarp();
#pragma unroll 8
        for (int i = lane; i < total; i += 32) {
            const uint32_t key = keys[i];
            if ((key & known) == prefix) atomicAdd(&hist[(key >> shift) & dmask], 1);
        }
        __syncwarp();
        int h[8], mine = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            h[j] = hist[lane * 8 + j];
            mine += h[j];
        }
        int suf = mine;   // inclusive suffix sum over lanes (higher lanes = larger digits)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_down_sync(kFull, suf, d);
            if (lane + d < 32) suf += o;
        }
        const int above = suf - mine;
        const unsigned who = __ballot_sync(kFull, above < need && suf >= need);
        const int src = __ffs(who) - 1;
        int digit = lane * 8, acc = above;
        bool found = false;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
            if (!found) {
                if (acc + h[j] >= need) {
                    digit = lane * 8 + j;
                    found = true;
                } else {
                    acc += h[j];
                }
            }
        }
        digit = __shfl_sync(kFull, digit, src);
        acc = __shfl_sync(kFull, acc, src);
        prefix |= (uint32_t)digit << shift;
        known |= dmask << shift;
        need -= acc;
        top = shift;
        __syncwarp();
    }
    return prefix;
}

template <int NV>
struct Log2 {
    static constexpr int value = 1 + Log2<NV / 2>::value;
};
template <>
struct Log2<1> {
    static constexpr int value = 0;
};

}  // namespace archi
