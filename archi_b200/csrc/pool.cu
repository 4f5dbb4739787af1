// pool.cu -- fused masked mean-pool + L2-normalise + cast on the encoder's last hidden state.
//
// Replaces the tail of Embeddings.embed_documents / embed_query (reference call sites
// manager.py:373, postgres_vectorstore.py:143,245,390), i.e. sentence-transformers'
// Pooling(mean) + Normalize modules [external]:
//     pooled = sum_t h[t]*m[t] / max(sum_t m[t], 1e-9);   out = pooled / max(|pooled|_2, 1e-12)
// HBM-bound: the [L, H] slab of a sequence is read once (tokens behind the last live one are never fetched, masked
// tokens inside are never accumulated), the pooled row never leaves the SM, and the normalised row is written once in
// each requested format -- optionally straight into the tail of the corpus matrix together with its |row|^2 and live
// bit, so add_documents needs no further kernel.  Two kernels:
//   * pool_ring_kernel (batches that fill the machine): persistent CTAs, a producer warp streams the slabs of the CTA's
//     sequences through a shared-memory ring with cp.async.bulk (the TMA engine: deep prefetch at no register cost,
//     and the ring keeps filling with the NEXT sequence while the consumer warps reduce / normalise / write the
//     current one), eight consumer warps accumulate from shared memory;
//   * pool_normalize_kernel (small batches, rows too wide for the ring): one CTA per sequence, 16-byte streaming loads.
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace archi {

// CTA width: 256 threads for a handful of sequences (more loads in flight per sequence), 128 for large batches --
// at 40-48 registers a 256-thread CTA fits 6 times on an SM, so 1024 sequences would run as 1.15 waves (the last 136
// CTAs alone on the machine); 128-thread CTAs fit 10 times: one balanced wave.

struct PoolParams {
    const void *hidden;
    const void *mask;
    int L, H;
    // sinks (any may be null)
    void *store_rows;       // row b -> store_rows + b*store_ld elements of store dtype
    int store_is_bf16;
    int store_ld;
    float *store_norm2;     // [B]
    uint32_t *alive;        // tombstone bitmask of the store, indexed by first_row + b
    long long first_row;
    __nv_bfloat16 *out_bf16;  // [B, H]
    float *out_f32;           // [B, H]
};

template <typename MT>
__device__ __forceinline__ float mask_at(const void *mask, size_t i)
{
    return (float)reinterpret_cast<const MT *>(mask)[i];
}

// barrier of the NT threads that run the epilogue: the whole CTA, or the consumer warps of the ring kernel (named)
template <int NT, bool NAMED>
__device__ __forceinline__ void pool_sync()
{
    if constexpr (NAMED) ptx::named_bar_sync(1, NT);
    else __syncthreads();
}

// Epilogue of one sequence, run by threads tid = 0..NT-1: spart[LS][H] holds the masked sums of the token slices,
// msum = max(sum of the mask, 1e-9).  spart is free again when it returns (trailing barrier).
template <int NT, bool NAMED>
__device__ __forceinline__ void pool_finish(const PoolParams &p, int b, int tid, int LS, float *spart, float *s_red,
                                            float *s_scalar, float msum)
{
    const int lane = tid & 31, warp = tid >> 5;
    // reduce the token slices, divide by the mask sum, accumulate |pooled|^2
    const float inv_m = 1.0f / msum;
    float ss = 0.f;
    for (int h = tid; h < p.H; h += NT) {
        float v = 0.f;
        for (int l = 0; l < LS; ++l) v += spart[(size_t)l * p.H + h];
        v *= inv_m;
        ss = fmaf(v, v, ss);
        spart[h] = v;  // slice 0 is only read by this same thread for column h
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    pool_sync<NT, NAMED>();
    if (lane == 0) s_red[warp] = ss;
    pool_sync<NT, NAMED>();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < NT / 32; ++w) t += s_red[w];
        s_scalar[1] = 1.0f / fmaxf(sqrtf(t), 1e-12f);
    }
    pool_sync<NT, NAMED>();
    const float inv_n = s_scalar[1];

    // write the normalised row to every sink; |stored row|^2 uses the values as stored
    float stored_ss = 0.f;
    for (int h = tid; h < p.H; h += NT) {
        const float v = spart[h] * inv_n;
        if (p.out_f32) p.out_f32[(size_t)b * p.H + h] = v;
        if (p.out_bf16) p.out_bf16[(size_t)b * p.H + h] = __float2bfloat16_rn(v);
        if (p.store_rows) {
            if (p.store_is_bf16) {
                const __nv_bfloat16 o = __float2bfloat16_rn(v);
                reinterpret_cast<__nv_bfloat16 *>(p.store_rows)[(size_t)b * p.store_ld + h] = o;
                const float f = __bfloat162float(o);
                stored_ss = fmaf(f, f, stored_ss);
            } else {
                reinterpret_cast<float *>(p.store_rows)[(size_t)b * p.store_ld + h] = v;
                stored_ss = fmaf(v, v, stored_ss);
            }
        }
    }
    if (p.store_rows) {
        // zero the row padding [H, ld)
        for (int h = p.H + tid; h < p.store_ld; h += NT) {
            if (p.store_is_bf16)
                reinterpret_cast<__nv_bfloat16 *>(p.store_rows)[(size_t)b * p.store_ld + h] =
                    __float2bfloat16_rn(0.f);
            else
                reinterpret_cast<float *>(p.store_rows)[(size_t)b * p.store_ld + h] = 0.f;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) stored_ss += __shfl_xor_sync(0xffffffffu, stored_ss, d);
        pool_sync<NT, NAMED>();
        if (lane == 0) s_red[warp] = stored_ss;
        pool_sync<NT, NAMED>();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < NT / 32; ++w) t += s_red[w];
            if (p.store_norm2) p.store_norm2[b] = t;
            if (p.alive) {
                const long long row = p.first_row + b;
                atomicOr(&p.alive[row >> 5], 1u << (row & 31));
            }
        }
    }
    pool_sync<NT, NAMED>();
}

// HT = float (VEC 4) or __nv_bfloat16 (VEC 8); H % VEC == 0 is required by the launcher.
template <typename HT, typename MT, int kPoolThreads>
__global__ void __launch_bounds__(kPoolThreads) pool_normalize_kernel(const PoolParams p)
{
    constexpr int VEC = sizeof(HT) == 4 ? 4 : 8;
    extern __shared__ __align__(16) float spart[];  // [LS][H] partial sums, then pooled row in [0,H)
    __shared__ float s_red[kPoolThreads / 32];
    __shared__ float s_scalar[2];

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nvec = p.H / VEC;
    const int VT = nvec < kPoolThreads ? nvec : kPoolThreads;  // threads across one row
    const int LS = kPoolThreads / VT;                           // token slices
    const int ls = tid / VT, vc0 = tid - ls * VT;
    const unsigned char *hid = reinterpret_cast<const unsigned char *>(p.hidden) +
                               (size_t)b * p.L * p.H * sizeof(HT);
    const size_t mbase = (size_t)b * p.L;

    // ---- the mask row: weights to shared memory, their sum and the end of the live range (one barrier).
    //      Tokens whose weight is zero are never read: padding behind the last live token is not even visited,
    //      holes inside the range are skipped by a (warp-coherent) branch ----
    float *s_w = spart + (size_t)LS * p.H;               // [L]
    __shared__ int s_end[kPoolThreads / 32];
    float msum = 0.f;
    int last = 0;
    for (int t = tid; t < p.L; t += kPoolThreads) {
        const float w = mask_at<MT>(p.mask, mbase + t);
        s_w[t] = w;
        msum += w;
        if (w != 0.f) last = t + 1;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        msum += __shfl_xor_sync(0xffffffffu, msum, d);
        last = max(last, __shfl_xor_sync(0xffffffffu, last, d));
    }
    if (lane == 0) {
        s_red[warp] = msum;
        s_end[warp] = last;
    }
    __syncthreads();
    int n_tok = 0;
    {
        float m = 0.f;
#pragma unroll
        for (int w = 0; w < kPoolThreads / 32; ++w) {
            m += s_red[w];
            n_tok = max(n_tok, s_end[w]);
        }
        if (tid == 0) s_scalar[0] = fmaxf(m, 1e-9f);
    }

    // masked sums: thread (ls, vc) accumulates tokens ls, ls+LS, ... of 16-byte column group vc, eight
    // independent 16-byte streaming loads in flight
    if (ls < LS) {
        for (int vc = vc0; vc < nvec; vc += VT) {
            float acc[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
            const unsigned char *col = hid + (size_t)vc * VEC * sizeof(HT);
            for (int t0 = ls; t0 < n_tok; t0 += 8 * LS) {
                uint4 d[8];
                float m[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int t = t0 + u * LS;
                    m[u] = t < n_tok ? s_w[t] : 0.f;
                    if (m[u] != 0.f) {
                        const unsigned char *ptr = col + (size_t)t * p.H * sizeof(HT);
                        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(d[u].x), "=r"(d[u].y), "=r"(d[u].z), "=r"(d[u].w)
                                     : "l"(ptr));
                    } else {
                        d[u] = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if constexpr (VEC == 4) {
                        acc[0] = fmaf(__uint_as_float(d[u].x), m[u], acc[0]);
                        acc[1] = fmaf(__uint_as_float(d[u].y), m[u], acc[1]);
                        acc[2] = fmaf(__uint_as_float(d[u].z), m[u], acc[2]);
                        acc[3] = fmaf(__uint_as_float(d[u].w), m[u], acc[3]);
                    } else {
                        acc[0] = fmaf(__uint_as_float(d[u].x << 16), m[u], acc[0]);
                        acc[1] = fmaf(__uint_as_float(d[u].x & 0xffff0000u), m[u], acc[1]);
                        acc[2] = fmaf(__uint_as_float(d[u].y << 16), m[u], acc[2]);
                        acc[3] = fmaf(__uint_as_float(d[u].y & 0xffff0000u), m[u], acc[3]);
                        acc[4] = fmaf(__uint_as_float(d[u].z << 16), m[u], acc[4]);
                        acc[5] = fmaf(__uint_as_float(d[u].z & 0xffff0000u), m[u], acc[5]);
                        acc[6] = fmaf(__uint_as_float(d[u].w << 16), m[u], acc[6]);
                        acc[7] = fmaf(__uint_as_float(d[u].w & 0xffff0000u), m[u], acc[7]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) spart[(size_t)ls * p.H + vc * VEC + i] = acc[i];
        }
    }
    __syncthreads();
    pool_finish<kPoolThreads, false>(p, b, tid, LS, spart, s_red, s_scalar, s_scalar[0]);
}

// ---------------------------------------------------------------------------------------------------------------
// Ring kernel.  Thread layout: warps 0..7 consume, warp 8 produces.  Shared memory:
//   ring  [4][chunk_bytes]   chunks of TOK consecutive token rows, filled by cp.async.bulk (16-byte aligned sizes)
//   spart [LS][H] f32        the token slices' sums at the end of a sequence
//   s_w   [2][L]  f32        mask weights of the current and the next sequence
// Barriers: full[s] / empty[s] per ring stage (tx bytes / one arrival per consumer warp); wfull[b] / wempty[b] per
// mask buffer.  The producer may run a whole sequence ahead of the consumers.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRingConsumers = 256;
constexpr int kRingThreads = kRingConsumers + 32;
constexpr int kRingStages = 4;        // power of two: stage and phase of a chunk are a mask and a shift

struct RingShape {
    int B;              // sequences
    int tok;            // token rows per chunk
    int chunk_bytes;    // tok * row bytes
    int *ctr;           // [2] device counters of the launching stream: sequences claimed beyond the first one of
                        // every CTA, CTAs finished (the last one zeroes both for the next launch)
};

template <typename HT, typename MT>
__global__ void __launch_bounds__(kRingThreads) pool_ring_kernel(const PoolParams p, const RingShape g)
{
    constexpr int VEC = sizeof(HT) == 4 ? 4 : 8;
    extern __shared__ __align__(128) unsigned char rsm[];
    __shared__ __align__(8) unsigned long long s_bar[2 * kRingStages + 4];
    __shared__ float s_red[kRingConsumers / 32];
    __shared__ float s_scalar[2];
    __shared__ int s_ntok[2];
    __shared__ int s_seq[2];
    __shared__ float s_msum[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nvec = p.H / VEC;                                  // <= kRingConsumers (launcher)
    const int LS = kRingConsumers / nvec;                        // token slices
    const size_t row_bytes = (size_t)p.H * sizeof(HT);
    unsigned char *ring = rsm;
    float *spart = reinterpret_cast<float *>(rsm + (size_t)kRingStages * g.chunk_bytes);
    float *s_w = spart + (size_t)LS * p.H;
    const uint32_t bar0 = ptx::smem_u32(s_bar);
    auto full = [&](int st) { return bar0 + 8u * st; };
    auto empty = [&](int st) { return bar0 + 8u * (kRingStages + st); };
    auto wfull = [&](int b) { return bar0 + 8u * (2 * kRingStages + b); };
    auto wempty = [&](int b) { return bar0 + 8u * (2 * kRingStages + 2 + b); };

    if (tid == 0) {
        for (int st = 0; st < kRingStages; ++st) {
            ptx::mbar_init(full(st), 1);
            ptx::mbar_init(empty(st), kRingConsumers / 32);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(wfull(b), 1);
            ptx::mbar_init(wempty(b), 1);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    if (warp == kRingConsumers / 32) {
        // ---------------- producer ----------------
        // Sequences are claimed dynamically (the first one is the CTA's own index): lengths differ, and a CTA that
        // finishes early takes work the slower ones would otherwise still hold when the machine drains.
        unsigned it = 0;
        int b = blockIdx.x;
        for (int i = 0;; ++i) {
            const int buf = i & 1;
            ptx::mbar_wait(wempty(buf), ((i >> 1) & 1) ^ 1);
            if (b >= g.B) {                                       // nothing left: tell the consumers and leave
                if (lane == 0) {
                    s_seq[buf] = -1;
                    ptx::mbar_arrive(wfull(buf));
                    if (atomicAdd(g.ctr + 1, 1) == (int)gridDim.x - 1) {     // every CTA is past its last claim
                        g.ctr[0] = 0;
                        g.ctr[1] = 0;
                    }
                }
                break;
            }
            int b_next = 0;                                       // claimed now, needed after this sequence is issued
            if (lane == 0) b_next = (int)gridDim.x + atomicAdd(g.ctr, 1);
            float msum = 0.f;
            int last = 0;
            const size_t mbase = (size_t)b * p.L;
            for (int t = lane; t < p.L; t += 32) {
                const float w = mask_at<MT>(p.mask, mbase + t);
                s_w[(size_t)buf * p.L + t] = w;
                msum += w;
                if (w != 0.f) last = t + 1;
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                msum += __shfl_xor_sync(0xffffffffu, msum, d);
                last = max(last, __shfl_xor_sync(0xffffffffu, last, d));
            }
            if (lane == 0) {
                s_ntok[buf] = last;
                s_seq[buf] = b;
                s_msum[buf] = fmaxf(msum, 1e-9f);
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(wfull(buf));
            const int n_fetch = last > 0 ? last : 1;              // a sequence always owns at least one chunk
            const int nch = (n_fetch + g.tok - 1) / g.tok;
            const unsigned char *src = reinterpret_cast<const unsigned char *>(p.hidden) + (size_t)b * p.L * row_bytes;
            for (int c = 0; c < nch; ++c, ++it) {
                const int st = it & (kRingStages - 1);
                const unsigned ph = (it / kRingStages) & 1u;
                ptx::mbar_wait(empty(st), ph ^ 1u);
                if (lane == 0) {
                    const int ntk = n_fetch - c * g.tok < g.tok ? n_fetch - c * g.tok : g.tok;
                    const uint32_t bytes = (uint32_t)(ntk * row_bytes);
                    ptx::mbar_arrive_expect_tx(full(st), bytes);
                    ptx::bulk_load(ptx::smem_u32(ring + (size_t)st * g.chunk_bytes), src + (size_t)c * g.tok * row_bytes, bytes,
                                   full(st));
                }
            }
            b = __shfl_sync(0xffffffffu, b_next, 0);
        }
        return;
    }

    // ---------------- consumers ----------------
    // Thread (ls, vc): 16-byte column group vc of the tokens ls, ls + LS, ... of every chunk.  Shared memory is
    // addressed in its own state space (32-bit addresses, no generic-pointer conversion in the loop), two tokens
    // per trip with all four loads issued before the first use.
    const int ls = tid / nvec, vc = tid - ls * nvec;
    const bool active = ls < LS;
    const uint32_t ring_s = ptx::smem_u32(ring), w_s = ptx::smem_u32(s_w);
    const uint32_t tok_step = (uint32_t)(LS * row_bytes);
    const uint32_t col_off = (uint32_t)(vc * 16 + ls * row_bytes);
    auto lds128 = [](uint32_t addr) {
        uint4 d;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "r"(addr));
        return d;
    };
    auto lds32 = [](uint32_t addr) {
        float f;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(addr));
        return f;
    };
    unsigned it = 0;
    for (int i = 0;; ++i) {
        const int buf = i & 1;
        ptx::mbar_wait(wfull(buf), (i >> 1) & 1);
        const int b = s_seq[buf];
        if (b < 0) break;
        const int n_tok = s_ntok[buf];
        const float msum = s_msum[buf];
        const uint32_t w_seq = w_s + (uint32_t)(buf * p.L + ls) * 4u;
        const int nch = ((n_tok > 0 ? n_tok : 1) + g.tok - 1) / g.tok;
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
        auto add = [&](const uint4 &d, float m) {
            if (m != 0.f) {                                  // masked tokens are never accumulated (0 * inf = nan)
                if constexpr (VEC == 4) {
                    acc[0] = fmaf(__uint_as_float(d.x), m, acc[0]);
                    acc[1] = fmaf(__uint_as_float(d.y), m, acc[1]);
                    acc[2] = fmaf(__uint_as_float(d.z), m, acc[2]);
                    acc[3] = fmaf(__uint_as_float(d.w), m, acc[3]);
                } else {
                    acc[0] = fmaf(__uint_as_float(d.x << 16), m, acc[0]);
                    acc[1] = fmaf(__uint_as_float(d.x & 0xffff0000u), m, acc[1]);
                    acc[2] = fmaf(__uint_as_float(d.y << 16), m, acc[2]);
                    acc[3] = fmaf(__uint_as_float(d.y & 0xffff0000u), m, acc[3]);
                    acc[4] = fmaf(__uint_as_float(d.z << 16), m, acc[4]);
                    acc[5] = fmaf(__uint_as_float(d.z & 0xffff0000u), m, acc[5]);
                    acc[6] = fmaf(__uint_as_float(d.w << 16), m, acc[6]);
                    acc[7] = fmaf(__uint_as_float(d.w & 0xffff0000u), m, acc[7]);
                }
            }
        };
        for (int c = 0; c < nch; ++c, ++it) {
            const int st = it & (kRingStages - 1);
            const unsigned ph = (it / kRingStages) & 1u;
            ptx::mbar_wait(full(st), ph);
            if (active) {
                const int t0 = c * g.tok;
                const int cnt = n_tok - t0 < g.tok ? n_tok - t0 : g.tok;
                uint32_t a = ring_s + (uint32_t)st * (uint32_t)g.chunk_bytes + col_off;
                uint32_t wa = w_seq + (uint32_t)t0 * 4u;
                int tl = ls;
                for (; tl + LS < cnt; tl += 2 * LS) {
                    const float m0 = lds32(wa), m1 = lds32(wa + 4u * LS);
                    const uint4 d0 = lds128(a), d1 = lds128(a + tok_step);
                    a += 2u * tok_step;
                    wa += 8u * LS;
                    add(d0, m0);
                    add(d1, m1);
                }
                if (tl < cnt) add(lds128(a), lds32(wa));
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(empty(st));
        }
        if (active) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) spart[(size_t)ls * p.H + vc * VEC + e] = acc[e];
        }
        ptx::named_bar_sync(1, kRingConsumers);           // sums visible; nobody reads this sequence's weights any more
        if (tid == 0) ptx::mbar_arrive(wempty(buf));
        pool_finish<kRingConsumers, true>(p, b, tid, LS, spart, s_red, s_scalar, msum);
    }
}

// Two device counters per (device, launching stream) for the ring kernel's dynamic sequence claims.  Launches on
// one stream are serial and every launch leaves its pair zeroed, so a pair is never shared by two running kernels.
static int ring_counters(cudaStream_t st, int **out)
{
    constexpr int kSlots = 256, kMaxDev = 64;
    static std::mutex mu;
    static int *base[kMaxDev] = {};
    static std::unordered_map<cudaStream_t, int> slot_of[kMaxDev];
    int dev = 0;
    ARCHI_CUDA(cudaGetDevice(&dev));
    ARCHI_REQUIRE(dev >= 0 && dev < kMaxDev, "pool_normalize: device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(mu);
    if (!base[dev]) {
        // the one allocation of this path: not while a stream capture is under way (the caller then takes the other kernel)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
            cudaGetLastError();
            *out = nullptr;
            return ARCHI_OK;
        }
        ARCHI_CUDA(cudaMalloc(&base[dev], kSlots * 2 * sizeof(int)));
        ARCHI_CUDA(cudaMemset(base[dev], 0, kSlots * 2 * sizeof(int)));
    }
    auto itr = slot_of[dev].find(st);
    int slot;
    if (itr == slot_of[dev].end()) {
        slot = (int)(slot_of[dev].size() % kSlots);
        slot_of[dev].emplace(st, slot);
    } else {
        slot = itr->second;
    }
    *out = base[dev] + 2 * slot;
    return ARCHI_OK;
}

// shape of the ring for (row bytes, L); false when the ring kernel does not apply
static bool ring_shape(int B, int L, int H, int elt, bool forced, RingShape *g, size_t *smem)
{
    const int VEC = 16 / elt;
    const int nvec = H / VEC;
    if (nvec > kRingConsumers) return false;
    const size_t row_bytes = (size_t)H * elt;
    const int LS = kRingConsumers / nvec;
    // ~11 KB chunks: four stages + sums fit four CTAs per SM (ARCHI_POOL_CHUNK_KB: experiments)
    const char *ck = getenv("ARCHI_POOL_CHUNK_KB");
    const size_t chunk_target = ck && atoi(ck) > 0 ? (size_t)atoi(ck) * 1024 : 11264;
    int tok = (int)(chunk_target / row_bytes);
    if (tok < LS) tok = LS;
    if (tok < 1) tok = 1;
    if (tok > L) tok = L;
    g->B = B;
    g->tok = tok;
    g->chunk_bytes = (int)(tok * row_bytes);
    *smem = (size_t)kRingStages * g->chunk_bytes + ((size_t)LS * H + 2 * (size_t)L) * sizeof(float);
    if (*smem > 200 * 1024) return false;
    int dev = 0, sm_count = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return false;
    if (B < 2 * sm_count && !forced) return false;          // small batches: one CTA per sequence
    return true;
}

int launch_pool_normalize(const void *hidden, int hidden_dtype, const void *mask, int mask_dtype,
                          int B, int L, int H, void *store_rows, int store_dtype, int store_ld,
                          float *store_norm2, uint32_t *alive, long long first_row, void *out_bf16,
                          float *out_f32, cudaStream_t st)
{
    ARCHI_REQUIRE(hidden_dtype == ARCHI_F32 || hidden_dtype == ARCHI_BF16,
                  "pool_normalize: hidden dtype must be f32 or bf16");
    ARCHI_REQUIRE(mask_dtype == ARCHI_I32 || mask_dtype == ARCHI_I64,
                  "pool_normalize: mask dtype must be i32 or i64");
    ARCHI_REQUIRE(B >= 0 && L >= 1 && H >= 1, "pool_normalize: bad shape B=%d L=%d H=%d", B, L, H);
    const int VEC = hidden_dtype == ARCHI_F32 ? 4 : 8;
    ARCHI_REQUIRE(H % VEC == 0, "pool_normalize: H=%d must be a multiple of %d", H, VEC);
    if (B == 0) return ARCHI_OK;
    const int nvec = H / VEC;

    PoolParams p;
    p.hidden = hidden;
    p.mask = mask;
    p.L = L;
    p.H = H;
    p.store_rows = store_rows;
    p.store_is_bf16 = store_dtype == ARCHI_BF16;
    p.store_ld = store_ld;
    p.store_norm2 = store_norm2;
    p.alive = alive;
    p.first_row = first_row;
    p.out_bf16 = reinterpret_cast<__nv_bfloat16 *>(out_bf16);
    p.out_f32 = out_f32;

    // ARCHI_POOL_RING: 0 = never use the ring kernel, 1 = whenever the shape allows (also small batches), unset = auto
    const char *env = getenv("ARCHI_POOL_RING");
    const int mode = env ? atoi(env) : -1;
    RingShape g;
    g.ctr = nullptr;
    size_t ring_smem = 0;
    if (mode != 0 && (reinterpret_cast<uintptr_t>(hidden) & 15) == 0 &&
        ring_shape(B, L, H, hidden_dtype == ARCHI_F32 ? 4 : 2, mode == 1, &g, &ring_smem)) {
        int rc = ring_counters(st, &g.ctr);
        if (rc != ARCHI_OK) return rc;
    }
    if (g.ctr != nullptr) {
        void (*rfn)(const PoolParams, const RingShape);
        if (hidden_dtype == ARCHI_F32)
            rfn = mask_dtype == ARCHI_I64 ? pool_ring_kernel<float, long long> : pool_ring_kernel<float, int>;
        else
            rfn = mask_dtype == ARCHI_I64 ? pool_ring_kernel<__nv_bfloat16, long long> : pool_ring_kernel<__nv_bfloat16, int>;
        if (ring_smem > 48 * 1024)
            ARCHI_CUDA(cudaFuncSetAttribute((const void *)rfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
        // resident CTAs only (every slot of the machine); they claim sequences until none is left
        int per_sm = 0, dev = 0, sm_count = 0;
        ARCHI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)rfn, kRingThreads, ring_smem));
        ARCHI_CUDA(cudaGetDevice(&dev));
        ARCHI_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        const char *cps_env = getenv("ARCHI_POOL_CTAS_PER_SM");        // experiments: fewer, fatter CTAs
        if (cps_env && atoi(cps_env) > 0 && atoi(cps_env) < per_sm) per_sm = atoi(cps_env);
        const long long max_grid = (long long)sm_count * (per_sm > 0 ? per_sm : 1);
        const int ring_grid = (int)(B < max_grid ? B : max_grid);
        rfn<<<ring_grid, kRingThreads, ring_smem, st>>>(p, g);
        ARCHI_CHECK_LAUNCH();
        return ARCHI_OK;
    }

    const int threads = B >= 512 ? 128 : 256;
    const int VT = nvec < threads ? nvec : threads;
    const int LS = threads / VT;
    const size_t smem = ((size_t)LS * H + (size_t)L) * sizeof(float);   // partial sums + mask weights
    ARCHI_REQUIRE(smem <= 200 * 1024, "pool_normalize: H=%d, L=%d need %zu B of shared memory", H, L, smem);

    void (*fn)(const PoolParams);
    if (threads == 128) {
        if (hidden_dtype == ARCHI_F32)
            fn = mask_dtype == ARCHI_I64 ? pool_normalize_kernel<float, long long, 128> : pool_normalize_kernel<float, int, 128>;
        else
            fn = mask_dtype == ARCHI_I64 ? pool_normalize_kernel<__nv_bfloat16, long long, 128>
                                         : pool_normalize_kernel<__nv_bfloat16, int, 128>;
    } else {
        if (hidden_dtype == ARCHI_F32)
            fn = mask_dtype == ARCHI_I64 ? pool_normalize_kernel<float, long long, 256> : pool_normalize_kernel<float, int, 256>;
        else
            fn = mask_dtype == ARCHI_I64 ? pool_normalize_kernel<__nv_bfloat16, long long, 256>
                                         : pool_normalize_kernel<__nv_bfloat16, int, 256>;
    }
    if (smem > 48 * 1024)
        ARCHI_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    fn<<<B, threads, smem, st>>>(p);
    ARCHI_CHECK_LAUNCH();
    return ARCHI_OK;
}

}  // namespace archi
