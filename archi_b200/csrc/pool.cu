// pool.cu -- fused masked mean-pool + L2-normalise + cast on the encoder's last hidden state.
//
// Replaces the tail of Embeddings.embed_documents / embed_query (reference call sites
// manager.py:373, postgres_vectorstore.py:143,245,390), i.e. sentence-transformers'
// Pooling(mean) + Normalize modules [external]:
//     pooled = sum_t h[t]*m[t] / max(sum_t m[t], 1e-9);   out = pooled / max(|pooled|_2, 1e-12)
// One CTA per sequence.  HBM-bound: the [L, H] slab of a sequence is read once with 16-byte streaming loads
// (tokens whose mask is 0 are not read at all: padding behind the last live token is not visited), the pooled row never leaves the SM, and the
// normalised row is written once in each requested format -- optionally straight into the tail
// of the corpus matrix together with its |row|^2 and live bit, so add_documents needs no further
// kernel.
#include "common.cuh"

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

    // reduce the token slices, divide by the mask sum, accumulate |pooled|^2
    const float inv_m = 1.0f / s_scalar[0];
    float ss = 0.f;
    for (int h = tid; h < p.H; h += kPoolThreads) {
        float v = 0.f;
        for (int l = 0; l < LS; ++l) v += spart[(size_t)l * p.H + h];
        v *= inv_m;
        ss = fmaf(v, v, ss);
        spart[h] = v;  // slice 0 is only read by this same thread for column h
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
    __syncthreads();
    if (lane == 0) s_red[warp] = ss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < kPoolThreads / 32; ++w) t += s_red[w];
        s_scalar[1] = 1.0f / fmaxf(sqrtf(t), 1e-12f);
    }
    __syncthreads();
    const float inv_n = s_scalar[1];

    // write the normalised row to every sink; |stored row|^2 uses the values as stored
    float stored_ss = 0.f;
    for (int h = tid; h < p.H; h += kPoolThreads) {
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
        for (int h = p.H + tid; h < p.store_ld; h += kPoolThreads) {
            if (p.store_is_bf16)
                reinterpret_cast<__nv_bfloat16 *>(p.store_rows)[(size_t)b * p.store_ld + h] =
                    __float2bfloat16_rn(0.f);
            else
                reinterpret_cast<float *>(p.store_rows)[(size_t)b * p.store_ld + h] = 0.f;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) stored_ss += __shfl_xor_sync(0xffffffffu, stored_ss, d);
        __syncthreads();
        if (lane == 0) s_red[warp] = stored_ss;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < kPoolThreads / 32; ++w) t += s_red[w];
            if (p.store_norm2) p.store_norm2[b] = t;
            if (p.alive) {
                const long long row = p.first_row + b;
                atomicOr(&p.alive[row >> 5], 1u << (row & 31));
            }
        }
    }
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
    const int threads = B >= 512 ? 128 : 256;
    const int VT = nvec < threads ? nvec : threads;
    const int LS = threads / VT;
    const size_t smem = ((size_t)LS * H + (size_t)L) * sizeof(float);   // partial sums + mask weights
    ARCHI_REQUIRE(smem <= 200 * 1024, "pool_normalize: H=%d, L=%d need %zu B of shared memory", H, L, smem);

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
