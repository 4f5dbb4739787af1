// exchange.cu -- shard exchange + merge in ONE kernel over NVLink peer memory.
//
// The reference has no distributed path; this replaces "NCCL all-gather of the per-shard k-lists, then
// merge" (sharded.py) when the ranks of a box can map each other's memory (CUDA IPC).  Every rank owns
// a gather buffer [2 parities][world][record] plus one arrival flag per peer.  One kernel per search:
//   1. push: the rank's record {ids [nq,k] int64 | scores [nq,k] fp32} is stored straight into slot
//      [parity][rank] of EVERY peer's buffer (16-byte stores over NVLink / NVSwitch);
//   2. signal: when the last CTA has pushed, it release-stores the call's epoch into flag [rank] of
//      every peer;
//   3. wait: every CTA acquires flags [0, world) >= epoch (bounded spin: a missing peer ends in a status
//      word, never in a hang);
//   4. merge: one warp per query merges the world lists from the local buffer (merge.cuh).
// Buffers are double-buffered by epoch parity: a peer can be at most one call ahead of the slowest rank
// (its next call needs this rank's flag of the current one), so slot [parity] is never overwritten
// while it is still being merged.  All exchange kernels of a rank must be issued in order (one stream).
#include <string.h>

#include "common.cuh"
#include "merge.cuh"

namespace archi {

constexpr int kExchMaxWorld = 32;
constexpr int kExchThreads = 256;

struct ExchangeParams {
    unsigned char *peer_base[kExchMaxWorld];   // every rank's gather buffer as mapped into this process
    const unsigned char *record;               // this rank's record (device)
    size_t rec_bytes;                          // bytes actually used by a record (multiple of 16)
    size_t slot_bytes;                         // distance between slots (max record bytes, multiple of 16)
    size_t flags_off;                          // byte offset of flags [world] inside a gather buffer
    int rank, world, parity;
    uint32_t epoch;
    int nq, k, larger;
    float *out_scores;
    long long *out_ids;
    unsigned int *counter;                     // local: CTAs that finished pushing
    int *status;                               // local: 1 = a wait timed out
    unsigned long long timeout_ns;
    int fence_all;                             // experiments (ARCHI_EXCH_FENCE_ALL=1): every thread fences after its pushes
    int host_status;                           // experiments (ARCHI_EXCH_HOST_STATUS=1): read the sticky status from host memory
    int *dev_status;                           // device copy of the sticky status word
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(kExchThreads) exchange_merge_kernel(const ExchangeParams p)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int s_last;
    __shared__ int s_bad;      // a wait gave up in this call or in an earlier one: the gather buffer cannot be trusted
    // the sticky status lives in pinned host memory (the host checks it before every launch); the kernel reads its
    // device copy -- a read of host memory from every CTA of every exchange is a PCIe round trip on the critical path
    if (tid == 0) s_bad = p.host_status ? *reinterpret_cast<volatile int *>(p.status) : *reinterpret_cast<volatile int *>(p.dev_status);

    // ---- 1. push my record into slot [parity][rank] of every peer (myself included) ----
    const size_t vecs = p.rec_bytes / 16;
    const size_t slot = ((size_t)p.parity * p.world + p.rank) * p.slot_bytes;
    const uint4 *src = reinterpret_cast<const uint4 *>(p.record);
    for (size_t i = (size_t)blockIdx.x * kExchThreads + tid; i < vecs * p.world; i += (size_t)gridDim.x * kExchThreads) {
        const int peer = (int)(i / vecs);
        const size_t v = i - (size_t)peer * vecs;
        reinterpret_cast<uint4 *>(p.peer_base[peer] + slot)[v] = src[v];
    }
    // one system-scope fence per CTA, after the barrier that orders every thread's stores before it (fences are
    // cumulative); a fence by each of the 32k threads serialises in the memory system
    if (p.fence_all) __threadfence_system();
    __syncthreads();

    // ---- 2. the last CTA to finish signals every peer ----
    if (tid == 0) {
        __threadfence_system();
        s_last = atomicAdd(p.counter, 1u) == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (tid < p.world)
            st_release_sys(reinterpret_cast<uint32_t *>(p.peer_base[tid] + p.flags_off) + p.rank, p.epoch);
        if (tid == 0) *p.counter = 0u;         // the next call on this stream starts from zero
    }

    // ---- 3. wait for every rank's record of this call ----
    if (tid < p.world) {
        const uint32_t *flag = reinterpret_cast<const uint32_t *>(p.peer_base[p.rank] + p.flags_off) + tid;
        const unsigned long long t0 = global_timer_ns();
        while ((int)(ld_acquire_sys(flag) - p.epoch) < 0) {
            if (global_timer_ns() - t0 > p.timeout_ns) {
                *reinterpret_cast<volatile int *>(p.status) = 1;      // host-mapped and sticky: the next call fails on the host
                *reinterpret_cast<volatile int *>(p.dev_status) = 1;
                __threadfence_system();
                s_bad = 1;
                break;
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (s_bad) {
        // never merge a stale or half-written slot into plausible results: poison this call's outputs
        const size_t n_out = (size_t)p.nq * p.k;
        for (size_t i = (size_t)blockIdx.x * kExchThreads + tid; i < n_out; i += (size_t)gridDim.x * kExchThreads) {
            p.out_scores[i] = __int_as_float(0x7fc00000);
            p.out_ids[i] = -1;
        }
        return;
    }

    // ---- 4. merge the world lists of each query from the local gather buffer ----
    const unsigned char *mine = p.peer_base[p.rank] + (size_t)p.parity * p.world * p.slot_bytes;
    const size_t n = (size_t)p.nq * p.k;
    const long long *ids = reinterpret_cast<const long long *>(mine);
    const float *scores = reinterpret_cast<const float *>(mine + n * 8);
    for (int q = blockIdx.x * (kExchThreads / 32) + warp; q < p.nq; q += gridDim.x * (kExchThreads / 32))
        merge_query_lists(scores, ids, p.slot_bytes / 4, p.slot_bytes / 8, p.world, p.nq, q, p.k, p.larger,
                          p.out_scores, p.out_ids, lane);
}

}  // namespace archi

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
struct archi_exchange {
    int device = 0, rank = 0, world = 1;
    size_t slot_bytes = 0, flags_off = 0, total_bytes = 0;
    unsigned char *local = nullptr;               // cudaMalloc'ed: [2][world][slot] | flags [world] | counter | status
    unsigned char *peer[archi::kExchMaxWorld] = {};
    int *h_status = nullptr;                      // pinned + mapped: 1 once any wait timed out (sticky)
    int *d_status = nullptr;                      // the device alias the kernel writes
    bool connected = false;
    uint32_t epoch = 0;
    std::mutex mu;
};

namespace {
struct ScopedDevice {
    int prev = -1;
    bool ok = true;
    explicit ScopedDevice(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
    }
    ~ScopedDevice()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
}  // namespace

#define EXCH_DEVICE(dev)                                             \
    ScopedDevice _guard(dev);                                        \
    if (!_guard.ok) {                                                \
        archi::set_error("cudaSetDevice(%d) failed", (int)(dev));    \
        return ARCHI_ECUDA;                                          \
    }

extern "C" {

int archi_exchange_create(int device, int rank, int world, int64_t max_record_bytes, archi_exchange_t **out)
{
    ARCHI_REQUIRE(out != nullptr, "exchange_create: null out");
    ARCHI_REQUIRE(world >= 1 && world <= archi::kExchMaxWorld && rank >= 0 && rank < world,
                  "exchange_create: rank %d / world %d out of range (world <= %d)", rank, world, archi::kExchMaxWorld);
    ARCHI_REQUIRE(max_record_bytes > 0, "exchange_create: max_record_bytes must be positive");
    EXCH_DEVICE(device);
    archi_exchange *x = new archi_exchange();
    x->device = device;
    x->rank = rank;
    x->world = world;
    x->slot_bytes = ((size_t)max_record_bytes + 15) / 16 * 16;
    x->flags_off = 2 * (size_t)world * x->slot_bytes;
    x->total_bytes = x->flags_off + (size_t)world * 4 + 64;   // + counter, status (local use only)
    cudaError_t e = cudaMalloc(&x->local, x->total_bytes);
    if (e == cudaSuccess) e = cudaMemset(x->local, 0, x->total_bytes);
    if (e == cudaSuccess) e = cudaHostAlloc(&x->h_status, sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *x->h_status = 0;
        e = cudaHostGetDevicePointer(&x->d_status, x->h_status, 0);
    }
    if (e != cudaSuccess) {
        archi::set_error("exchange_create: allocating %zu bytes failed: %s", x->total_bytes, cudaGetErrorString(e));
        if (x->local) cudaFree(x->local);
        if (x->h_status) cudaFreeHost(x->h_status);
        delete x;
        return ARCHI_ECUDA;
    }
    x->peer[rank] = x->local;
    *out = x;
    return ARCHI_OK;
}

int archi_exchange_local_handle(archi_exchange_t *x, void *handle_out)
{
    ARCHI_REQUIRE(x && handle_out, "exchange_local_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == ARCHI_EXCHANGE_HANDLE_BYTES, "IPC handle size");
    EXCH_DEVICE(x->device);
    cudaIpcMemHandle_t h;
    ARCHI_CUDA(cudaIpcGetMemHandle(&h, x->local));
    memcpy(handle_out, &h, sizeof(h));
    return ARCHI_OK;
}

int archi_exchange_connect(archi_exchange_t *x, const void *handles)
{
    ARCHI_REQUIRE(x && handles, "exchange_connect: null argument");
    std::lock_guard<std::mutex> lock(x->mu);
    ARCHI_REQUIRE(!x->connected, "exchange_connect: already connected");
    EXCH_DEVICE(x->device);
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char *)handles + (size_t)r * sizeof(h), sizeof(h));
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            archi::set_error("exchange_connect: cannot map rank %d's buffer: %s", r, cudaGetErrorString(e));
            cudaGetLastError();
            for (int j = 0; j < r; ++j)
                if (j != x->rank && x->peer[j]) {
                    cudaIpcCloseMemHandle(x->peer[j]);
                    x->peer[j] = nullptr;
                }
            return ARCHI_EUNSUPPORTED;
        }
        x->peer[r] = (unsigned char *)ptr;
    }
    x->connected = true;
    return ARCHI_OK;
}

int archi_exchange_merge_topk(archi_exchange_t *x, const void *record_dev, int nq, int k, int larger_is_better,
                              float *out_scores_dev, int64_t *out_ids_dev, void *stream)
{
    ARCHI_REQUIRE(x && record_dev && out_scores_dev && out_ids_dev, "exchange_merge_topk: null argument");
    ARCHI_REQUIRE(nq >= 1 && k >= 1, "exchange_merge_topk: nq=%d k=%d must be positive", nq, k);
    std::lock_guard<std::mutex> lock(x->mu);
    ARCHI_REQUIRE(x->connected || x->world == 1, "exchange_merge_topk: peers are not connected");
    const size_t rec = ((size_t)nq * k * 12 + 15) / 16 * 16;
    ARCHI_REQUIRE(rec <= x->slot_bytes, "exchange_merge_topk: record of %zu bytes exceeds the slot (%zu bytes)", rec,
                  x->slot_bytes);
    ARCHI_REQUIRE(((uintptr_t)record_dev & 15) == 0, "exchange_merge_topk: record must be 16-byte aligned");
    if (*reinterpret_cast<volatile int *>(x->h_status) != 0) {
        archi::set_error("exchange_merge_topk: an earlier exchange timed out waiting for a rank; the results of that call "
                         "were invalidated (id -1 / NaN) and the ranks' epochs may be skewed -- rebuild the exchange");
        return ARCHI_ECUDA;
    }
    EXCH_DEVICE(x->device);
    archi::ExchangeParams p;
    for (int r = 0; r < archi::kExchMaxWorld; ++r) p.peer_base[r] = r < x->world ? x->peer[r] : nullptr;
    p.record = (const unsigned char *)record_dev;
    p.rec_bytes = rec;
    p.slot_bytes = x->slot_bytes;
    p.flags_off = x->flags_off;
    p.rank = x->rank;
    p.world = x->world;
    x->epoch += 1;
    p.epoch = x->epoch;
    p.parity = (int)(x->epoch & 1u);
    p.nq = nq;
    p.k = k;
    p.larger = larger_is_better;
    p.out_scores = out_scores_dev;
    p.out_ids = (long long *)out_ids_dev;
    p.counter = reinterpret_cast<unsigned int *>(x->local + x->flags_off + (size_t)x->world * 4);
    p.status = x->d_status;
    p.dev_status = reinterpret_cast<int *>(x->local + x->flags_off + (size_t)x->world * 4 + 16);
    static const int fence_all = getenv("ARCHI_EXCH_FENCE_ALL") ? atoi(getenv("ARCHI_EXCH_FENCE_ALL")) : 0;
    static const int host_status = getenv("ARCHI_EXCH_HOST_STATUS") ? atoi(getenv("ARCHI_EXCH_HOST_STATUS")) : 0;
    p.fence_all = fence_all;
    p.host_status = host_status;
    p.timeout_ns = 5ull * 1000 * 1000 * 1000;
    int grid = (nq + archi::kExchThreads / 32 - 1) / (archi::kExchThreads / 32);
    if (grid < 8) grid = 8;
    if (grid > 128) grid = 128;                 // all CTAs co-resident: the waits cannot starve the pushes
    archi::exchange_merge_kernel<<<grid, archi::kExchThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    ARCHI_CHECK_LAUNCH();
    return ARCHI_OK;
}

int archi_exchange_status(archi_exchange_t *x, int *timed_out)
{
    ARCHI_REQUIRE(x && timed_out, "exchange_status: null argument");
    EXCH_DEVICE(x->device);
    ARCHI_CUDA(cudaDeviceSynchronize());
    *timed_out = *reinterpret_cast<volatile int *>(x->h_status);
    return ARCHI_OK;
}

int archi_exchange_destroy(archi_exchange_t *x)
{
    if (!x) return ARCHI_OK;
    {
        ScopedDevice guard(x->device);
        cudaDeviceSynchronize();
        for (int r = 0; r < x->world; ++r)
            if (r != x->rank && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
        if (x->local) cudaFree(x->local);
        if (x->h_status) cudaFreeHost(x->h_status);
    }
    delete x;
    return ARCHI_OK;
}

}  // extern "C"
