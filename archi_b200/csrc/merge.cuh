// merge.cuh -- shard merge of sorted k-lists, shared by merge_lists_kernel (after an NCCL all-gather)
// and exchange_merge_kernel (peer-memory exchange).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "topk.cuh"

namespace archi {

// One warp merges the n_lists (<= 32) sorted k-lists of query q: lane l walks list l, each step a warp
// arg-best on (key desc, id asc) picks the next output.  List l starts s_stride fp32 / i_stride int64
// elements after list l-1.  Loads bypass L1 (__ldcg): the lists may have been written by a peer GPU.
__device__ __forceinline__ void merge_query_lists(const float *scores, const long long *ids, size_t s_stride,
                                                  size_t i_stride, int n_lists, int nq, int q, int k, int larger,
                                                  float *out_scores, long long *out_ids, int lane)
{
    (void)nq;
    const bool has_list = lane < n_lists;
    // list l starts s_stride (i_stride) elements after list l-1: dense [n_lists, nq, k] arrays or the
    // records of one packed all-gather buffer
    const size_t in_list = (size_t)q * k;
    const float *my_scores = scores + (has_list ? (size_t)lane * s_stride + in_list : 0);
    const long long *my_ids = ids + (has_list ? (size_t)lane * i_stride + in_list : 0);
    int pos = 0;
    for (int out = 0; out < k; ++out) {
        // head of my list
        float key = -CUDART_INF_F;
        long long id = -1;
        float sc = CUDART_NAN_F;
        if (has_list && pos < k) {
            id = __ldcg(my_ids + pos);
            sc = __ldcg(my_scores + pos);
            if (id >= 0 && sc == sc) key = larger ? sc : -sc;
            else id = -1;
        }
        // warp arg-best on (key desc, id asc); exhausted lists carry id -1
        float bk = key;
        long long bi = id;
        int bl = lane;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const float ok = __shfl_xor_sync(kFull, bk, d);
            const long long oi = __shfl_xor_sync(kFull, bi, d);
            const int ol = __shfl_xor_sync(kFull, bl, d);
            const bool mine_valid = bi >= 0, other_valid = oi >= 0;
            bool take = false;
            if (other_valid && !mine_valid) take = true;
            else if (other_valid && mine_valid)
                take = ok > bk || (ok == bk && (oi < bi || (oi == bi && ol < bl)));
            else if (!other_valid && !mine_valid)
                take = ol < bl;
            if (take) {
                bk = ok;
                bi = oi;
                bl = ol;
            }
        }
        const float win_sc = __shfl_sync(kFull, sc, bl);
        if (lane == 0) {
            out_scores[(size_t)q * k + out] = bi >= 0 ? win_sc : CUDART_NAN_F;
            out_ids[(size_t)q * k + out] = bi;
        }
        if (bi < 0) {
            // every list is exhausted: pad the rest
            if (lane == 0)
                for (int o = out + 1; o < k; ++o) {
                    out_scores[(size_t)q * k + o] = CUDART_NAN_F;
                    out_ids[(size_t)q * k + o] = -1;
                }
            break;
        }
        if (lane == bl) ++pos;
    }
}

}  // namespace archi
