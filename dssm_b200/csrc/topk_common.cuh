// Running top-k list of one query, kept by one warp in shared memory, shared by the exact path (topk.cu) and the
// rescoring pass of the tensor-core paths (topk_tc.cu).  Order: (score desc, id asc) -- tf.nn.top_k(sorted=True),
// utils/tf_ranking_utils.py:47 -- a strict total order because ids are unique.
#pragma once
#include "common.cuh"

namespace dssm {

__device__ __forceinline__ bool topk_ahead(float as, int ai, float bs, int bi) { return as > bs || (as == bs && ai < bi); }

// Merge up to 32 new (score, id) pairs -- one per lane, `elig` false for lanes without one -- into the sorted list
// (ls, li)[0..cnt) in ONE step; the result lands in (ls2, li2) and the two buffers are swapped.  Returns the new count.
//   1. lanes whose pair cannot beat the current worst of a full list drop out;
//   2. the others are sorted across the lanes, best first (bitonic network, 15 shuffle steps);
//   3. merge path: a list element moves to  own index + #batch elements ahead of it,  a batch element to  own rank +
//      #list elements ahead of it (binary search); ranks >= k fall off.
// Equal to inserting the pairs one by one (strict total order), at ~1/6 of the instructions when most of a batch enters.
__device__ __forceinline__ int topk_merge_batch(float*& ls, int*& li, float*& ls2, int*& li2, int cnt, int k, float s, int id, bool elig,
                                                int lane) {
    if (elig && cnt == k) elig = topk_ahead(s, id, ls[k - 1], li[k - 1]);
    const int m = __popc(__ballot_sync(0xffffffffu, elig));
    if (m == 0) return cnt;
    float bs = elig ? s : -INFINITY;
    int bi = elig ? id : 0x7fffffff;
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, j2);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, j2);
            const bool keep_best = ((lane & j2) == 0) == ((lane & k2) == 0);
            if (keep_best ? topk_ahead(os, oi, bs, bi) : topk_ahead(bs, bi, os, oi)) { bs = os; bi = oi; }
        }
    }
    // lane j < m now holds the j-th best of the batch
    const int newcnt = min(k, cnt + m);
    for (int i0 = 0; i0 < cnt; i0 += 32) {
        const int i = i0 + lane;
        const float xs = i < cnt ? ls[i] : 0.f;
        const int xi = i < cnt ? li[i] : 0;
        int ahead = 0;
        for (int j = 0; j < m; ++j) {
            const float ns = __shfl_sync(0xffffffffu, bs, j);
            const int ni = __shfl_sync(0xffffffffu, bi, j);
            ahead += topk_ahead(ns, ni, xs, xi) ? 1 : 0;
        }
        if (i < cnt && i + ahead < k) { ls2[i + ahead] = xs; li2[i + ahead] = xi; }
    }
    if (lane < m) {
        int lo = 0, hi = cnt;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (topk_ahead(ls[mid], li[mid], bs, bi)) lo = mid + 1; else hi = mid;
        }
        if (lane + lo < k) { ls2[lane + lo] = bs; li2[lane + lo] = bi; }
    }
    __syncwarp();
    float* tf = ls; ls = ls2; ls2 = tf;
    int* ti = li; li = li2; li2 = ti;
    return newcnt;
}

}  // namespace dssm
