// Corpus cosine top-k on the tensor cores (north_star item 5): dense Q.D^T contraction with fused norms and a
// per-query threshold filter in the TMEM epilogue, followed by exact fp32 rescoring of the few survivors so that the
// returned ids (and scores) are bit-identical to the exact path / oracle (oracle/retrieval_oracle.py).
//
//   pass 0   the first TKC_SEED docs go through the exact kernels of topk.cu -> an exact running top-k per query
//   pass i   topk_tc_filter_kernel: tcgen05.mma kind::tf32 directly on the fp32 rows (the tensor core reads the top
//            19 bits: |approx - exact| <= 2^-9 ||q|| ||d||), accumulators double-buffered in TMEM; the epilogue keeps
//            doc j for query q iff  dot_approx >= (tau_q - margin) * ||q|| * ||d_j||  where tau_q is the exact k-th
//            best so far (a lower bound of the final one, so no true top-k doc can be dropped) and appends its id
//            to the query's candidate list;
//            topk_rescore_select_kernel: one warp per query re-scores the candidates with the oracle's sequential
//            fp32 arithmetic and inserts them into the running list ordered by (score desc, id asc); tau_q rises.
//   The corpus is visited in geometrically growing chunks so the expected number of candidates per query per pass
//   stays ~ k * (chunk / docs seen so far).  A candidate list that would overflow raises a device flag and the caller
//   falls back to the exact path (never observed on the synthetic corpora).
//
// CTA = 256 threads: 128 queries (TMEM lanes) x a range of doc tiles of 128 docs; Q tile resident in shared memory
// (64 KB, SW128 K-major), doc tiles double-buffered (2 x 64 KB) and filled with cp.async straight from the fp32
// corpus -- no conversion pass, no extra copy of the corpus.
#include "common.cuh"
#include "tc_common.cuh"
#include "topk_common.cuh"
#include <math.h>

namespace dssm {
namespace tkc {

using namespace dssm::tc;

constexpr int QT = 128, DT = 128, DIM = 128, THREADS = 256;
constexpr int KBLKS = DIM / 32;              // 128-byte swizzle rows per operand row
constexpr int TILE_BYTES = QT * DIM * 4;     // 64 KB
constexpr int KBLK_BYTES = QT * 128;         // one [128 rows x 32 floats] block
constexpr int CAP = 2048;                    // candidate ids per query per pass
constexpr float MARGIN = 2.5e-3f;            // > 2^-9 (tf32 truncation of both operands) + fp32 slack
constexpr int SEED_DOCS = 1024;              // docs of the first filter pass (thresholds still -inf: all of them become candidates)

// stage one [128 x 128] fp32 tile (rows row0.. of X, `rows_valid` of them real) into SW128 K-major blocks with cp.async.
// Chunk i of a thread always goes to the same place of the tile: 16-byte chunk id = tid + 256*i  ->  row id/32,
// k-block (id%32)/8, chunk-in-row id%8 (consecutive threads read consecutive 16 B of a 512-byte row).
__device__ __forceinline__ uint32_t tile_dst_off(int id) {
    const int r = id >> 5, cc = id & 31;
    return (uint32_t)((cc >> 3) * KBLK_BYTES + r * 128 + (((cc & 7) ^ (r & 7)) << 4));
}
__device__ __forceinline__ void load_tile_async(uint32_t smem_tile, const float* __restrict__ X, int64_t row0, int rows_valid, int tid) {
    const float* src0 = X + row0 * DIM + (size_t)tid * 4;  // chunk id -> element offset id*4 (row-major tile is contiguous)
    if (rows_valid >= QT) {
#pragma unroll
        for (int i = 0; i < (QT * DIM / 4) / THREADS; ++i) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_tile + tile_dst_off(tid + i * THREADS)),
                         "l"(src0 + (size_t)i * THREADS * 4)
                         : "memory");
        }
    } else {
#pragma unroll
        for (int i = 0; i < (QT * DIM / 4) / THREADS; ++i) {
            const int id = tid + i * THREADS;
            const bool ok = (id >> 5) < rows_valid;  // zero-fill rows past the end
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_tile + tile_dst_off(id)),
                         "l"(ok ? src0 + (size_t)i * THREADS * 4 : X), "r"(ok ? 16 : 0)
                         : "memory");
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1)
topk_tc_filter_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ docs, int64_t doc_lo, int64_t doc_hi,
                      const float* __restrict__ dn, const float* __restrict__ tq /* (tau - margin) * ||q|| */,
                      const float* __restrict__ qn, int id_base /* id of doc row 0 */, int2* __restrict__ cand, int* __restrict__ cand_cnt,
                      int* __restrict__ overflow, int tiles_per_split, int reserve /* candidate slots per reservation */) {
    extern __shared__ char smem_raw[];
    __shared__ uint64_t mma_done[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_dn[3][DT];  // doc norms of the tiles in flight (t-1 being drained, t in the MMA, t+1 loading)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t q_tile = smem_u32(smem);
    const uint32_t d_tile0 = q_tile + TILE_BYTES;  // stage s at d_tile0 + s * TILE_BYTES
    float4* stash = reinterpret_cast<float4*>(smem + 3 * TILE_BYTES);  // [THREADS][8] float4, survivors' slow path

    const int q0 = blockIdx.x * QT;
    const int64_t n_docs = doc_hi - doc_lo;
    const int total_tiles = (int)((n_docs + DT - 1) / DT);
    const int t_begin = blockIdx.y * tiles_per_split;
    const int t_end = min(total_tiles, t_begin + tiles_per_split);
    if (t_begin >= t_end) return;

    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    if (tid == 0) {
        mbar_init(&mma_done[0], 1);
        mbar_init(&mma_done[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // resident Q tile + first doc tile
    load_tile_async(q_tile, Q, q0, min(QT, nq - q0), tid);
    {
        const int64_t d0 = doc_lo + (int64_t)t_begin * DT;
        load_tile_async(d_tile0, docs, d0, (int)((doc_hi - d0) < (int64_t)DT ? (doc_hi - d0) : (int64_t)DT), tid);
        if (tid < DT) s_dn[0][tid] = (d0 + tid < doc_hi) ? __ldg(dn + d0 + tid) : 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = make_idesc_tf32(QT, DT);

    // this thread's query (TMEM lane) and its scaled threshold
    const int lane_grp = warp & 3, half = warp >> 2;  // 8 warps: 4 TMEM lane groups x 2 column halves
    const int q = q0 + lane_grp * 32 + lane;
    const float t_q = q < nq ? __ldg(tq + q) : INFINITY;
    const float inv_qn = q < nq ? 1.0f / __ldg(qn + q) : 0.f;

    // Candidate slots are reserved in runs of `reserve` per (thread, CTA) -- one atomic per run, not per candidate: a
    // survivor is rare per thread but nearly every chunk has one somewhere in the warp, and every lane of the warp then
    // waited out the atomic's L2 round trip (see topk_bf16.cu).  Unused slots of the last run get the id -1 (skipped by
    // the rescoring pass).
    int res_pos = 0, res_left = 0;
    // drain the accumulator of this CTA's tile number `tr` (relative index): approx dot -> threshold test -> append
    auto wait_mma = [&](int tr) {
        mbar_wait(&mma_done[tr & 1], (uint32_t)((tr >> 1) & 1));
        tc_fence_after();
    };
    auto drain = [&](int tr) {
        const int bb = tr & 1;
        const float4* nrm4 = reinterpret_cast<const float4*>(s_dn[tr % 3]);
        const int64_t d0 = doc_lo + (int64_t)(t_begin + tr) * DT;
        const int nvalid = (int)((doc_hi - d0) < (int64_t)DT ? (doc_hi - d0) : (int64_t)DT);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            uint32_t r[32];
            const int col0 = half * 64 + ch * 32;
            tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(bb * DT + col0), r);
            // keep iff cos_approx >= tau - margin  <=>  dot >= (tau - margin) * ||q|| * ||d||  (norms >= 0).  Survivors
            // are rare: collect the 32 verdicts in a bit mask, branch once.
            // verdicts -> bit mask with one FSET.BF (1.0f / 0.0f) and one FFMA (k = 2k + verdict, exact: 16 bits per
            // accumulator) per score, highest column first -- see topk_bf16.cu
            float k_hi = 0.f, k_lo = 0.f;
#pragma unroll
            for (int j4 = 3; j4 >= 0; --j4) {
                const float4 nl = nrm4[(col0 >> 2) + j4], nh = nrm4[(col0 >> 2) + 4 + j4];
                const float tl[4] = {t_q * nl.x, t_q * nl.y, t_q * nl.z, t_q * nl.w};
                const float th[4] = {t_q * nh.x, t_q * nh.y, t_q * nh.z, t_q * nh.w};
#pragma unroll
                for (int e = 3; e >= 0; --e) {
                    float v_hi, v_lo;
                    asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(v_hi) : "f"(__uint_as_float(r[16 + 4 * j4 + e])), "f"(th[e]));
                    asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(v_lo) : "f"(__uint_as_float(r[4 * j4 + e])), "f"(tl[e]));
                    k_hi = fmaf(k_hi, 2.0f, v_hi);
                    k_lo = fmaf(k_lo, 2.0f, v_lo);
                }
            }
            uint32_t keep = ((uint32_t)k_hi << 16) | (uint32_t)k_lo;
            if (nvalid - col0 < 32) keep &= (nvalid - col0 <= 0) ? 0u : (0xffffffffu >> (32 - (nvalid - col0)));  // partial last tile
            if (keep) {
                // this lane has survivors in the chunk: park its 32 dots in its own shared-memory row (XOR-swizzled
                // float4 slots) so that the few survivors can be fetched by index without forcing r[] out of registers
                float4* row = stash + (size_t)tid * 8;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    row[i ^ (lane & 7)] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                      __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                const float* rowf = reinterpret_cast<const float*>(row);
                while (keep) {
                    const int j = __ffs(keep) - 1;
                    keep &= keep - 1;
                    if (res_left == 0) {
                        res_pos = atomicAdd(cand_cnt + q, reserve);
                        res_left = reserve;
                        if (res_pos + reserve > CAP) {
                            *overflow = 1;
                            res_left = res_pos < CAP ? CAP - res_pos : 0;
                        }
                    }
                    if (res_left > 0) {
                        // id + the approximate cosine (the rescoring pass skips candidates that can no longer enter the list)
                        const float dot = rowf[(((j >> 2) ^ (lane & 7)) << 2) + (j & 3)];
                        const float nd = s_dn[tr % 3][col0 + j];
                        cand[(size_t)q * CAP + res_pos] = make_int2(id_base + (int)(d0 + col0 + j), __float_as_int(dot * inv_qn / nd));
                        ++res_pos;
                        --res_left;
                    }
                }
            }
        }
        tc_fence_before();
    };

    for (int t = t_begin; t < t_end; ++t) {
        const int b = (t - t_begin) & 1;
        // tile t has landed (cp.async of this thread) -> make it visible to the tensor core, then everybody syncs
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < KBLKS; ++kb) {
                const uint64_t da = make_desc_k_sw128(q_tile + kb * KBLK_BYTES);
                const uint64_t db = make_desc_k_sw128(d_tile0 + b * TILE_BYTES + kb * KBLK_BYTES);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma_tf32(tmem_d + (uint32_t)(b * DT), da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (kb | ks) ? 1u : 0u);
            }
            mma_commit(&mma_done[b]);
        }
        // While the tensor core works on tile t: once MMA(t-1) has retired, its smem stage is free -> start the
        // cp.async prefetch of tile t+1 into it FIRST, then drain accumulator t-1 (TMEM reads do not touch smem), so
        // the copy overlaps the epilogue instead of being waited for right after it was issued.
        const int tr = t - t_begin;
        if (tr > 0) wait_mma(tr - 1);
        if (t + 1 < t_end) {
            const int64_t d0 = doc_lo + (int64_t)(t + 1) * DT;
            load_tile_async(d_tile0 + (b ^ 1) * TILE_BYTES, docs, d0, (int)((doc_hi - d0) < (int64_t)DT ? (doc_hi - d0) : (int64_t)DT), tid);
            if (tid < DT) s_dn[(tr + 1) % 3][tid] = (d0 + tid < doc_hi) ? __ldg(dn + d0 + tid) : 0.f;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (tr > 0) drain(tr - 1);
    }
    wait_mma(t_end - 1 - t_begin);
    drain(t_end - 1 - t_begin);  // last tile
    for (; res_left > 0; --res_left, ++res_pos) cand[(size_t)q * CAP + res_pos] = make_int2(-1, __float_as_int(-INFINITY));
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 256);
}

// One warp per query: exact scores of the candidates (sequential fp32 multiply-then-add, as the oracle), insertion
// into the running top-k ordered by (score desc, id asc), new threshold, candidate counter reset.
__global__ void __launch_bounds__(128)
topk_rescore_select_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ docs /* row 0 = id id_base */, int id_base,
                           int d, const float* __restrict__ qn, const float* __restrict__ dn, int k, const int2* __restrict__ cand,
                           int* __restrict__ cand_cnt, float* __restrict__ run_s, int* __restrict__ run_i,
                           int* __restrict__ run_cnt, float* __restrict__ tq, float margin /* bound on |approx - exact| of the filter */) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * 4 + w;
    if (q >= nq) return;
    float* ls = sm + (size_t)w * 4 * k;  // two (score, id) list buffers per warp, ping-ponged by the batch merge
    int* li = reinterpret_cast<int*>(ls + k);
    float* ls2 = ls + 2 * k;
    int* li2 = reinterpret_cast<int*>(ls2 + k);
    int cnt = run_cnt[q];
    for (int i = lane; i < cnt; i += 32) {
        ls[i] = run_s[(size_t)q * k + i];
        li[i] = run_i[(size_t)q * k + i];
    }
    __syncwarp();
    const int n = min(cand_cnt[q], CAP);
    const float nqv = qn[q];
    // per-warp staging of the query row only.  Every lane then walks ITS candidate's fp32 row straight from L2 with
    // 128-bit loads, 8 in flight, and sums strictly in t order (the oracle's arithmetic).  (Round 1 staged 32 candidate
    // rows per warp in shared memory: 17 KB per warp capped the SM at 12 warps, and ncu showed the kernel issue-bound at
    // 33 % issue utilisation with 15 % of the warp slots filled -- the rows are L2 hits (84 %), DRAM is idle.)
    float* sq = reinterpret_cast<float*>(sm + (size_t)4 * 4 * k) + (size_t)w * (d + 4);
    for (int t = lane; t < d; t += 32) sq[t] = __ldg(Q + (size_t)q * d + t);
    __syncwarp();
    const bool vec = (d % 32 == 0);
    for (int base = 0; base < n; base += 32) {
        // candidates whose approximate cosine cannot reach the current k-th best any more are dropped before the
        // (expensive) exact scoring: |approx - exact| <= margin, and the k-th best only rises
        int cid0 = 0x7fffffff;
        bool alive = false;
        if (base + lane < n) {
            const int2 c2 = cand[(size_t)q * CAP + base + lane];
            cid0 = c2.x;
            const float approx = __int_as_float(c2.y);
            // id -1: an unused slot of a filter thread's reserved run (topk_bf16.cu); NaN approx (zero-norm doc) stays alive
            alive = cid0 >= 0 && ((cnt < k) || !(approx + margin < ls[k - 1]));
        }
        const unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
        const int na = __popc(alive_mask);
        if (na == 0) continue;
        // compact: lane c takes the c-th alive candidate
        const int src_lane = lane < na ? __fns(alive_mask, 0, lane + 1) : 0;
        int cid = __shfl_sync(0xffffffffu, cid0, src_lane);
        if (lane >= na) cid = 0x7fffffff;
        float s = -INFINITY;
        if (lane < na) {
            const float* mine = docs + (size_t)(cid - id_base) * d;
            float acc = 0.f;
            if (vec) {
                const float4* m4 = reinterpret_cast<const float4*>(mine);
                for (int t0 = 0; t0 < d; t0 += 32) {
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = __ldg(m4 + (t0 >> 2) + u);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 qv = *reinterpret_cast<const float4*>(sq + t0 + 4 * u);
                        acc = __fadd_rn(acc, __fmul_rn(qv.x, v[u].x));
                        acc = __fadd_rn(acc, __fmul_rn(qv.y, v[u].y));
                        acc = __fadd_rn(acc, __fmul_rn(qv.z, v[u].z));
                        acc = __fadd_rn(acc, __fmul_rn(qv.w, v[u].w));
                    }
                }
            } else {
                for (int t = 0; t < d; ++t) acc = __fadd_rn(acc, __fmul_rn(sq[t], __ldg(mine + t)));
            }
            s = __fdiv_rn(acc, __fmul_rn(nqv, dn[cid - id_base]));
            if (s != s) s = -INFINITY;
            s = s + 0.0f;
        }
        __syncwarp();
        // merge the batch into the running list in one step (topk_common.cuh; round 1 inserted candidate by candidate, a
        // warp-wide O(k) shift each: ~120 instructions x ~250 insertions per query and pass made the kernel issue-bound)
        cnt = topk_merge_batch(ls, li, ls2, li2, cnt, k, s, cid, lane < na, lane);
    }
    for (int i = lane; i < cnt; i += 32) {
        run_s[(size_t)q * k + i] = ls[i];
        run_i[(size_t)q * k + i] = li[i];
    }
    // fewer than k entries so far: the tail reads (-inf, max id), so a caller can tell an incomplete list from a full one
    for (int i = cnt + lane; i < k; i += 32) {
        run_s[(size_t)q * k + i] = -INFINITY;
        run_i[(size_t)q * k + i] = 0x7fffffff;
    }
    if (lane == 0) {
        run_cnt[q] = cnt;
        cand_cnt[q] = 0;
        // threshold of the next filter pass, pre-multiplied by ||q||: keep iff dot >= (tau - margin) * ||q|| * ||d||
        const float tau = cnt == k ? ls[k - 1] : -INFINITY;
        tq[q] = (tau - margin) * nqv;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Between two filter passes: new thresholds from the APPROXIMATE scores, no row fetches.  cand[q][0 .. cand_cnt[q]) holds
// the query's "bag" (survivors of earlier selections) followed by the candidates the last filter pass appended.  One warp
// per query:
//   key      monotone uint image of the approximate cosine; 0 = not usable for the bound (unused reserved slot, NaN, doc of
//            zero norm -- its exact cosine is NaN = -inf whatever the filter computed)
//   tau      a lower bound of the k-th largest usable approximate cosine: 3 rounds of a 256-bucket select over the key range,
//            i.e. the lower edge of the last bucket that holds the k-th largest (-inf while fewer than k usable entries exist)
//   bound    k docs have exact >= approx - m >= tau - m, so the FINAL exact k-th best is >= tau - m, so any doc of the final
//            top-k has approx >= exact - m >= tau - 2m: the bag keeps exactly the entries with approx >= tau - 2m (all of
//            them while tau = -inf) and the next filter pass keeps docs with dot >= (tau - 2m) * ||q|| * ||d||.
// tau only rises, so an entry dropped here can never qualify again; after the last pass the bag (k plus the docs within
// 2m of the k-th best) is re-scored exactly ONCE by topk_rescore_select_kernel.  (Round 2 first version re-scored the
// ~4k candidates of EVERY pass exactly: ~2500 random 512-byte row fetches per query, 1.1 of the 3.0 ms at C5.)
__device__ __forceinline__ uint32_t key_of(float a) {
    const uint32_t u = __float_as_uint(a);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float float_of_key(uint32_t key) {
    const uint32_t u = (key & 0x80000000u) ? (key ^ 0x80000000u) : ~key;
    return __uint_as_float(u);
}

__global__ void __launch_bounds__(128)
topk_approx_select_kernel(int nq, int k, int2* __restrict__ cand, int* __restrict__ cand_cnt, const float* __restrict__ dn, int id_base,
                          const float* __restrict__ qn, float* __restrict__ tq, float margin2) {
    extern __shared__ uint32_t sel_sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * 4 + w;
    if (q >= nq) return;
    uint32_t* keys = sel_sm + (size_t)w * CAP;
    int* hist = reinterpret_cast<int*>(sel_sm + (size_t)4 * CAP) + w * 256;
    int2* mine = cand + (size_t)q * CAP;
    const int n = min(cand_cnt[q], CAP);
    int usable = 0;
    for (int i0 = lane; i0 < n; i0 += 4 * 32) {  // 4 entries per lane in flight: entry, then its doc norm (two dependent loads)
        int2 c[4];
        float nv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = (i0 + 32 * u < n) ? mine[i0 + 32 * u] : make_int2(-1, 0);
#pragma unroll
        for (int u = 0; u < 4; ++u) nv[u] = c[u].x >= 0 ? __ldg(dn + (c[u].x - id_base)) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + 32 * u >= n) continue;
            const float a = __int_as_float(c[u].y);
            const uint32_t key = (c[u].x >= 0 && a == a && nv[u] > 0.f) ? key_of(a) : 0u;
            keys[i0 + 32 * u] = key;
            usable += key != 0 ? 1 : 0;
        }
    }
    usable = __reduce_add_sync(0xffffffffu, usable);
    __syncwarp();
    float tau = -INFINITY;
    uint32_t thr_key = 0;  // keep every entry with a real id
    // A query of zero norm has NaN (= -inf) scores everywhere: its top-k is the k lowest ids.  The first pass (thresholds
    // -inf) covers docs [0, 1024) -- at least the first k -- so keep exactly those and close the threshold (+inf).
    const float nqv = qn[q];
    const bool dead_query = !(nqv > 0.f);
    if (!dead_query && usable >= k) {
        // range of the usable keys: the cosines of one query share their leading key bits, so the buckets of every round are
        // laid over [lo, hi], not over fixed digit positions (with fixed digits the first round put every entry into one
        // or two bins: a 32-way shared-memory atomic conflict per step)
        uint32_t lo = 0xffffffffu, hi = 0u;
        for (int i = lane; i < n; i += 32) {
            const uint32_t key = keys[i];
            if (key != 0) { lo = min(lo, key); hi = max(hi, key); }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        int want = k;  // rank (from the top) still to be located inside [lo, hi]
#pragma unroll 1
        for (int round = 0; round < 3 && hi > lo; ++round) {
            const uint32_t span = hi - lo;
            const int shift = max(0, 32 - __clz(span) - 8);  // (key - lo) >> shift  in [0, 255]
            for (int b = lane; b < 256; b += 32) hist[b] = 0;
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                const uint32_t key = keys[i];
                if (key >= lo && key <= hi && key != 0) atomicAdd(hist + ((key - lo) >> shift), 1);
            }
            __syncwarp();
            // lane l owns bins 8l .. 8l+7; suffix sums from the top bin down
            int c8[8], mine_sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c8[j] = hist[lane * 8 + j]; mine_sum += c8[j]; }
            int suf = mine_sum;  // inclusive suffix sum over the lanes >= this one (Hillis-Steele)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_down_sync(0xffffffffu, suf, o);
                if (lane + o < 32) suf += v;
            }
            const int above = suf - mine_sum;  // entries in bins owned by higher lanes
            int digit = -1, rank_in = 0;
            if (above < want && above + mine_sum >= want) {  // the wanted rank falls into this lane's bins
                int acc = above;
#pragma unroll
                for (int j = 7; j >= 0; --j) {
                    if (digit < 0 && acc + c8[j] >= want) { digit = lane * 8 + j; rank_in = want - acc; }
                    acc += c8[j];
                }
            }
            const unsigned who = __ballot_sync(0xffffffffu, digit >= 0);
            const int src = __ffs(who) - 1;
            digit = __shfl_sync(0xffffffffu, digit, src);
            want = __shfl_sync(0xffffffffu, rank_in, src);
            const uint32_t new_lo = lo + ((uint32_t)digit << shift);
            const uint32_t width = shift >= 32 ? 0xffffffffu : ((1u << shift) - 1u);
            hi = min(hi, new_lo + width);
            lo = new_lo;
            __syncwarp();
        }
        const uint32_t tau_key = lo;  // lower edge of the last bucket: <= the k-th largest key
        tau = float_of_key(tau_key);
        thr_key = key_of(tau - margin2);
        if (thr_key == 0) thr_key = 1;
    }
    // compaction in place: chunk by chunk, every lane reads its entry before anybody writes (targets are <= sources)
    // (the next chunk's entries are fetched before this chunk's survivors are written: its positions lie behind every target)
    int out = 0;
    int2 nxt = lane < n ? mine[lane] : make_int2(-1, 0);
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const int2 c = nxt;
        if (base + 32 < n) nxt = (i + 32 < n) ? mine[i + 32] : make_int2(-1, 0);
        const bool keep = i < n && c.x >= 0 && (dead_query ? (c.x - id_base < k) : (thr_key == 0 || keys[i] >= thr_key));
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) mine[out + __popc(m & ((1u << lane) - 1u))] = c;
        out += __popc(m);
        __syncwarp();
    }
    if (lane == 0) {
        cand_cnt[q] = out;
        tq[q] = dead_query ? (n > 0 ? INFINITY : -INFINITY) : (tau - margin2) * nqv;
    }
}

struct Ws {
    float *qn, *dn, *S, *run_s, *tq;
    int *run_i, *run_cnt, *cand_cnt, *overflow;
    int2* cand;
    size_t bytes;
};
static Ws carve(void* ws, int nq, int64_t nd, int k) {
    Arena a(ws, (size_t)-1);
    Ws w;
    const int nq_pad = (nq + QT - 1) / QT * QT;
    w.qn = a.take<float>(nq_pad);
    w.dn = a.take<float>((size_t)nd);
    const int seed = nd < SEED_DOCS ? (int)nd : SEED_DOCS;
    w.S = a.take<float>((size_t)nq * seed);
    w.run_s = a.take<float>((size_t)nq * k);
    w.run_i = a.take<int>((size_t)nq * k);
    w.run_cnt = a.take<int>(nq_pad);
    w.tq = a.take<float>(nq_pad);
    w.cand = a.take<int2>((size_t)nq_pad * CAP);
    w.cand_cnt = a.take<int>(nq_pad);
    w.overflow = a.take<int>(4);
    w.bytes = a.off;
    return w;
}

}  // namespace tkc

// rescoring + selection + next thresholds, shared by the tf32 (fp32 corpus) and bf16 (corpus index) filters
int topk_rescore_select(const float* Q, int nq, const float* docs, int id_base, int d, const float* qn, const float* dn, int k,
                        const int2* cand, int* cand_cnt, float* run_s, int* run_i, int* run_cnt, float* tq, float margin, cudaStream_t st) {
    static PerDeviceOnce once;
    if (once.need()) CUDA_TRY(cudaFuncSetAttribute(tkc::topk_rescore_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const size_t sel_smem = (size_t)4 * 4 * k * sizeof(float) + (size_t)4 * (d + 4) * sizeof(float);  // 2 lists per warp + its query row
    tkc::topk_rescore_select_kernel<<<cdiv(nq, 4), 128, sel_smem, st>>>(Q, nq, docs, id_base, d, qn, dn, k, cand, cand_cnt, run_s, run_i, run_cnt,
                                                                       tq, margin);
    LAUNCH_CHECK("topk_rescore_select");
    return DSSM_OK;
}

int topk_approx_select(int nq, int k, int2* cand, int* cand_cnt, const float* dn, int id_base, const float* qn, float* tq, float margin,
                       cudaStream_t st) {
    const size_t smem = (size_t)4 * tkc::CAP * sizeof(uint32_t) + (size_t)4 * 256 * sizeof(int);
    tkc::topk_approx_select_kernel<<<cdiv(nq, 4), 128, smem, st>>>(nq, k, cand, cand_cnt, dn, id_base, qn, tq, 2.0f * margin);
    LAUNCH_CHECK("topk_approx_select");
    return DSSM_OK;
}

// exact path helpers (topk.cu)
int topk_row_norms(const float* X, int64_t n, int d, float* out, cudaStream_t st);
int topk_exact_chunk(const float* Q, int nq, const float* docs, int64_t doc0, int cd, int d, const float* qn, const float* dn, float* S,
                     int ldS, int id0, int k, float* run_s, int* run_i, int* run_cnt, cudaStream_t st);

}  // namespace dssm

using namespace dssm;

extern "C" size_t dssm_corpus_topk_tc_workspace_bytes(int32_t nq, int64_t nd, int32_t d, int32_t k) {
    (void)d;
    if (nq <= 0 || nd <= 0 || k <= 0) return 0;
    return tkc::carve(nullptr, nq, nd, k).bytes;
}

// Tensor-core filtered top-k; same outputs as dssm_corpus_topk.  *overflow_flag (device int) is set to 1 if a
// candidate list overflowed (results then incomplete: rerun with dssm_corpus_topk).  Requires d == 128.
extern "C" int dssm_corpus_topk_tc(const float* Q, int32_t nq, const float* docs, int64_t nd, int32_t d, int32_t k, int32_t id_offset,
                                   float* out_scores, int32_t* out_ids, int32_t* overflow_flag, void* workspace, size_t workspace_bytes,
                                   dssm_stream_t stream) {
    DSSM_REQUIRE(Q && docs && out_scores && out_ids && overflow_flag && workspace, DSSM_ERR_BAD_ARG, "dssm_corpus_topk_tc: null pointer");
    DSSM_REQUIRE(d == tkc::DIM, DSSM_ERR_BAD_SHAPE, "dssm_corpus_topk_tc: d must be %d (got %d); use dssm_corpus_topk", tkc::DIM, d);
    DSSM_REQUIRE(nq > 0 && nd > 0 && k > 0 && k <= nd && k <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_corpus_topk_tc: bad shape");
    DSSM_REQUIRE(nd + (int64_t)id_offset < (int64_t)1 << 31, DSSM_ERR_BAD_SHAPE, "dssm_corpus_topk_tc: ids overflow int32");
    DSSM_REQUIRE(aligned16(Q) && aligned16(docs), DSSM_ERR_BAD_ALIGN, "dssm_corpus_topk_tc: Q/docs must be 16-byte aligned");
    DSSM_REQUIRE(workspace_bytes >= tkc::carve(nullptr, nq, nd, k).bytes, DSSM_ERR_WORKSPACE, "dssm_corpus_topk_tc: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    tkc::Ws w = tkc::carve(workspace, nq, nd, k);
    const int nq_pad = (nq + tkc::QT - 1) / tkc::QT * tkc::QT;
    static PerDeviceOnce once;
    const size_t smem = 3 * (size_t)tkc::TILE_BYTES + (size_t)tkc::THREADS * 32 * sizeof(float) + 1024;
    if (once.need()) CUDA_TRY(cudaFuncSetAttribute(tkc::topk_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int rc = topk_row_norms(Q, nq, d, w.qn, st);
    if (rc != DSSM_OK) return rc;
    rc = topk_row_norms(docs, nd, d, w.dn, st);
    if (rc != DSSM_OK) return rc;
    CUDA_TRY(cudaMemsetAsync(w.run_cnt, 0, (size_t)nq_pad * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(w.cand_cnt, 0, (size_t)nq_pad * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(w.overflow, 0, sizeof(int), st));
    // thresholds start at -inf (empty bags): the first, short pass lets every doc through and seeds the bound
    rc = topk_approx_select(nq, k, w.cand, w.cand_cnt, w.dn, id_offset, w.qn, w.tq, tkc::MARGIN, st);
    if (rc != DSSM_OK) return rc;
    const int n_qtiles = nq_pad / tkc::QT;
    const int growth = k <= 160 ? 4 : 2;  // candidates per pass ~ growth * k on top of the bag: keep them inside CAP
    int64_t lo = 0, chunk = tkc::SEED_DOCS;
    while (lo < nd) {
        const int64_t hi = (nd - lo <= chunk + chunk / 2) ? nd : lo + chunk;  // fold a short tail into the last pass
        const int tiles = (int)((hi - lo + tkc::DT - 1) / tkc::DT);
        int splits = sm_count() / n_qtiles;
        if (splits < 1) splits = 1;
        if (splits > tiles) splits = tiles;
        const int tps = (tiles + splits - 1) / splits;
        dim3 grid(n_qtiles, (tiles + tps - 1) / tps);
        int reserve = tkc::CAP / (8 * (int)grid.y);  // unused tails of the reserved runs stay below CAP / 8 per query
        reserve = reserve > 32 ? 32 : (reserve < 1 ? 1 : reserve);
        tkc::topk_tc_filter_kernel<<<grid, tkc::THREADS, smem, st>>>(Q, nq, docs, lo, hi, w.dn, w.tq, w.qn, id_offset, w.cand, w.cand_cnt,
                                                                     w.overflow, tps, reserve);
        LAUNCH_CHECK("topk_tc_filter");
        rc = topk_approx_select(nq, k, w.cand, w.cand_cnt, w.dn, id_offset, w.qn, w.tq, tkc::MARGIN, st);
        if (rc != DSSM_OK) return rc;
        lo = hi;
        chunk *= growth;
    }
    // the bags hold every doc that can be in the exact top-k: re-score them exactly, once
    rc = topk_rescore_select(Q, nq, docs, id_offset, d, w.qn, w.dn, k, w.cand, w.cand_cnt, w.run_s, w.run_i, w.run_cnt, w.tq, tkc::MARGIN, st);
    if (rc != DSSM_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_scores, w.run_s, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(out_ids, w.run_i, (size_t)nq * k * sizeof(int), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(overflow_flag, w.overflow, sizeof(int), cudaMemcpyDeviceToDevice, st));
    return DSSM_OK;
}
