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
constexpr int SEED_DOCS = 16384;

// stage one [128 x 128] fp32 tile (rows row0.. of X, `rows_valid` of them real) into SW128 K-major blocks with cp.async
__device__ __forceinline__ void load_tile_async(char* smem_tile, const float* __restrict__ X, int64_t row0, int rows_valid, int tid) {
#pragma unroll
    for (int i = 0; i < (QT * DIM / 4) / THREADS; ++i) {  // 4096 16-byte chunks / 256 threads
        const int id = tid + i * THREADS;
        const int r = id >> 5, cc = id & 31;  // 32 chunks per row: consecutive threads read consecutive 16 B
        const int kb = cc >> 3, c = cc & 7;
        const uint32_t dst = smem_u32(smem_tile + kb * KBLK_BYTES + r * 128 + ((c ^ (r & 7)) << 4));
        const float* src = X + (row0 + (r < rows_valid ? r : 0)) * DIM + cc * 4;
        const int bytes = r < rows_valid ? 16 : 0;  // zero-fill rows past the end
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
}

__global__ void __launch_bounds__(THREADS, 1)
topk_tc_filter_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ docs, int64_t doc_lo, int64_t doc_hi,
                      const float* __restrict__ dn, const float* __restrict__ tq /* (tau - margin) * ||q|| */,
                      int id_base /* id of doc row 0 */, int* __restrict__ cand, int* __restrict__ cand_cnt,
                      int* __restrict__ overflow, int tiles_per_split) {
    extern __shared__ char smem_raw[];
    __shared__ uint64_t mma_done[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_dn[2][DT];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    char* q_tile = smem;
    char* d_tile[2] = {smem + TILE_BYTES, smem + 2 * TILE_BYTES};

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
        load_tile_async(d_tile[0], docs, d0, (int)((doc_hi - d0) < (int64_t)DT ? (doc_hi - d0) : (int64_t)DT), tid);
        if (tid < DT) s_dn[0][tid] = (d0 + tid < doc_hi) ? __ldg(dn + d0 + tid) : 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = make_idesc_tf32(QT, DT);

    // this thread's query (TMEM lane) and its scaled threshold
    const int lane_grp = warp & 3, half = warp >> 2;
    const int q = q0 + lane_grp * 32 + lane;
    const float t_q = q < nq ? __ldg(tq + q) : INFINITY;

    // drain the accumulator of this CTA's tile number `tr` (relative index): approx dot -> threshold test -> append
    auto drain = [&](int tr) {
        const int bb = tr & 1;
        mbar_wait(&mma_done[bb], (uint32_t)((tr >> 1) & 1));
        tc_fence_after();
        const int64_t d0 = doc_lo + (int64_t)(t_begin + tr) * DT;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            uint32_t r[32];
            const int col0 = half * 64 + ch * 32;
            tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(bb * DT + col0), r);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float dot = __uint_as_float(r[j]);
                const float nd = s_dn[bb][col0 + j];
                // keep iff cos_approx >= tau - margin  <=>  dot >= (tau - margin) * ||q|| * ||d||   (norms are >= 0)
                if (dot >= t_q * nd && d0 + col0 + j < doc_hi) {
                    const int pos = atomicAdd(cand_cnt + q, 1);
                    if (pos < CAP) cand[(size_t)q * CAP + pos] = id_base + (int)(d0 + col0 + j);
                    else *overflow = 1;
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
                const uint64_t da = make_desc_k_sw128(smem_u32(q_tile + kb * KBLK_BYTES));
                const uint64_t db = make_desc_k_sw128(smem_u32(d_tile[b] + kb * KBLK_BYTES));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma_tf32(tmem_d + (uint32_t)(b * DT), da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (kb | ks) ? 1u : 0u);
            }
            mma_commit(&mma_done[b]);
        }
        // while the tensor core works on tile t: drain tile t-1 (its smem stage and accumulator become free) ...
        if (t > t_begin) drain(t - 1 - t_begin);
        __syncthreads();  // everyone is done with stage (t-1)&1 and its norms
        // ... and prefetch tile t+1 into the stage tile t-1 used
        if (t + 1 < t_end) {
            const int nb = b ^ 1;
            const int64_t d0 = doc_lo + (int64_t)(t + 1) * DT;
            load_tile_async(d_tile[nb], docs, d0, (int)((doc_hi - d0) < (int64_t)DT ? (doc_hi - d0) : (int64_t)DT), tid);
            if (tid < DT) s_dn[nb][tid] = (d0 + tid < doc_hi) ? __ldg(dn + d0 + tid) : 0.f;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    drain(t_end - 1 - t_begin);  // last tile
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 256);
}

// One warp per query: exact scores of the candidates (sequential fp32 multiply-then-add, as the oracle), insertion
// into the running top-k ordered by (score desc, id asc), new threshold, candidate counter reset.
__global__ void __launch_bounds__(128)
topk_rescore_select_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ docs /* row 0 = id id_base */, int id_base,
                           int d, const float* __restrict__ qn, const float* __restrict__ dn, int k, const int* __restrict__ cand,
                           int* __restrict__ cand_cnt, float* __restrict__ run_s, int* __restrict__ run_i,
                           int* __restrict__ run_cnt, float* __restrict__ tq) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * 4 + w;
    if (q >= nq) return;
    float* ls = sm + (size_t)w * 2 * k;
    int* li = reinterpret_cast<int*>(ls + k);
    int cnt = run_cnt[q];
    for (int i = lane; i < cnt; i += 32) {
        ls[i] = run_s[(size_t)q * k + i];
        li[i] = run_i[(size_t)q * k + i];
    }
    __syncwarp();
    const int n = min(cand_cnt[q], CAP);
    const float* qrow = Q + (size_t)q * d;
    const float nqv = qn[q];
    for (int base = 0; base < n; base += 32) {
        int cid = 0x7fffffff;
        float s = -INFINITY;
        if (base + lane < n) {
            cid = cand[(size_t)q * CAP + base + lane];
            const float* drow = docs + (size_t)(cid - id_base) * d;
            float acc = 0.f;
            for (int t = 0; t < d; ++t) acc = __fadd_rn(acc, __fmul_rn(__ldg(qrow + t), __ldg(drow + t)));
            s = __fdiv_rn(acc, __fmul_rn(nqv, dn[cid - id_base]));
            if (s != s) s = -INFINITY;
            s = s + 0.0f;
        }
        unsigned mask = __ballot_sync(0xffffffffu, base + lane < n);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const float cs = __shfl_sync(0xffffffffu, s, src);
            const int ci = __shfl_sync(0xffffffffu, cid, src);
            if (cnt == k) {  // must beat the current worst under (score desc, id asc)
                const float wsc = ls[k - 1];
                const int wid = li[k - 1];
                if (!(cs > wsc || (cs == wsc && ci < wid))) continue;
            }
            int ahead = 0;
            for (int i = lane; i < cnt; i += 32) ahead += (ls[i] > cs || (ls[i] == cs && li[i] < ci)) ? 1 : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ahead += __shfl_xor_sync(0xffffffffu, ahead, o);
            const int newcnt = cnt < k ? cnt + 1 : k;
            float tmp_s[32];
            int tmp_i[32];
            int nt = 0;
            for (int i = ahead + lane; i < newcnt - 1; i += 32) {
                if (nt < 32) { tmp_s[nt] = ls[i]; tmp_i[nt] = li[i]; }
                ++nt;
            }
            __syncwarp();
            nt = 0;
            for (int i = ahead + lane; i < newcnt - 1; i += 32) {
                if (nt < 32) { ls[i + 1] = tmp_s[nt]; li[i + 1] = tmp_i[nt]; }
                ++nt;
            }
            if (lane == 0) { ls[ahead] = cs; li[ahead] = ci; }
            cnt = newcnt;
            __syncwarp();
        }
    }
    for (int i = lane; i < cnt; i += 32) {
        run_s[(size_t)q * k + i] = ls[i];
        run_i[(size_t)q * k + i] = li[i];
    }
    if (lane == 0) {
        run_cnt[q] = cnt;
        cand_cnt[q] = 0;
        // threshold of the next filter pass, pre-multiplied by ||q||: keep iff dot >= (tau - margin) * ||q|| * ||d||
        const float tau = cnt == k ? ls[k - 1] : -INFINITY;
        tq[q] = (tau - MARGIN) * nqv;
    }
}

struct Ws {
    float *qn, *dn, *S, *run_s, *tq;
    int *run_i, *run_cnt, *cand, *cand_cnt, *overflow;
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
    w.cand = a.take<int>((size_t)nq_pad * CAP);
    w.cand_cnt = a.take<int>(nq_pad);
    w.overflow = a.take<int>(4);
    w.bytes = a.off;
    return w;
}

}  // namespace tkc

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
    static bool attr_set = false;
    const size_t smem = 3 * (size_t)tkc::TILE_BYTES + 1024;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(tkc::topk_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int rc = topk_row_norms(Q, nq, d, w.qn, st);
    if (rc != DSSM_OK) return rc;
    rc = topk_row_norms(docs, nd, d, w.dn, st);
    if (rc != DSSM_OK) return rc;
    CUDA_TRY(cudaMemsetAsync(w.run_cnt, 0, (size_t)nq_pad * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(w.cand_cnt, 0, (size_t)nq_pad * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(w.overflow, 0, sizeof(int), st));
    // pass 0: exact top-k of the first `seed` docs
    const int seed = nd < tkc::SEED_DOCS ? (int)nd : tkc::SEED_DOCS;
    rc = topk_exact_chunk(Q, nq, docs, 0, seed, d, w.qn, w.dn, w.S, seed, id_offset, k, w.run_s, w.run_i, w.run_cnt, st);
    if (rc != DSSM_OK) return rc;
    const size_t sel_smem = (size_t)4 * 2 * k * sizeof(float);
    // thresholds from the seed (no candidates yet)
    tkc::topk_rescore_select_kernel<<<cdiv(nq, 4), 128, sel_smem, st>>>(Q, nq, docs, id_offset, d, w.qn, w.dn, k, w.cand, w.cand_cnt,
                                                                       w.run_s, w.run_i, w.run_cnt, w.tq);
    LAUNCH_CHECK("topk_rescore_select(seed)");
    const int n_qtiles = nq_pad / tkc::QT;
    int64_t lo = seed, chunk = 4 * (int64_t)tkc::SEED_DOCS;
    while (lo < nd) {
        const int64_t hi = (nd - lo <= chunk + chunk / 2) ? nd : lo + chunk;  // fold a short tail into the last pass
        const int tiles = (int)((hi - lo + tkc::DT - 1) / tkc::DT);
        int splits = sm_count() / n_qtiles;
        if (splits < 1) splits = 1;
        if (splits > tiles) splits = tiles;
        const int tps = (tiles + splits - 1) / splits;
        dim3 grid(n_qtiles, (tiles + tps - 1) / tps);
        tkc::topk_tc_filter_kernel<<<grid, tkc::THREADS, smem, st>>>(Q, nq, docs, lo, hi, w.dn, w.tq, id_offset, w.cand, w.cand_cnt,
                                                                     w.overflow, tps);
        LAUNCH_CHECK("topk_tc_filter");
        tkc::topk_rescore_select_kernel<<<cdiv(nq, 4), 128, sel_smem, st>>>(Q, nq, docs, id_offset, d, w.qn, w.dn, k, w.cand,
                                                                           w.cand_cnt, w.run_s, w.run_i, w.run_cnt, w.tq);
        LAUNCH_CHECK("topk_rescore_select");
        lo = hi;
        chunk *= 4;
    }
    CUDA_TRY(cudaMemcpyAsync(out_scores, w.run_s, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(out_ids, w.run_i, (size_t)nq * k * sizeof(int), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(overflow_flag, w.overflow, sizeof(int), cudaMemcpyDeviceToDevice, st));
    return DSSM_OK;
}
