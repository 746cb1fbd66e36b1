// Corpus cosine top-k over a bf16 corpus INDEX (north_star item 5; SURVEY 8d "report fp32-storage and bf16-storage
// variants").  The index is built once per corpus (dssm_corpus_index_build): the fp32 row norms the exact rescoring
// needs, and the rows NORMALISED, rounded to bf16 and laid out as ready-to-use tensor-core tiles
//      [doc tile of 128][k-block of 64][128 rows x 128 B, SWIZZLE_128B]
// so that a query batch streams 256 B per document (half of the fp32 rows) with one cp.async.bulk per 32 KB tile and no
// conversion or swizzling at query time.  The filter then is a warp-specialised tcgen05 pipeline:
//      warp 16  lane 0: bulk copies of doc tiles into a 3-stage ring (mbarrier complete_tx)
//      warp 17  lane 0: tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), TWO resident query tiles (256 queries) per
//               doc tile, accumulators double-buffered: 2 tiles x 2 buffers x 128 columns = the 512 TMEM columns
//      warps 0-15 drain TMEM: approx cosine * ||q|| = <q_bf16, d_hat_bf16> against the query's threshold (the doc norm is
//               already inside d_hat: one compare per score), survivors appended to the query's candidate list
// |approx - exact| <= 2^-8 (both operands rounded to nearest bf16, Cauchy-Schwarz) < MARGIN_BF16, the threshold is a lower
// bound of the final k-th best, so no true top-k document is dropped; survivors are re-scored EXACTLY from the fp32 rows
// (topk_rescore_select_kernel, the oracle's sequential arithmetic), so ids and scores stay bit-identical to the exact path.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <math.h>

namespace dssm {
namespace tkb {

using namespace dssm::tc;

constexpr int QT = 128, DT = 128, DIM = 128;
constexpr int QTILES = 2;                      // query tiles resident per CTA
constexpr int KB = 2;                          // k-blocks of 64 bf16 (= one 128-byte swizzle row)
constexpr int KBLK_BYTES = 128 * 128;          // [128 rows x 128 B]
constexpr int TILE_BYTES = KB * KBLK_BYTES;    // 32 KB per 128 x 128 bf16 tile
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 16;                  // (TMEM lane quadrant) x (query tile) x (column half of the doc tile)
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = EPI_THREADS + 64;      // + loader warp + MMA warp
constexpr int RESERVE = 32;                    // candidate slots a thread reserves per atomic (see the epilogue); fewer when a
                                               // query is split over many CTAs, so the unused tails stay below CAP / 4
constexpr int CAP = 2048;                      // candidate ids per query per pass (same as topk_tc.cu)
constexpr float MARGIN_BF16 = 4.5e-3f;         // > 2^-8 + fp32 slack

// InstrDescriptor for kind::f16 with bf16 operands: c_format F32 (1) [4,6), a/b_format BF16 (1) [7,10)/[10,13), K-major,
// n_dim = N>>3 [17,23), m_dim = M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte offset of (row r, 16-byte chunk c of the 128-byte row) inside a [128 x 128 B] SW128 block
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// ---- index build: rows -> (optionally normalised) bf16 tile image ---------------------------------------------------------
// one warp per row: lane l holds elements 4l..4l+3; inv = 1/norm (or 1 for the query image)
__global__ void __launch_bounds__(256)
build_image_kernel(const float* __restrict__ X, int64_t n, int64_t n_pad /* rows of the image: [n, n_pad) are zero */,
                   const float* __restrict__ norms /* NULL: no normalisation */, char* __restrict__ img) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_pad) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n) {
        v = __ldg(reinterpret_cast<const float4*>(X + row * DIM) + lane);
        if (norms) {
            const float nr = __ldg(norms + row);
            const float inv = nr > 0.f ? 1.0f / nr : 0.f;  // zero rows stay zero (their exact cosine is NaN = -inf anyway)
            v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        }
    }
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    // element e = 4*lane: k-block e/64, 16-byte chunk (e%64)/8, 8 bytes into the chunk if lane is odd
    const int64_t tile = row / DT;
    const int r = (int)(row - tile * DT);
    const int kb = lane >> 4, c = (lane & 15) >> 1, half8 = lane & 1;
    char* dst = img + tile * TILE_BYTES + kb * KBLK_BYTES + sw128_off(r, c) + half8 * 8;
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst) = pk;
}

// ---- the filter -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 1)
topk_bf16_filter_kernel(const char* __restrict__ q_img, int nq, const char* __restrict__ d_img, int64_t doc_lo, int64_t doc_hi,
                        const float* __restrict__ tq /* (tau - margin) * ||q|| */, const float* __restrict__ qn, int id_base,
                        int2* __restrict__ cand, int* __restrict__ cand_cnt, int* __restrict__ overflow, int tiles_per_split,
                        int reserve /* candidate slots per reservation, <= RESERVE */) {
    extern __shared__ char smem_raw[];
    __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2], q_full;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    char* q_sm = smem;                                  // QTILES x 32 KB
    char* d_sm = smem + QTILES * TILE_BYTES;            // STAGES x 32 KB
    float4* stash = reinterpret_cast<float4*>(d_sm + STAGES * TILE_BYTES);  // [EPI_THREADS][4] float4

    // doc tiles are addressed relative to the index (doc_lo is a multiple of DT: caller's contract)
    const int64_t tile_lo = doc_lo / DT;
    const int total_tiles = (int)((doc_hi - doc_lo + DT - 1) / DT);
    const int t_begin = blockIdx.y * tiles_per_split;
    const int t_end = min(total_tiles, t_begin + tiles_per_split);
    const int n_tiles = t_end - t_begin;
    if (n_tiles <= 0) return;
    const int qt0 = blockIdx.x * QTILES;  // first query tile of this CTA

    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], EPI_THREADS);
        }
        mbar_init(&q_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp == EPI_WARPS) {
        // ------------------------------------------------------------------ loader
        if (lane == 0) {
            // the query image is padded (with zero rows) to a multiple of QTILES tiles, so both tiles can always be copied
            bulk_copy_g2s(q_sm, q_img + (size_t)qt0 * TILE_BYTES, (uint32_t)(QTILES * TILE_BYTES), &q_full);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % STAGES, use = t / STAGES;
                if (use > 0) mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
                bulk_copy_g2s(d_sm + s * TILE_BYTES, d_img + (size_t)(tile_lo + t_begin + t) * TILE_BYTES, (uint32_t)TILE_BYTES, &full[s]);
            }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(QT, DT);
            mbar_wait(&q_full, 0);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % STAGES, b = t & 1;
                mbar_wait(&full[s], (uint32_t)((t / STAGES) & 1));
                if (t >= 2) mbar_wait(&acc_empty[b], (uint32_t)(((t >> 1) - 1) & 1));  // the epilogue has drained this buffer
                tc_fence_after();
#pragma unroll
                for (int qt = 0; qt < QTILES; ++qt) {
                    const uint32_t acc = tmem_d + (uint32_t)((qt * 2 + b) * DT);
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) {
                        const uint64_t da = make_desc_k_sw128(smem_u32(q_sm + qt * TILE_BYTES + kb * KBLK_BYTES));
                        const uint64_t db = make_desc_k_sw128(smem_u32(d_sm + s * TILE_BYTES + kb * KBLK_BYTES));
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)  // 16 bf16 = 32 bytes per MMA
                            mma_bf16(acc, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (kb | ks) ? 1u : 0u);
                    }
                }
                mma_commit(&empty[s]);     // the stage may be refilled once these MMAs have read it
                mma_commit(&acc_full[b]);  // both accumulators of buffer b are complete
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 0-15)
        // warp w: TMEM lanes 32*(w%4).., query tile (w/4)%2, columns 64*(w/8).. of the doc tile.  A thread owns ONE query.
        //
        // Candidate slots are RESERVED in runs of RESERVE per (thread, CTA) with one atomic, not one atomic per candidate:
        // a survivor is rare per thread (a few % of the 32-column chunks) but almost every chunk has one somewhere in the
        // warp, so with an atomic per candidate every warp sat out a full L2 atomic round trip on nearly every chunk --
        // ncu (profiles/r2_ncu_bf16_filter.txt): 3100 cycles per doc tile against ~1000 of MMA work, tensor pipe 22 %.
        // Slots of the last run that stay unused are filled with the id -1 the rescoring pass skips.
        const int lane_grp = warp & 3, qt = (warp >> 2) & 1, half = warp >> 3;
        const int q = (qt0 + qt) * QT + lane_grp * 32 + lane;
        const float t_q = q < nq ? __ldg(tq + q) : INFINITY;
        const float inv_qn = q < nq ? 1.0f / __ldg(qn + q) : 0.f;
        int res_pos = 0, res_left = 0;
        float4* row = stash + (size_t)tid * 4;  // 16 scores of this thread (one half chunk)
        const float* rowf = reinterpret_cast<const float*>(row);
        auto emit = [&](uint32_t keep16, const uint32_t* r16, int64_t doc_first) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                row[i ^ (lane & 3)] = make_float4(__uint_as_float(r16[4 * i]), __uint_as_float(r16[4 * i + 1]),
                                                  __uint_as_float(r16[4 * i + 2]), __uint_as_float(r16[4 * i + 3]));
            while (keep16) {
                const int j = __ffs(keep16) - 1;
                keep16 &= keep16 - 1;
                if (res_left == 0) {
                    res_pos = atomicAdd(cand_cnt + q, reserve);
                    res_left = reserve;
                    if (res_pos + reserve > CAP) {
                        *overflow = 1;
                        res_left = res_pos < CAP ? CAP - res_pos : 0;
                    }
                }
                if (res_left > 0) {
                    const float dot = rowf[(((j >> 2) ^ (lane & 3)) << 2) + (j & 3)];
                    cand[(size_t)q * CAP + res_pos] = make_int2(id_base + (int)(doc_first + j), __float_as_int(dot * inv_qn));
                    ++res_pos;
                    --res_left;
                }
            }
        };
        for (int t = 0; t < n_tiles; ++t) {
            const int b = t & 1;
            mbar_wait(&acc_full[b], (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            const int64_t d0 = doc_lo + (int64_t)(t_begin + t) * DT;
            const int nvalid = (int)((doc_hi - d0) < (int64_t)DT ? (doc_hi - d0) : (int64_t)DT);
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                const int ch = half * 2 + c2;
                uint32_t r[32];
                tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)((qt * 2 + b) * DT + ch * 32), r);
                // 32 verdicts -> bit mask.  One FSET (0 / 0xffffffff, the 16-lane ALU pipe) and one multiply-add
                // keep = 2 keep - verdict (IMAD: the FMA pipe, otherwise idle here) per score, highest column first so that
                // column j lands on bit j.  (FSETP + SEL + IADD3, all on the ALU pipe, made the epilogue the bound of the
                // kernel: profiles/r2_ncu_bf16_filter.txt.)
                float k_hi = 0.f, k_lo = 0.f;  // exact: 16 bits each
#pragma unroll
                for (int j = 15; j >= 0; --j) {
                    float v_hi, v_lo;
                    asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(v_hi) : "f"(__uint_as_float(r[16 + j])), "f"(t_q));
                    asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(v_lo) : "f"(__uint_as_float(r[j])), "f"(t_q));
                    k_hi = fmaf(k_hi, 2.0f, v_hi);
                    k_lo = fmaf(k_lo, 2.0f, v_lo);
                }
                uint32_t keep = ((uint32_t)k_hi << 16) | (uint32_t)k_lo;
                const int col0 = ch * 32;
                if (nvalid - col0 < 32) keep &= (nvalid - col0 <= 0) ? 0u : (0xffffffffu >> (32 - (nvalid - col0)));  // partial last tile
                if (keep & 0xffffu) emit(keep & 0xffffu, r, d0 + col0);
                if (keep >> 16) emit(keep >> 16, r + 16, d0 + col0 + 16);
            }
            tc_fence_before();
            mbar_arrive_cta(&acc_empty[b]);
        }
        for (; res_left > 0; --res_left, ++res_pos) cand[(size_t)q * CAP + res_pos] = make_int2(-1, __float_as_int(-INFINITY));
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 512);
}

}  // namespace tkb

// shared with topk_tc.cu / topk.cu
int topk_row_norms(const float* X, int64_t n, int d, float* out, cudaStream_t st);
int topk_exact_chunk(const float* Q, int nq, const float* docs, int64_t doc0, int cd, int d, const float* qn, const float* dn, float* S,
                     int ldS, int id0, int k, float* run_s, int* run_i, int* run_cnt, cudaStream_t st);
int topk_rescore_select(const float* Q, int nq, const float* docs, int id_base, int d, const float* qn, const float* dn, int k,
                        const int2* cand, int* cand_cnt, float* run_s, int* run_i, int* run_cnt, float* tq, float margin, cudaStream_t st);
int topk_approx_select(int nq, int k, int2* cand, int* cand_cnt, const float* dn, int id_base, const float* qn, float* tq, float margin,
                       cudaStream_t st);

}  // namespace dssm

using namespace dssm;

// index = [ row norms fp32 (nd, padded to 256 B) | bf16 tile image (ceil(nd/128) tiles x 32 KB) ]
static size_t index_norm_bytes(int64_t nd) { return align_up((size_t)nd * sizeof(float), 1024); }

extern "C" size_t dssm_corpus_index_bytes(int64_t nd, int32_t d) {
    if (nd <= 0 || d != tkb::DIM) return 0;
    return index_norm_bytes(nd) + (size_t)((nd + tkb::DT - 1) / tkb::DT) * tkb::TILE_BYTES;
}

extern "C" int dssm_corpus_index_build(const float* docs, int64_t nd, int32_t d, void* index, size_t index_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(docs && index, DSSM_ERR_BAD_ARG, "dssm_corpus_index_build: null pointer");
    DSSM_REQUIRE(d == tkb::DIM, DSSM_ERR_BAD_SHAPE, "dssm_corpus_index_build: d must be %d (got %d)", tkb::DIM, d);
    DSSM_REQUIRE(nd > 0 && index_bytes >= dssm_corpus_index_bytes(nd, d), DSSM_ERR_WORKSPACE, "dssm_corpus_index_build: index buffer too small");
    DSSM_REQUIRE(aligned16(docs) && (reinterpret_cast<uintptr_t>(index) & 1023u) == 0, DSSM_ERR_BAD_ALIGN,
                 "dssm_corpus_index_build: docs must be 16-byte and the index 1024-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    float* dn = (float*)index;
    int rc = topk_row_norms(docs, nd, d, dn, st);
    if (rc != DSSM_OK) return rc;
    const int64_t n_pad = (nd + tkb::DT - 1) / tkb::DT * tkb::DT;
    tkb::build_image_kernel<<<cdiv(n_pad, 8), 256, 0, st>>>(docs, nd, n_pad, dn, (char*)index + index_norm_bytes(nd));
    LAUNCH_CHECK("build_image(docs)");
    return DSSM_OK;
}

namespace {
struct WsB {
    float *qn, *S, *run_s, *tq;
    int *run_i, *run_cnt, *cand_cnt, *overflow;
    int2* cand;
    char* q_img;
    size_t bytes;
};
WsB carve_b(void* ws, int nq, int k) {
    Arena a(ws, (size_t)-1);
    WsB w;
    const int nq_pad = (nq + tkb::QT * tkb::QTILES - 1) / (tkb::QT * tkb::QTILES) * (tkb::QT * tkb::QTILES);
    w.q_img = a.take<char>((size_t)(nq_pad / tkb::QT) * tkb::TILE_BYTES + 1024);
    w.qn = a.take<float>(nq_pad);
    w.S = a.take<float>((size_t)nq * 4096);
    w.run_s = a.take<float>((size_t)nq * k);
    w.run_i = a.take<int>((size_t)nq * k);
    w.run_cnt = a.take<int>(nq_pad);
    w.tq = a.take<float>(nq_pad);
    w.cand = a.take<int2>((size_t)nq_pad * tkb::CAP);
    w.cand_cnt = a.take<int>(nq_pad);
    w.overflow = a.take<int>(4);
    w.bytes = a.off;
    return w;
}
}  // namespace

extern "C" size_t dssm_corpus_topk_indexed_workspace_bytes(int32_t nq, int32_t k) {
    if (nq <= 0 || k <= 0) return 0;
    return carve_b(nullptr, nq, k).bytes + 1024;
}

// Same contract and outputs as dssm_corpus_topk / dssm_corpus_topk_tc; `index` from dssm_corpus_index_build over the same docs.
extern "C" int dssm_corpus_topk_indexed(const float* Q, int32_t nq, const float* docs, const void* index, int64_t nd, int32_t d, int32_t k,
                                        int32_t id_offset, float* out_scores, int32_t* out_ids, int32_t* overflow_flag, void* workspace,
                                        size_t workspace_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(Q && docs && index && out_scores && out_ids && overflow_flag && workspace, DSSM_ERR_BAD_ARG, "dssm_corpus_topk_indexed: null pointer");
    DSSM_REQUIRE(d == tkb::DIM, DSSM_ERR_BAD_SHAPE, "dssm_corpus_topk_indexed: d must be %d (got %d)", tkb::DIM, d);
    DSSM_REQUIRE(nq > 0 && nd > 0 && k > 0 && k <= nd && k <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_corpus_topk_indexed: bad shape");
    DSSM_REQUIRE(nd + (int64_t)id_offset < (int64_t)1 << 31, DSSM_ERR_BAD_SHAPE, "dssm_corpus_topk_indexed: ids overflow int32");
    DSSM_REQUIRE(aligned16(Q) && aligned16(docs) && (reinterpret_cast<uintptr_t>(index) & 1023u) == 0, DSSM_ERR_BAD_ALIGN,
                 "dssm_corpus_topk_indexed: Q/docs must be 16-byte, the index 1024-byte aligned");
    DSSM_REQUIRE(workspace_bytes >= dssm_corpus_topk_indexed_workspace_bytes(nq, k), DSSM_ERR_WORKSPACE, "dssm_corpus_topk_indexed: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    void* ws_al = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
    WsB w = carve_b(ws_al, nq, k);
    const float* dn = (const float*)index;
    const char* d_img = (const char*)index + index_norm_bytes(nd);
    const int nq_pad = (nq + tkb::QT * tkb::QTILES - 1) / (tkb::QT * tkb::QTILES) * (tkb::QT * tkb::QTILES);
    static PerDeviceOnce once;
    const size_t smem = (size_t)(tkb::QTILES + tkb::STAGES) * tkb::TILE_BYTES + (size_t)tkb::EPI_THREADS * 16 * sizeof(float) + 1024;
    if (once.need()) CUDA_TRY(cudaFuncSetAttribute(tkb::topk_bf16_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int rc = topk_row_norms(Q, nq, d, w.qn, st);
    if (rc != DSSM_OK) return rc;
    tkb::build_image_kernel<<<cdiv(nq_pad, 8), 256, 0, st>>>(Q, nq, nq_pad, nullptr, w.q_img);  // rows past nq are zero
    LAUNCH_CHECK("build_image(Q)");
    CUDA_TRY(cudaMemsetAsync(w.run_cnt, 0, (size_t)nq_pad * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(w.cand_cnt, 0, (size_t)nq_pad * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(w.overflow, 0, sizeof(int), st));
    // thresholds start at -inf (empty bags): the first, short pass lets every doc through and seeds the bound; between the
    // passes topk_approx_select_kernel (topk_tc.cu) raises the thresholds from the APPROXIMATE scores and keeps, per query,
    // the bag of docs that can still be in the exact top-k; the bag is re-scored exactly once, at the end
    rc = topk_approx_select(nq, k, w.cand, w.cand_cnt, dn, id_offset, w.qn, w.tq, tkb::MARGIN_BF16, st);
    if (rc != DSSM_OK) return rc;
    const int n_qgroups = nq_pad / (tkb::QT * tkb::QTILES);
    const int growth = k <= 160 ? 4 : 2;  // candidates per pass ~ growth * k on top of the bag: keep them inside CAP
    int64_t lo = 0, chunk = 1024;
    while (lo < nd) {
        const int64_t hi = (nd - lo <= chunk + chunk / 2) ? nd : lo + chunk;  // fold a short tail into the last pass
        const int tiles = (int)((hi - lo + tkb::DT - 1) / tkb::DT);
        int splits = sm_count() / n_qgroups;
        if (splits < 1) splits = 1;
        if (splits > tiles) splits = tiles;
        const int tps = (tiles + splits - 1) / splits;
        dim3 grid(n_qgroups, (tiles + tps - 1) / tps);
        int reserve = tkb::CAP / (8 * (int)grid.y);  // unused tails of the reserved runs stay below CAP / 8 per query
        reserve = reserve > tkb::RESERVE ? tkb::RESERVE : (reserve < 1 ? 1 : reserve);
        tkb::topk_bf16_filter_kernel<<<grid, tkb::THREADS, smem, st>>>(w.q_img, nq, d_img, lo, hi, w.tq, w.qn, id_offset, w.cand, w.cand_cnt,
                                                                       w.overflow, tps, reserve);
        LAUNCH_CHECK("topk_bf16_filter");
        rc = topk_approx_select(nq, k, w.cand, w.cand_cnt, dn, id_offset, w.qn, w.tq, tkb::MARGIN_BF16, st);
        if (rc != DSSM_OK) return rc;
        lo = hi;
        chunk *= growth;
    }
    rc = topk_rescore_select(Q, nq, docs, id_offset, d, w.qn, dn, k, w.cand, w.cand_cnt, w.run_s, w.run_i, w.run_cnt, w.tq, tkb::MARGIN_BF16, st);
    if (rc != DSSM_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_scores, w.run_s, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(out_ids, w.run_i, (size_t)nq * k * sizeof(int), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(overflow_flag, w.overflow, sizeof(int), cudaMemcpyDeviceToDevice, st));
    return DSSM_OK;
}
