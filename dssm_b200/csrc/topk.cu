// Corpus cosine top-k (no reference function exists).  Exact-arithmetic path:
// every score is the sequential fp32 multiply-then-add dot product of oracle/retrieval_oracle.py divided by
// the product of the two sequential norms, so doc ids are bit-exact against the oracle, including ties
// (score descending, id ascending; NaN ranks as -inf, -0 as +0; tf.nn.top_k(sorted=True) contract,
// utils/tf_ranking_utils.py:47).
//
//   row_norm_seq      ||x|| with the oracle's summation order
//   exact_scores      64x64 (query x doc) register-tiled tile, 4x4 per thread, operands staged t-major in smem
//   chunk_select      one warp per query streams the chunk's scores against its running k-th best and
//                     inserts the rare survivors into a sorted list kept in shared memory
// The corpus is processed in chunks so the score scratch stays bounded (workspace query).
#include "common.cuh"
#include "topk_common.cuh"
#include <math.h>

namespace dssm {

constexpr int TK_QT = 64, TK_DT = 64, TK_TT = 32;
constexpr int TK_CHUNK = 16384;  // docs per chunk

// ||x|| with the oracle's strictly sequential fp32 sum of squares.  128 rows per block: 128 x 32 column slabs are
// staged through shared memory (coalesced 128-byte row segments in, conflict-free column walks out, stride 33), and
// thread r accumulates row r in t order across the slabs.
__global__ void __launch_bounds__(128) row_norm_seq_kernel(const float* __restrict__ X, int64_t n, int d, float* __restrict__ out) {
    __shared__ float slab[128][33];
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * 128;
    float acc = 0.f;
    for (int c0 = 0; c0 < d; c0 += 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {  // 4096 elements / 128 threads, column fastest
            const int e = tid + i * 128;
            const int rr = e >> 5, cc = e & 31;
            slab[rr][cc] = (r0 + rr < n && c0 + cc < d) ? __ldg(X + (r0 + rr) * d + c0 + cc) : 0.f;
        }
        __syncthreads();
        const int cmax = min(32, d - c0);
        for (int t = 0; t < cmax; ++t) acc = __fadd_rn(acc, __fmul_rn(slab[tid][t], slab[tid][t]));
        __syncthreads();
    }
    if (r0 + tid < n) out[r0 + tid] = __fsqrt_rn(acc);
}

// S[q, j] = key(score(q, doc0 + j)) for j < cd
__global__ void __launch_bounds__(256)
exact_scores_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ docs, int64_t doc0, int cd, int d,
                    const float* __restrict__ qn, const float* __restrict__ dn, float* __restrict__ S, int ldS) {
    __shared__ __align__(16) float Qs[TK_TT][TK_QT + 4];
    __shared__ __align__(16) float Ds[TK_TT][TK_DT + 4];
    const int tid = threadIdx.x;
    const int q0 = blockIdx.y * TK_QT, j0 = blockIdx.x * TK_DT;
    const int ty = tid / 16, tx = tid % 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int t0 = 0; t0 < d; t0 += TK_TT) {
        // stage 64 x 32 of each operand, t fastest over tid (coalesced 128 B row segments)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * 256;
            const int tt = e % TK_TT, rr = e / TK_TT;
            const int t = t0 + tt;
            const int q = q0 + rr;
            Qs[tt][rr] = (q < nq && t < d) ? __ldg(Q + (size_t)q * d + t) : 0.f;
            const int j = j0 + rr;
            Ds[tt][rr] = (j < cd && t < d) ? __ldg(docs + (size_t)(doc0 + j) * d + t) : 0.f;
        }
        __syncthreads();
        const int tmax = min(TK_TT, d - t0);
        for (int tt = 0; tt < tmax; ++tt) {
            const float4 a = *reinterpret_cast<const float4*>(&Qs[tt][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Ds[tt][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(av[i], bv[j]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ty * 4 + i;
        if (q >= nq) continue;
        const float nqv = qn[q];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jj = j0 + tx * 4 + j;
            if (jj >= cd) continue;
            float s = __fdiv_rn(acc[i][j], __fmul_rn(nqv, dn[doc0 + jj]));
            if (s != s) s = -INFINITY;  // NaN (zero-norm row) ranks last
            s = s + 0.0f;               // -0 -> +0
            S[(size_t)q * ldS + jj] = s;
        }
    }
}

// running list per query: scores/ids [k] sorted (score desc, id asc), cnt valid entries
__global__ void __launch_bounds__(128)
chunk_select_kernel(const float* __restrict__ S, int ldS, int nq, int cd, int id0, int k, float* __restrict__ run_s,
                    int* __restrict__ run_i, int* __restrict__ run_cnt) {
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
    const float* srow = S + (size_t)q * ldS;
    for (int base = 0; base < cd; base += 32) {
        const int j = base + lane;
        const float s = j < cd ? srow[j] : -INFINITY;
        cnt = topk_merge_batch(ls, li, ls2, li2, cnt, k, s, id0 + j, j < cd, lane);
    }
    for (int i = lane; i < cnt; i += 32) {
        run_s[(size_t)q * k + i] = ls[i];
        run_i[(size_t)q * k + i] = li[i];
    }
    if (lane == 0) run_cnt[q] = cnt;
}

// merge n_parts sorted lists per query (each [k], sorted by (score desc, id asc)) into the global top-k.
// One warp per query, repeated selection of the best head (n_parts <= 32).
__global__ void __launch_bounds__(128)
topk_merge_kernel(const float* __restrict__ ps, const int* __restrict__ pi, int n_parts, int nq, int k,
                  float* __restrict__ out_s, int* __restrict__ out_i) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * 4 + w;
    if (q >= nq) return;
    int head = 0;  // lane p walks part p
    for (int o = 0; o < k; ++o) {
        float s = -INFINITY;
        int id = 0x7fffffff;
        bool valid = lane < n_parts && head < k;
        if (valid) {
            s = ps[((size_t)lane * nq + q) * k + head];
            id = pi[((size_t)lane * nq + q) * k + head];
            if (s != s) s = -INFINITY;
            s = s + 0.0f;
        }
        // best = max score, then min id; invalid lanes lose against any valid one
        float bs = s;
        int bi = id, bv = valid ? 1 : 0, bl = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            const int ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int ol = __shfl_xor_sync(0xffffffffu, bl, off);
            const bool take = (ov && !bv) || (ov == bv && (os > bs || (os == bs && (oi < bi || (oi == bi && ol < bl)))));
            if (take) { bs = os; bi = oi; bv = ov; bl = ol; }
        }
        if (lane == 0) {
            out_s[(size_t)q * k + o] = bs;
            out_i[(size_t)q * k + o] = bi;
        }
        if (lane == bl) ++head;
    }
}

struct TopkWs {
    float *qn, *dn, *S, *run_s;
    int *run_i, *run_cnt;
    size_t bytes;
};
static TopkWs carve_topk(void* ws, int nq, int64_t nd, int k) {
    Arena a(ws, (size_t)-1);
    TopkWs w;
    w.qn = a.take<float>(nq);
    w.dn = a.take<float>((size_t)nd);
    const int chunk = nd < TK_CHUNK ? (int)nd : TK_CHUNK;
    w.S = a.take<float>((size_t)nq * chunk);
    w.run_s = a.take<float>((size_t)nq * k);
    w.run_i = a.take<int>((size_t)nq * k);
    w.run_cnt = a.take<int>(nq);
    w.bytes = a.off;
    return w;
}

// host helpers shared with the tensor-core path (topk_tc.cu)
// two [k] (score, id) lists per warp: 64 KB per block at the largest k (1024)
static cudaError_t chunk_select_attr() {
    static PerDeviceOnce once;
    if (!once.need()) return cudaSuccess;
    return cudaFuncSetAttribute(chunk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4 * 1024 * (int)sizeof(float));
}

int topk_row_norms(const float* X, int64_t n, int d, float* out, cudaStream_t st) {
    row_norm_seq_kernel<<<cdiv(n, 128), 128, 0, st>>>(X, n, d, out);
    LAUNCH_CHECK("row_norm_seq");
    return DSSM_OK;
}
int topk_exact_chunk(const float* Q, int nq, const float* docs, int64_t doc0, int cd, int d, const float* qn, const float* dn, float* S,
                     int ldS, int id0, int k, float* run_s, int* run_i, int* run_cnt, cudaStream_t st) {
    dim3 grid(cdiv(cd, TK_DT), cdiv(nq, TK_QT));
    exact_scores_kernel<<<grid, 256, 0, st>>>(Q, nq, docs, doc0, cd, d, qn, dn, S, ldS);
    LAUNCH_CHECK("exact_scores");
    CUDA_TRY(chunk_select_attr());
    chunk_select_kernel<<<cdiv(nq, 4), 128, (size_t)4 * 4 * k * sizeof(float), st>>>(S, ldS, nq, cd, id0, k, run_s, run_i, run_cnt);
    LAUNCH_CHECK("chunk_select");
    return DSSM_OK;
}

}  // namespace dssm

using namespace dssm;

extern "C" size_t dssm_corpus_topk_workspace_bytes(int32_t nq, int64_t nd, int32_t d, int32_t k) {
    (void)d;
    if (nq <= 0 || nd <= 0 || k <= 0) return 0;
    return carve_topk(nullptr, nq, nd, k).bytes;
}

extern "C" int dssm_corpus_topk(const float* Q, int32_t nq, const float* docs, int64_t nd, int32_t d, int32_t k,
                                int32_t id_offset, float* out_scores, int32_t* out_ids, void* workspace,
                                size_t workspace_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(Q && docs && out_scores && out_ids && workspace, DSSM_ERR_BAD_ARG, "dssm_corpus_topk: null pointer");
    DSSM_REQUIRE(nq > 0 && nd > 0 && d > 0 && k > 0, DSSM_ERR_BAD_SHAPE, "dssm_corpus_topk: bad shape");
    DSSM_REQUIRE(k <= nd, DSSM_ERR_BAD_SHAPE, "dssm_corpus_topk: k=%d exceeds corpus size %lld", k, (long long)nd);
    DSSM_REQUIRE(k <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_corpus_topk: k=%d > 1024 not supported", k);
    DSSM_REQUIRE(nd + (int64_t)id_offset < (int64_t)1 << 31, DSSM_ERR_BAD_SHAPE, "dssm_corpus_topk: ids overflow int32");
    DSSM_REQUIRE(workspace_bytes >= carve_topk(nullptr, nq, nd, k).bytes, DSSM_ERR_WORKSPACE, "dssm_corpus_topk: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    TopkWs w = carve_topk(workspace, nq, nd, k);
    row_norm_seq_kernel<<<cdiv(nq, 128), 128, 0, st>>>(Q, nq, d, w.qn);
    LAUNCH_CHECK("row_norm_seq(Q)");
    row_norm_seq_kernel<<<cdiv(nd, 128), 128, 0, st>>>(docs, nd, d, w.dn);
    LAUNCH_CHECK("row_norm_seq(docs)");
    CUDA_TRY(cudaMemsetAsync(w.run_cnt, 0, (size_t)nq * sizeof(int), st));
    const int chunk = nd < TK_CHUNK ? (int)nd : TK_CHUNK;
    const size_t sel_smem = (size_t)4 * 4 * k * sizeof(float);
    CUDA_TRY(chunk_select_attr());
    for (int64_t d0 = 0; d0 < nd; d0 += chunk) {
        const int cd = (int)((nd - d0) < chunk ? (nd - d0) : chunk);
        dim3 grid(cdiv(cd, TK_DT), cdiv(nq, TK_QT));
        exact_scores_kernel<<<grid, 256, 0, st>>>(Q, nq, docs, d0, cd, d, w.qn, w.dn, w.S, chunk);
        LAUNCH_CHECK("exact_scores");
        chunk_select_kernel<<<cdiv(nq, 4), 128, sel_smem, st>>>(w.S, chunk, nq, cd, id_offset + (int)d0, k, w.run_s,
                                                               w.run_i, w.run_cnt);
        LAUNCH_CHECK("chunk_select");
    }
    CUDA_TRY(cudaMemcpyAsync(out_scores, w.run_s, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(out_ids, w.run_i, (size_t)nq * k * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return DSSM_OK;
}

extern "C" int dssm_topk_merge(const float* part_scores, const int32_t* part_ids, int32_t n_parts, int32_t nq, int32_t k,
                               float* out_scores, int32_t* out_ids, dssm_stream_t stream) {
    DSSM_REQUIRE(part_scores && part_ids && out_scores && out_ids, DSSM_ERR_BAD_ARG, "dssm_topk_merge: null pointer");
    DSSM_REQUIRE(n_parts > 0 && n_parts <= 32 && nq > 0 && k > 0, DSSM_ERR_BAD_SHAPE, "dssm_topk_merge: need 1 <= n_parts <= 32");
    topk_merge_kernel<<<cdiv(nq, 4), 128, 0, (cudaStream_t)stream>>>(part_scores, part_ids, n_parts, nq, k, out_scores, out_ids);
    LAUNCH_CHECK("topk_merge");
    return DSSM_OK;
}
