// FC2.. dense layers (tf.matmul(x_out, weight2) + bias2, new_dssm.py:146-148) and their gradients, fp32 path.
//
// One register-tiled FFMA kernel serves the three contractions; the previous layer's BN + activation
// (new_dssm.py:87,134-136) is applied while the A operand is staged, so the normalised tensor never
// exists in HBM:
//   NN  Hout[R,N] = pro(Hprev)[R,K] . W[K,N] + bias                 (forward)
//   NT  dA[R,K]   = dH[R,N] . W[K,N]^T                              (gradient w.r.t. the layer input)
//   TN  dW[K,N]   = pro(Hprev)[R,K]^T . dH[R,N]   split over R      (gradient w.r.t. the weights)
// This is the 1e-5 parity mode (DSSM_GEMM_FP32); the tcgen05 bf16 path lives in fc_tc.cu.
#include "common.cuh"

namespace dssm {

constexpr int BM = 128, BN = 64, BK = 16, GEMM_THREADS = 256;
constexpr int TM = 8, TN_ = 4;
enum { MODE_NN = 0, MODE_NT = 1, MODE_TN = 2 };

struct GemmArgs {
    const float* A;
    const float* Bm;
    float* C;
    const float* bias;
    int M, N, K;  // C is [M,N]; K is the reduced dimension
    const float* scale;
    const float* shift;  // [2][feat] or NULL
    int act;             // applied to A in NN/TN (even when scale is NULL)
    int Bseg;            // rows < Bseg use segment 0
    int k_per_split;     // TN only
};

template <int MODE>
__device__ __forceinline__ float load_a(const GemmArgs& g, int m, int k) {
    if (MODE == MODE_NN) {
        if (m >= g.M || k >= g.K) return 0.f;
        float x = __ldg(g.A + (size_t)m * g.K + k);
        if (g.scale) {
            const int o = (m < g.Bseg ? 0 : g.K) + k;
            x = fmaf(x, __ldg(g.scale + o), __ldg(g.shift + o));
        }
        return act_fwd(x, g.act);
    } else if (MODE == MODE_NT) {
        if (m >= g.M || k >= g.K) return 0.f;
        return __ldg(g.A + (size_t)m * g.K + k);
    } else {  // TN: A = Hprev [K(rows), M(features)]
        if (m >= g.M || k >= g.K) return 0.f;
        float x = __ldg(g.A + (size_t)k * g.M + m);
        if (g.scale) {
            const int o = (k < g.Bseg ? 0 : g.M) + m;
            x = fmaf(x, __ldg(g.scale + o), __ldg(g.shift + o));
        }
        return act_fwd(x, g.act);
    }
}

template <int MODE>
__device__ __forceinline__ float load_b(const GemmArgs& g, int k, int n) {
    if (k >= g.K || n >= g.N) return 0.f;
    if (MODE == MODE_NT) return __ldg(g.Bm + (size_t)n * g.K + k);
    return __ldg(g.Bm + (size_t)k * g.N + n);
}

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_f32_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    int kbeg = 0, kend = g.K;
    if (MODE == MODE_TN) {
        kbeg = blockIdx.z * g.k_per_split;
        kend = min(g.K, kbeg + g.k_per_split);
    }
    // load mapping: the operand's contiguous index runs fastest over tid
    constexpr bool A_KFAST = (MODE != MODE_TN);
    constexpr bool B_KFAST = (MODE == MODE_NT);
    float ra[8], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * GEMM_THREADS;  // 0..2047
            const int kk = A_KFAST ? (e % BK) : (e / BM);
            const int mm = A_KFAST ? (e / BK) : (e % BM);
            const int k = k0 + kk;
            ra[i] = (k < kend) ? load_a<MODE>(g, m0 + mm, k) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * GEMM_THREADS;  // 0..1023
            const int kk = B_KFAST ? (e % BK) : (e / BN);
            const int nn = B_KFAST ? (e / BK) : (e % BN);
            const int k = k0 + kk;
            rb[i] = (k < kend) ? load_b<MODE>(g, k, n0 + nn) : 0.f;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * GEMM_THREADS;
            const int kk = A_KFAST ? (e % BK) : (e / BM);
            const int mm = A_KFAST ? (e / BK) : (e % BM);
            As[kk][mm] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * GEMM_THREADS;
            const int kk = B_KFAST ? (e % BK) : (e / BN);
            const int nn = B_KFAST ? (e / BK) : (e % BN);
            Bs[kk][nn] = rb[i];
        }
    };
    const int ty = tid / 16, tx = tid % 16;
    float acc[TM][TN_];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN_; ++j) acc[i][j] = 0.f;

    fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        stash();
        __syncthreads();
        if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN_]);
            const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[TN_] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN_; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* C = g.C;
    if (MODE == MODE_TN) C += (size_t)blockIdx.z * g.M * g.N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN_; ++j) {
            const int n = n0 + tx * TN_ + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (MODE == MODE_NN && g.bias) v += __ldg(g.bias + n);
            C[(size_t)m * g.N + n] = v;
        }
    }
}

// out[i] = sum_s part[s][i] in split order
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, size_t n, float* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += part[(size_t)k * n + i];
        out[i] = s;
    }
}

// ---- column sums (bias gradients) -------------------------------------------------------------------
constexpr int CS_ROWS = 256;
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ X, int R, int N, float* __restrict__ part) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx;
    const int r0 = blockIdx.y * CS_ROWS, r1 = min(R, r0 + CS_ROWS);
    float s = 0.f;
    if (col < N) {
        int r = r0 + ty;
        for (; r + 7 * 8 < r1; r += 64) {  // 8 independent loads in flight per thread
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(X + (size_t)(r + u * 8) * N + col);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; r < r1; r += 8) s += __ldg(X + (size_t)r * N + col);
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && col < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][tx];
        part[(size_t)blockIdx.y * N + col] = t;
    }
}

static int pick_splits(int R) {
    // enough CTAs to fill the chip on a [K,N] ~ 300x300 output (15 tiles), at least 128 rows per split
    int s = cdiv(R, 256);
    if (s > 64) s = 64;
    if (s < 1) s = 1;
    return s;
}

}  // namespace dssm

using namespace dssm;

// tcgen05 path (fc_tc.cu)
extern "C" size_t dssm_fc_tc_workspace_bytes(int32_t K, int32_t N);
extern "C" int dssm_fc_fwd_tc(const float*, int32_t, int32_t, int32_t, const float*, const float*, int32_t,
                              const float*, const float*, int32_t, float*, void*, size_t, int32_t, dssm_stream_t);
extern "C" int dssm_fc_bwd_dx_tc(const float*, int32_t, int32_t, const float*, int32_t, float*, void*, size_t, int32_t, dssm_stream_t);
extern "C" size_t dssm_fc_bwd_dw_tc_workspace_bytes(int32_t R, int32_t K, int32_t N);
extern "C" int dssm_fc_bwd_dw_tc(const float*, int32_t, int32_t, int32_t, const float*, const float*, int32_t, const float*,
                                 int32_t, float*, int32_t*, int32_t, dssm_stream_t);
static inline bool is_tc(int mode) { return mode == DSSM_GEMM_TC_3XTF32 || mode == DSSM_GEMM_TC_TF32; }
static inline int tc_passes(int mode) { return mode == DSSM_GEMM_TC_TF32 ? 1 : 3; }

extern "C" size_t dssm_fc_fwd_workspace_bytes(int32_t K, int32_t N, int32_t gemm_mode) {
    return is_tc(gemm_mode) ? dssm_fc_tc_workspace_bytes(K, N) : 0;
}

extern "C" int dssm_fc_fwd(const float* Hprev, int32_t R, int32_t K, int32_t B, const float* scale, const float* shift,
                           int32_t act, const float* W, const float* bias, int32_t N, float* Hout, int32_t gemm_mode,
                           void* workspace, size_t workspace_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(Hprev && W && Hout, DSSM_ERR_BAD_ARG, "dssm_fc_fwd: null pointer");
    DSSM_REQUIRE((scale == nullptr) == (shift == nullptr), DSSM_ERR_BAD_ARG, "dssm_fc_fwd: scale/shift must both be set or both NULL");
    DSSM_REQUIRE(R > 0 && K > 0 && N > 0, DSSM_ERR_BAD_SHAPE, "dssm_fc_fwd: bad shape R=%d K=%d N=%d", R, K, N);
    if (is_tc(gemm_mode))
        return dssm_fc_fwd_tc(Hprev, R, K, B, scale, shift, act, W, bias, N, Hout, workspace, workspace_bytes, tc_passes(gemm_mode), stream);
    DSSM_REQUIRE(gemm_mode == DSSM_GEMM_FP32, DSSM_ERR_BAD_ARG, "dssm_fc_fwd: unknown gemm_mode %d", gemm_mode);
    GemmArgs g{Hprev, W, Hout, bias, R, N, K, scale, shift, act, B, 0};
    dim3 grid(cdiv(N, BN), cdiv(R, BM), 1);
    gemm_f32_kernel<MODE_NN><<<grid, GEMM_THREADS, 0, (cudaStream_t)stream>>>(g);
    LAUNCH_CHECK("fc_fwd");
    return DSSM_OK;
}

extern "C" int dssm_fc_bwd_dx(const float* dH, int32_t R, int32_t N, const float* W, int32_t K, float* dA,
                              int32_t gemm_mode, void* workspace, size_t workspace_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(dH && W && dA, DSSM_ERR_BAD_ARG, "dssm_fc_bwd_dx: null pointer");
    DSSM_REQUIRE(R > 0 && K > 0 && N > 0, DSSM_ERR_BAD_SHAPE, "dssm_fc_bwd_dx: bad shape");
    if (is_tc(gemm_mode)) return dssm_fc_bwd_dx_tc(dH, R, N, W, K, dA, workspace, workspace_bytes, tc_passes(gemm_mode), stream);
    // C[R,K] = dH[R,N] . W[K,N]^T : reduce over N
    GemmArgs g{dH, W, dA, nullptr, R, K, N, nullptr, nullptr, DSSM_ACT_NONE, 0, 0};
    dim3 grid(cdiv(K, BN), cdiv(R, BM), 1);
    gemm_f32_kernel<MODE_NT><<<grid, GEMM_THREADS, 0, (cudaStream_t)stream>>>(g);
    LAUNCH_CHECK("fc_bwd_dx");
    return DSSM_OK;
}

extern "C" size_t dssm_colsum_workspace_bytes(int32_t R, int32_t N) {
    if (R <= 0 || N <= 0) return 0;
    return align_up((size_t)cdiv(R, CS_ROWS) * N * sizeof(float), 256);
}

extern "C" int dssm_colsum(const float* X, int32_t R, int32_t N, float* out, void* workspace, size_t workspace_bytes,
                           dssm_stream_t stream) {
    DSSM_REQUIRE(X && out, DSSM_ERR_BAD_ARG, "dssm_colsum: null pointer");
    DSSM_REQUIRE(R > 0 && N > 0, DSSM_ERR_BAD_SHAPE, "dssm_colsum: bad shape");
    const int chunks = cdiv(R, CS_ROWS);
    DSSM_REQUIRE(workspace && workspace_bytes >= (size_t)chunks * N * sizeof(float), DSSM_ERR_WORKSPACE, "dssm_colsum: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(cdiv(N, 32), chunks);
    colsum_partial_kernel<<<grid, 256, 0, st>>>(X, R, N, (float*)workspace);
    LAUNCH_CHECK("colsum_partial");
    splitk_reduce_kernel<<<cdiv(N, 256), 256, 0, st>>>((const float*)workspace, chunks, (size_t)N, out);
    LAUNCH_CHECK("colsum_reduce");
    return DSSM_OK;
}

extern "C" size_t dssm_fc_bwd_dw_workspace_bytes(int32_t R, int32_t K, int32_t N) {
    if (R <= 0 || K <= 0 || N <= 0) return 0;
    size_t a = align_up((size_t)pick_splits(R) * K * N * sizeof(float), 256);
    const size_t t = dssm_fc_bwd_dw_tc_workspace_bytes(R, K, N);  // sized for either gemm_mode
    if (t > a) a = t;
    return a + dssm_colsum_workspace_bytes(R, N);
}

extern "C" int dssm_fc_bwd_dw(const float* Hprev, int32_t R, int32_t K, int32_t B, const float* scale,
                              const float* shift, int32_t act, const float* dH, int32_t N, float* dW, float* db,
                              int32_t gemm_mode, void* workspace, size_t workspace_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(Hprev && dH && dW, DSSM_ERR_BAD_ARG, "dssm_fc_bwd_dw: null pointer");
    DSSM_REQUIRE((scale == nullptr) == (shift == nullptr), DSSM_ERR_BAD_ARG, "dssm_fc_bwd_dw: scale/shift must both be set or both NULL");
    DSSM_REQUIRE(R > 0 && K > 0 && N > 0, DSSM_ERR_BAD_SHAPE, "dssm_fc_bwd_dw: bad shape");
    DSSM_REQUIRE(workspace && workspace_bytes >= dssm_fc_bwd_dw_workspace_bytes(R, K, N), DSSM_ERR_WORKSPACE,
                 "dssm_fc_bwd_dw: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int splits = pick_splits(R);
    float* part = (float*)workspace;
    const size_t part_bytes = dssm_fc_bwd_dw_workspace_bytes(R, K, N) - dssm_colsum_workspace_bytes(R, N);
    if (is_tc(gemm_mode)) {
        int rc = dssm_fc_bwd_dw_tc(Hprev, R, K, B, scale, shift, act, dH, N, part, &splits, tc_passes(gemm_mode), stream);
        if (rc != DSSM_OK) return rc;
    } else {
        const int kps = cdiv(R, splits);
        // C[K,N] = pro(Hprev)^T . dH : M=K (features), reduce over R
        GemmArgs g{Hprev, dH, part, nullptr, K, N, R, scale, shift, act, B, kps};
        dim3 grid(cdiv(N, BN), cdiv(K, BM), splits);
        gemm_f32_kernel<MODE_TN><<<grid, GEMM_THREADS, 0, st>>>(g);
        LAUNCH_CHECK("fc_bwd_dw");
    }
    const size_t n = (size_t)K * N;
    int rb = cdiv((int64_t)n, 256);
    splitk_reduce_kernel<<<rb, 256, 0, st>>>(part, splits, n, dW);
    LAUNCH_CHECK("fc_bwd_dw_reduce");
    if (db) {
        char* cs = (char*)workspace + part_bytes;
        return dssm_colsum(dH, R, N, db, cs, dssm_colsum_workspace_bytes(R, N), stream);
    }
    return DSSM_OK;
}
