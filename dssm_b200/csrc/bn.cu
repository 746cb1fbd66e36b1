// batch_normalization (new_dssm.py:62-88) for the two instances of one layer (query rows [0,B), doc rows
// [B,R)), its EMA update, the fused normalise+activation, and the backward through act(BN(x)).
//
// Moments are computed per row chunk from shifted data (pivot = the chunk's first row) and the chunk triples
// (n, mean, M2) are merged by a fixed shuffle tree with Chan's formula -- no E[x^2]-E[x]^2 cancellation, and the
// result is deterministic.
#include "common.cuh"

namespace dssm {

constexpr int BN_CHUNK_ROWS = 256;  // minimum rows per partial (workspace sizing uses this)
constexpr int BN_MAX_CHUNKS = 32;   // per segment: keeps the serial merge in the finalize kernels short
constexpr int BN_TX = 32;           // columns per block
constexpr int BN_TY = 8;            // row lanes per block
constexpr int MLP = 8;              // independent row loads in flight per thread in the column reductions

// chunk table: chunks never straddle the segment boundary
__host__ __device__ inline int bn_chunks_of(int rows, int chunk_rows = BN_CHUNK_ROWS) { return (rows + chunk_rows - 1) / chunk_rows; }
// rows per chunk for a batch: >= BN_CHUNK_ROWS and at most BN_MAX_CHUNKS chunks in the larger segment
static inline int bn_chunk_rows(int R, int B) {
    const int seg = B > R - B ? B : R - B;
    int cr = (seg + BN_MAX_CHUNKS - 1) / BN_MAX_CHUNKS;
    cr = (cr + 7) / 8 * 8;
    return cr < BN_CHUNK_ROWS ? BN_CHUNK_ROWS : cr;
}

// partial layout: [3][n_chunks_total][L]  (count as float, mean, M2)
__global__ void __launch_bounds__(BN_TX * BN_TY)
bn_stats_kernel(const float* __restrict__ X, int R, int L, int B, float* __restrict__ part, int n_chunks_total, int chunk_rows) {
    __shared__ float red[BN_TY][BN_TX + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int col = blockIdx.x * BN_TX + tx;
    const int nq = bn_chunks_of(B, chunk_rows);
    int chunk = blockIdx.y, r0, r1;
    if (chunk < nq) {
        r0 = chunk * chunk_rows;
        r1 = min(B, r0 + chunk_rows);
    } else {
        r0 = B + (chunk - nq) * chunk_rows;
        r1 = min(R, r0 + chunk_rows);
    }
    const int n = r1 - r0;
    // single pass over the chunk with shifted data: d = x - K, K = the chunk's first row (a sample, so |mean-K| is
    // O(sigma) and  M2 = sum d^2 - (sum d)^2/n  loses at most a digit -- unlike E[x^2]-E[x]^2 around zero)
    float s = 0.f, q = 0.f;
    float K = 0.f;
    if (col < L) {
        K = __ldg(X + (size_t)r0 * L + col);
        int r = r0 + ty;
        for (; r + (MLP - 1) * BN_TY < r1; r += MLP * BN_TY) {  // MLP independent loads in flight per thread
            float v[MLP];
#pragma unroll
            for (int u = 0; u < MLP; ++u) v[u] = __ldg(X + (size_t)(r + u * BN_TY) * L + col);
#pragma unroll
            for (int u = 0; u < MLP; ++u) {
                const float d = v[u] - K;
                s += d;
                q = fmaf(d, d, q);
            }
        }
#pragma unroll 1
        for (; r < r1; r += BN_TY) {
            const float d = __ldg(X + (size_t)r * L + col) - K;
            s += d;
            q = fmaf(d, d, q);
        }
    }
    __shared__ float red2[BN_TY][BN_TX + 1];
    red[ty][tx] = s;
    red2[ty][tx] = q;
    __syncthreads();
    if (ty == 0 && col < L) {
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int i = 0; i < BN_TY; ++i) { ts += red[i][tx]; tq += red2[i][tx]; }
        const float md = ts / (float)n;
        part[((size_t)0 * n_chunks_total + chunk) * L + col] = (float)n;
        part[((size_t)1 * n_chunks_total + chunk) * L + col] = K + md;
        part[((size_t)2 * n_chunks_total + chunk) * L + col] = fmaxf(tq - ts * md, 0.f);
    }
}

// one warp per (segment, column): lane c holds chunk c's (n, mean, M2) -- at most BN_MAX_CHUNKS = 32 chunks per
// segment by construction -- merged by a fixed shuffle tree with Chan's formula; lane 0 writes EMA, scale/shift
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float* __restrict__ part, int n_chunks_total, int nq_chunks, int L, int on_train, int update_ema,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ ema_mean,
                   float* __restrict__ ema_var, float eps, float decay, float* __restrict__ mean_o, float* __restrict__ var_o,
                   float* __restrict__ rstd_o, float* __restrict__ scale_o, float* __restrict__ shift_o) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= 2 * L) return;
    const int seg = i / L, col = i - seg * L;
    float mean, var;
    if (on_train) {
        const int c0 = seg == 0 ? 0 : nq_chunks;
        const int c1 = seg == 0 ? nq_chunks : n_chunks_total;
        if (c0 == c1) return;  // empty segment (single-instance call, B == R)
        float n = 0.f, mu = 0.f, m2 = 0.f;
        if (c0 + lane < c1) {
            const int c = c0 + lane;
            n = part[((size_t)0 * n_chunks_total + c) * L + col];
            mu = part[((size_t)1 * n_chunks_total + c) * L + col];
            m2 = part[((size_t)2 * n_chunks_total + c) * L + col];
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {  // lane l absorbs lane l+o: chunks stay merged in ascending order
            const float nb = __shfl_down_sync(0xffffffffu, n, o);
            const float mb = __shfl_down_sync(0xffffffffu, mu, o);
            const float qb = __shfl_down_sync(0xffffffffu, m2, o);
            const float nt = n + nb;
            if (nb > 0.f && (lane & (2 * o - 1)) == 0) {
                const float delta = mb - mu;
                mu = mu + delta * (nb / nt);
                m2 = m2 + qb + delta * delta * (n * nb / nt);
                n = nt;
            }
        }
        if (lane != 0) return;
        mean = mu;
        var = m2 / n;  // biased (tf.nn.moments)
        if (update_ema) {
            // ExponentialMovingAverage.apply: shadow -= (1 - decay) * (shadow - value), new_dssm.py:78-81
            const float em = ema_mean[i], ev = ema_var[i];
            ema_mean[i] = em - (1.f - decay) * (em - mean);
            ema_var[i] = ev - (1.f - decay) * (ev - var);
        }
    } else {
        if (lane != 0) return;
        mean = ema_mean[i];
        var = ema_var[i];
    }
    const float rstd = 1.0f / sqrtf(var + eps);
    const float sc = rstd * gamma[i];
    mean_o[i] = mean;
    var_o[i] = var;
    rstd_o[i] = rstd;
    scale_o[i] = sc;
    shift_o[i] = beta[i] - mean * sc;
}

constexpr int EW_ROWS = 4;  // rows per block iteration of the elementwise kernels

__global__ void __launch_bounds__(256)
bn_act_apply_kernel(const float* __restrict__ X, int R, int L, int B, const float* __restrict__ scale,
                    const float* __restrict__ shift, int act, float* __restrict__ Y) {
    for (int r0 = blockIdx.x * EW_ROWS; r0 < R; r0 += gridDim.x * EW_ROWS) {
        for (int c = threadIdx.x; c < L; c += blockDim.x) {
#pragma unroll
            for (int j = 0; j < EW_ROWS; ++j) {
                const int r = r0 + j;
                if (r < R) {
                    float x = __ldg(X + (size_t)r * L + c);
                    if (scale) {
                        const int o = (r < B ? 0 : L) + c;
                        x = fmaf(x, __ldg(scale + o), __ldg(shift + o));
                    }
                    Y[(size_t)r * L + c] = act_fwd(x, act);
                }
            }
        }
    }
}

// ---- backward -------------------------------------------------------------------------------------
// pass 1: per chunk column sums of g = dA*act'(a) and g*xhat  -> part [2][n_chunks_total][L]
__global__ void __launch_bounds__(BN_TX * BN_TY)
bn_bwd_reduce_kernel(const float* __restrict__ dA, const float* __restrict__ H, int R, int L, int B, int act,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ scale,
                     const float* __restrict__ shift, float* __restrict__ part, int n_chunks_total, int chunk_rows) {
    __shared__ float red0[BN_TY][BN_TX + 1];
    __shared__ float red1[BN_TY][BN_TX + 1];
    __shared__ float red2[BN_TY][BN_TX + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int col = blockIdx.x * BN_TX + tx;
    const int nq = bn_chunks_of(B, chunk_rows);
    int chunk = blockIdx.y, r0, r1, seg;
    if (chunk < nq) {
        seg = 0; r0 = chunk * chunk_rows; r1 = min(B, r0 + chunk_rows);
    } else {
        seg = 1; r0 = B + (chunk - nq) * chunk_rows; r1 = min(R, r0 + chunk_rows);
    }
    float sg = 0.f, sgx = 0.f, sx = 0.f;
    if (col < L) {
        const int o = seg * L + col;
        const float mu = __ldg(mean + o), rs = __ldg(rstd + o), sc = __ldg(scale + o), sh = __ldg(shift + o);
        int r = r0 + ty;
        for (; r + (MLP - 1) * BN_TY < r1; r += MLP * BN_TY) {
            float hv[MLP], dv[MLP];
#pragma unroll
            for (int u = 0; u < MLP; ++u) {
                hv[u] = __ldg(H + (size_t)(r + u * BN_TY) * L + col);
                dv[u] = __ldg(dA + (size_t)(r + u * BN_TY) * L + col);
            }
#pragma unroll
            for (int u = 0; u < MLP; ++u) {
                const float a = act_fwd(fmaf(hv[u], sc, sh), act);
                const float g = dv[u] * act_grad_from_out(a, act);
                const float xh = (hv[u] - mu) * rs;
                sg += g;
                sgx = fmaf(g, xh, sgx);
                sx += xh;
            }
        }
        for (; r < r1; r += BN_TY) {
            const float h = __ldg(H + (size_t)r * L + col);
            const float a = act_fwd(fmaf(h, sc, sh), act);
            const float g = __ldg(dA + (size_t)r * L + col) * act_grad_from_out(a, act);
            const float xh = (h - mu) * rs;
            sg += g;
            sgx = fmaf(g, xh, sgx);
            sx += xh;
        }
    }
    red0[ty][tx] = sg;
    red1[ty][tx] = sgx;
    red2[ty][tx] = sx;
    __syncthreads();
    if (ty == 0 && col < L) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int i = 0; i < BN_TY; ++i) { t0 += red0[i][tx]; t1 += red1[i][tx]; t2 += red2[i][tx]; }
        part[((size_t)0 * n_chunks_total + chunk) * L + col] = t0;
        part[((size_t)1 * n_chunks_total + chunk) * L + col] = t1;
        part[((size_t)2 * n_chunks_total + chunk) * L + col] = t2;
    }
}

// dgamma, dbeta per (segment, column); optionally the pre-BN bias gradient db[c] = sum_r dH[r,c].  Under BN that sum
// is identically  -gamma*rstd*dgamma*(sum_r xhat)/n  per segment (the g and dbeta terms cancel exactly), i.e. rounding
// noise around zero -- evaluated here from sum xhat accumulated in the same pass instead of another sweep over dH.
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ part, int n_chunks_total, int nq_chunks, int L, int B, int R,
                       const float* __restrict__ gamma, const float* __restrict__ rstd, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ db_seg /* [2][L] or NULL */) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= 2 * L) return;
    const int seg = i / L, col = i - seg * L;
    const int c0 = seg == 0 ? 0 : nq_chunks;
    const int c1 = seg == 0 ? nq_chunks : n_chunks_total;
    float b = 0.f, g = 0.f, x = 0.f;
    for (int c = c0 + lane; c < c1; c += 32) {
        b += part[((size_t)0 * n_chunks_total + c) * L + col];
        g += part[((size_t)1 * n_chunks_total + c) * L + col];
        x += part[((size_t)2 * n_chunks_total + c) * L + col];
    }
    b = warp_sum(b);
    g = warp_sum(g);
    x = warp_sum(x);
    if (lane == 0) {
        dbeta[i] = b;
        dgamma[i] = g;
        if (db_seg) {
            const float n = seg == 0 ? (float)B : (float)(R - B);
            db_seg[i] = n > 0.f ? -(gamma[i] * rstd[i]) * g * (x / n) : 0.f;
        }
    }
}

__global__ void db_combine_kernel(const float* __restrict__ db_seg, int L, float* __restrict__ db) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < L) db[c] = db_seg[c] + db_seg[L + c];
}

// pass 2 (in place): dH = gamma*rstd * (g - dbeta/n - xhat*dgamma/n)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(float* __restrict__ dA, const float* __restrict__ H, int R, int L, int B, int act,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ dgamma,
                    const float* __restrict__ dbeta) {
    for (int r0 = blockIdx.x * EW_ROWS; r0 < R; r0 += gridDim.x * EW_ROWS) {
        for (int c = threadIdx.x; c < L; c += blockDim.x) {
            float h[EW_ROWS], d[EW_ROWS];
#pragma unroll
            for (int j = 0; j < EW_ROWS; ++j) {
                const int r = r0 + j;
                h[j] = r < R ? __ldg(H + (size_t)r * L + c) : 0.f;
                d[j] = r < R ? dA[(size_t)r * L + c] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < EW_ROWS; ++j) {
                const int r = r0 + j;
                if (r >= R) continue;
                const int seg = r < B ? 0 : 1;
                const int o = seg * L + c;
                const float n = seg == 0 ? (float)B : (float)(R - B);
                const float a = act_fwd(fmaf(h[j], __ldg(scale + o), __ldg(shift + o)), act);
                const float g = d[j] * act_grad_from_out(a, act);
                const float rs = __ldg(rstd + o);
                const float xhat = (h[j] - __ldg(mean + o)) * rs;
                dA[(size_t)r * L + c] = (__ldg(gamma + o) * rs) * (g - __ldg(dbeta + o) / n - xhat * (__ldg(dgamma + o) / n));
            }
        }
    }
}

// float4 variant (L % 4 == 0): one thread per 4 consecutive columns, EW4_ROWS rows in flight
constexpr int EW4_ROWS = 8;
__global__ void __launch_bounds__(128)
bn_bwd_apply_v4_kernel(float4* __restrict__ dA, const float4* __restrict__ H, int R, int L4, int B, int act,
                       const float4* __restrict__ gamma, const float4* __restrict__ mean, const float4* __restrict__ rstd,
                       const float4* __restrict__ scale, const float4* __restrict__ shift, const float4* __restrict__ dgamma,
                       const float4* __restrict__ dbeta) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= L4) return;
    const int r0 = blockIdx.x * EW4_ROWS;
    float4 h[EW4_ROWS], d[EW4_ROWS];
#pragma unroll
    for (int j = 0; j < EW4_ROWS; ++j) {
        const int r = r0 + j;
        if (r < R) {
            h[j] = __ldg(H + (size_t)r * L4 + c);
            d[j] = dA[(size_t)r * L4 + c];
        }
    }
#pragma unroll
    for (int j = 0; j < EW4_ROWS; ++j) {
        const int r = r0 + j;
        if (r >= R) continue;
        const int seg = r < B ? 0 : 1;
        const int o = seg * L4 + c;
        const float inv_n = 1.0f / (seg == 0 ? (float)B : (float)(R - B));
        const float4 sc = __ldg(scale + o), sh = __ldg(shift + o), rs = __ldg(rstd + o), mu = __ldg(mean + o), ga = __ldg(gamma + o),
                     dg = __ldg(dgamma + o), db = __ldg(dbeta + o);
        float4 out;
#define BN_BWD1(x)                                                                  \
    {                                                                               \
        const float a = act_fwd(fmaf(h[j].x, sc.x, sh.x), act);                     \
        const float g = d[j].x * act_grad_from_out(a, act);                         \
        const float xhat = (h[j].x - mu.x) * rs.x;                                  \
        out.x = (ga.x * rs.x) * (g - db.x * inv_n - xhat * (dg.x * inv_n));        \
    }
        BN_BWD1(x) BN_BWD1(y) BN_BWD1(z) BN_BWD1(w)
#undef BN_BWD1
        dA[(size_t)r * L4 + c] = out;
    }
}

// no-BN mode: dH = dA * act'(act(H))
__global__ void act_bwd_kernel(float* __restrict__ dA, const float* __restrict__ H, size_t total, int act) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float a = act_fwd(__ldg(H + i), act);
        dA[i] = dA[i] * act_grad_from_out(a, act);
    }
}

static int row_blocks(int R) {
    int b = (R + EW_ROWS - 1) / EW_ROWS;
    const int cap = sm_count() * 16;
    return b < cap ? (b ? b : 1) : cap;
}

static int ew_blocks(size_t total) {
    size_t b = (total + 255) / 256;
    const size_t cap = (size_t)sm_count() * 16;
    return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace dssm

using namespace dssm;

extern "C" size_t dssm_bn_workspace_bytes(int32_t R, int32_t L) {
    if (R <= 0 || L <= 0) return 0;
    // worst case split of R rows into two segments adds one chunk
    const size_t chunks = (size_t)bn_chunks_of(R) + 2;
    return align_up((3 * chunks + 2) * (size_t)L * sizeof(float), 256);
}

extern "C" int dssm_bn_forward(const float* X, int32_t R, int32_t L, int32_t B, int32_t on_train, int32_t update_ema,
                               const float* gamma, const float* beta, float* ema_mean, float* ema_var, float eps,
                               float ema_decay, float* mean, float* var, float* rstd, float* scale, float* shift,
                               void* workspace, size_t workspace_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(X && gamma && beta && ema_mean && ema_var && mean && var && rstd && scale && shift, DSSM_ERR_BAD_ARG,
                 "dssm_bn_forward: null pointer");
    DSSM_REQUIRE(R > 0 && L > 0 && B > 0 && B <= R, DSSM_ERR_BAD_SHAPE, "dssm_bn_forward: need 0 < B <= R (R=%d B=%d)", R, B);
    cudaStream_t st = (cudaStream_t)stream;
    const int cr = bn_chunk_rows(R, B);
    const int nq = bn_chunks_of(B, cr), nd = bn_chunks_of(R - B, cr), nt = nq + nd;
    float* part = (float*)workspace;
    if (on_train) {
        DSSM_REQUIRE(workspace && workspace_bytes >= (size_t)3 * nt * L * sizeof(float), DSSM_ERR_WORKSPACE,
                     "dssm_bn_forward: workspace too small");
        dim3 grid(cdiv(L, BN_TX), nt), block(BN_TX, BN_TY);
        bn_stats_kernel<<<grid, block, 0, st>>>(X, R, L, B, part, nt, cr);
        LAUNCH_CHECK("bn_stats");
    }
    bn_finalize_kernel<<<cdiv(2 * L, 8), 256, 0, st>>>(part, nt, nq, L, on_train, update_ema, gamma, beta, ema_mean,
                                                         ema_var, eps, ema_decay, mean, var, rstd, scale, shift);
    LAUNCH_CHECK("bn_finalize");
    return DSSM_OK;
}

extern "C" int dssm_bn_act_apply(const float* X, int32_t R, int32_t L, int32_t B, const float* scale,
                                 const float* shift, int32_t act, float* Y, dssm_stream_t stream) {
    DSSM_REQUIRE(X && Y, DSSM_ERR_BAD_ARG, "dssm_bn_act_apply: null pointer");
    DSSM_REQUIRE((scale == nullptr) == (shift == nullptr), DSSM_ERR_BAD_ARG, "dssm_bn_act_apply: scale/shift must both be set or both NULL");
    DSSM_REQUIRE(R >= 0 && L > 0, DSSM_ERR_BAD_SHAPE, "dssm_bn_act_apply: bad shape");
    if (R == 0) return DSSM_OK;
    bn_act_apply_kernel<<<row_blocks(R), 256, 0, (cudaStream_t)stream>>>(X, R, L, B, scale, shift, act, Y);
    LAUNCH_CHECK("bn_act_apply");
    return DSSM_OK;
}

extern "C" int dssm_bn_act_backward(float* dA, const float* H, int32_t R, int32_t L, int32_t B, int32_t act,
                                    const float* gamma, const float* mean, const float* rstd, const float* scale,
                                    const float* shift, float* dgamma, float* dbeta, float* db, void* workspace,
                                    size_t workspace_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(dA && H, DSSM_ERR_BAD_ARG, "dssm_bn_act_backward: null pointer");
    DSSM_REQUIRE(R > 0 && L > 0, DSSM_ERR_BAD_SHAPE, "dssm_bn_act_backward: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (!scale) {
        act_bwd_kernel<<<ew_blocks((size_t)R * L), 256, 0, st>>>(dA, H, (size_t)R * L, act);
        LAUNCH_CHECK("act_bwd");
        return DSSM_OK;
    }
    DSSM_REQUIRE(gamma && mean && rstd && shift && dgamma && dbeta, DSSM_ERR_BAD_ARG, "dssm_bn_act_backward: null BN pointer");
    DSSM_REQUIRE(B > 0 && B <= R, DSSM_ERR_BAD_SHAPE, "dssm_bn_act_backward: need 0 < B <= R");
    const int cr = bn_chunk_rows(R, B);
    const int nq = bn_chunks_of(B, cr), nd = bn_chunks_of(R - B, cr), nt = nq + nd;
    DSSM_REQUIRE(workspace && workspace_bytes >= ((size_t)3 * nt + 2) * L * sizeof(float), DSSM_ERR_WORKSPACE,
                 "dssm_bn_act_backward: workspace too small");
    float* part = (float*)workspace;
    float* db_seg = db ? part + (size_t)3 * nt * L : nullptr;
    dim3 grid(cdiv(L, BN_TX), nt), block(BN_TX, BN_TY);
    bn_bwd_reduce_kernel<<<grid, block, 0, st>>>(dA, H, R, L, B, act, mean, rstd, scale, shift, part, nt, cr);
    LAUNCH_CHECK("bn_bwd_reduce");
    bn_bwd_finalize_kernel<<<cdiv(2 * L, 8), 256, 0, st>>>(part, nt, nq, L, B, R, gamma, rstd, dgamma, dbeta, db_seg);
    LAUNCH_CHECK("bn_bwd_finalize");
    if (db) {
        db_combine_kernel<<<cdiv(L, 128), 128, 0, st>>>(db_seg, L, db);
        LAUNCH_CHECK("db_combine");
    }
    if (L % 4 == 0 && aligned16(dA) && aligned16(H) && aligned16(gamma) && aligned16(mean) && aligned16(rstd) && aligned16(scale) &&
        aligned16(shift) && aligned16(dgamma) && aligned16(dbeta)) {
        dim3 grid(cdiv(R, EW4_ROWS), cdiv(L / 4, 128));
        bn_bwd_apply_v4_kernel<<<grid, 128, 0, st>>>((float4*)dA, (const float4*)H, R, L / 4, B, act, (const float4*)gamma,
                                                    (const float4*)mean, (const float4*)rstd, (const float4*)scale,
                                                    (const float4*)shift, (const float4*)dgamma, (const float4*)dbeta);
    } else {
        bn_bwd_apply_kernel<<<row_blocks(R), 256, 0, st>>>(dA, H, R, L, B, act, gamma, mean, rstd, scale, shift, dgamma, dbeta);
    }
    LAUNCH_CHECK("bn_bwd_apply");
    return DSSM_OK;
}
