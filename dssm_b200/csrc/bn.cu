// batch_normalization (new_dssm.py:62-88) for the two instances of one layer (query rows [0,B), doc rows
// [B,R)), its EMA update, the fused normalise+activation, and the backward through act(BN(x)).
//
// Moments are computed per row chunk from shifted data (pivot = the chunk's first row) and the chunk triples
// (n, mean, M2) are merged by a fixed shuffle tree with Chan's formula -- no E[x^2]-E[x]^2 cancellation, and the
// result is deterministic.
#include "common.cuh"
#include "bn_common.cuh"

namespace dssm {

constexpr int BN_CHUNK_ROWS = 256;  // minimum rows per partial (workspace sizing uses this)
constexpr int BN_MAX_CHUNKS = 32;   // per segment: keeps the serial merge in the finalize kernels short
constexpr int BN_TX = 32;           // columns per block
constexpr int BN_TY = 32;           // row lanes per block, forward: a 256-row chunk is ONE round of MLP loads per thread
constexpr int BN_TY_BWD = 16;       // backward (two operands, ~64 registers): 512 threads so that two blocks share an SM
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

// "last block finalizes": every block of one 32-column strip takes a ticket after publishing its partials; the block
// that draws the last ticket merges the strip (fixed ascending chunk order, so the result does not depend on which
// block that is) and puts the ticket counter back to zero.  Saves a dependent launch per reduction.
__device__ __forceinline__ bool last_block_of_strip(int* tickets, int n_blocks) {
    __shared__ int s_last;
    // partials of this block visible device-wide before the ticket is taken.  Only thread row 0 wrote partials: the other
    // 31 rows skip the fence (ncu: the membar of all 1024 threads was 14 % of the kernel's stall samples)
    if (threadIdx.y == 0) __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const int prev = atomicAdd(tickets + blockIdx.x, 1);
        s_last = prev == n_blocks - 1;
        if (s_last) tickets[blockIdx.x] = 0;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

// The strip's partials [planes][n_chunks][32 columns] staged in shared memory by the whole block (one L2 round trip
// instead of one per chunk in a serial merge).
template <int PLANES>
__device__ __forceinline__ void load_strip_partials(const float* __restrict__ part, int n_chunks_total, int L, float* s_part) {
    const int t = threadIdx.y * BN_TX + threadIdx.x;
    const int total = PLANES * n_chunks_total * BN_TX;
    const int col0 = blockIdx.x * BN_TX;
    for (int e = t; e < total; e += BN_TX * (int)blockDim.y) {
        const int pc = e / BN_TX, cx = e - pc * BN_TX;  // pc = plane * n_chunks_total + chunk
        s_part[e] = col0 + cx < L ? __ldcg(part + (size_t)pc * L + col0 + cx) : 0.f;
    }
    __syncthreads();
}

// partial layout: [3][n_chunks_total][L]  (count as float, mean, M2)
__global__ void __launch_bounds__(BN_TX * BN_TY)
bn_stats_kernel(const float* __restrict__ X, int R, int L, int B, float* __restrict__ part, int n_chunks_total, int chunk_rows,
                int* __restrict__ tickets, BnFinalize fin) {
    __shared__ float red[BN_TY][BN_TX + 1];
    __shared__ float red2[BN_TY][BN_TX + 1];
    __shared__ float s_part[3 * 2 * BN_MAX_CHUNKS * BN_TX];
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
        {   // remainder (< MLP rows per thread), still issued together
            float v[MLP];
#pragma unroll
            for (int u = 0; u < MLP; ++u) v[u] = (r + u * BN_TY < r1) ? __ldg(X + (size_t)(r + u * BN_TY) * L + col) : K;
#pragma unroll
            for (int u = 0; u < MLP; ++u) {
                const float d = v[u] - K;  // exactly 0 for the padding
                s += d;
                q = fmaf(d, d, q);
            }
        }
    }
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
    if (!last_block_of_strip(tickets, n_chunks_total)) return;
    load_strip_partials<3>(part, n_chunks_total, L, s_part);
    // thread rows 0 / 1 merge the query / doc segment of their column: chunk triples (n, mean, M2) in ascending
    // order with Chan's formula
    if (ty >= 2 || col >= L) return;
    const int seg = ty;
    const int c0 = seg == 0 ? 0 : fin.nq_chunks;
    const int c1 = seg == 0 ? fin.nq_chunks : n_chunks_total;
    if (c0 == c1) return;  // empty segment (single-instance call, B == R)
    float cn = 0.f, mu = 0.f, m2 = 0.f;
    for (int c = c0; c < c1; ++c)
        chan_merge(cn, mu, m2, s_part[(0 * n_chunks_total + c) * BN_TX + tx], s_part[(1 * n_chunks_total + c) * BN_TX + tx],
                   s_part[(2 * n_chunks_total + c) * BN_TX + tx]);
    bn_finalize_column(fin, seg * L + col, mu, m2 / cn);  // biased variance (tf.nn.moments)
}

// inference (new_dssm.py:85-86): the moments are the EMA shadows
__global__ void __launch_bounds__(256) bn_eval_affine_kernel(int L, BnFinalize fin) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * L) bn_write_affine(fin, i, fin.ema_mean[i], fin.ema_var[i]);
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
__global__ void __launch_bounds__(BN_TX * BN_TY_BWD)
bn_bwd_reduce_kernel(const float* __restrict__ dA, const float* __restrict__ H, int R, int L, int B, int act,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ scale,
                     const float* __restrict__ shift, float* __restrict__ part, int n_chunks_total, int chunk_rows,
                     int* __restrict__ tickets, const float* __restrict__ gamma, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ db /* [L] or NULL */,
                     float* __restrict__ sumx_out /* [2][L] or NULL: sum of xhat per instance (SyncBN needs it for db) */) {
    __shared__ float red0[BN_TY_BWD][BN_TX + 1];
    __shared__ float red1[BN_TY_BWD][BN_TX + 1];
    __shared__ float red2[BN_TY_BWD][BN_TX + 1];
    __shared__ float s_part[3 * 2 * BN_MAX_CHUNKS * BN_TX];
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
        for (; r + (MLP - 1) * BN_TY_BWD < r1; r += MLP * BN_TY_BWD) {
            float hv[MLP], dv[MLP];
#pragma unroll
            for (int u = 0; u < MLP; ++u) {
                hv[u] = __ldg(H + (size_t)(r + u * BN_TY_BWD) * L + col);
                dv[u] = __ldg(dA + (size_t)(r + u * BN_TY_BWD) * L + col);
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
        {   // remainder (< MLP rows per thread), still issued together
            float hv[MLP], dv[MLP];
#pragma unroll
            for (int u = 0; u < MLP; ++u) {
                const bool ok = r + u * BN_TY_BWD < r1;
                hv[u] = ok ? __ldg(H + (size_t)(r + u * BN_TY_BWD) * L + col) : 0.f;
                dv[u] = ok ? __ldg(dA + (size_t)(r + u * BN_TY_BWD) * L + col) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < MLP; ++u) {
                if (r + u * BN_TY_BWD < r1) {
                    const float a = act_fwd(fmaf(hv[u], sc, sh), act);
                    const float g = dv[u] * act_grad_from_out(a, act);
                    const float xh = (hv[u] - mu) * rs;
                    sg += g;
                    sgx = fmaf(g, xh, sgx);
                    sx += xh;
                }
            }
        }
    }
    red0[ty][tx] = sg;
    red1[ty][tx] = sgx;
    red2[ty][tx] = sx;
    __syncthreads();
    if (ty == 0 && col < L) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int i = 0; i < BN_TY_BWD; ++i) { t0 += red0[i][tx]; t1 += red1[i][tx]; t2 += red2[i][tx]; }
        part[((size_t)0 * n_chunks_total + chunk) * L + col] = t0;
        part[((size_t)1 * n_chunks_total + chunk) * L + col] = t1;
        part[((size_t)2 * n_chunks_total + chunk) * L + col] = t2;
    }
    if (!last_block_of_strip(tickets, n_chunks_total)) return;
    // dgamma, dbeta per (segment, column); optionally the pre-BN bias gradient db[c] = sum_r dH[r,c].  Under BN that
    // sum is identically  -gamma*rstd*dgamma*(sum_r xhat)/n  per segment (the g and dbeta terms cancel exactly), i.e.
    // rounding noise around zero -- evaluated from sum xhat of the same pass instead of another sweep over dH.
    load_strip_partials<3>(part, n_chunks_total, L, s_part);
    float* s_db = &red0[0][0];  // [2][BN_TX + 1], free after the block reduction above
    if (ty < 2) {
        float dbv = 0.f;
        if (col < L) {
            const int sg_ = ty;
            const int c0 = sg_ == 0 ? 0 : nq;
            const int c1 = sg_ == 0 ? nq : n_chunks_total;
            float b = 0.f, g = 0.f, x = 0.f;
            for (int c = c0; c < c1; ++c) {
                b += s_part[(0 * n_chunks_total + c) * BN_TX + tx];
                g += s_part[(1 * n_chunks_total + c) * BN_TX + tx];
                x += s_part[(2 * n_chunks_total + c) * BN_TX + tx];
            }
            const int i = sg_ * L + col;
            if (c0 < c1) {
                dbeta[i] = b;
                dgamma[i] = g;
                if (sumx_out) sumx_out[i] = x;
                const float n = sg_ == 0 ? (float)B : (float)(R - B);
                dbv = -(__ldg(gamma + i) * __ldg(rstd + i)) * g * (x / n);
            }
        }
        s_db[ty * (BN_TX + 1) + tx] = dbv;
    }
    __syncthreads();
    if (db && ty == 0 && col < L) db[col] = s_db[tx] + s_db[(BN_TX + 1) + tx];
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

// workspace = [ticket counters, one per 32-column strip | chunk partials]
static inline size_t bn_ticket_bytes(int L) { return align_up((size_t)cdiv(L, BN_TX) * sizeof(int), 256); }

extern "C" size_t dssm_bn_workspace_bytes(int32_t R, int32_t L) {
    if (R <= 0 || L <= 0) return 0;
    // worst case split of R rows into two segments adds one chunk
    const size_t chunks = (size_t)bn_chunks_of(R) + 2;
    return bn_ticket_bytes(L) + align_up(3 * chunks * (size_t)L * sizeof(float), 256);
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
    BnFinalize fin{gamma, beta, ema_mean, ema_var, mean, var, rstd, scale, shift, eps, ema_decay, update_ema, nq};
    if (on_train) {
        DSSM_REQUIRE(workspace && workspace_bytes >= bn_ticket_bytes(L) + (size_t)3 * nt * L * sizeof(float), DSSM_ERR_WORKSPACE,
                     "dssm_bn_forward: workspace too small");
        int* tickets = (int*)workspace;
        float* part = (float*)((char*)workspace + bn_ticket_bytes(L));
        dim3 grid(cdiv(L, BN_TX), nt), block(BN_TX, BN_TY);
        bn_stats_kernel<<<grid, block, 0, st>>>(X, R, L, B, part, nt, cr, tickets, fin);
        LAUNCH_CHECK("bn_stats");
    } else {
        bn_eval_affine_kernel<<<cdiv(2 * L, 256), 256, 0, st>>>(L, fin);
        LAUNCH_CHECK("bn_eval_affine");
    }
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

// The two passes of the backward, separately (SyncBN exchanges [dbeta | dgamma] between them; tower.cu).
extern "C" int dssm_bn_bwd_reduce_only(const float* dA, const float* H, int32_t R, int32_t L, int32_t B, int32_t act, const float* gamma,
                                       const float* mean, const float* rstd, const float* scale, const float* shift, float* dgamma,
                                       float* dbeta, float* db, float* sumx, void* workspace, size_t workspace_bytes,
                                       dssm_stream_t stream) {
    DSSM_REQUIRE(dA && H && gamma && mean && rstd && scale && shift && dgamma && dbeta, DSSM_ERR_BAD_ARG, "dssm_bn_act_backward: null BN pointer");
    DSSM_REQUIRE(R > 0 && L > 0 && B > 0 && B <= R, DSSM_ERR_BAD_SHAPE, "dssm_bn_act_backward: need 0 < B <= R");
    const int cr = bn_chunk_rows(R, B);
    const int nq = bn_chunks_of(B, cr), nd = bn_chunks_of(R - B, cr), nt = nq + nd;
    DSSM_REQUIRE(workspace && workspace_bytes >= bn_ticket_bytes(L) + (size_t)3 * nt * L * sizeof(float), DSSM_ERR_WORKSPACE,
                 "dssm_bn_act_backward: workspace too small");
    int* tickets = (int*)workspace;
    float* part = (float*)((char*)workspace + bn_ticket_bytes(L));
    dim3 grid(cdiv(L, BN_TX), nt), block(BN_TX, BN_TY_BWD);
    bn_bwd_reduce_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(dA, H, R, L, B, act, mean, rstd, scale, shift, part, nt, cr, tickets, gamma,
                                                                    dgamma, dbeta, db, sumx);
    LAUNCH_CHECK("bn_bwd_reduce");
    return DSSM_OK;
}

extern "C" int dssm_bn_bwd_apply_only(float* dA, const float* H, int32_t R, int32_t L, int32_t B, int32_t act, const float* gamma,
                                      const float* mean, const float* rstd, const float* scale, const float* shift, const float* dgamma,
                                      const float* dbeta, dssm_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
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
    int rc = dssm_bn_bwd_reduce_only(dA, H, R, L, B, act, gamma, mean, rstd, scale, shift, dgamma, dbeta, db, nullptr, workspace,
                                     workspace_bytes, stream);
    if (rc != DSSM_OK) return rc;
    return dssm_bn_bwd_apply_only(dA, H, R, L, B, act, gamma, mean, rstd, scale, shift, dgamma, dbeta, stream);
}
