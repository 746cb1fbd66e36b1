// Data-parallel exchange of the FC1 weight gradient over NVLink peer memory, fused with Adam (no reference
// counterpart: the reference is single-process; this is the reduce-scatter -> sharded Adam -> all-gather of SURVEY 8e
// done by ONE kernel).  Every rank has written its dense dW1 into a buffer its peers can address (symmetric memory);
// rank r owns the rows [row_begin, row_end) of W1.  For each owned row a warp PULLS the row from every rank over
// NVLink (peer loads, all in flight together), sums them in rank order, averages, applies TF-Adam with its local m, v
// and PUSHES the new weight row into every rank's W1.  Per GPU and step that is (n-1)/n * |W1| in and out -- the wire
// volume of a ring all-reduce -- but the optimizer state traffic shrinks n-fold and nothing is staged twice.
// The caller orders it between two cross-device barriers (all dW1 complete / all W1 rows landed).
#include "common.cuh"

namespace dssm {

constexpr int NV_THREADS = 256;

struct PeerPtrs {
    const float4* dW[DSSM_MAX_PEERS];
    float4* W[DSSM_MAX_PEERS];
};

// NR = compile-time bound on the rank count (2, 4, 8, 16).  HOIST: all NCH * n_ranks peer loads of a row are issued before
// the first use -- a peer load costs an NVLink round trip (~2 us), so bytes in flight per SM, not occupancy, decide
// whether the pull runs at link rate (at n = 2 a lane has only NCH remote loads to overlap).  For NR = 16 the row is
// processed one float4 chunk at a time to stay inside the register file.
template <int NCH, int NR, bool HOIST>
__global__ void __launch_bounds__(NV_THREADS)
w1_shard_reduce_adam_kernel(PeerPtrs p, int n_ranks, int self, int L4, int row_begin, int row_end, float4* __restrict__ m,
                            float4* __restrict__ v, const float* __restrict__ beta_pow, float lr, float b1, float b2, float eps) {
    const float b1p = __ldg(beta_pow), b2p = __ldg(beta_pow + 1);
    const float lr_t = lr * sqrtf(1.f - b2p) / (1.f - b1p);
    const float inv_n = 1.0f / (float)n_ranks;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto adam4 = [&](float4 g, float4& pp, float4& mm, float4& vv) {
#define ADAM1(x)                                          \
    {                                                     \
        const float gr = g.x * inv_n;                     \
        mm.x = b1 * mm.x + (1.f - b1) * gr;               \
        vv.x = b2 * vv.x + (1.f - b2) * (gr * gr);        \
        pp.x = pp.x - lr_t * mm.x / (sqrtf(vv.x) + eps);  \
    }
        ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
    };
    for (int row = row_begin + blockIdx.x * wpb + (threadIdx.x >> 5); row < row_end; row += gridDim.x * wpb) {
        if (HOIST) {
            float4 part[NCH][NR], pp[NCH], mm[NCH], vv[NCH];
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int col = lane + 32 * k;
                const size_t i = (size_t)row * L4 + col;
#pragma unroll
                for (int r = 0; r < NR; ++r) part[k][r] = (col < L4 && r < n_ranks) ? __ldcv(p.dW[r] + i) : z4;
            }
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int col = lane + 32 * k;
                const size_t i = (size_t)row * L4 + col;
                if (col < L4) { pp[k] = p.W[self][i]; mm[k] = m[i]; vv[k] = v[i]; }
            }
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int col = lane + 32 * k;
                if (col >= L4) continue;
                const size_t i = (size_t)row * L4 + col;
                float4 g = z4;
#pragma unroll
                for (int r = 0; r < NR; ++r)  // rank order: the one owner of a row fixes the summation order for everybody
                    if (r < n_ranks) { g.x += part[k][r].x; g.y += part[k][r].y; g.z += part[k][r].z; g.w += part[k][r].w; }
                adam4(g, pp[k], mm[k], vv[k]);
                m[i] = mm[k];
                v[i] = vv[k];
#pragma unroll
                for (int r = 0; r < NR; ++r)
                    if (r < n_ranks) p.W[r][i] = pp[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int col = lane + 32 * k;
                if (col >= L4) continue;
                const size_t i = (size_t)row * L4 + col;
                float4 part[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r)
                    if (r < n_ranks) part[r] = __ldcv(p.dW[r] + i);  // peer memory: never from a stale cache line
                float4 pp = p.W[self][i], mm = m[i], vv = v[i];
                float4 g = z4;
#pragma unroll
                for (int r = 0; r < NR; ++r)
                    if (r < n_ranks) { g.x += part[r].x; g.y += part[r].y; g.z += part[r].z; g.w += part[r].w; }
                adam4(g, pp, mm, vv);
                m[i] = mm;
                v[i] = vv;
#pragma unroll
                for (int r = 0; r < NR; ++r)
                    if (r < n_ranks) p.W[r][i] = pp;
            }
        }
    }
}

template <int NCH>
static void launch_shard(const PeerPtrs& p, int n_ranks, int self, int L1, int row_begin, int row_end, float* m, float* v,
                         const float* beta_pow, float lr, float b1, float b2, float eps, cudaStream_t st) {
    int blocks = sm_count() * 2;
    const int need = cdiv(row_end - row_begin, NV_THREADS / 32);
    if (blocks > need) blocks = need;
#define NV_ARGS p, n_ranks, self, L1 / 4, row_begin, row_end, (float4*)m, (float4*)v, beta_pow, lr, b1, b2, eps
    if (n_ranks <= 2) w1_shard_reduce_adam_kernel<NCH, 2, true><<<blocks, NV_THREADS, 0, st>>>(NV_ARGS);
    else if (n_ranks <= 4) w1_shard_reduce_adam_kernel<NCH, 4, true><<<blocks, NV_THREADS, 0, st>>>(NV_ARGS);
    else if (n_ranks <= 8 && NCH <= 4) w1_shard_reduce_adam_kernel<NCH, 8, true><<<blocks, NV_THREADS, 0, st>>>(NV_ARGS);
    else if (n_ranks <= 8) w1_shard_reduce_adam_kernel<NCH, 8, false><<<blocks, NV_THREADS, 0, st>>>(NV_ARGS);
    else w1_shard_reduce_adam_kernel<NCH, DSSM_MAX_PEERS, false><<<blocks, NV_THREADS, 0, st>>>(NV_ARGS);
#undef NV_ARGS
}

}  // namespace dssm

using namespace dssm;

extern "C" int dssm_w1_shard_reduce_adam(const float* const* peer_dW1, float* const* peer_W1, int32_t n_ranks, int32_t self, int32_t D,
                                         int32_t L1, int32_t row_begin, int32_t row_end, float* m1, float* v1, const float* beta_pow,
                                         float lr, float beta1, float beta2, float eps, dssm_stream_t stream) {
    DSSM_REQUIRE(peer_dW1 && peer_W1 && m1 && v1 && beta_pow, DSSM_ERR_BAD_ARG, "dssm_w1_shard_reduce_adam: null pointer");
    DSSM_REQUIRE(n_ranks >= 1 && n_ranks <= DSSM_MAX_PEERS && self >= 0 && self < n_ranks, DSSM_ERR_BAD_ARG,
                 "dssm_w1_shard_reduce_adam: n_ranks=%d self=%d (at most %d peers)", n_ranks, self, DSSM_MAX_PEERS);
    DSSM_REQUIRE(D > 0 && L1 > 0 && L1 % 4 == 0 && L1 <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_w1_shard_reduce_adam: bad shape");
    DSSM_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= D, DSSM_ERR_BAD_ARG, "dssm_w1_shard_reduce_adam: bad row range");
    PeerPtrs p{};
    for (int r = 0; r < n_ranks; ++r) {
        DSSM_REQUIRE(peer_dW1[r] && peer_W1[r] && aligned16(peer_dW1[r]) && aligned16(peer_W1[r]), DSSM_ERR_BAD_ALIGN,
                     "dssm_w1_shard_reduce_adam: peer buffer %d null or not 16-byte aligned", r);
        p.dW[r] = (const float4*)peer_dW1[r];
        p.W[r] = (float4*)peer_W1[r];
    }
    DSSM_REQUIRE(aligned16(m1) && aligned16(v1), DSSM_ERR_BAD_ALIGN, "dssm_w1_shard_reduce_adam: m1/v1 must be 16-byte aligned");
    if (row_begin == row_end) return DSSM_OK;
    const int nch = cdiv(L1 / 4, 32);
    DISPATCH_NCH(nch, launch_shard<N_>(p, n_ranks, self, L1, row_begin, row_end, m1, v1, beta_pow, lr, beta1, beta2, eps, (cudaStream_t)stream));
    LAUNCH_CHECK("w1_shard_reduce_adam");
    return DSSM_OK;
}
