// Merge_Negative_Doc (new_dssm.py:160-180), Cosine_Similarity (:182-201) and Loss (:203-213) with the
// gradient w.r.t. the embeddings, one warp per query group.  The merge is never materialised on the
// training path: group j reads its positive at row B+j and negative slot i at row 2B + j*NEG + i, which
// is exactly doc_y[(i+1)*B + j] of the reference's concat chain.  The standalone gather kernel and its
// index form exist for callers that want doc_y itself (and for the bit-exact ordering test).
#include "common.cuh"
#include <math.h>

namespace dssm {

constexpr int CL_WARPS = 4;

// dynamic smem per warp: 3*(1+NEG) floats (dot, dnorm, coefficient)
__global__ void __launch_bounds__(CL_WARPS * 32)
cos_softmax_loss_kernel(const float* __restrict__ Hin, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                        float* __restrict__ Y, int B, int NEG, int L, float gamma, float loss_eps, float inv_denom,
                        float* __restrict__ query_norm_single, float* __restrict__ doc_norm,
                        float* __restrict__ cos_sim_raw, float* __restrict__ cos_sim, float* __restrict__ prob,
                        float* __restrict__ loss_terms, float* __restrict__ dY) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int K1 = NEG + 1;
    float* s_dot = smem + (size_t)w * 3 * K1;
    float* s_dn = s_dot + K1;
    float* s_c = s_dn + K1;
    const int j = blockIdx.x * CL_WARPS + w;
    if (j >= B) return;
    if (Hin) {
        // fused last-layer BN + activation (new_dssm.py:87,156-158): this warp owns the query row and the 1+NEG doc rows
        // of group j, so it materialises their embeddings first; every lane later re-reads only what it wrote itself
        for (int k = -1; k <= NEG; ++k) {
            const size_t row = (k < 0) ? (size_t)j : (k == 0) ? (size_t)(B + j) : (size_t)(2 * B) + (size_t)j * NEG + (k - 1);
            const int o = (k < 0) ? 0 : L;
            for (int c = lane; c < L; c += 32) {
                float x = __ldg(Hin + row * L + c);
                if (scale) x = fmaf(x, __ldg(scale + o + c), __ldg(shift + o + c));
                Y[row * L + c] = act_fwd(x, act);
            }
        }
        __syncwarp();
    }
    const float* q = Y + (size_t)j * L;
    // ||q||
    float qq = 0.f;
    for (int c = lane; c < L; c += 32) {
        const float x = q[c];
        qq = fmaf(x, x, qq);
    }
    qq = warp_sum(qq);
    const float qn = sqrtf(qq);
    // dots and doc norms
    for (int k = 0; k < K1; ++k) {
        const size_t drow = (k == 0) ? (size_t)(B + j) : (size_t)(2 * B) + (size_t)j * NEG + (k - 1);
        const float* d = Y + drow * L;
        float dd = 0.f, dq = 0.f;
        for (int c = lane; c < L; c += 32) {
            const float x = d[c];
            dd = fmaf(x, x, dd);
            dq = fmaf(x, q[c], dq);
        }
        dd = warp_sum(dd);
        dq = warp_sum(dq);
        if (lane == 0) {
            s_dot[k] = dq;
            s_dn[k] = sqrtf(dd);
        }
    }
    __syncwarp();
    // softmax over the 1+NEG logits (lanes stride over k)
    float mx = -INFINITY;
    for (int k = lane; k < K1; k += 32) {
        const float raw = s_dot[k] / (qn * s_dn[k]);  // tf.truediv, no epsilon: 0/0 = NaN
        mx = fmaxf(mx, raw * gamma);
    }
    mx = warp_max(mx);
    float se = 0.f;
    for (int k = lane; k < K1; k += 32) {
        const float raw = s_dot[k] / (qn * s_dn[k]);
        se += expf(raw * gamma - mx);
    }
    se = warp_sum(se);
    const float raw0 = s_dot[0] / (qn * s_dn[0]);
    const float p0 = expf(raw0 * gamma - mx) / se;
    // NaN anywhere in the group poisons the softmax exactly as in TF (max/exp/sum propagate NaN);
    // fmaxf drops NaNs, so re-inject: if any logit is NaN the sum must be NaN.
    float any_nan = 0.f;
    for (int k = lane; k < K1; k += 32) {
        const float raw = s_dot[k] / (qn * s_dn[k]);
        if (raw != raw) any_nan = 1.f;
    }
    any_nan = warp_sum(any_nan);
    const float poison = any_nan > 0.f ? NAN : 0.f;
    const float wgt = p0 / (p0 + loss_eps);
    float sum_c_raw = 0.f;
    for (int k = lane; k < K1; k += 32) {
        const float raw = s_dot[k] / (qn * s_dn[k]);
        const float logit = raw * gamma;
        const float p = expf(logit - mx) / se + poison;
        if (cos_sim_raw) cos_sim_raw[(size_t)k * B + j] = raw;
        if (doc_norm) doc_norm[(size_t)k * B + j] = s_dn[k];
        if (cos_sim) cos_sim[(size_t)j * K1 + k] = logit;
        if (prob) prob[(size_t)j * K1 + k] = p;
        // dLoss/dcos_k = gamma * w * (p_k - [k==0]) / denom
        const float dcos = gamma * (wgt * (p - (k == 0 ? 1.f : 0.f)) * inv_denom);
        s_c[k] = dcos;
        sum_c_raw = fmaf(dcos, raw, sum_c_raw);
    }
    sum_c_raw = warp_sum(sum_c_raw);
    if (lane == 0) {
        if (query_norm_single) query_norm_single[j] = qn;
        loss_terms[j] = -logf(p0 + poison + loss_eps);
    }
    if (!dY) return;
    __syncwarp();
    // dq = sum_k c_k/(qn*dn_k) * d_k - (sum_k c_k raw_k)/qn^2 * q ;  dd_k = c_k/(qn*dn_k) * q - c_k raw_k/dn_k^2 * d_k
    const float qcoef = sum_c_raw / (qn * qn);
    for (int c = lane; c < L; c += 32) {
        const float qv = q[c];
        float dqv = 0.f;
        for (int k = 0; k < K1; ++k) {
            const size_t drow = (k == 0) ? (size_t)(B + j) : (size_t)(2 * B) + (size_t)j * NEG + (k - 1);
            const float dv = Y[drow * L + c];
            const float ck = s_c[k], dn = s_dn[k];
            const float inv_qd = 1.f / (qn * dn);
            const float raw = s_dot[k] * inv_qd;
            dqv = fmaf(ck * inv_qd, dv, dqv);
            dY[drow * L + c] = (ck * inv_qd) * qv - (ck * raw / (dn * dn)) * dv;
        }
        dY[(size_t)j * L + c] = dqv - qcoef * qv;
    }
}

// 128-bit variant for L % 4 == 0, L <= 128*NV: the same arithmetic with every row held as NV float4 per lane and
// CL_DB doc rows in flight per warp (loads and the 2*CL_DB shuffle reductions of a batch are independent chains), so
// a warp's latency is ~(1+NEG)/CL_DB round trips instead of ~4*(1+NEG).  B warps is all the parallelism a batch
// offers (7 warps per SM at B = 1024): the kernel is latency-bound, not bandwidth-bound.
template <int NV>
__global__ void __launch_bounds__(CL_WARPS * 32)
cos_softmax_loss_v4_kernel(const float4* __restrict__ Hin, const float4* __restrict__ scale, const float4* __restrict__ shift,
                           int act, float4* __restrict__ Y, int B, int NEG, int L4, float gamma, float loss_eps, float inv_denom,
                           float* __restrict__ query_norm_single, float* __restrict__ doc_norm,
                           float* __restrict__ cos_sim_raw, float* __restrict__ cos_sim, float* __restrict__ prob,
                           float* __restrict__ loss_terms, float4* __restrict__ dY) {
    constexpr int DB = 8 / NV;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int K1 = NEG + 1;
    float* s_dot = smem + (size_t)w * 3 * K1;
    float* s_dn = s_dot + K1;
    float* s_c = s_dn + K1;
    const int j = blockIdx.x * CL_WARPS + w;
    if (j >= B) return;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* src = Hin ? Hin : Y;
    const bool affine = Hin && scale;
    auto doc_row = [&](int k) { return (k == 0) ? (size_t)(B + j) : (size_t)(2 * B) + (size_t)j * NEG + (k - 1); };
    auto embed = [&](float4 x, float4 sc, float4 sh) {
        if (affine) { x.x = fmaf(x.x, sc.x, sh.x); x.y = fmaf(x.y, sc.y, sh.y); x.z = fmaf(x.z, sc.z, sh.z); x.w = fmaf(x.w, sc.w, sh.w); }
        if (Hin) { x.x = act_fwd(x.x, act); x.y = act_fwd(x.y, act); x.z = act_fwd(x.z, act); x.w = act_fwd(x.w, act); }
        return x;
    };
    // query row (BN instance 0) and the doc instance's affine in registers
    float4 qv[NV], scd[NV], shd[NV];
    float qq = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int c = lane + 32 * v;
        qv[v] = z4; scd[v] = z4; shd[v] = z4;
        if (c < L4) {
            float4 sc = z4, sh = z4;
            if (affine) { sc = __ldg(scale + c); sh = __ldg(shift + c); scd[v] = __ldg(scale + L4 + c); shd[v] = __ldg(shift + L4 + c); }
            qv[v] = embed(src[(size_t)j * L4 + c], sc, sh);
            if (Hin) Y[(size_t)j * L4 + c] = qv[v];
            qq = fmaf(qv[v].x, qv[v].x, qq); qq = fmaf(qv[v].y, qv[v].y, qq); qq = fmaf(qv[v].z, qv[v].z, qq); qq = fmaf(qv[v].w, qv[v].w, qq);
        }
    }
    qq = warp_sum(qq);
    const float qn = sqrtf(qq);
    // dots and doc norms, DB rows per round
    for (int k0 = 0; k0 < K1; k0 += DB) {
        float4 d[DB][NV];
#pragma unroll
        for (int u = 0; u < DB; ++u)
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + 32 * v;
                d[u][v] = (k0 + u < K1 && c < L4) ? src[doc_row(k0 + u) * L4 + c] : z4;
            }
        float dd[DB], dq[DB];
#pragma unroll
        for (int u = 0; u < DB; ++u) {
            dd[u] = 0.f; dq[u] = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + 32 * v;
                if (k0 + u < K1 && c < L4) {
                    const float4 x = embed(d[u][v], scd[v], shd[v]);
                    if (Hin) Y[doc_row(k0 + u) * L4 + c] = x;
                    dd[u] = fmaf(x.x, x.x, dd[u]); dd[u] = fmaf(x.y, x.y, dd[u]); dd[u] = fmaf(x.z, x.z, dd[u]); dd[u] = fmaf(x.w, x.w, dd[u]);
                    dq[u] = fmaf(x.x, qv[v].x, dq[u]); dq[u] = fmaf(x.y, qv[v].y, dq[u]); dq[u] = fmaf(x.z, qv[v].z, dq[u]); dq[u] = fmaf(x.w, qv[v].w, dq[u]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < DB; ++u) {
                dd[u] += __shfl_xor_sync(0xffffffffu, dd[u], o);
                dq[u] += __shfl_xor_sync(0xffffffffu, dq[u], o);
            }
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < DB; ++u)
                if (k0 + u < K1) { s_dot[k0 + u] = dq[u]; s_dn[k0 + u] = sqrtf(dd[u]); }
        }
    }
    __syncwarp();
    // softmax over the 1+NEG logits (lanes stride over k); s_c holds raw until the coefficients replace it
    float mx = -INFINITY, any_nan = 0.f;
    for (int k = lane; k < K1; k += 32) {
        const float raw = s_dot[k] / (qn * s_dn[k]);  // tf.truediv, no epsilon: 0/0 = NaN
        s_c[k] = raw;
        mx = fmaxf(mx, raw * gamma);
        if (raw != raw) any_nan = 1.f;
    }
    mx = warp_max(mx);
    any_nan = warp_sum(any_nan);
    float se = 0.f;
    for (int k = lane; k < K1; k += 32) se += expf(s_c[k] * gamma - mx);
    se = warp_sum(se);
    __syncwarp();
    const float p0 = expf(s_c[0] * gamma - mx) / se;
    // NaN anywhere in the group poisons the softmax exactly as in TF; fmaxf drops NaNs, so re-inject
    const float poison = any_nan > 0.f ? NAN : 0.f;
    const float wgt = p0 / (p0 + loss_eps);
    float sum_c_raw = 0.f;
    __syncwarp();
    for (int k = lane; k < K1; k += 32) {
        const float raw = s_c[k];
        const float logit = raw * gamma;
        const float p = expf(logit - mx) / se + poison;
        if (cos_sim_raw) cos_sim_raw[(size_t)k * B + j] = raw;
        if (doc_norm) doc_norm[(size_t)k * B + j] = s_dn[k];
        if (cos_sim) cos_sim[(size_t)j * K1 + k] = logit;
        if (prob) prob[(size_t)j * K1 + k] = p;
        // dLoss/dcos_k = gamma * w * (p_k - [k==0]) / denom
        const float dcos = gamma * (wgt * (p - (k == 0 ? 1.f : 0.f)) * inv_denom);
        sum_c_raw = fmaf(dcos, raw, sum_c_raw);
        // row coefficients of the backward: dd_k = a_k * q - b_k * d_k,  dq = sum_k a_k * d_k - qcoef * q
        const float dn = s_dn[k];
        const float inv_qd = 1.f / (qn * dn);
        s_dot[k] = dcos * inv_qd;                            // a_k
        s_c[k] = dcos * raw / (dn * dn);                     // b_k
    }
    sum_c_raw = warp_sum(sum_c_raw);
    if (lane == 0) {
        if (query_norm_single) query_norm_single[j] = qn;
        loss_terms[j] = -logf(p0 + poison + loss_eps);
    }
    if (!dY) return;
    __syncwarp();
    const float qcoef = sum_c_raw / (qn * qn);
    float4 dqa[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) dqa[v] = z4;
    for (int k0 = 0; k0 < K1; k0 += DB) {
        float4 d[DB][NV];
#pragma unroll
        for (int u = 0; u < DB; ++u)
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + 32 * v;
                d[u][v] = (k0 + u < K1 && c < L4) ? Y[doc_row(k0 + u) * L4 + c] : z4;
            }
#pragma unroll
        for (int u = 0; u < DB; ++u) {
            if (k0 + u >= K1) continue;
            const float a = s_dot[k0 + u], b = s_c[k0 + u];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + 32 * v;
                if (c < L4) {
                    const float4 x = d[u][v];
                    float4 o;
                    o.x = a * qv[v].x - b * x.x; o.y = a * qv[v].y - b * x.y; o.z = a * qv[v].z - b * x.z; o.w = a * qv[v].w - b * x.w;
                    dY[doc_row(k0 + u) * L4 + c] = o;
                    dqa[v].x = fmaf(a, x.x, dqa[v].x); dqa[v].y = fmaf(a, x.y, dqa[v].y);
                    dqa[v].z = fmaf(a, x.z, dqa[v].z); dqa[v].w = fmaf(a, x.w, dqa[v].w);
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int c = lane + 32 * v;
        if (c < L4) {
            float4 o;
            o.x = dqa[v].x - qcoef * qv[v].x; o.y = dqa[v].y - qcoef * qv[v].y;
            o.z = dqa[v].z - qcoef * qv[v].z; o.w = dqa[v].w - qcoef * qv[v].w;
            dY[(size_t)j * L4 + c] = o;
        }
    }
}

// loss = sum_j terms[j] * inv_denom, single block, fixed tree
__global__ void __launch_bounds__(1024) loss_reduce_kernel(const float* __restrict__ terms, int B, float inv_denom,
                                                            float* __restrict__ loss) {
    __shared__ float s[1024];
    float a = 0.f;
    for (int i = threadIdx.x; i < B; i += 1024) a += terms[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = s[0] * inv_denom;
}

__global__ void merge_negative_doc_kernel(const float* __restrict__ pos, const float* __restrict__ neg, int B, int NEG,
                                          int L, float* __restrict__ doc_y) {
    const size_t total = (size_t)(1 + NEG) * B * L;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / L;
        const int c = (int)(i - r * L);
        float v;
        if (r < (size_t)B) {
            v = __ldg(pos + r * L + c);
        } else {
            const size_t t = r - B;              // t = i_slot*B + j
            const size_t slot = t / B, j = t - slot * B;
            v = __ldg(neg + (j * NEG + slot) * L + c);
        }
        doc_y[i] = v;
    }
}

__global__ void merge_negative_doc_index_kernel(int B, int NEG, int* __restrict__ src) {
    const int total = (1 + NEG) * B;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += gridDim.x * blockDim.x) {
        if (r < B) {
            src[r] = r;
        } else {
            const int t = r - B, slot = t / B, j = t - slot * B;
            src[r] = B + j * NEG + slot;
        }
    }
}

}  // namespace dssm

using namespace dssm;

static int cos_softmax_loss_impl(const float* Hin, const float* scale, const float* shift, int act, float* Y, int32_t B, int32_t NEG,
                                 int32_t L, float gamma, float loss_eps, int32_t loss_div_bs, float* query_norm_single,
                                 float* doc_norm, float* cos_sim_raw, float* cos_sim, float* prob, float* loss_terms, float* loss,
                                 float* dY, dssm_stream_t stream) {
    DSSM_REQUIRE(Y && loss_terms, DSSM_ERR_BAD_ARG, "dssm_cos_softmax_loss: null pointer");
    DSSM_REQUIRE(B > 0 && NEG > 0 && L > 0, DSSM_ERR_BAD_SHAPE, "dssm_cos_softmax_loss: bad shape B=%d NEG=%d L=%d", B, NEG, L);
    const size_t smem = (size_t)CL_WARPS * 3 * (NEG + 1) * sizeof(float);
    DSSM_REQUIRE(smem <= 48 * 1024, DSSM_ERR_BAD_SHAPE, "dssm_cos_softmax_loss: NEG=%d too large", NEG);
    cudaStream_t st = (cudaStream_t)stream;
    const float inv_denom = loss_div_bs ? 1.0f / (float)B : 1.0f;
    const bool v4 = L % 4 == 0 && L <= 256 && aligned16(Y) && (!Hin || aligned16(Hin)) && (!scale || (aligned16(scale) && aligned16(shift))) &&
                    (!dY || aligned16(dY));
#define CL_V4_ARGS (const float4*)Hin, (const float4*)scale, (const float4*)shift, act, (float4*)Y, B, NEG, L / 4, gamma, loss_eps, inv_denom, \
                   query_norm_single, doc_norm, cos_sim_raw, cos_sim, prob, loss_terms, (float4*)dY
    if (v4 && L <= 128)
        cos_softmax_loss_v4_kernel<1><<<cdiv(B, CL_WARPS), CL_WARPS * 32, smem, st>>>(CL_V4_ARGS);
    else if (v4)
        cos_softmax_loss_v4_kernel<2><<<cdiv(B, CL_WARPS), CL_WARPS * 32, smem, st>>>(CL_V4_ARGS);
    else
        cos_softmax_loss_kernel<<<cdiv(B, CL_WARPS), CL_WARPS * 32, smem, st>>>(Hin, scale, shift, act, Y, B, NEG, L, gamma, loss_eps, inv_denom,
                                                                                query_norm_single, doc_norm, cos_sim_raw,
                                                                                cos_sim, prob, loss_terms, dY);
#undef CL_V4_ARGS
    LAUNCH_CHECK("cos_softmax_loss");
    if (loss) {
        loss_reduce_kernel<<<1, 1024, 0, st>>>(loss_terms, B, inv_denom, loss);
        LAUNCH_CHECK("loss_reduce");
    }
    return DSSM_OK;
}

extern "C" int dssm_cos_softmax_loss(const float* Y, int32_t B, int32_t NEG, int32_t L, float gamma, float loss_eps,
                                     int32_t loss_div_bs, float* query_norm_single, float* doc_norm, float* cos_sim_raw,
                                     float* cos_sim, float* prob, float* loss_terms, float* loss, float* dY,
                                     dssm_stream_t stream) {
    return cos_softmax_loss_impl(nullptr, nullptr, nullptr, DSSM_ACT_NONE, const_cast<float*>(Y), B, NEG, L, gamma, loss_eps, loss_div_bs,
                                 query_norm_single, doc_norm, cos_sim_raw, cos_sim, prob, loss_terms, loss, dY, stream);
}

// Same, with the last layer's BN + activation fused in front: reads the pre-BN activations H [R,L] and the [2][L]
// scale/shift (NULL = identity), WRITES the embeddings Y, then proceeds as above.
extern "C" int dssm_cos_softmax_loss_fused(const float* H, const float* scale, const float* shift, int32_t act, float* Y, int32_t B,
                                           int32_t NEG, int32_t L, float gamma, float loss_eps, int32_t loss_div_bs,
                                           float* query_norm_single, float* doc_norm, float* cos_sim_raw, float* cos_sim,
                                           float* prob, float* loss_terms, float* loss, float* dY, dssm_stream_t stream) {
    DSSM_REQUIRE(H, DSSM_ERR_BAD_ARG, "dssm_cos_softmax_loss_fused: null pointer");
    DSSM_REQUIRE((scale == nullptr) == (shift == nullptr), DSSM_ERR_BAD_ARG, "dssm_cos_softmax_loss_fused: scale/shift must both be set or both NULL");
    return cos_softmax_loss_impl(H, scale, shift, act, Y, B, NEG, L, gamma, loss_eps, loss_div_bs, query_norm_single, doc_norm,
                                 cos_sim_raw, cos_sim, prob, loss_terms, loss, dY, stream);
}

extern "C" int dssm_merge_negative_doc(const float* doc_positive_y, const float* doc_negative_y, int32_t B, int32_t NEG,
                                       int32_t L, float* doc_y, dssm_stream_t stream) {
    DSSM_REQUIRE(doc_positive_y && doc_negative_y && doc_y, DSSM_ERR_BAD_ARG, "dssm_merge_negative_doc: null pointer");
    DSSM_REQUIRE(B > 0 && NEG >= 0 && L > 0, DSSM_ERR_BAD_SHAPE, "dssm_merge_negative_doc: bad shape");
    const size_t total = (size_t)(1 + NEG) * B * L;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    merge_negative_doc_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(doc_positive_y, doc_negative_y, B, NEG, L, doc_y);
    LAUNCH_CHECK("merge_negative_doc");
    return DSSM_OK;
}

extern "C" int dssm_merge_negative_doc_index(int32_t B, int32_t NEG, int32_t* src, dssm_stream_t stream) {
    DSSM_REQUIRE(src, DSSM_ERR_BAD_ARG, "dssm_merge_negative_doc_index: null pointer");
    DSSM_REQUIRE(B > 0 && NEG >= 0, DSSM_ERR_BAD_SHAPE, "dssm_merge_negative_doc_index: bad shape");
    merge_negative_doc_index_kernel<<<cdiv((int64_t)(1 + NEG) * B, 256), 256, 0, (cudaStream_t)stream>>>(B, NEG, src);
    LAUNCH_CHECK("merge_negative_doc_index");
    return DSSM_OK;
}
