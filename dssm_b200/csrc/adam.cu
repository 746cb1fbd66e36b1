// Training (new_dssm.py:215-217): tf.train.AdamOptimizer over the flat parameter buffer.
// Pure streaming: 16 B of reads (g,w,m,v) and 12 B of writes (w,m,v) per parameter = 28 B/param,
// 128-bit accesses, grid sized to the SM count.  The bias-correction powers live on the device so a
// captured CUDA graph can be replayed for every step.
#include "common.cuh"

namespace dssm {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            const float* __restrict__ beta_pow, float lr, float b1, float b2, float eps, float gscale) {
    const float b1p = __ldg(beta_pow), b2p = __ldg(beta_pow + 1);
    const float lr_t = lr * sqrtf(1.f - b2p) / (1.f - b1p);
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 gg = ldg_stream4(g4 + i);
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
#define ADAM1(c)                                              \
    {                                                         \
        const float gr = gg.c * gscale;                       \
        mm.c = b1 * mm.c + (1.f - b1) * gr;                   \
        vv.c = b2 * vv.c + (1.f - b2) * (gr * gr);            \
        pp.c = pp.c - lr_t * mm.c / (sqrtf(vv.c) + eps);      \
    }
        ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
        p4[i] = pp;
        m4[i] = mm;
        v4[i] = vv;
    }
    // tail (n not a multiple of 4)
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gr = g[i] * gscale;
        const float mm = b1 * m[i] + (1.f - b1) * gr;
        const float vv = b2 * v[i] + (1.f - b2) * (gr * gr);
        m[i] = mm;
        v[i] = vv;
        p[i] = p[i] - lr_t * mm / (sqrtf(vv) + eps);
    }
}

__global__ void adam_advance_kernel(float* beta_pow, float b1, float b2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        beta_pow[0] *= b1;
        beta_pow[1] *= b2;
    }
}

}  // namespace dssm

using namespace dssm;

extern "C" int dssm_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, const float* beta_pow,
                              float lr, float beta1, float beta2, float eps, float grad_scale, dssm_stream_t stream) {
    DSSM_REQUIRE(params && grads && m && v && beta_pow, DSSM_ERR_BAD_ARG, "dssm_adam_step: null pointer");
    DSSM_REQUIRE(n >= 0, DSSM_ERR_BAD_ARG, "dssm_adam_step: negative size");
    if (n == 0) return DSSM_OK;
    DSSM_REQUIRE(aligned16(params) && aligned16(grads) && aligned16(m) && aligned16(v), DSSM_ERR_BAD_ALIGN,
                 "dssm_adam_step: buffers must be 16-byte aligned");
    int64_t blocks = ((n >> 2) + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, m, v, n, beta_pow, lr, beta1, beta2, eps, grad_scale);
    LAUNCH_CHECK("adam");
    return DSSM_OK;
}

extern "C" int dssm_adam_advance(float* beta_pow, float beta1, float beta2, dssm_stream_t stream) {
    DSSM_REQUIRE(beta_pow, DSSM_ERR_BAD_ARG, "dssm_adam_advance: null pointer");
    adam_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(beta_pow, beta1, beta2);
    LAUNCH_CHECK("adam_advance");
    return DSSM_OK;
}
