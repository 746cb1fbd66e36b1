// Pieces of batch_normalization (new_dssm.py:62-88) shared by the stand-alone BN kernels (bn.cu) and the GEMM epilogue
// that takes the column moments of its own output (fc_tc.cu).
#pragma once
#include "common.cuh"

namespace dssm {

struct BnFinalize {  // outputs of the forward finalize, all [2][L]
    const float *gamma, *beta;
    float *ema_mean, *ema_var, *mean, *var, *rstd, *scale, *shift;
    float eps, decay;
    int update_ema, nq_chunks;
};

// Column moments taken in the epilogue of the GEMM that PRODUCES the tensor: every CTA writes
// the (count, mean, M2) triple of its 128-row tile per output column; the CTA that draws the last ticket of its N tile
// merges the tiles of each BN instance in ascending order (Chan) and finalizes -- mean, biased variance, EMA, scale, shift
// -- so the consumer finds them ready and the separate bn_stats launch disappears.  Needs Bseg % 128 == 0 (a tile never
// straddles the query / doc boundary); fin.nq_chunks = number of M tiles of the query instance.
struct FusedBnStats {
    float* part;   // [3][n_mtiles][N]
    int* tickets;  // one per N tile; zero-filled once, self-resetting
    BnFinalize fin;
    int on;
};

#ifdef __CUDACC__
__device__ __forceinline__ void bn_write_affine(const BnFinalize& f, int i, float mean, float var) {
    const float rstd = 1.0f / sqrtf(var + f.eps);
    const float sc = rstd * f.gamma[i];
    f.mean[i] = mean;
    f.var[i] = var;
    f.rstd[i] = rstd;
    f.scale[i] = sc;
    f.shift[i] = f.beta[i] - mean * sc;
}

// batch moments of one (instance, column) are final: EMA shadows (ExponentialMovingAverage.apply: shadow -= (1 - decay) *
// (shadow - value), new_dssm.py:78-81) and the affine form consumers apply
__device__ __forceinline__ void bn_finalize_column(const BnFinalize& f, int i, float mu, float var) {
    if (f.update_ema) {
        const float em = f.ema_mean[i], ev = f.ema_var[i];
        f.ema_mean[i] = em - (1.f - f.decay) * (em - mu);
        f.ema_var[i] = ev - (1.f - f.decay) * (ev - var);
    }
    bn_write_affine(f, i, mu, var);
}

// Chan's parallel update of (n, mean, M2) with another triple
__device__ __forceinline__ void chan_merge(float& cn, float& mu, float& m2, float nb, float mb, float qb) {
    if (nb <= 0.f) return;
    const float nt = cn + nb;
    const float delta = mb - mu;
    mu = mu + delta * (nb / nt);
    m2 = m2 + qb + delta * delta * (cn * nb / nt);
    cn = nt;
}
#endif

}  // namespace dssm
