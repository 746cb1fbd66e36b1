// Accuracy / Auc stages of the reference graph (new_dssm.py:219-231) on the device.
//
// tf.metrics.auc(labels, predictions, num_thresholds=T) keeps four confusion counters per threshold and the reference
// never re-initialises them (local variables are initialised once, new_dssm.py:252, and auc_op runs on every eval batch
// of every epoch, :281,308).  A prediction is positive at threshold t iff prediction > t, so all four counters follow
// from ONE histogram per class over "how many thresholds does this prediction exceed":
//      tp[i] = #positives whose bucket > i,   fp[i] likewise,   fn = P - tp,   tn = N - fp.
// auc_update_kernel adds one batch to the two (never reset) 64-bit histograms; auc_result_kernel turns them into the
// trapezoidal ROC area exactly as TF-1.x does (epsilon 1e-6 in the rates, float64 arithmetic here).
#include "common.cuh"

namespace dssm {

constexpr int AUC_THREADS = 256;
constexpr int AUC_MAX_T = 2048;  // thresholds (static shared memory of the result kernel: 2 x (T+1) doubles)

// number of thresholds strictly below p (thresholds ascending); NaN exceeds none, like TF's `predictions > t`
__device__ __forceinline__ int auc_bucket(const float* __restrict__ thr, int T, float p) {
    int lo = 0, hi = T;  // first index with !(thr[i] < p)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(thr + mid) < p) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// predictions [n]: the first n_pos entries carry label 1, the rest label 0 (new_dssm.py:163-165: label = [1]*B + [0]*B*NEG,
// aligned with cos_sim_raw's row order)
__global__ void __launch_bounds__(AUC_THREADS)
auc_update_kernel(const float* __restrict__ pred, int n_pos, int n, const float* __restrict__ thr, int T,
                  unsigned long long* __restrict__ pos_hist, unsigned long long* __restrict__ neg_hist) {
    extern __shared__ unsigned int auc_sh[];  // [2][T+1]
    unsigned int* hp = auc_sh;
    unsigned int* hn = auc_sh + (T + 1);
    for (int i = threadIdx.x; i < 2 * (T + 1); i += blockDim.x) auc_sh[i] = 0u;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = auc_bucket(thr, T, __ldg(pred + i));
        atomicAdd((i < n_pos ? hp : hn) + b, 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= T; i += blockDim.x) {
        if (hp[i]) atomicAdd(pos_hist + i, (unsigned long long)hp[i]);
        if (hn[i]) atomicAdd(neg_hist + i, (unsigned long long)hn[i]);
    }
}

// one block; out[0] = auc, out[1] = positives seen, out[2] = negatives seen (doubles)
__global__ void __launch_bounds__(1024)
auc_result_kernel(const unsigned long long* __restrict__ pos_hist, const unsigned long long* __restrict__ neg_hist, int T,
                  double* __restrict__ out) {
    __shared__ double tp[AUC_MAX_T + 1], fp[AUC_MAX_T + 1];
    __shared__ double red[1024];
    const int t = threadIdx.x;
    // suffix sums by one warp-free pass per class: T is a few thousand, the loop is 2(T+1) dependent adds
    if (t < 2) {
        const unsigned long long* h = t == 0 ? pos_hist : neg_hist;
        double* dst = t == 0 ? tp : fp;
        unsigned long long run = 0;
        for (int b = T; b >= 0; --b) {
            dst[b] = (double)run;  // entries with bucket > b
            run += h[b];
        }
        red[t] = (double)run;  // class total
    }
    __syncthreads();
    const double P = red[0], N = red[1];
    __syncthreads();
    const double eps = 1e-6;
    double acc = 0.0;
    for (int i = t; i < T - 1; i += blockDim.x) {
        const double tpr0 = (tp[i] + eps) / (P + eps), tpr1 = (tp[i + 1] + eps) / (P + eps);
        const double fpr0 = fp[i] / (N + eps), fpr1 = fp[i + 1] / (N + eps);
        acc += (fpr0 - fpr1) * (tpr0 + tpr1) / 2.0;
    }
    red[t] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {  // fixed tree: deterministic
        if (t < s) red[t] += red[t + s];
        __syncthreads();
    }
    if (t == 0) {
        out[0] = red[0];
        out[1] = P;
        out[2] = N;
    }
}

// accuracy = mean(argmax(prob, 1) == 0)  (new_dssm.py:220-221; argmax returns the first maximum)
__global__ void __launch_bounds__(256)
accuracy_kernel(const float* __restrict__ prob, int B, int C, float* __restrict__ out) {
    __shared__ int red[256];
    int hits = 0;
    for (int j = threadIdx.x; j < B; j += blockDim.x) {
        const float p0 = __ldg(prob + (size_t)j * C);
        bool first = true;  // column 0 is the argmax iff no later column is strictly larger ... and p0 is not beaten by NaN rules
        for (int k = 1; k < C; ++k) first = first && !(__ldg(prob + (size_t)j * C + k) > p0);
        hits += first ? 1 : 0;
    }
    red[threadIdx.x] = hits;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)red[0] / (float)B;
}

}  // namespace dssm

using namespace dssm;

extern "C" int dssm_auc_update(const float* predictions, int32_t n_pos, int32_t n, const float* thresholds, int32_t num_thresholds,
                               uint64_t* pos_hist, uint64_t* neg_hist, dssm_stream_t stream) {
    DSSM_REQUIRE(predictions && thresholds && pos_hist && neg_hist, DSSM_ERR_BAD_ARG, "dssm_auc_update: null pointer");
    DSSM_REQUIRE(n >= 0 && n_pos >= 0 && n_pos <= n, DSSM_ERR_BAD_ARG, "dssm_auc_update: need 0 <= n_pos <= n");
    DSSM_REQUIRE(num_thresholds >= 2 && num_thresholds <= AUC_MAX_T, DSSM_ERR_BAD_SHAPE, "dssm_auc_update: num_thresholds=%d out of [2,%d]",
                 num_thresholds, AUC_MAX_T);
    if (n == 0) return DSSM_OK;
    int blocks = cdiv(n, AUC_THREADS * 4);
    if (blocks > sm_count()) blocks = sm_count();
    const size_t smem = (size_t)2 * (num_thresholds + 1) * sizeof(unsigned int);
    auc_update_kernel<<<blocks, AUC_THREADS, smem, (cudaStream_t)stream>>>(predictions, n_pos, n, thresholds, num_thresholds,
                                                                         (unsigned long long*)pos_hist, (unsigned long long*)neg_hist);
    LAUNCH_CHECK("auc_update");
    return DSSM_OK;
}

extern "C" int dssm_auc_result(const uint64_t* pos_hist, const uint64_t* neg_hist, int32_t num_thresholds, double* out,
                               dssm_stream_t stream) {
    DSSM_REQUIRE(pos_hist && neg_hist && out, DSSM_ERR_BAD_ARG, "dssm_auc_result: null pointer");
    DSSM_REQUIRE(num_thresholds >= 2 && num_thresholds <= AUC_MAX_T, DSSM_ERR_BAD_SHAPE, "dssm_auc_result: num_thresholds=%d out of [2,%d]",
                 num_thresholds, AUC_MAX_T);
    auc_result_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((const unsigned long long*)pos_hist, (const unsigned long long*)neg_hist,
                                                           num_thresholds, out);
    LAUNCH_CHECK("auc_result");
    return DSSM_OK;
}

extern "C" int dssm_accuracy(const float* prob, int32_t B, int32_t n_classes, float* out, dssm_stream_t stream) {
    DSSM_REQUIRE(prob && out, DSSM_ERR_BAD_ARG, "dssm_accuracy: null pointer");
    DSSM_REQUIRE(B > 0 && n_classes > 0, DSSM_ERR_BAD_SHAPE, "dssm_accuracy: bad shape");
    accuracy_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(prob, B, n_classes, out);
    LAUNCH_CHECK("accuracy");
    return DSSM_OK;
}
