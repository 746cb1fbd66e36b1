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

// NR = compile-time bound on the rank count (2, 4, 8, 16); ROWS = rows a warp handles per iteration.  All
// ROWS * NCH * n_ranks peer loads of an iteration are issued before the first use: a peer load costs an NVLink round
// trip (~2-3 us), so the bytes in flight per SM -- not occupancy -- decide whether the pull runs at link rate
// (measured at n = 2 with one row per iteration: 120 us for 30 MB in + 30 MB out).  ROWS * NR is kept at 8.
template <int NCH, int NR, int ROWS>
__global__ void __launch_bounds__(NV_THREADS)
w1_shard_reduce_adam_kernel(PeerPtrs p, int n_ranks, int self, int L4, int row_begin, int row_end, float4* __restrict__ m,
                            float4* __restrict__ v, const float* __restrict__ beta_pow, float lr, float b1, float b2, float eps) {
    const float b1p = __ldg(beta_pow), b2p = __ldg(beta_pow + 1);
    const float lr_t = lr * sqrtf(1.f - b2p) / (1.f - b1p);
    const float inv_n = 1.0f / (float)n_ranks;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int warp_global = blockIdx.x * wpb + (threadIdx.x >> 5), n_warps = gridDim.x * wpb;
    for (int row0 = row_begin + warp_global * ROWS; row0 < row_end; row0 += n_warps * ROWS) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int col = lane + 32 * k;
            if (col >= L4) continue;
            float4 part[ROWS][NR], pp[ROWS], mm[ROWS], vv[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const bool okr = row0 + j < row_end;
                const size_t i = (size_t)(okr ? row0 + j : row0) * L4 + col;
#pragma unroll
                for (int r = 0; r < NR; ++r) part[j][r] = (okr && r < n_ranks) ? __ldcv(p.dW[r] + i) : z4;  // never a stale line
            }
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const size_t i = (size_t)(row0 + j < row_end ? row0 + j : row0) * L4 + col;
                pp[j] = p.W[self][i];
                mm[j] = m[i];
                vv[j] = v[i];
            }
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                if (row0 + j >= row_end) continue;
                const size_t i = (size_t)(row0 + j) * L4 + col;
                float4 g = z4;
#pragma unroll
                for (int r = 0; r < NR; ++r)  // rank order: the one owner of a row fixes the summation order for everybody
                    if (r < n_ranks) { g.x += part[j][r].x; g.y += part[j][r].y; g.z += part[j][r].z; g.w += part[j][r].w; }
#define ADAM1(x)                                                   \
    {                                                              \
        const float gr = g.x * inv_n;                              \
        mm[j].x = b1 * mm[j].x + (1.f - b1) * gr;                  \
        vv[j].x = b2 * vv[j].x + (1.f - b2) * (gr * gr);           \
        pp[j].x = pp[j].x - lr_t * mm[j].x / (sqrtf(vv[j].x) + eps); \
    }
                ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
                m[i] = mm[j];
                v[i] = vv[j];
#pragma unroll
                for (int r = 0; r < NR; ++r)
                    if (r < n_ranks) p.W[r][i] = pp[j];
            }
        }
    }
}

// Owner pass of the PUSH exchange (spmm.cu: PushTarget): the gradient rows of the W1 rows this rank owns have been
// written by every rank into this rank's LOCAL slot buffer [n_ranks][per][L1] (valid[r][row] == stamp iff rank r's batch
// touched the row).  For each owned row: sum the valid slots in rank order, average, TF-Adam with the local m, v, and
// replicate the new weight row into every rank's W1 -- peer stores, or ONE multimem.st through the NVSwitch multicast
// mapping (mc_W != NULL).  All loads are local; the only NVLink traffic of this kernel is posted stores.
template <int NCH, int NR, int ROWS>
__global__ void __launch_bounds__(NV_THREADS)
w1_slots_reduce_adam_kernel(const float4* __restrict__ slots, const uint32_t* __restrict__ valid, const uint32_t* __restrict__ epoch,
                            PeerPtrs p, float4* __restrict__ mc_W, int n_ranks, int self, int L4, int per, int row_begin, int row_end,
                            float4* __restrict__ m, float4* __restrict__ v, const float* __restrict__ beta_pow, float lr, float b1, float b2,
                            float eps) {
    const float b1p = __ldg(beta_pow), b2p = __ldg(beta_pow + 1);
    const float lr_t = lr * sqrtf(1.f - b2p) / (1.f - b1p);
    const float inv_n = 1.0f / (float)n_ranks;
    const uint32_t stamp = __ldg(epoch) + 1u;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int warp_global = blockIdx.x * wpb + (threadIdx.x >> 5), n_warps = gridDim.x * wpb;
    for (int row0 = row_begin + warp_global * ROWS; row0 < row_end; row0 += n_warps * ROWS) {
        bool ok[ROWS][NR];
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
#pragma unroll
            for (int r = 0; r < NR; ++r)
                ok[j][r] = (row0 + j < row_end) && r < n_ranks && __ldcv(valid + (size_t)r * per + (row0 + j - row_begin)) == stamp;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int col = lane + 32 * k;
            if (col >= L4) continue;
            float4 part[ROWS][NR], pp[ROWS], mm[ROWS], vv[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j)
#pragma unroll
                for (int r = 0; r < NR; ++r)
                    part[j][r] = ok[j][r] ? __ldcv(slots + ((size_t)r * per + (row0 + j - row_begin)) * L4 + col) : z4;
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const size_t i = (size_t)(row0 + j < row_end ? row0 + j : row0) * L4 + col;
                pp[j] = p.W[self][i];
                mm[j] = m[i];
                vv[j] = v[i];
            }
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                if (row0 + j >= row_end) continue;
                const size_t i = (size_t)(row0 + j) * L4 + col;
                float4 g = z4;
#pragma unroll
                for (int r = 0; r < NR; ++r)  // rank order; absent slots contribute exact zeros
                    if (r < n_ranks) { g.x += part[j][r].x; g.y += part[j][r].y; g.z += part[j][r].z; g.w += part[j][r].w; }
#define ADAM1(x)                                                   \
    {                                                              \
        const float gr = g.x * inv_n;                              \
        mm[j].x = b1 * mm[j].x + (1.f - b1) * gr;                  \
        vv[j].x = b2 * vv[j].x + (1.f - b2) * (gr * gr);           \
        pp[j].x = pp[j].x - lr_t * mm[j].x / (sqrtf(vv[j].x) + eps); \
    }
                ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
                m[i] = mm[j];
                v[i] = vv[j];
                if (mc_W) {
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_W + i), "f"(pp[j].x), "f"(pp[j].y),
                                 "f"(pp[j].z), "f"(pp[j].w)
                                 : "memory");
                } else {
#pragma unroll
                    for (int r = 0; r < NR; ++r)
                        if (r < n_ranks) p.W[r][i] = pp[j];
                }
            }
        }
    }
}

template <int NCH>
static void launch_slots(const float* slots, const uint32_t* valid, const uint32_t* epoch, const PeerPtrs& p, float* mc_W, int n_ranks, int self,
                         int L1, int per, int row_begin, int row_end, float* m, float* v, const float* beta_pow, float lr, float b1, float b2,
                         float eps, cudaStream_t st) {
    int blocks = sm_count() * 4;
    const int need = cdiv(row_end - row_begin, NV_THREADS / 32);
    if (blocks > need) blocks = need;
#define SL_ARGS (const float4*)slots, valid, epoch, p, (float4*)mc_W, n_ranks, self, L1 / 4, per, row_begin, row_end, (float4*)m, (float4*)v, \
                beta_pow, lr, b1, b2, eps
    if (n_ranks <= 2) w1_slots_reduce_adam_kernel<NCH, 2, 4><<<blocks, NV_THREADS, 0, st>>>(SL_ARGS);
    else if (n_ranks <= 4) w1_slots_reduce_adam_kernel<NCH, 4, 2><<<blocks, NV_THREADS, 0, st>>>(SL_ARGS);
    else if (n_ranks <= 8) w1_slots_reduce_adam_kernel<NCH, 8, 1><<<blocks, NV_THREADS, 0, st>>>(SL_ARGS);
    else w1_slots_reduce_adam_kernel<NCH, DSSM_MAX_PEERS, 1><<<blocks, NV_THREADS, 0, st>>>(SL_ARGS);
#undef SL_ARGS
}

// NVLS variant: the buffers are also mapped through an NVSwitch MULTICAST address.  multimem.ld_reduce makes the switch
// fetch the row from every replica and return the fp32 sum (one response instead of n), multimem.st makes it replicate
// the new weight row into every replica (one store instead of n): per GPU and direction the wire carries
// (n-1)/n * |W1| + |W1|/n instead of 2 (n-1)/n * |W1|.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(mc)
                 : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st(float4* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int NCH, int ROWS>
__global__ void __launch_bounds__(NV_THREADS)
w1_shard_reduce_adam_mc_kernel(const float4* __restrict__ mc_dW, float4* __restrict__ mc_W, const float4* __restrict__ W_local, float inv_n,
                               int L4, int row_begin, int row_end, float4* __restrict__ m, float4* __restrict__ v,
                               const float* __restrict__ beta_pow, float lr, float b1, float b2, float eps) {
    const float b1p = __ldg(beta_pow), b2p = __ldg(beta_pow + 1);
    const float lr_t = lr * sqrtf(1.f - b2p) / (1.f - b1p);
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int warp_global = blockIdx.x * wpb + (threadIdx.x >> 5), n_warps = gridDim.x * wpb;
    for (int row0 = row_begin + warp_global * ROWS; row0 < row_end; row0 += n_warps * ROWS) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int col = lane + 32 * k;
            if (col >= L4) continue;
            float4 g[ROWS], pp[ROWS], mm[ROWS], vv[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const size_t i = (size_t)(row0 + j < row_end ? row0 + j : row0) * L4 + col;
                g[j] = multimem_ld_reduce_add(mc_dW + i);
            }
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const size_t i = (size_t)(row0 + j < row_end ? row0 + j : row0) * L4 + col;
                pp[j] = W_local[i];
                mm[j] = m[i];
                vv[j] = v[i];
            }
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                if (row0 + j >= row_end) continue;
                const size_t i = (size_t)(row0 + j) * L4 + col;
#define ADAM1(x)                                                   \
    {                                                              \
        const float gr = g[j].x * inv_n;                           \
        mm[j].x = b1 * mm[j].x + (1.f - b1) * gr;                  \
        vv[j].x = b2 * vv[j].x + (1.f - b2) * (gr * gr);           \
        pp[j].x = pp[j].x - lr_t * mm[j].x / (sqrtf(vv[j].x) + eps); \
    }
                ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
                m[i] = mm[j];
                v[i] = vv[j];
                multimem_st(mc_W + i, pp[j]);
            }
        }
    }
}

template <int NCH>
static void launch_shard_mc(const float* mc_dW, float* mc_W, const float* W_local, int n_ranks, int L1, int row_begin, int row_end, float* m,
                            float* v, const float* beta_pow, float lr, float b1, float b2, float eps, cudaStream_t st) {
    int blocks = sm_count() * 4;
    const int need = cdiv(row_end - row_begin, NV_THREADS / 32);
    if (blocks > need) blocks = need;
    w1_shard_reduce_adam_mc_kernel<NCH, 4><<<blocks, NV_THREADS, 0, st>>>((const float4*)mc_dW, (float4*)mc_W, (const float4*)W_local,
                                                                          1.0f / (float)n_ranks, L1 / 4, row_begin, row_end, (float4*)m,
                                                                          (float4*)v, beta_pow, lr, b1, b2, eps);
}

template <int NCH>
static void launch_shard(const PeerPtrs& p, int n_ranks, int self, int L1, int row_begin, int row_end, float* m, float* v,
                         const float* beta_pow, float lr, float b1, float b2, float eps, cudaStream_t st) {
    int blocks = sm_count() * 4;
    const int need = cdiv(row_end - row_begin, NV_THREADS / 32);
    if (blocks > need) blocks = need;
#define NV_ARGS p, n_ranks, self, L1 / 4, row_begin, row_end, (float4*)m, (float4*)v, beta_pow, lr, b1, b2, eps
    if (n_ranks <= 2) w1_shard_reduce_adam_kernel<NCH, 2, 4><<<blocks, NV_THREADS, 0, st>>>(NV_ARGS);
    else if (n_ranks <= 4) w1_shard_reduce_adam_kernel<NCH, 4, 2><<<blocks, NV_THREADS, 0, st>>>(NV_ARGS);
    else if (n_ranks <= 8) w1_shard_reduce_adam_kernel<NCH, 8, 1><<<blocks, NV_THREADS, 0, st>>>(NV_ARGS);
    else w1_shard_reduce_adam_kernel<NCH, DSSM_MAX_PEERS, 1><<<blocks, NV_THREADS, 0, st>>>(NV_ARGS);
#undef NV_ARGS
}


// ---------------------------------------------------------------------------------------------------------------------
// SyncBN: global-batch BatchNorm moments under data parallelism (new_dssm.py:77 computes tf.nn.moments over the WHOLE
// batch; n replicas with per-replica moments are a different model).  Every rank owns a small peer-mapped exchange
// buffer:   [ flags u32[DSSM_MAX_PEERS] | epoch u32 | pad to 512 B | slots[point][rank][4*Lmax] floats ]
// One single-CTA kernel per BN instance pair and direction: PUSH this rank's [2][L] statistics into slot[point][self] of
// every rank, cross-GPU barrier (flag[self] = epoch in every rank's buffer with st.release.sys; spin on the own flags
// with ld.acquire.sys), then merge the n slots locally in rank order -- every rank computes identical bits.
//   forward   mean = avg_r mean_r ;  var = avg_r (var_r + (mean_r - mean)^2)          (equal row counts: Chan's merge)
//   backward  dbeta, dgamma = avg_r of the local column sums (each replica's loss is divided by its LOCAL query_BS, so
//             the average restores the global-batch sums; oracle/syncbn.py states the contract)
// A slot region is written again only one full step later, and the other exchange points of the step lie in between, so
// no reader can still be on it.
constexpr int SYNCBN_HEADER_BYTES = 512;
constexpr int SYNCBN_THREADS = 512;

struct SyncBnPeers {
    char* buf[DSSM_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all threads of the (single) CTA call this after their pushes
__device__ __forceinline__ void syncbn_barrier(const SyncBnPeers& p, int n, int self) {
    __shared__ uint32_t s_epoch;
    __threadfence_system();  // this thread's pushes are visible system-wide before the flag goes up
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t* ep = reinterpret_cast<uint32_t*>(p.buf[self] + DSSM_MAX_PEERS * sizeof(uint32_t));
        s_epoch = *ep + 1u;
        *ep = s_epoch;
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    if ((int)threadIdx.x < n) {
        st_release_sys(reinterpret_cast<uint32_t*>(p.buf[threadIdx.x]) + self, epoch);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(p.buf[self]) + threadIdx.x;
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float* syncbn_slot(const SyncBnPeers& p, int owner, int point, int n, int rank, int slot_floats) {
    return reinterpret_cast<float*>(p.buf[owner] + SYNCBN_HEADER_BYTES) + ((size_t)point * n + rank) * slot_floats;
}

__global__ void __launch_bounds__(SYNCBN_THREADS)
syncbn_fwd_kernel(SyncBnPeers p, int n, int self, int point, int slot_floats, int L, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ ema_mean, float* __restrict__ ema_var, float* __restrict__ mean,
                  float* __restrict__ var, float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift, float eps,
                  float decay, int update_ema) {
    const int L2 = 2 * L;
    for (int i = threadIdx.x; i < L2; i += blockDim.x) {
        const float m = mean[i], v = var[i];
        for (int r = 0; r < n; ++r) {
            float* dst = syncbn_slot(p, r, point, n, self, slot_floats);
            dst[i] = m;
            dst[L2 + i] = v;
        }
    }
    syncbn_barrier(p, n, self);
    const float inv_n = 1.0f / (float)n;
    for (int i = threadIdx.x; i < L2; i += blockDim.x) {
        float mu = 0.f;
        for (int r = 0; r < n; ++r) mu += __ldcv(syncbn_slot(p, self, point, n, r, slot_floats) + i);
        mu *= inv_n;
        float vv = 0.f;
        for (int r = 0; r < n; ++r) {
            const float* s = syncbn_slot(p, self, point, n, r, slot_floats);
            const float d = __ldcv(s + i) - mu;
            vv += __ldcv(s + L2 + i) + d * d;
        }
        vv *= inv_n;
        if (update_ema) {  // the shadows follow the GLOBAL moments: identical on every replica
            const float em = ema_mean[i], ev = ema_var[i];
            ema_mean[i] = em - (1.f - decay) * (em - mu);
            ema_var[i] = ev - (1.f - decay) * (ev - vv);
        }
        const float rs = 1.0f / sqrtf(vv + eps);
        const float sc = rs * gamma[i];
        mean[i] = mu;
        var[i] = vv;
        rstd[i] = rs;
        scale[i] = sc;
        shift[i] = beta[i] - mu * sc;
    }
}

__global__ void __launch_bounds__(SYNCBN_THREADS)
syncbn_bwd_kernel(SyncBnPeers p, int n, int self, int point, int slot_floats, int L, int n_q, int n_d, const float* __restrict__ gamma,
                  const float* __restrict__ rstd, const float* __restrict__ sumx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                  float* __restrict__ db) {
    __shared__ float s_db[2 * 1024];
    const int L2 = 2 * L;
    for (int i = threadIdx.x; i < L2; i += blockDim.x) {
        const float b = dbeta[i], g = dgamma[i];
        for (int r = 0; r < n; ++r) {
            float* dst = syncbn_slot(p, r, point, n, self, slot_floats);
            dst[i] = b;
            dst[L2 + i] = g;
        }
    }
    syncbn_barrier(p, n, self);
    const float inv_n = 1.0f / (float)n;
    for (int i = threadIdx.x; i < L2; i += blockDim.x) {
        float b = 0.f, g = 0.f;
        for (int r = 0; r < n; ++r) {
            const float* s = syncbn_slot(p, self, point, n, r, slot_floats);
            b += __ldcv(s + i);
            g += __ldcv(s + L2 + i);
        }
        b *= inv_n;
        g *= inv_n;
        const float b_loc = dbeta[i];
        const float rows = i < L ? (float)n_q : (float)n_d;
        // sum over the LOCAL rows of dH = gamma*rstd*(g - dbeta/n - xhat*dgamma/n): the pre-BN bias gradient of this replica
        if (i < 2 * 1024) s_db[i] = (gamma[i] * rstd[i]) * ((b_loc - b) - (sumx[i] / rows) * g);
        dbeta[i] = b;
        dgamma[i] = g;
    }
    __syncthreads();
    if (db)
        for (int c = threadIdx.x; c < L; c += blockDim.x) db[c] = s_db[c] + s_db[L + c];
}

// ---- cross-GPU flags for the chunked dW1 exchange ----------------------------------------------------------------------
// Every rank owns a peer-mapped flag block  [ flags u32[DSSM_MAX_PEERS] | epoch u32 ]  (zero-filled once).  A step's
// sync points are numbered idx = 0 .. stride-1; rank r "signals idx" by writing  epoch*stride + idx + 1  into flags[r] of
// EVERY rank's block (st.release.sys after a system-scope fence: everything this stream wrote before -- e.g. a chunk of
// dW1 -- is visible to a peer that acquires the flag); "wait idx" spins until all n flags of the own block have reached
// that value.  The epoch lives in device memory and is advanced by a kernel once per step, so the whole sequence replays
// unchanged inside a CUDA graph.  Values only grow, so a rank that is already past a sync point never confuses a waiter.
__global__ void peer_signal_kernel(SyncBnPeers p, int n, int self, int idx, int stride) {
    __threadfence_system();
    const uint32_t epoch = *reinterpret_cast<const uint32_t*>(p.buf[self] + DSSM_MAX_PEERS * sizeof(uint32_t));
    const uint32_t value = epoch * (uint32_t)stride + (uint32_t)idx + 1u;
    if ((int)threadIdx.x < n) st_release_sys(reinterpret_cast<uint32_t*>(p.buf[threadIdx.x]) + self, value);
}
__global__ void peer_wait_kernel(const uint32_t* own, int n, int idx, int stride) {
    const uint32_t epoch = own[DSSM_MAX_PEERS];
    const uint32_t value = epoch * (uint32_t)stride + (uint32_t)idx + 1u;
    if ((int)threadIdx.x < n)
        while ((int32_t)(ld_acquire_sys(own + threadIdx.x) - value) < 0) {
        }
    __syncthreads();
}
__global__ void peer_epoch_advance_kernel(uint32_t* own) { own[DSSM_MAX_PEERS] += 1u; }
// signal(idx) + wait(idx) (+ the epoch advance that ends a step) in ONE launch: the two sync points of the push exchange
// sit on the critical path of every step, and each launch saved there is ~3-4 us
__global__ void peer_barrier_kernel(SyncBnPeers p, int n, int self, int idx, int stride, int advance) {
    __threadfence_system();
    uint32_t* own = reinterpret_cast<uint32_t*>(p.buf[self]);
    const uint32_t epoch = own[DSSM_MAX_PEERS];
    const uint32_t value = epoch * (uint32_t)stride + (uint32_t)idx + 1u;
    if ((int)threadIdx.x < n) {
        st_release_sys(reinterpret_cast<uint32_t*>(p.buf[threadIdx.x]) + self, value);
        while ((int32_t)(ld_acquire_sys(own + threadIdx.x) - value) < 0) {
        }
    }
    __syncthreads();
    if (advance && threadIdx.x == 0) own[DSSM_MAX_PEERS] = epoch + 1u;
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

extern "C" int dssm_w1_shard_reduce_adam_mc(const float* mc_dW1, float* mc_W1, const float* W1_local, int32_t n_ranks, int32_t D,
                                            int32_t L1, int32_t row_begin, int32_t row_end, float* m1, float* v1,
                                            const float* beta_pow, float lr, float beta1, float beta2, float eps, dssm_stream_t stream) {
    DSSM_REQUIRE(mc_dW1 && mc_W1 && W1_local && m1 && v1 && beta_pow, DSSM_ERR_BAD_ARG, "dssm_w1_shard_reduce_adam_mc: null pointer");
    DSSM_REQUIRE(n_ranks >= 1, DSSM_ERR_BAD_ARG, "dssm_w1_shard_reduce_adam_mc: n_ranks=%d", n_ranks);
    DSSM_REQUIRE(D > 0 && L1 > 0 && L1 % 4 == 0 && L1 <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_w1_shard_reduce_adam_mc: bad shape");
    DSSM_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= D, DSSM_ERR_BAD_ARG, "dssm_w1_shard_reduce_adam_mc: bad row range");
    DSSM_REQUIRE(aligned16(mc_dW1) && aligned16(mc_W1) && aligned16(W1_local) && aligned16(m1) && aligned16(v1), DSSM_ERR_BAD_ALIGN,
                 "dssm_w1_shard_reduce_adam_mc: buffers must be 16-byte aligned");
    if (row_begin == row_end) return DSSM_OK;
    const int nch = cdiv(L1 / 4, 32);
    DISPATCH_NCH(nch, launch_shard_mc<N_>(mc_dW1, mc_W1, W1_local, n_ranks, L1, row_begin, row_end, m1, v1, beta_pow, lr, beta1, beta2, eps,
                                          (cudaStream_t)stream));
    LAUNCH_CHECK("w1_shard_reduce_adam_mc");
    return DSSM_OK;
}


// ---- SyncBN exchange (internal: called by the tower, tower.cu) --------------------------------------------------------
extern "C" size_t dssm_syncbn_buffer_bytes(int32_t n_ranks, int32_t n_points, int32_t Lmax) {
    return align_up((size_t)SYNCBN_HEADER_BYTES + (size_t)n_points * n_ranks * 4 * Lmax * sizeof(float), 256);
}

static int syncbn_peers(void* const* peer_bufs, int n_ranks, int self, SyncBnPeers* out) {
    DSSM_REQUIRE(peer_bufs && n_ranks >= 1 && n_ranks <= DSSM_MAX_PEERS && self >= 0 && self < n_ranks, DSSM_ERR_BAD_ARG,
                 "syncbn: n_ranks=%d self=%d (at most %d peers)", n_ranks, self, DSSM_MAX_PEERS);
    for (int r = 0; r < n_ranks; ++r) {
        DSSM_REQUIRE(peer_bufs[r] && aligned16(peer_bufs[r]), DSSM_ERR_BAD_ALIGN, "syncbn: peer buffer %d null or unaligned", r);
        out->buf[r] = (char*)peer_bufs[r];
    }
    return DSSM_OK;
}

extern "C" int dssm_syncbn_forward(void* const* peer_bufs, int32_t n_ranks, int32_t self, int32_t point, int32_t Lmax, int32_t L,
                                   const float* gamma, const float* beta, float* ema_mean, float* ema_var, float* mean, float* var,
                                   float* rstd, float* scale, float* shift, float eps, float decay, int32_t update_ema,
                                   dssm_stream_t stream) {
    SyncBnPeers p{};
    int rc = syncbn_peers(peer_bufs, n_ranks, self, &p);
    if (rc != DSSM_OK) return rc;
    DSSM_REQUIRE(L > 0 && L <= Lmax, DSSM_ERR_BAD_SHAPE, "syncbn: L=%d exceeds Lmax=%d", L, Lmax);
    syncbn_fwd_kernel<<<1, SYNCBN_THREADS, 0, (cudaStream_t)stream>>>(p, n_ranks, self, point, 4 * Lmax, L, gamma, beta, ema_mean, ema_var, mean,
                                                                     var, rstd, scale, shift, eps, decay, update_ema);
    LAUNCH_CHECK("syncbn_fwd");
    return DSSM_OK;
}

extern "C" int dssm_syncbn_backward(void* const* peer_bufs, int32_t n_ranks, int32_t self, int32_t point, int32_t Lmax, int32_t L,
                                    int32_t n_q, int32_t n_d, const float* gamma, const float* rstd, const float* sumx, float* dgamma,
                                    float* dbeta, float* db, dssm_stream_t stream) {
    SyncBnPeers p{};
    int rc = syncbn_peers(peer_bufs, n_ranks, self, &p);
    if (rc != DSSM_OK) return rc;
    DSSM_REQUIRE(L > 0 && L <= Lmax && L <= 1024, DSSM_ERR_BAD_SHAPE, "syncbn: L=%d exceeds min(Lmax=%d, 1024)", L, Lmax);
    syncbn_bwd_kernel<<<1, SYNCBN_THREADS, 0, (cudaStream_t)stream>>>(p, n_ranks, self, point, 4 * Lmax, L, n_q, n_d, gamma, rstd, sumx, dgamma,
                                                                     dbeta, db);
    LAUNCH_CHECK("syncbn_bwd");
    return DSSM_OK;
}

// ---- flag synchronisation of the chunked exchange (public: include/dssm_b200.h) -----------------------------------------
extern "C" size_t dssm_peer_flags_bytes(void) { return 256; }

extern "C" int dssm_peer_signal(void* const* host_peer_flags, int32_t n_ranks, int32_t self, int32_t idx, int32_t stride, dssm_stream_t stream) {
    SyncBnPeers p{};
    int rc = syncbn_peers(host_peer_flags, n_ranks, self, &p);
    if (rc != DSSM_OK) return rc;
    DSSM_REQUIRE(idx >= 0 && idx < stride, DSSM_ERR_BAD_ARG, "dssm_peer_signal: idx=%d outside [0,%d)", idx, stride);
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, n_ranks, self, idx, stride);
    LAUNCH_CHECK("peer_signal");
    return DSSM_OK;
}

extern "C" int dssm_peer_wait(const void* own_flags, int32_t n_ranks, int32_t idx, int32_t stride, dssm_stream_t stream) {
    DSSM_REQUIRE(own_flags && n_ranks >= 1 && n_ranks <= DSSM_MAX_PEERS, DSSM_ERR_BAD_ARG, "dssm_peer_wait: bad arguments");
    DSSM_REQUIRE(idx >= 0 && idx < stride, DSSM_ERR_BAD_ARG, "dssm_peer_wait: idx=%d outside [0,%d)", idx, stride);
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const uint32_t*)own_flags, n_ranks, idx, stride);
    LAUNCH_CHECK("peer_wait");
    return DSSM_OK;
}

extern "C" int dssm_peer_barrier(void* const* host_peer_flags, int32_t n_ranks, int32_t self, int32_t idx, int32_t stride,
                                 int32_t advance_epoch, dssm_stream_t stream) {
    SyncBnPeers p{};
    int rc = syncbn_peers(host_peer_flags, n_ranks, self, &p);
    if (rc != DSSM_OK) return rc;
    DSSM_REQUIRE(idx >= 0 && idx < stride, DSSM_ERR_BAD_ARG, "dssm_peer_barrier: idx=%d outside [0,%d)", idx, stride);
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, n_ranks, self, idx, stride, advance_epoch);
    LAUNCH_CHECK("peer_barrier");
    return DSSM_OK;
}

extern "C" int dssm_peer_epoch_advance(void* own_flags, dssm_stream_t stream) {
    DSSM_REQUIRE(own_flags, DSSM_ERR_BAD_ARG, "dssm_peer_epoch_advance: null pointer");
    peer_epoch_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((uint32_t*)own_flags);
    LAUNCH_CHECK("peer_epoch_advance");
    return DSSM_OK;
}


// ---- owner pass of the push exchange (public: include/dssm_b200.h) ----------------------------------------------------
extern "C" int dssm_w1_slots_reduce_adam(const float* slots, const uint32_t* valid, const uint32_t* epoch, float* const* peer_W1,
                                         float* mc_W1, int32_t n_ranks, int32_t self, int32_t D, int32_t L1, int32_t per, float* m1,
                                         float* v1, const float* beta_pow, float lr, float beta1, float beta2, float eps,
                                         dssm_stream_t stream) {
    DSSM_REQUIRE(slots && valid && epoch && peer_W1 && m1 && v1 && beta_pow, DSSM_ERR_BAD_ARG, "dssm_w1_slots_reduce_adam: null pointer");
    DSSM_REQUIRE(n_ranks >= 1 && n_ranks <= DSSM_MAX_PEERS && self >= 0 && self < n_ranks, DSSM_ERR_BAD_ARG,
                 "dssm_w1_slots_reduce_adam: n_ranks=%d self=%d (at most %d peers)", n_ranks, self, DSSM_MAX_PEERS);
    DSSM_REQUIRE(D > 0 && L1 > 0 && L1 % 4 == 0 && L1 <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_w1_slots_reduce_adam: bad shape");
    DSSM_REQUIRE(per > 0 && (int64_t)per * n_ranks >= D, DSSM_ERR_BAD_ARG, "dssm_w1_slots_reduce_adam: per=%d does not cover D=%d", per, D);
    PeerPtrs p{};
    for (int r = 0; r < n_ranks; ++r) {
        DSSM_REQUIRE(peer_W1[r] && aligned16(peer_W1[r]), DSSM_ERR_BAD_ALIGN, "dssm_w1_slots_reduce_adam: peer W1 %d null or unaligned", r);
        p.W[r] = (float4*)peer_W1[r];
    }
    DSSM_REQUIRE(aligned16(slots) && aligned16(m1) && aligned16(v1) && (!mc_W1 || aligned16(mc_W1)), DSSM_ERR_BAD_ALIGN,
                 "dssm_w1_slots_reduce_adam: buffers must be 16-byte aligned");
    const int row_begin = self * per < D ? self * per : D;
    const int row_end = row_begin + per < D ? row_begin + per : D;
    if (row_begin == row_end) return DSSM_OK;
    const int nch = cdiv(L1 / 4, 32);
    DISPATCH_NCH(nch, launch_slots<N_>(slots, valid, epoch, p, mc_W1, n_ranks, self, L1, per, row_begin, row_end, m1, v1, beta_pow, lr, beta1,
                                       beta2, eps, (cudaStream_t)stream));
    LAUNCH_CHECK("w1_slots_reduce_adam");
    return DSSM_OK;
}
