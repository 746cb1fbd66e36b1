// Dense-layer contractions on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM)
// with error-compensated operands ("3xTF32"): every fp32 operand x is split into hi = rna_tf32(x) and
// lo = rna_tf32(x - hi) and the product is accumulated as a_hi*b_lo + a_lo*b_hi + a_hi*b_hi in fp32.  The
// dropped a_lo*b_lo term and the rounding of lo are O(2^-22) relative, so this path keeps the 1e-5 parity
// bar of the fp32 mode while the FFMA work moves to the tensor pipe.
//
//   D[M,N] = pro(A)[M,K] . B[N,K]^T (+ bias[n])       A and B both K-major (row-major, K contiguous)
//
// pro() is the previous layer's BN scale/shift + activation (new_dssm.py:87,134-136), applied in registers
// while the A tile is staged, so the normalised tensor never exists in HBM.  Used for
//   forward   Hout = pro(Hprev) . W        with B = W^T
//   backward  dA   = dH . W^T              with B = W ([K_layer, N_layer] is K-major for this product)
// In both, B is the (small) weight matrix: it is split and swizzled ONCE per call into a tile image
// (make_b_image_kernel) that each CTA pulls into shared memory with cp.async.bulk + mbarrier complete_tx, so the
// threads only stage the activation operand.
//
// Kernels in this file (one 128 x BN output tile per CTA, BK = 32 fp32 per k-block, 3 stages):
//   gemm_tc3_tma_kernel  K-major, default: raw activation tiles by TMA (cp.async.bulk.tensor.2d), producers read their row
//                        from shared memory, A operand written into TENSOR MEMORY (tcgen05.st), TS-form MMAs
//   gemm_tc3_ws_kernel   K-major, SS form: warps 0-7 stage A (global -> registers -> split -> swizzled smem), warp 8 issues
//                        the MMAs, warp 9 streams the weight image; fallback (K > 384, fused BN moments, DSSM_GEMM_TS=0)
//   gemm_tc3_kernel<MN>  MN = true: the dW contraction, both operands MN-major as they lie in memory, reduction over the
//                        rows split over blockIdx.z; MN = false: K-major without a weight image (generic entry points).
//                        16 warps stage both operands, thread 0 issues the MMAs after a CTA barrier
// In all of them tcgen05.commit releases a stage through an mbarrier and the epilogue goes TMEM -> registers -> shared
// memory -> full rows to global.
#include "common.cuh"
#include "tc_common.cuh"
#include "bn_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace dssm {
namespace tc {

constexpr int BM = 128, BK = 32, STAGES = 3;
// 16 warps for gemm_tc3_kernel: ncu on the dW contraction at C3 (profiles/r2_ncu_dw_c3.txt) showed 8 warps -- two per
// scheduler -- issuing 30 % of the time on dependent split/convert chains (stalls: wait 25 %, long scoreboard 23 %); with
// four warps per scheduler a thread stages half as many chunks and the schedulers have twice the warps to pick from.
constexpr int THREADS = 512;
constexpr int A_PER = BM * BK / 4 / THREADS;                        // 16-byte A chunks per thread and k-block (2)
constexpr int B_PER = (160 * BK / 4 + THREADS - 1) / THREADS;       // B chunks per thread at the widest tile (3, guarded)
constexpr int MAX_BN = 160;
constexpr int TMEM_COLS = 512;
// The tensor core adds into its fp32 accumulator with truncation, so a long chain of MMAs into ONE accumulator
// drifts by ~0.5 ulp(acc) per step (measured: 1.5e-5 relative on a 2000-row reduction).  k-blocks therefore go
// round-robin into NACC accumulators (3 x 160 columns of the 512) that the epilogue adds in registers (RN).
constexpr int NACC = 3;
constexpr int A_TILE_BYTES = BM * BK * 4;  // 16 KB

// MN-major operand tile (the MN index is the contiguous one in memory).  For 32-bit operands the only MN-major
// layout the tensor core accepts is SWIZZLE_128B_BASE32B (cutlass sm100_common.inl: "for mn-major tf32 operands,
// SW128_32B is the only available smem layout"): a 128-byte line holds 32 consecutive MN elements of one k, 4
// consecutive k lines form the 512-byte swizzle atom in which the 32-byte granule index is XORed with (k & 3)
// (Swizzle<2,5,2> on byte addresses); the next 32 MN elements are LBO bytes away, the next 4 k are SBO bytes away.
// Tiles here are laid out [mn_block][k = 0..BK-1][128 B], so SBO = 512 and LBO = BK*128.
constexpr int MN_LBO = BK * 128;
constexpr int MN_SBO = 512;
constexpr int MN_K8_BYTES = 8 * 128;  // one tf32 MMA consumes 8 k lines
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(MN_LBO >> 4) << 16;
    d |= (uint64_t)(MN_SBO >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
    return d;
}

struct Args {
    const float* A;   // [M, K] row-major (lda = K)
    const float* Bm;  // [N, K] row-major (ldb = ldb)
    float* D;         // [M, N] row-major
    const float* bias;
    const float* scale;
    const float* shift;  // [2][K] or NULL
    int M, N, K, ldb, act, Bseg, BN;
    int k_per_split;  // MN mode: rows of the reduction handled by one blockIdx.z (multiple of BK)
    const char* Bimg;  // K-major mode: pre-split, pre-swizzled image of B (make_b_image_kernel); NULL = stage B in registers
    int passes;        // 3: error-compensated 3xTF32 (1e-5 parity); 1: one tf32 MMA per product (DSSM_GEMM_TC_TF32)
    FusedBnStats fbn;  // ws kernel: column moments of D taken in the epilogue (fbn.on)
};

// write one float4 (hi and lo parts) into a swizzled K-major tile: row r, 16-byte chunk c
__device__ __forceinline__ void stage_chunk(char* tile_hi, char* tile_lo, int r, int c, float4 v) {
    float4 hi, lo;
    hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
    lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y); lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
    const int off = r * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<float4*>(tile_hi + off) = hi;
    *reinterpret_cast<float4*>(tile_lo + off) = lo;
}

// MN-major tile: reduction index kk (0..BK-1), 16-byte chunk c along MN
__device__ __forceinline__ void stage_chunk_mn(char* tile_hi, char* tile_lo, int kk, int c, float4 v) {
    float4 hi, lo;
    hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
    lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y); lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
    const int off = (c >> 3) * MN_LBO + kk * 128 + (((((c & 7) >> 1) ^ (kk & 3))) << 5) + ((c & 1) << 4);
    *reinterpret_cast<float4*>(tile_hi + off) = hi;
    *reinterpret_cast<float4*>(tile_lo + off) = lo;
}

// MN = false:  D[M,N] = pro(A)[M,K] . B[N,K]^T (+bias)           A, B row-major with K contiguous
// MN = true :  D[M,N] = sum_r pro(A)[r,M]^T . B[r,N]  over the rows r of this split (A = [R,M], B = [R,N] row-major)
template <bool MN>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc3_kernel(Args g) {
    extern __shared__ char smem_raw[];
    __shared__ uint64_t empty_bar[STAGES];
    __shared__ uint64_t full_bar[STAGES];  // B tile landed (bulk async copy), K-major mode with a B image
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_base_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * g.BN;
    const int BN = g.BN;
    const int b_tile_bytes = BN * BK * 4;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
    char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&empty_bar[s], 1);
            mbar_init(&full_bar[s], 1);
        }
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t idesc = make_idesc_tf32(BM, BN, MN);
    int kbeg = 0, kend = g.K;
    if (MN) {
        kbeg = blockIdx.z * g.k_per_split;
        kend = min(g.K, kbeg + g.k_per_split);
    }
    const int nkb = max(1, (kend - kbeg + BK - 1) / BK);  // at least one (zero-filled) block so the accumulator is defined
    // short reductions (K <= 384: the forward and dX contractions) stay well inside the 1e-5 bar with one
    // accumulator (120 chained MMAs); only long ones (dW over thousands of rows) need the rotation
    const int nacc_used = nkb <= 12 ? 1 : NACC;
    const int bchunks = BN / 4;                            // 16-byte chunks per B row in MN mode
    const bool use_img = g.Bimg != nullptr;
    // image layout: [n_tile][k_block][hi | lo][BN rows x 128 B, SW128]
    const char* b_img = use_img ? g.Bimg + (size_t)blockIdx.x * nkb * 2 * b_tile_bytes : nullptr;

    // ---- operand staging, software-pipelined one k-block ahead in registers --------------------------------
    struct Regs {
        float4 a[A_PER], b[B_PER], sc[A_PER], sh[A_PER];
    };
    auto load_regs = [&](int kb, Regs& q) {
        const int k0 = kbeg + kb * BK;
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int id = tid + i * THREADS;
            q.a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            q.sc[i] = make_float4(1.f, 1.f, 1.f, 1.f);
            q.sh[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!MN) {
                const int r = id >> 3, c = id & 7;
                const int m = m0 + r, k = k0 + c * 4;
                if (m < g.M && k < g.K) {
                    q.a[i] = __ldg(reinterpret_cast<const float4*>(g.A + (size_t)m * g.K + k));
                    if (g.scale) {
                        const int o = (m < g.Bseg ? 0 : g.K) + k;
                        q.sc[i] = __ldg(reinterpret_cast<const float4*>(g.scale + o));
                        q.sh[i] = __ldg(reinterpret_cast<const float4*>(g.shift + o));
                    }
                }
            } else {
                const int kk = id >> 5, c = id & 31;  // 32 reduction rows x 32 chunks of 4 features
                const int r = k0 + kk, m = m0 + c * 4;
                if (r < kend && m < g.M) {
                    q.a[i] = __ldg(reinterpret_cast<const float4*>(g.A + (size_t)r * g.M + m));
                    if (g.scale) {
                        const int o = (r < g.Bseg ? 0 : g.M) + m;
                        q.sc[i] = __ldg(reinterpret_cast<const float4*>(g.scale + o));
                        q.sh[i] = __ldg(reinterpret_cast<const float4*>(g.shift + o));
                    }
                }
            }
        }
        if (!MN && use_img) return;
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int id = tid + i * THREADS;
            q.b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!MN) {
                const int r = id >> 3, c = id & 7;
                const int n = n0 + r, k = k0 + c * 4;
                if (r < BN && n < g.N && k < g.K) q.b[i] = __ldg(reinterpret_cast<const float4*>(g.Bm + (size_t)n * g.ldb + k));
            } else {
                const int kk = id / bchunks, c = id - kk * bchunks;
                const int r = k0 + kk, n = n0 + c * 4;
                if (kk < BK && r < kend && n < g.N) q.b[i] = __ldg(reinterpret_cast<const float4*>(g.Bm + (size_t)r * g.N + n));
            }
        }
    };
    auto process = [&](int kb, const Regs& q) {
        const int st = kb % STAGES, use = kb / STAGES;
        char* a_hi = smem + st * stage_bytes;
        char* a_lo = a_hi + A_TILE_BYTES;
        char* b_hi = a_lo + A_TILE_BYTES;
        char* b_lo = b_hi + b_tile_bytes;
        const int k0 = kbeg + kb * BK;
        // the stage must have been drained by the MMAs that last read it
        if (use > 0) mbar_wait(&empty_bar[st], (uint32_t)((use - 1) & 1));
        // prologue (BN scale/shift + activation) + hi/lo split + swizzled store
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int id = tid + i * THREADS;
            float4 v = q.a[i];
            bool valid;
            if (!MN) {
                valid = (m0 + (id >> 3) < g.M) && (k0 + (id & 7) * 4 < g.K);
            } else {
                valid = (k0 + (id >> 5) < kend) && (m0 + (id & 31) * 4 < g.M);
            }
            if (valid) {  // padding stays exactly zero (act(shift) must not leak into it)
                v.x = act_fwd(fmaf(v.x, q.sc[i].x, q.sh[i].x), g.act);
                v.y = act_fwd(fmaf(v.y, q.sc[i].y, q.sh[i].y), g.act);
                v.z = act_fwd(fmaf(v.z, q.sc[i].z, q.sh[i].z), g.act);
                v.w = act_fwd(fmaf(v.w, q.sc[i].w, q.sh[i].w), g.act);
            }
            if (!MN) stage_chunk(a_hi, a_lo, id >> 3, id & 7, v);
            else stage_chunk_mn(a_hi, a_lo, id >> 5, id & 31, v);
        }
        if (MN || !use_img) {
#pragma unroll
            for (int i = 0; i < B_PER; ++i) {
                const int id = tid + i * THREADS;
                if (!MN) {
                    const int r = id >> 3, c = id & 7;
                    if (r < BN) stage_chunk(b_hi, b_lo, r, c, q.b[i]);
                } else {
                    const int kk = id / bchunks, c = id - kk * bchunks;
                    if (kk < BK) stage_chunk_mn(b_hi, b_lo, kk, c, q.b[i]);
                }
            }
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            if (!MN && use_img) mbar_wait(&full_bar[st], (uint32_t)(use & 1));  // this block's B tile has landed
            tc_fence_after();
            const uint64_t da_hi = MN ? make_desc_mn_sw128(smem_u32(a_hi)) : make_desc_k_sw128(smem_u32(a_hi));
            const uint64_t da_lo = MN ? make_desc_mn_sw128(smem_u32(a_lo)) : make_desc_k_sw128(smem_u32(a_lo));
            const uint64_t db_hi = MN ? make_desc_mn_sw128(smem_u32(b_hi)) : make_desc_k_sw128(smem_u32(b_hi));
            const uint64_t db_lo = MN ? make_desc_mn_sw128(smem_u32(b_lo)) : make_desc_k_sw128(smem_u32(b_lo));
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                // one MMA consumes 8 tf32 along the reduction: 32 B inside the swizzled row (K-major) or eight
                // 128-byte k lines (MN-major)
                const uint64_t adv = MN ? (uint64_t)((ks * MN_K8_BYTES) >> 4) : (uint64_t)((ks * 8 * 4) >> 4);
                const uint32_t acc = tmem_d + (uint32_t)((kb % nacc_used) * MAX_BN);
                const uint32_t first = (kb >= nacc_used || ks > 0) ? 1u : 0u;
                if (g.passes == 3) {
                    mma_tf32(acc, da_hi + adv, db_lo + adv, idesc, first);
                    mma_tf32(acc, da_lo + adv, db_hi + adv, idesc, 1u);
                    mma_tf32(acc, da_hi + adv, db_hi + adv, idesc, 1u);
                } else {
                    mma_tf32(acc, da_hi + adv, db_hi + adv, idesc, first);
                }
            }
            mma_commit(&empty_bar[st]);                // stage reusable when these MMAs have read it
            if (kb == nkb - 1) mma_commit(&done_bar);  // accumulators complete
            if (!MN && use_img && kb + 2 < nkb) {      // prefetch the B tile two k-blocks ahead
                const int st2 = (kb + 2) % STAGES, use2 = (kb + 2) / STAGES;
                if (use2 > 0) mbar_wait(&empty_bar[st2], (uint32_t)((use2 - 1) & 1));
                bulk_copy_g2s(smem + st2 * stage_bytes + 2 * A_TILE_BYTES, b_img + (size_t)(kb + 2) * 2 * b_tile_bytes,
                              (uint32_t)(2 * b_tile_bytes), &full_bar[st2]);
            }
        }
    };

    if (!MN && use_img && tid == 0) {
        for (int kb = 0; kb < 2 && kb < nkb; ++kb)
            bulk_copy_g2s(smem + kb * stage_bytes + 2 * A_TILE_BYTES, b_img + (size_t)kb * 2 * b_tile_bytes,
                          (uint32_t)(2 * b_tile_bytes), &full_bar[kb]);
    }
    Regs r0, r1;
    load_regs(0, r0);
    for (int kb = 0; kb < nkb; kb += 2) {
        if (kb + 1 < nkb) load_regs(kb + 1, r1);  // in flight while block kb is transformed, stored and issued
        process(kb, r0);
        if (kb + 1 < nkb) {
            if (kb + 2 < nkb) load_regs(kb + 2, r0);
            process(kb + 1, r1);
        }
    }
    mbar_wait(&done_bar, 0);
    tc_fence_after();

    // ---- epilogue: TMEM -> registers -> shared-memory tile -> (+bias) -> global, full rows per warp (coalesced; see the
    // warp-specialised kernel below for the measurement).  Warp w owns TMEM lanes 32*(w%4)..+31; the pipeline stages are
    // free once done_bar has completed.
    const int lane_grp = warp & 3;
    const int nchunks = BN / 32;  // BN is a multiple of 32 on this path
    const int nacc = nkb < nacc_used ? nkb : nacc_used;
    float* tile = reinterpret_cast<float*>(smem);
    const int ld = BN + 4;
    for (int ch = (warp >> 2); ch < nchunks; ch += THREADS / 128) {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;  // +0.0f
        for (int a = 0; a < nacc; ++a) {
            uint32_t t[32];
            tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(a * MAX_BN + ch * 32), t);
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(t[j]));
        }
        float* dst = tile + (size_t)(lane_grp * 32 + lane) * ld + ch * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                              __uint_as_float(r[j + 3]));
    }
    tc_fence_before();
    __syncthreads();
    {
        float* Dz = g.D + (MN ? (size_t)blockIdx.z * g.M * g.N : 0);
        const int n4 = BN / 4;
        const int nvalid = min(BN, g.N - n0);  // N is a multiple of 4 on this path
        for (int c4 = lane; c4 < n4; c4 += 32) {
            if (c4 * 4 >= nvalid) continue;
            float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.bias) bz = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + c4 * 4));
            for (int rr = warp; rr < BM; rr += THREADS / 32) {
                if (m0 + rr >= g.M) break;
                float4 o = *reinterpret_cast<const float4*>(tile + (size_t)rr * ld + c4 * 4);
                o.x += bz.x; o.y += bz.y; o.z += bz.z; o.w += bz.w;
                *reinterpret_cast<float4*>(Dz + (size_t)(m0 + rr) * g.N + n0 + c4 * 4) = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Warp-specialised variant for the K-major contractions whose B operand is a pre-built weight image (forward, dX):
//   warps 0-7  producers: stage the activation tile (register ring WS_PF k-blocks ahead, BN scale/shift from shared
//              memory, hi/lo split, swizzled store), fence.proxy.async, arrive on full_a[stage]          (256 arrivals)
//   warp 8     lane 0 issues the MMAs as soon as full_a / full_b of a stage have completed, tcgen05.commit -> empty
//   warp 9     lane 0 streams the B image tiles with cp.async.bulk -> full_b (complete_tx), waiting on empty for reuse
//   warps 0-7 drain TMEM at the end.
// No CTA-wide barrier inside the main loop: producers run up to STAGES k-blocks ahead of the tensor core.
constexpr int WS_MAX_K = 512;  // scale/shift of both BN instances live in shared memory (8 KB static next to 217 KB dynamic)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int WS_PF = 4;                       // k-blocks of activation loads in flight per producer thread
constexpr int WS_PRODUCERS = 256;             // warps 0-7
constexpr int WS_THREADS = WS_PRODUCERS + 64;  // + warp 8 (MMA issuer) + warp 9 (B image loader)

template <int ACT>
__device__ __forceinline__ float act_t(float x) {
    if (ACT == DSSM_ACT_RELU) return fmaxf(x, 0.f);
    if (ACT == DSSM_ACT_TANH) return tanhf(x);
    return x;
}

template <int ACT>
__global__ void __launch_bounds__(WS_THREADS, 1) gemm_tc3_ws_kernel(Args g) {
    extern __shared__ char smem_raw[];
    __shared__ uint64_t empty_bar[STAGES];
    __shared__ uint64_t full_a[STAGES];
    __shared__ uint64_t full_b[STAGES];
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_scale[2][WS_MAX_K];
    __shared__ __align__(16) float s_shift[2][WS_MAX_K];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * g.BN;
    const int BN = g.BN;
    const int b_tile_bytes = BN * BK * 4;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
    char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int nkb = (g.K + BK - 1) / BK;
    const int kpad = nkb * BK;

    // Producers (warps 0-7): loop-invariant addressing, and the first WS_PF k-blocks of activation loads go out BEFORE the
    // CTA's setup (TMEM allocation, barrier init, scale/shift to shared memory): their L2 round trip (~2-3 us when all
    // CTAs of the launch pull at once) then overlaps the ~1.2 us of setup instead of following it.
    // chunk i of a producer thread: tile row r_i = (tid + 256 i) / 8, 16-byte chunk c = tid % 8
    const int c = tid & 7;
    const float* rowp[4];
    uint32_t off[4];
    bool okm[4];
    int seg[4];
    // register ring, WS_PF k-blocks of loads in flight per thread (64 KB per CTA): one k-block ahead leaves most of
    // the L2 round trip exposed, because converting a k-block takes far less time than fetching one
    float4 buf[WS_PF][4];
    auto load4 = [&](int kb, float4 (&q)[4]) {
        const bool okk = kb < nkb && kb * BK + c * 4 < g.K;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            q[i] = (okm[i] && okk) ? __ldg(reinterpret_cast<const float4*>(rowp[i] + kb * BK)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (warp < 8) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = (tid >> 3) + 32 * i;
            const int m = m0 + r;
            okm[i] = m < g.M;
            seg[i] = m < g.Bseg ? 0 : 1;
            rowp[i] = g.A + (size_t)(okm[i] ? m : 0) * g.K + c * 4;
            off[i] = (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
        }
#pragma unroll
        for (int j = 0; j < WS_PF; ++j) load4(j, buf[j]);
    }

    if (warp == 0) tmem_alloc(&tmem_base_slot, TMEM_COLS);
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&empty_bar[s], 1);
            mbar_init(&full_a[s], WS_PRODUCERS);
            mbar_init(&full_b[s], 1);
        }
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * kpad; i += WS_THREADS) {  // identity prologue when there is no BN
        const int seg = i / kpad, k = i - seg * kpad;
        s_scale[seg][k] = (g.scale && k < g.K) ? __ldg(g.scale + seg * g.K + k) : 1.f;
        s_shift[seg][k] = (g.scale && k < g.K) ? __ldg(g.shift + seg * g.K + k) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    const char* b_img = g.Bimg + (size_t)blockIdx.x * nkb * 2 * b_tile_bytes;

    if (warp < 8) {
        // ------------------------------------------------------------------ producers (256 threads, 4 chunks each)
        for (int kb0 = 0; kb0 < nkb; kb0 += WS_PF) {
#pragma unroll
            for (int j = 0; j < WS_PF; ++j) {
                const int kb = kb0 + j;
                if (kb >= nkb) break;
                const int st = kb % STAGES, use = kb / STAGES;
                if (use > 0) mbar_wait(&empty_bar[st], (uint32_t)((use - 1) & 1));
                char* a_hi = smem + st * stage_bytes;
                char* a_lo = a_hi + A_TILE_BYTES;
                const int k = kb * BK + c * 4;
                const bool okk = k < g.K;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 v = buf[j][i];
                    if (okm[i] && okk) {  // padding stays exactly zero
                        const float4 sc = *reinterpret_cast<const float4*>(&s_scale[seg[i]][k]);
                        const float4 sh = *reinterpret_cast<const float4*>(&s_shift[seg[i]][k]);
                        v.x = act_t<ACT>(fmaf(v.x, sc.x, sh.x));
                        v.y = act_t<ACT>(fmaf(v.y, sc.y, sh.y));
                        v.z = act_t<ACT>(fmaf(v.z, sc.z, sh.z));
                        v.w = act_t<ACT>(fmaf(v.w, sc.w, sh.w));
                    }
                    float4 hi, lo;
                    hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
                    lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y); lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
                    *reinterpret_cast<float4*>(a_hi + off[i]) = hi;
                    if (g.passes == 3) *reinterpret_cast<float4*>(a_lo + off[i]) = lo;
                }
                fence_proxy_async();  // this thread's generic-proxy writes -> visible to the tensor core
                mbar_arrive(&full_a[st]);
                load4(kb + WS_PF, buf[j]);  // refill the slot just consumed
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, BN);
            const int nacc_used = nkb <= 12 ? 1 : NACC;
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % STAGES, use = kb / STAGES;
                mbar_wait(&full_a[st], (uint32_t)(use & 1));
                mbar_wait(&full_b[st], (uint32_t)(use & 1));
                tc_fence_after();
                char* a_hi = smem + st * stage_bytes;
                const uint64_t da_hi = make_desc_k_sw128(smem_u32(a_hi)), da_lo = make_desc_k_sw128(smem_u32(a_hi + A_TILE_BYTES));
                const uint64_t db_hi = make_desc_k_sw128(smem_u32(a_hi + 2 * A_TILE_BYTES));
                const uint64_t db_lo = make_desc_k_sw128(smem_u32(a_hi + 2 * A_TILE_BYTES + b_tile_bytes));
                const uint32_t acc = tmem_d + (uint32_t)((kb % nacc_used) * MAX_BN);
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint64_t adv = (uint64_t)((ks * 8 * 4) >> 4);
                    const uint32_t first = (kb >= nacc_used || ks > 0) ? 1u : 0u;
                    if (g.passes == 3) {
                        mma_tf32(acc, da_hi + adv, db_lo + adv, idesc, first);
                        mma_tf32(acc, da_lo + adv, db_hi + adv, idesc, 1u);
                        mma_tf32(acc, da_hi + adv, db_hi + adv, idesc, 1u);
                    } else {
                        mma_tf32(acc, da_hi + adv, db_hi + adv, idesc, first);
                    }
                }
                mma_commit(&empty_bar[st]);
                if (kb == nkb - 1) mma_commit(&done_bar);
            }
        }
    } else {
        // ------------------------------------------------------------------ B image loader (warp 9)
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % STAGES, use = kb / STAGES;
                if (use > 0) mbar_wait(&empty_bar[st], (uint32_t)((use - 1) & 1));
                bulk_copy_g2s(smem + st * stage_bytes + 2 * A_TILE_BYTES, b_img + (size_t)kb * 2 * b_tile_bytes,
                              (uint32_t)(2 * b_tile_bytes), &full_b[st]);
            }
        }
    }
    __syncwarp();
    if (warp < 8) {
        mbar_wait(&done_bar, 0);
        tc_fence_after();
        // ---- epilogue (warps 0-7): TMEM -> registers -> (+bias) -> global.  Warp w owns TMEM lanes 32*(w%4)..+31 ----
        const int lane_grp = warp & 3;
        const int row = m0 + lane_grp * 32 + lane;
        const int nchunks = BN / 32;
        const int nacc_e = nkb <= 12 ? 1 : (nkb < NACC ? nkb : NACC);
        // scratch for the fused column moments: the pipeline stages are free once done_bar has completed
        float* sc_n = reinterpret_cast<float*>(smem);
        float* sc_mu = sc_n + 4 * MAX_BN;
        float* sc_m2 = sc_mu + 4 * MAX_BN;
        if (!g.fbn.on) {
            // Coalesced epilogue: TMEM -> registers -> the (now free) pipeline shared memory as a [128][BN + 4] fp32 tile ->
            // full rows to global, a warp writing 512 contiguous bytes per instruction.  Writing straight from the TMEM
            // layout (a lane = a row) made every 128-bit store instruction touch 32 different rows, 16 bytes each: ~2 us per
            // 32-column chunk (measured with in-kernel timestamps: 4.1 us of a 17 us tile at BN = 128).
            float* tile = reinterpret_cast<float*>(smem);
            const int ld = BN + 4;
            for (int ch = (warp >> 2); ch < nchunks; ch += 2) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(ch * 32), r);
                for (int a = 1; a < nacc_e; ++a) {
                    uint32_t t[32];
                    tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(a * MAX_BN + ch * 32), t);
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(t[j]));
                }
                float* dst = tile + (size_t)(lane_grp * 32 + lane) * ld + ch * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                                      __uint_as_float(r[j + 3]));
            }
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps
            const int n4 = BN / 4;                          // float4 columns of the tile (BN is a multiple of 32)
            const int nvalid = min(BN, g.N - n0);           // N is a multiple of 4 on this path
            for (int c4 = lane; c4 < n4; c4 += 32) {
                if (c4 * 4 >= nvalid) continue;
                float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.bias) bz = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + c4 * 4));
                for (int rr = warp; rr < BM; rr += 8) {
                    if (m0 + rr >= g.M) break;
                    float4 o = *reinterpret_cast<const float4*>(tile + (size_t)rr * ld + c4 * 4);
                    o.x += bz.x; o.y += bz.y; o.z += bz.z; o.w += bz.w;
                    *reinterpret_cast<float4*>(g.D + (size_t)(m0 + rr) * g.N + n0 + c4 * 4) = o;
                }
            }
        } else
        for (int ch = (warp >> 2); ch < nchunks; ch += 2) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(ch * 32), r);
            for (int a = 1; a < nacc_e; ++a) {
                uint32_t t[32];
                tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(a * MAX_BN + ch * 32), t);
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(t[j]));
            }
            const int nb = n0 + ch * 32;
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]) + ((g.bias && nb + j < g.N) ? __ldg(g.bias + nb + j) : 0.f);
            if (row < g.M) {
                float* out = g.D + (size_t)row * g.N + nb;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (nb + j + 3 < g.N) {
                        *reinterpret_cast<float4*>(out + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (nb + j + q < g.N) out[j + q] = o[j + q];
                    }
                }
            }
            if (g.fbn.on && !(g.fbn.on & 8)) {
                // (count, mean, M2) of this warp's 32 rows for the 32 columns of the chunk: shifted by the warp's first row
                // (no E[x^2] - E[x]^2 cancellation), then a butterfly transpose-reduce -- 31 shuffles per statistic leave
                // lane j with the totals of column j
                const bool valid = row < g.M;
                const int n_w = __popc(__ballot_sync(0xffffffffu, valid));
                float sq[32], kmine = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float kj = __shfl_sync(0xffffffffu, o[j], 0);
                    if (lane == j) kmine = kj;
                    const float d = valid ? o[j] - kj : 0.f;
                    o[j] = d;
                    sq[j] = d * d;
                }
#define XSTEP(OFF, NN)                                                            \
    _Pragma("unroll") for (int i = 0; i < NN; ++i) {                              \
        const bool up = (lane & OFF) != 0;                                        \
        const float send_s = up ? o[i] : o[i + NN], keep_s = up ? o[i + NN] : o[i]; \
        const float send_q = up ? sq[i] : sq[i + NN], keep_q = up ? sq[i + NN] : sq[i]; \
        o[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, OFF);                \
        sq[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, OFF);               \
    }
                XSTEP(16, 16) XSTEP(8, 8) XSTEP(4, 4) XSTEP(2, 2) XSTEP(1, 1)
#undef XSTEP
                const int cidx = lane_grp * MAX_BN + ch * 32 + lane;
                if (n_w > 0) {
                    const float md = o[0] / (float)n_w;
                    sc_n[cidx] = (float)n_w;
                    sc_mu[cidx] = kmine + md;
                    sc_m2[cidx] = fmaxf(sq[0] - o[0] * md, 0.f);
                } else {
                    sc_n[cidx] = 0.f;
                    sc_mu[cidx] = 0.f;
                    sc_m2[cidx] = 0.f;
                }
            }
        }
        if (g.fbn.on && !(g.fbn.on & 4)) {
            const int n_mtiles = gridDim.y;
            asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps
            if (tid < BN && n0 + tid < g.N) {  // the tile's triple per column: lane groups merged in row order
                float cn = 0.f, mu = 0.f, m2 = 0.f;
#pragma unroll
                for (int lg = 0; lg < 4; ++lg) chan_merge(cn, mu, m2, sc_n[lg * MAX_BN + tid], sc_mu[lg * MAX_BN + tid], sc_m2[lg * MAX_BN + tid]);
                float* part = g.fbn.part;
                part[((size_t)0 * n_mtiles + blockIdx.y) * g.N + n0 + tid] = cn;
                part[((size_t)1 * n_mtiles + blockIdx.y) * g.N + n0 + tid] = mu;
                part[((size_t)2 * n_mtiles + blockIdx.y) * g.N + n0 + tid] = m2;
            }
            // "last CTA of this N tile finalizes" (same ticket scheme as bn.cu)
            __shared__ int s_last;
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) {
                const int prev = atomicAdd(g.fbn.tickets + blockIdx.x, 1);
                s_last = prev == n_mtiles - 1;
                if (s_last) g.fbn.tickets[blockIdx.x] = 0;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (s_last && !(g.fbn.on & 2)) {
                // Finalize this N tile's columns.  The partials were written by other SMs, so every read is an L2 round
                // trip: a thread-per-column serial merge over the M tiles costs one round trip per tile, and staging through
                // a generic-pointer smem store per load serialises the same way (measured: +14..29 us per GEMM).  So: the 256
                // threads pull [3][M tiles][BN] into the (now free) pipeline smem with 128-bit loads, STAGE_MLP of them in
                // flight per thread before the first store, then one thread per (column, instance) merges its tiles in
                // ascending order from shared memory -- fixed order, deterministic.  (n_mtiles <= 64: caller's contract.)
                __threadfence();
                const int nq = g.fbn.fin.nq_chunks;
                float4* st4 = reinterpret_cast<float4*>(smem);
                const float* st = reinterpret_cast<const float*>(smem);
                const int bn4 = BN / 4;
                const int total4 = 3 * n_mtiles * bn4;
                const float* part = g.fbn.part;
                constexpr int STAGE_MLP = 8;
                for (int e0 = tid; e0 < total4; e0 += 256 * STAGE_MLP) {
                    float4 v[STAGE_MLP];
#pragma unroll
                    for (int u = 0; u < STAGE_MLP; ++u) {
                        const int e = e0 + u * 256;
                        const int pc = e / bn4, c4 = e - pc * bn4;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (e < total4 && n0 + c4 * 4 < g.N) v[u] = __ldcg(reinterpret_cast<const float4*>(part + (size_t)pc * g.N + n0 + c4 * 4));
                    }
#pragma unroll
                    for (int u = 0; u < STAGE_MLP; ++u)
                        if (e0 + u * 256 < total4) st4[e0 + u * 256] = v[u];
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                for (int pr = tid; pr < 2 * BN; pr += 256) {
                    const int seg = pr / BN, cl = pr - seg * BN, col = n0 + cl;
                    const int c0 = seg == 0 ? 0 : nq, c1 = seg == 0 ? nq : n_mtiles;
                    if (col >= g.N || c0 >= c1) continue;
                    float cn = 0.f, mu = 0.f, m2 = 0.f;
                    for (int c = c0; c < c1; ++c)
                        chan_merge(cn, mu, m2, st[(0 * n_mtiles + c) * BN + cl], st[(1 * n_mtiles + c) * BN + cl], st[(2 * n_mtiles + c) * BN + cl]);
                    bn_finalize_column(g.fbn.fin, seg * g.N + col, mu, m2 / cn);  // biased variance (tf.nn.moments)
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// TMA + TMEM variant of the warp-specialised kernel (the default for K <= 384).  The SS kernel above is bound by
// shared-memory bandwidth (160 KB per k-block -- 96 KB of operand reads by the 12 MMAs, 32 KB of hi/lo stores,
// 32 KB of weight image); here the A operand lives in TENSOR MEMORY: the raw fp32 activation tile is fetched by the TMA unit (cp.async.bulk.tensor.2d, one box of
// [128 rows x 32 floats] per k-block, SWIZZLE_128B, out-of-range rows / columns zero-filled) next to the weight image tile;
// the producers read THEIR row from shared memory (conflict-free through the swizzle), apply BN + activation, split and
// write the A operand into tensor memory (tcgen05.st; a warp reaches only the TMEM lanes 32 (warp % 4)..+31, hence one row
// per thread), and the MMAs read A from TMEM, B from shared memory (TS form).  No global-load instructions, no register
// ring, and a k-block moves 112 KB through shared memory instead of 160 KB.  TMEM columns: accumulator [0, BN), A stages at
// TS_A_COL0 + 64 st (hi: 32 columns, lo: 32 columns).  (A first TS version loaded each thread's row straight from global
// memory -- 64 contiguous bytes per thread, 32 rows per warp instruction: ncu showed the load/store unit throttled, lg 19 %
// of the stalls, and it was slower than the SS kernel, 18.2 vs 14.8 us; the TMA removes those instructions altogether.)
constexpr int TS_A_COL0 = 256;
constexpr int TMA_MAX_STAGES = 4;
static_assert(TS_A_COL0 >= MAX_BN && TS_A_COL0 + TMA_MAX_STAGES * 64 <= TMEM_COLS, "TMEM column budget of the TS / TMA kernels");
// stages of the TMA kernel: a stage is 16 KB of raw A + the two weight image tiles; four fit up to 128-column tiles
// (measured: four stages at BN <= 128 are not faster than three -- 14.65 vs 14.17 us at C2 -- so three everywhere)
__host__ __device__ inline int tma_stages(int BN) { return BN <= 160 ? 3 : 3; }

template <int ACT>
__global__ void __launch_bounds__(WS_THREADS, 1) gemm_tc3_tma_kernel(Args g, const __grid_constant__ CUtensorMap tmap_a) {
    extern __shared__ char smem_raw[];
    __shared__ uint64_t empty_bar[TMA_MAX_STAGES];
    __shared__ uint64_t full_a[TMA_MAX_STAGES];
    __shared__ uint64_t full_ld[TMA_MAX_STAGES];  // raw A tile (TMA) + weight image tile (bulk copy) of the stage have landed
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_scale[2][WS_MAX_K];
    __shared__ __align__(16) float s_shift[2][WS_MAX_K];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * g.BN;
    const int BN = g.BN;
    const int b_tile_bytes = BN * BK * 4;
    const int stage_bytes = A_TILE_BYTES + 2 * b_tile_bytes;  // [raw fp32 A tile, SWIZZLE_128B | B hi | B lo]
    char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int nkb = (g.K + BK - 1) / BK;
    const int kpad = nkb * BK;
    const int nst = tma_stages(BN);  // pipeline depth: TMA keeps nst k-blocks in flight without any registers

    // Producers (warps 0-7): a thread owns ONE row of the tile -- TMEM lane 32 (warp % 4) + lane, the only lanes its warp can
    // reach -- and 16 of the k-block's 32 columns (half = warp / 4); the first WS_PF k-blocks of loads go out before the setup.
    const int quad = warp & 3, half = warp >> 2;
    const int trow = quad * 32 + lane;
    const int mrow = m0 + trow;
    const bool okm = warp < 8 && mrow < g.M;
    const int seg = mrow < g.Bseg ? 0 : 1;

    if (warp == 0) tmem_alloc(&tmem_base_slot, TMEM_COLS);
    if (tid == 0) {
        for (int s = 0; s < nst; ++s) {
            mbar_init(&empty_bar[s], 1);
            mbar_init(&full_a[s], WS_PRODUCERS);
            mbar_init(&full_ld[s], 1);
        }
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * kpad; i += WS_THREADS) {  // identity prologue when there is no BN
        const int seg = i / kpad, k = i - seg * kpad;
        s_scale[seg][k] = (g.scale && k < g.K) ? __ldg(g.scale + seg * g.K + k) : 1.f;
        s_shift[seg][k] = (g.scale && k < g.K) ? __ldg(g.shift + seg * g.K + k) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    const char* b_img = g.Bimg + (size_t)blockIdx.x * nkb * 2 * b_tile_bytes;

    if (warp < 8) {
        // ------------------------------------------------------------------ producers (256 threads, 4 chunks each)
        {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % nst, use = kb / nst;
                mbar_wait(&full_ld[st], (uint32_t)(use & 1));  // the TMA has landed this k-block's raw tile
                // own row, chunks 4 half .. 4 half + 3 of its 128 bytes; SWIZZLE_128B: chunk c of row r sits at c ^ (r & 7)
                const char* raw = smem + st * stage_bytes + trow * 128;
                float4 bufk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) bufk[i] = *reinterpret_cast<const float4*>(raw + (((half * 4 + i) ^ (trow & 7)) << 4));
                if (use > 0) tc_fence_after();  // (the stage's TMEM slice was released through empty_bar -> loader -> full_ld)
                const int k = kb * BK + half * 16;
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 v = bufk[i];
                    if (okm && k + i * 4 < g.K) {  // padding stays exactly zero
                        const float4 sc = *reinterpret_cast<const float4*>(&s_scale[seg][k + i * 4]);
                        const float4 sh = *reinterpret_cast<const float4*>(&s_shift[seg][k + i * 4]);
                        v.x = act_t<ACT>(fmaf(v.x, sc.x, sh.x));
                        v.y = act_t<ACT>(fmaf(v.y, sc.y, sh.y));
                        v.z = act_t<ACT>(fmaf(v.z, sc.z, sh.z));
                        v.w = act_t<ACT>(fmaf(v.w, sc.w, sh.w));
                    }
                    const float h0 = tf32_rna(v.x), h1 = tf32_rna(v.y), h2 = tf32_rna(v.z), h3 = tf32_rna(v.w);
                    hi[4 * i + 0] = __float_as_uint(h0); hi[4 * i + 1] = __float_as_uint(h1);
                    hi[4 * i + 2] = __float_as_uint(h2); hi[4 * i + 3] = __float_as_uint(h3);
                    lo[4 * i + 0] = __float_as_uint(tf32_rna(v.x - h0)); lo[4 * i + 1] = __float_as_uint(tf32_rna(v.y - h1));
                    lo[4 * i + 2] = __float_as_uint(tf32_rna(v.z - h2)); lo[4 * i + 3] = __float_as_uint(tf32_rna(v.w - h3));
                }
                // straight into tensor memory: A never touches shared memory (half of the traffic that bounded the SS kernel)
                const uint32_t ta = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(TS_A_COL0 + st * 64 + half * 16);
                tmem_st_32x16(ta, hi);
                if (g.passes == 3) tmem_st_32x16(ta + 32, lo);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&full_a[st]);
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % nst, use = kb / nst;
                mbar_wait(&full_a[st], (uint32_t)(use & 1));
                mbar_wait(&full_ld[st], (uint32_t)(use & 1));
                tc_fence_after();
                char* b_st = smem + st * stage_bytes + A_TILE_BYTES;
                const uint64_t db_hi = make_desc_k_sw128(smem_u32(b_st));
                const uint64_t db_lo = make_desc_k_sw128(smem_u32(b_st + b_tile_bytes));
                const uint32_t ta_hi = tmem_d + (uint32_t)(TS_A_COL0 + st * 64), ta_lo = ta_hi + 32;
                const uint32_t acc = tmem_d;
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint64_t adv = (uint64_t)((ks * 8 * 4) >> 4);
                    const uint32_t first = (kb > 0 || ks > 0) ? 1u : 0u;
                    if (g.passes == 3) {
                        mma_tf32_ts(acc, ta_hi + ks * 8, db_lo + adv, idesc, first);
                        mma_tf32_ts(acc, ta_lo + ks * 8, db_hi + adv, idesc, 1u);
                        mma_tf32_ts(acc, ta_hi + ks * 8, db_hi + adv, idesc, 1u);
                    } else {
                        mma_tf32_ts(acc, ta_hi + ks * 8, db_hi + adv, idesc, first);
                    }
                }
                mma_commit(&empty_bar[st]);
                if (kb == nkb - 1) mma_commit(&done_bar);
            }
        }
    } else {
        // ------------------------------------------------------------------ B image loader (warp 9)
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % nst, use = kb / nst;
                if (use > 0) mbar_wait(&empty_bar[st], (uint32_t)((use - 1) & 1));
                char* stg = smem + st * stage_bytes;
                const uint32_t bar = smem_u32(&full_ld[st]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(A_TILE_BYTES + 2 * b_tile_bytes))
                             : "memory");
                // A: a [128 rows x 32 floats] box of the row-major activation matrix, 128-byte swizzled, zero-filled outside
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                                 smem_u32(stg)),
                             "l"(reinterpret_cast<uint64_t>(&tmap_a)), "r"(kb * BK), "r"(m0), "r"(bar)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(stg + A_TILE_BYTES)),
                             "l"(b_img + (size_t)kb * 2 * b_tile_bytes), "r"((uint32_t)(2 * b_tile_bytes)), "r"(bar)
                             : "memory");
            }
        }
    }
    __syncwarp();
    if (warp < 8) {
        mbar_wait(&done_bar, 0);
        tc_fence_after();
        // ---- epilogue (warps 0-7): TMEM -> registers -> (+bias) -> global.  Warp w owns TMEM lanes 32*(w%4)..+31 ----
        const int lane_grp = warp & 3;
        const int row = m0 + lane_grp * 32 + lane;
        const int nchunks = BN / 32;
        const int nacc_e = 1;  // K <= 384 on this path: one accumulator
        // scratch for the fused column moments: the pipeline stages are free once done_bar has completed
        float* sc_n = reinterpret_cast<float*>(smem);
        float* sc_mu = sc_n + 4 * MAX_BN;
        float* sc_m2 = sc_mu + 4 * MAX_BN;
        if (!g.fbn.on) {
            // Coalesced epilogue: TMEM -> registers -> the (now free) pipeline shared memory as a [128][BN + 4] fp32 tile ->
            // full rows to global, a warp writing 512 contiguous bytes per instruction.  Writing straight from the TMEM
            // layout (a lane = a row) made every 128-bit store instruction touch 32 different rows, 16 bytes each: ~2 us per
            // 32-column chunk (measured with in-kernel timestamps: 4.1 us of a 17 us tile at BN = 128).
            float* tile = reinterpret_cast<float*>(smem);
            const int ld = BN + 4;
            for (int ch = (warp >> 2); ch < nchunks; ch += 2) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(ch * 32), r);
                for (int a = 1; a < nacc_e; ++a) {
                    uint32_t t[32];
                    tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(a * MAX_BN + ch * 32), t);
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(t[j]));
                }
                float* dst = tile + (size_t)(lane_grp * 32 + lane) * ld + ch * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                                      __uint_as_float(r[j + 3]));
            }
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps
            const int n4 = BN / 4;                          // float4 columns of the tile (BN is a multiple of 32)
            const int nvalid = min(BN, g.N - n0);           // N is a multiple of 4 on this path
            for (int c4 = lane; c4 < n4; c4 += 32) {
                if (c4 * 4 >= nvalid) continue;
                float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.bias) bz = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + c4 * 4));
                for (int rr = warp; rr < BM; rr += 8) {
                    if (m0 + rr >= g.M) break;
                    float4 o = *reinterpret_cast<const float4*>(tile + (size_t)rr * ld + c4 * 4);
                    o.x += bz.x; o.y += bz.y; o.z += bz.z; o.w += bz.w;
                    *reinterpret_cast<float4*>(g.D + (size_t)(m0 + rr) * g.N + n0 + c4 * 4) = o;
                }
            }
        } else
        for (int ch = (warp >> 2); ch < nchunks; ch += 2) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(ch * 32), r);
            for (int a = 1; a < nacc_e; ++a) {
                uint32_t t[32];
                tmem_ld_32x32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(a * MAX_BN + ch * 32), t);
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(t[j]));
            }
            const int nb = n0 + ch * 32;
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]) + ((g.bias && nb + j < g.N) ? __ldg(g.bias + nb + j) : 0.f);
            if (row < g.M) {
                float* out = g.D + (size_t)row * g.N + nb;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (nb + j + 3 < g.N) {
                        *reinterpret_cast<float4*>(out + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (nb + j + q < g.N) out[j + q] = o[j + q];
                    }
                }
            }
            if (g.fbn.on && !(g.fbn.on & 8)) {
                // (count, mean, M2) of this warp's 32 rows for the 32 columns of the chunk: shifted by the warp's first row
                // (no E[x^2] - E[x]^2 cancellation), then a butterfly transpose-reduce -- 31 shuffles per statistic leave
                // lane j with the totals of column j
                const bool valid = row < g.M;
                const int n_w = __popc(__ballot_sync(0xffffffffu, valid));
                float sq[32], kmine = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float kj = __shfl_sync(0xffffffffu, o[j], 0);
                    if (lane == j) kmine = kj;
                    const float d = valid ? o[j] - kj : 0.f;
                    o[j] = d;
                    sq[j] = d * d;
                }
#define XSTEP(OFF, NN)                                                            \
    _Pragma("unroll") for (int i = 0; i < NN; ++i) {                              \
        const bool up = (lane & OFF) != 0;                                        \
        const float send_s = up ? o[i] : o[i + NN], keep_s = up ? o[i + NN] : o[i]; \
        const float send_q = up ? sq[i] : sq[i + NN], keep_q = up ? sq[i + NN] : sq[i]; \
        o[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, OFF);                \
        sq[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, OFF);               \
    }
                XSTEP(16, 16) XSTEP(8, 8) XSTEP(4, 4) XSTEP(2, 2) XSTEP(1, 1)
#undef XSTEP
                const int cidx = lane_grp * MAX_BN + ch * 32 + lane;
                if (n_w > 0) {
                    const float md = o[0] / (float)n_w;
                    sc_n[cidx] = (float)n_w;
                    sc_mu[cidx] = kmine + md;
                    sc_m2[cidx] = fmaxf(sq[0] - o[0] * md, 0.f);
                } else {
                    sc_n[cidx] = 0.f;
                    sc_mu[cidx] = 0.f;
                    sc_m2[cidx] = 0.f;
                }
            }
        }
        if (g.fbn.on && !(g.fbn.on & 4)) {
            const int n_mtiles = gridDim.y;
            asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps
            if (tid < BN && n0 + tid < g.N) {  // the tile's triple per column: lane groups merged in row order
                float cn = 0.f, mu = 0.f, m2 = 0.f;
#pragma unroll
                for (int lg = 0; lg < 4; ++lg) chan_merge(cn, mu, m2, sc_n[lg * MAX_BN + tid], sc_mu[lg * MAX_BN + tid], sc_m2[lg * MAX_BN + tid]);
                float* part = g.fbn.part;
                part[((size_t)0 * n_mtiles + blockIdx.y) * g.N + n0 + tid] = cn;
                part[((size_t)1 * n_mtiles + blockIdx.y) * g.N + n0 + tid] = mu;
                part[((size_t)2 * n_mtiles + blockIdx.y) * g.N + n0 + tid] = m2;
            }
            // "last CTA of this N tile finalizes" (same ticket scheme as bn.cu)
            __shared__ int s_last;
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) {
                const int prev = atomicAdd(g.fbn.tickets + blockIdx.x, 1);
                s_last = prev == n_mtiles - 1;
                if (s_last) g.fbn.tickets[blockIdx.x] = 0;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (s_last && !(g.fbn.on & 2)) {
                // Finalize this N tile's columns.  The partials were written by other SMs, so every read is an L2 round
                // trip: a thread-per-column serial merge over the M tiles costs one round trip per tile, and staging through
                // a generic-pointer smem store per load serialises the same way (measured: +14..29 us per GEMM).  So: the 256
                // threads pull [3][M tiles][BN] into the (now free) pipeline smem with 128-bit loads, STAGE_MLP of them in
                // flight per thread before the first store, then one thread per (column, instance) merges its tiles in
                // ascending order from shared memory -- fixed order, deterministic.  (n_mtiles <= 64: caller's contract.)
                __threadfence();
                const int nq = g.fbn.fin.nq_chunks;
                float4* st4 = reinterpret_cast<float4*>(smem);
                const float* st = reinterpret_cast<const float*>(smem);
                const int bn4 = BN / 4;
                const int total4 = 3 * n_mtiles * bn4;
                const float* part = g.fbn.part;
                constexpr int STAGE_MLP = 8;
                for (int e0 = tid; e0 < total4; e0 += 256 * STAGE_MLP) {
                    float4 v[STAGE_MLP];
#pragma unroll
                    for (int u = 0; u < STAGE_MLP; ++u) {
                        const int e = e0 + u * 256;
                        const int pc = e / bn4, c4 = e - pc * bn4;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (e < total4 && n0 + c4 * 4 < g.N) v[u] = __ldcg(reinterpret_cast<const float4*>(part + (size_t)pc * g.N + n0 + c4 * 4));
                    }
#pragma unroll
                    for (int u = 0; u < STAGE_MLP; ++u)
                        if (e0 + u * 256 < total4) st4[e0 + u * 256] = v[u];
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                for (int pr = tid; pr < 2 * BN; pr += 256) {
                    const int seg = pr / BN, cl = pr - seg * BN, col = n0 + cl;
                    const int c0 = seg == 0 ? 0 : nq, c1 = seg == 0 ? nq : n_mtiles;
                    if (col >= g.N || c0 >= c1) continue;
                    float cn = 0.f, mu = 0.f, m2 = 0.f;
                    for (int c = c0; c < c1; ++c)
                        chan_merge(cn, mu, m2, st[(0 * n_mtiles + c) * BN + cl], st[(1 * n_mtiles + c) * BN + cl], st[(2 * n_mtiles + c) * BN + cl]);
                    bn_finalize_column(g.fbn.fin, seg * g.N + col, mu, m2 / cn);  // biased variance (tf.nn.moments)
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

// Pre-split, pre-swizzled image of the weight operand B[n][k] for the K-major kernel, built once per call
// (<= 820 KB): [n_tile][k_block][hi | lo][BN rows x 128 B].  transposed: B[n][k] = src[k*ld + n], else src[n*ld + k].
__global__ void __launch_bounds__(256)
make_b_image_kernel(const float* __restrict__ src, int N, int K, int ld, int transposed, int BN, int nkb, char* __restrict__ img) {
    const int nt = blockIdx.x, kb = blockIdx.y;
    const int b_tile_bytes = BN * BK * 4;
    char* hi = img + ((size_t)nt * nkb + kb) * 2 * b_tile_bytes;
    char* lo = hi + b_tile_bytes;
    for (int id = threadIdx.x; id < BN * 8; id += blockDim.x) {
        const int r = id >> 3, c = id & 7;
        const int n = nt * BN + r, k = kb * BK + c * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = 0.f;
            if (n < N && k + j < K) v[j] = transposed ? __ldg(src + (size_t)(k + j) * ld + n) : __ldg(src + (size_t)n * ld + k + j);
        }
        stage_chunk(hi, lo, r, c, make_float4(v[0], v[1], v[2], v[3]));
    }
}

// Width of the N tiles for an [M, N] output.  With many M tiles (C3: 384) the widest tile (160 columns, two tiles for
// N = 300) keeps the per-CTA fixed costs low.  With few M tiles (C2: 48) the GEMMs are latency-bound with most SMs idle:
// narrower tiles put more CTAs on the chip (N = 300 -> 3 x 128 -> 144 CTAs on 148 SMs; N = 128 -> 2 x 64 -> 96) and leave
// ~34 KB of every SM's shared memory to the side-stream kernels that run beside them (160-wide tiles leave 10 KB, which
// kept CSC-build blocks and GEMM CTAs off each other's SMs).  M <= 0 selects the wide policy.
static int pick_bn(int N, int M) {
    const int wide_tiles = (N + MAX_BN - 1) / MAX_BN;
    int tiles = wide_tiles;
    if (M > 0) {
        const int m_tiles = (M + BM - 1) / BM;
        const int nsm = sm_count();
        for (int t = wide_tiles + 1; t <= 8; ++t) {
            const int bn = ((N + t - 1) / t + 31) / 32 * 32;
            if (bn < 64 || m_tiles * t > nsm || bn * (t - 1) >= N) break;
            tiles = t;
        }
    }
    int bn = ((N + tiles - 1) / tiles + 31) / 32 * 32;
    return bn > MAX_BN ? MAX_BN : bn;
}

// which kernel runs the K-major contractions: 2 (default) = TMA + TMEM (A operand in tensor memory), 0 = SS (both operands
// in shared memory; also used for K > 384 and for the fused-BN-moments epilogue)
static int ts_mode() {
    const char* e = getenv("DSSM_GEMM_TS");
    return e ? atoi(e) : 2;  // default: TMA + TMEM (C2: -0.7 / -0.9 us on the 300-wide GEMMs, C3: -13 % / -17 %)
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// tensor map of a row-major fp32 matrix [rows, cols] for boxes of [BM rows x BK columns], 128-byte swizzle, zero fill
static int make_tmap_a(const float* A, int rows, int cols, CUtensorMap* out) {
    EncodeTiledFn fn = encode_tiled_fn();
    DSSM_REQUIRE(fn != nullptr, DSSM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSSM_REQUIRE(r == CUDA_SUCCESS, DSSM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a [%d x %d] fp32 matrix", (int)r, rows, cols);
    return DSSM_OK;
}

static size_t smem_bytes(int BN) { return (size_t)STAGES * (2 * A_TILE_BYTES + 2 * (size_t)BN * BK * 4) + 1024; }

static int launch(const Args& a, cudaStream_t st, int splits = 0) {
    static PerDeviceOnce once;
    if (once.need()) {
        CUDA_TRY(cudaFuncSetAttribute(gemm_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_BN)));
        CUDA_TRY(cudaFuncSetAttribute(gemm_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_BN)));
        CUDA_TRY(cudaFuncSetAttribute(gemm_tc3_ws_kernel<DSSM_ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_BN)));
        CUDA_TRY(cudaFuncSetAttribute(gemm_tc3_ws_kernel<DSSM_ACT_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_BN)));
        CUDA_TRY(cudaFuncSetAttribute(gemm_tc3_ws_kernel<DSSM_ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_BN)));
        CUDA_TRY(cudaFuncSetAttribute(gemm_tc3_tma_kernel<DSSM_ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_BN)));
        CUDA_TRY(cudaFuncSetAttribute(gemm_tc3_tma_kernel<DSSM_ACT_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_BN)));
        CUDA_TRY(cudaFuncSetAttribute(gemm_tc3_tma_kernel<DSSM_ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_BN)));
    }
    if (splits > 0) {
        dim3 grid(cdiv(a.N, a.BN), cdiv(a.M, BM), splits);
        gemm_tc3_kernel<true><<<grid, THREADS, smem_bytes(a.BN), st>>>(a);
    } else if (a.Bimg && a.K <= 384 && !a.fbn.on && ts_mode() == 2 && encode_tiled_fn() != nullptr) {
        alignas(64) CUtensorMap tm;
        const int rc = make_tmap_a(a.A, a.M, a.K, &tm);
        if (rc != DSSM_OK) return rc;
        dim3 grid(cdiv(a.N, a.BN), cdiv(a.M, BM));
        // stages, and never less than the epilogue's [128][BN + 4] staging tile
        size_t sm = (size_t)tma_stages(a.BN) * (A_TILE_BYTES + 2 * (size_t)a.BN * BK * 4);
        const size_t epi = (size_t)BM * (a.BN + 4) * sizeof(float);
        sm = (sm > epi ? sm : epi) + 1024;
        if (a.act == DSSM_ACT_RELU) gemm_tc3_tma_kernel<DSSM_ACT_RELU><<<grid, WS_THREADS, sm, st>>>(a, tm);
        else if (a.act == DSSM_ACT_TANH) gemm_tc3_tma_kernel<DSSM_ACT_TANH><<<grid, WS_THREADS, sm, st>>>(a, tm);
        else gemm_tc3_tma_kernel<DSSM_ACT_NONE><<<grid, WS_THREADS, sm, st>>>(a, tm);
    } else if (a.Bimg && a.K <= WS_MAX_K - BK) {
        dim3 grid(cdiv(a.N, a.BN), cdiv(a.M, BM));
        if (a.act == DSSM_ACT_RELU) gemm_tc3_ws_kernel<DSSM_ACT_RELU><<<grid, WS_THREADS, smem_bytes(a.BN), st>>>(a);
        else if (a.act == DSSM_ACT_TANH) gemm_tc3_ws_kernel<DSSM_ACT_TANH><<<grid, WS_THREADS, smem_bytes(a.BN), st>>>(a);
        else gemm_tc3_ws_kernel<DSSM_ACT_NONE><<<grid, WS_THREADS, smem_bytes(a.BN), st>>>(a);
    } else {
        dim3 grid(cdiv(a.N, a.BN), cdiv(a.M, BM));
        gemm_tc3_kernel<false><<<grid, THREADS, smem_bytes(a.BN), st>>>(a);
    }
    LAUNCH_CHECK("gemm_tc3");
    return DSSM_OK;
}

// number of reduction splits for the dW contraction: fill the chip, keep >= 4 k-blocks per split
static int pick_splits_tc(int R, int tiles) {
    int s = (sm_count() + tiles - 1) / tiles;
    const int max_s = (R + 4 * BK - 1) / (4 * BK);
    if (s > max_s) s = max_s;
    if (s > 64) s = 64;
    return s < 1 ? 1 : s;
}

}  // namespace tc
}  // namespace dssm

using namespace dssm;

// image of an operand with N rows (tiled by BN) and K reduction columns
static size_t image_bytes(int N, int K, int M) {
    const int bn = tc::pick_bn(N, M);
    return align_up((size_t)cdiv(N, bn) * cdiv(K, tc::BK) * 2 * bn * tc::BK * 4, 256);
}

extern "C" size_t dssm_fc_tc_workspace_bytes(int32_t K, int32_t N) {
    if (K <= 0 || N <= 0) return 0;
    // forward (B = W^T) and dX (B = W), for any row count (the tile width depends on it): generous bound on the padding
    const size_t f = align_up((size_t)cdiv(K, tc::BK) * 2 * (size_t)(N + 2 * tc::MAX_BN) * tc::BK * 4, 256);
    const size_t b = align_up((size_t)cdiv(N, tc::BK) * 2 * (size_t)(K + 2 * tc::MAX_BN) * tc::BK * 4, 256);
    return f > b ? f : b;
}

static int build_image(const float* src, int N, int K, int ld, int transposed, int M, char* img, cudaStream_t st) {
    const int bn = tc::pick_bn(N, M), nkb = cdiv(K, tc::BK);
    dim3 grid(cdiv(N, bn), nkb);
    tc::make_b_image_kernel<<<grid, 256, 0, st>>>(src, N, K, ld, transposed, bn, nkb, img);
    LAUNCH_CHECK("make_b_image");
    return DSSM_OK;
}

// image of W for the forward (transposed = 1: B = W^T) or for dX (transposed = 0: B = W); internal, used by the tower to
// build all images of a step beside the forward
// R = rows of the activation operand the image will be multiplied with (decides the tile width, pick_bn)
extern "C" size_t dssm_fc_tc_image_bytes(int32_t K, int32_t N, int32_t for_dx, int32_t R) {
    return for_dx ? image_bytes(K, N, R) : image_bytes(N, K, R);
}
extern "C" int dssm_fc_tc_build_image(const float* W, int32_t K, int32_t N, int32_t for_dx, int32_t R, void* img, dssm_stream_t stream) {
    return for_dx ? build_image(W, K, N, N, 0, R, (char*)img, (cudaStream_t)stream) : build_image(W, N, K, N, 1, R, (char*)img, (cudaStream_t)stream);
}
// fused_bn: NULL, or a dssm::FusedBnStats (host struct, bn_common.cuh) -- the epilogue then also takes the column moments
// of Hout for the layer's two BN instances and the last CTA of every N tile finalizes them
extern "C" int dssm_fc_fwd_tc_img(const float* Hprev, int32_t R, int32_t K, int32_t B, const float* scale, const float* shift,
                                  int32_t act, const void* img, const float* bias, int32_t N, float* Hout, int32_t passes,
                                  const void* fused_bn, dssm_stream_t stream) {
    tc::Args a{Hprev, nullptr, Hout, bias, scale, shift, R, N, K, 0, act, B, tc::pick_bn(N, R), 0, (const char*)img, passes};
    if (fused_bn) {
        DSSM_REQUIRE(K <= tc::WS_MAX_K - tc::BK, DSSM_ERR_BAD_SHAPE, "dssm_fc_fwd_tc_img: fused BN moments need the warp-specialised kernel (K <= %d)", tc::WS_MAX_K - tc::BK);
        a.fbn = *reinterpret_cast<const FusedBnStats*>(fused_bn);
    }
    return tc::launch(a, (cudaStream_t)stream);
}
extern "C" int dssm_fc_bwd_dx_tc_img(const float* dH, int32_t R, int32_t N, const void* img, int32_t K, float* dA, int32_t passes,
                                     dssm_stream_t stream) {
    tc::Args a{dH, nullptr, dA, nullptr, nullptr, nullptr, R, K, N, 0, DSSM_ACT_NONE, 0, tc::pick_bn(K, R), 0, (const char*)img, passes};
    return tc::launch(a, (cudaStream_t)stream);
}

// forward on the tensor cores: B = W^T as a pre-split swizzled image in `workspace`
extern "C" int dssm_fc_fwd_tc(const float* Hprev, int32_t R, int32_t K, int32_t B, const float* scale, const float* shift,
                              int32_t act, const float* W, const float* bias, int32_t N, float* Hout, void* workspace,
                              size_t workspace_bytes, int32_t passes, dssm_stream_t stream) {
    DSSM_REQUIRE(K % 4 == 0 && N % 4 == 0, DSSM_ERR_BAD_SHAPE, "tensor-core dense path needs K and N multiples of 4 (K=%d N=%d)", K, N);
    DSSM_REQUIRE(aligned16(Hprev) && aligned16(W) && aligned16(Hout) && (!bias || aligned16(bias)) && (!scale || (aligned16(scale) && aligned16(shift))),
                 DSSM_ERR_BAD_ALIGN, "tensor-core dense path needs 16-byte aligned buffers");
    DSSM_REQUIRE(workspace && workspace_bytes >= image_bytes(N, K, R) && (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, DSSM_ERR_WORKSPACE,
                 "dssm_fc_fwd (tc): workspace too small or unaligned");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = build_image(W, N, K, N, 1, R, (char*)workspace, st);  // B[n][k] = W[k][n]
    if (rc != DSSM_OK) return rc;
    tc::Args a{Hprev, nullptr, Hout, bias, scale, shift, R, N, K, 0, act, B, tc::pick_bn(N, R), 0, (const char*)workspace, passes};
    return tc::launch(a, st);
}

// dA[R,K] = dH[R,N] . W[K,N]^T on the tensor cores: B = W (rows = K_layer, reduction = N_layer) as an image
extern "C" int dssm_fc_bwd_dx_tc(const float* dH, int32_t R, int32_t N, const float* W, int32_t K, float* dA, void* workspace,
                                 size_t workspace_bytes, int32_t passes, dssm_stream_t stream) {
    DSSM_REQUIRE(K % 4 == 0 && N % 4 == 0, DSSM_ERR_BAD_SHAPE, "tensor-core dense path needs K and N multiples of 4 (K=%d N=%d)", K, N);
    DSSM_REQUIRE(aligned16(dH) && aligned16(W) && aligned16(dA), DSSM_ERR_BAD_ALIGN, "tensor-core dense path needs 16-byte aligned buffers");
    DSSM_REQUIRE(workspace && workspace_bytes >= image_bytes(K, N, R) && (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, DSSM_ERR_WORKSPACE,
                 "dssm_fc_bwd_dx (tc): workspace too small or unaligned");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = build_image(W, K, N, N, 0, R, (char*)workspace, st);  // B[n'=k_layer][k'=n_layer] = W[k_layer][n_layer]
    if (rc != DSSM_OK) return rc;
    // D[M=R, N'=K] = A[M=R, K'=N] . B[N'=K, K'=N]^T
    tc::Args a{dH, nullptr, dA, nullptr, nullptr, nullptr, R, K, N, 0, DSSM_ACT_NONE, 0, tc::pick_bn(K, R), 0, (const char*)workspace, passes};
    return tc::launch(a, st);
}

// dW[K,N] partials = pro(Hprev)[rows of split]^T . dH[rows of split] on the tensor cores (both operands MN-major
// as they lie in memory); `partials` receives [splits][K][N]; returns the split count through *splits_out.
extern "C" size_t dssm_fc_bwd_dw_tc_workspace_bytes(int32_t R, int32_t K, int32_t N) {
    if (R <= 0 || K <= 0 || N <= 0) return 0;
    const int tiles = cdiv(N, tc::pick_bn(N, 0)) * cdiv(K, tc::BM);
    return align_up((size_t)tc::pick_splits_tc(R, tiles) * K * N * sizeof(float), 256);
}

extern "C" int dssm_fc_bwd_dw_tc(const float* Hprev, int32_t R, int32_t K, int32_t B, const float* scale, const float* shift,
                                 int32_t act, const float* dH, int32_t N, float* partials, int32_t* splits_out, int32_t passes,
                                 dssm_stream_t stream) {
    DSSM_REQUIRE(K % 4 == 0 && N % 4 == 0, DSSM_ERR_BAD_SHAPE, "tensor-core dense path needs K and N multiples of 4 (K=%d N=%d)", K, N);
    DSSM_REQUIRE(aligned16(Hprev) && aligned16(dH) && aligned16(partials) && (!scale || (aligned16(scale) && aligned16(shift))),
                 DSSM_ERR_BAD_ALIGN, "tensor-core dense path needs 16-byte aligned buffers");
    const int bn = tc::pick_bn(N, 0);  // dW: the parallelism comes from the split of the reduction over R
    const int tiles = cdiv(N, bn) * cdiv(K, tc::BM);
    const int splits = tc::pick_splits_tc(R, tiles);
    int kps = cdiv(R, splits);
    kps = (kps + tc::BK - 1) / tc::BK * tc::BK;
    // D[M=K_layer, N] ; reduction over the R rows
    tc::Args a{Hprev, dH, partials, nullptr, scale, shift, K, N, R, 0, act, B, bn, kps, nullptr, passes};
    *splits_out = splits;
    return tc::launch(a, (cudaStream_t)stream, splits);
}
