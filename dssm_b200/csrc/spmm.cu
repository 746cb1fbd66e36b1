// FC1: CSR bag-of-grams x W1 gather-accumulate (forward) and dW1 = X^T dH (backward).
// Replaces tf.sparse_tensor_dense_matmul and its adjoint gradient, new_dssm.py:124-126.
//
// HBM-bound integer/gather work: one warp owns one output row, lanes cover the L1 columns with
// 128-bit loads (L1=300 -> 75 float4, three per lane), column indices/values are loaded coalesced by
// the warp and broadcast with shuffles, and GATHER_UNROLL independent W1-row loads are in flight per
// lane before the first FMA.  Algorithmic bytes: nnz*(4*L1+8) + R*4*L1 + (R+1)*4.
#include "common.cuh"
#include <stdlib.h>

namespace dssm {

constexpr int GATHER_UNROLL = 4;
constexpr int SPMM_THREADS = 256;
constexpr int CSC_CHUNK = 128;  // entries of one column handled by one warp in the dW1 gather
constexpr int ITEM_GRAB = 4;       // single-item columns claimed per atomic in the gather
constexpr int MAX_W1_CHUNKS = 64;  // column chunks the gather can be issued in (data-parallel comm overlap)

// acc[k] += sum_{p in [s,e)} val[p] * src4[idx[p]*L4 + lane + 32k]   (sequential in p)
template <int NCH>
__device__ __forceinline__ void gather_accumulate(const int* __restrict__ idx, const float* __restrict__ val, int s,
                                                  int e, const float4* __restrict__ src4, int L4, int lane,
                                                  float4 (&acc)[NCH]) {
    for (int base = s; base < e; base += 32) {
        const int n = min(32, e - base);
        int my_c = 0;
        float my_v = 0.f;
        if (lane < n) {
            my_c = __ldg(idx + base + lane);
            my_v = __ldg(val + base + lane);
        }
        for (int t = 0; t < n; t += GATHER_UNROLL) {
            float4 w[GATHER_UNROLL][NCH];
            float vv[GATHER_UNROLL];
#pragma unroll
            for (int u = 0; u < GATHER_UNROLL; ++u) {
                const int tt = t + u;
                const int c = __shfl_sync(0xffffffffu, my_c, tt & 31);
                vv[u] = __shfl_sync(0xffffffffu, my_v, tt & 31);
                const float4* rowp = src4 + (size_t)c * L4;
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    const int col = lane + 32 * k;
                    if (tt < n && col < L4)
                        w[u][k] = __ldg(rowp + col);
                    else
                        w[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (tt >= n) vv[u] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < GATHER_UNROLL; ++u) {
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    acc[k].x = fmaf(vv[u], w[u][k].x, acc[k].x);
                    acc[k].y = fmaf(vv[u], w[u][k].y, acc[k].y);
                    acc[k].z = fmaf(vv[u], w[u][k].z, acc[k].z);
                    acc[k].w = fmaf(vv[u], w[u][k].w, acc[k].w);
                }
            }
        }
    }
}

// Same contract as gather_accumulate, but the gathered rows travel global -> shared with cp.async (LDGSTS): no
// registers are tied up by loads in flight, so every warp keeps RING rows (RING*NCH*512 B) outstanding instead of
// GATHER_UNROLL.  A lane reads back only the chunks it copied itself, so cp.async.wait_group is the only sync needed.
constexpr int RING = 4;
template <int NCH>
__device__ __forceinline__ void gather_accumulate_async(const int* __restrict__ idx, const float* __restrict__ val, int s,
                                                        int e, const float4* __restrict__ src4, int L4, int lane,
                                                        float4* ring /* this warp's [RING][NCH*32] */, float4 (&acc)[NCH]) {
    auto issue = [&](int slot, int c) {
        const float4* rowp = src4 + (size_t)c * L4;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int col = lane + 32 * k;
            if (col < L4) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + slot * (NCH * 32) + col);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rowp + col) : "memory");
            }
        }
    };
    for (int base = s; base < e; base += 32) {
        const int n = min(32, e - base);
        int my_c = 0;
        float my_v = 0.f;
        if (lane < n) {
            my_c = __ldg(idx + base + lane);
            my_v = __ldg(val + base + lane);
        }
#pragma unroll
        for (int t = 0; t < RING; ++t) {  // prologue: RING rows in flight (empty groups keep the count uniform)
            const int c = __shfl_sync(0xffffffffu, my_c, t);
            if (t < n) issue(t, c);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int t = 0; t < n; ++t) {
            asm volatile("cp.async.wait_group %0;" ::"n"(RING - 1) : "memory");
            const float v = __shfl_sync(0xffffffffu, my_v, t);
            const int slot = t % RING;
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int col = lane + 32 * k;
                if (col < L4) {
                    const float4 w = ring[slot * (NCH * 32) + col];
                    acc[k].x = fmaf(v, w.x, acc[k].x);
                    acc[k].y = fmaf(v, w.y, acc[k].y);
                    acc[k].z = fmaf(v, w.z, acc[k].z);
                    acc[k].w = fmaf(v, w.w, acc[k].w);
                }
            }
            const int tn = t + RING;
            const int c = __shfl_sync(0xffffffffu, my_c, tn & 31);
            if (tn < n) issue(slot, c);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int NCH>
__global__ void __launch_bounds__(SPMM_THREADS)
spmm_fwd_v4_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ values,
                   const float4* __restrict__ W4, const float4* __restrict__ bias4, float4* __restrict__ Y4, int R,
                   int L4) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int stride = gridDim.x * wpb;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < R; row += stride) {
        const int s = __ldg(indptr + row), e = __ldg(indptr + row + 1);
        float4 acc[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        gather_accumulate<NCH>(indices, values, s, e, W4, L4, lane, acc);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int col = lane + 32 * k;
            if (col < L4) {
                float4 o = acc[k];
                if (bias4) {
                    const float4 b = __ldg(bias4 + col);
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                }
                Y4[(size_t)row * L4 + col] = o;
            }
        }
    }
}

// any L1 (not a multiple of 4, or wider than the vector kernels cover)
__global__ void __launch_bounds__(SPMM_THREADS)
spmm_fwd_scalar_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                       const float* __restrict__ values, const float* __restrict__ W, const float* __restrict__ bias,
                       float* __restrict__ Y, int R, int L1) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int stride = gridDim.x * wpb;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < R; row += stride) {
        const int s = __ldg(indptr + row), e = __ldg(indptr + row + 1);
        for (int col = lane; col < L1; col += 32) {
            float acc = 0.f;
            for (int p = s; p < e; ++p)
                acc = fmaf(__ldg(values + p), __ldg(W + (size_t)__ldg(indices + p) * L1 + col), acc);
            Y[(size_t)row * L1 + col] = acc + (bias ? __ldg(bias + col) : 0.f);
        }
    }
}

template <int NCH>
static void launch_fwd_v4(const int* indptr, const int* indices, const float* values, const float* W, const float* b,
                          float* Y, int R, int L1, cudaStream_t st) {
    const int wpb = SPMM_THREADS / 32;
    int blocks = cdiv(R, wpb);
    const int cap = sm_count() * 32;  // grid-stride beyond this
    if (blocks > cap) blocks = cap;
    spmm_fwd_v4_kernel<NCH><<<blocks, SPMM_THREADS, 0, st>>>(indptr, indices, values, (const float4*)W,
                                                             (const float4*)b, (float4*)Y, R, L1 / 4);
}

// ------------------------------------------------------------------------------------------------
// backward, method 1: zero-fill + vector red scatter (cross-check path)
template <int NCH>
__global__ void __launch_bounds__(SPMM_THREADS)
spmm_bwd_scatter_v4_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                           const float* __restrict__ values, const float4* __restrict__ dH4, float* __restrict__ dW,
                           int R, int L4) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int stride = gridDim.x * wpb;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < R; row += stride) {
        const int s = __ldg(indptr + row), e = __ldg(indptr + row + 1);
        float4 g[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int col = lane + 32 * k;
            g[k] = col < L4 ? __ldg(dH4 + (size_t)row * L4 + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int p = s; p < e; ++p) {
            const int c = __ldg(indices + p);
            const float v = __ldg(values + p);
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int col = lane + 32 * k;
                if (col < L4)
                    red_add4(dW + ((size_t)c * L4 + col) * 4, make_float4(v * g[k].x, v * g[k].y, v * g[k].z, v * g[k].w));
            }
        }
    }
}

__global__ void __launch_bounds__(SPMM_THREADS)
spmm_bwd_scatter_scalar_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                               const float* __restrict__ values, const float* __restrict__ dH, float* __restrict__ dW,
                               int R, int L1) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int stride = gridDim.x * wpb;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < R; row += stride) {
        const int s = __ldg(indptr + row), e = __ldg(indptr + row + 1);
        for (int p = s; p < e; ++p) {
            const int c = __ldg(indices + p);
            const float v = __ldg(values + p);
            for (int col = lane; col < L1; col += 32)
                atomicAdd(dW + (size_t)c * L1 + col, v * __ldg(dH + (size_t)row * L1 + col));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward, method 0: per-batch CSC (device-built) + gather-by-column.
//
//  csc_hist     colcnt[c] += 1 for every non-zero                       (int atomics, exact)
//  csc_scan_*   colptr = exclusive_scan(colcnt); itemptr = exclusive_scan(max(1, ceil(cnt/CSC_CHUNK)))
//  csc_fill     (row, value) of every non-zero into its column segment of a scratch copy (slot by atomic cursor)
//  csc_sort_*   every column segment is put in ROW order (a row occurs at most once per column, so this is a
//               counting sort with unique keys: row bitmap in shared memory + popcount prefix = rank), written to
//               csc_row / csc_val.  Slots handed out by the atomics differ from run to run; the sorted segments do not.
//  dw_gather    one warp per item (<= CSC_CHUNK entries of one column): sum value * dH[row,:] in entry order;
//               single-item columns write their dW1 row directly; multi-item columns park partial sums
//               and the last-arriving item adds them in item order.
// Every dW1 row is therefore summed in one fixed order (rows ascending, chunk partials in chunk order): the train
// step is bit-reproducible run to run, like the reference on TF-CPU.

__global__ void csc_hist_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, int R,
                                int* __restrict__ colcnt) {
    const int nnz = __ldg(indptr + R);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += gridDim.x * blockDim.x)
        atomicAdd(colcnt + __ldg(indices + p), 1);
}

// Two-pass multi-block exclusive scan of (colcnt, items-per-column) over the D columns.
//   pass A: every block scans SCAN_TILE consecutive columns (coalesced 4-per-thread loads, shuffle scans) and
//           writes block-local exclusive prefixes plus its two block totals;
//   pass B: every block adds the sum of the preceding blocks' totals and emits colptr / cursor / itemptr.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

// work items of a column: ceil(cnt / CSC_CHUNK); empty columns have none (their dW1 rows are zero-filled by a memset)
__device__ __forceinline__ int items_of(int n) { return (n + CSC_CHUNK - 1) / CSC_CHUNK; }

__global__ void __launch_bounds__(SCAN_THREADS)
csc_scan_local_kernel(const int* __restrict__ colcnt, int D, int* __restrict__ colptr, int* __restrict__ itemptr,
                      int2* __restrict__ block_totals) {
    __shared__ int2 warp_tot[SCAN_THREADS / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int base = blockIdx.x * SCAN_TILE + t * SCAN_PER_THREAD;
    int c[SCAN_PER_THREAD], it[SCAN_PER_THREAD];
    int sa = 0, sb = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        const int col = base + i;
        c[i] = col < D ? colcnt[col] : 0;
        it[i] = col < D ? items_of(c[i]) : 0;
        sa += c[i];
        sb += it[i];
    }
    // inclusive warp scan of the per-thread sums
    int xa = sa, xb = sb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o);
        if (lane >= o) { xa += ya; xb += yb; }
    }
    if (lane == 31) warp_tot[w] = make_int2(xa, xb);
    __syncthreads();
    int wa = 0, wb = 0;
    for (int i = 0; i < w; ++i) { wa += warp_tot[i].x; wb += warp_tot[i].y; }
    int ra = wa + xa - sa, rb = wb + xb - sb;  // exclusive prefix of this thread inside the block
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        const int col = base + i;
        if (col < D) { colptr[col] = ra; itemptr[col] = rb; }
        ra += c[i];
        rb += it[i];
    }
    if (t == SCAN_THREADS - 1) block_totals[blockIdx.x] = make_int2(ra, rb);
}

__global__ void __launch_bounds__(SCAN_THREADS)
csc_scan_add_kernel(int D, int n_blocks, const int2* __restrict__ block_totals, int* __restrict__ colptr,
                    int* __restrict__ cursor, int* __restrict__ itemptr, const int* __restrict__ colcnt,
                    int4* __restrict__ item_rec, int* __restrict__ n_heavy, int* __restrict__ heavy_list) {
    __shared__ int2 red[SCAN_THREADS / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    int oa = 0, ob = 0;
    for (int b = t; b < (int)blockIdx.x; b += SCAN_THREADS) { oa += block_totals[b].x; ob += block_totals[b].y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { oa += __shfl_xor_sync(0xffffffffu, oa, o); ob += __shfl_xor_sync(0xffffffffu, ob, o); }
    if (lane == 0) red[w] = make_int2(oa, ob);
    __syncthreads();
    oa = 0; ob = 0;
#pragma unroll
    for (int i = 0; i < SCAN_THREADS / 32; ++i) { oa += red[i].x; ob += red[i].y; }
    const int base = blockIdx.x * SCAN_TILE + t * SCAN_PER_THREAD;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        const int col = base + i;
        if (col < D) {
            const int p = colptr[col] + oa;
            colptr[col] = p;
            cursor[col] = p;
            const int first = itemptr[col] + ob;
            itemptr[col] = first;
            // one record per item {column, first entry, end entry, items of the column}: the gather reads it with
            // a single 16-byte load instead of chasing item -> column -> colptr
            const int cnt = colcnt[col];
            const int ni = items_of(cnt);
            for (int i = 0; i < ni; ++i)
                item_rec[first + i] = make_int4(col, p + i * CSC_CHUNK, min(p + cnt, p + (i + 1) * CSC_CHUNK), ni);
            if (ni > 1) {  // items of multi-item columns also go on the "heavy" list the full-range gather starts with
                const int hb = atomicAdd(n_heavy, ni);
                for (int i = 0; i < ni; ++i) heavy_list[hb + i] = first + i;
            }
        }
    }
    if ((int)blockIdx.x == n_blocks - 1 && t == 0) {
        colptr[D] = oa + block_totals[n_blocks - 1].x;
        itemptr[D] = ob + block_totals[n_blocks - 1].y;
    }
}

__global__ void __launch_bounds__(SPMM_THREADS)
csc_fill_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ values,
                int R, int* __restrict__ cursor, int* __restrict__ csc_row, float* __restrict__ csc_val) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int stride = gridDim.x * wpb;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < R; row += stride) {
        const int s = __ldg(indptr + row), e = __ldg(indptr + row + 1);
        for (int p = s + lane; p < e; p += 32) {
            const int slot = atomicAdd(cursor + __ldg(indices + p), 1);
            csc_row[slot] = row;
            csc_val[slot] = __ldg(values + p);
        }
    }
}

// Row-order every column segment of the scratch CSC (tmp_row / tmp_val, slots as the atomics gave them) into
// csc_row / csc_val.  Keys are unique inside a column.  Small kernel: one warp per column of up to SORT_WARP_MAX entries,
// ranks by comparison counting in registers (no shared memory: these kernels run on a side stream beside the tcgen05
// GEMMs, whose CTAs leave ~10-30 KB of an SM's shared memory free -- a side kernel that needs more than that would keep
// GEMM CTAs off every SM it touches); longer columns are queued for the block kernel, which ranks through a row bitmap
// (rank(row) = number of set bits below `row`).
constexpr int SORT_WARP_MAX = 128;  // = 4 entries per lane
constexpr int SORT_BIG_THREADS = 256;

__device__ __forceinline__ void bitmap_rank_scatter(const int* __restrict__ tmp_row, const float* __restrict__ tmp_val, int p,
                                                    int cnt, const uint32_t* bm, const int* pre, int* __restrict__ csc_row,
                                                    float* __restrict__ csc_val, int tid, int nthreads) {
    for (int i = tid; i < cnt; i += nthreads) {
        const int r = tmp_row[p + i];
        const float v = tmp_val[p + i];
        const int rank = pre[r >> 5] + __popc(bm[r >> 5] & ((1u << (r & 31)) - 1u));
        csc_row[p + rank] = r;
        csc_val[p + rank] = v;
    }
}

__global__ void __launch_bounds__(SPMM_THREADS)
csc_sort_small_kernel(const int* __restrict__ colcnt, const int* __restrict__ colptr, const int* __restrict__ tmp_row,
                      const float* __restrict__ tmp_val, int* __restrict__ csc_row, float* __restrict__ csc_val, int D,
                      int* __restrict__ big_ctl, int* __restrict__ big_list) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int c = blockIdx.x * wpb + w; c < D; c += gridDim.x * wpb) {
        const int cnt = __ldg(colcnt + c);
        if (cnt == 0) continue;
        if (cnt > SORT_WARP_MAX) {
            if (lane == 0) big_list[atomicAdd(big_ctl, 1)] = c;
            continue;
        }
        const int p = __ldg(colptr + c);
        constexpr int PER = SORT_WARP_MAX / 32;
        int r[PER], rank[PER];
        float v[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = lane + 32 * j;
            r[j] = 0x7fffffff;
            v[j] = 0.f;
            rank[j] = 0;
            if (i < cnt) { r[j] = tmp_row[p + i]; v[j] = tmp_val[p + i]; }
        }
#pragma unroll
        for (int jj = 0; jj < PER; ++jj) {
            if (jj * 32 >= cnt) break;  // warp-uniform
            const int lim = min(32, cnt - jj * 32);
            for (int sl = 0; sl < lim; ++sl) {
                const int key = __shfl_sync(0xffffffffu, r[jj], sl);
#pragma unroll
                for (int j = 0; j < PER; ++j) rank[j] += key < r[j] ? 1 : 0;
            }
        }
#pragma unroll
        for (int j = 0; j < PER; ++j)
            if (lane + 32 * j < cnt) { csc_row[p + rank[j]] = r[j]; csc_val[p + rank[j]] = v[j]; }
    }
}

// one block per long column (a hot gram occurs in a large share of the rows)
__global__ void __launch_bounds__(SORT_BIG_THREADS)
csc_sort_big_kernel(const int* __restrict__ colcnt, const int* __restrict__ colptr, const int* __restrict__ tmp_row,
                    const float* __restrict__ tmp_val, int* __restrict__ csc_row, float* __restrict__ csc_val, int words,
                    const int* __restrict__ big_ctl, const int* __restrict__ big_list) {
    extern __shared__ uint32_t sort_smem[];
    __shared__ int warp_tot[SORT_BIG_THREADS / 32];
    uint32_t* bm = sort_smem;
    int* pre = reinterpret_cast<int*>(bm + words);
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int nbig = __ldg(big_ctl);
    for (int b = blockIdx.x; b < nbig; b += gridDim.x) {
        const int c = __ldg(big_list + b), cnt = __ldg(colcnt + c), p = __ldg(colptr + c);
        for (int i = t; i < words; i += SORT_BIG_THREADS) bm[i] = 0u;
        __syncthreads();
        for (int i = t; i < cnt; i += SORT_BIG_THREADS) {
            const int r = tmp_row[p + i];
            atomicOr(bm + (r >> 5), 1u << (r & 31));
        }
        __syncthreads();
        const int per = (words + SORT_BIG_THREADS - 1) / SORT_BIG_THREADS, lo = min(words, t * per), hi = min(words, lo + per);
        int sum = 0;
        for (int i = lo; i < hi; ++i) sum += __popc(bm[i]);
        int x = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[w] = x;
        __syncthreads();
        int base = 0;
        for (int i = 0; i < w; ++i) base += warp_tot[i];
        int run = base + x - sum;
        for (int i = lo; i < hi; ++i) { pre[i] = run; run += __popc(bm[i]); }
        __syncthreads();
        bitmap_rank_scatter(tmp_row, tmp_val, p, cnt, bm, pre, csc_row, csc_val, t, SORT_BIG_THREADS);
        __syncthreads();
    }
}

// Optional fusion of TF-Adam on W1 into the gather (single-GPU train step): the finished gradient row never goes to
// HBM -- the warp that holds it updates w, m, v of that row in place.  Same arithmetic as adam_kernel (adam.cu).
struct AdamW1 {
    float4 *w, *m, *v;      // W1 and its Adam slots, [D, L1]
    const float* beta_pow;  // device: beta1^t, beta2^t
    float lr, b1, b2, eps;
    const int* colcnt;      // columns without entries get the g = 0 update (dense-Adam semantics of the reference)
    int absent_done;        // ... unless dssm_spmm_bwd_adam_absent already gave it to them
};

// 128-bit accesses that mark their L2 lines evict-first: the absent-column Adam streams ~0.3 GB of W1 / m / v once per step
// beside the dense layers, whose 7 MB activation tensors must stay L2-resident
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld_evict_first4(const float4* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_evict_first4(float4* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

template <int NCH, bool STREAM = false>
__device__ __forceinline__ void adam_row(const AdamW1& a, int c, int L4, int lane, const float4 (&g)[NCH], float lr_t) {
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int col = lane + 32 * k;
        if (col < L4) {
            const size_t i = (size_t)c * L4 + col;
            float4 pp, mm, vv;
            const uint64_t pol = STREAM ? l2_evict_first_policy() : 0;
            if (STREAM) { pp = ld_evict_first4(a.w + i, pol); mm = ld_evict_first4(a.m + i, pol); vv = ld_evict_first4(a.v + i, pol); }
            else { pp = a.w[i]; mm = a.m[i]; vv = a.v[i]; }
#define ADAM1(x)                                               \
    {                                                          \
        const float gr = g[k].x;                               \
        mm.x = a.b1 * mm.x + (1.f - a.b1) * gr;                \
        vv.x = a.b2 * vv.x + (1.f - a.b2) * (gr * gr);         \
        pp.x = pp.x - lr_t * mm.x / (sqrtf(vv.x) + a.eps);     \
    }
            ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
            if (STREAM) { st_evict_first4(a.w + i, pp, pol); st_evict_first4(a.m + i, mm, pol); st_evict_first4(a.v + i, vv, pol); }
            else { a.w[i] = pp; a.m[i] = mm; a.v[i] = vv; }
        }
    }
}

// Data-parallel "push" output of the gather: instead of a local dense dW1 (which the owners would have to pull over
// NVLink, one latency-bound round trip per row), every finished gradient row goes straight to the rank that OWNS the
// row -- rank o owns the W1 rows [o*per, (o+1)*per) -- into slot [self][row - o*per] of that rank's peer-mapped slot
// buffer [n_ranks][per][L1], together with a validity stamp (the step's epoch + 1).  Posted stores: the warp does not
// wait for them, rows of columns absent from the batch are never written (no zero-fill, no traffic), and the exchange
// needs only the owner pass over LOCAL memory afterwards (nvlink.cu: w1_slots_reduce_adam_kernel).
struct PushTarget {
    float4* slot[DSSM_MAX_PEERS];
    uint32_t* valid[DSSM_MAX_PEERS];
    const uint32_t* epoch;  // device: this step's epoch (own flag block)
    int n_ranks, self, per;
};

// ASYNC: rows travel through the cp.async ring (best when dH is L2-resident and latency-bound, e.g. R = 6144);
// otherwise through registers with GATHER_UNROLL loads in flight (measured faster when dH streams from HBM, R = 49152)
template <int NCH, bool ASYNC, bool FUSE_ADAM>
__global__ void __launch_bounds__(SPMM_THREADS)
dw_gather_v4_kernel(const int* __restrict__ colptr, const int* __restrict__ itemptr, const int4* __restrict__ item_rec,
                    const int* __restrict__ csc_row,
                    const float* __restrict__ csc_val, const float4* __restrict__ dH4, float4* __restrict__ dW4,
                    float4* __restrict__ partial4, int* __restrict__ done, int* __restrict__ next_item, int D, int L4,
                    int col_begin, int col_end, AdamW1 adam, const int* __restrict__ heavy_ctl /* {count} */,
                    int* __restrict__ heavy_cursor /* this launch's cursor into the heavy list */, const int* __restrict__ heavy_list,
                    PushTarget push) {
    extern __shared__ float4 ring_smem[];
    const uint32_t stamp = push.n_ranks > 0 ? __ldg(push.epoch) + 1u : 0u;
    // where the finished gradient row of column c goes: the local dense dW1, or the owner's slot buffer
    auto store_row = [&](int c, const float4 (&acc)[NCH]) {
        float4* dst;
        if (push.n_ranks > 0) {
            const int owner = c / push.per, local = c - owner * push.per;
            const size_t slot_row = (size_t)push.self * push.per + local;
            dst = push.slot[owner] + slot_row * L4;
            if ((threadIdx.x & 31) == 0) push.valid[owner][slot_row] = stamp;
        } else {
            dst = dW4 + (size_t)c * L4;
        }
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int col = (threadIdx.x & 31) + 32 * k;
            if (col < L4) dst[col] = acc[k];
        }
    };
    float lr_t = 0.f;
    if (FUSE_ADAM) {
        const float b1p = __ldg(adam.beta_pow), b2p = __ldg(adam.beta_pow + 1);
        lr_t = adam.lr * sqrtf(1.f - b2p) / (1.f - b1p);
    }
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int stride = gridDim.x * wpb;
    float4* ring = ring_smem + (size_t)(threadIdx.x >> 5) * RING * NCH * 32;
    // items of the columns [col_begin, col_end): a contiguous item range (chunked gather for comm overlap)
    const int item_lo = __ldg(itemptr + col_begin), n_items = __ldg(itemptr + col_end);
    (void)wpb; (void)stride; (void)D;
    // Dynamic work distribution (item costs range from 1 to CSC_CHUNK gathered rows; a static round-robin leaves half of
    // the SMs idle behind the stragglers).  Two refinements: (1) every launch first drains the "heavy" list -- the items of
    // multi-item columns, ~CSC_CHUNK rows each, half of all entries under a Zipf vocabulary -- so that their long
    // dependent chains start at t = 0 and spread over the warps (a column-chunk launch takes the list items inside its
    // range; without this a warp that grabbed ITEM_GRAB consecutive items of one hot column chained 4 x 128 rows and a
    // quarter-range launch took as long as the full range); (2) the remaining
    // single-item columns are claimed ITEM_GRAB at a time: one same-address atomic per item caps the kernel at the
    // L2's serialised atomic rate.
    auto process = [&](int item, bool skip_multi) {
        const int4 rec = __ldg(item_rec + item);
        const int c = rec.x, s = rec.y, e = rec.z, n_col_items = rec.w;
        if (skip_multi && n_col_items > 1) return;
        if (FUSE_ADAM) {
            // start pulling this column's w / m / v rows from HBM into L2 now: the Adam loads at the end of the item then
            // cost an L2 hit instead of a DRAM round trip that nothing else in the warp could hide
            const int row_bytes = L4 * 16;
            const int lines = (row_bytes + 127) / 128;
            for (int i = lane; i < 3 * lines; i += 32) {
                const int arr = i / lines, ln = i - arr * lines;
                const char* base = reinterpret_cast<const char*>((arr == 0 ? adam.w : arr == 1 ? adam.m : adam.v) + (size_t)c * L4);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + ln * 128));
            }
        }
        float4 acc[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < e) {
            if (ASYNC) gather_accumulate_async<NCH>(csc_row, csc_val, s, e, dH4, L4, lane, ring, acc);
            else gather_accumulate<NCH>(csc_row, csc_val, s, e, dH4, L4, lane, acc);
        }
        if (n_col_items == 1) {
            if (FUSE_ADAM) {
                adam_row<NCH>(adam, c, L4, lane, acc, lr_t);
            } else {
                store_row(c, acc);
            }
        } else {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int col = lane + 32 * k;
                if (col < L4) partial4[(size_t)item * L4 + col] = acc[k];
            }
            __threadfence();
            __syncwarp();
            int prev = 0;
            if (lane == 0) prev = atomicAdd(done + c, 1);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if (prev == n_col_items - 1) {  // last item of the column: fold the partials in item order
                __threadfence();
                const int first_item = item - (s - __ldg(colptr + c)) / CSC_CHUNK;
#pragma unroll
                for (int k = 0; k < NCH; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int i = 0; i < n_col_items; ++i) {
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        const int col = lane + 32 * k;
                        if (col < L4) {
                            const float4 p = __ldcg(partial4 + (size_t)(first_item + i) * L4 + col);
                            acc[k].x += p.x; acc[k].y += p.y; acc[k].z += p.z; acc[k].w += p.w;
                        }
                    }
                }
                if (FUSE_ADAM) {
                    adam_row<NCH>(adam, c, L4, lane, acc, lr_t);
                } else {
                    store_row(c, acc);
                }
                if (lane == 0) done[c] = 0;  // leave the counters clean for the next step
            }
        }
    };
    {   // heavy items first (every chunk scans the whole list and takes the items that fall into its range)
        const int nh = __ldg(heavy_ctl);
        for (;;) {
            int h = 0;
            if (lane == 0) h = atomicAdd(heavy_cursor, 1);
            h = __shfl_sync(0xffffffffu, h, 0);
            if (h >= nh) break;
            const int item = __ldg(heavy_list + h);
            if (item >= item_lo && item < n_items) process(item, false);
        }
    }
    for (;;) {
        int base = 0;
        if (lane == 0) base = item_lo + atomicAdd(next_item, ITEM_GRAB);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_items) break;
        const int end = min(base + ITEM_GRAB, n_items);
        for (int item = base; item < end; ++item) process(item, true);
    }
    if (FUSE_ADAM && !adam.absent_done) {
        // columns absent from the batch: zero gradient, but m, v decay and w keeps moving on its momentum
        float4 zero[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) zero[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int wpb2 = blockDim.x >> 5;
        for (int c = col_begin + blockIdx.x * wpb2 + (threadIdx.x >> 5); c < col_end; c += gridDim.x * wpb2)
            if (__ldg(adam.colcnt + c) == 0) adam_row<NCH>(adam, c, L4, lane, zero, lr_t);
    }
}

// The g = 0 Adam update of the W1 rows whose column does not occur in the batch.  It needs the column histogram only,
// and neither the forward (which reads just the rows of occurring columns) nor the backward touches those rows, so
// the train step runs it on the side stream right after the CSC build, under the dense layers.
template <int NCH>
__global__ void __launch_bounds__(SPMM_THREADS)
adam_absent_columns_kernel(int D, int L4, AdamW1 adam) {
    const float b1p = __ldg(adam.beta_pow), b2p = __ldg(adam.beta_pow + 1);
    const float lr_t = adam.lr * sqrtf(1.f - b2p) / (1.f - b1p);
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    float4 zero[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) zero[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < D; c += gridDim.x * wpb)
        if (__ldg(adam.colcnt + c) == 0) adam_row<NCH, true>(adam, c, L4, lane, zero, lr_t);
}

template <int NCH>
static void launch_adam_absent(int D, int L1, const AdamW1& ad, cudaStream_t st) {
    // ONE 256-thread block per SM.  The blocks are resident for the whole kernel (~100 us) and must leave room for the
    // main stream's CTAs on every SM: a tcgen05 GEMM CTA is 320 threads x 144 registers (46 K of the 64 K), the MN-major
    // dW CTA 256 x 199 (51 K); one block of this kernel holds 12 K, two hold 25 K -- with two resident, a GEMM launched
    // while this kernel runs finds NO SM with room and waits for it to end (measured: fc_fwd3 17 -> 86 us when the timeline
    // shifted so that they met).  8 warps x 9 float4 loads in flight per lane are ~36 KB per SM, enough to stream.
    int blocks = getenv("DSSM_ABSENT_BLOCKS") ? sm_count() * atoi(getenv("DSSM_ABSENT_BLOCKS")) : sm_count();
    const int need = cdiv(D, SPMM_THREADS / 32);
    if (blocks > need) blocks = need;
    adam_absent_columns_kernel<NCH><<<blocks, SPMM_THREADS, 0, st>>>(D, L1 / 4, ad);
}

struct CscWorkspace {
    int *colcnt, *done, *next_item, *heavy_ctl, *big_ctl, *colptr, *cursor, *itemptr, *csc_row, *heavy_list, *tmp_row, *big_list;
    int2* block_totals;
    int4* item_rec;
    float *csc_val, *tmp_val;
    float* partial;
    size_t bytes;
};

static CscWorkspace carve_csc(void* ws, int D, int L1, int64_t max_nnz) {
    Arena a(ws, (size_t)-1);
    CscWorkspace w;
    w.colcnt = a.take<int>(D + 1);   // colcnt and done are contiguous: one memset clears both
    w.done = a.take<int>(D + 1);
    w.next_item = a.take<int>(MAX_W1_CHUNKS);  // one work counter per column chunk; cleared together with colcnt / done
    w.heavy_ctl = a.take<int>(2 + MAX_W1_CHUNKS);  // {heavy items, -, one cursor per column chunk}, cleared with them
    w.big_ctl = a.take<int>(2);                // {columns queued for the block sort}, cleared with them
    w.colptr = a.take<int>(D + 1);
    w.cursor = a.take<int>(D + 1);
    w.itemptr = a.take<int>(D + 1);
    w.block_totals = a.take<int2>((size_t)(D + SCAN_TILE - 1) / SCAN_TILE + 1);
    w.csc_row = a.take<int>((size_t)max_nnz);
    w.csc_val = a.take<float>((size_t)max_nnz);
    w.tmp_row = a.take<int>((size_t)max_nnz);  // segments in atomic-slot order, before the row sort
    w.tmp_val = a.take<float>((size_t)max_nnz);
    w.big_list = a.take<int>((size_t)(max_nnz / SORT_WARP_MAX) + 2);  // columns with > SORT_WARP_MAX entries
    w.heavy_list = a.take<int>(2 * (size_t)(max_nnz / CSC_CHUNK) + 2);  // items of columns with > CSC_CHUNK entries
    const size_t max_items = (size_t)D + (size_t)(max_nnz / CSC_CHUNK) + 1;
    w.item_rec = a.take<int4>(max_items);
    w.partial = a.take<float>(max_items * (size_t)L1);
    w.bytes = a.off;
    return w;
}

template <int NCH>
static void launch_dw_gather(const CscWorkspace& w, const float* dH, float* dW, int D, int L1, int col_begin, int col_end,
                             int chunk, bool async, const AdamW1* adam, cudaStream_t st, const PushTarget* push_to = nullptr) {
    PushTarget push{};
    if (push_to) push = *push_to;
    int blocks = sm_count() * 8;
    const int cols = col_end - col_begin;
    if (blocks > cols) blocks = cols > 0 ? cols : 1;
    const size_t smem = (size_t)(SPMM_THREADS / 32) * RING * NCH * 32 * sizeof(float4);
    static PerDeviceOnce once;
    if (smem > 48 * 1024 && once.need())
        cudaFuncSetAttribute(dw_gather_v4_kernel<NCH, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    AdamW1 ad{};
    if (adam) {
        ad = *adam;
        ad.colcnt = w.colcnt;
        static PerDeviceOnce once2;
        if (smem > 48 * 1024 && once2.need())
            cudaFuncSetAttribute(dw_gather_v4_kernel<NCH, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
#define DW_ARGS w.colptr, w.itemptr, w.item_rec, w.csc_row, w.csc_val, (const float4*)dH, (float4*)dW, (float4*)w.partial, w.done, \
                w.next_item + chunk, D, L1 / 4, col_begin, col_end, ad, w.heavy_ctl, w.heavy_ctl + 2 + chunk, w.heavy_list, push
    if (async && adam) dw_gather_v4_kernel<NCH, true, true><<<blocks, SPMM_THREADS, smem, st>>>(DW_ARGS);
    else if (async) dw_gather_v4_kernel<NCH, true, false><<<blocks, SPMM_THREADS, smem, st>>>(DW_ARGS);
    else if (adam) dw_gather_v4_kernel<NCH, false, true><<<blocks, SPMM_THREADS, 0, st>>>(DW_ARGS);
    else dw_gather_v4_kernel<NCH, false, false><<<blocks, SPMM_THREADS, 0, st>>>(DW_ARGS);
#undef DW_ARGS
}

template <int NCH>
static void launch_scatter(const int* indptr, const int* indices, const float* values, const float* dH, float* dW,
                           int R, int L1, cudaStream_t st) {
    const int wpb = SPMM_THREADS / 32;
    int blocks = cdiv(R, wpb);
    const int cap = sm_count() * 32;
    if (blocks > cap) blocks = cap;
    spmm_bwd_scatter_v4_kernel<NCH><<<blocks, SPMM_THREADS, 0, st>>>(indptr, indices, values, (const float4*)dH, dW,
                                                                   R, L1 / 4);
}

// set by the tower's profiling step: recorded between the CSC build and the gather kernel
thread_local cudaEvent_t g_spmm_bwd_mid_event = nullptr;
// set by the tower around dssm_spmm_bwd_csc_build: recorded right after the column histogram, which is all that the
// absent-column Adam needs -- it then runs on its own stream beside the rest of the build (scans, fill, row sort)
thread_local cudaEvent_t g_csc_hist_done_event = nullptr;

}  // namespace dssm

using namespace dssm;

extern "C" int dssm_spmm_fwd(const int32_t* indptr, const int32_t* indices, const float* values, int32_t R, int32_t D,
                             const float* W1, const float* b1, int32_t L1, float* Y, dssm_stream_t stream) {
    DSSM_REQUIRE(indptr && W1 && Y, DSSM_ERR_BAD_ARG, "dssm_spmm_fwd: null pointer");
    DSSM_REQUIRE(R >= 0 && D > 0 && L1 > 0, DSSM_ERR_BAD_ARG, "dssm_spmm_fwd: bad sizes R=%d D=%d L1=%d", R, D, L1);
    if (R == 0) return DSSM_OK;
    DSSM_REQUIRE(indices && values, DSSM_ERR_BAD_ARG, "dssm_spmm_fwd: null indices/values");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (L1 % 4 == 0) && L1 <= 1024;
    if (vec) {
        DSSM_REQUIRE(aligned16(W1) && aligned16(Y) && (!b1 || aligned16(b1)), DSSM_ERR_BAD_ALIGN,
                     "dssm_spmm_fwd: W1/b1/Y must be 16-byte aligned when L1 %% 4 == 0");
        const int nch = cdiv(L1 / 4, 32);
        DISPATCH_NCH(nch, launch_fwd_v4<N_>(indptr, indices, values, W1, b1, Y, R, L1, st));
    } else {
        const int wpb = SPMM_THREADS / 32;
        int blocks = cdiv(R, wpb);
        const int cap = sm_count() * 32;
        if (blocks > cap) blocks = cap;
        spmm_fwd_scalar_kernel<<<blocks, SPMM_THREADS, 0, st>>>(indptr, indices, values, W1, b1, Y, R, L1);
    }
    LAUNCH_CHECK("spmm_fwd");
    return DSSM_OK;
}

extern "C" size_t dssm_spmm_bwd_dw_workspace_bytes(int32_t R, int32_t D, int32_t L1, int64_t max_nnz) {
    (void)R;
    if (D <= 0 || L1 <= 0 || max_nnz < 0) return 0;
    return carve_csc(nullptr, D, L1, max_nnz).bytes;
}

static int csc_from_workspace(int D, int L1, void* workspace, size_t workspace_bytes, CscWorkspace* out) {
    DSSM_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, DSSM_ERR_BAD_ALIGN,
                 "spmm_bwd: workspace must be 256-byte aligned");
    const size_t fixed = carve_csc(nullptr, D, L1, 0).bytes;
    DSSM_REQUIRE(workspace_bytes >= fixed, DSSM_ERR_WORKSPACE, "spmm_bwd: workspace %zu < %zu", workspace_bytes, fixed);
    // the caller sized the workspace for some max_nnz: recover the largest one whose carve fits (monotone)
    int64_t lo = 0, hi = (int64_t)1 << 40;
    while (hi - lo > 1) {
        const int64_t mid = lo + (hi - lo) / 2;
        if (carve_csc(nullptr, D, L1, mid).bytes <= workspace_bytes) lo = mid; else hi = mid;
    }
    *out = carve_csc(workspace, D, L1, lo);
    return DSSM_OK;
}

// Per-batch CSC of X (colptr / csc_row / csc_val / item table) into the workspace.  nnz lives on the device
// (indptr[R]); the caller guarantees nnz <= the max_nnz the workspace was sized for.
extern "C" int dssm_spmm_bwd_csc_build(const int32_t* indptr, const int32_t* indices, const float* values, int32_t R,
                                       int32_t D, int32_t L1, float* dW1, void* workspace, size_t workspace_bytes,
                                       dssm_stream_t stream) {
    DSSM_REQUIRE(indptr && indices && values, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_csc_build: null pointer");
    DSSM_REQUIRE(R > 0 && D > 0 && L1 > 0 && L1 % 4 == 0 && L1 <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_spmm_bwd_csc_build: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    CscWorkspace w;
    int rc = csc_from_workspace(D, L1, workspace, workspace_bytes, &w);
    if (rc != DSSM_OK) return rc;
    CUDA_TRY(cudaMemsetAsync(w.colcnt, 0, (size_t)((char*)w.colptr - (char*)w.colcnt), st));
    // rows of dW1 whose column is absent from the batch have no work item: they are zero-filled here
    if (dW1) CUDA_TRY(cudaMemsetAsync(dW1, 0, (size_t)D * L1 * sizeof(float), st));
    const int nsm = sm_count();
    csc_hist_kernel<<<nsm * 8, 256, 0, st>>>(indptr, indices, R, w.colcnt);
    LAUNCH_CHECK("csc_hist");
    if (g_csc_hist_done_event) CUDA_TRY(cudaEventRecord(g_csc_hist_done_event, st));
    const int scan_blocks = cdiv(D, SCAN_TILE);
    csc_scan_local_kernel<<<scan_blocks, SCAN_THREADS, 0, st>>>(w.colcnt, D, w.colptr, w.itemptr, w.block_totals);
    LAUNCH_CHECK("csc_scan_local");
    csc_scan_add_kernel<<<scan_blocks, SCAN_THREADS, 0, st>>>(D, scan_blocks, w.block_totals, w.colptr, w.cursor, w.itemptr,
                                                              w.colcnt, w.item_rec, w.heavy_ctl, w.heavy_list);
    LAUNCH_CHECK("csc_scan_add");
    const int wpb = SPMM_THREADS / 32;
    int blocks = cdiv(R, wpb);
    if (blocks > nsm * 32) blocks = nsm * 32;
    csc_fill_kernel<<<blocks, SPMM_THREADS, 0, st>>>(indptr, indices, values, R, w.cursor, w.tmp_row, w.tmp_val);
    LAUNCH_CHECK("csc_fill");
    // row-order the segments: fixed summation order in the gather (bit-reproducible dW1)
    const int words = cdiv(R, 32);
    const size_t bm_bytes = (size_t)2 * words * sizeof(uint32_t);  // row bitmap + popcount prefix of the block kernel
    DSSM_REQUIRE(bm_bytes <= ((size_t)160 << 10), DSSM_ERR_BAD_SHAPE, "dssm_spmm_bwd_csc_build: R=%d rows exceed the row-bitmap sort (max %d)", R,
                 (int)(((size_t)160 << 10) * 4));
    static PerDeviceOnce once;
    if (once.need()) CUDA_TRY(cudaFuncSetAttribute(csc_sort_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 << 10));
    int sblocks = cdiv(D, SPMM_THREADS / 32);
    if (sblocks > nsm * 8) sblocks = nsm * 8;
    csc_sort_small_kernel<<<sblocks, SPMM_THREADS, 0, st>>>(w.colcnt, w.colptr, w.tmp_row, w.tmp_val, w.csc_row, w.csc_val, D, w.big_ctl,
                                                            w.big_list);
    LAUNCH_CHECK("csc_sort_small");
    csc_sort_big_kernel<<<nsm * 2, SORT_BIG_THREADS, bm_bytes, st>>>(w.colcnt, w.colptr, w.tmp_row, w.tmp_val, w.csc_row, w.csc_val, words,
                                                                     w.big_ctl, w.big_list);
    LAUNCH_CHECK("csc_sort_big");
    return DSSM_OK;
}

// dW1 rows [col_begin, col_end) from the CSC in the workspace (every row of the range is written).  `chunk` selects
// the work counter; use a different chunk id (< 64) for every range issued after one csc_build.
extern "C" int dssm_spmm_bwd_dw_range(const float* dH, int32_t R, int32_t D, int32_t L1, float* dW1, int32_t col_begin, int32_t col_end,
                                      int32_t chunk, void* workspace, size_t workspace_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(dH && dW1, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_dw_range: null pointer");
    DSSM_REQUIRE(0 <= col_begin && col_begin <= col_end && col_end <= D, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_dw_range: bad column range");
    DSSM_REQUIRE(chunk >= 0 && chunk < MAX_W1_CHUNKS, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_dw_range: chunk id out of range");
    DSSM_REQUIRE(L1 % 4 == 0 && L1 <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_spmm_bwd_dw_range: bad L1");
    if (col_begin == col_end) return DSSM_OK;
    CscWorkspace w;
    int rc = csc_from_workspace(D, L1, workspace, workspace_bytes, &w);
    if (rc != DSSM_OK) return rc;
    const int nch = cdiv(L1 / 4, 32);
    const bool async = (size_t)R * L1 * sizeof(float) <= ((size_t)32 << 20);  // dH comfortably L2-resident
    DISPATCH_NCH(nch, launch_dw_gather<N_>(w, dH, dW1, D, L1, col_begin, col_end, chunk, async, nullptr, (cudaStream_t)stream));
    LAUNCH_CHECK("dw_gather");
    return DSSM_OK;
}

// Data-parallel gather with the rows PUSHED to their owners (see PushTarget): host arrays of n_ranks device pointers to
// every rank's slot buffer [n_ranks][per][L1] and validity array [n_ranks][per]; epoch = the device word holding this
// step's epoch (the flag block's).  The CSC must have been built with dW1 = NULL (no zero-fill is needed).
extern "C" int dssm_spmm_bwd_dw_push(const float* dH, int32_t R, int32_t D, int32_t L1, float* const* host_peer_slots,
                                     uint32_t* const* host_peer_valid, const uint32_t* epoch, int32_t n_ranks, int32_t self, int32_t per,
                                     void* workspace, size_t workspace_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(dH && host_peer_slots && host_peer_valid && epoch, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_dw_push: null pointer");
    DSSM_REQUIRE(n_ranks >= 1 && n_ranks <= DSSM_MAX_PEERS && self >= 0 && self < n_ranks, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_dw_push: bad ranks");
    DSSM_REQUIRE(L1 % 4 == 0 && L1 <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_spmm_bwd_dw_push: bad L1");
    DSSM_REQUIRE(per > 0 && (int64_t)per * n_ranks >= D, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_dw_push: per=%d does not cover D=%d over %d ranks", per, D, n_ranks);
    PushTarget pt{};
    for (int r = 0; r < n_ranks; ++r) {
        DSSM_REQUIRE(host_peer_slots[r] && host_peer_valid[r] && aligned16(host_peer_slots[r]), DSSM_ERR_BAD_ALIGN,
                     "dssm_spmm_bwd_dw_push: peer buffer %d null or unaligned", r);
        pt.slot[r] = (float4*)host_peer_slots[r];
        pt.valid[r] = host_peer_valid[r];
    }
    pt.epoch = epoch;
    pt.n_ranks = n_ranks;
    pt.self = self;
    pt.per = per;
    CscWorkspace w;
    int rc = csc_from_workspace(D, L1, workspace, workspace_bytes, &w);
    if (rc != DSSM_OK) return rc;
    const int nch = cdiv(L1 / 4, 32);
    const bool async = (size_t)R * L1 * sizeof(float) <= ((size_t)32 << 20);
    DISPATCH_NCH(nch, launch_dw_gather<N_>(w, dH, nullptr, D, L1, 0, D, 0, async, nullptr, (cudaStream_t)stream, &pt));
    LAUNCH_CHECK("dw_gather_push");
    return DSSM_OK;
}

// Gather fused with TF-Adam on W1 (single-GPU train step): the gradient rows are consumed in registers, W1 / m / v are
// updated in place (absent columns get the g = 0 update), dW1 is NOT produced.  beta_pow is read, not advanced.
extern "C" int dssm_spmm_bwd_dw_adam(const float* dH, int32_t R, int32_t D, int32_t L1, float* W1, float* m1, float* v1,
                                     const float* beta_pow, float lr, float beta1, float beta2, float eps, int32_t absent_done,
                                     void* workspace, size_t workspace_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(dH && W1 && m1 && v1 && beta_pow, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_dw_adam: null pointer");
    DSSM_REQUIRE(L1 % 4 == 0 && L1 <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_spmm_bwd_dw_adam: bad L1");
    DSSM_REQUIRE(aligned16(dH) && aligned16(W1) && aligned16(m1) && aligned16(v1), DSSM_ERR_BAD_ALIGN, "dssm_spmm_bwd_dw_adam: buffers must be 16-byte aligned");
    CscWorkspace w;
    int rc = csc_from_workspace(D, L1, workspace, workspace_bytes, &w);
    if (rc != DSSM_OK) return rc;
    AdamW1 ad{(float4*)W1, (float4*)m1, (float4*)v1, beta_pow, lr, beta1, beta2, eps, nullptr, absent_done != 0};
    const int nch = cdiv(L1 / 4, 32);
    const bool async = (size_t)R * L1 * sizeof(float) <= ((size_t)32 << 20);
    DISPATCH_NCH(nch, launch_dw_gather<N_>(w, dH, nullptr, D, L1, 0, D, 0, async, &ad, (cudaStream_t)stream));
    LAUNCH_CHECK("dw_gather_adam");
    return DSSM_OK;
}

// First half of the fused W1 update: TF-Adam with g = 0 on the rows of W1 / m1 / v1 whose column is absent from the
// batch whose CSC is in the workspace.  Order it after dssm_spmm_bwd_csc_build; it may run concurrently with
// dssm_spmm_fwd and the dense layers of the same step (disjoint rows); follow with dssm_spmm_bwd_dw_adam(absent_done=1).
extern "C" int dssm_spmm_bwd_adam_absent(int32_t D, int32_t L1, float* W1, float* m1, float* v1, const float* beta_pow, float lr,
                                         float beta1, float beta2, float eps, void* workspace, size_t workspace_bytes,
                                         dssm_stream_t stream) {
    DSSM_REQUIRE(W1 && m1 && v1 && beta_pow, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_adam_absent: null pointer");
    DSSM_REQUIRE(D > 0 && L1 % 4 == 0 && L1 <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_spmm_bwd_adam_absent: bad shape");
    DSSM_REQUIRE(aligned16(W1) && aligned16(m1) && aligned16(v1), DSSM_ERR_BAD_ALIGN, "dssm_spmm_bwd_adam_absent: buffers must be 16-byte aligned");
    CscWorkspace w;
    int rc = csc_from_workspace(D, L1, workspace, workspace_bytes, &w);
    if (rc != DSSM_OK) return rc;
    AdamW1 ad{(float4*)W1, (float4*)m1, (float4*)v1, beta_pow, lr, beta1, beta2, eps, w.colcnt, 0};
    const int nch = cdiv(L1 / 4, 32);
    DISPATCH_NCH(nch, launch_adam_absent<N_>(D, L1, ad, (cudaStream_t)stream));
    LAUNCH_CHECK("adam_absent_columns");
    return DSSM_OK;
}

extern "C" int dssm_spmm_bwd_dw(const int32_t* indptr, const int32_t* indices, const float* values, int32_t R,
                                int32_t D, const float* dH, int32_t L1, float* dW1, int32_t method, void* workspace,
                                size_t workspace_bytes, dssm_stream_t stream) {
    DSSM_REQUIRE(indptr && dH && dW1, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_dw: null pointer");
    DSSM_REQUIRE(R >= 0 && D > 0 && L1 > 0, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_dw: bad sizes");
    DSSM_REQUIRE(method == 0 || method == 1, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_dw: unknown method %d", method);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (L1 % 4 == 0) && L1 <= 1024;
    if (vec)
        DSSM_REQUIRE(aligned16(dH) && aligned16(dW1), DSSM_ERR_BAD_ALIGN, "dssm_spmm_bwd_dw: dH/dW1 must be 16-byte aligned");
    if (method == 1 || !vec || R == 0) {
        CUDA_TRY(cudaMemsetAsync(dW1, 0, (size_t)D * L1 * sizeof(float), st));
        if (R == 0) return DSSM_OK;
        DSSM_REQUIRE(indices && values, DSSM_ERR_BAD_ARG, "dssm_spmm_bwd_dw: null indices/values");
        if (vec) {
            const int nch = cdiv(L1 / 4, 32);
            DISPATCH_NCH(nch, launch_scatter<N_>(indptr, indices, values, dH, dW1, R, L1, st));
        } else {
            const int wpb = SPMM_THREADS / 32;
            int blocks = cdiv(R, wpb);
            const int cap = sm_count() * 32;
            if (blocks > cap) blocks = cap;
            spmm_bwd_scatter_scalar_kernel<<<blocks, SPMM_THREADS, 0, st>>>(indptr, indices, values, dH, dW1, R, L1);
        }
        LAUNCH_CHECK("spmm_bwd_scatter");
        return DSSM_OK;
    }
    int rc = dssm_spmm_bwd_csc_build(indptr, indices, values, R, D, L1, dW1, workspace, workspace_bytes, stream);
    if (rc != DSSM_OK) return rc;
    if (g_spmm_bwd_mid_event) CUDA_TRY(cudaEventRecord(g_spmm_bwd_mid_event, st));
    return dssm_spmm_bwd_dw_range(dH, R, D, L1, dW1, 0, D, 0, workspace, workspace_bytes, stream);
}
