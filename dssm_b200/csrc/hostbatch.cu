// Host-side batch assembly for the training loop: the stacked CSR [query ; doc_pos ; doc_neg] of one step, written
// straight into (pinned) staging arrays.  Replaces what the reference does per sess.run on its one Python thread --
// three scipy row slices, three tocoo() conversions, an np.mat transpose (utils/utils.py:20-24,45-61) -- by one C call
// that releases the GIL: pull_batch takes CONTIGUOUS row ranges, so every part is one contiguous run of the epoch
// matrix's indices / data arrays (memcpy, or a converting copy for int64 counts / float64 tf-idf values) plus an offset
// indptr.  Large batches (C3: 13 MB per step) are split over a few std::threads.
#include "common.cuh"
#include <string.h>
#include <thread>
#include <vector>

namespace dssm {

template <typename S, typename D>
static void convert_copy(const S* src, D* dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = (D)src[i];
}

// kind: 0 = float32, 1 = float64, 2 = int64, 3 = int32
static void copy_values(const void* src, int kind, int64_t off, float* dst, int64_t n) {
    switch (kind) {
        case 0: memcpy(dst, (const float*)src + off, (size_t)n * sizeof(float)); break;
        case 1: convert_copy((const double*)src + off, dst, n); break;
        case 2: convert_copy((const int64_t*)src + off, dst, n); break;
        default: convert_copy((const int32_t*)src + off, dst, n); break;
    }
}
// kind: 0 = int32, 1 = int64
static void copy_indices(const void* src, int kind, int64_t off, int32_t* dst, int64_t n) {
    if (kind == 0) memcpy(dst, (const int32_t*)src + off, (size_t)n * sizeof(int32_t));
    else convert_copy((const int64_t*)src + off, dst, n);
}
static inline int64_t ip_at(const void* ip, int kind, int64_t i) { return kind == 0 ? (int64_t)((const int32_t*)ip)[i] : ((const int64_t*)ip)[i]; }

}  // namespace dssm

using namespace dssm;

extern "C" int64_t dssm_host_stack_csr(int32_t n_parts, const void* const* part_indptr, const void* const* part_indices,
                                       const void* const* part_values, const int32_t* index_kind, const int32_t* value_kind,
                                       const int64_t* row_lo, const int64_t* row_hi, int32_t* out_indptr, int32_t* out_indices,
                                       float* out_values, int64_t capacity, int32_t n_threads) {
    if (!(part_indptr && part_indices && part_values && index_kind && value_kind && row_lo && row_hi && out_indptr && out_indices && out_values)) {
        fail(DSSM_ERR_BAD_ARG, "dssm_host_stack_csr: null pointer");
        return -1;
    }
    if (n_parts <= 0 || n_parts > 16) {
        fail(DSSM_ERR_BAD_ARG, "dssm_host_stack_csr: n_parts=%d out of [1,16]", n_parts);
        return -1;
    }
    struct Job { int part; int64_t src_off, dst_off, n; };
    std::vector<Job> jobs;
    int64_t o = 0, r = 0;
    out_indptr[0] = 0;
    for (int p = 0; p < n_parts; ++p) {
        const int64_t lo = row_lo[p], hi = row_hi[p];
        if (lo < 0 || hi < lo) {
            fail(DSSM_ERR_BAD_ARG, "dssm_host_stack_csr: bad row range [%lld,%lld) of part %d", (long long)lo, (long long)hi, p);
            return -1;
        }
        const int64_t s = ip_at(part_indptr[p], index_kind[p], lo), e = ip_at(part_indptr[p], index_kind[p], hi);
        const int64_t n = e - s;
        if (o + n > capacity) {
            fail(DSSM_ERR_WORKSPACE, "dssm_host_stack_csr: batch has more than the %lld non-zeros the buffers were sized for", (long long)capacity);
            return -1;
        }
        for (int64_t i = lo + 1; i <= hi; ++i) out_indptr[r + (i - lo)] = (int32_t)(ip_at(part_indptr[p], index_kind[p], i) - s + o);
        // split a part's copy into pieces of at most 1 M entries so that the threads below get even shares
        for (int64_t c = 0; c < n; c += (1 << 20)) jobs.push_back({p, s + c, o + c, (n - c) < (1 << 20) ? (n - c) : (1 << 20)});
        o += n;
        r += hi - lo;
    }
    auto run = [&](size_t j) {
        const Job& jb = jobs[j];
        copy_indices(part_indices[jb.part], index_kind[jb.part], jb.src_off, out_indices + jb.dst_off, jb.n);
        copy_values(part_values[jb.part], value_kind[jb.part], jb.src_off, out_values + jb.dst_off, jb.n);
    };
    int nt = n_threads < 1 ? 1 : n_threads;
    if ((size_t)nt > jobs.size()) nt = (int)jobs.size();
    if (nt <= 1 || o < (1 << 19)) {
        for (size_t j = 0; j < jobs.size(); ++j) run(j);
    } else {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t)
            th.emplace_back([&, t]() { for (size_t j = t; j < jobs.size(); j += nt) run(j); });
        for (size_t j = 0; j < jobs.size(); j += nt) run(j);
        for (auto& x : th) x.join();
    }
    return o;
}
