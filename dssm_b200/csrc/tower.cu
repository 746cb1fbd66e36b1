// Tower handle: the graph of new_dssm.py:104-217 (input -> FC1 -> BN1 -> FC2 -> BN2 -> Merge_Negative_Doc ->
// Cosine_Similarity -> Loss -> Training) as one sequence of kernel launches over caller-owned buffers.
// One C call per sess.run(train_step): no Python between the kernels, and the whole step can be captured
// into a CUDA graph because every launch geometry depends only on (B, NEG, D, layers), never on nnz.
#include "common.cuh"
#include "bn_common.cuh"
#include <string>
#include <vector>
#include <string.h>
#include <stdlib.h>

namespace dssm {

struct TensorInfo {
    std::string name;
    int64_t off;  // floats; for kind 2 relative to the workspace base
    int64_t rows, cols;
};

static int64_t pad4(int64_t n) { return (n + 3) / 4 * 4; }

extern thread_local cudaEvent_t g_spmm_bwd_mid_event;
extern thread_local cudaEvent_t g_csc_hist_done_event;

}  // namespace dssm
// tensor-core dense path with prebuilt weight images (fc_tc.cu)
extern "C" size_t dssm_fc_tc_image_bytes(int32_t K, int32_t N, int32_t for_dx, int32_t R);
extern "C" int dssm_fc_tc_build_image(const float* W, int32_t K, int32_t N, int32_t for_dx, int32_t R, void* img, dssm_stream_t stream);
extern "C" int dssm_fc_fwd_tc_img(const float* Hprev, int32_t R, int32_t K, int32_t B, const float* scale, const float* shift,
                                  int32_t act, const void* img, const float* bias, int32_t N, float* Hout, int32_t passes, const void* fused_bn,
                                  dssm_stream_t stream);
extern "C" int dssm_fc_bwd_dx_tc_img(const float* dH, int32_t R, int32_t N, const void* img, int32_t K, float* dA, int32_t passes, dssm_stream_t stream);
static inline bool is_tc_mode(int mode) { return mode == DSSM_GEMM_TC_3XTF32 || mode == DSSM_GEMM_TC_TF32; }
static inline int tc_passes_of(int mode) { return mode == DSSM_GEMM_TC_TF32 ? 1 : 3; }
// the two passes of the BN backward (bn.cu) and the SyncBN exchange between replicas (nvlink.cu)
extern "C" int dssm_bn_bwd_reduce_only(const float* dA, const float* H, int32_t R, int32_t L, int32_t B, int32_t act, const float* gamma,
                                       const float* mean, const float* rstd, const float* scale, const float* shift, float* dgamma,
                                       float* dbeta, float* db, float* sumx, void* workspace, size_t workspace_bytes, dssm_stream_t stream);
extern "C" int dssm_bn_bwd_apply_only(float* dA, const float* H, int32_t R, int32_t L, int32_t B, int32_t act, const float* gamma,
                                      const float* mean, const float* rstd, const float* scale, const float* shift, const float* dgamma,
                                      const float* dbeta, dssm_stream_t stream);
extern "C" int dssm_spmm_bwd_dw_push(const float* dH, int32_t R, int32_t D, int32_t L1, float* const* host_peer_slots,
                                     uint32_t* const* host_peer_valid, const uint32_t* epoch, int32_t n_ranks, int32_t self, int32_t per,
                                     void* workspace, size_t workspace_bytes, dssm_stream_t stream);
extern "C" size_t dssm_syncbn_buffer_bytes(int32_t n_ranks, int32_t n_points, int32_t Lmax);
extern "C" int dssm_syncbn_forward(void* const* peer_bufs, int32_t n_ranks, int32_t self, int32_t point, int32_t Lmax, int32_t L,
                                   const float* gamma, const float* beta, float* ema_mean, float* ema_var, float* mean, float* var,
                                   float* rstd, float* scale, float* shift, float eps, float decay, int32_t update_ema, dssm_stream_t stream);
extern "C" int dssm_syncbn_backward(void* const* peer_bufs, int32_t n_ranks, int32_t self, int32_t point, int32_t Lmax, int32_t L,
                                    int32_t n_q, int32_t n_d, const float* gamma, const float* rstd, const float* sumx, float* dgamma,
                                    float* dbeta, float* db, dssm_stream_t stream);
namespace dssm {

__global__ void stamp_kernel(unsigned long long* slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}

// phase boundaries of one profiled step (dssm_tower_profile_step)
enum { PH_START = 0, PH_SPMM_FWD, PH_DENSE_FWD, PH_COSLOSS, PH_DENSE_BWD, PH_CSC_BUILD, PH_DW_GATHER, PH_B1, PH_ADAM, PH_COUNT };
struct PhaseTimer {
    cudaEvent_t ev[PH_COUNT];
    cudaStream_t st;
    bool on;
    bool serial;  // true: side-stream work and the W1 Adam fusion are folded back so every phase is timed alone
    std::vector<cudaEvent_t>* fine_ev;  // optional per-call timeline (dssm_tower_profile_timeline)
    std::vector<std::string>* fine_name;
    unsigned long long* stamps;  // device: when set, labels are %globaltimer stamps by a 1-thread kernel (graph-capturable)
    void mark(int i) {
        if (on && !stamps) cudaEventRecord(ev[i], st);
    }
    void markf(const std::string& name) {
        if (!on || !fine_name) return;
        if (stamps) {
            stamp_kernel<<<1, 1, 0, st>>>(stamps + fine_name->size());
            fine_name->push_back(name);
            return;
        }
        if (!fine_ev) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, st);
        fine_ev->push_back(e);
        fine_name->push_back(name);
    }
};
static thread_local PhaseTimer* g_timer = nullptr;
static inline void mark(int i) {
    if (g_timer) g_timer->mark(i);
}
static inline void markf(const std::string& name) {
    if (g_timer) g_timer->markf(name);
}

}  // namespace dssm

using namespace dssm;

struct dssm_tower {
    dssm_config cfg;
    int n_layers, R, B, NEG, D;
    int L[DSSM_MAX_LAYERS + 1];  // L[0] = D, L[l] = width of layer l
    std::vector<TensorInfo> params, ema, wst;
    int64_t P, E;
    // bound buffers
    float *params_p, *grads_p, *m_p, *v_p, *ema_p, *beta_pow_p;
    char* ws;
    size_t ws_bytes;
    int64_t max_nnz;
    bool bound, fwd_train_done;
    // workspace carve (device pointers)
    int32_t *st_indptr, *st_indices;
    float* st_values;
    float* h[DSSM_MAX_LAYERS + 1];
    float* dh[DSSM_MAX_LAYERS + 1];
    float *bn_mean[DSSM_MAX_LAYERS + 1], *bn_var[DSSM_MAX_LAYERS + 1], *bn_rstd[DSSM_MAX_LAYERS + 1],
        *bn_scale[DSSM_MAX_LAYERS + 1], *bn_shift[DSSM_MAX_LAYERS + 1];
    float* bn_sumx[DSSM_MAX_LAYERS + 1];  // [2][L] sum of xhat per BN instance (SyncBN backward)
    // SyncBN: global-batch moments over n replicas (dssm_tower_set_syncbn); 0 / 1 = per-replica moments
    int sync_n, sync_rank;
    void* sync_bufs[DSSM_MAX_PEERS];
    bool w1_push;  // data-parallel push exchange: the gather writes no local dW1, so the CSC build need not zero-fill it
    float *Y, *qnorm, *dnorm, *cos_raw, *cos_sim, *prob, *loss_terms, *loss;
    void* img_fwd[DSSM_MAX_LAYERS + 1];  // pre-split weight images of layer l (tensor-core mode), rebuilt every step
    void* img_dx[DSSM_MAX_LAYERS + 1];
    cudaEvent_t ev_img;
    bool img_forked;
    void *bn_ws, *dw_ws, *sp_ws, *fc_ws;
    size_t bn_ws_bytes, dw_ws_bytes, sp_ws_bytes, fc_ws_bytes;
    // BN moments taken in the producing GEMM's epilogue (fc_tc.cu, FusedBnStats): [tickets (256 B) | 3 x M tiles x maxL floats]
    char* fbn_ws;
    size_t fbn_ws_bytes;
    // last forward's CSR (backward reuses it)
    const int32_t *cur_indptr, *cur_indices;
    const float* cur_values;
    // graph
    cudaGraph_t graph;
    cudaGraphExec_t graph_exec;
    int64_t launches_per_step;
    // The per-batch CSC of X depends only on the input batch: it is built on a side stream, forked at the start
    // of the training forward and joined right before the dW1 gather (also inside graph capture).
    cudaStream_t side;
    // side2: the zero-gradient Adam of the W1 rows absent from the batch -- it needs only the column histogram, so it forks
    //        off `side` right after csc_hist and no longer queues behind the scans / fill / row sort (or they behind it);
    // side3: the dW GEMMs of the dense backward (independent of the dX chain once dh_l exists), joined before Adam.
    cudaStream_t side2, side3;
    // capture stream of the whole-step graphs (most urgent priority; see dssm_tower_bind)
    cudaStream_t main_hi;
    cudaEvent_t ev_hist, ev_join2, ev_dh, ev_join3;
    bool absent_forked, dw_forked;
    // pipelined host feed (dssm_tower_train_step_host_async): upload buffers k%2 filled on the copy stream
    cudaStream_t copy;
    cudaEvent_t ev_up_ready[2], ev_up_free[2], ev_step_done[2];
    int32_t* up_indptr[2];
    int32_t* up_indices[2];
    float* up_values[2];
    int64_t feed_k;
    cudaEvent_t ev_fork, ev_join;
    bool csc_forked;
    bool fuse_w1_adam;  // set for the duration of tower_step_impl: gather + Adam on W1 in one kernel, dW1 never stored
    cudaGraph_t graph_dp;  // forward + backward_begin (data-parallel pipeline)
    cudaGraphExec_t graph_dp_exec;
    int64_t launches_per_dp;
    int64_t launches;

    int find(const std::vector<TensorInfo>& v, const std::string& n) const {
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i].name == n) return (int)i;
        return -1;
    }
    float* P_(const std::string& n) const { return params_p + params[find(params, n)].off; }
    float* G_(const std::string& n) const { return grads_p + params[find(params, n)].off; }
    float* E_(const std::string& n) const { return ema_p + ema[find(ema, n)].off; }
};

static size_t tower_carve(dssm_tower* t, char* base, int64_t max_nnz) {
    Arena a(base, (size_t)-1);
    const int R = t->R, B = t->B, NEG = t->NEG, n = t->n_layers;
    t->wst.clear();
    auto reg = [&](const std::string& name, float* p, int64_t rows, int64_t cols) {
        t->wst.push_back({name, (int64_t)(((char*)p - base) / (int64_t)sizeof(float)), rows, cols});
    };
    t->st_indptr = a.take<int32_t>(R + 1);
    t->st_indices = a.take<int32_t>((size_t)max_nnz);
    t->st_values = a.take<float>((size_t)max_nnz);
    for (int b = 0; b < 2; ++b) {
        t->up_indptr[b] = a.take<int32_t>(R + 1);
        t->up_indices[b] = a.take<int32_t>((size_t)max_nnz);
        t->up_values[b] = a.take<float>((size_t)max_nnz);
    }
    for (int l = 1; l <= n; ++l) {
        t->h[l] = a.take<float>((size_t)R * t->L[l]);
        reg("h" + std::to_string(l), t->h[l], R, t->L[l]);
    }
    for (int l = 1; l <= n; ++l) {
        t->dh[l] = a.take<float>((size_t)R * t->L[l]);
        reg("dh" + std::to_string(l), t->dh[l], R, t->L[l]);
    }
    if (t->cfg.use_bn) {
        for (int l = 1; l <= n; ++l) {
            const std::string pre = "bn" + std::to_string(l) + "_";
            float** arrs[5] = {&t->bn_mean[l], &t->bn_var[l], &t->bn_rstd[l], &t->bn_scale[l], &t->bn_shift[l]};
            const char* nm[5] = {"mean", "var", "rstd", "scale", "shift"};
            for (int i = 0; i < 5; ++i) {
                *arrs[i] = a.take<float>((size_t)2 * t->L[l]);
                reg(pre + nm[i], *arrs[i], 2, t->L[l]);
            }
            t->bn_sumx[l] = a.take<float>((size_t)2 * t->L[l]);
        }
    }
    const int Ll = t->L[n];
    t->Y = a.take<float>((size_t)R * Ll);
    reg("Y", t->Y, R, Ll);
    t->qnorm = a.take<float>(B);
    reg("query_norm_single", t->qnorm, B, 1);
    t->dnorm = a.take<float>((size_t)(1 + NEG) * B);
    reg("doc_norm", t->dnorm, (int64_t)(1 + NEG) * B, 1);
    t->cos_raw = a.take<float>((size_t)(1 + NEG) * B);
    reg("cos_sim_raw", t->cos_raw, (int64_t)(1 + NEG) * B, 1);
    t->cos_sim = a.take<float>((size_t)(1 + NEG) * B);
    reg("cos_sim", t->cos_sim, B, 1 + NEG);
    t->prob = a.take<float>((size_t)(1 + NEG) * B);
    reg("prob", t->prob, B, 1 + NEG);
    t->loss_terms = a.take<float>(B);
    reg("loss_terms", t->loss_terms, B, 1);
    t->loss = a.take<float>(4);
    reg("loss", t->loss, 1, 1);
    // scratch
    int maxL = 0;
    for (int l = 1; l <= n; ++l) maxL = t->L[l] > maxL ? t->L[l] : maxL;
    t->bn_ws_bytes = dssm_bn_workspace_bytes(R, maxL);
    t->bn_ws = a.take<char>(t->bn_ws_bytes);
    t->fbn_ws_bytes = 256 + (size_t)3 * ((R + 127) / 128) * maxL * sizeof(float);
    t->fbn_ws = a.take<char>(t->fbn_ws_bytes);
    size_t dw = dssm_colsum_workspace_bytes(R, maxL);
    for (int l = 2; l <= n; ++l) {
        const size_t x = dssm_fc_bwd_dw_workspace_bytes(R, t->L[l - 1], t->L[l]);
        dw = x > dw ? x : dw;
    }
    t->dw_ws_bytes = dw;
    t->dw_ws = a.take<char>(dw);
    size_t fcw = 0;
    for (int l = 2; l <= n; ++l) {
        const size_t x = dssm_fc_fwd_workspace_bytes(t->L[l - 1], t->L[l], t->cfg.gemm_mode);
        fcw = x > fcw ? x : fcw;
    }
    t->fc_ws_bytes = fcw;
    t->fc_ws = a.take<char>(fcw ? fcw : 256);
    for (int l = 2; l <= n; ++l) {
        t->img_fwd[l] = t->img_dx[l] = nullptr;
        if (is_tc_mode(t->cfg.gemm_mode)) {
            t->img_fwd[l] = a.take<char>(dssm_fc_tc_image_bytes(t->L[l - 1], t->L[l], 0, R));
            t->img_dx[l] = a.take<char>(dssm_fc_tc_image_bytes(t->L[l - 1], t->L[l], 1, R));
        }
    }
    t->sp_ws_bytes = dssm_spmm_bwd_dw_workspace_bytes(R, t->D, t->L[1], max_nnz);
    t->sp_ws = a.take<char>(t->sp_ws_bytes);
    return a.off;
}

extern "C" int dssm_tower_create(const dssm_config* cfg, dssm_tower** out) {
    DSSM_REQUIRE(cfg && out, DSSM_ERR_BAD_ARG, "dssm_tower_create: null pointer");
    DSSM_REQUIRE(cfg->n_layers >= 1 && cfg->n_layers <= DSSM_MAX_LAYERS, DSSM_ERR_BAD_ARG, "dssm_tower_create: n_layers=%d out of [1,%d]", cfg->n_layers, DSSM_MAX_LAYERS);
    DSSM_REQUIRE(cfg->TRIGRAM_D > 0 && cfg->NEG > 0 && cfg->query_BS > 0, DSSM_ERR_BAD_ARG, "dssm_tower_create: TRIGRAM_D, NEG, query_BS must be positive");
    DSSM_REQUIRE(cfg->act == DSSM_ACT_RELU || cfg->act == DSSM_ACT_TANH || cfg->act == DSSM_ACT_NONE, DSSM_ERR_BAD_ARG, "dssm_tower_create: unknown act %d", cfg->act);
    DSSM_REQUIRE(cfg->gemm_mode == DSSM_GEMM_FP32 || is_tc_mode(cfg->gemm_mode), DSSM_ERR_BAD_ARG, "dssm_tower_create: unknown gemm_mode %d", cfg->gemm_mode);
    for (int l = 0; l < cfg->n_layers; ++l)
        DSSM_REQUIRE(cfg->layers[l] > 0, DSSM_ERR_BAD_ARG, "dssm_tower_create: layer %d width %d", l + 1, cfg->layers[l]);
    const int64_t rows = (int64_t)(2 + cfg->NEG) * cfg->query_BS;
    DSSM_REQUIRE(rows < (int64_t)1 << 30, DSSM_ERR_BAD_SHAPE, "dssm_tower_create: (2+NEG)*query_BS too large");
    dssm_tower* t = new dssm_tower();
    t->cfg = *cfg;
    t->n_layers = cfg->n_layers;
    t->B = cfg->query_BS;
    t->NEG = cfg->NEG;
    t->D = cfg->TRIGRAM_D;
    t->R = (int)rows;
    t->L[0] = t->D;
    for (int l = 1; l <= t->n_layers; ++l) t->L[l] = cfg->layers[l - 1];
    int64_t off = 0;
    for (int l = 1; l <= t->n_layers; ++l) {
        t->params.push_back({"W" + std::to_string(l), off, t->L[l - 1], t->L[l]});
        off += pad4((int64_t)t->L[l - 1] * t->L[l]);
        t->params.push_back({"b" + std::to_string(l), off, 1, t->L[l]});
        off += pad4(t->L[l]);
    }
    int64_t eoff = 0;
    if (cfg->use_bn) {
        for (int l = 1; l <= t->n_layers; ++l) {
            t->params.push_back({"bn" + std::to_string(l) + "_gamma", off, 2, t->L[l]});
            off += pad4(2 * (int64_t)t->L[l]);
            t->params.push_back({"bn" + std::to_string(l) + "_beta", off, 2, t->L[l]});
            off += pad4(2 * (int64_t)t->L[l]);
            t->ema.push_back({"bn" + std::to_string(l) + "_ema_mean", eoff, 2, t->L[l]});
            eoff += pad4(2 * (int64_t)t->L[l]);
            t->ema.push_back({"bn" + std::to_string(l) + "_ema_var", eoff, 2, t->L[l]});
            eoff += pad4(2 * (int64_t)t->L[l]);
        }
    }
    t->P = off;
    t->E = eoff;
    t->bound = false;
    t->fwd_train_done = false;
    t->graph = nullptr;
    t->graph_exec = nullptr;
    t->launches_per_step = 0;
    t->graph_dp = nullptr;
    t->graph_dp_exec = nullptr;
    t->launches_per_dp = 0;
    t->side = nullptr;
    t->side2 = t->side3 = nullptr;
    t->main_hi = nullptr;
    t->ev_hist = t->ev_join2 = t->ev_dh = t->ev_join3 = nullptr;
    t->absent_forked = t->dw_forked = false;
    t->copy = nullptr;
    for (int b = 0; b < 2; ++b) t->ev_up_ready[b] = t->ev_up_free[b] = t->ev_step_done[b] = nullptr;
    t->feed_k = 0;
    t->ev_fork = nullptr;
    t->ev_join = nullptr;
    t->ev_img = nullptr;
    t->csc_forked = false;
    t->img_forked = false;
    t->launches = 0;
    t->fuse_w1_adam = false;
    t->sync_n = 0;
    t->sync_rank = 0;
    t->w1_push = false;
    tower_carve(t, nullptr, 0);  // populate the workspace tensor table (offsets are final after bind)
    *out = t;
    return DSSM_OK;
}

extern "C" void dssm_tower_destroy(dssm_tower* t) {
    if (!t) return;
    if (t->graph_exec) cudaGraphExecDestroy(t->graph_exec);
    if (t->graph) cudaGraphDestroy(t->graph);
    if (t->graph_dp_exec) cudaGraphExecDestroy(t->graph_dp_exec);
    if (t->graph_dp) cudaGraphDestroy(t->graph_dp);
    if (t->ev_fork) cudaEventDestroy(t->ev_fork);
    if (t->ev_join) cudaEventDestroy(t->ev_join);
    if (t->ev_img) cudaEventDestroy(t->ev_img);
    if (t->side) cudaStreamDestroy(t->side);
    if (t->side2) cudaStreamDestroy(t->side2);
    if (t->side3) cudaStreamDestroy(t->side3);
    if (t->main_hi) cudaStreamDestroy(t->main_hi);
    for (cudaEvent_t e : {t->ev_hist, t->ev_join2, t->ev_dh, t->ev_join3})
        if (e) cudaEventDestroy(e);
    for (int b = 0; b < 2; ++b) {
        if (t->ev_up_ready[b]) cudaEventDestroy(t->ev_up_ready[b]);
        if (t->ev_up_free[b]) cudaEventDestroy(t->ev_up_free[b]);
        if (t->ev_step_done[b]) cudaEventDestroy(t->ev_step_done[b]);
    }
    if (t->copy) cudaStreamDestroy(t->copy);
    delete t;
}

extern "C" int64_t dssm_tower_param_count(const dssm_tower* t) { return t ? t->P : -1; }
extern "C" int64_t dssm_tower_ema_count(const dssm_tower* t) { return t ? t->E : -1; }

extern "C" size_t dssm_tower_workspace_bytes(const dssm_tower* t, int64_t max_nnz) {
    if (!t || max_nnz < 0) return 0;
    dssm_tower tmp = *t;
    tmp.graph = nullptr;
    tmp.graph_exec = nullptr;
    tmp.graph_dp = nullptr;
    tmp.graph_dp_exec = nullptr;
    tmp.side = nullptr;
    tmp.side2 = tmp.side3 = nullptr;
    tmp.main_hi = nullptr;
    tmp.ev_hist = tmp.ev_join2 = tmp.ev_dh = tmp.ev_join3 = nullptr;
    tmp.ev_fork = nullptr;
    tmp.ev_join = nullptr;
    tmp.ev_img = nullptr;
    return tower_carve(&tmp, nullptr, max_nnz);
}

extern "C" int32_t dssm_tower_num_tensors(const dssm_tower* t, int32_t kind) {
    if (!t) return -1;
    if (kind == 0) return (int32_t)t->params.size();
    if (kind == 1) return (int32_t)t->ema.size();
    if (kind == 2) return (int32_t)t->wst.size();
    return -1;
}

extern "C" int dssm_tower_tensor_info(const dssm_tower* t, int32_t kind, int32_t index, char* name, int32_t name_cap,
                                      int64_t* offset_floats, int64_t* rows, int64_t* cols) {
    DSSM_REQUIRE(t, DSSM_ERR_BAD_ARG, "dssm_tower_tensor_info: null tower");
    const std::vector<TensorInfo>* v = kind == 0 ? &t->params : kind == 1 ? &t->ema : kind == 2 ? &t->wst : nullptr;
    DSSM_REQUIRE(v && index >= 0 && index < (int)v->size(), DSSM_ERR_BAD_ARG, "dssm_tower_tensor_info: bad kind/index %d/%d", kind, index);
    const TensorInfo& ti = (*v)[index];
    if (name && name_cap > 0) {
        strncpy(name, ti.name.c_str(), name_cap - 1);
        name[name_cap - 1] = 0;
    }
    if (offset_floats) *offset_floats = ti.off;
    if (rows) *rows = ti.rows;
    if (cols) *cols = ti.cols;
    return DSSM_OK;
}

extern "C" int dssm_tower_bind(dssm_tower* t, float* params, float* grads, float* m, float* v, float* ema,
                               float* beta_pow, void* workspace, size_t workspace_bytes, int64_t max_nnz) {
    DSSM_REQUIRE(t && params && workspace, DSSM_ERR_BAD_ARG, "dssm_tower_bind: null pointer");
    DSSM_REQUIRE(!t->cfg.use_bn || ema, DSSM_ERR_BAD_ARG, "dssm_tower_bind: ema buffer required with use_bn");
    DSSM_REQUIRE(aligned16(params) && (!grads || aligned16(grads)) && (!m || aligned16(m)) && (!v || aligned16(v)) && (!ema || aligned16(ema)),
                 DSSM_ERR_BAD_ALIGN, "dssm_tower_bind: buffers must be 16-byte aligned");
    DSSM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, DSSM_ERR_BAD_ALIGN, "dssm_tower_bind: workspace must be 256-byte aligned");
    DSSM_REQUIRE(max_nnz > 0 && max_nnz < (int64_t)1 << 31, DSSM_ERR_BAD_ARG, "dssm_tower_bind: max_nnz out of range");
    const size_t need = dssm_tower_workspace_bytes(t, max_nnz);
    DSSM_REQUIRE(workspace_bytes >= need, DSSM_ERR_WORKSPACE, "dssm_tower_bind: workspace %zu < required %zu", workspace_bytes, need);
    if (t->graph_exec) { cudaGraphExecDestroy(t->graph_exec); t->graph_exec = nullptr; }
    if (t->graph) { cudaGraphDestroy(t->graph); t->graph = nullptr; }
    if (t->graph_dp_exec) { cudaGraphExecDestroy(t->graph_dp_exec); t->graph_dp_exec = nullptr; }
    if (t->graph_dp) { cudaGraphDestroy(t->graph_dp); t->graph_dp = nullptr; }
    t->params_p = params; t->grads_p = grads; t->m_p = m; t->v_p = v; t->ema_p = ema; t->beta_pow_p = beta_pow;
    t->ws = (char*)workspace;
    t->ws_bytes = workspace_bytes;
    t->max_nnz = max_nnz;
    tower_carve(t, t->ws, max_nnz);
    // the reduction kernels keep self-resetting ticket counters at the head of their workspaces
    CUDA_TRY(cudaMemset(t->bn_ws, 0, t->bn_ws_bytes));
    CUDA_TRY(cudaMemset(t->fbn_ws, 0, 256));
    CUDA_TRY(cudaMemset(t->dw_ws, 0, t->dw_ws_bytes));
    if (!t->side) {
        // Priorities: the side streams carry work with slack (CSC build, absent-row Adam, dW GEMMs -- consumed only at the end
        // of the step), the main chain has none.  Side streams get the LEAST urgent priority and the step is captured on an
        // internal stream of the MOST urgent one (graph kernel nodes inherit the capturing stream's priority), so the block
        // scheduler places main-chain CTAs first whenever both have blocks pending.  Without it a dW GEMM (144 CTAs, one per
        // SM) launched beside the dX GEMM of the same layer held the SMs for its whole life and the dX GEMM on the critical
        // path ran after it (fc_dx3: 23 us inside the step, 12 alone).  DSSM_SIDE_PRIORITY=0: one priority for all (A/B runs).
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));  // lo = numerically greatest = least urgent
        const char* sp = getenv("DSSM_SIDE_PRIORITY");
        if (sp && sp[0] == '0') prio_hi = prio_lo;
        CUDA_TRY(cudaStreamCreateWithPriority(&t->side, cudaStreamNonBlocking, prio_lo));
        CUDA_TRY(cudaStreamCreateWithPriority(&t->side2, cudaStreamNonBlocking, prio_lo));
        CUDA_TRY(cudaStreamCreateWithPriority(&t->side3, cudaStreamNonBlocking, prio_lo));
        CUDA_TRY(cudaStreamCreateWithPriority(&t->main_hi, cudaStreamNonBlocking, prio_hi));
        CUDA_TRY(cudaEventCreateWithFlags(&t->ev_hist, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&t->ev_join2, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&t->ev_dh, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&t->ev_join3, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&t->ev_img, cudaEventDisableTiming));
        CUDA_TRY(cudaStreamCreateWithFlags(&t->copy, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            CUDA_TRY(cudaEventCreateWithFlags(&t->ev_up_ready[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&t->ev_up_free[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&t->ev_step_done[b], cudaEventDisableTiming));
        }
    }
    t->feed_k = 0;
    t->csc_forked = false;
    t->bound = true;
    t->fwd_train_done = false;
    return DSSM_OK;
}

extern "C" size_t dssm_tower_syncbn_bytes(const dssm_tower* t, int32_t n_ranks) {
    if (!t || n_ranks < 1 || n_ranks > DSSM_MAX_PEERS) return 0;
    int m = 0;
    for (int l = 1; l <= t->n_layers; ++l) m = t->L[l] > m ? t->L[l] : m;
    return dssm_syncbn_buffer_bytes(n_ranks, 2 * t->n_layers, m);
}

extern "C" int dssm_tower_set_syncbn(dssm_tower* t, int32_t n_ranks, int32_t rank, void* const* host_peer_bufs) {
    DSSM_REQUIRE(t, DSSM_ERR_BAD_ARG, "dssm_tower_set_syncbn: null tower");
    // captured graphs hold the previous mode's launches
    if (t->graph_exec) { cudaGraphExecDestroy(t->graph_exec); t->graph_exec = nullptr; }
    if (t->graph) { cudaGraphDestroy(t->graph); t->graph = nullptr; }
    if (t->graph_dp_exec) { cudaGraphExecDestroy(t->graph_dp_exec); t->graph_dp_exec = nullptr; }
    if (t->graph_dp) { cudaGraphDestroy(t->graph_dp); t->graph_dp = nullptr; }
    if (n_ranks <= 1) {
        t->sync_n = 0;
        return DSSM_OK;
    }
    DSSM_REQUIRE(t->cfg.use_bn, DSSM_ERR_STATE, "dssm_tower_set_syncbn: the tower has no BatchNorm");
    DSSM_REQUIRE(n_ranks <= DSSM_MAX_PEERS && rank >= 0 && rank < n_ranks && host_peer_bufs, DSSM_ERR_BAD_ARG,
                 "dssm_tower_set_syncbn: n_ranks=%d rank=%d (at most %d peers)", n_ranks, rank, DSSM_MAX_PEERS);
    for (int l = 1; l <= t->n_layers; ++l)
        DSSM_REQUIRE(t->L[l] <= 1024, DSSM_ERR_BAD_SHAPE, "dssm_tower_set_syncbn: layer width %d > 1024", t->L[l]);
    for (int r = 0; r < n_ranks; ++r) {
        DSSM_REQUIRE(host_peer_bufs[r] && aligned16(host_peer_bufs[r]), DSSM_ERR_BAD_ALIGN, "dssm_tower_set_syncbn: peer buffer %d null or unaligned", r);
        t->sync_bufs[r] = host_peer_bufs[r];
    }
    t->sync_n = n_ranks;
    t->sync_rank = rank;
    return DSSM_OK;
}

#define TRY(call)                  \
    do {                           \
        int _rc = (call);          \
        if (_rc != DSSM_OK) return _rc; \
    } while (0)

static int max_width(const dssm_tower* t) {
    int m = 0;
    for (int l = 1; l <= t->n_layers; ++l) m = t->L[l] > m ? t->L[l] : m;
    return m;
}

static int tower_forward_impl(dssm_tower* t, const int32_t* indptr, const int32_t* indices, const float* values,
                              int on_train, int update_ema, bool want_grad, dssm_stream_t s) {
    const dssm_config& c = t->cfg;
    const int n = t->n_layers, R = t->R, B = t->B;
    mark(PH_START);
    markf("start");
    t->csc_forked = false;
    t->img_forked = false;
    const bool tc = is_tc_mode(c.gemm_mode) && n >= 2;
    const bool timing = g_timer && g_timer->on && g_timer->serial;
    const bool train = want_grad && on_train;
    const bool fork_csc = train && t->L[1] % 4 == 0 && t->L[1] <= 1024 && !timing;
    cudaStream_t main_st = (cudaStream_t)s;
    // Side stream, forked here and joined where its results are consumed: (1) the pre-split weight images of the
    // tensor-core dense layers (they depend on the parameters only), (2) the per-batch CSC of X for the dW1 gather
    // (it depends on the batch only).  Both run beside the FC1 SpMM / the dense stack.
    cudaStream_t aux = timing ? main_st : t->side;
    if (!timing && (tc || fork_csc)) {
        CUDA_TRY(cudaEventRecord(t->ev_fork, main_st));
        CUDA_TRY(cudaStreamWaitEvent(t->side, t->ev_fork, 0));
    }
    auto build_images = [&]() -> int {
        for (int l = 2; l <= n; ++l) {
            const std::string ls = std::to_string(l);
            TRY(dssm_fc_tc_build_image(t->P_("W" + ls), t->L[l - 1], t->L[l], 0, R, t->img_fwd[l], (dssm_stream_t)aux));
            if (train) TRY(dssm_fc_tc_build_image(t->P_("W" + ls), t->L[l - 1], t->L[l], 1, R, t->img_dx[l], (dssm_stream_t)aux));
        }
        return DSSM_OK;
    };
    if (tc && !timing) {
        TRY(build_images());
        CUDA_TRY(cudaEventRecord(t->ev_img, t->side));
        t->img_forked = true;
    }
    TRY(dssm_spmm_fwd(indptr, indices, values, R, t->D, t->P_("W1"), t->P_("b1"), t->L[1], t->h[1], s));
    if (fork_csc) {
        // forked AFTER the FC1 SpMM: that kernel is bandwidth-bound and loses what the CSC build takes from it, while the
        // dense layers that follow are latency-bound and leave most of the memory system idle
        CUDA_TRY(cudaEventRecord(t->ev_fork, main_st));
        CUDA_TRY(cudaStreamWaitEvent(t->side, t->ev_fork, 0));
        g_csc_hist_done_event = t->fuse_w1_adam ? t->ev_hist : nullptr;
        const int rc_build = dssm_spmm_bwd_csc_build(indptr, indices, values, R, t->D, t->L[1],
                                                     (t->grads_p && !t->fuse_w1_adam && !t->w1_push) ? t->G_("W1") : nullptr, t->sp_ws,
                                                     t->sp_ws_bytes, (dssm_stream_t)t->side);
        g_csc_hist_done_event = nullptr;
        TRY(rc_build);
        if (t->fuse_w1_adam) {  // the rows of W1 this batch does not touch take their Adam step under the dense layers
            const int64_t wo = t->params[t->find(t->params, "W1")].off;
            // Measured (profiles/r2_absent_adam_placement.txt): this kernel streams 0.3 GB and slows whatever latency-bound
            // kernel runs beside it; started right after the histogram it sat on the first GEMM (+25..50 us on fc_fwd2),
            // started after the CSC build it costs ~10 us spread over cos/loss and the first backward GEMMs.  So it waits
            // for the end of the build (DSSM_ABSENT_EARLY=1 keeps the early start for experiments).
            if (!getenv("DSSM_ABSENT_EARLY")) CUDA_TRY(cudaEventRecord(t->ev_hist, t->side));
            CUDA_TRY(cudaStreamWaitEvent(t->side2, t->ev_hist, 0));
            TRY(dssm_spmm_bwd_adam_absent(t->D, t->L[1], t->params_p + wo, t->m_p + wo, t->v_p + wo, t->beta_pow_p, c.learning_rate,
                                          c.beta1, c.beta2, c.adam_eps, t->sp_ws, t->sp_ws_bytes, (dssm_stream_t)t->side2));
            CUDA_TRY(cudaEventRecord(t->ev_join2, t->side2));
            t->absent_forked = true;
        }
        CUDA_TRY(cudaEventRecord(t->ev_join, t->side));
        t->csc_forked = true;
    }
    mark(PH_SPMM_FWD);
    markf("spmm_fwd");
    if (tc && timing) TRY(build_images());  // profiled step: serial, accounted to the dense forward
    if (t->img_forked) CUDA_TRY(cudaStreamWaitEvent(main_st, t->ev_img, 0));
    bool stats_fused = false;  // BN moments of h[l] already taken (and finalized) by the GEMM that produced it
    for (int l = 1; l <= n; ++l) {
        const std::string ls = std::to_string(l);
        if (c.use_bn) {
            const bool sync = t->sync_n > 1 && on_train;  // the shadows then move with the GLOBAL moments, in the exchange kernel
            if (!stats_fused)
                TRY(dssm_bn_forward(t->h[l], R, t->L[l], B, on_train, sync ? 0 : update_ema, t->P_("bn" + ls + "_gamma"),
                                    t->P_("bn" + ls + "_beta"), t->E_("bn" + ls + "_ema_mean"), t->E_("bn" + ls + "_ema_var"),
                                    c.bn_eps, c.ema_decay, t->bn_mean[l], t->bn_var[l], t->bn_rstd[l], t->bn_scale[l],
                                    t->bn_shift[l], t->bn_ws, t->bn_ws_bytes, s));
            if (sync)
                TRY(dssm_syncbn_forward(t->sync_bufs, t->sync_n, t->sync_rank, l - 1, max_width(t), t->L[l], t->P_("bn" + ls + "_gamma"),
                                        t->P_("bn" + ls + "_beta"), t->E_("bn" + ls + "_ema_mean"), t->E_("bn" + ls + "_ema_var"),
                                        t->bn_mean[l], t->bn_var[l], t->bn_rstd[l], t->bn_scale[l], t->bn_shift[l], c.bn_eps, c.ema_decay,
                                        update_ema, s));
            markf("bn_fwd" + ls);
        }
        stats_fused = false;
        const float* sc = c.use_bn ? t->bn_scale[l] : nullptr;
        const float* sh = c.use_bn ? t->bn_shift[l] : nullptr;
        if (l < n) {
            const std::string ns = std::to_string(l + 1);
            if (tc && t->L[l] % 4 == 0 && t->L[l + 1] % 4 == 0) {
                // training under BN: the epilogue of this GEMM also takes the batch moments of h[l+1] (both instances) and its
                // last CTAs finalize them -- the separate bn_stats launch of the next layer disappears.  Needs B % 128 == 0
                // (an M tile never straddles the query / doc boundary) and the warp-specialised kernel.
                FusedBnStats fb{};
                const bool fuse = c.use_bn && on_train && B % 128 == 0 && t->L[l] <= 512 - 32 && !timing && R <= 128 * 64 &&
                                  getenv("DSSM_FUSED_BN") != nullptr;  // R bound: the finalize stages 3 x M tiles x BN floats in smem
                // OFF by default -- measured (profiles/r2_fused_bn_epilogue.txt): the moments cost the epilogue +3 us, but the
                // finalize by the LAST CTA of an N tile is a serial tail of 7-9 us (one L2 round trip to pull 48 tile triples,
                // a 40-deep Chan merge per column), which is what the stand-alone bn_stats launch costs.  C2 step 0.354 ms
                // with it, 0.351 ms without.  DSSM_FUSED_BN=1 turns it on (tests/test_gpu_tower.py runs both).
                if (fuse) {
                    const std::string pre = "bn" + ns + "_";
                    const bool sync = t->sync_n > 1;
                    fb.part = reinterpret_cast<float*>(t->fbn_ws + 256);
                    fb.tickets = reinterpret_cast<int*>(t->fbn_ws);
                    fb.fin = BnFinalize{t->P_(pre + "gamma"), t->P_(pre + "beta"), t->E_(pre + "ema_mean"), t->E_(pre + "ema_var"),
                                        t->bn_mean[l + 1], t->bn_var[l + 1], t->bn_rstd[l + 1], t->bn_scale[l + 1], t->bn_shift[l + 1],
                                        c.bn_eps, c.ema_decay, sync ? 0 : update_ema, B / 128};
                    fb.on = 1;
                }
                TRY(dssm_fc_fwd_tc_img(t->h[l], R, t->L[l], B, sc, sh, c.act, t->img_fwd[l + 1], t->P_("b" + ns), t->L[l + 1],
                                       t->h[l + 1], tc_passes_of(c.gemm_mode), fuse ? &fb : nullptr, s));
                stats_fused = fuse;
            } else {
                TRY(dssm_fc_fwd(t->h[l], R, t->L[l], B, sc, sh, c.act, t->P_("W" + ns), t->P_("b" + ns), t->L[l + 1],
                                t->h[l + 1], c.gemm_mode, t->fc_ws, t->fc_ws_bytes, s));
            }
            markf("fc_fwd" + ns);
        } else {
            mark(PH_DENSE_FWD);
            // last layer: BN + activation fused into the cosine / loss kernel, which also writes the embeddings Y
            TRY(dssm_cos_softmax_loss_fused(t->h[l], sc, sh, c.act, t->Y, B, t->NEG, t->L[n], c.gamma, c.loss_eps, c.loss_div_bs,
                                            t->qnorm, t->dnorm, t->cos_raw, t->cos_sim, t->prob, t->loss_terms, t->loss,
                                            want_grad ? t->dh[n] : nullptr, s));
            mark(PH_COSLOSS);
            markf("cos_loss");
        }
    }
    t->cur_indptr = indptr; t->cur_indices = indices; t->cur_values = values;
    t->fwd_train_done = want_grad && on_train;
    return DSSM_OK;
}

static void w1_chunk_cols(const dssm_tower* t, int chunk, int n_chunks, int* c0, int* c1) {
    const int per = ((t->D + n_chunks - 1) / n_chunks + 3) / 4 * 4;  // multiple of 4 columns: float4-aligned slices
    *c0 = chunk * per < t->D ? chunk * per : t->D;
    *c1 = *c0 + per < t->D ? *c0 + per : t->D;
}

// w1_mode: 0 = whole dW1 here, 1 = stop after the CSC build (dssm_tower_backward_w1 produces dW1 chunk by chunk)
static int tower_backward_impl(dssm_tower* t, dssm_stream_t s, int w1_mode = 0) {
    const dssm_config& c = t->cfg;
    const int n = t->n_layers, R = t->R, B = t->B;
    // dW GEMMs whose launch is deferred to the start of the dW1 gather (see fork_dw below)
    int late_dw[DSSM_MAX_LAYERS + 1], n_late = 0;
    const bool dw_late = getenv("DSSM_DW_LATE") != nullptr;
    auto launch_dw = [&](int l, dssm_stream_t st_dw) -> int {
        const std::string ls = std::to_string(l);
        const float* sc = c.use_bn ? t->bn_scale[l - 1] : nullptr;
        const float* sh = c.use_bn ? t->bn_shift[l - 1] : nullptr;
        return dssm_fc_bwd_dw(t->h[l - 1], R, t->L[l - 1], B, sc, sh, c.act, t->dh[l], t->L[l], t->G_("W" + ls),
                              c.use_bn ? nullptr : t->G_("b" + ls), c.gemm_mode, t->dw_ws, t->dw_ws_bytes, st_dw);
    };
    for (int l = n; l >= 1; --l) {
        const std::string ls = std::to_string(l);
        if (c.use_bn && t->sync_n > 1) {
            // SyncBN: local column sums -> average over the replicas (and this replica's db) -> dH with the global sums
            TRY(dssm_bn_bwd_reduce_only(t->dh[l], t->h[l], R, t->L[l], B, c.act, t->P_("bn" + ls + "_gamma"), t->bn_mean[l], t->bn_rstd[l],
                                        t->bn_scale[l], t->bn_shift[l], t->G_("bn" + ls + "_gamma"), t->G_("bn" + ls + "_beta"), nullptr,
                                        t->bn_sumx[l], t->bn_ws, t->bn_ws_bytes, s));
            TRY(dssm_syncbn_backward(t->sync_bufs, t->sync_n, t->sync_rank, n + l - 1, max_width(t), t->L[l], B, R - B,
                                     t->P_("bn" + ls + "_gamma"), t->bn_rstd[l], t->bn_sumx[l], t->G_("bn" + ls + "_gamma"),
                                     t->G_("bn" + ls + "_beta"), t->G_("b" + ls), s));
            TRY(dssm_bn_bwd_apply_only(t->dh[l], t->h[l], R, t->L[l], B, c.act, t->P_("bn" + ls + "_gamma"), t->bn_mean[l], t->bn_rstd[l],
                                       t->bn_scale[l], t->bn_shift[l], t->G_("bn" + ls + "_gamma"), t->G_("bn" + ls + "_beta"), s));
        } else if (c.use_bn) {
            TRY(dssm_bn_act_backward(t->dh[l], t->h[l], R, t->L[l], B, c.act, t->P_("bn" + ls + "_gamma"), t->bn_mean[l],
                                     t->bn_rstd[l], t->bn_scale[l], t->bn_shift[l], t->G_("bn" + ls + "_gamma"),
                                     t->G_("bn" + ls + "_beta"), t->G_("b" + ls), t->bn_ws, t->bn_ws_bytes, s));
        } else {
            TRY(dssm_bn_act_backward(t->dh[l], t->h[l], R, t->L[l], B, c.act, nullptr, nullptr, nullptr, nullptr, nullptr,
                                     nullptr, nullptr, nullptr, nullptr, 0, s));
        }
        markf("bn_bwd" + ls);
        if (l > 1) {
            const float* sc = c.use_bn ? t->bn_scale[l - 1] : nullptr;
            const float* sh = c.use_bn ? t->bn_shift[l - 1] : nullptr;
            // under BN the bias gradient came out of dssm_bn_act_backward; without BN it is the column sum of dH
            // dW_l needs dh_l and h_{l-1} only: under BN (no colsum sharing its workspace) it runs on side3 beside the dX chain.
            // The dX GEMM -- the one on the critical path -- is issued FIRST, so its CTAs get the SMs first.
            // (the in-graph %globaltimer timeline keeps the fork: its stamps sit between the main-stream calls only)
            const bool fork_dw = c.use_bn && !(g_timer && g_timer->on && !g_timer->stamps) && getenv("DSSM_NO_DW_FORK") == nullptr;
            dssm_stream_t dw_st = s;
            if (fork_dw && !dw_late) {
                CUDA_TRY(cudaEventRecord(t->ev_dh, (cudaStream_t)s));
                CUDA_TRY(cudaStreamWaitEvent(t->side3, t->ev_dh, 0));
                dw_st = (dssm_stream_t)t->side3;
                t->dw_forked = true;
            }
            auto issue_dw = [&]() -> int {
                if (fork_dw && dw_late) {  // launched on side3 when the dX chain is through (beside the dW1 gather)
                    late_dw[n_late++] = l;
                    return DSSM_OK;
                }
                TRY(launch_dw(l, dw_st));
                if (!fork_dw) markf("fc_dw" + ls);
                return DSSM_OK;
            };
            if (!fork_dw) TRY(issue_dw());  // same stream: dW first (the profiled timeline keeps its order)
            if (is_tc_mode(c.gemm_mode) && t->img_dx[l] && t->L[l] % 4 == 0 && t->L[l - 1] % 4 == 0) {
                TRY(dssm_fc_bwd_dx_tc_img(t->dh[l], R, t->L[l], t->img_dx[l], t->L[l - 1], t->dh[l - 1], tc_passes_of(c.gemm_mode), s));
            } else {
                TRY(dssm_fc_bwd_dx(t->dh[l], R, t->L[l], t->P_("W" + ls), t->L[l - 1], t->dh[l - 1], c.gemm_mode, t->fc_ws,
                                   t->fc_ws_bytes, s));
            }
            markf("fc_dx" + ls);
            if (fork_dw) TRY(issue_dw());
        } else {
            mark(PH_DENSE_BWD);
            if (n_late > 0) {
                CUDA_TRY(cudaEventRecord(t->ev_dh, (cudaStream_t)s));
                CUDA_TRY(cudaStreamWaitEvent(t->side3, t->ev_dh, 0));
                for (int i = 0; i < n_late; ++i) TRY(launch_dw(late_dw[i], (dssm_stream_t)t->side3));
                t->dw_forked = true;
            }
            if (t->csc_forked) {  // join: the CSC built beside the forward is ready (or will be) -- gather only
                CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)s, t->ev_join, 0));
                if (t->absent_forked) CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)s, t->ev_join2, 0));
                t->absent_forked = false;
                t->csc_forked = false;
                mark(PH_CSC_BUILD);  // overlapped profile: time the main stream waited for the side stream
                markf("join_side_stream");
                if (w1_mode == 0 && t->fuse_w1_adam) {
                    const int wi = t->find(t->params, "W1");
                    const int64_t wo = t->params[wi].off;
                    TRY(dssm_spmm_bwd_dw_adam(t->dh[1], R, t->D, t->L[1], t->params_p + wo, t->m_p + wo, t->v_p + wo, t->beta_pow_p,
                                              c.learning_rate, c.beta1, c.beta2, c.adam_eps, 1, t->sp_ws, t->sp_ws_bytes, s));
                } else if (w1_mode == 0) {
                    TRY(dssm_spmm_bwd_dw_range(t->dh[1], R, t->D, t->L[1], t->G_("W1"), 0, t->D, 0, t->sp_ws, t->sp_ws_bytes, s));
                }
            } else if (w1_mode == 1) {
                DSSM_REQUIRE(t->L[1] % 4 == 0 && t->L[1] <= 1024, DSSM_ERR_BAD_SHAPE, "chunked dW1 needs L1 %% 4 == 0");
                TRY(dssm_spmm_bwd_csc_build(t->cur_indptr, t->cur_indices, t->cur_values, R, t->D, t->L[1], t->G_("W1"), t->sp_ws,
                                            t->sp_ws_bytes, s));
            } else {
                if (g_timer && g_timer->on && g_timer->serial) g_spmm_bwd_mid_event = g_timer->ev[PH_CSC_BUILD];
                const int rc = dssm_spmm_bwd_dw(t->cur_indptr, t->cur_indices, t->cur_values, R, t->D, t->dh[1], t->L[1],
                                                t->G_("W1"), 0, t->sp_ws, t->sp_ws_bytes, s);
                g_spmm_bwd_mid_event = nullptr;
                TRY(rc);
            }
            mark(PH_DW_GATHER);
            markf("dw1_gather");
            if (!c.use_bn) TRY(dssm_colsum(t->dh[1], R, t->L[1], t->G_("b1"), t->dw_ws, t->dw_ws_bytes, s));
            mark(PH_B1);
        }
    }
    if (t->dw_forked) {  // the dW GEMMs on side3 are done before anything downstream (Adam, gradient exchange) reads them
        CUDA_TRY(cudaEventRecord(t->ev_join3, t->side3));
        CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)s, t->ev_join3, 0));
        t->dw_forked = false;
    }
    return DSSM_OK;
}

static int tower_adam_impl(dssm_tower* t, float grad_scale, dssm_stream_t s) {
    const dssm_config& c = t->cfg;
    TRY(dssm_adam_step(t->params_p, t->grads_p, t->m_p, t->v_p, t->P, t->beta_pow_p, c.learning_rate, c.beta1, c.beta2,
                       c.adam_eps, grad_scale, s));
    TRY(dssm_adam_advance(t->beta_pow_p, c.beta1, c.beta2, s));
    return DSSM_OK;
}

struct LaunchScope {
    dssm_tower* t;
    int64_t start;
    explicit LaunchScope(dssm_tower* t_) : t(t_), start(g_launch_count) {}
    ~LaunchScope() { t->launches += g_launch_count - start; }
};

extern "C" int dssm_tower_forward(dssm_tower* t, const int32_t* indptr, const int32_t* indices, const float* values,
                                  int32_t on_train, int32_t update_ema, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_forward: tower not bound");
    DSSM_REQUIRE(indptr && indices && values, DSSM_ERR_BAD_ARG, "dssm_tower_forward: null CSR pointer");
    LaunchScope ls(t);
    return tower_forward_impl(t, indptr, indices, values, on_train, update_ema, on_train != 0, stream);
}

extern "C" int dssm_tower_backward(dssm_tower* t, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_backward: tower not bound");
    DSSM_REQUIRE(t->grads_p, DSSM_ERR_STATE, "dssm_tower_backward: no grads buffer bound");
    DSSM_REQUIRE(t->fwd_train_done, DSSM_ERR_STATE, "dssm_tower_backward: needs a preceding training-mode forward");
    LaunchScope ls(t);
    t->fwd_train_done = false;
    return tower_backward_impl(t, stream);
}

extern "C" int dssm_tower_backward_begin(dssm_tower* t, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_backward_begin: tower not bound");
    DSSM_REQUIRE(t->grads_p, DSSM_ERR_STATE, "dssm_tower_backward_begin: no grads buffer bound");
    DSSM_REQUIRE(t->fwd_train_done, DSSM_ERR_STATE, "dssm_tower_backward_begin: needs a preceding training-mode forward");
    LaunchScope ls(t);
    t->fwd_train_done = false;
    return tower_backward_impl(t, stream, 1);
}

extern "C" int dssm_tower_w1_chunk(const dssm_tower* t, int32_t chunk, int32_t n_chunks, int64_t* offset_floats, int64_t* count_floats) {
    DSSM_REQUIRE(t && n_chunks > 0 && n_chunks <= 64 && chunk >= 0 && chunk < n_chunks, DSSM_ERR_BAD_ARG, "dssm_tower_w1_chunk: bad chunk");
    int c0, c1;
    w1_chunk_cols(t, chunk, n_chunks, &c0, &c1);
    if (offset_floats) *offset_floats = t->params[t->find(t->params, "W1")].off + (int64_t)c0 * t->L[1];
    if (count_floats) *count_floats = (int64_t)(c1 - c0) * t->L[1];
    return DSSM_OK;
}

extern "C" int dssm_tower_backward_w1(dssm_tower* t, int32_t chunk, int32_t n_chunks, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound && t->grads_p, DSSM_ERR_STATE, "dssm_tower_backward_w1: tower not bound");
    DSSM_REQUIRE(n_chunks > 0 && n_chunks <= 64 && chunk >= 0 && chunk < n_chunks, DSSM_ERR_BAD_ARG, "dssm_tower_backward_w1: bad chunk");
    LaunchScope ls(t);
    int c0, c1;
    w1_chunk_cols(t, chunk, n_chunks, &c0, &c1);
    return dssm_spmm_bwd_dw_range(t->dh[1], t->R, t->D, t->L[1], t->G_("W1"), c0, c1, chunk, t->sp_ws, t->sp_ws_bytes, stream);
}

extern "C" int dssm_tower_set_w1_push(dssm_tower* t, int32_t enabled) {
    DSSM_REQUIRE(t, DSSM_ERR_BAD_ARG, "dssm_tower_set_w1_push: null tower");
    if (t->graph_dp_exec) { cudaGraphExecDestroy(t->graph_dp_exec); t->graph_dp_exec = nullptr; }
    if (t->graph_dp) { cudaGraphDestroy(t->graph_dp); t->graph_dp = nullptr; }
    t->w1_push = enabled != 0;
    return DSSM_OK;
}

extern "C" int dssm_tower_backward_w1_push(dssm_tower* t, float* const* host_peer_slots, uint32_t* const* host_peer_valid,
                                           const uint32_t* epoch, int32_t n_ranks, int32_t self, int32_t per, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound && t->grads_p, DSSM_ERR_STATE, "dssm_tower_backward_w1_push: tower not bound");
    DSSM_REQUIRE(t->w1_push, DSSM_ERR_STATE, "dssm_tower_backward_w1_push: call dssm_tower_set_w1_push(t, 1) first");
    LaunchScope ls(t);
    return dssm_spmm_bwd_dw_push(t->dh[1], t->R, t->D, t->L[1], host_peer_slots, host_peer_valid, epoch, n_ranks, self, per, t->sp_ws,
                                 t->sp_ws_bytes, stream);
}

extern "C" int dssm_tower_adam_range(dssm_tower* t, int64_t offset_floats, int64_t count_floats, float grad_scale,
                                     dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_adam_range: tower not bound");
    DSSM_REQUIRE(t->grads_p && t->m_p && t->v_p && t->beta_pow_p, DSSM_ERR_STATE, "dssm_tower_adam_range: optimizer buffers not bound");
    DSSM_REQUIRE(offset_floats >= 0 && count_floats >= 0 && offset_floats + count_floats <= t->P && offset_floats % 4 == 0,
                 DSSM_ERR_BAD_ARG, "dssm_tower_adam_range: bad range");
    LaunchScope ls(t);
    const dssm_config& c = t->cfg;
    return dssm_adam_step(t->params_p + offset_floats, t->grads_p + offset_floats, t->m_p + offset_floats, t->v_p + offset_floats,
                          count_floats, t->beta_pow_p, c.learning_rate, c.beta1, c.beta2, c.adam_eps, grad_scale, stream);
}

extern "C" int dssm_tower_adam_advance(dssm_tower* t, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound && t->beta_pow_p, DSSM_ERR_STATE, "dssm_tower_adam_advance: tower not bound");
    LaunchScope ls(t);
    return dssm_adam_advance(t->beta_pow_p, t->cfg.beta1, t->cfg.beta2, stream);
}

extern "C" int dssm_tower_adam(dssm_tower* t, float grad_scale, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_adam: tower not bound");
    DSSM_REQUIRE(t->grads_p && t->m_p && t->v_p && t->beta_pow_p, DSSM_ERR_STATE, "dssm_tower_adam: optimizer buffers not bound");
    LaunchScope ls(t);
    return tower_adam_impl(t, grad_scale, stream);
}

static int tower_step_impl(dssm_tower* t, const int32_t* indptr, const int32_t* indices, const float* values,
                           dssm_stream_t s) {
    const dssm_config& c = t->cfg;
    // single-GPU step: W1's Adam update rides on the dW1 gather (the CSC is built beside the forward) unless the
    // step is being phase-profiled or FC1's width rules out the vector kernels
    const int wi = t->find(t->params, "W1");
    const bool fuse = t->L[1] % 4 == 0 && t->L[1] <= 1024 && !(g_timer && g_timer->on && g_timer->serial) && t->params[wi].off == 0;
    t->fuse_w1_adam = fuse;
    int rc = tower_forward_impl(t, indptr, indices, values, 1, 1, true, s);
    if (rc == DSSM_OK) rc = tower_backward_impl(t, s);
    t->fuse_w1_adam = false;
    t->fwd_train_done = false;
    if (rc != DSSM_OK) return rc;
    if (!fuse) return tower_adam_impl(t, 1.0f, s);
    // the rest of the flat buffer (everything behind W1), then the beta powers
    const int64_t w1_end = pad4((int64_t)t->D * t->L[1]);
    TRY(dssm_adam_step(t->params_p + w1_end, t->grads_p + w1_end, t->m_p + w1_end, t->v_p + w1_end, t->P - w1_end, t->beta_pow_p,
                       c.learning_rate, c.beta1, c.beta2, c.adam_eps, 1.0f, s));
    TRY(dssm_adam_advance(t->beta_pow_p, c.beta1, c.beta2, s));
    return DSSM_OK;
}

extern "C" int dssm_tower_train_step(dssm_tower* t, const int32_t* indptr, const int32_t* indices, const float* values,
                                     dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_train_step: tower not bound");
    DSSM_REQUIRE(t->grads_p && t->m_p && t->v_p && t->beta_pow_p, DSSM_ERR_STATE, "dssm_tower_train_step: optimizer buffers not bound");
    DSSM_REQUIRE(indptr && indices && values, DSSM_ERR_BAD_ARG, "dssm_tower_train_step: null CSR pointer");
    LaunchScope ls(t);
    return tower_step_impl(t, indptr, indices, values, stream);
}

extern "C" int dssm_tower_staging(dssm_tower* t, int32_t** indptr, int32_t** indices, float** values) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_staging: tower not bound");
    if (indptr) *indptr = t->st_indptr;
    if (indices) *indices = t->st_indices;
    if (values) *values = t->st_values;
    return DSSM_OK;
}

extern "C" int dssm_tower_capture_graph(dssm_tower* t, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_capture_graph: tower not bound");
    DSSM_REQUIRE(t->grads_p && t->m_p && t->v_p && t->beta_pow_p, DSSM_ERR_STATE, "dssm_tower_capture_graph: optimizer buffers not bound");
    cudaStream_t st = (cudaStream_t)stream;
    DSSM_REQUIRE(st != nullptr, DSSM_ERR_BAD_ARG, "dssm_tower_capture_graph: needs a non-default stream");
    if (t->graph_exec) { cudaGraphExecDestroy(t->graph_exec); t->graph_exec = nullptr; }
    if (t->graph) { cudaGraphDestroy(t->graph); t->graph = nullptr; }
    const int64_t before = g_launch_count;
    // captured on the tower's own most-urgent-priority stream (nothing executes during capture; the graph is launched on
    // the caller's stream): the main chain's kernel nodes outrank the side streams' (see dssm_tower_bind)
    st = t->main_hi;
    CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = tower_step_impl(t, t->st_indptr, t->st_indices, t->st_values, (dssm_stream_t)st);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc != DSSM_OK) {
        if (g) cudaGraphDestroy(g);
        return rc;
    }
    if (e != cudaSuccess) return fail(DSSM_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    t->launches_per_step = g_launch_count - before;
    t->graph = g;
    CUDA_TRY(cudaGraphInstantiate(&t->graph_exec, t->graph, 0));
    return DSSM_OK;
}

// Data-parallel pipeline, first half: training forward + backward_begin on the staging CSR, as one CUDA graph.
extern "C" int dssm_tower_capture_graph_dp(dssm_tower* t, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound && t->grads_p, DSSM_ERR_STATE, "dssm_tower_capture_graph_dp: tower not bound");
    cudaStream_t st = (cudaStream_t)stream;
    DSSM_REQUIRE(st != nullptr, DSSM_ERR_BAD_ARG, "dssm_tower_capture_graph_dp: needs a non-default stream");
    if (t->graph_dp_exec) { cudaGraphExecDestroy(t->graph_dp_exec); t->graph_dp_exec = nullptr; }
    if (t->graph_dp) { cudaGraphDestroy(t->graph_dp); t->graph_dp = nullptr; }
    const int64_t before = g_launch_count;
    st = t->main_hi;  // as in dssm_tower_capture_graph
    CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = tower_forward_impl(t, t->st_indptr, t->st_indices, t->st_values, 1, 1, true, (dssm_stream_t)st);
    if (rc == DSSM_OK) rc = tower_backward_impl(t, (dssm_stream_t)st, 1);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(st, &g);
    t->fwd_train_done = false;
    if (rc != DSSM_OK) {
        if (g) cudaGraphDestroy(g);
        return rc;
    }
    if (e != cudaSuccess) return fail(DSSM_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    t->launches_per_dp = g_launch_count - before;
    t->graph_dp = g;
    CUDA_TRY(cudaGraphInstantiate(&t->graph_dp_exec, t->graph_dp, 0));
    return DSSM_OK;
}

// forward (training) + backward_begin on the staging CSR; graph replay when captured
extern "C" int dssm_tower_fwd_bwd_begin_staged(dssm_tower* t, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound && t->grads_p, DSSM_ERR_STATE, "dssm_tower_fwd_bwd_begin_staged: tower not bound");
    if (t->graph_dp_exec) {
        CUDA_TRY(cudaGraphLaunch(t->graph_dp_exec, (cudaStream_t)stream));
        t->launches += t->launches_per_dp;
        t->cur_indptr = t->st_indptr; t->cur_indices = t->st_indices; t->cur_values = t->st_values;
        return DSSM_OK;
    }
    LaunchScope ls(t);
    TRY(tower_forward_impl(t, t->st_indptr, t->st_indices, t->st_values, 1, 1, true, stream));
    t->fwd_train_done = false;
    return tower_backward_impl(t, stream, 1);
}

extern "C" int dssm_tower_train_step_staged(dssm_tower* t, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_train_step_staged: tower not bound");
    if (t->graph_exec) {
        CUDA_TRY(cudaGraphLaunch(t->graph_exec, (cudaStream_t)stream));
        t->launches += t->launches_per_step;
        return DSSM_OK;
    }
    return dssm_tower_train_step(t, t->st_indptr, t->st_indices, t->st_values, stream);
}

extern "C" int dssm_tower_train_step_host(dssm_tower* t, const int32_t* host_indptr, const int32_t* host_indices,
                                          const float* host_values, int64_t nnz, float* host_loss, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_train_step_host: tower not bound");
    DSSM_REQUIRE(host_indptr && host_indices && host_values, DSSM_ERR_BAD_ARG, "dssm_tower_train_step_host: null CSR pointer");
    DSSM_REQUIRE(nnz >= 0 && nnz <= t->max_nnz, DSSM_ERR_WORKSPACE, "dssm_tower_train_step_host: nnz %lld exceeds bound max_nnz %lld",
                 (long long)nnz, (long long)t->max_nnz);
    DSSM_REQUIRE(host_indptr[0] == 0 && host_indptr[t->R] == nnz, DSSM_ERR_BAD_SHAPE,
                 "dssm_tower_train_step_host: batch must have exactly (2+NEG)*query_BS = %d rows and indptr[R] == nnz", t->R);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(t->st_indptr, host_indptr, (size_t)(t->R + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (nnz > 0) {
        CUDA_TRY(cudaMemcpyAsync(t->st_indices, host_indices, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(t->st_values, host_values, (size_t)nnz * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    TRY(dssm_tower_train_step_staged(t, stream));
    if (host_loss) {
        CUDA_TRY(cudaMemcpyAsync(host_loss, t->loss, sizeof(float), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return DSSM_OK;
}

// Pipelined host feed, step k (k counts uploads since bind): the CSR travels pinned host -> upload buffer k%2 on the
// tower's copy stream while the previous step still computes; on `stream` it is then moved into the staging CSR (the
// train-step graphs read fixed addresses).  Nothing synchronises here.
extern "C" int64_t dssm_tower_feed_upload_async(dssm_tower* t, const int32_t* host_indptr, const int32_t* host_indices,
                                                const float* host_values, int64_t nnz, dssm_stream_t stream) {
    if (!(t && t->bound)) { fail(DSSM_ERR_STATE, "dssm_tower_feed_upload_async: tower not bound"); return -1; }
    if (!(host_indptr && host_indices && host_values)) { fail(DSSM_ERR_BAD_ARG, "dssm_tower_feed_upload_async: null CSR pointer"); return -1; }
    if (!(nnz >= 0 && nnz <= t->max_nnz)) {
        fail(DSSM_ERR_WORKSPACE, "dssm_tower_feed_upload_async: nnz %lld exceeds bound max_nnz %lld", (long long)nnz, (long long)t->max_nnz);
        return -1;
    }
    if (!(host_indptr[0] == 0 && host_indptr[t->R] == nnz)) {
        fail(DSSM_ERR_BAD_SHAPE, "dssm_tower_feed_upload_async: batch must have exactly (2+NEG)*query_BS = %d rows and indptr[R] == nnz", t->R);
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t k = t->feed_k;
    const int b = (int)(k & 1);
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
    if (k >= 2) ok(cudaStreamWaitEvent(t->copy, t->ev_up_free[b], 0));  // step k-2 has drained this upload buffer
    ok(cudaMemcpyAsync(t->up_indptr[b], host_indptr, (size_t)(t->R + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, t->copy));
    if (nnz > 0) {
        ok(cudaMemcpyAsync(t->up_indices[b], host_indices, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, t->copy));
        ok(cudaMemcpyAsync(t->up_values[b], host_values, (size_t)nnz * sizeof(float), cudaMemcpyHostToDevice, t->copy));
    }
    ok(cudaEventRecord(t->ev_up_ready[b], t->copy));
    ok(cudaStreamWaitEvent(st, t->ev_up_ready[b], 0));
    ok(cudaMemcpyAsync(t->st_indptr, t->up_indptr[b], (size_t)(t->R + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (nnz > 0) {
        ok(cudaMemcpyAsync(t->st_indices, t->up_indices[b], (size_t)nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        ok(cudaMemcpyAsync(t->st_values, t->up_values[b], (size_t)nnz * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    ok(cudaEventRecord(t->ev_up_free[b], st));
    if (e != cudaSuccess) { fail(DSSM_ERR_CUDA, "dssm_tower_feed_upload_async: %s", cudaGetErrorString(e)); return -1; }
    t->feed_k = k + 1;
    return k;
}

// Closes pipelined step `step` (the value dssm_tower_feed_upload_async returned) after the caller has enqueued the work
// that consumes the staging CSR: copies the loss to host_loss (may be NULL) and records the step's completion event.
extern "C" int dssm_tower_feed_step_done(dssm_tower* t, int64_t step, float* host_loss, dssm_stream_t stream) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_feed_step_done: tower not bound");
    DSSM_REQUIRE(step == t->feed_k - 1, DSSM_ERR_BAD_ARG, "dssm_tower_feed_step_done: step %lld is not the last upload (%lld)",
                 (long long)step, (long long)(t->feed_k - 1));
    cudaStream_t st = (cudaStream_t)stream;
    if (host_loss) CUDA_TRY(cudaMemcpyAsync(host_loss, t->loss, sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(t->ev_step_done[step & 1], st));
    return DSSM_OK;
}

// upload + single-GPU train step + loss read-back, all asynchronous; returns the step id (< 0 on error)
extern "C" int64_t dssm_tower_train_step_host_async(dssm_tower* t, const int32_t* host_indptr, const int32_t* host_indices,
                                                    const float* host_values, int64_t nnz, float* host_loss, dssm_stream_t stream) {
    const int64_t k = dssm_tower_feed_upload_async(t, host_indptr, host_indices, host_values, nnz, stream);
    if (k < 0) return k;
    if (dssm_tower_train_step_staged(t, stream) != DSSM_OK) return -1;
    if (dssm_tower_feed_step_done(t, k, host_loss, stream) != DSSM_OK) return -1;
    return k;
}

// Blocks until pipelined step `step` (a value returned by dssm_tower_train_step_host_async; only the last two are
// waitable) has finished: its host_loss is valid and its host CSR buffers may be reused.
extern "C" int dssm_tower_feed_wait(dssm_tower* t, int64_t step) {
    DSSM_REQUIRE(t && t->bound, DSSM_ERR_STATE, "dssm_tower_feed_wait: tower not bound");
    DSSM_REQUIRE(step >= 0 && step < t->feed_k && step >= t->feed_k - 2, DSSM_ERR_BAD_ARG,
                 "dssm_tower_feed_wait: step %lld is not one of the last two issued (%lld issued)", (long long)step, (long long)t->feed_k);
    CUDA_TRY(cudaEventSynchronize(t->ev_step_done[step & 1]));
    return DSSM_OK;
}

extern "C" int64_t dssm_tower_launch_count(const dssm_tower* t) { return t ? t->launches : -1; }

static int profile_step_impl(dssm_tower* t, float* host_phase_ms, bool serial, dssm_stream_t stream,
                             std::vector<cudaEvent_t>* fine_ev = nullptr, std::vector<std::string>* fine_name = nullptr) {
    DSSM_REQUIRE(t && t->bound && host_phase_ms, DSSM_ERR_STATE, "dssm_tower_profile_step: tower not bound / null output");
    DSSM_REQUIRE(t->grads_p && t->m_p && t->v_p && t->beta_pow_p, DSSM_ERR_STATE, "dssm_tower_profile_step: optimizer buffers not bound");
    PhaseTimer pt;
    pt.st = (cudaStream_t)stream;
    pt.on = true;
    pt.serial = serial;
    pt.fine_ev = fine_ev;
    pt.fine_name = fine_name;
    pt.stamps = nullptr;
    for (int i = 0; i < PH_COUNT; ++i) CUDA_TRY(cudaEventCreate(&pt.ev[i]));
    g_timer = &pt;
    int rc;
    {
        LaunchScope ls(t);
        rc = tower_step_impl(t, t->st_indptr, t->st_indices, t->st_values, stream);
        if (rc == DSSM_OK) { mark(PH_ADAM); markf("adam_rest"); }
    }
    g_timer = nullptr;
    if (rc == DSSM_OK) {
        cudaError_t e = cudaStreamSynchronize(pt.st);
        if (e != cudaSuccess) rc = fail(DSSM_ERR_CUDA, "profile step failed: %s", cudaGetErrorString(e));
    }
    if (rc == DSSM_OK)
        for (int i = 1; i < PH_COUNT; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, pt.ev[i - 1], pt.ev[i]);
            host_phase_ms[i - 1] = ms;
        }
    for (int i = 0; i < PH_COUNT; ++i) cudaEventDestroy(pt.ev[i]);
    return rc;
}

extern "C" int dssm_tower_profile_step(dssm_tower* t, float* host_phase_ms, dssm_stream_t stream) {
    return profile_step_impl(t, host_phase_ms, true, stream);
}

// Same events around the step as it really runs (side stream, fused W1 Adam): phase i is then main-stream time between
// its boundaries, "csc_build" becomes the time the main stream waited at the join, "adam" the non-W1 parameters.
extern "C" int dssm_tower_profile_step_overlapped(dssm_tower* t, float* host_phase_ms, dssm_stream_t stream) {
    return profile_step_impl(t, host_phase_ms, false, stream);
}

// Timeline of the step as it really runs, INSIDE a CUDA graph: the step is captured with a 1-thread %globaltimer stamp
// after every call on the main stream (each costs a graph node, ~1.5 us), replayed three times, and the stamps of the
// last replay are read back.  names receives the labels joined by ';', ms[i] the time between label i-1 and label i
// (ms[0] = 0); *n_out their count.  Profiling helper: allocates its own small stamp buffer and synchronises.
extern "C" int dssm_tower_profile_timeline(dssm_tower* t, char* names, int32_t names_cap, float* ms, int32_t max_n, int32_t* n_out,
                                           dssm_stream_t stream) {
    DSSM_REQUIRE(names && ms && n_out && names_cap > 0 && max_n > 0, DSSM_ERR_BAD_ARG, "dssm_tower_profile_timeline: null output");
    DSSM_REQUIRE(t && t->bound && t->grads_p && t->m_p && t->v_p && t->beta_pow_p, DSSM_ERR_STATE, "dssm_tower_profile_timeline: tower not bound");
    cudaStream_t st = (cudaStream_t)stream;
    DSSM_REQUIRE(st != nullptr, DSSM_ERR_BAD_ARG, "dssm_tower_profile_timeline: needs a non-default stream");
    constexpr int MAX_STAMPS = 128;
    unsigned long long* d_stamps = nullptr;
    CUDA_TRY(cudaMalloc(&d_stamps, MAX_STAMPS * sizeof(unsigned long long)));
    std::vector<std::string> nm;
    PhaseTimer pt;
    cudaStream_t cap = t->main_hi;  // captured like the real step (dssm_tower_capture_graph), launched on the caller's stream
    pt.st = cap;
    pt.on = true;
    pt.serial = false;
    pt.fine_ev = nullptr;
    pt.fine_name = &nm;
    pt.stamps = d_stamps;
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    int rc = DSSM_OK;
    cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        g_timer = &pt;
        rc = tower_step_impl(t, t->st_indptr, t->st_indices, t->st_values, (dssm_stream_t)cap);
        if (rc == DSSM_OK) markf("adam_rest");
        g_timer = nullptr;
        e = cudaStreamEndCapture(cap, &g);
    }
    if (rc == DSSM_OK && e == cudaSuccess) e = cudaGraphInstantiate(&ge, g, 0);
    for (int i = 0; rc == DSSM_OK && e == cudaSuccess && i < 3; ++i) e = cudaGraphLaunch(ge, st);
    if (rc == DSSM_OK && e == cudaSuccess) e = cudaStreamSynchronize(st);
    int n = 0;
    if (rc == DSSM_OK && e == cudaSuccess) {
        std::vector<unsigned long long> h(MAX_STAMPS);
        e = cudaMemcpy(h.data(), d_stamps, MAX_STAMPS * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        std::string joined;
        for (size_t i = 0; e == cudaSuccess && i < nm.size() && n < max_n && n < MAX_STAMPS; ++i, ++n) {
            ms[n] = i ? (float)((double)(h[i] - h[i - 1]) * 1e-6) : 0.f;
            if (i) joined += ";";
            joined += nm[i];
        }
        snprintf(names, (size_t)names_cap, "%s", joined.c_str());
    }
    if (ge) cudaGraphExecDestroy(ge);
    if (g) cudaGraphDestroy(g);
    cudaFree(d_stamps);
    *n_out = n;
    if (rc != DSSM_OK) return rc;
    if (e != cudaSuccess) return fail(DSSM_ERR_CUDA, "dssm_tower_profile_timeline: %s", cudaGetErrorString(e));
    return DSSM_OK;
}
