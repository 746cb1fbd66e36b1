// tcgen05 (bf16 operands, fp32 accumulate in TMEM) path of the dense layers -- placeholder entry point.
// The product refuses the mode loudly until the kernel lands; there is no silent fallback to the fp32 path.
#include "common.cuh"

extern "C" int dssm_fc_fwd_tc(const float*, int32_t, int32_t, int32_t, const float*, const float*, int32_t,
                              const float*, const float*, int32_t, float*, dssm_stream_t) {
    return dssm::fail(DSSM_ERR_BAD_ARG, "DSSM_GEMM_BF16_TC: tcgen05 dense-layer kernel not built in this revision");
}
