// tcgen05 + TMA bf16 GEMM / implicit-GEMM conv (EGR_PREC_BF16 dense stage).  PLACEHOLDER: filled in next.
#include "gemm.cuh"
namespace egr {
int gemm_tc_init() { return fail(EGR_ERR_UNSUPPORTED, "tcgen05 GEMM not built yet; use egr_set_option(\"tc\", 0)"); }
int gemm_tc(const GemmDesc&, int, cudaStream_t) { return fail(EGR_ERR_UNSUPPORTED, "tcgen05 GEMM not built yet"); }
}  // namespace egr
