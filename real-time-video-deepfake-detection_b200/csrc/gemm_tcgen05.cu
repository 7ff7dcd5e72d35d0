// placeholder until the tcgen05 GEMM lands
#include "dfd_internal.cuh"
bool dfd_gemm_bf16_enabled() { return false; }
int dfd_gemm_bf16(dfd_ctx* ctx, const __nv_bfloat16*, const __nv_bfloat16*, const float*, const __nv_bfloat16*,
                  __nv_bfloat16*, int, int, int, int, cudaStream_t) { ctx->err = "tcgen05 GEMM not built"; return DFD_ERR_INVALID; }
void dfd_gemm_free(dfd_ctx*) {}
extern "C" int dfd_gemm_selftest(dfd_ctx* ctx, int, int, int, int, int, void*, void*) { ctx->err = "tcgen05 GEMM not built"; return DFD_ERR_INVALID; }
