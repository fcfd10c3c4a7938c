// bf16 tcgen05 implicit-GEMM path of the Network2 conv stack (placeholder until the kernel lands).
#include "snv_model.cuh"

namespace mural {
int snv_tc_prepare(mural_snv_model* m, const float* h_blob) { (void)m; (void)h_blob; return 0; }
void snv_tc_destroy(mural_snv_model* m) { (void)m; }
int snv_forward_tc(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta,
                   const uint8_t* d_sym, const int64_t* d_cat, int64_t n, float* d_logp, cudaStream_t st) {
  (void)m; (void)G; (void)d_pos; (void)d_meta; (void)d_sym; (void)d_cat; (void)n; (void)d_logp; (void)st;
  MURAL_FAIL("MURAL_MODE_BF16 is not available in this build");
}
}  // namespace mural

extern "C" int mural_snv_tc_available(const mural_snv_model_t* m) { (void)m; return 0; }
