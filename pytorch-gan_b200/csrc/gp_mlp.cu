// gp_mlp.cu -- WGAN-GP critic gradient penalty (placeholder until the fused kernel lands).
#include "common.cuh"
using namespace b200gan;
extern "C" size_t b200gan_gp_mlp_workspace_floats(const b200gan_gp_mlp_desc *) { return 0; }
extern "C" int b200gan_gp_mlp_fwd_bwd(const b200gan_gp_mlp_desc *, const float *, const float *, const float *,
                                      const float *, const float *, const float *, float *, float *, float *,
                                      float *, float *, void *) {
  B2_UNSUPPORTED("gp_mlp_fwd_bwd: not built in this revision");
}
