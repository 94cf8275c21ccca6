// GP-surrogate refit (gplite_post / gplite_nlZ -> gplite_core).  Placeholder until the Gram +
// Cholesky kernels land.
#include "common.cuh"
extern "C" {
int vbmc_b200_gp_post(vbmc_b200_ctx*, const vbmc_b200_gp_desc*, double*, double*, double*, double*, int*) {
  VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:NotYet: gp_post is not built yet");
}
int vbmc_b200_gp_nlz(vbmc_b200_ctx*, const vbmc_b200_gp_desc*, const vbmc_b200_hprior*, double*, double*) {
  VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:NotYet: gp_nlz is not built yet");
}
}
