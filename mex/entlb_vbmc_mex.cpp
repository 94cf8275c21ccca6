// MEX gateway: shadows ent/entlb_vbmc.m.
//   [H,dH] = entlb_vbmc(vp,grad_flags,jacobian_flag)                          (ent/entlb_vbmc.m:1-147)
// Defaults: grad_flags = nargout > 1 (scalar expands to all four blocks, :8-13), jacobian_flag = true (:16).
// The third output of the .m (gammasum) is a debugging aid no caller in VBMC uses; it is not shadowed.
// Build: mex -R2018a mex/entlb_vbmc_mex.cpp -Iinclude -Imex -Lvbmc_b200/lib -lvbmc_b200 -output ent/entlb_vbmc
#include "vbmc_b200_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  using namespace vbmex;
  if (nrhs < 1) mexErrMsgIdAndTxt("entlb_vbmc:nargin", "vp is required.");
  if (nlhs > 2) mexErrMsgIdAndTxt("vbmc_b200:OutOfScope", "entlb_vbmc's third output (gammasum) is not shadowed.");
  vbmc_b200_ctx* c = context();
  VpHold vh;
  vp_set(c, prhs[0], &vh);
  int gf[4] = {nlhs > 1, nlhs > 1, nlhs > 1, nlhs > 1};
  if (given(nrhs, prhs, 1)) {
    const size_t n = mxGetNumberOfElements(prhs[1]);
    const double* g = mxGetDoubles(prhs[1]);
    for (int i = 0; i < 4; ++i) gf[i] = (n == 1 ? g[0] : (i < (int)n ? g[i] : 0.0)) != 0.0;
  }
  if (nlhs < 2) gf[0] = gf[1] = gf[2] = gf[3] = 0;
  const int jac = given(nrhs, prhs, 2) ? (mxGetScalar(prhs[2]) != 0) : 1;
  const int D = vh.d.D, K = vh.d.K;
  const int n = D * K * gf[0] + K * gf[1] + D * gf[2] + K * gf[3];
  double H = 0.0;
  plhs[0] = mxCreateDoubleScalar(0);
  double* dH = nullptr;
  if (nlhs > 1) {
    plhs[1] = mxCreateDoubleMatrix(n, n ? 1 : 0, mxREAL);
    dH = n ? mxGetDoubles(plhs[1]) : nullptr;
  }
  check(vbmc_b200_entlb(c, gf, jac, &H, dH));
  *mxGetDoubles(plhs[0]) = H;
}
