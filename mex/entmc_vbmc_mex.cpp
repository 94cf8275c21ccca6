// MEX gateway: shadows ent/entmc_vbmc.m.
//   [H,dH] = entmc_vbmc(vp,Ns,grad_flags,jacobian_flag)                       (ent/entmc_vbmc.m:1-14)
// Defaults: grad_flags = nargout > 1 (scalar expands to all four blocks, :6-10), jacobian_flag = true (:11).
// Draws: the device generator, keyed by this gateway's call counter (replaces randn(D,1,Ns/2), :53).
// Build: mex -R2018a mex/entmc_vbmc_mex.cpp -Iinclude -Lvbmc_b200/lib -lvbmc_b200 -output ent/entmc_vbmc
#include "vbmc_b200_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  using namespace vbmex;
  if (nrhs < 2) mexErrMsgIdAndTxt("entmc_vbmc:nargin", "vp and Ns are required.");
  vbmc_b200_ctx* c = context();
  VpHold vh;
  vp_set(c, prhs[0], &vh);
  const int Ns = (int)mxGetScalar(prhs[1]);
  int gf[4] = {nlhs > 1, nlhs > 1, nlhs > 1, nlhs > 1};
  if (given(nrhs, prhs, 2)) {
    const size_t n = mxGetNumberOfElements(prhs[2]);
    const double* g = mxGetDoubles(prhs[2]);
    for (int i = 0; i < 4; ++i) gf[i] = (n == 1 ? g[0] : (i < (int)n ? g[i] : 0.0)) != 0.0;   // :8-10
  }
  if (nlhs < 2) gf[0] = gf[1] = gf[2] = gf[3] = 0;
  const int jac = given(nrhs, prhs, 3) ? (mxGetScalar(prhs[3]) != 0) : 1;
  const int D = vh.d.D, K = vh.d.K;
  const int n = D * K * gf[0] + K * gf[1] + D * gf[2] + K * gf[3];
  double H = 0.0;
  plhs[0] = mxCreateDoubleScalar(0);
  double* dH = nullptr;
  if (nlhs > 1) {
    plhs[1] = mxCreateDoubleMatrix(n, n ? 1 : 0, mxREAL);
    dH = n ? mxGetDoubles(plhs[1]) : nullptr;
  }
  check(vbmc_b200_entmc(c, Ns, gf, jac, VBMC_B200_EPS_PHILOX, nullptr, kSeed + 1, next_stream(), &H, dH));
  *mxGetDoubles(plhs[0]) = H;
}
