// MEX gateway: shadows misc/gplogjoint.m.
//   [F,dF,varF,dvarF,varss,I_sk,J_sjk] = gplogjoint(vp,gp,grad_flags,avg_flag,jacobian_flag,compute_var,separate_K)
// nargin/nargout defaults of misc/gplogjoint.m:14-22: grad_flags = nargout > 1 (scalar expands), avg_flag = true,
// jacobian_flag = true, compute_var = nargout > 2, separate_K = nargout > 5; gradients off when nargout < 2.
// Build: mex -R2018a mex/gplogjoint_mex.cpp -Iinclude -Lvbmc_b200/lib -lvbmc_b200 -output misc/gplogjoint
#include "vbmc_b200_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  using namespace vbmex;
  if (nrhs < 2) mexErrMsgIdAndTxt("gplogjoint:nargin", "vp and gp are required.");
  vbmc_b200_ctx* c = context();
  VpHold vh;
  vp_set(c, prhs[0], &vh);
  int gf[4] = {nlhs > 1, nlhs > 1, nlhs > 1, nlhs > 1};
  if (given(nrhs, prhs, 2)) {
    const size_t n = mxGetNumberOfElements(prhs[2]);
    const double* g = mxGetDoubles(prhs[2]);
    for (int i = 0; i < 4; ++i) gf[i] = (n == 1 ? g[0] : (i < (int)n ? g[i] : 0.0)) != 0.0;
  }
  if (nlhs < 2) gf[0] = gf[1] = gf[2] = gf[3] = 0;
  const int avg = given(nrhs, prhs, 3) ? (mxGetScalar(prhs[3]) != 0) : 1;
  const int jac = given(nrhs, prhs, 4) ? (mxGetScalar(prhs[4]) != 0) : 1;
  const int cvar = given(nrhs, prhs, 5) ? (int)mxGetScalar(prhs[5]) : (nlhs > 2);
  const int sepK = given(nrhs, prhs, 6) ? (mxGetScalar(prhs[6]) != 0) : (nlhs > 5);
  const int S = gp_attach(c, prhs[1], cvar != 0);
  const int D = vh.d.D, K = vh.d.K;
  const int n = D * K * gf[0] + K * gf[1] + D * gf[2] + K * gf[3];
  double F = 0, varF = 0, varss = 0;
  double *dF = nullptr, *dvar = nullptr, *Isk = nullptr, *Jsjk = nullptr;
  plhs[0] = mxCreateDoubleScalar(0);
  if (nlhs > 1) { plhs[1] = mxCreateDoubleMatrix(n, n ? 1 : 0, mxREAL); dF = n ? mxGetDoubles(plhs[1]) : nullptr; }
  if (nlhs > 3) {
    const bool want = n && cvar;
    plhs[3] = mxCreateDoubleMatrix(want ? n : 0, want ? 1 : 0, mxREAL);
    dvar = want ? mxGetDoubles(plhs[3]) : nullptr;
  }
  if (nlhs > 5) { plhs[5] = mxCreateDoubleMatrix(sepK ? S : 0, sepK ? K : 0, mxREAL); Isk = sepK ? mxGetDoubles(plhs[5]) : nullptr; }
  if (nlhs > 6) {
    const bool want = sepK && cvar;
    mwSize dims[3] = {(mwSize)(want ? S : 0), (mwSize)(want ? K : 0), (mwSize)(want ? K : 0)};
    plhs[6] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
    Jsjk = want ? mxGetDoubles(plhs[6]) : nullptr;
  }
  check(vbmc_b200_gplogjoint(c, gf, avg, jac, cvar, &F, dF, &varF, dvar, &varss, Isk, Jsjk));
  *mxGetDoubles(plhs[0]) = F;
  if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(varF);
  if (nlhs > 4) plhs[4] = mxCreateDoubleScalar(varss);
}
