// MEX gateway: shadows gplite/gplite_nlZ.m for the outputs VBMC uses.
//   [nlZ,dnlZ] = gplite_nlZ(hyp,gp,hprior)                                    (gplite/gplite_nlZ.m:27-66)
// hyp Nhyp x 1: value (and gradient when nargout > 1).  hyp Nhyp x Ns with nargout == 1: one batched Gram +
// Cholesky for all columns (what gplite_train.m:200-204,318-330 evaluate one call at a time); with nargout > 1 the
// library raises gplite_nlZ:NoSampling like the reference (:40-43).  nargout > 2 (post, K_mat, Q) is not shadowed:
// keep a copy of the .m under another name for those diagnostic calls.
// Build: mex -R2018a mex/gplite_nlZ_mex.cpp -Iinclude -Lvbmc_b200/lib -lvbmc_b200 -output gplite/gplite_nlZ
#include "vbmc_b200_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  using namespace vbmex;
  if (nrhs < 2) mexErrMsgIdAndTxt("gplite_nlZ:nargin", "hyp and gp are required.");
  if (nlhs > 2) mexErrMsgIdAndTxt("vbmc_b200:OutOfScope", "gplite_nlZ with more than two outputs is not shadowed.");
  vbmc_b200_ctx* c = context();
  vbmc_b200_gp_desc g;
  gp_model(prhs[1], &g);
  g.Nhyp = (int)mxGetM(prhs[0]);
  g.S = (int)mxGetN(prhs[0]);
  g.hyp = mxGetDoubles(prhs[0]);
  vbmc_b200_hprior hp, *php = nullptr;
  if (given(nrhs, prhs, 2)) {
    memset(&hp, 0, sizeof(hp));
    hp.mu = dbl(prhs[2], 0, "mu"); hp.sigma = dbl(prhs[2], 0, "sigma"); hp.df = dbl(prhs[2], 0, "df");
    php = &hp;
  }
  if (g.S > 1 && nlhs < 2) {
    plhs[0] = mxCreateDoubleMatrix(1, g.S, mxREAL);
    check(vbmc_b200_gp_nlz_batch(c, &g, php, mxGetDoubles(plhs[0])));
    return;
  }
  if (g.S > 1)
    mexErrMsgIdAndTxt("gplite_nlZ:NoSampling",
                      "Computation of the log marginal likelihood is available only for one-sample hyperparameter inputs.");
  plhs[0] = mxCreateDoubleScalar(0);
  double* d = nullptr;
  if (nlhs > 1) { plhs[1] = mxCreateDoubleMatrix(g.Nhyp, 1, mxREAL); d = mxGetDoubles(plhs[1]); }
  check(vbmc_b200_gp_nlz(c, &g, php, mxGetDoubles(plhs[0]), d));
}
