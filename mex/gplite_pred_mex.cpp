// MEX gateway: shadows gplite/gplite_pred.m.
//   [ymu,ys2,fmu,fs2,lp] = gplite_pred(gp,Xstar,ystar,s2star,ssflag,nowarpflag)   (gplite/gplite_pred.m:1-163)
// The posterior is attached (with its factors) unless it is the one already resident — e.g. left there by the
// gplite_post gateway.  Output warping / integrated mean functions are outside the build (not VBMC defaults).
// Build: mex -R2018a mex/gplite_pred_mex.cpp -Iinclude -Lvbmc_b200/lib -lvbmc_b200 -output gplite/gplite_pred
#include "vbmc_b200_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  using namespace vbmex;
  if (nrhs < 2) mexErrMsgIdAndTxt("gplite_pred:nargin", "gp and Xstar are required.");
  vbmc_b200_ctx* c = context();
  const mxArray* gp = prhs[0];
  if (fld(gp, 0, "outwarpfun") || num(gp, "intmeanfun", 0) > 0)
    mexErrMsgIdAndTxt("vbmc_b200:OutOfScope", "gplite_pred with output warping or an integrated mean function is not shadowed.");
  const int S = gp_attach(c, gp, nlhs > 1);
  const int Nstar = (int)mxGetM(prhs[1]);
  const double* ystar = given(nrhs, prhs, 2) ? mxGetDoubles(prhs[2]) : nullptr;
  const double* s2star = given(nrhs, prhs, 3) ? mxGetDoubles(prhs[3]) : nullptr;
  if (ystar && (int)mxGetM(prhs[2]) != Nstar)
    mexErrMsgIdAndTxt("gplite_pred:ydimmismatch", "YSTAR should be empty or a column vector of NSTAR observations.");
  if (s2star && (int)mxGetM(prhs[3]) != Nstar)
    mexErrMsgIdAndTxt("gplite_pred:s2dimmismatch", "S2STAR should be empty or a column vector of NSTAR estimated variances.");
  const int ssflag = given(nrhs, prhs, 4) ? (mxGetScalar(prhs[4]) != 0) : 0;
  const int cols = (ssflag || S == 1) ? S : 1;
  double* out[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int o = 0; o < 4 && o < (nlhs > 0 ? nlhs : 1); ++o) {
    plhs[o] = mxCreateDoubleMatrix(Nstar, cols, mxREAL);
    out[o] = mxGetDoubles(plhs[o]);
  }
  if (nlhs > 4) {
    const bool want = ystar != nullptr;
    plhs[4] = mxCreateDoubleMatrix(want ? Nstar : 0, want ? S : 0, mxREAL);
    out[4] = want ? mxGetDoubles(plhs[4]) : nullptr;
  }
  check(vbmc_b200_gp_pred(c, Nstar, mxGetDoubles(prhs[1]), ystar, s2star, ssflag, nlhs > 1, out[0], out[1], out[2], out[3], out[4]));
}
