// MEX entry point for the whole stochastic optimisation of misc/vpoptimize_vbmc.m:126-127 in ONE call:
//   [x,f,xtab,ftab,iter] = fminadam_negelcbo_mex(theta0,beta,vp,gp,Ns,compute_var,thetabnd,LB,UB,TolFun,MaxIter,master_stepsize)
// == fminadam(@(t) negelcbo_vbmc(t,beta,vp,gp,Ns,1,compute_var,altent,thetabnd,entropy_alpha), theta0, LB, UB, TolFun,
//             MaxIter, master_stepsize)                                                         (utils/fminadam.m:1-102)
// fminadam receives a function handle, which cannot be evaluated on the device, so the integration point is the call
// site.  beta ~= 0 (compute_var == 2) runs the variance-penalised objective of negelcbo_vbmc.m:119-130
// every iteration; the factors gp.post(s).L are attached for it.
// Build: mex -R2018a mex/fminadam_negelcbo_mex.cpp -Iinclude -Lvbmc_b200/lib -lvbmc_b200 -output utils/fminadam_negelcbo_mex
#include "vbmc_b200_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  using namespace vbmex;
  if (nrhs < 5) mexErrMsgIdAndTxt("fminadam:nargin", "theta0, beta, vp, gp and Ns are required.");
  vbmc_b200_ctx* c = context();
  VpHold vh;
  vp_set(c, prhs[2], &vh);
  const int cvar = given(nrhs, prhs, 5) ? (int)mxGetScalar(prhs[5]) : 0;
  gp_attach(c, prhs[3], /*want_L=*/cvar != 0);
  const mxArray* tb = given(nrhs, prhs, 6) ? prhs[6] : nullptr;
  thetabnd_set(c, tb);
  const int n = (int)mxGetNumberOfElements(prhs[0]);
  vbmc_b200_fminadam_args a;
  memset(&a, 0, sizeof(a));
  a.x0 = mxGetDoubles(prhs[0]);
  a.nvars = n;
  a.beta = mxIsEmpty(prhs[1]) ? 0.0 : mxGetScalar(prhs[1]);
  a.Ns = (int)mxGetScalar(prhs[4]);
  a.compute_var = cvar;
  a.use_thetabnd = tb != nullptr;
  if (given(nrhs, prhs, 7)) {
    if ((int)mxGetNumberOfElements(prhs[7]) != n) mexErrMsgIdAndTxt("fminadam:bounds", "LB must have numel(x0) entries.");
    a.LB = mxGetDoubles(prhs[7]);
  }
  if (given(nrhs, prhs, 8)) {
    if ((int)mxGetNumberOfElements(prhs[8]) != n) mexErrMsgIdAndTxt("fminadam:bounds", "UB must have numel(x0) entries.");
    a.UB = mxGetDoubles(prhs[8]);
  }
  a.TolFun = given(nrhs, prhs, 9) ? mxGetScalar(prhs[9]) : 0.0;          // <= 0 -> 0.001 (fminadam.m:6)
  a.MaxIter = given(nrhs, prhs, 10) ? (int)mxGetScalar(prhs[10]) : 0;    // <= 0 -> 10000 (fminadam.m:7)
  if (given(nrhs, prhs, 11)) {                                           // master_stepsize struct (fminadam.m:11-18)
    a.stepsize_max = num(prhs[11], "max", 0.0);
    a.stepsize_min = num(prhs[11], "min", 0.0);
    a.stepsize_decay = num(prhs[11], "decay", 0.0);
  }
  a.eps_mode = VBMC_B200_EPS_PHILOX;
  a.seed = kSeed + 2;
  static unsigned long long stream = 0;                                   // iteration i of this call draws from stream + i
  a.stream = stream;
  const int maxit = a.MaxIter > 0 ? a.MaxIter : 10000;
  std::vector<double> xtab((size_t)n * maxit), ftab(maxit);
  int iter = 0;
  double f = 0.0;
  const bool row = mxGetM(prhs[0]) == 1 && n > 1;                         // x = reshape(x,size(x0)) (fminadam.m:101)
  plhs[0] = mxCreateDoubleMatrix(row ? 1 : n, row ? n : 1, mxREAL);
  a.x = mxGetDoubles(plhs[0]);
  a.f = &f;
  a.xtab = xtab.data();
  a.ftab = ftab.data();
  a.iter = &iter;
  check(vbmc_b200_fminadam(c, &a));
  stream += (unsigned long long)iter;
  if (nlhs > 1) plhs[1] = mxCreateDoubleScalar(f);
  if (nlhs > 2) {                              // xtab(:,1:iter) (fminadam.m:98), transposed for a row x0 (fminadam.m:103)
    plhs[2] = mxCreateDoubleMatrix(row ? iter : n, row ? n : iter, mxREAL);
    double* o = mxGetDoubles(plhs[2]);
    if (!row) {
      memcpy(o, xtab.data(), sizeof(double) * (size_t)n * iter);
    } else {
      for (int it = 0; it < iter; ++it)
        for (int i = 0; i < n; ++i) o[(size_t)i * iter + it] = xtab[(size_t)it * n + i];
    }
  }
  if (nlhs > 3) {
    plhs[3] = mxCreateDoubleMatrix(1, iter, mxREAL);
    memcpy(mxGetDoubles(plhs[3]), ftab.data(), sizeof(double) * iter);
  }
  if (nlhs > 4) plhs[4] = mxCreateDoubleScalar((double)iter);
}
