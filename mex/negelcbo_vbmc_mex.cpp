// MEX gateway: shadows misc/negelcbo_vbmc.m (a negelcbo_vbmc.mexa64 in the same folder wins over the .m).
//   [F,dF,G,H,varF,dH,varGss,varG,varH,I_sk,J_sjk] =
//       negelcbo_vbmc(theta,beta,vp,gp,Ns,compute_grad,compute_var,altent_flag,thetabnd,entropy_alpha)
// Build (on a box that has MATLAB; not buildable in the CI image, which has no mex.h):
//   mex -R2018a mex/negelcbo_vbmc_mex.cpp -Iinclude -Lvbmc_b200/lib -lvbmc_b200 -output misc/negelcbo_vbmc
// The gateway only marshals MATLAB arrays (column-major doubles) into the C ABI of include/vbmc_b200.h;
// nargin/nargout defaults follow negelcbo_vbmc.m:9-17; altent_flag and entropy_alpha are accepted and ignored (:19).
// Per Adam step this costs one H2D of theta and one D2H of [F; dF]: vp, gp and thetabnd are re-sent only when they
// changed (gp: fingerprint kept by the library, see vbmc_b200_mex_common.h).
#include "vbmc_b200_mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  using namespace vbmex;
  if (nrhs < 4) mexErrMsgIdAndTxt("negelcbo_vbmc:nargin", "theta, beta, vp, gp are required.");
  vbmc_b200_ctx* c = context();
  const mxArray* theta = prhs[0];
  double beta = mxIsEmpty(prhs[1]) ? 0.0 : mxGetScalar(prhs[1]);
  if (!mxIsFinite(beta)) beta = 0.0;                                                      // :15
  const int Ns = given(nrhs, prhs, 4) ? (int)mxGetScalar(prhs[4]) : 0;
  const int compute_grad = given(nrhs, prhs, 5) ? (mxGetScalar(prhs[5]) != 0) : (nlhs > 1);   // :10
  const int compute_var = given(nrhs, prhs, 6) ? (int)mxGetScalar(prhs[6]) : (beta != 0.0 || nlhs > 4);  // :16
  const mxArray* tb = given(nrhs, prhs, 8) ? prhs[8] : nullptr;

  VpHold vh;
  vp_set(c, prhs[2], &vh);
  const int S = gp_attach(c, prhs[3], compute_var != 0);
  thetabnd_set(c, tb);

  const int ntheta = (int)mxGetNumberOfElements(theta);
  vbmc_b200_negelcbo_args a;
  memset(&a, 0, sizeof(a));
  a.theta = mxGetDoubles(theta); a.ntheta = ntheta; a.beta = beta; a.Ns = Ns;
  a.compute_grad = compute_grad; a.compute_var = compute_var; a.separate_K = nlhs > 9; a.use_thetabnd = tb != nullptr;
  a.eps_mode = VBMC_B200_EPS_PHILOX;      // replaces MATLAB's global randn stream (entmc_vbmc.m:53)
  a.seed = kSeed; a.stream = next_stream();
  double sc[7] = {0};
  a.F = &sc[0]; a.G = &sc[1]; a.H = &sc[2]; a.varF = &sc[3]; a.varGss = &sc[4]; a.varG = &sc[5]; a.varH = &sc[6];
  plhs[0] = mxCreateDoubleScalar(0);
  if (nlhs > 1) { plhs[1] = mxCreateDoubleMatrix(compute_grad ? ntheta : 0, compute_grad ? 1 : 0, mxREAL); a.dF = compute_grad ? mxGetDoubles(plhs[1]) : nullptr; }
  if (nlhs > 5) { plhs[5] = mxCreateDoubleMatrix(compute_grad ? ntheta : 0, compute_grad ? 1 : 0, mxREAL); a.dH = compute_grad ? mxGetDoubles(plhs[5]) : nullptr; }
  if (nlhs > 9) { plhs[9] = mxCreateDoubleMatrix(S, vh.d.K, mxREAL); a.I_sk = mxGetDoubles(plhs[9]); }
  if (nlhs > 10) {
    mwSize dims[3] = {(mwSize)S, (mwSize)vh.d.K, (mwSize)vh.d.K};
    plhs[10] = compute_var ? mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL) : mxCreateDoubleMatrix(0, 0, mxREAL);
    a.J_sjk = compute_var ? mxGetDoubles(plhs[10]) : nullptr;
  }
  check(vbmc_b200_negelcbo(c, &a));
  *mxGetDoubles(plhs[0]) = sc[0];
  const int sidx[] = {-1, -1, 1, 2, 3, -1, 4, 5, 6};  // G,H,varF,(dH),varGss,varG,varH
  for (int o = 2; o < 9 && o < nlhs; ++o)
    if (sidx[o] >= 0) plhs[o] = mxCreateDoubleScalar(sc[sidx[o]]);
}
