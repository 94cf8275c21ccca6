// MEX gateway: shadows misc/negelcbo_vbmc.m (a negelcbo_vbmc.mexa64 in the same folder wins over the .m).
//   [F,dF,G,H,varF,dH,varGss,varG,varH,I_sk,J_sjk] =
//       negelcbo_vbmc(theta,beta,vp,gp,Ns,compute_grad,compute_var,altent_flag,thetabnd,entropy_alpha)
// Build (on a box that has MATLAB; not buildable in the CI image, which has no mex.h):
//   mex -R2018a mex/negelcbo_vbmc_mex.cpp -Iinclude -Lvbmc_b200/lib -lvbmc_b200 -output misc/negelcbo_vbmc
// The gateway only marshals MATLAB arrays (column-major doubles) into the C ABI of include/vbmc_b200.h;
// nargin/nargout defaults follow negelcbo_vbmc.m:9-17.
#include <math.h>
#include <string.h>

#include <vector>

#include "mex.h"
#include "vbmc_b200.h"

static vbmc_b200_ctx* g_ctx = nullptr;
static const void* g_gp_key = nullptr;  // data pointer of gp.post(1).alpha: cheap fingerprint of the resident GP
static unsigned long long g_call = 0;

static void cleanup() {
  if (g_ctx) vbmc_b200_destroy(g_ctx);
  g_ctx = nullptr;
}
static void check(int rc) {
  if (rc == VBMC_B200_OK) return;
  const char* msg = vbmc_b200_last_error();  // "<matlab:id>: text"
  const char* sep = strstr(msg, ": ");
  char id[128] = "vbmc_b200:error";
  if (sep && sep - msg < 120) { memcpy(id, msg, sep - msg); id[sep - msg] = 0; }
  mexErrMsgIdAndTxt(id, "%s", msg);
}
static double* field(const mxArray* s, mwIndex i, const char* name) {
  const mxArray* f = mxGetField(s, i, name);
  return (f && !mxIsEmpty(f)) ? mxGetDoubles(f) : nullptr;
}
static double scalar(const mxArray* s, const char* name, double dflt) {
  const mxArray* f = mxGetField(s, 0, name);
  return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 4) mexErrMsgIdAndTxt("negelcbo_vbmc:nargin", "theta, beta, vp, gp are required.");
  if (!g_ctx) { check(vbmc_b200_create(&g_ctx, 0)); mexAtExit(cleanup); mexLock(); }
  const mxArray *theta = prhs[0], *vp = prhs[2], *gp = prhs[3];
  double beta = mxIsEmpty(prhs[1]) ? 0.0 : mxGetScalar(prhs[1]);
  int Ns = (nrhs > 4 && !mxIsEmpty(prhs[4])) ? (int)mxGetScalar(prhs[4]) : 0;
  int compute_grad = (nrhs > 5 && !mxIsEmpty(prhs[5])) ? (mxGetScalar(prhs[5]) != 0) : (nlhs > 1);
  if (!mxIsFinite(beta)) beta = 0.0;
  int compute_var = (nrhs > 6 && !mxIsEmpty(prhs[6])) ? (int)mxGetScalar(prhs[6]) : (beta != 0.0 || nlhs > 4);
  const mxArray* tb = (nrhs > 8 && !mxIsEmpty(prhs[8])) ? prhs[8] : nullptr;

  // ---- VP (misc/setupvars_vbmc.m:78-99) ----
  vbmc_b200_vp_desc v;
  memset(&v, 0, sizeof(v));
  v.D = (int)scalar(vp, "D", 0); v.K = (int)scalar(vp, "K", 0);
  v.mu = field(vp, 0, "mu"); v.sigma = field(vp, 0, "sigma"); v.lambda = field(vp, 0, "lambda");
  v.w = field(vp, 0, "w"); v.eta = field(vp, 0, "eta"); v.delta = field(vp, 0, "delta");
  v.optimize_mu = scalar(vp, "optimize_mu", 1) != 0; v.optimize_sigma = scalar(vp, "optimize_sigma", 1) != 0;
  v.optimize_lambda = scalar(vp, "optimize_lambda", 1) != 0; v.optimize_weights = scalar(vp, "optimize_weights", 0) != 0;
  std::vector<double> dl;
  if (v.delta && mxGetNumberOfElements(mxGetField(vp, 0, "delta")) == 1) { dl.assign(v.D, v.delta[0]); v.delta = dl.data(); }
  check(vbmc_b200_vp_set(g_ctx, &v));

  // ---- GP (gplite_post.m:94-151): upload only when the posterior changed ----
  const mxArray* post = mxGetField(gp, 0, "post");
  const int S = (int)mxGetNumberOfElements(post);
  const void* key = mxGetData(mxGetField(post, 0, "alpha"));
  if (key != g_gp_key || compute_var) {
    const mxArray* X = mxGetField(gp, 0, "X");
    vbmc_b200_gp_desc g;
    memset(&g, 0, sizeof(g));
    g.N = (int)mxGetM(X); g.D = (int)mxGetN(X); g.S = S;
    g.Nhyp = (int)mxGetNumberOfElements(mxGetField(post, 0, "hyp"));
    g.covfun = (int)mxGetDoubles(mxGetField(gp, 0, "covfun"))[0];
    g.meanfun = (int)scalar(gp, "meanfun", 1);
    const double* nf = field(gp, 0, "noisefun");
    for (int i = 0; i < 3; ++i) g.noisefun[i] = (int)nf[i];
    g.X = mxGetDoubles(X);
    std::vector<double> hyp((size_t)g.Nhyp * S), alpha((size_t)g.N * S), sW1(S), L;
    std::vector<int> Lchol(S);
    if (compute_var) L.resize((size_t)g.N * g.N * S);
    for (int s = 0; s < S; ++s) {
      memcpy(&hyp[(size_t)s * g.Nhyp], field(post, s, "hyp"), sizeof(double) * g.Nhyp);
      memcpy(&alpha[(size_t)s * g.N], field(post, s, "alpha"), sizeof(double) * g.N);
      sW1[s] = field(post, s, "sW")[0];
      Lchol[s] = mxGetScalar(mxGetField(post, s, "Lchol")) != 0;
      if (compute_var) memcpy(&L[(size_t)s * g.N * g.N], field(post, s, "L"), sizeof(double) * g.N * g.N);
    }
    g.hyp = hyp.data();
    check(vbmc_b200_gp_attach(g_ctx, &g, alpha.data(), sW1.data(), Lchol.data(), compute_var ? L.data() : nullptr));
    g_gp_key = key;
  }
  // ---- thetabnd (misc/vpbounds.m:32-52) ----
  if (tb) {
    const mxArray* lb = mxGetField(tb, 0, "lb");
    check(vbmc_b200_thetabnd_set(g_ctx, (int)mxGetNumberOfElements(lb), mxGetDoubles(lb), field(tb, 0, "ub"),
                                 scalar(tb, "TolCon", 0), scalar(tb, "WeightThreshold", 0), scalar(tb, "WeightPenalty", 0)));
  } else {
    check(vbmc_b200_thetabnd_set(g_ctx, 0, nullptr, nullptr, 0, 0, 0));
  }
  // ---- the call ----
  const int ntheta = (int)mxGetNumberOfElements(theta);
  vbmc_b200_negelcbo_args a;
  memset(&a, 0, sizeof(a));
  a.theta = mxGetDoubles(theta); a.ntheta = ntheta; a.beta = beta; a.Ns = Ns;
  a.compute_grad = compute_grad; a.compute_var = compute_var; a.separate_K = nlhs > 9; a.use_thetabnd = tb != nullptr;
  a.eps_mode = VBMC_B200_EPS_PHILOX;      // replaces MATLAB's global randn stream (entmc_vbmc.m:53)
  a.seed = 0x5eed; a.stream = g_call++;
  double sc[7] = {0};
  a.F = &sc[0]; a.G = &sc[1]; a.H = &sc[2]; a.varF = &sc[3]; a.varGss = &sc[4]; a.varG = &sc[5]; a.varH = &sc[6];
  plhs[0] = mxCreateDoubleScalar(0);
  if (nlhs > 1) { plhs[1] = mxCreateDoubleMatrix(compute_grad ? ntheta : 0, compute_grad ? 1 : 0, mxREAL); a.dF = compute_grad ? mxGetDoubles(plhs[1]) : nullptr; }
  if (nlhs > 5) { plhs[5] = mxCreateDoubleMatrix(compute_grad ? ntheta : 0, compute_grad ? 1 : 0, mxREAL); a.dH = compute_grad ? mxGetDoubles(plhs[5]) : nullptr; }
  if (nlhs > 9) { plhs[9] = mxCreateDoubleMatrix(S, v.K, mxREAL); a.I_sk = mxGetDoubles(plhs[9]); }
  if (nlhs > 10) {
    mwSize dims[3] = {(mwSize)S, (mwSize)v.K, (mwSize)v.K};
    plhs[10] = compute_var ? mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL) : mxCreateDoubleMatrix(0, 0, mxREAL);
    a.J_sjk = compute_var ? mxGetDoubles(plhs[10]) : nullptr;
  }
  check(vbmc_b200_negelcbo(g_ctx, &a));
  *mxGetDoubles(plhs[0]) = sc[0];
  const int sidx[] = {-1, -1, 1, 2, 3, -1, 4, 5, 6};  // G,H,varF,(dH),varGss,varG,varH
  for (int o = 2; o < 9 && o < nlhs; ++o)
    if (sidx[o] >= 0) plhs[o] = mxCreateDoubleScalar(sc[sidx[o]]);
}
