// MEX helper behind gplite/gplite_post.m: the posterior of ALL hyper-parameter samples in one batched call.
//
//   post = gplite_post_core_mex(gp)                 replaces the per-sample loop of gplite_post.m:165-170
//                                                   (for s = 1:Ns, [~,~,gp.post(s)] = gplite_core(gp.post(s).hyp,gp,0,0); end)
//   post = gplite_post_core_mex(gp, xstar, ystar)   replaces the per-sample loop of the rank-one branch, gplite_post.m:186-246
//                                                   (Cholesky posteriors without s2; gp.X / gp.y are appended by the .m as before)
//
// gplite_post builds a struct with a dozen bookkeeping fields through gplite_covfun/meanfun/noisefun('info', ...)
// (gplite_post.m:94-151); that part stays in MATLAB — only the two loops that do the linear algebra are replaced, a
// two-line change at the call sites quoted above.  `post` is the 1 x Ns struct array with the fields
// gplite_core fills (gplite_core.m:278-285): hyp, alpha, sW, L, sn2_mult, Lchol.  The posterior also stays resident on
// the device, tagged with the address of post(1).alpha, so that negelcbo_vbmc / gplite_pred find it there.
// Build: mex -R2018a mex/gplite_post_core_mex.cpp -Iinclude -Lvbmc_b200/lib -lvbmc_b200 -output gplite/gplite_post_core_mex
#include "vbmc_b200_mex_common.h"

static const char* kFields[] = {"hyp", "alpha", "sW", "L", "sn2_mult", "Lchol"};

static mxArray* column(const double* src, size_t n) {
  mxArray* a = mxCreateDoubleMatrix(n, 1, mxREAL);
  memcpy(mxGetDoubles(a), src, sizeof(double) * n);
  return a;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  using namespace vbmex;
  (void)nlhs;
  if (nrhs < 1 || !mxIsStruct(prhs[0])) mexErrMsgIdAndTxt("gplite_post:NoGP", "gplite_post_core_mex needs the gp struct.");
  vbmc_b200_ctx* c = context();
  const mxArray* gp = prhs[0];
  const mxArray* post_in = fld(gp, 0, "post");
  if (!post_in) mexErrMsgIdAndTxt("vbmc_b200:gp", "gp.post is missing or empty.");
  const int S = (int)mxGetNumberOfElements(post_in);
  vbmc_b200_gp_desc g;
  gp_model(gp, &g);
  g.S = S;
  g.Nhyp = (int)mxGetNumberOfElements(mxGetField(post_in, 0, "hyp"));
  const size_t N = (size_t)g.N;

  if (nrhs < 2) {
    // ---- full refit: S x gplite_core(hyp,gp,0,0) as one batched Gram + Cholesky ----
    std::vector<double> hyp((size_t)g.Nhyp * S), alpha(N * S), L(N * N * S), sW1(S), mult(S);
    std::vector<int> Lchol(S);
    for (int s = 0; s < S; ++s) memcpy(&hyp[(size_t)s * g.Nhyp], dbl(post_in, s, "hyp"), sizeof(double) * g.Nhyp);
    g.hyp = hyp.data();
    check(vbmc_b200_gp_post(c, &g, alpha.data(), L.data(), sW1.data(), mult.data(), Lchol.data()));
    plhs[0] = mxCreateStructMatrix(1, S, 6, kFields);
    for (int s = 0; s < S; ++s) {
      mxSetFieldByNumber(plhs[0], s, 0, column(&hyp[(size_t)s * g.Nhyp], g.Nhyp));
      mxSetFieldByNumber(plhs[0], s, 1, column(&alpha[s * N], N));
      mxArray* sW = mxCreateDoubleMatrix(N, 1, mxREAL);
      for (size_t i = 0; i < N; ++i) mxGetDoubles(sW)[i] = sW1[s];      // ones(N,1)/sqrt(min(sn2)*sn2_mult)  (gplite_core.m:281)
      mxSetFieldByNumber(plhs[0], s, 2, sW);
      mxArray* Ls = mxCreateDoubleMatrix(N, N, mxREAL);
      memcpy(mxGetDoubles(Ls), &L[s * N * N], sizeof(double) * N * N);
      mxSetFieldByNumber(plhs[0], s, 3, Ls);
      mxSetFieldByNumber(plhs[0], s, 4, mxCreateDoubleScalar(mult[s]));
      mxSetFieldByNumber(plhs[0], s, 5, mxCreateLogicalScalar(Lchol[s] != 0));
    }
  } else {
    // ---- rank-one update with the observation (xstar, ystar) ----
    if (nrhs < 3) mexErrMsgIdAndTxt("gplite_post:NotRankOne", "xstar and ystar are required for the rank-one update.");
    if (mxGetM(prhs[1]) > 1)
      mexErrMsgIdAndTxt("gplite_post:NotRankOne", "GPLITE_POST with this input format only supports rank-one updates.");
    if ((int)mxGetNumberOfElements(prhs[1]) != g.D) mexErrMsgIdAndTxt("gplite_post:dimmismatch", "xstar must have D entries.");
    gp_attach(c, gp, /*want_L=*/true);
    std::vector<double> alpha((N + 1) * S), col((N + 1) * S), sWn(S);
    check(vbmc_b200_gp_post_update1(c, mxGetDoubles(prhs[1]), mxGetScalar(prhs[2]), alpha.data(), col.data(), sWn.data()));
    plhs[0] = mxCreateStructMatrix(1, S, 6, kFields);
    for (int s = 0; s < S; ++s) {
      mxSetFieldByNumber(plhs[0], s, 0, column(dbl(post_in, s, "hyp"), g.Nhyp));
      mxSetFieldByNumber(plhs[0], s, 1, column(&alpha[s * (N + 1)], N + 1));
      mxArray* sW = mxCreateDoubleMatrix(N + 1, 1, mxREAL);
      memcpy(mxGetDoubles(sW), dbl(post_in, s, "sW"), sizeof(double) * N);
      mxGetDoubles(sW)[N] = sWn[s];                                     // [sW; 1/sqrt(sn2_eff)]   (gplite_post.m:238)
      mxSetFieldByNumber(plhs[0], s, 2, sW);
      mxArray* Ls = mxCreateDoubleMatrix(N + 1, N + 1, mxREAL);
      double* Ld = mxGetDoubles(Ls);
      const mxArray* lc = mxGetField(post_in, s, "Lchol");
      const bool is_chol = !lc || mxGetScalar(lc) != 0;   // mxGetScalar reads logicals as 0/1
      if (is_chol) {                                                    // [L, c; 0, d]            (gplite_post.m:228-230)
        const double* Lo = dbl(post_in, s, "L");
        for (size_t j = 0; j < N; ++j) memcpy(Ld + j * (N + 1), Lo + j * N, sizeof(double) * N);   // last row stays 0
        memcpy(Ld + N * (N + 1), &col[s * (N + 1)], sizeof(double) * (N + 1));
      } else {                                                          // [L + v*a', -v; -v', -1/vstar]  (:234-238)
        check(vbmc_b200_gp_get_factor(c, s, Ld));
      }
      mxSetFieldByNumber(plhs[0], s, 3, Ls);
      const double* m = dbl(post_in, s, "sn2_mult");
      mxSetFieldByNumber(plhs[0], s, 4, mxCreateDoubleScalar(m ? m[0] : 1.0));
      mxSetFieldByNumber(plhs[0], s, 5, mxCreateLogicalScalar(is_chol));
    }
  }
  // the device holds exactly this posterior, factors included: tag it with the address MATLAB will hand back
  const unsigned long long key = (unsigned long long)(size_t)mxGetData(mxGetFieldByNumber(plhs[0], 0, 1)) & ~1ULL;
  check(vbmc_b200_gp_tag_set(c, key | 1ULL));
}
