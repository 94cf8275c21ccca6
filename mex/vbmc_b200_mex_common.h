// Shared pieces of the MEX gateways in this folder: the process-wide library context, MATLAB-style error
// forwarding, and the marshalling of the reference's VP / GP / thetabnd structs into the C ABI
// (include/vbmc_b200.h).  Not compiled in the CI image (no mex.h there); tests/test_mex_sources.py compiles every
// gateway against tests/stubs/mex.h, a declaration-only stand-in, so that the sources stay in step with the C ABI.
#pragma once
#include <math.h>
#include <string.h>

#include <vector>

#include "mex.h"
#include "vbmc_b200.h"

namespace vbmex {

// One context per process, owned by libvbmc_b200 (every gateway is its own .mexa64 but links the same library), so the
// posterior gplite_post leaves on the device is the one negelcbo_vbmc / gplite_pred evaluate.
inline void check(int rc) {
  if (rc == VBMC_B200_OK) return;
  const char* msg = vbmc_b200_last_error();  // "<matlab:id>: text"
  const char* sep = strstr(msg, ": ");
  char id[128] = "vbmc_b200:error";
  if (sep && sep - msg < 120) {
    memcpy(id, msg, sep - msg);
    id[sep - msg] = 0;
  }
  mexErrMsgIdAndTxt(id, "%s", msg);
}
inline void release_all() { vbmc_b200_shared_release(); }
inline vbmc_b200_ctx* context() {
  static bool registered = false;
  vbmc_b200_ctx* c = nullptr;
  check(vbmc_b200_shared(&c, 0));
  if (!registered) {
    mexAtExit(release_all);
    mexLock();
    registered = true;
  }
  return c;
}

inline const mxArray* fld(const mxArray* s, mwIndex i, const char* name) {
  const mxArray* f = s ? mxGetField(s, i, name) : nullptr;
  return (f && !mxIsEmpty(f)) ? f : nullptr;
}
inline double* dbl(const mxArray* s, mwIndex i, const char* name) {
  const mxArray* f = fld(s, i, name);
  return f ? mxGetDoubles(f) : nullptr;
}
inline double num(const mxArray* s, const char* name, double dflt) {
  const mxArray* f = fld(s, 0, name);
  return f ? mxGetScalar(f) : dflt;
}
inline bool given(int nrhs, const mxArray* prhs[], int i) { return nrhs > i && !mxIsEmpty(prhs[i]); }

// ---- VP struct (misc/setupvars_vbmc.m:78-99) ----
struct VpHold {
  vbmc_b200_vp_desc d;
  std::vector<double> delta;
};
inline void vp_set(vbmc_b200_ctx* c, const mxArray* vp, VpHold* h) {
  vbmc_b200_vp_desc& v = h->d;
  memset(&v, 0, sizeof(v));
  v.D = (int)num(vp, "D", 0);
  v.K = (int)num(vp, "K", 0);
  v.mu = dbl(vp, 0, "mu"); v.sigma = dbl(vp, 0, "sigma"); v.lambda = dbl(vp, 0, "lambda");
  v.w = dbl(vp, 0, "w"); v.eta = dbl(vp, 0, "eta"); v.delta = dbl(vp, 0, "delta");
  v.optimize_mu = num(vp, "optimize_mu", 1) != 0; v.optimize_sigma = num(vp, "optimize_sigma", 1) != 0;
  v.optimize_lambda = num(vp, "optimize_lambda", 1) != 0; v.optimize_weights = num(vp, "optimize_weights", 0) != 0;
  if (v.delta && mxGetNumberOfElements(fld(vp, 0, "delta")) == 1) {  // scalar delta applies to every dimension
    h->delta.assign(v.D, v.delta[0]);
    v.delta = h->delta.data();
  }
  check(vbmc_b200_vp_set(c, &v));
}

// ---- GP struct (gplite/gplite_post.m:94-151) ----
// descriptor of the model part (X, y, s2, covfun, meanfun, noisefun); hyp is filled by the caller
inline void gp_model(const mxArray* gp, vbmc_b200_gp_desc* g) {
  memset(g, 0, sizeof(*g));
  const mxArray* X = fld(gp, 0, "X");
  if (!X) mexErrMsgIdAndTxt("vbmc_b200:gp", "gp.X is missing or empty.");
  g->N = (int)mxGetM(X);
  g->D = (int)mxGetN(X);
  g->X = mxGetDoubles(X);
  g->y = dbl(gp, 0, "y");
  g->s2 = dbl(gp, 0, "s2");
  const double* cf = dbl(gp, 0, "covfun");
  g->covfun = cf ? (int)cf[0] : 1;
  g->meanfun = (int)num(gp, "meanfun", 1);
  const mxArray* nf = fld(gp, 0, "noisefun");
  g->noisefun[0] = 1; g->noisefun[1] = g->s2 ? 1 : 0; g->noisefun[2] = 0;       // gplite_post.m:103-105 defaults
  if (nf)
    for (int i = 0; i < 3 && i < (int)mxGetNumberOfElements(nf); ++i) g->noisefun[i] = (int)mxGetDoubles(nf)[i];
}

// Fingerprint of a gp struct (FNV-1a): shape, model switches, the data address AND first/last value of every gp.post(s).alpha, the
// address of gp.X, and the content of every hyper-parameter vector, sn2_mult and Lchol.  MATLAB arrays are copy-on-write, so any
// edit of alpha / X gives a new address; the small fields are hashed by value.  Low bit = "the factors L were uploaded too".
inline unsigned long long fnv1a(unsigned long long h, const void* p, size_t n) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; ++i) {
    h ^= b[i];
    h *= 1099511628211ULL;
  }
  return h;
}
inline unsigned long long gp_fingerprint(const mxArray* gp, const mxArray* post, int S) {
  unsigned long long h = 1469598103934665603ULL;
  vbmc_b200_gp_desc g;
  gp_model(gp, &g);
  const int dims[8] = {S, g.N, g.D, g.covfun, g.meanfun, g.noisefun[0], g.noisefun[1], g.noisefun[2]};
  h = fnv1a(h, dims, sizeof(dims));
  const void* px[3] = {g.X, g.y, g.s2};
  h = fnv1a(h, px, sizeof(px));
  for (int s = 0; s < S; ++s) {
    const mxArray* a = mxGetField(post, s, "alpha");
    const mxArray* hy = mxGetField(post, s, "hyp");
    if (!a || !hy) mexErrMsgIdAndTxt("vbmc_b200:gp", "gp.post(%d) has no alpha / hyp.", s + 1);
    const void* pa = mxGetData(a);
    const size_t na = mxGetNumberOfElements(a), nh = mxGetNumberOfElements(hy);
    h = fnv1a(h, &pa, sizeof(pa));
    h = fnv1a(h, &na, sizeof(na));
    if (na) {
      h = fnv1a(h, mxGetDoubles(a), sizeof(double));
      h = fnv1a(h, mxGetDoubles(a) + na - 1, sizeof(double));
    }
    h = fnv1a(h, mxGetDoubles(hy), sizeof(double) * nh);
    const double m = dbl(post, s, "sn2_mult") ? dbl(post, s, "sn2_mult")[0] : 1.0;
    const double lc = mxGetField(post, s, "Lchol") ? mxGetScalar(mxGetField(post, s, "Lchol")) : 1.0;
    h = fnv1a(h, &m, sizeof(m));
    h = fnv1a(h, &lc, sizeof(lc));
  }
  return h & ~1ULL;
}

// Make gp.post resident when it is not already.  The fingerprint is kept by the library next to the posterior itself
// (vbmc_b200_gp_tag_*) -- every gateway is a separate shared object, a static here would not be seen by the others -- and the
// library clears it whenever the resident posterior changes.  A hit is only trusted when the RESIDENT shape equals the struct's
// (the gateways size their outputs from the struct; the library computes with the resident N, D, S).
inline int gp_attach(vbmc_b200_ctx* c, const mxArray* gp, bool want_L) {
  const mxArray* post = fld(gp, 0, "post");
  if (!post) mexErrMsgIdAndTxt("vbmc_b200:gp", "gp.post is missing or empty.");
  const int S = (int)mxGetNumberOfElements(post);
  const unsigned long long key = gp_fingerprint(gp, post, S);
  unsigned long long have = 0;
  check(vbmc_b200_gp_tag_get(c, &have));
  if (have && (have & ~1ULL) == key && ((have & 1ULL) || !want_L)) {
    int rN = 0, rD = 0, rS = 0;
    check(vbmc_b200_gp_shape(c, &rN, &rD, &rS));
    const mxArray* X = fld(gp, 0, "X");
    if (X && rS == S && rN == (int)mxGetM(X) && rD == (int)mxGetN(X)) return S;
  }
  vbmc_b200_gp_desc g;
  gp_model(gp, &g);
  g.S = S;
  g.Nhyp = (int)mxGetNumberOfElements(mxGetField(post, 0, "hyp"));
  std::vector<double> hyp((size_t)g.Nhyp * S), alpha((size_t)g.N * S), sW1(S), mult(S, 1.0), L;
  std::vector<int> Lchol(S);
  if (want_L) L.resize((size_t)g.N * g.N * S);
  for (int s = 0; s < S; ++s) {
    memcpy(&hyp[(size_t)s * g.Nhyp], dbl(post, s, "hyp"), sizeof(double) * g.Nhyp);
    memcpy(&alpha[(size_t)s * g.N], dbl(post, s, "alpha"), sizeof(double) * g.N);
    sW1[s] = dbl(post, s, "sW")[0];
    Lchol[s] = mxGetScalar(mxGetField(post, s, "Lchol")) != 0;
    if (const double* m = dbl(post, s, "sn2_mult")) mult[s] = m[0];
    if (want_L) {
      const double* Ls = dbl(post, s, "L");
      if (!Ls) mexErrMsgIdAndTxt("vbmc_b200:noL", "gp.post(%d).L is empty (gplite_clean was called?).", s + 1);
      memcpy(&L[(size_t)s * g.N * g.N], Ls, sizeof(double) * g.N * g.N);
    }
  }
  g.hyp = hyp.data();
  check(vbmc_b200_gp_attach(c, &g, alpha.data(), sW1.data(), Lchol.data(), want_L ? L.data() : nullptr));
  check(vbmc_b200_gp_set_sn2_mult(c, mult.data()));
  check(vbmc_b200_gp_tag_set(c, key | (want_L ? 1ULL : 0ULL)));
  return S;
}

// ---- thetabnd struct (misc/vpbounds.m:32-52) ----
inline void thetabnd_set(vbmc_b200_ctx* c, const mxArray* tb) {
  if (!tb || mxIsEmpty(tb)) {
    check(vbmc_b200_thetabnd_set(c, 0, nullptr, nullptr, 0, 0, 0));
    return;
  }
  const mxArray* lb = mxGetField(tb, 0, "lb");
  check(vbmc_b200_thetabnd_set(c, (int)mxGetNumberOfElements(lb), mxGetDoubles(lb), dbl(tb, 0, "ub"), num(tb, "TolCon", 0),
                               num(tb, "WeightThreshold", 0), num(tb, "WeightPenalty", 0)));
}

// call counter of the device draw generator (replaces MATLAB's global randn stream, ent/entmc_vbmc.m:53); advancing by
// one per call is what lets the library generate the next call's draws ahead of time
inline unsigned long long next_stream() {
  static unsigned long long n = 0;
  return n++;
}
const unsigned long long kSeed = 0x5eed;

}  // namespace vbmex
