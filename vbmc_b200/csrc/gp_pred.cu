// gplite_pred on sm_100a (reference: gplite/gplite_pred.m:52-163): posterior mean and variance of
// the GP surrogate at Nstar test points, for every hyper-parameter sample of the attached posterior.
//   Ks = sf2 exp(-0.5 sq_dist(X/ell, Xstar/ell))            (:68-72)    fmu = mstar + Ks' alpha        (:80)
//   V  = L' \ (sW .* Ks),  fs2 = max(kss - sum(V.*V), 0)    (:96-99,118)  ys2 = fs2 + sn2_star*sn2_mult  (:119)
// The cross-kernel column of one (test point, sample) is produced by one CTA, which also takes its dot product with
// alpha (the same streaming contraction as gplogjoint's z.alpha); the Nstar columns are the right-hand sides of ONE
// forward substitution per sample (var_fwd_kernel, factor read once per 8 columns), not Nstar separate solves.
// Low-noise posteriors handed over as L = -inv(K+Sigma) use the symmetric product instead (:100-102).
#include <math.h>

#include <utility>

#include "common.cuh"

namespace vb {

struct PredArgs {
  int N, D, S, T;        // training points, dimension, samples, test points in this chunk
  int meanfun, Ncov, Nnoise, Nhyp;
  const double* X;       // [D][N]
  const double* Xs;      // [D][T]  test points of this chunk (column-major T x D)
  const double* hyp;     // [S][Nhyp]
  const double* alpha;   // [S][N]
  const double* sn2eff;  // [S]
  const int* isfac;      // [S]
  double* Z;             // [S][T][N] or null (mean only)
  const double* W;       // [S][T][N] K^-1-products of the samples without a factor
  double* fmu;           // [S][T]
  double* fs2;           // [S][T]
};

__device__ __forceinline__ double block_sum_256(double v, double* part) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((tid & 31) == 0) part[tid >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < 8; ++i) t += part[i];
  return t;
}

// grid (T, S), 256 threads
__global__ void __launch_bounds__(256) pred_cross_kernel(const PredArgs a) {
  __shared__ double xs[32], il[32], part[8];
  __shared__ double s_m;
  const int t = blockIdx.x, s = blockIdx.y, tid = threadIdx.x, D = a.D, N = a.N;
  const double* h = a.hyp + static_cast<size_t>(s) * a.Nhyp;
  if (tid < D) {
    il[tid] = exp(-h[tid]);
    xs[tid] = a.Xs[static_cast<size_t>(tid) * a.T + t] * il[tid];
  }
  if (tid == 32) {  // mstar (gplite_meanfun.m cases 0, 1, 4)
    const double* hm = h + a.Ncov + a.Nnoise;
    double m = 0.0;
    if (a.meanfun == 1) m = hm[0];
    if (a.meanfun == 4) {
      double z2 = 0.0;
      for (int d = 0; d < D; ++d) {
        const double z = (a.Xs[static_cast<size_t>(d) * a.T + t] - hm[1 + d]) / exp(hm[1 + D + d]);
        z2 = fma(z, z, z2);
      }
      m = hm[0] - 0.5 * z2;
    }
    s_m = m;
  }
  __syncthreads();
  const double sf2 = exp(2.0 * h[D]);
  const double zscale = a.isfac[s] ? 1.0 / sqrt(a.sn2eff[s]) : 1.0;  // sW (gplite_pred.m:97)
  const double* al = a.alpha + static_cast<size_t>(s) * N;
  double* z = a.Z ? a.Z + (static_cast<size_t>(s) * a.T + t) * N : nullptr;
  double acc = 0.0;
  for (int n = tid; n < N; n += 256) {
    double ss = 0.0;
    for (int d = 0; d < D; ++d) {
      const double df = a.X[static_cast<size_t>(d) * N + n] * il[d] - xs[d];
      ss = fma(df, df, ss);
    }
    const double k = sf2 * exp(-0.5 * ss);
    if (z) z[n] = k * zscale;
    acc = fma(k, al[n], acc);
  }
  acc = block_sum_256(acc, part);
  if (tid == 0) a.fmu[static_cast<size_t>(s) * a.T + t] = s_m + acc;  // (:80)
}

// grid (T, S): fs2 = max(kss - sum(V.*V), 0)  /  max(kss + sum(Ks.*(L*Ks)), 0)
__global__ void __launch_bounds__(256) pred_var_kernel(const PredArgs a) {
  __shared__ double part[8];
  const int t = blockIdx.x, s = blockIdx.y, tid = threadIdx.x, N = a.N;
  const double* z = a.Z + (static_cast<size_t>(s) * a.T + t) * N;
  double acc = 0.0;
  if (a.isfac[s]) {
    for (int n = tid; n < N; n += 256) acc = fma(z[n], z[n], acc);
  } else {
    const double* w = a.W + (static_cast<size_t>(s) * a.T + t) * N;
    for (int n = tid; n < N; n += 256) acc = fma(z[n], w[n], acc);
  }
  acc = block_sum_256(acc, part);
  if (tid == 0) {
    const double sf2 = exp(2.0 * a.hyp[static_cast<size_t>(s) * a.Nhyp + a.D]);  // kss (:72)
    a.fs2[static_cast<size_t>(s) * a.T + t] = fmax(sf2 - acc, 0.0);               // (:99,118)
  }
}


// ---- rank-one update of the posterior (gplite_post.m:173-251) ----
// par[s] = {gamma = (mstar - ystar)/vstar, colscale, diagadd, 1/vstar}
// Device factor (Cholesky branch :226-233, and the low-noise samples a device refit keeps as R'R = K + sn2_mult*sn2*I):
// new_L_column = (L'\Ks)/sn2_eff (:229-230) is the forward-substitution result V = L'\(sW.*Ks) rescaled by
// colscale = sqrt(sn2_eff_old)/sn2_eff_new (1 for the unscaled low-noise factor); L(N+1,N+1) = sqrt(diagadd - c'c) with
// diagadd = 1 + K/sn2_eff (:231-233), respectively K + sn2_eff.  grid (S), 256 threads; inverse-form samples are skipped.
__global__ void __launch_bounds__(256) rank1_column_kernel(int N, int ld, double* L, double* Z, const double* par, const int* isfac) {
  __shared__ double part[8];
  const int s = blockIdx.x, tid = threadIdx.x;
  if (!isfac[s]) return;
  double* z = Z + static_cast<size_t>(s) * N;
  double* col = L + static_cast<size_t>(s) * ld * ld + static_cast<size_t>(N) * ld;
  const double cs = par[4 * s + 1];
  double acc = 0.0;
  for (int n = tid; n < N; n += 256) {
    const double cv = z[n] * cs;
    z[n] = cv;      // right-hand side of the backward solve: alpha_update = L\c  (== (L\(L'\Ks))/sn2_eff, :228)
    col[n] = cv;
    acc = fma(cv, cv, acc);
  }
  acc = block_sum_256(acc, part);
  if (tid == 0) col[N] = sqrt(par[4 * s + 2] - acc);
}

// Low-noise posterior in inverse form, L = -inv(K + diag) handed over by gp_attach (:234-238): with a = -L*Ks (= W of
// var_symv_kernel) and v = -a/vstar,  L <- [L + v a', -v; -v', -1/vstar].  grid (ceil((N+1)/256), N+1, S).
__global__ void __launch_bounds__(256) rank1_inverse_kernel(int N, int ld, double* L, const double* W, const double* par, const int* isfac) {
  const int s = blockIdx.z, j = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (isfac[s] || i > N) return;
  const double* w = W + static_cast<size_t>(s) * N;
  const double iv = par[4 * s + 3];
  double* e = L + static_cast<size_t>(s) * ld * ld + static_cast<size_t>(j) * ld + i;
  if (i < N && j < N)
    *e = fma(-w[i] * iv, w[j], *e);
  else if (i == N && j == N)
    *e = -iv;
  else
    *e = w[i < N ? i : j] * iv;   // -v = a/vstar
}

// alpha = [alpha; 0] + (mstar - ystar)/vstar * [alpha_update; -1]   (:245-247); grid (S)
// alpha_update = L\(L'\Ks)/sn2_eff in Z (factor samples) or -L*Ks in Winv (inverse-form samples)
__global__ void __launch_bounds__(256) rank1_alpha_kernel(int N, const double* alpha_old, const double* Z, const double* Winv,
                                                          const int* isfac, const double* par, double* alpha_new) {
  const int s = blockIdx.x;
  const double g = par[4 * s];
  const double* ao = alpha_old + static_cast<size_t>(s) * N;
  const double* w = (isfac[s] ? Z : Winv) + static_cast<size_t>(s) * N;
  double* an = alpha_new + static_cast<size_t>(s) * (N + 1);
  for (int n = threadIdx.x; n < N; n += 256) an[n] = fma(g, w[n], ao[n]);
  if (threadIdx.x == 0) an[N] = -g;
}

}  // namespace vb

using namespace vb;

extern "C" int vbmc_b200_gp_pred(vbmc_b200_ctx* c, int Nstar, const double* Xstar, const double* ystar, const double* s2star,
                                 int ssflag, int want_var, double* ymu, double* ys2, double* fmu, double* fs2, double* lp) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  if (!c->gp_ready) VB_FAIL(VBMC_B200_ESTATE, "gplite_pred: call vbmc_b200_gp_attach or vbmc_b200_gp_post first");
  if (Nstar <= 0 || !Xstar) VB_FAIL(VBMC_B200_EINVAL, "gplite_pred: Xstar (Nstar x D) is required");
  if (want_var && !c->gpHasL)
    VB_FAIL(VBMC_B200_ESTATE, "gplite_pred: predictive variances need the factors gp.post(s).L on the device (gp_attach with L, or gp_post)");
  if ((c->gp_noisefun[1] == 1 || c->gp_noisefun[1] == 2) && want_var && !s2star)
    VB_FAIL(VBMC_B200_EINVAL, "gplite_pred: this GP has user-provided noise (noisefun(2) > 0): s2star is required for ys2");
  VB_CUDA(cudaSetDevice(c->device));
  const int N = c->gp.N, D = c->gp.D, S = c->gp.S;
  cudaStream_t st = c->stream;
  // chunk the test points so that the right-hand-side block stays below ~2 GB
  size_t per_t = static_cast<size_t>(S) * N * sizeof(double) * (want_var ? 1 : 0);
  bool any_inv = false;
  for (int s = 0; s < S; ++s) any_inv = any_inv || !c->gpLfactor[s];
  if (any_inv) per_t *= 2;
  int chunk = Nstar;
  if (per_t > 0) {
    const size_t cap = (2ull << 30) / per_t;
    chunk = static_cast<int>(cap < 8 ? 8 : (cap > static_cast<size_t>(Nstar) ? static_cast<size_t>(Nstar) : cap));
  }
  const size_t nz = want_var ? static_cast<size_t>(S) * chunk * N : 0;
  const size_t nhead = static_cast<size_t>(D) * chunk + 2 * static_cast<size_t>(S) * chunk + (static_cast<size_t>(S) + 1) / 2;
  VB_TRY(c->predWork.reserve(sizeof(double) * (nhead + nz * (any_inv ? 2 : 1))));
  double* d_xs = c->predWork.d();
  double* d_fmu = d_xs + static_cast<size_t>(D) * chunk;
  double* d_fs2 = d_fmu + static_cast<size_t>(S) * chunk;
  int* d_isfac = reinterpret_cast<int*>(d_fs2 + static_cast<size_t>(S) * chunk);
  double* d_Z = want_var ? d_fs2 + static_cast<size_t>(S) * chunk + (static_cast<size_t>(S) + 1) / 2 : nullptr;
  double* d_W = (want_var && any_inv) ? d_Z + nz : nullptr;
  VB_CUDA(cudaMemcpyAsync(d_isfac, c->gpLfactor.data(), sizeof(int) * S, cudaMemcpyHostToDevice, st));
  std::vector<double> hf(static_cast<size_t>(S) * Nstar), hv(want_var ? static_cast<size_t>(S) * Nstar : 0), xs_chunk, tmp;
  for (int t0 = 0; t0 < Nstar; t0 += chunk) {
    const int T = (Nstar - t0) < chunk ? (Nstar - t0) : chunk;
    xs_chunk.resize(static_cast<size_t>(D) * T);
    for (int d = 0; d < D; ++d) memcpy(&xs_chunk[static_cast<size_t>(d) * T], Xstar + static_cast<size_t>(d) * Nstar + t0, sizeof(double) * T);
    VB_CUDA(cudaMemcpyAsync(d_xs, xs_chunk.data(), sizeof(double) * D * T, cudaMemcpyHostToDevice, st));
    PredArgs a;
    a.N = N; a.D = D; a.S = S; a.T = T;
    a.meanfun = c->gp.meanfun; a.Ncov = c->gp.Ncov; a.Nnoise = c->gp.Nnoise; a.Nhyp = c->gp.Nhyp;
    a.X = c->gp.X; a.Xs = d_xs; a.hyp = c->gp.hyp; a.alpha = c->gp.alpha; a.sn2eff = c->gp.sn2eff;
    a.isfac = d_isfac; a.Z = d_Z; a.W = d_W; a.fmu = d_fmu; a.fs2 = d_fs2;
    {
      KernelScope ks(c, "pred_cross", st);
      pred_cross_kernel<<<dim3(T, S), 256, 0, st>>>(a);
      VB_CUDA(cudaGetLastError());
    }
    if (want_var) {
      VB_TRY(run_rhs_solve(c, T, d_Z, d_W, d_isfac, st));
      KernelScope ks(c, "pred_var", st);
      pred_var_kernel<<<dim3(T, S), 256, 0, st>>>(a);
      VB_CUDA(cudaGetLastError());
    }
    tmp.resize(static_cast<size_t>(S) * T);
    VB_CUDA(cudaMemcpyAsync(tmp.data(), d_fmu, sizeof(double) * S * T, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    VB_TRY(trsv1_check(c));
    for (int s = 0; s < S; ++s) memcpy(&hf[static_cast<size_t>(s) * Nstar + t0], &tmp[static_cast<size_t>(s) * T], sizeof(double) * T);
    if (want_var) {
      VB_CUDA(cudaMemcpyAsync(tmp.data(), d_fs2, sizeof(double) * S * T, cudaMemcpyDeviceToHost, st));
      VB_CUDA(cudaStreamSynchronize(st));
      for (int s = 0; s < S; ++s) memcpy(&hv[static_cast<size_t>(s) * Nstar + t0], &tmp[static_cast<size_t>(s) * T], sizeof(double) * T);
    }
  }
  // ---- O(Nstar S) host epilogue: observation noise at the test points, log density, average over samples ----
  const bool separate = ssflag || S == 1;
  const double TWO_PI = 6.283185307179586;
  std::vector<double> ysv(want_var ? static_cast<size_t>(S) * Nstar : 0);
  for (int s = 0; s < S && want_var; ++s) {
    const double* hn = c->gpHypHost.data() + static_cast<size_t>(s) * c->gp.Nhyp + c->gp.Ncov;
    const double mult = s < static_cast<int>(c->gpSn2mult.size()) ? c->gpSn2mult[s] : 1.0;
    for (int t = 0; t < Nstar; ++t) {
      int idx = 0;
      double sn2 = 2.220446049250313e-16;                      // gplite_noisefun.m:177-184
      if (c->gp_noisefun[0] == 1) sn2 = exp(2.0 * hn[idx++]);
      if (c->gp_noisefun[1] == 1) sn2 += s2star[t];            // :186-194
      else if (c->gp_noisefun[1] == 2) sn2 += exp(hn[idx++]) * s2star[t];
      if (c->gp_noisefun[2] == 1 && ystar) {                   // :196-209
        const double zz = fmax(0.0, hn[idx] - ystar[t]);
        sn2 += exp(2.0 * hn[idx + 1]) * zz * zz;
      }
      const size_t i = static_cast<size_t>(s) * Nstar + t;
      ysv[i] = hv[i] + sn2 * mult;                             // gplite_pred.m:119
      if (lp && ystar) {                                       // :124 (per sample, never averaged)
        const double r = ystar[t] - hf[i];
        lp[i] = -0.5 * r * r / ysv[i] - 0.5 * log(TWO_PI * ysv[i]);
      }
    }
  }
  if (separate) {
    const size_t n = static_cast<size_t>(S) * Nstar;
    if (fmu) memcpy(fmu, hf.data(), sizeof(double) * n);
    if (ymu) memcpy(ymu, hf.data(), sizeof(double) * n);       // ymu = fmu (:92)
    if (want_var && fs2) memcpy(fs2, hv.data(), sizeof(double) * n);
    if (want_var && ys2) memcpy(ys2, ysv.data(), sizeof(double) * n);
  } else {                                                      // :153-163
    for (int t = 0; t < Nstar; ++t) {
      double fbar = 0.0;
      for (int s = 0; s < S; ++s) fbar += hf[static_cast<size_t>(s) * Nstar + t];
      fbar /= S;
      if (fmu) fmu[t] = fbar;
      if (ymu) ymu[t] = fbar;
      if (want_var) {
        double vf = 0.0, sf = 0.0, sy = 0.0;
        for (int s = 0; s < S; ++s) {
          const size_t i = static_cast<size_t>(s) * Nstar + t;
          vf += (hf[i] - fbar) * (hf[i] - fbar);
          sf += hv[i];
          sy += ysv[i];
        }
        vf /= (S - 1);
        if (fs2) fs2[t] = sf / S + vf;
        if (ys2) ys2[t] = sy / S + vf;                         // ymu == fmu => vy == vf
      }
    }
  }
  return VBMC_B200_OK;
}

extern "C" int vbmc_b200_gp_set_sn2_mult(vbmc_b200_ctx* c, const double* sn2_mult) {
  if (!c || !sn2_mult) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  if (!c->gp_ready) VB_FAIL(VBMC_B200_ESTATE, "gp_set_sn2_mult: no GP attached");
  c->gpSn2mult.assign(sn2_mult, sn2_mult + c->gp.S);
  return VBMC_B200_OK;
}

// gp = gplite_post(gp,xstar,ystar,[],[],[],[],1) — rank-one update, gplite/gplite_post.m:50-92,173-251.
// One new training point is folded into the resident posterior: prediction at xstar (one cross-kernel column + forward
// substitution per sample), the new factor column, a backward substitution for alpha_update, O(N) updates.  N^2 work
// per sample instead of the N^3/3 refit.  Outputs (optional): alpha (N+1) x S, the new factor column (N+1) x S, the new sW entry.
extern "C" int vbmc_b200_gp_post_update1(vbmc_b200_ctx* c, const double* xstar, double ystar, double* alpha_out,
                                         double* Lcol_out, double* sW_out) {
  if (!c || !xstar) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  c->gp_tag = 0;  // the resident posterior is about to change: a caller's fingerprint of it no longer holds
  if (!c->gp_ready)
    VB_FAIL(VBMC_B200_EREFERENCE, "gplite_post:NoGP: GPLITE_POST can perform rank-one update only with an existing GP struct.");
  if (!c->gpHasL) VB_FAIL(VBMC_B200_ESTATE, "gplite_post (rank-one): the factors gp.post(s).L must be resident (gp_attach with L, or gp_post)");
  if (c->gp_noisefun[1] != 0)
    VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:FullUpdate: rank-one updates are not defined for heteroskedastic noise (gplite_post.m:78-81): refit with gplite_post");
  const int N = c->gp.N, D = c->gp.D, S = c->gp.S;
  bool any_inv = false;
  for (int s = 0; s < S; ++s) any_inv = any_inv || !c->gpLfactor[s];
  VB_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  // ---- [mstar,vstar] = gplite_pred(gp,xstar,y,s2,1,1)  (:191) ----
  const size_t nhead = static_cast<size_t>(D) + 2 * S + (static_cast<size_t>(S) + 1) / 2 + 4 * static_cast<size_t>(S);
  VB_TRY(c->predWork.reserve(sizeof(double) * (nhead + (any_inv ? 2 : 1) * static_cast<size_t>(S) * N)));
  double* d_xs = c->predWork.d();
  double* d_fmu = d_xs + D;
  double* d_fs2 = d_fmu + S;
  int* d_isfac = reinterpret_cast<int*>(d_fs2 + S);
  double* d_par = d_fs2 + S + (static_cast<size_t>(S) + 1) / 2;
  double* d_Z = d_par + 4 * static_cast<size_t>(S);
  double* d_W = any_inv ? d_Z + static_cast<size_t>(S) * N : nullptr;   // -L*Ks of the inverse-form samples
  VB_CUDA(cudaMemcpyAsync(d_isfac, c->gpLfactor.data(), sizeof(int) * S, cudaMemcpyHostToDevice, st));
  VB_CUDA(cudaMemcpyAsync(d_xs, xstar, sizeof(double) * D, cudaMemcpyHostToDevice, st));
  PredArgs a;
  a.N = N; a.D = D; a.S = S; a.T = 1;
  a.meanfun = c->gp.meanfun; a.Ncov = c->gp.Ncov; a.Nnoise = c->gp.Nnoise; a.Nhyp = c->gp.Nhyp;
  a.X = c->gp.X; a.Xs = d_xs; a.hyp = c->gp.hyp; a.alpha = c->gp.alpha; a.sn2eff = c->gp.sn2eff;
  a.isfac = d_isfac; a.Z = d_Z; a.W = d_W; a.fmu = d_fmu; a.fs2 = d_fs2;
  {
    KernelScope ks(c, "pred_cross", st);
    pred_cross_kernel<<<dim3(1, S), 256, 0, st>>>(a);
    VB_CUDA(cudaGetLastError());
  }
  VB_TRY(run_rhs_solve(c, 1, d_Z, d_W, d_isfac, st));
  {
    KernelScope ks(c, "pred_var", st);
    pred_var_kernel<<<dim3(1, S), 256, 0, st>>>(a);
    VB_CUDA(cudaGetLastError());
  }
  std::vector<double> hf(S), hv(S), par(4 * static_cast<size_t>(S)), sw(S);
  VB_CUDA(cudaMemcpyAsync(hf.data(), d_fmu, sizeof(double) * S, cudaMemcpyDeviceToHost, st));
  VB_CUDA(cudaMemcpyAsync(hv.data(), d_fs2, sizeof(double) * S, cudaMemcpyDeviceToHost, st));
  VB_CUDA(cudaStreamSynchronize(st));
  for (int s = 0; s < S; ++s) {
    const double* h = c->gpHypHost.data() + static_cast<size_t>(s) * c->gp.Nhyp;
    const double* hn = h + c->gp.Ncov;
    int idx = 0;
    double sn2 = 2.220446049250313e-16;                       // gplite_noisefun.m:177-184
    if (c->gp_noisefun[0] == 1) sn2 = exp(2.0 * hn[idx++]);
    if (c->gp_noisefun[2] == 1) {                             // :196-209
      const double zz = fmax(0.0, hn[idx] - ystar);
      sn2 += exp(2.0 * hn[idx + 1]) * zz * zz;
    }
    const double sn2_eff = sn2 * c->gpSn2mult[s];             // gplite_post.m:209
    const double vstar = hv[s] + sn2 * c->gpSn2mult[s];       // ys2 of gplite_pred (:119)
    par[4 * s] = (hf[s] - ystar) / vstar;                     // :247
    if (c->gpLchol[s]) {
      par[4 * s + 1] = sqrt(c->gpSn2effHost[s]) / sn2_eff;    // V = (L'\Ks)/sqrt(sn2_eff_old)  ->  (L'\Ks)/sn2_eff
      par[4 * s + 2] = 1.0 + exp(2.0 * h[D]) / sn2_eff;       // :233
    } else {                                                  // low-noise sample kept as the unscaled factor of K + sn2_eff*I:
      par[4 * s + 1] = 1.0;                                   // the bordered matrix [A Ks; Ks' K + sn2_eff] has the factor [R c; 0 d],
      par[4 * s + 2] = exp(2.0 * h[D]) + sn2_eff;             // c = R'\Ks, d^2 = K + sn2_eff - c'c; its negated inverse is :236
    }
    par[4 * s + 3] = 1.0 / vstar;
    sw[s] = 1.0 / sqrt(sn2_eff);                              // :242
  }
  VB_CUDA(cudaMemcpyAsync(d_par, par.data(), sizeof(double) * par.size(), cudaMemcpyHostToDevice, st));
  // ---- grow the resident arrays by one point ----
  const int N1 = N + 1;
  int ld = c->gpLd;
  if (N1 > ld) {
    const int ld2 = (N1 + 63) / 64 * 64;
    DevBuf nb;
    VB_TRY(nb.reserve(sizeof(double) * static_cast<size_t>(S) * ld2 * ld2));
    VB_CUDA(cudaMemsetAsync(nb.p, 0, sizeof(double) * static_cast<size_t>(S) * ld2 * ld2, st));
    for (int s = 0; s < S; ++s)
      VB_CUDA(cudaMemcpy2DAsync(nb.d() + static_cast<size_t>(s) * ld2 * ld2, sizeof(double) * ld2,
                                c->gpL.d() + static_cast<size_t>(s) * ld * ld, sizeof(double) * ld, sizeof(double) * N, N,
                                cudaMemcpyDeviceToDevice, st));
    VB_TRY(pad_identity(nb.d(), N, ld2, S, st));   // unit diagonal in the padding (trsm.cu reads whole tiles)
    VB_CUDA(cudaStreamSynchronize(st));
    c->gpL.release();
    c->gpL = nb;
    c->gpLd = ld = ld2;
  }
  {
    KernelScope ks(c, "rank1", st);
    rank1_column_kernel<<<S, 256, 0, st>>>(N, ld, c->gpL.d(), d_Z, d_par, d_isfac);
    if (any_inv) rank1_inverse_kernel<<<dim3(N / 256 + 1, N + 1, S), 256, 0, st>>>(N, ld, c->gpL.d(), d_W, d_par, d_isfac);
    VB_CUDA(cudaGetLastError());
  }
  VB_TRY(run_rhs_backsolve(c, 1, d_Z, st, d_isfac));          // alpha_update (old N x N factor)
  // the grown alpha / X go into the ping-pong partners of the resident buffers (reserved with room for 256 more points)
  DevBuf& na = c->gpAlphaAlt;
  DevBuf& nx = c->gpXalt;
  if (na.cap < sizeof(double) * static_cast<size_t>(S) * N1) VB_TRY(na.reserve(sizeof(double) * static_cast<size_t>(S) * (N1 + 256)));
  if (nx.cap < sizeof(double) * static_cast<size_t>(D) * N1) VB_TRY(nx.reserve(sizeof(double) * static_cast<size_t>(D) * (N1 + 256)));
  {
    KernelScope ks(c, "rank1", st);
    rank1_alpha_kernel<<<S, 256, 0, st>>>(N, c->gp.alpha, d_Z, d_W ? d_W : d_Z, d_isfac, d_par, na.d());
    VB_CUDA(cudaGetLastError());
  }
  VB_CUDA(cudaMemcpy2DAsync(nx.p, sizeof(double) * N1, c->gpX.p, sizeof(double) * N, sizeof(double) * N, D, cudaMemcpyDeviceToDevice, st));
  VB_CUDA(cudaMemcpy2DAsync(nx.d() + N, sizeof(double) * N1, d_xs, sizeof(double), sizeof(double), D, cudaMemcpyDeviceToDevice, st));
  if (alpha_out) VB_CUDA(cudaMemcpyAsync(alpha_out, na.p, sizeof(double) * static_cast<size_t>(S) * N1, cudaMemcpyDeviceToHost, st));
  if (Lcol_out)
    VB_CUDA(cudaMemcpy2DAsync(Lcol_out, sizeof(double) * N1, c->gpL.d() + static_cast<size_t>(N) * ld, sizeof(double) * ld * ld,
                              sizeof(double) * N1, S, cudaMemcpyDeviceToHost, st));
  VB_CUDA(cudaStreamSynchronize(st));
  VB_TRY(trsv1_check(c));
  if (sW_out) memcpy(sW_out, sw.data(), sizeof(double) * S);
  std::swap(c->gpAlpha, c->gpAlphaAlt);
  std::swap(c->gpX, c->gpXalt);
  c->gp.N = N1;
  c->gp.X = c->gpX.d();
  c->gp.alpha = c->gpAlpha.d();
  return VBMC_B200_OK;
}
