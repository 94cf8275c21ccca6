/*
 * vbmc_oracle.c — plain-C (C99 + OpenMP) CPU restatement of the VBMC hot path.  TEST INFRASTRUCTURE:
 * only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may link or call it; the
 * product (vbmc_b200/) never does.  PARITY UNPINNED (see oracle/vbmc_oracle.py header): the reference
 * has no golden vectors for this path and MATLAB cannot run here; this port is pinned by three-way
 * agreement with the NumPy oracle and the CUDA path plus the analytic identities in tests/.
 *
 * It follows the reference's own formulas and loop nest (not the CUDA kernels' algebra):
 *   entmc     ent/entmc_vbmc.m:49-125   xi = mu_j + sigma_j*lambda.*eps, N_k(xi) by direct exp, lsum, q_j
 *   gplogjoint misc/gplogjoint.m:98-271 + :352-413, mean functions 0/1/4
 *   penalties misc/vpbndloss.m:9-71, utils/softbndloss.m:9-28, misc/negelcbo_vbmc.m:136-164
 * The loops are fused (no D x Ns x K temporaries) and parallelised over draws / (s,k) pairs with
 * OpenMP: this is the "optimised CPU" baseline of BASELINE.md §2.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif


/* Arithmetic type.  Default: double (the CPU baseline / three-way parity port).  With -DVBMC_ORACLE_QUAD the SAME source is
 * compiled in IEEE binary128 (libquadmath) and exported as vbmc_truth128_*: the "truth" the FP64 implementations (NumPy oracle,
 * this port in double, the CUDA path) are measured against on ill-conditioned posteriors, where plain FP64 evaluation of
 * sum_n z_n*alpha_n loses 4-6 digits to cancellation (|alpha| ~ 1e4).  Inputs/outputs stay double; every intermediate is `real`. */
#ifdef VBMC_ORACLE_QUAD
#include <quadmath.h>
typedef __float128 real;
#define R_EXP expq
#define R_LOG logq
#define R_SQRT sqrtq
#define R_POW powq
#define FN(name) vbmc_truth128_##name
#define PI 3.14159265358979323846264338327950288Q
#else
typedef double real;
#define R_EXP exp
#define R_LOG log
#define R_SQRT sqrt
#define R_POW pow
#define FN(name) vbmc_oracle_##name
#define PI 3.14159265358979323846
#endif

int FN(threads)(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* explicit thread count: torchrun exports OMP_NUM_THREADS=1 to every rank and libgomp may already be initialised */
void FN(set_threads)(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

typedef struct {
  int D, K;
  real *mu, *sigma, *lambda, *w, *eta; /* mu[K][D] */
} vp_t;

/* misc/negelcbo_vbmc.m:32-48 */
static void unpack(const double* theta, int ntheta, const int* opt, const double* bmu, const double* bsig,
                   const double* blam, const double* bw, const double* beta_, vp_t* v) {
  const int D = v->D, K = v->K;
  int idx = 0;
  for (int i = 0; i < D * K; ++i) v->mu[i] = opt[0] ? theta[i] : bmu[i];
  if (opt[0]) idx = D * K;
  for (int k = 0; k < K; ++k) v->sigma[k] = opt[1] ? R_EXP(theta[idx + k]) : bsig[k];
  if (opt[1]) idx += K;
  for (int d = 0; d < D; ++d) v->lambda[d] = opt[2] ? R_EXP(theta[idx + d]) : blam[d];
  if (opt[3]) {
    real es = 0.0;
    for (int k = 0; k < K; ++k) { v->eta[k] = theta[ntheta - K + k]; es += R_EXP(v->eta[k]); }
    for (int k = 0; k < K; ++k) v->w[k] = R_EXP(v->eta[k]) / es;
  } else {
    for (int k = 0; k < K; ++k) { v->eta[k] = beta_[k]; v->w[k] = bw[k]; }
  }
}

/* (J_w g): J_w = diag(e/es) - e e'/es^2, e = R_EXP(eta)  (gplogjoint.m:366-368) */
static void softmax_jac(const real* eta, int K, const real* g, real* out) {
  real es = 0.0, dot = 0.0;
  for (int k = 0; k < K; ++k) es += R_EXP(eta[k]);
  for (int k = 0; k < K; ++k) dot += R_EXP(eta[k]) / es * g[k];
  for (int k = 0; k < K; ++k) out[k] = R_EXP(eta[k]) / es * (g[k] - dot);
}

/* ent/entmc_vbmc.m: raw (pre-Jacobian) gradient pieces.  eps: [K][half][D]. */
static void entmc(const vp_t* v, int Ns, const double* eps, int grad, real* H, real* mu_grad /*[K][D]*/,
                  real* sigma_grad /*[K]*/, real* lambda_grad /*[D]*/, real* w_grad /*[K]*/) {
  const int D = v->D, K = v->K, half = Ns / 2;
  real prodl = 1.0;
  for (int d = 0; d < D; ++d) prodl *= v->lambda[d];
  const real nf = 1.0 / R_POW(2.0 * PI, 0.5 * D) / prodl; /* :40 */
  real* cn = (real*)malloc(sizeof(real) * K);
  for (int k = 0; k < K; ++k) cn[k] = nf / R_POW(v->sigma[k], (real)D);
  *H = 0.0;
  if (grad) {
    memset(mu_grad, 0, sizeof(real) * D * K);
    memset(sigma_grad, 0, sizeof(real) * K);
    memset(lambda_grad, 0, sizeof(real) * D);
    memset(w_grad, 0, sizeof(real) * K);
  }
  const int nacc = 1 + 2 * D + K; /* log q, lsum/q [D], lsum.*eps/q [D], N_l/q [K] */
  int nt = FN(threads)();
  real* acc = (real*)malloc(sizeof(real) * nacc * nt);
  for (int j = 0; j < K; ++j) {
    memset(acc, 0, sizeof(real) * nacc * nt);
#pragma omp parallel
    {
#ifdef _OPENMP
      const int tid = omp_get_thread_num();
#else
      const int tid = 0;
#endif
      real* a = acc + (size_t)tid * nacc;
      real* x = (real*)malloc(sizeof(real) * (3 * D + K));
      real* lsum = x + D;
      real* es = lsum + D;
      real* Nl = es + D;
#pragma omp for schedule(static)
      for (int s = 0; s < Ns; ++s) {
        const int p = s < half ? s : s - half;
        const real sgn = s < half ? 1.0 : -1.0; /* antithetic, :53-54 */
        const double* e = eps + ((size_t)j * half + p) * D;
        for (int d = 0; d < D; ++d) {
          es[d] = sgn * e[d];
          x[d] = es[d] * v->lambda[d] * v->sigma[j] + v->mu[j * D + d]; /* :55 */
          lsum[d] = 0.0;
        }
        real q = 0.0;
        for (int k = 0; k < K; ++k) {
          real d2 = 0.0;
          for (int d = 0; d < D; ++d) {
            const real z = (x[d] - v->mu[k * D + d]) / (v->sigma[k] * v->lambda[d]);
            d2 += z * z;
          }
          const real nn = cn[k] * R_EXP(-0.5 * d2); /* :63 */
          Nl[k] = nn;
          q += v->w[k] * nn;
          if (grad)
            for (int d = 0; d < D; ++d) {
              const real sl = v->sigma[k] * v->lambda[d];
              lsum[d] += (x[d] - v->mu[k * D + d]) / (sl * sl) * (nn * v->w[k]); /* :77-79 */
            }
        }
        a[0] += R_LOG(q);
        if (grad) {
          for (int d = 0; d < D; ++d) {
            a[1 + d] += lsum[d] / q;
            a[1 + D + d] += lsum[d] * es[d] / q;
          }
          for (int k = 0; k < K; ++k) a[1 + 2 * D + k] += Nl[k] / q;
        }
      }
      free(x);
    }
    for (int t = 1; t < nt; ++t)
      for (int i = 0; i < nacc; ++i) acc[i] += acc[(size_t)t * nacc + i];
    *H -= v->w[j] * acc[0] / Ns; /* :67 */
    if (grad) {
      real isum = 0.0;
      for (int d = 0; d < D; ++d) {
        mu_grad[j * D + d] = v->w[j] * acc[1 + d] / Ns;                   /* :82 */
        isum += acc[1 + D + d] * v->lambda[d];                             /* :87 */
        lambda_grad[d] += v->w[j] * v->sigma[j] * acc[1 + D + d] / Ns;     /* :93 */
      }
      sigma_grad[j] = v->w[j] * isum / Ns;                                  /* :88 */
      w_grad[j] -= acc[0] / Ns;                                             /* :97 */
      for (int l = 0; l < K; ++l) w_grad[l] -= v->w[j] * acc[1 + 2 * D + l] / Ns; /* :100 */
    }
  }
  if (grad)
    for (int d = 0; d < D; ++d) lambda_grad[d] *= v->lambda[d]; /* :106-108 */
  free(acc);
  free(cn);
}

/* misc/gplogjoint.m (no variance).  X: N x D column-major, hyp: Nhyp x S, alpha: N x S.
 * Outputs averaged over s, Jacobians NOT yet applied except as noted by the caller. */
static void gplogjoint(const vp_t* v, int N, int S, int Nhyp, int Ncov, int Nnoise, int meanfun, const double* X,
                       const double* hyp, const double* alpha, const double* delta, int grad, real* Fs /*[S]*/,
                       real* I_sk /*[S][K]*/, real* mu_grad /*[S][K][D]*/, real* sigma_grad /*[S][K]*/,
                       real* lambda_grad /*[S][D]*/) {
  const int D = v->D, K = v->K;
  if (grad) memset(lambda_grad, 0, sizeof(real) * S * D);
#pragma omp parallel for collapse(2) schedule(dynamic)
  for (int s = 0; s < S; ++s) {
    for (int k = 0; k < K; ++k) {
      const double* h = hyp + (size_t)s * Nhyp;
      const double* al = alpha + (size_t)s * N;
      real tau[64], mu[64], accB[64], accC[64];
      real sum_lnell = 0.0, slt = 0.0;
      for (int d = 0; d < D; ++d) {
        const real ell = R_EXP(h[d]);
        const real dl = delta ? delta[d] : 0.0;
        sum_lnell += h[d];
        tau[d] = R_SQRT(v->sigma[k] * v->sigma[k] * v->lambda[d] * v->lambda[d] + ell * ell + dl * dl); /* :164 */
        slt += R_LOG(tau[d]);
        mu[d] = v->mu[k * D + d];
        accB[d] = accC[d] = 0.0;
      }
      const real lnnf = 2.0 * h[D] + sum_lnell - slt; /* :165 */
      real A = 0.0;
      for (int n = 0; n < N; ++n) {
        real ss = 0.0, dk[64];
        for (int d = 0; d < D; ++d) {
          dk[d] = (mu[d] - X[(size_t)d * N + n]) / tau[d]; /* :167 */
          ss += dk[d] * dk[d];
        }
        const real za = R_EXP(lnnf - 0.5 * ss) * al[n];
        A += za;
        if (grad)
          for (int d = 0; d < D; ++d) {
            accB[d] += -(dk[d] / tau[d]) * za;                 /* dz_dmu*alpha, :206-208 */
            accC[d] += (dk[d] * dk[d] - 1.0) * za;
          }
      }
      const real m0 = meanfun > 0 ? h[Ncov + Nnoise] : 0.0;
      real Ik = A + m0; /* :169 */
      real gs = 0.0;
      for (int d = 0; d < D; ++d) {
        const real lam = v->lambda[d], sg = v->sigma[k];
        real gmu = accB[d];
        real glam = (sg / tau[d]) * (sg / tau[d]) * lam * accC[d]; /* :248-249 */
        gs += (lam / tau[d]) * (lam / tau[d]) * accC[d];            /* :227 */
        if (meanfun == 4) {
          const real xm = h[Ncov + Nnoise + 1 + d], om = R_EXP(h[Ncov + Nnoise + D + 1 + d]);
          const real dl = delta ? delta[d] : 0.0;
          Ik += -0.5 / (om * om) * (mu[d] * mu[d] + sg * sg * lam * lam - 2.0 * mu[d] * xm + xm * xm + dl * dl); /* :172 */
          gmu -= (mu[d] - xm) / (om * om);    /* :210 */
          gs -= lam * lam / (om * om);        /* :231 (sigma factor applied below) */
          glam -= sg * sg / (om * om) * lam;  /* :252 */
        }
        if (grad) {
          mu_grad[((size_t)s * K + k) * D + d] = v->w[k] * gmu;
#pragma omp atomic
          lambda_grad[s * D + d] += v->w[k] * glam;
        }
      }
      if (grad) sigma_grad[s * K + k] = v->w[k] * v->sigma[k] * gs; /* :228-231 */
      I_sk[s * K + k] = Ik;
    }
  }
  for (int s = 0; s < S; ++s) {
    real f = 0.0;
    for (int k = 0; k < K; ++k) f += v->w[k] * I_sk[s * K + k]; /* :203 */
    Fs[s] = f;
  }
}

static real softbnd(real x, real lb, real ub, real tol, real* dy) { /* softbndloss.m:12-27 */
  const real ell = (ub - lb) * tol;
  *dy = 0.0;
  if (x < lb) { *dy = (x - lb) / (ell * ell); return 0.5 * ((lb - x) / ell) * ((lb - x) / ell); }
  if (x > ub) { *dy = (x - ub) / (ell * ell); return 0.5 * ((x - ub) / ell) * ((x - ub) / ell); }
  return 0.0;
}

/*
 * [F,dF,G,H,~,dH] = negelcbo_vbmc(theta,0,vp,gp,Ns,compute_grad,0,0,thetabnd)   (compute_var = 0 path)
 * Layouts: X N x D col-major; hyp Nhyp x S; alpha N x S; base mu D x K col-major ([K][D]); eps [K][Ns/2][D].
 * nbnd == 0 => thetabnd = [].  Returns 0, or -1 for unsupported D.
 */
int FN(negelcbo)(int D, int K, int N, int S, int Nhyp, int Ncov, int Nnoise, int meanfun, const double* X,
                         const double* hyp, const double* alpha, const double* theta, int ntheta, const int* opt,
                         const double* bmu, const double* bsig, const double* blam, const double* bw, const double* beta_,
                         const double* delta, int Ns, const double* eps, int nbnd, const double* lb, const double* ub,
                         double TolCon_, double WThresh_, double WPen_, int compute_grad, double* F_out, double* dF_out,
                         double* G_out, double* H_out, double* dH_out, double* I_sk_out) {
  if (D > 64) return -1;
  const real TolCon = TolCon_, WThresh = WThresh_, WPen = WPen_;
  real Fbuf, Gbuf, Hbuf;
  real *F = &Fbuf, *G = &Gbuf, *H = &Hbuf;
  real* dF = (real*)malloc(sizeof(real) * 2 * (ntheta > 0 ? ntheta : 1));
  real* dH = dH_out ? dF + ntheta : NULL;
  Ns = (Ns + 1) / 2 * 2; /* entmc_vbmc.m:45 */
  vp_t v;
  v.D = D; v.K = K;
  v.mu = (real*)malloc(sizeof(real) * (D * K + 3 * K + D));
  v.sigma = v.mu + D * K; v.lambda = v.sigma + K; v.w = v.lambda + D; v.eta = v.w + K;
  unpack(theta, ntheta, opt, bmu, bsig, blam, bw, beta_, &v);
  const int gf0 = compute_grad && opt[0], gf1 = compute_grad && opt[1], gf2 = compute_grad && opt[2], gf3 = compute_grad && opt[3];
  /* ---- G ---- */
  real* Fs = (real*)malloc(sizeof(real) * (S + (size_t)S * K * (D + 2) + S * D));
  real* I_sk = Fs + S;
  real* gmu = I_sk + S * K;
  real* gsig = gmu + (size_t)S * K * D;
  real* glam = gsig + S * K;
  gplogjoint(&v, N, S, Nhyp, Ncov, Nnoise, meanfun, X, hyp, alpha, delta, compute_grad, Fs, I_sk, gmu, gsig, glam);
  real Gv = 0.0;
  for (int s = 0; s < S; ++s) Gv += Fs[s];
  Gv /= S; /* :398-399 */
  /* ---- H ---- */
  real Hv;
  real* emu = (real*)malloc(sizeof(real) * (D * K + 2 * K + D + 3 * K));
  real* esig = emu + D * K; real* elam = esig + K; real* ew = elam + D;
  real* tmp = ew + K; real* tmp2 = tmp + K; real* pen_w = tmp2 + K;
  entmc(&v, Ns, eps, compute_grad, &Hv, emu, esig, elam, ew);
  *G = Gv; *H = Hv;
  real Fv = -Gv - Hv;
  /* ---- gradients: Jacobians (gplogjoint.m:352-369, entmc_vbmc.m:110-124), average over s ---- */
  int n = 0;
  if (compute_grad) {
    for (int i = 0; i < ntheta; ++i) dF[i] = 0.0;
    if (gf0) {
      for (int i = 0; i < D * K; ++i) {
        real g = 0.0;
        for (int s = 0; s < S; ++s) g += gmu[(size_t)s * K * D + i];
        dF[n + i] = -(g / S) - emu[i];
        if (dH) dH[n + i] = emu[i];
      }
      n += D * K;
    }
    if (gf1) {
      for (int k = 0; k < K; ++k) {
        real g = 0.0;
        for (int s = 0; s < S; ++s) g += gsig[s * K + k] * v.sigma[k];
        const real h = esig[k] * v.sigma[k];
        dF[n + k] = -(g / S) - h;
        if (dH) dH[n + k] = h;
      }
      n += K;
    }
    if (gf2) {
      for (int d = 0; d < D; ++d) {
        real g = 0.0;
        for (int s = 0; s < S; ++s) g += glam[s * D + d] * v.lambda[d];
        dF[n + d] = -(g / S) - elam[d];
        if (dH) dH[n + d] = elam[d];
      }
      n += D;
    }
    if (gf3) {
      for (int k = 0; k < K; ++k) {
        real g = 0.0;
        for (int s = 0; s < S; ++s) g += I_sk[s * K + k];
        tmp[k] = g / S;
      }
      softmax_jac(v.eta, K, tmp, tmp2);
      softmax_jac(v.eta, K, ew, tmp);
      for (int k = 0; k < K; ++k) {
        dF[n + k] = -tmp2[k] - tmp[k];
        if (dH) dH[n + k] = tmp[k];
      }
    }
  }
  /* ---- penalties ---- */
  if (nbnd > 0) {
    int b = 0, o = 0;
    real L = 0.0, dy;
    if (opt[0]) {
      for (int i = 0; i < D * K; ++i) {
        L += softbnd(v.mu[i], lb[b + i], ub[b + i], TolCon, &dy);
        if (gf0) dF[o + i] += dy;
      }
      b += D * K; o += D * K;
    }
    if (opt[1] || opt[2]) {
      const int o_sig = o, o_lam = o + (opt[1] ? K : 0);
      for (int k = 0; k < K; ++k)
        for (int d = 0; d < D; ++d) {
          const real lns = opt[1] ? theta[o_sig + k] : R_LOG(bsig[k]);
          const real lnl = opt[2] ? theta[o_lam + d] : R_LOG(blam[d]);
          L += softbnd(lns + lnl, lb[b + k * D + d], ub[b + k * D + d], TolCon, &dy);
          if (gf1) dF[o_sig + k] += dy;
          if (gf2) dF[o_lam + d] += dy;
        }
      b += D * K;
    }
    if (opt[3]) {
      const int o_w = ntheta - K;
      real Lw = 0.0;
      for (int k = 0; k < K; ++k) {
        L += softbnd(v.eta[k], lb[b + k], ub[b + k], TolCon, &dy);
        if (gf3) dF[o_w + k] += dy;
        Lw += (v.w[k] < WThresh) ? v.w[k] : WThresh;   /* negelcbo_vbmc.m:150 */
        pen_w[k] = WPen * ((v.w[k] < WThresh) ? 1.0 : 0.0);
      }
      L += Lw * WPen;
      if (gf3) {
        softmax_jac(v.eta, K, pen_w, tmp);
        for (int k = 0; k < K; ++k) dF[o_w + k] += tmp[k];
      }
    }
    Fv += L;
  }
  *F = Fv;
  *F_out = (double)*F; *G_out = (double)*G; *H_out = (double)*H;
  if (compute_grad) {
    for (int i = 0; i < ntheta; ++i) {
      if (dF_out) dF_out[i] = (double)dF[i];
      if (dH_out) dH_out[i] = (double)dH[i];
    }
  }
  if (I_sk_out)
    for (int i = 0; i < S * K; ++i) I_sk_out[i] = (double)I_sk[i];
  free(dF); free(emu); free(Fs); free(v.mu);
  return 0;
}
