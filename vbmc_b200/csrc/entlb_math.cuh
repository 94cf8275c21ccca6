// Deterministic entropy lower bound of the variational mixture and its gradient — ent/entlb_vbmc.m:1-147
// (Gershman et al. 2012), what negelcbo_vbmc evaluates instead of entmc_vbmc when Ns == 0 (negelcbo_vbmc.m:102-109):
// misc/vpsieve_vbmc.m:76 scores every candidate posterior with it, and the whole optimisation uses it when K == 1 or
// EntropySwitch is on (vpsieve_vbmc.m:29-33).  O(K^2 D) work on a few hundred numbers: latency, not throughput.
//
// Written as element functions — one call computes ONE output element from global arrays, no shared memory, no
// inter-thread communication — so that the very same source runs as CUDA kernels (entlb.cu: one thread per element,
// four tiny launches) and, compiled by g++ with VB_HD empty, as plain loops in tests/test_entlb_host.py, which checks
// every element against the NumPy restatement on the CPU.
//
//   gamma_jk = nconst / sumsigma_jk^D * exp(-0.5 d2_jk),  sumsigma2_jk = sigma_j^2 + sigma_k^2            (:76-82)
//   gsum_k   = sum_j w_j gamma_jk                                                                       (:83)
//   H        = -sum_k w_k log gsum_k                                                                    (:85)
#pragma once
#include <math.h>

#ifndef VB_HD
#ifdef __CUDACC__
#define VB_HD __host__ __device__ __forceinline__
#else
#define VB_HD inline
#endif
#endif

namespace vb {

struct EntlbArgs {
  int D, K;
  int gf[4];      // grad_flags: mu, sigma, lambda, w
  int jacobian;   // jacobian_flag
  const double* mu;      // [K][D]  (== MATLAB D x K column-major)
  const double* sigma;   // [K]
  const double* lambda;  // [D]
  const double* w;       // [K]
  const double* eta;     // [K]
  double* gamma;         // [K][K] work
  double* gsum;          // [K]    work
  double* wraw;          // [K]    work: w-gradient before the softmax Jacobian
  double* out;           // [1 + n]: H, then dH = [mu_grad(:); sigma_grad; lambda_grad; w_grad] for the requested blocks
};

VB_HD int entlb_ngrad(const EntlbArgs& a) {
  return (a.gf[0] ? a.D * a.K : 0) + (a.gf[1] ? a.K : 0) + (a.gf[2] ? a.D : 0) + (a.gf[3] ? a.K : 0);
}

// ---- stage A: idx in [0, K*K) ----
VB_HD void entlb_gamma_elem(const EntlbArgs& a, int idx) {
  const int D = a.D, K = a.K, j = idx / K, k = idx - j * K;
  double pl = 1.0;
  for (int d = 0; d < D; ++d) pl *= a.lambda[d];
  const double nconst = 1.0 / pow(2.0 * 3.14159265358979323846, 0.5 * D) / pl;         // :79
  const double ss = sqrt(a.sigma[j] * a.sigma[j] + a.sigma[k] * a.sigma[k]);            // :76-77
  double d2 = 0.0;
  for (int d = 0; d < D; ++d) {
    const double t = (a.mu[j * D + d] - a.mu[k * D + d]) / (ss * a.lambda[d]);
    d2 += t * t;                                                                        // :81
  }
  a.gamma[idx] = nconst / pow(ss, (double)D) * exp(-0.5 * d2);                           // :82
}

// ---- stage B: idx in [0, K) ----
VB_HD void entlb_gsum_elem(const EntlbArgs& a, int k) {
  double s = 0.0;
  for (int j = 0; j < a.K; ++j) s += a.w[j] * a.gamma[j * a.K + k];                      // :83
  a.gsum[k] = s;
}

// ---- stage C: idx in [0, 1 + D*K + K + D + K): H, mu_grad[j][d], sigma_grad[j], lambda_grad[d], wraw[j] ----
VB_HD void entlb_grad_elem(const EntlbArgs& a, int idx) {
  const int D = a.D, K = a.K;
  int o_mu = 1, o_sig = o_mu + (a.gf[0] ? D * K : 0), o_lam = o_sig + (a.gf[1] ? K : 0), o_w = o_lam + (a.gf[2] ? D : 0);
  if (idx == 0) {                                                                       // H
    if (K == 1) {
      double sl = 0.0;
      for (int d = 0; d < D; ++d) sl += log(a.lambda[d]);
      a.out[0] = 0.5 * D * (1.0 + log(2.0 * 3.14159265358979323846)) + D * log(a.sigma[0]) + sl;   // :34
    } else {
      double H = 0.0;
      for (int k = 0; k < K; ++k) H -= a.w[k] * log(a.gsum[k]);                          // :85
      a.out[0] = H;
    }
    return;
  }
  int r = idx - 1;
  if (r < D * K) {                                                                      // mu_grad(d,j)
    if (!a.gf[0]) return;
    const int j = r / D, d = r - j * D;
    double g = 0.0;
    if (K > 1) {
      double m1 = 0.0, m2 = 0.0;
      for (int k = 0; k < K; ++k) {
        const double ss2 = a.sigma[j] * a.sigma[j] + a.sigma[k] * a.sigma[k];
        const double dmu = (a.mu[k * D + d] - a.mu[j * D + d]) / (ss2 * a.lambda[d] * a.lambda[d]);   // :94
        const double gm = a.gamma[j * K + k];
        m1 += a.w[k] * gm / a.gsum[k] * dmu;                                              // :104
        m2 += dmu * gm * a.w[k];                                                          // :105
      }
      g = -a.w[j] * (m1 + m2 / a.gsum[j]);                                                // :106
    }
    a.out[o_mu + r] = g;
    return;
  }
  r -= D * K;
  if (r < K) {                                                                          // sigma_grad(j)
    if (!a.gf[1]) return;
    const int j = r;
    double g;
    if (K == 1) {
      g = D / a.sigma[0];                                                               // :38-39
    } else {
      double s1 = 0.0, s2 = 0.0;
      for (int k = 0; k < K; ++k) {
        const double ss2 = a.sigma[j] * a.sigma[j] + a.sigma[k] * a.sigma[k];
        double q = 0.0;
        for (int d = 0; d < D; ++d) {
          const double t = (a.mu[j * D + d] - a.mu[k * D + d]) / a.lambda[d];
          q += t * t;
        }
        const double ds = -D / ss2 + 1.0 / (ss2 * ss2) * q;                              // :97
        const double gm = a.gamma[j * K + k];
        s1 += a.w[k] * gm / a.gsum[k] * ds;                                               // :111
        s2 += ds * gm * a.w[k];                                                           // :112
      }
      g = -a.w[j] * a.sigma[j] * (s1 + s2 / a.gsum[j]);                                   // :113
    }
    if (a.jacobian) g *= a.sigma[j];                                                    // :139-141
    a.out[o_sig + j] = g;
    return;
  }
  r -= K;
  if (r < D) {                                                                          // lambda_grad(d)
    if (!a.gf[2]) return;
    const int d = r;
    double g;
    if (K == 1) {
      g = 1.0;                                                                          // :42-43
    } else {
      g = 0.0;
      for (int k = 0; k < K; ++k) {
        double inner = 0.0;
        for (int j = 0; j < K; ++j) {
          const double ss2 = a.sigma[j] * a.sigma[j] + a.sigma[k] * a.sigma[k];
          const double t = a.mu[k * D + d] - a.mu[j * D + d];
          const double dmu2 = t * t / (ss2 * a.lambda[d] * a.lambda[d]);                  // :118
          inner += (dmu2 - 1.0) * (a.gamma[j * K + k] * a.w[j]);                          // :120
        }
        g -= a.w[k] * inner / a.gsum[k];                                                  // :119-121
      }
    }
    if (!a.jacobian) g /= a.lambda[d];                                                  // :143-145
    a.out[o_lam + d] = g;
    return;
  }
  r -= D;
  if (r < K) {                                                                          // w_grad(j) before J_w
    if (!a.gf[3]) return;
    const int j = r;
    double g = 0.0;
    if (K > 1) {
      double s = 0.0;
      for (int k = 0; k < K; ++k) s += a.w[k] * a.gamma[j * K + k] / a.gsum[k];
      g = -log(a.gsum[j]) - s;                                                          // :126
    }
    a.wraw[j] = g;
    if (!a.jacobian) a.out[o_w + j] = g;
  }
}

// ---- stage D: idx in [0, K): softmax Jacobian of the w block (entlb_vbmc.m:147-151) ----
VB_HD void entlb_jw_elem(const EntlbArgs& a, int i) {
  if (!a.gf[3] || !a.jacobian) return;
  const int D = a.D, K = a.K;
  const int o_w = 1 + (a.gf[0] ? D * K : 0) + (a.gf[1] ? K : 0) + (a.gf[2] ? D : 0);
  double es = 0.0, dot = 0.0;
  for (int l = 0; l < K; ++l) es += exp(a.eta[l]);
  for (int l = 0; l < K; ++l) dot += exp(a.eta[l]) * a.wraw[l];
  const double ei = exp(a.eta[i]);
  a.out[o_w + i] = ei / es * a.wraw[i] - ei / (es * es) * dot;   // J_w = diag(e/es) - e e'/es^2
}

VB_HD int entlb_stage_size(const EntlbArgs& a, int stage) {
  switch (stage) {
    case 0: return a.K > 1 ? a.K * a.K : 0;
    case 1: return a.K > 1 ? a.K : 0;
    case 2: return 1 + a.D * a.K + a.K + a.D + a.K;
    default: return a.K;
  }
}
VB_HD void entlb_stage_elem(const EntlbArgs& a, int stage, int idx) {
  switch (stage) {
    case 0: entlb_gamma_elem(a, idx); break;
    case 1: entlb_gsum_elem(a, idx); break;
    case 2: entlb_grad_elem(a, idx); break;
    default: entlb_jw_elem(a, idx); break;
  }
}

}  // namespace vb
