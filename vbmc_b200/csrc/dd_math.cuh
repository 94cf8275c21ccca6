// Two-word ("double + error word") arithmetic for the expected log-joint contraction (gplogjoint.cu).
//
// Why: the sums  A = sum_n zeta_n,  B_d = sum_n zeta_n Delta_dn,  C_d = sum_n zeta_n (Delta_dn^2 - 1)  of
// misc/gplogjoint.m:164-252 weight the kernel column z_n by the GP weights alpha_n, whose magnitude is ~1e4 for
// VBMC's default noise floor (sn2 = 1e-5) while the results are O(1): 5-8 digits cancel, and a plain FP64 evaluation
// -- the reference's own included -- is only good to 1e-11..5e-9 relative to the exact value
// of the formula on the same double inputs (tests/test_truth128.py measures it against an IEEE binary128 evaluation).
// What decides the error is NOT the summation order (an exactly rounded sum of the FP64 terms is no better) but the
// n-dependent rounding of every term: mu - x, the scaling by 1/tau, the squares, the exponent and the exponential.
// Quantities that are common to all n of a sum (1/tau_d, lnnf) act as a relative perturbation of an input and stay
// plain doubles.  Everything that depends on n is carried as an unevaluated sum (h, l) of two doubles with error-free
// transformations (Knuth TwoSum, Dekker/FMA TwoProduct); low words are combined with plain operations (no
// renormalisation), which keeps ~100 significant bits at ~5x the FP64 instruction count.
//
// Host + device: tests/host_harness/dd_math_host.cpp compiles this header with g++ (-ffp-contract=off) and checks
// exp_dd and the term sums against mpmath / binary128 on the CPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "dd_exp2_table.inc"

#ifdef __CUDACC__
#define VB_HD __host__ __device__ __forceinline__
#else
#define VB_HD inline
#endif

namespace vb {

// s + e == a + b exactly (6 flops, no magnitude assumption)
VB_HD void two_sum(double a, double b, double& s, double& e) {
#ifdef __CUDA_ARCH__
  s = __dadd_rn(a, b);
  const double bb = __dsub_rn(s, a);
  e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
#else
  s = a + b;
  const double bb = s - a;
  e = (a - (s - bb)) + (b - bb);
#endif
}
// s + e == a + b exactly, requires |a| >= |b| (or a == 0)
VB_HD void fast_two_sum(double a, double b, double& s, double& e) {
#ifdef __CUDA_ARCH__
  s = __dadd_rn(a, b);
  e = __dsub_rn(b, __dsub_rn(s, a));
#else
  s = a + b;
  e = b - (s - a);
#endif
}
// p + e == a * b exactly
VB_HD void two_prod(double a, double b, double& p, double& e) {
#ifdef __CUDA_ARCH__
  p = __dmul_rn(a, b);
#else
  p = a * b;
#endif
  e = fma(a, b, -p);
}
// (h, l) += (ph, pl): high words by TwoSum, low words plainly
VB_HD void acc_add(double& h, double& l, double ph, double pl) {
  double s, e;
  two_sum(h, ph, s, e);
  h = s;
  l += e + pl;
}
// (h, l) + (ph, pl) renormalised (used in the reductions)
VB_HD void dd_add(double ah, double al, double bh, double bl, double& h, double& l) {
  double s, e;
  two_sum(ah, bh, s, e);
  e += al + bl;
  fast_two_sum(s, e, h, l);
}

#ifndef __CUDACC__
struct double2 {
  double x, y;
};
#endif

// exp(ah + al): returns (eh, el), eh = the nearest double, eh + el = exp(ah + al) to ~1e-23 relative, for -700 <= ah <= 700;
// below -700 the result is 0 (such a term is < 1e-304: nothing of it survives in any sum it enters).
// tab: 64 x (hi, lo) of 2^(j/64) (shared memory on the device).
VB_HD void exp_dd(double ah, double al, const double2* tab, double& eh, double& el) {
  if (!(ah >= -700.0)) {  // also catches NaN
    eh = ah != ah ? ah : 0.0;
    el = 0.0;
    return;
  }
  if (ah > 700.0) ah = 700.0;
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: round-to-nearest-integer by addition
  double kd = fma(ah, VB_64_LN2, MAGIC);
  int64_t kbits;
#ifdef __CUDA_ARCH__
  kbits = __double_as_longlong(kd);
  kd = __dsub_rn(kd, MAGIC);
#else
  memcpy(&kbits, &kd, 8);
  kd = kd - MAGIC;
#endif
  const int ki = static_cast<int>(static_cast<int32_t>(kbits));  // low 32 bits of the mantissa hold the integer (two's complement)
  const int j = ki & 63, m = ki >> 6;
  // r = a - k ln2/64 in three pieces: the first two products are exact
  const double r0 = fma(-kd, VB_LN2_64_C1, ah);  // exact (Cody-Waite)
  double p2, dummy;
  two_prod(kd, VB_LN2_64_C2, p2, dummy);          // exact: 21 + 32 bits
  double rh, re;
  two_sum(r0, -p2, rh, re);
  const double rl = fma(-kd, VB_LN2_64_C3, re) + al;  // |rl| <~ 1e-16 |a|: enters as the factor (1 + rl) at the end
  // exp(r) - 1 = r + r^2/2 + r^3 (1/6 + r/24 + ...): |r| <= ln2/128 = 0.0054, the cubic tail (<= 2.7e-8) in plain FP64
  double h2, h2e;
  two_prod(rh, rh, h2, h2e);
  double t = 1.0 / 40320.0;
  t = fma(t, rh, 1.0 / 5040.0);
  t = fma(t, rh, 1.0 / 720.0);
  t = fma(t, rh, 1.0 / 120.0);
  t = fma(t, rh, 1.0 / 24.0);
  t = fma(t, rh, 1.0 / 6.0);
  t = t * (h2 * rh);
  double qh, qe;
  fast_two_sum(rh, 0.5 * h2, qh, qe);
  const double ql = qe + (0.5 * h2e + t);
  // 2^(j/64) * (1 + q)
  const double Th = tab[j].x, Tl = tab[j].y;
  double mh, me;
  two_prod(Th, qh, mh, me);
  double Eh, Ee;
  fast_two_sum(Th, mh, Eh, Ee);
  const double El = Ee + (me + (Tl + fma(Th, ql, Tl * qh)));
  double Nh, Nl;
  fast_two_sum(Eh, El, Nh, Nl);  // the cubic tail (<= 2.7e-8 Eh) sits in El: renormalise
  Nl = fma(Nh, rl, Nl);          // * (1 + rl)
  // 2^m, m in [-1011, 1010]
  const int64_t sb = static_cast<int64_t>(m + 1023) << 52;
  double sc;
#ifdef __CUDA_ARCH__
  sc = __longlong_as_double(sb);
#else
  memcpy(&sc, &sb, 8);
#endif
  eh = Nh * sc;
  el = Nl * sc;
}

// One training point of one (s,k) pair, first half: Delta_d = (mu_d - x_d)/tau_d as (dh, dl) for the DP dimensions this
// thread owns, and their contribution sum_d Delta_d^2 as (ssh, ssl).
//   mu[d], itau[d]: plain doubles (common to all n);  x[d]: this point's coordinates.
template <int DP>
VB_HD void glj_delta(const double* mu, const double* itau, const double* x, double* dh, double* dl, double& ssh, double& ssl) {
  ssh = 0.0;
  ssl = 0.0;
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    double th, tl;
    two_sum(mu[d], -x[d], th, tl);  // mu - x exactly
    double h, l;
    two_prod(th, itau[d], h, l);
    l = fma(tl, itau[d], l);
    dh[d] = h;
    dl[d] = l;
    double p, e;
    two_prod(h, h, p, e);
    e = fma(h + h, l, e);
    acc_add(ssh, ssl, p, e);
  }
}
// second half: zeta = exp(lnnf - 0.5*(ssh + ssl)) * alpha as (zh, zl)
VB_HD void glj_zeta(double ssh, double ssl, double lnnf, double alpha, const double2* tab, double& zh, double& zl) {
  double ahh, ae;
  two_sum(lnnf, -0.5 * ssh, ahh, ae);
  const double al = fma(-0.5, ssl, ae);
  double eh, el;
  exp_dd(ahh, al, tab, eh, el);
  two_prod(eh, alpha, zh, zl);
  zl = fma(el, alpha, zl);
}

// accumulate A += zeta, B_d += zeta*Delta_d, Q_d += zeta*Delta_d^2  (C_d = Q_d - A afterwards)
template <int DP>
VB_HD void glj_accumulate(const double* dh, const double* dl, double zh, double zl, double* Bh, double* Bl, double* Qh, double* Ql) {
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    const double dhd = dh[d], dld = dl[d];
    double ph, pl;
    two_prod(zh, dhd, ph, pl);
    pl = fma(zh, dld, fma(zl, dhd, pl));
    acc_add(Bh[d], Bl[d], ph, pl);
    double rh, rl;
    two_prod(ph, dhd, rh, rl);
    rl = fma(ph, dld, fma(pl, dhd, rl));
    acc_add(Qh[d], Ql[d], rh, rl);
  }
}

}  // namespace vb
