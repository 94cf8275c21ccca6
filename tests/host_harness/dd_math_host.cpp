// Host build of vbmc_b200/csrc/dd_math.cuh (the two-word arithmetic of the expected-log-joint kernel) for
// tests/test_dd_math_host.py.  Compile with -ffp-contract=off: the error-free transformations need every
// product and sum rounded separately.
#include "../../vbmc_b200/csrc/dd_math.cuh"

static const vb::double2 kTab[64] = {VB_EXP2_TABLE_ROWS};

extern "C" void dd_exp_host(double ah, double al, double* eh, double* el) { vb::exp_dd(ah, al, kTab, *eh, *el); }

// sums of one (s,k) pair over n = 0..N-1: out = [Ah, Al, Bh[D], Bl[D], Qh[D], Ql[D]]; X is N x D column-major
extern "C" void dd_glj_sums_host(int N, int D, const double* mu, const double* itau, const double* X, const double* alpha,
                                 double lnnf, double* out) {
  constexpr int DP = 24;
  double m[DP] = {0}, it[DP] = {0}, x[DP] = {0}, dh[DP], dl[DP];
  double Ah = 0, Al = 0, Bh[DP] = {0}, Bl[DP] = {0}, Qh[DP] = {0}, Ql[DP] = {0};
  for (int d = 0; d < D; ++d) {
    m[d] = mu[d];
    it[d] = itau[d];
  }
  for (int n = 0; n < N; ++n) {
    for (int d = 0; d < D; ++d) x[d] = X[static_cast<size_t>(d) * N + n];
    double zh, zl, ssh, ssl;
    vb::glj_delta<DP>(m, it, x, dh, dl, ssh, ssl);
    vb::glj_zeta(ssh, ssl, lnnf, alpha[n], kTab, zh, zl);
    vb::acc_add(Ah, Al, zh, zl);
    vb::glj_accumulate<DP>(dh, dl, zh, zl, Bh, Bl, Qh, Ql);
  }
  out[0] = Ah;
  out[1] = Al;
  for (int d = 0; d < D; ++d) {
    out[2 + d] = Bh[d];
    out[2 + D + d] = Bl[d];
    out[2 + 2 * D + d] = Qh[d];
    out[2 + 3 * D + d] = Ql[d];
  }
}
