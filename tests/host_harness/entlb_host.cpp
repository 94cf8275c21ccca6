// Host build of the entlb element functions (vbmc_b200/csrc/entlb_math.cuh) — the same source the CUDA kernels in
// csrc/entlb.cu execute one element per thread — run as plain loops so that tests/test_entlb_host.py can check every
// element against the NumPy restatement without a GPU.  Test infrastructure only; never part of the product library.
#include <vector>

// libvbmc_b200.so (often loaded in the same test process) exports nvcc's host-side launch stubs under the same mangled names as
// these kernels; without a private namespace the dynamic linker would bind the calls below to those stubs.
#define vb vb_host_harness

#define VB_HD inline
#include "../../vbmc_b200/csrc/entlb_math.cuh"

extern "C" int entlb_host(int D, int K, const int* gf, int jacobian, const double* mu, const double* sigma, const double* lambda,
                          const double* w, const double* eta, double* out) {
  std::vector<double> gamma((size_t)K * K), gsum(K), wraw(K);
  vb::EntlbArgs a;
  a.D = D; a.K = K; a.jacobian = jacobian;
  for (int i = 0; i < 4; ++i) a.gf[i] = gf[i];
  a.mu = mu; a.sigma = sigma; a.lambda = lambda; a.w = w; a.eta = eta;
  a.gamma = gamma.data(); a.gsum = gsum.data(); a.wraw = wraw.data(); a.out = out;
  for (int stage = 0; stage < 4; ++stage) {
    const int n = vb::entlb_stage_size(a, stage);
    for (int idx = 0; idx < n; ++idx) vb::entlb_stage_elem(a, stage, idx);   // a kernel launch runs exactly these calls, one per thread
  }
  return vb::entlb_ngrad(a);
}
