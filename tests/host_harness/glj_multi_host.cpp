// Host run of vbmc_b200/csrc/glj_multi.cuh through the thread-per-CUDA-thread shim (see cuda_shim.h).
#include "cuda_shim.h"

// libvbmc_b200.so (often loaded in the same test process) exports nvcc's host-side launch stubs under the same mangled names as
// these kernels; without a private namespace the dynamic linker would bind the calls below to those stubs.
#define vb vb_host_harness
#include "../../vbmc_b200/csrc/glj_multi.cuh"

template <int DP, int P>
static void run(const vb::GljMultiArgs& a, int nchunks, int kgroups, int s_count) {
  const size_t smem = sizeof(double) * vb::glj_multi_smem_doubles(a.D, DP, a.kg);
  vbshim::launch(vb::glj_multi_kernel<DP, P>, vbshim::Dim3{static_cast<unsigned>(nchunks), static_cast<unsigned>(kgroups),
                                                           static_cast<unsigned>(s_count)},
                 vb::GLJM_THREADS, smem, a);
}

// part must hold nchunks * (1 + 2D) * s_count * K doubles; returns nchunks (or -1 for an unsupported DP / P)
extern "C" int glj_multi_host(int N, int D, int K, int s_begin, int s_count, int kg, int DP, int P, const double* X, const double* alpha,
                              const double* ell, const double* lnc, const double* mu, const double* sigma, const double* lambda,
                              const double* delta, double* part) {
  vb::GljMultiArgs a;
  a.N = N; a.D = D; a.K = K; a.s_begin = s_begin; a.npairs = s_count * K; a.kg = kg;
  a.X = X; a.alpha = alpha; a.ell = ell; a.lnc = lnc; a.mu = mu; a.sigma = sigma; a.lambda = lambda; a.delta = delta; a.part = part;
  const int nchunks = (N + P * vb::GLJM_THREADS - 1) / (P * vb::GLJM_THREADS), kgroups = (K + kg - 1) / kg;
  if (DP == 2 && P == 4) run<2, 4>(a, nchunks, kgroups, s_count);
  else if (DP == 4 && P == 4) run<4, 4>(a, nchunks, kgroups, s_count);
  else if (DP == 10 && P == 4) run<10, 4>(a, nchunks, kgroups, s_count);
  else if (DP == 4 && P == 2) run<4, 2>(a, nchunks, kgroups, s_count);
  else return -1;
  return nchunks;
}
