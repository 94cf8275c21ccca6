// Expected log-joint contraction, variant "multi": one CTA = one hyper-parameter sample s, one chunk of P*128 training
// points and a GROUP of components.  A thread keeps its P points (D coordinates each) in registers and sweeps the
// group's components, so X is read from L2 once per (s, chunk, group) instead of once per (s, k): S*K = 1000 re-reads of
// X per step at c3 (160 MB of L2 traffic) become S*K/kg.  The 2D+1 sums of a component are reduced across the warp
// through a per-warp transposed scratch in shared memory (fixed order => deterministic), the four warps are added in
// order, and the chunk partials go to the same [chunk][value][pair] array the thread-per-pair variant uses, so the chunk
// sum and the per-(s,k) epilogue kernels of gplogjoint.cu are shared (misc/gplogjoint.m:164-252).
//
// Selected with VBMC_B200_GLJ_VARIANT=multi (default: one CTA per (s,k)); written at the end of round 1 without GPU time
// left, so it is OFF by default.  The very same source is run on the CPU by tests/test_glj_multi_host.py through a
// thread-per-CUDA-thread shim (tests/host_harness/cuda_shim.h: std::barrier for __syncthreads/__syncwarp), which checks
// its indexing, barriers and sums against a plain triple loop.
#pragma once
#include <math.h>

namespace vb {

struct GljMultiArgs {
  int N, D, K;
  int s_begin;         // first hyper-parameter sample of this rank; blockIdx.z counts from it
  int npairs;          // s_count * K
  int kg;              // components per CTA
  const double* X;      // [D][N]
  const double* alpha;  // [S][N]
  const double* ell;    // [S][D]
  const double* lnc;    // [S]
  const double* mu;     // [K][D]
  const double* sigma;  // [K]
  const double* lambda; // [D]
  const double* delta;  // [D]
  double* part;         // [nchunks][1 + 2D][npairs]
};

constexpr int GLJM_THREADS = 128;

#ifdef __CUDACC__
#define VB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define VB_LDG(p) __ldg(p)
#else
#define VB_DYN_SMEM(name) unsigned char* name = vbshim::dynamic_smem()
#define VB_LDG(p) (*(p))
#endif

// dynamic shared memory in doubles: mu[kg][DP] | itau[kg][DP] | lnnf[kg] (padded to even) | scratch[4][V][33] | wsum[4][kg][V]
inline size_t glj_multi_smem_doubles(int D, int DP, int kg) {
  const int V = 1 + 2 * D;
  return static_cast<size_t>(2) * kg * DP + ((kg + 1) & ~1) + static_cast<size_t>(4) * V * 33 + static_cast<size_t>(4) * kg * V;
}

template <int DP, int P>
__global__ void __launch_bounds__(GLJM_THREADS) glj_multi_kernel(const GljMultiArgs a) {
  VB_DYN_SMEM(smem_raw);
  const int D = a.D, N = a.N, K = a.K, V = 1 + 2 * D, kg = a.kg;
  double* s_mu = reinterpret_cast<double*>(smem_raw);   // [kg][DP]
  double* s_itau = s_mu + kg * DP;                       // [kg][DP]
  double* s_lnnf = s_itau + kg * DP;                     // [kg]
  double* s_scr = s_lnnf + ((kg + 1) & ~1);              // [4][V][33]
  double* s_wsum = s_scr + 4 * V * 33;                   // [4][kg][V]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.x, k0 = blockIdx.y * kg, sl = blockIdx.z, s = a.s_begin + sl;
  const int nk = (K - k0) < kg ? (K - k0) : kg;

  // ---- per-component constants of this group (gplogjoint.m:164-165) ----
  for (int i = tid; i < nk * DP; i += GLJM_THREADS) {
    const int kk = i / DP, d = i - kk * DP, k = k0 + kk;
    double m = 0.0, it = 0.0;
    if (d < D) {
      const double sg = a.sigma[k], lam = a.lambda[d], el = a.ell[s * D + d], dl = a.delta[d];
      it = 1.0 / sqrt(sg * sg * lam * lam + el * el + dl * dl);
      m = a.mu[k * D + d];
    }
    s_mu[i] = m;
    s_itau[i] = it;
  }
  __syncthreads();
  for (int kk = tid; kk < nk; kk += GLJM_THREADS) {
    double slt = 0.0;
    for (int d = 0; d < D; ++d) slt += log(1.0 / s_itau[kk * DP + d]);
    s_lnnf[kk] = a.lnc[s] - slt;
  }
  __syncthreads();

  // ---- this thread's P points ----
  double x[P][DP], al[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int n = (chunk * P + p) * GLJM_THREADS + tid;
    const bool valid = n < N;
    al[p] = valid ? VB_LDG(a.alpha + static_cast<size_t>(s) * N + n) : 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) x[p][d] = (valid && d < D) ? VB_LDG(a.X + static_cast<size_t>(d) * N + n) : 0.0;
  }

  double* scr = s_scr + warp * V * 33;
#pragma unroll 1
  for (int kk = 0; kk < nk; ++kk) {
    double mu[DP], itau[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      mu[d] = s_mu[kk * DP + d];
      itau[d] = s_itau[kk * DP + d];
    }
    const double lnnf = s_lnnf[kk];
    double A = 0.0, B[DP], C[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) B[d] = C[d] = 0.0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      double dl[DP], ss = 0.0;
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        dl[d] = (mu[d] - x[p][d]) * itau[d];
        ss = fma(dl[d], dl[d], ss);
      }
      const double zeta = exp(lnnf - 0.5 * ss) * al[p];   // z_k(n)*alpha(n)  (:167-169)
      A += zeta;
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        B[d] = fma(zeta, dl[d], B[d]);
        C[d] = fma(zeta, fma(dl[d], dl[d], -1.0), C[d]);
      }
    }
    // warp reduction through the transposed scratch: lane v adds row v in lane order
    scr[0 * 33 + lane] = A;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      if (d < D) {
        scr[(1 + d) * 33 + lane] = B[d];
        scr[(1 + D + d) * 33 + lane] = C[d];
      }
    }
    __syncwarp();
    for (int v = lane; v < V; v += 32) {
      const double* row = scr + v * 33;
      double t = 0.0;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) t += row[l];
      s_wsum[(warp * kg + kk) * V + v] = t;
    }
    __syncwarp();
  }
  __syncthreads();
  // ---- the four warps in order -> chunk partial [chunk][value][pair] ----
  for (int i = tid; i < nk * V; i += GLJM_THREADS) {
    const int kk = i / V, v = i - kk * V;
    const double t = ((s_wsum[(0 * kg + kk) * V + v] + s_wsum[(1 * kg + kk) * V + v]) + s_wsum[(2 * kg + kk) * V + v]) +
                     s_wsum[(3 * kg + kk) * V + v];
    const int pair = sl * K + k0 + kk;
    a.part[(static_cast<size_t>(chunk) * V + v) * a.npairs + pair] = t;
  }
}

}  // namespace vb
