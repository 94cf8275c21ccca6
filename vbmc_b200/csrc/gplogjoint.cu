// gplogjoint on sm_100a: expected GP log-joint by Bayesian quadrature and its gradient
// (reference: misc/gplogjoint.m:98-271).
//
// For each hyper-parameter sample s and mixture component k the reference builds D x N
// temporaries; everything it needs collapses to 2D+1 reductions over the N training points
// (SURVEY.md Appendix B.1):
//     zeta_n = exp(lnnf_sk - 0.5*sum_d Delta_dn^2) * alpha_sn,  Delta_dn = (mu_kd - X_nd)/tau_skd
//     A = sum zeta_n,   B_d = sum zeta_n Delta_dn,   C_d = sum zeta_n (Delta_dn^2 - 1)
// One CTA per (s,k): X (N x D column-major => unit stride over n) and alpha_s are streamed once,
// coalesced; the CTA then turns (A,B,C) into I_sk and the un-weighted gradient pieces
// (gplogjoint.m:169-174, 206-210, 227-231, 248-252).  The whole working set (X, alpha) is < 1 MB
// and stays in L2; the kernel is latency/FP64-pipe bound, not HBM bound.
#include <stdlib.h>

#include "common.cuh"
#include "glj_multi.cuh"

namespace vb {

struct GljArgs {
  int N, D, K, S;
  int s_begin, s_count;  // this rank's shard of the hyper-parameter samples
  int meanfun;
  int ostride;  // 2 + 2*D doubles per (s,k): [I, gsig, gmu[D], glam[D]]
  GpDev gp;
  VpDev vp;
  double* out;  // [S][K][ostride] (only rows s_begin..s_begin+s_count-1 written)
  const double* wvec;  // optional [S][K][N] weight vectors replacing alpha_s (variance gradient: K^-1 z_k)
  int raw;             // 1: epilogue without the mean-function terms (derivative contractions only)
};

constexpr int GLJ_THREADS = 128;

template <int DP>
__global__ void __launch_bounds__(GLJ_THREADS) glj_kernel(const GljArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int D = a.D, N = a.N;
  const int k = blockIdx.x;
  const int s = a.s_begin + blockIdx.y;
  const int tid = threadIdx.x;
  double* s_mu = sm;            // [DP]
  double* s_itau = sm + DP;     // [DP]
  double* s_misc = sm + 2 * DP; // [4]: lnnf
  double* red = sm + 2 * DP + 4;  // [(1+2D)][GLJ_THREADS+1]

  const double sigk = a.vp.sigma[k];
  if (tid < DP) {
    double mu = 0.0, itau = 0.0;
    if (tid < D) {
      const double lam = a.vp.lambda[tid], ell = a.gp.ell[s * D + tid], dl = a.vp.delta[tid];
      const double tau = sqrt(sigk * sigk * lam * lam + ell * ell + dl * dl);  // gplogjoint.m:164
      mu = a.vp.mu[k * D + tid];
      itau = 1.0 / tau;
    }
    s_mu[tid] = mu;
    s_itau[tid] = itau;
  }
  __syncthreads();
  if (tid == 0) {
    double slt = 0.0;
    for (int d = 0; d < D; ++d) slt += log(1.0 / s_itau[d]);
    s_misc[0] = a.gp.lnc[s] - slt;  // lnnf_k = ln_sf2 + sum_lnell - sum(log(tau_k))  (:165)
  }
  __syncthreads();
  const double lnnf = s_misc[0];
  double mu[DP], itau[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    mu[d] = s_mu[d];
    itau[d] = s_itau[d];
  }
  double A = 0.0, B[DP], C[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) B[d] = C[d] = 0.0;
  const double* __restrict__ X = a.gp.X;
  const double* __restrict__ alpha = a.wvec ? a.wvec + (static_cast<size_t>(s) * a.K + k) * N : a.gp.alpha + static_cast<size_t>(s) * N;
  for (int n = tid; n < N; n += GLJ_THREADS) {
    double dl[DP];
    double ss = 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const double x = d < D ? __ldg(X + static_cast<size_t>(d) * N + n) : 0.0;
      dl[d] = (mu[d] - x) * itau[d];
      ss = fma(dl[d], dl[d], ss);
    }
    const double zeta = exp(lnnf - 0.5 * ss) * __ldg(alpha + n);  // z_k(n)*alpha(n)  (:167-169)
    A += zeta;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      B[d] = fma(zeta, dl[d], B[d]);
      C[d] = fma(zeta, fma(dl[d], dl[d], -1.0), C[d]);
    }
  }
  // ---- block reduction in fixed order ----
  constexpr int RS = GLJ_THREADS + 1;
  red[0 * RS + tid] = A;
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    if (d < D) {
      red[(1 + d) * RS + tid] = B[d];
      red[(1 + D + d) * RS + tid] = C[d];
    }
  }
  __syncthreads();
  const int nval = 1 + 2 * D;
  // two-stage ordered sum: 4 partial sums of 32 per value, then the 4 parts in order (deterministic)
  double* red2 = red + static_cast<size_t>(nval) * RS;  // [nval][4]
  for (int idx = tid; idx < nval * 4; idx += GLJ_THREADS) {
    const int i = idx >> 2, part = idx & 3;
    const double* rr = red + i * RS + part * (GLJ_THREADS / 4);
    double sacc = 0.0;
    for (int t = 0; t < GLJ_THREADS / 4; ++t) sacc += rr[t];
    red2[idx] = sacc;
  }
  __syncthreads();
  for (int i = tid; i < nval; i += GLJ_THREADS) red[i * RS] = (red2[4 * i] + red2[4 * i + 1]) + (red2[4 * i + 2] + red2[4 * i + 3]);
  __syncthreads();
  // ---- per-(s,k) epilogue ----
  double* o = a.out + (static_cast<size_t>(s) * a.K + k) * a.ostride;
  const bool quad = a.meanfun == 4 && !a.raw;
  if (tid < D) {
    const int d = tid;
    const double lam = a.vp.lambda[d];
    const double it = s_itau[d];
    const double Bd = red[(1 + d) * RS], Cd = red[(1 + D + d) * RS];
    double gmu = -Bd * it;                                  // w(k)*dz_dmu*alpha / w(k)   (:206-208)
    double glam = sigk * sigk * lam * (Cd * it * it);       // (:248-249) / w(k)
    if (quad) {
      const double io2 = a.gp.iom2[s * D + d], xm = a.gp.xm[s * D + d];
      gmu -= io2 * (s_mu[d] - xm);                          // (:210)
      glam -= sigk * sigk * lam * io2;                      // (:252)
    }
    o[2 + d] = gmu;
    o[2 + D + d] = glam;
  }
  if (tid == 32) {
    double I = red[0] + ((a.meanfun > 0 && !a.raw) ? a.gp.m0[s] : 0.0);  // I_k = z_k*alpha + m0   (:169)
    double gs = 0.0;
    for (int d = 0; d < D; ++d) {
      const double lam = a.vp.lambda[d], it = s_itau[d], dl = a.vp.delta[d];
      gs += (lam * it) * (lam * it) * red[(1 + D + d) * RS];  // sum (lambda/tau)^2 (Delta^2-1) z alpha (:227-229)
      if (quad) {
        const double io2 = a.gp.iom2[s * D + d], xm = a.gp.xm[s * D + d], m = s_mu[d];
        I -= 0.5 * io2 * (m * m + sigk * sigk * lam * lam - 2.0 * m * xm + xm * xm + dl * dl);  // nu_k (:172-174)
        gs -= io2 * lam * lam;                                                                 // (:231)
      }
    }
    o[0] = I;
    o[1] = sigk * gs;
  }
}

// R.I[s][k] (all S rows: zero outside the local shard), R.Gmu[k][d], R.Gsig[k], R.GlamK[k][d] = w_k * sum_s glam[s][k][d]
// (finalize adds the K rows of GlamK).  One thread per output, the <= S loads of a thread are independent.
__global__ void __launch_bounds__(128) glj_reduce_kernel(const double* __restrict__ out, int ostride, int D, int K, int S,
                                                         int s_begin, int s_count, const double* __restrict__ w,
                                                         double* __restrict__ RI, double* __restrict__ Gmu,
                                                         double* __restrict__ Gsig, double* __restrict__ GlamK) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nI = S * K, nM = K * D;
  if (i < nI) {
    const int s = i / K;
    RI[i] = (s >= s_begin && s < s_begin + s_count) ? out[static_cast<size_t>(i) * ostride] : 0.0;
    return;
  }
  int r = i - nI, off, k;
  double scale = 1.0;
  double* dst;
  if (r < nM) {
    k = r / D; off = 2 + (r - k * D); dst = Gmu + r;
  } else if (r < nM + K) {
    k = r - nM; off = 1; dst = Gsig + k;
  } else if (r < 2 * nM + K) {
    r -= nM + K;
    k = r / D; off = 2 + D + (r - k * D); dst = GlamK + r; scale = w[k];
  } else {
    return;
  }
  const double* o = out + (static_cast<size_t>(s_begin) * K + k) * ostride + off;
  const size_t step = static_cast<size_t>(K) * ostride;
  double acc = 0.0;
#pragma unroll 8
  for (int s = 0; s < s_count; ++s) acc += o[s * step];
  *dst = scale * acc;
}

static int pick_dp(int D) {
  static const int opts[] = {2, 4, 6, 8, 10, 12, 16, 20, 24};
  for (int o : opts)
    if (D <= o) return o;
  return -1;
}

template <int DP>
static int launch_glj(vbmc_b200_ctx* c, const GljArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(double) * (2 * DP + 4 + static_cast<size_t>(1 + 2 * a.D) * (GLJ_THREADS + 1 + 4));
  auto kern = glj_kernel<DP>;
  if (smem > 48 * 1024)
    VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(a.K, a.s_count);
  KernelScope ks(c, "gplogjoint", st);
  kern<<<grid, GLJ_THREADS, smem, st>>>(a);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}


// ------------------------------------------------------------------------------------------------
// Fast path of the step (alpha-weighted contraction, all samples of this rank): one THREAD per (s,k) pair, the CTA's
// threads sweep the same chunk of training points, so X[n][:] is a CTA-uniform (broadcast) load and the 2D+1 sums of
// a pair live in registers with no cross-thread reduction at all.  The N axis is split into chunks across CTAs
// (grid.x) to fill the machine; a second kernel adds the chunk partials in chunk order (deterministic) and applies
// the per-(s,k) epilogue.  Compared with one CTA per (s,k) this removes 1000 block reductions and the S*K-fold
// re-read of X from L2: 60-80 us -> ~10 us at c3, short enough to hide behind the draw generator.
// ------------------------------------------------------------------------------------------------
constexpr int GLJ2_THREADS = 128;

template <int DP>
__global__ void __launch_bounds__(GLJ2_THREADS) glj_pairs_kernel(const GljArgs a, int chunk, double* __restrict__ part) {
  const int D = a.D, N = a.N, K = a.K;
  const int npairs = a.s_count * K;
  const int i = blockIdx.y * GLJ2_THREADS + threadIdx.x;
  const bool live = i < npairs;
  const int sl = live ? i / K : 0, k = live ? i - sl * K : 0;
  const int s = a.s_begin + sl;
  const double sigk = a.vp.sigma[k];
  double mu[DP], itau[DP];
  double slt = 0.0;
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    mu[d] = 0.0;
    itau[d] = 0.0;
    if (d < D) {
      const double lam = a.vp.lambda[d], ell = a.gp.ell[s * D + d], dl = a.vp.delta[d];
      const double tau = sqrt(sigk * sigk * lam * lam + ell * ell + dl * dl);  // gplogjoint.m:164
      mu[d] = a.vp.mu[k * D + d];
      itau[d] = 1.0 / tau;
      slt += log(1.0 / itau[d]);
    }
  }
  const double lnnf = a.gp.lnc[s] - slt;  // lnnf_k = ln_sf2 + sum_lnell - sum(log(tau_k))  (:165)
  const int n0 = blockIdx.x * chunk;
  int n1 = n0 + chunk;
  n1 = n1 > N ? N : n1;
  const double* __restrict__ X = a.gp.X;
  const double* __restrict__ alpha = a.gp.alpha + static_cast<size_t>(s) * N;
  double A = 0.0, B[DP], C[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) B[d] = C[d] = 0.0;
  int n = n0;
  for (; n + 1 < n1; n += 2) {  // two points per iteration: two independent exp chains
    double d0[DP], d1[DP];
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const double x0 = d < D ? __ldg(X + static_cast<size_t>(d) * N + n) : 0.0;
      const double x1 = d < D ? __ldg(X + static_cast<size_t>(d) * N + n + 1) : 0.0;
      d0[d] = (mu[d] - x0) * itau[d];
      d1[d] = (mu[d] - x1) * itau[d];
      s0 = fma(d0[d], d0[d], s0);
      s1 = fma(d1[d], d1[d], s1);
    }
    const double z0 = exp(lnnf - 0.5 * s0) * __ldg(alpha + n);  // z_k(n)*alpha(n)  (:167-169)
    const double z1 = exp(lnnf - 0.5 * s1) * __ldg(alpha + n + 1);
    A += z0;
    A += z1;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      B[d] = fma(z0, d0[d], B[d]);
      C[d] = fma(z0, fma(d0[d], d0[d], -1.0), C[d]);
      B[d] = fma(z1, d1[d], B[d]);
      C[d] = fma(z1, fma(d1[d], d1[d], -1.0), C[d]);
    }
  }
  if (n < n1) {
    double d0[DP];
    double s0 = 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const double x0 = d < D ? __ldg(X + static_cast<size_t>(d) * N + n) : 0.0;
      d0[d] = (mu[d] - x0) * itau[d];
      s0 = fma(d0[d], d0[d], s0);
    }
    const double z0 = exp(lnnf - 0.5 * s0) * __ldg(alpha + n);
    A += z0;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      B[d] = fma(z0, d0[d], B[d]);
      C[d] = fma(z0, fma(d0[d], d0[d], -1.0), C[d]);
    }
  }
  if (!live) return;
  // partials: [chunk][value][pair] => consecutive threads write consecutive addresses
  double* o = part + static_cast<size_t>(blockIdx.x) * (1 + 2 * D) * npairs + i;
  o[0] = A;
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    if (d < D) {
      o[static_cast<size_t>(1 + d) * npairs] = B[d];
      o[static_cast<size_t>(1 + D + d) * npairs] = C[d];
    }
  }
}

// chunk partials -> totals, one thread per (value, pair): the nchunks loads of a thread are independent and coalesced
// across the warp; fixed chunk order => deterministic
__global__ void __launch_bounds__(128) glj_pairs_sum_kernel(int nchunks, int nval, int npairs, const double* __restrict__ part,
                                                            double* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (i >= npairs) return;
  const double* p = part + static_cast<size_t>(v) * npairs + i;
  const size_t stride = static_cast<size_t>(nval) * npairs;
  double acc = 0.0;
#pragma unroll 8
  for (int c = 0; c < nchunks; ++c) acc += p[c * stride];
  sums[static_cast<size_t>(v) * npairs + i] = acc;
}

// totals -> [I, gsig, gmu[D], glam[D]] per (s,k)   (gplogjoint.m:169-174, 206-210, 227-231, 248-252)
__global__ void __launch_bounds__(128) glj_pairs_epilogue_kernel(const GljArgs a, const double* __restrict__ sums) {
  const int D = a.D, K = a.K;
  const int npairs = a.s_count * K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npairs) return;
  const int sl = i / K, k = i - sl * K, s = a.s_begin + sl;
  const double sigk = a.vp.sigma[k];
  double* o = a.out + (static_cast<size_t>(s) * K + k) * a.ostride;
  const bool quad = a.meanfun == 4;
  double I = sums[i] + (a.meanfun > 0 ? a.gp.m0[s] : 0.0);  // I_k = z_k*alpha + m0   (:169)
  double gs = 0.0;
  for (int d = 0; d < D; ++d) {
    const double lam = a.vp.lambda[d], ell = a.gp.ell[s * D + d], dl = a.vp.delta[d];
    const double it = 1.0 / sqrt(sigk * sigk * lam * lam + ell * ell + dl * dl);
    const double Bd = sums[static_cast<size_t>(1 + d) * npairs + i], Cd = sums[static_cast<size_t>(1 + D + d) * npairs + i];
    const double m = a.vp.mu[k * D + d];
    double gmu = -Bd * it;                             // (:206-208) / w(k)
    double glam = sigk * sigk * lam * (Cd * it * it);  // (:248-249) / w(k)
    gs += (lam * it) * (lam * it) * Cd;                // (:227-229)
    if (quad) {
      const double io2 = a.gp.iom2[s * D + d], xm = a.gp.xm[s * D + d];
      gmu -= io2 * (m - xm);                           // (:210)
      glam -= sigk * sigk * lam * io2;                 // (:252)
      I -= 0.5 * io2 * (m * m + sigk * sigk * lam * lam - 2.0 * m * xm + xm * xm + dl * dl);  // nu_k (:172-174)
      gs -= io2 * lam * lam;                           // (:231)
    }
    o[2 + d] = gmu;
    o[2 + D + d] = glam;
  }
  o[0] = I;
  o[1] = sigk * gs;
}

template <int DP>
static int launch_glj_pairs(vbmc_b200_ctx* c, const GljArgs& a, cudaStream_t st) {
  const int npairs = a.s_count * a.K;
  const int by = (npairs + GLJ2_THREADS - 1) / GLJ2_THREADS;
  int nchunks = (4 * c->num_sms + by - 1) / by;          // ~4 CTAs (16 warps) per SM
  const int max_chunks = (a.N + 15) / 16;                // at least 16 points per chunk
  nchunks = nchunks > max_chunks ? max_chunks : (nchunks < 1 ? 1 : nchunks);
  const int chunk = (a.N + nchunks - 1) / nchunks;
  nchunks = (a.N + chunk - 1) / chunk;
  const size_t nval = 1 + 2 * a.D;
  VB_TRY(c->glj_part.reserve(sizeof(double) * (nchunks + 1) * nval * npairs));
  double* sums = c->glj_part.d() + static_cast<size_t>(nchunks) * nval * npairs;
  {
    KernelScope ks(c, "gplogjoint", st);
    glj_pairs_kernel<DP><<<dim3(nchunks, by), GLJ2_THREADS, 0, st>>>(a, chunk, c->glj_part.d());
    VB_CUDA(cudaGetLastError());
  }
  {
    KernelScope ks(c, "gplogjoint_epilogue", st);
    glj_pairs_sum_kernel<<<dim3((npairs + 127) / 128, static_cast<unsigned>(nval)), 128, 0, st>>>(nchunks, static_cast<int>(nval), npairs,
                                                                                                 c->glj_part.d(), sums);
    VB_CUDA(cudaGetLastError());
    c->launches++;
    glj_pairs_epilogue_kernel<<<(npairs + 127) / 128, 128, 0, st>>>(a, sums);
    VB_CUDA(cudaGetLastError());
  }
  return VBMC_B200_OK;
}

// Variant "multi" (glj_multi.cuh): several components per CTA with the points in registers; shares the chunk-sum and epilogue
// kernels of the thread-per-pair variant.  VBMC_B200_GLJ_VARIANT=multi, group size VBMC_B200_GLJ_KG (default 10).
template <int DP>
static int launch_glj_multi(vbmc_b200_ctx* c, const GljArgs& a, cudaStream_t st) {
  constexpr int P = 4;
  static const int kg_env = getenv("VBMC_B200_GLJ_KG") ? atoi(getenv("VBMC_B200_GLJ_KG")) : 10;
  const int kg = kg_env < 1 ? 1 : (kg_env > a.K ? a.K : kg_env);
  const int npairs = a.s_count * a.K;
  const int nchunks = (a.N + P * GLJM_THREADS - 1) / (P * GLJM_THREADS), kgroups = (a.K + kg - 1) / kg;
  const size_t nval = 1 + 2 * a.D;
  VB_TRY(c->glj_part.reserve(sizeof(double) * (nchunks + 1) * nval * npairs));
  double* sums = c->glj_part.d() + static_cast<size_t>(nchunks) * nval * npairs;
  GljMultiArgs m;
  m.N = a.N; m.D = a.D; m.K = a.K; m.s_begin = a.s_begin; m.npairs = npairs; m.kg = kg;
  m.X = a.gp.X; m.alpha = a.gp.alpha; m.ell = a.gp.ell; m.lnc = a.gp.lnc;
  m.mu = a.vp.mu; m.sigma = a.vp.sigma; m.lambda = a.vp.lambda; m.delta = a.vp.delta;
  m.part = c->glj_part.d();
  const size_t smem = sizeof(double) * glj_multi_smem_doubles(a.D, DP, kg);
  auto kern = glj_multi_kernel<DP, P>;
  if (smem > 48 * 1024)
    VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  {
    KernelScope ks(c, "gplogjoint", st);
    kern<<<dim3(nchunks, kgroups, a.s_count), GLJM_THREADS, smem, st>>>(m);
    VB_CUDA(cudaGetLastError());
  }
  {
    KernelScope ks(c, "gplogjoint_epilogue", st);
    glj_pairs_sum_kernel<<<dim3((npairs + 127) / 128, static_cast<unsigned>(nval)), 128, 0, st>>>(nchunks, static_cast<int>(nval), npairs,
                                                                                                 c->glj_part.d(), sums);
    VB_CUDA(cudaGetLastError());
    c->launches++;
    glj_pairs_epilogue_kernel<<<(npairs + 127) / 128, 128, 0, st>>>(a, sums);
    VB_CUDA(cudaGetLastError());
  }
  return VBMC_B200_OK;
}

int launch_gplogjoint(vbmc_b200_ctx* c, int need_grad, cudaStream_t st) {
  (void)need_grad;
  GljArgs a;
  a.N = c->gp.N; a.D = c->D; a.K = c->K; a.S = c->gp.S;
  shard_range(a.S, c->nranks, c->rank, &a.s_begin, &a.s_count);
  a.s_count -= a.s_begin;
  a.meanfun = c->gp.meanfun;
  a.ostride = 2 + 2 * a.D;
  a.gp = c->gp;
  a.vp = c->vp;
  a.wvec = nullptr;
  a.raw = 0;
  VB_TRY(c->glj_out.reserve(sizeof(double) * static_cast<size_t>(a.S) * a.K * a.ostride));
  a.out = c->glj_out.d();
  if (a.s_count <= 0) return VBMC_B200_OK;
  static const bool multi = getenv("VBMC_B200_GLJ_VARIANT") && !strcmp(getenv("VBMC_B200_GLJ_VARIANT"), "multi");
  if (multi) {
    switch (pick_dp(a.D)) {
      case 2: return launch_glj_multi<2>(c, a, st);
      case 4: return launch_glj_multi<4>(c, a, st);
      case 6: return launch_glj_multi<6>(c, a, st);
      case 8: return launch_glj_multi<8>(c, a, st);
      case 10: return launch_glj_multi<10>(c, a, st);
      case 12: return launch_glj_multi<12>(c, a, st);
      case 16: return launch_glj_multi<16>(c, a, st);
      case 20: return launch_glj_multi<20>(c, a, st);
      case 24: return launch_glj_multi<24>(c, a, st);
    }
  }
  if (!getenv("VBMC_B200_GLJ_THREAD_PER_PAIR")) {  // default: one CTA per (s,k); the thread-per-pair variant measured slower end to end
    switch (pick_dp(a.D)) {
      case 2: return launch_glj<2>(c, a, st);
      case 4: return launch_glj<4>(c, a, st);
      case 6: return launch_glj<6>(c, a, st);
      case 8: return launch_glj<8>(c, a, st);
      case 10: return launch_glj<10>(c, a, st);
      case 12: return launch_glj<12>(c, a, st);
      case 16: return launch_glj<16>(c, a, st);
      case 20: return launch_glj<20>(c, a, st);
      case 24: return launch_glj<24>(c, a, st);
    }
  }
  switch (pick_dp(a.D)) {
    case 2: return launch_glj_pairs<2>(c, a, st);
    case 4: return launch_glj_pairs<4>(c, a, st);
    case 6: return launch_glj_pairs<6>(c, a, st);
    case 8: return launch_glj_pairs<8>(c, a, st);
    case 10: return launch_glj_pairs<10>(c, a, st);
    case 12: return launch_glj_pairs<12>(c, a, st);
    case 16: return launch_glj_pairs<16>(c, a, st);
    case 20: return launch_glj_pairs<20>(c, a, st);
    case 24: return launch_glj_pairs<24>(c, a, st);
  }
  VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:gplogjoint: D=%d > 24 is not supported by this build", a.D);
}

// same contraction with per-(s,k) weight vectors w_sk (N each) and the raw epilogue; all S samples, out [S][K][2+2D]
int launch_gplogjoint_weighted(vbmc_b200_ctx* c, const double* wvec, double* out, cudaStream_t st) {
  GljArgs a;
  a.N = c->gp.N; a.D = c->D; a.K = c->K; a.S = c->gp.S;
  a.s_begin = 0; a.s_count = a.S;
  a.meanfun = c->gp.meanfun;
  a.ostride = 2 + 2 * a.D;
  a.gp = c->gp;
  a.vp = c->vp;
  a.wvec = wvec;
  a.raw = 1;
  a.out = out;
  switch (pick_dp(a.D)) {
    case 2: return launch_glj<2>(c, a, st);
    case 4: return launch_glj<4>(c, a, st);
    case 6: return launch_glj<6>(c, a, st);
    case 8: return launch_glj<8>(c, a, st);
    case 10: return launch_glj<10>(c, a, st);
    case 12: return launch_glj<12>(c, a, st);
    case 16: return launch_glj<16>(c, a, st);
    case 20: return launch_glj<20>(c, a, st);
    case 24: return launch_glj<24>(c, a, st);
  }
  VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:gplogjoint: D=%d > 24 is not supported by this build", a.D);
}

int launch_glj_reduce(vbmc_b200_ctx* c, cudaStream_t st) {
  RLayout rl;
  rl.init(c->D, c->K, c->gp.S);
  int sb, se;
  shard_range(c->gp.S, c->nranks, c->rank, &sb, &se);
  double* R = c->R_dev.d();
  KernelScope ks(c, "reduce", st);
  const int nout = c->gp.S * c->K + 2 * c->K * c->D + c->K;
  glj_reduce_kernel<<<(nout + 127) / 128, 128, 0, st>>>(c->glj_out.d(), 2 + 2 * c->D, c->D, c->K, c->gp.S, sb, se - sb, c->vp.w,
                                                       R + rl.oI, R + rl.oGmu, R + rl.oGsig, R + rl.oGlam);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

}  // namespace vb
