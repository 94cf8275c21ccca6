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
#include "common.cuh"

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

// R.I[s][k] (all S rows: zero outside the local shard), R.Gmu[k][d], R.Gsig[k], R.Glam[d]
__global__ void __launch_bounds__(256) glj_reduce_kernel(const double* __restrict__ out, int ostride, int D, int K, int S,
                                                         int s_begin, int s_count, const double* __restrict__ w,
                                                         double* __restrict__ RI, double* __restrict__ Gmu,
                                                         double* __restrict__ Gsig, double* __restrict__ Glam) {
  extern __shared__ double tmp[];  // [K*D]  w_k * sum_s glam[s][k][d]
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < S * K; i += nt) {
    const int s = i / K;
    RI[i] = (s >= s_begin && s < s_begin + s_count) ? out[static_cast<size_t>(i) * ostride] : 0.0;
  }
  for (int i = tid; i < K * D; i += nt) {
    const int k = i / D, d = i - k * D;
    double gm = 0.0, gl = 0.0;
#pragma unroll 4
    for (int s = s_begin; s < s_begin + s_count; ++s) {
      const double* o = out + (static_cast<size_t>(s) * K + k) * ostride;
      gm += o[2 + d];
      gl += o[2 + D + d];
    }
    Gmu[i] = gm;
    tmp[i] = w[k] * gl;
  }
  for (int k = tid; k < K; k += nt) {
    double acc = 0.0;
#pragma unroll 4
    for (int s = s_begin; s < s_begin + s_count; ++s) acc += out[(static_cast<size_t>(s) * K + k) * ostride + 1];
    Gsig[k] = acc;
  }
  __syncthreads();
  for (int d = tid; d < D; d += nt) {
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc += tmp[k * D + d];
    Glam[d] = acc;
  }
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
  glj_reduce_kernel<<<1, 256, sizeof(double) * c->K * c->D, st>>>(c->glj_out.d(), 2 + 2 * c->D, c->D, c->K, c->gp.S, sb, se - sb, c->vp.w,
                                       R + rl.oI, R + rl.oGmu, R + rl.oGsig, R + rl.oGlam);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

}  // namespace vb
