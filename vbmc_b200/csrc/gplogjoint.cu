// gplogjoint on sm_100a: expected GP log-joint by Bayesian quadrature and its gradient
// (reference: misc/gplogjoint.m:98-271).
//
// For each hyper-parameter sample s and mixture component k the reference builds D x N
// temporaries; everything it needs collapses to 2D+1 reductions over the N training points
// (SURVEY.md Appendix B.1):
//     zeta_n = exp(lnnf_sk - 0.5*sum_d Delta_dn^2) * alpha_sn,  Delta_dn = (mu_kd - X_nd)/tau_skd
//     A = sum zeta_n,   B_d = sum zeta_n Delta_dn,   C_d = sum zeta_n (Delta_dn^2 - 1)
// A CTA owns one (s, k, slice of N): X (N x D column-major => unit stride over n) and alpha_s are streamed
// coalesced; the working set (X, alpha) is < 1 MB and stays in L2, the kernel is FP64-pipe bound.
//
// Accuracy (dd_math.cuh): alpha_n ~ 1e4 while the sums are O(1), so every n-dependent quantity is carried as a
// two-word (h, l) value -- the terms, the three running sums, the block and slice reductions -- and the results
// are the correctly rounded sums of the exact formula to ~1e-20 * sum|zeta_n|.  tests/test_truth128.py checks the
// outputs against an IEEE binary128 evaluation on ill-conditioned posteriors (plain FP64 is off by 1e-11..5e-9 there).
//
// The N axis is split over gridDim.z slices when S_local*K CTAs would not fill the GPU (multi-GPU shards of S);
// slice partials go to global memory and the LAST slice to finish (atomic ticket) adds them in slice order and
// runs the per-(s,k) epilogue (gplogjoint.m:169-174, 206-210, 227-231, 248-252) -- no second launch, deterministic.
#include <stdlib.h>

#include "common.cuh"
#include "dd_math.cuh"

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
  int nsplit;          // slices of the N axis (gridDim.z)
  double* part;        // [S*K][nsplit][2*(1+2D)] slice partials (nsplit > 1)
  unsigned* ticket;    // [S*K] arrival counters, zero between launches
};

constexpr int GLJ_THREADS = 128;
constexpr int GLJ_PARTS = 8;  // block reduction: 8 ordered partial sums of 16 threads each, then the 8 in order

__device__ const double2 g_exp2_tab[64] = {VB_EXP2_TABLE_ROWS};

// dynamic shared memory (doubles): exp2 table [128] | mu [DP] | itau [DP] | misc [4] | red [2*nval][COLS+1] | red2 [2*nval][GLJ_PARTS]
static size_t glj_smem_bytes(int D, int DP, int halves) {
  const size_t nval2 = 2 * (1 + 2 * static_cast<size_t>(D));
  return sizeof(double) * (128 + 2 * DP + 4 + nval2 * (GLJ_THREADS / halves + 1) + nval2 * GLJ_PARTS);
}

// DPH dimensions per thread, HALVES threads per training point (HALVES = 2 for D > 12: the 4*D + 2 accumulator words of one
// thread would not fit the register file; the lane pair exchanges its two partial sum_d Delta^2 with one shuffle and both
// lanes evaluate the exponential).
template <int DPH, int HALVES>
__global__ void __launch_bounds__(GLJ_THREADS) glj_kernel(const GljArgs a) {
  extern __shared__ __align__(16) double sm[];
  constexpr int DP = DPH * HALVES;
  constexpr int COLS = GLJ_THREADS / HALVES;  // training points per sweep iteration
  constexpr int RS = COLS + 1;
  const int D = a.D, N = a.N;
  const int k = blockIdx.x;
  const int s = a.s_begin + blockIdx.y;
  const int tid = threadIdx.x;
  const int half = HALVES == 2 ? (tid & 1) : 0, col = HALVES == 2 ? (tid >> 1) : tid;
  const int d0 = half * DPH;
  double2* s_tab = reinterpret_cast<double2*>(sm);  // [64]
  double* s_mu = sm + 128;          // [DP]
  double* s_itau = s_mu + DP;       // [DP]
  double* s_misc = s_itau + DP;     // [4]: lnnf, last-slice flag
  double* red = s_misc + 4;         // [2*nval][RS]
  const int nval = 1 + 2 * D, nval2 = 2 * nval;
  double* red2 = red + static_cast<size_t>(nval2) * RS;  // [2*nval][GLJ_PARTS]

  if (tid < 64) s_tab[tid] = g_exp2_tab[tid];
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
  double Ah = 0.0, Al = 0.0, Bh[DPH], Bl[DPH], Qh[DPH], Ql[DPH];
#pragma unroll
  for (int d = 0; d < DPH; ++d) Bh[d] = Bl[d] = Qh[d] = Ql[d] = 0.0;
  const double* __restrict__ X = a.gp.X;
  const double* __restrict__ alpha = a.wvec ? a.wvec + (static_cast<size_t>(s) * a.K + k) * N : a.gp.alpha + static_cast<size_t>(s) * N;
  // slice of the N axis owned by this CTA (multiples of the block size except the last)
  const int per = ((N + a.nsplit - 1) / a.nsplit + GLJ_THREADS - 1) / GLJ_THREADS * GLJ_THREADS;
  const int n0 = blockIdx.z * per;
  const int n1 = min(N, n0 + per);
  for (int base = n0; base < n1; base += COLS) {  // warp-uniform trip count: the lane pairs shuffle inside
    const int n = base + col;
    const bool live = n < n1;
    double x[DPH], dh[DPH], dl[DPH];
#pragma unroll
    for (int d = 0; d < DPH; ++d) x[d] = (live && d0 + d < D) ? __ldg(X + static_cast<size_t>(d0 + d) * N + n) : 0.0;
    const double al_n = live ? __ldg(alpha + n) : 0.0;
    double ssh, ssl;
    glj_delta<DPH>(s_mu + d0, s_itau + d0, x, dh, dl, ssh, ssl);
    if (HALVES == 2) {
      const double oh = __shfl_xor_sync(0xffffffffu, ssh, 1), ol = __shfl_xor_sync(0xffffffffu, ssl, 1);
      double th, tl;
      two_sum(ssh, oh, th, tl);  // symmetric: both lanes of the pair get the same (ssh, ssl)
      ssl = tl + (ssl + ol);
      ssh = th;
    }
    double zh, zl;
    glj_zeta(ssh, ssl, lnnf, al_n, s_tab, zh, zl);  // z_k(n)*alpha(n)  (:167-169)
    if (half == 0) acc_add(Ah, Al, zh, zl);
    glj_accumulate<DPH>(dh, dl, zh, zl, Bh, Bl, Qh, Ql);
  }
  // ---- block reduction in fixed order, two-word: value v of column `col` -> red[2v][col], red[2v+1][col] ----
  if (half == 0) {
    red[0 * RS + col] = Ah;
    red[1 * RS + col] = Al;
  }
#pragma unroll
  for (int d = 0; d < DPH; ++d) {
    if (d0 + d < D) {
      red[(2 + 2 * (d0 + d)) * RS + col] = Bh[d];
      red[(3 + 2 * (d0 + d)) * RS + col] = Bl[d];
      red[(2 + 2 * D + 2 * (d0 + d)) * RS + col] = Qh[d];
      red[(3 + 2 * D + 2 * (d0 + d)) * RS + col] = Ql[d];
    }
  }
  __syncthreads();
  constexpr int PER = COLS / GLJ_PARTS;
  for (int idx = tid; idx < nval * GLJ_PARTS; idx += GLJ_THREADS) {
    const int i = idx / GLJ_PARTS, part = idx - i * GLJ_PARTS;
    const double* rh = red + (2 * i) * RS + part * PER;
    const double* rl = red + (2 * i + 1) * RS + part * PER;
    double h = rh[0], l = rl[0];
    for (int t = 1; t < PER; ++t) acc_add(h, l, rh[t], rl[t]);
    red2[(2 * i) * GLJ_PARTS + part] = h;
    red2[(2 * i + 1) * GLJ_PARTS + part] = l;
  }
  __syncthreads();
  // totals of this CTA -> red[2*i*RS], red[(2*i+1)*RS] (renormalised)
  for (int i = tid; i < nval; i += GLJ_THREADS) {
    double h = red2[(2 * i) * GLJ_PARTS], l = red2[(2 * i + 1) * GLJ_PARTS];
    for (int p = 1; p < GLJ_PARTS; ++p) acc_add(h, l, red2[(2 * i) * GLJ_PARTS + p], red2[(2 * i + 1) * GLJ_PARTS + p]);
    double nh, nl;
    fast_two_sum(h, l, nh, nl);
    red[(2 * i) * RS] = nh;
    red[(2 * i + 1) * RS] = nl;
  }
  __syncthreads();
  const int pair = s * a.K + k;
  if (a.nsplit > 1) {
    double* mine = a.part + (static_cast<size_t>(pair) * a.nsplit + blockIdx.z) * nval2;
    for (int i = tid; i < nval2; i += GLJ_THREADS) __stcg(mine + i, red[i * RS]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned t = atomicAdd(a.ticket + pair, 1u);
      const bool last = t == static_cast<unsigned>(a.nsplit - 1);
      if (last) a.ticket[pair] = 0;  // ready for the next launch (kernels of one context are stream-ordered)
      s_misc[1] = last ? 1.0 : 0.0;
    }
    __syncthreads();
    if (s_misc[1] == 0.0) return;
    __threadfence();
    const double* all = a.part + static_cast<size_t>(pair) * a.nsplit * nval2;
    for (int i = tid; i < nval; i += GLJ_THREADS) {
      double h = __ldcg(all + 2 * i), l = __ldcg(all + 2 * i + 1);
      for (int z = 1; z < a.nsplit; ++z) acc_add(h, l, __ldcg(all + static_cast<size_t>(z) * nval2 + 2 * i), __ldcg(all + static_cast<size_t>(z) * nval2 + 2 * i + 1));
      double nh, nl;
      fast_two_sum(h, l, nh, nl);
      red[(2 * i) * RS] = nh;
      red[(2 * i + 1) * RS] = nl;
    }
    __syncthreads();
  }
  // ---- C_d = Q_d - A in two-word arithmetic, then everything rounds to one double ----
  // red2[0] = A, red2[1 + d] = B_d, red2[1 + D + d] = C_d
  if (tid < nval) {
    const double Ah_ = red[0], Al_ = red[RS];
    double v;
    if (tid == 0) {
      v = Ah_ + Al_;
    } else if (tid <= D) {
      v = red[(2 * tid) * RS] + red[(2 * tid + 1) * RS];
    } else {
      double h, l;
      dd_add(red[(2 * tid) * RS], red[(2 * tid + 1) * RS], -Ah_, -Al_, h, l);
      v = h + l;
    }
    red2[tid] = v;
  }
  __syncthreads();
  // ---- per-(s,k) epilogue ----
  double* o = a.out + static_cast<size_t>(pair) * a.ostride;
  const bool quad = a.meanfun == 4 && !a.raw;
  if (tid < D) {
    const int d = tid;
    const double lam = a.vp.lambda[d];
    const double it = s_itau[d];
    const double Bd = red2[1 + d], Cd = red2[1 + D + d];
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
    double I = red2[0] + ((a.meanfun > 0 && !a.raw) ? a.gp.m0[s] : 0.0);  // I_k = z_k*alpha + m0   (:169)
    double gs = 0.0;
    for (int d = 0; d < D; ++d) {
      const double lam = a.vp.lambda[d], it = s_itau[d], dl = a.vp.delta[d];
      gs += (lam * it) * (lam * it) * red2[1 + D + d];  // sum (lambda/tau)^2 (Delta^2-1) z alpha (:227-229)
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

// R.I[s][k] (all S rows: zero outside the local shard), R.Gmu[k][d], R.Gsig[k]: one thread per output, the <= S loads of a thread
// are independent.  R.Glam[d] = sum_k w_k sum_s glam[s][k][d]: one warp per d (blocks >= nb1), lanes take k = lane, lane + 32, ...
// Every value also goes to the peers' exchange inboxes when xc.peer is set (common.cuh xchg_push).
__global__ void __launch_bounds__(128) glj_reduce_kernel(const double* __restrict__ out, int ostride, int D, int K, int S,
                                                         int s_begin, int s_count, const double* __restrict__ w, double* __restrict__ R,
                                                         int oI, int oGmu, int oGsig, int oGlam, int nb1, XchgDev xc) {
  const int nI = S * K, nM = K * D;
  if (static_cast<int>(blockIdx.x) < nb1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int at;
    double v;
    if (i < nI) {
      const int s = i / K;
      at = oI + i;
      v = (s >= s_begin && s < s_begin + s_count) ? out[static_cast<size_t>(i) * ostride] : 0.0;
    } else {
      int r = i - nI, off, k;
      if (r < nM) {
        k = r / D; off = 2 + (r - k * D); at = oGmu + r;
      } else if (r < nM + K) {
        k = r - nM; off = 1; at = oGsig + k;
      } else {
        return;
      }
      const double* o = out + (static_cast<size_t>(s_begin) * K + k) * ostride + off;
      const size_t step = static_cast<size_t>(K) * ostride;
      double acc = 0.0;
#pragma unroll 8
      for (int s = 0; s < s_count; ++s) acc += o[s * step];
      v = acc;
    }
    R[at] = v;
    xchg_push(xc, at, v);
  } else {
    const int d = (blockIdx.x - nb1) * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (d >= D) return;
    double acc = 0.0;
    for (int k = lane; k < K; k += 32) {
      const double* o = out + (static_cast<size_t>(s_begin) * K + k) * ostride + 2 + D + d;
      const size_t step = static_cast<size_t>(K) * ostride;
      double a2 = 0.0;
      for (int s = 0; s < s_count; ++s) a2 += o[s * step];
      acc = fma(w[k], a2, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      R[oGlam + d] = acc;
      xchg_push(xc, oGlam + d, acc);
    }
  }
  if (xc.peer) __threadfence_system();
}

static int pick_dp(int D) {
  static const int opts[] = {2, 4, 6, 8, 10, 12, 16, 20, 24};
  for (int o : opts)
    if (D <= o) return o;
  return -1;
}

template <int DPH, int HALVES>
static int launch_glj_t(vbmc_b200_ctx* c, const GljArgs& a, cudaStream_t st) {
  const size_t smem = glj_smem_bytes(a.D, DPH * HALVES, HALVES);
  auto kern = glj_kernel<DPH, HALVES>;
  if (smem > 48 * 1024)
    VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(a.K, a.s_count, a.nsplit);
  KernelScope ks(c, "gplogjoint", st);
  kern<<<grid, GLJ_THREADS, smem, st>>>(a);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

// slices of the N axis: enough CTAs for ~4 per SM, at least 2*GLJ_THREADS points per slice; region 0 of the scratch buffers
// belongs to the step's own launch, region 1 to the weighted (variance-gradient) launch, which may be enqueued on another stream.
static int glj_dispatch(vbmc_b200_ctx* c, GljArgs& a, int region, cudaStream_t st) {
  const int pairs = a.K * a.s_count;
  int nsplit = (4 * c->num_sms + pairs - 1) / pairs;
  const int max_split = (a.N + 2 * GLJ_THREADS - 1) / (2 * GLJ_THREADS);
  nsplit = nsplit > max_split ? max_split : nsplit;
  nsplit = nsplit > 16 ? 16 : (nsplit < 1 ? 1 : nsplit);
  static const int force = getenv("VBMC_B200_GLJ_NSPLIT") ? atoi(getenv("VBMC_B200_GLJ_NSPLIT")) : 0;
  if (force > 0) nsplit = force > 16 ? 16 : force;
  a.nsplit = nsplit;
  const size_t npair_all = static_cast<size_t>(a.S) * a.K;
  const size_t nval2 = 2 * (1 + 2 * static_cast<size_t>(a.D));
  const size_t part_doubles = npair_all * 16 * nval2;  // capacity for the largest nsplit: the buffer does not move between launches
  if (c->glj_part.cap < 2 * part_doubles * sizeof(double)) {
    VB_CUDA(cudaStreamSynchronize(st));
    VB_TRY(c->glj_part.reserve(2 * part_doubles * sizeof(double)));
  }
  // arrival counters live in their own buffer: they must stay zero between launches whatever shape comes next
  if (c->glj_ticket.cap < 2 * npair_all * sizeof(unsigned)) {
    VB_CUDA(cudaStreamSynchronize(st));
    VB_TRY(c->glj_ticket.reserve(2 * npair_all * sizeof(unsigned) + 4096));
    VB_CUDA(cudaMemset(c->glj_ticket.p, 0, c->glj_ticket.cap));
  }
  a.part = c->glj_part.d() + static_cast<size_t>(region) * part_doubles;
  a.ticket = static_cast<unsigned*>(c->glj_ticket.p) + static_cast<size_t>(region) * (c->glj_ticket.cap / (2 * sizeof(unsigned)));
  switch (pick_dp(a.D)) {
    case 2: return launch_glj_t<2, 1>(c, a, st);
    case 4: return launch_glj_t<4, 1>(c, a, st);
    case 6: return launch_glj_t<6, 1>(c, a, st);
    case 8: return launch_glj_t<8, 1>(c, a, st);
    case 10: return launch_glj_t<10, 1>(c, a, st);
    case 12: return launch_glj_t<12, 1>(c, a, st);
    case 16: return launch_glj_t<8, 2>(c, a, st);
    case 20: return launch_glj_t<10, 2>(c, a, st);
    case 24: return launch_glj_t<12, 2>(c, a, st);
  }
  VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:gplogjoint: D=%d > 24 is not supported by this build", a.D);
}

// all_samples: ignore the rank's shard and evaluate every hyper-parameter sample (the variance gradient needs the per-sample
// gradients of ALL samples on every rank: gplogjoint.m:407-410)
int launch_gplogjoint(vbmc_b200_ctx* c, int all_samples, cudaStream_t st) {
  GljArgs a;
  a.N = c->gp.N; a.D = c->D; a.K = c->K; a.S = c->gp.S;
  if (all_samples) {
    a.s_begin = 0;
    a.s_count = a.S;
  } else {
    shard_range(a.S, c->nranks, c->rank, &a.s_begin, &a.s_count);
    a.s_count -= a.s_begin;
  }
  a.meanfun = c->gp.meanfun;
  a.ostride = 2 + 2 * a.D;
  a.gp = c->gp;
  a.vp = c->vp;
  a.wvec = nullptr;
  a.raw = 0;
  VB_TRY(c->glj_out.reserve(sizeof(double) * static_cast<size_t>(a.S) * a.K * a.ostride));
  a.out = c->glj_out.d();
  if (a.s_count <= 0) return VBMC_B200_OK;
  return glj_dispatch(c, a, 0, st);
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
  return glj_dispatch(c, a, 1, st);
}

int launch_glj_reduce(vbmc_b200_ctx* c, cudaStream_t st, bool whole_step) {
  RLayout rl;
  rl.init(c->D, c->K, c->gp.S);
  int sb, se;
  shard_range(c->gp.S, c->nranks, c->rank, &sb, &se);
  double* R = c->R_dev.d();
  KernelScope ks(c, "reduce", st);
  const int nout = c->gp.S * c->K + c->K * c->D + c->K;
  const int nb1 = (nout + 127) / 128;
  glj_reduce_kernel<<<nb1 + (c->D + 3) / 4, 128, 0, st>>>(c->glj_out.d(), 2 + 2 * c->D, c->D, c->K, c->gp.S, sb, se - sb, c->vp.w, R, rl.oI,
                                                         rl.oGmu, rl.oGsig, rl.oGlam, nb1, step_push_target(c, c->gp.S, whole_step));
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

}  // namespace vb
