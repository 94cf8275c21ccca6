// gplogjoint on sm_100a: expected GP log-joint by Bayesian quadrature and its gradient
// (reference: misc/gplogjoint.m:98-271).
//
// For each hyper-parameter sample s and mixture component k the reference builds D x N
// temporaries; everything it needs collapses to 2D+1 reductions over the N training points
// (SURVEY.md Appendix B.1):
//     zeta_n = exp(lnnf_sk - 0.5*sum_d Delta_dn^2) * alpha_sn,  Delta_dn = (mu_kd - X_nd)/tau_skd
//     A = sum zeta_n,   B_d = sum zeta_n Delta_dn,   C_d = sum zeta_n (Delta_dn^2 - 1)
// X (N x D column-major => unit stride over n) and alpha_s are streamed coalesced; the working set (X, alpha) is
// < 1 MB and stays in L2, the kernel is FP64-pipe bound.
//
// Accuracy (dd_math.cuh): alpha_n ~ 1e4 while the sums are O(1), so every n-dependent quantity is carried as a
// two-word (h, l) value -- the terms, the three running sums, the warp and segment reductions -- and the results
// are the correctly rounded sums of the exact formula to ~1e-20 * sum|zeta_n|.  tests/test_truth128.py checks the
// outputs against an IEEE binary128 evaluation on ill-conditioned posteriors (plain FP64 is off by 1e-11..5e-9 there).
//
// Schedule: the flattened (pair, n) space of this rank -- pair = (local s, k), n = training point -- is cut into W
// equal contiguous ranges, ONE WARP per range (4 warps per CTA, no block barrier after the table load).  A range touches
// at most a few pairs; for each it sweeps its segment of the N axis, reduces the 2D+1 two-word sums inside the warp and
// writes them to the pair's segment slot.  The LAST segment of a pair to finish (atomic ticket) adds the segments in
// order and runs the per-(s,k) epilogue (gplogjoint.m:169-174, 206-210, 227-231, 248-252): perfectly balanced whatever
// S_local * K is (a multi-GPU shard of 2-3 samples fills the machine like the full problem), one launch, deterministic.
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
  int W;               // number of ranges (= warps in the grid)
  int maxseg;          // segment slots per pair
  double* part;        // [s_count*K][maxseg][2*(1+2D)] segment partials
  unsigned* ticket;    // [s_count*K] arrival counters, zero between launches
};

constexpr int GLJ_THREADS = 128;   // 4 independent warps

__device__ const double2 g_exp2_tab[64] = {VB_EXP2_TABLE_ROWS};

// per-warp shared memory (doubles): mu [DP] | itau [DP] | red [2*(1+2D)][COLS+1]   (+ the CTA's exp2 table [128] in front)
static size_t glj_warp_doubles(int D, int DP, int halves) {
  return 2 * static_cast<size_t>(DP) + 2 * (1 + 2 * static_cast<size_t>(D)) * (32 / halves + 1);
}
static size_t glj_smem_bytes(int D, int DP, int halves) { return sizeof(double) * (128 + 4 * glj_warp_doubles(D, DP, halves)); }

// first range that owns a point of flattened index q:  range w covers [ceil(w*T/W), ceil((w+1)*T/W))  <=>  owner(q) = floor(q*W/T)
__device__ __forceinline__ int glj_owner(long long q, long long T, int W) { return static_cast<int>((q * W) / T); }

// DPH dimensions per lane, HALVES lanes per training point (HALVES = 2 for D > 12: the 4*D + 2 accumulator words of one
// thread would not fit the register file; the lane pair exchanges its two partial sum_d Delta^2 with one shuffle and both
// lanes evaluate the exponential).
template <int DPH, int HALVES>
__global__ void __launch_bounds__(GLJ_THREADS, 2) glj_kernel(const GljArgs a) {
  extern __shared__ __align__(16) double sm[];
  constexpr int DP = DPH * HALVES;
  constexpr int COLS = 32 / HALVES;  // training points per warp iteration
  constexpr int RS = COLS + 1;
  const int D = a.D, N = a.N, K = a.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = HALVES == 2 ? (lane & 1) : 0, col = HALVES == 2 ? (lane >> 1) : lane;
  const int d0 = half * DPH;
  const int nval = 1 + 2 * D, nval2 = 2 * nval;
  double2* s_tab = reinterpret_cast<double2*>(sm);  // [64]
  double* wsm = sm + 128 + static_cast<size_t>(warp) * (2 * DP + static_cast<size_t>(nval2) * RS);
  double* s_mu = wsm;              // [DP]
  double* s_itau = s_mu + DP;      // [DP]
  double* red = s_itau + DP;       // [2*nval][RS]

  if (tid < 64) s_tab[tid] = g_exp2_tab[tid];
  __syncthreads();

  const int w = blockIdx.x * (GLJ_THREADS / 32) + warp;
  if (w >= a.W) return;
  const long long T = static_cast<long long>(a.s_count) * K * N;
  const long long q0 = (static_cast<long long>(w) * T + a.W - 1) / a.W, q1 = (static_cast<long long>(w + 1) * T + a.W - 1) / a.W;
  const double* __restrict__ X = a.gp.X;

  for (long long q = q0; q < q1;) {
    const int pl = static_cast<int>(q / N);             // local pair index
    const int n0 = static_cast<int>(q - static_cast<long long>(pl) * N);
    const int n1 = static_cast<int>(min(static_cast<long long>(N), q1 - static_cast<long long>(pl) * N));
    q += n1 - n0;
    const int sl = pl / K, k = pl - sl * K, s = a.s_begin + sl;
    const int pair = s * K + k;
    // ---- per-pair constants: lane d owns dimension d (gplogjoint.m:164-165) ----
    const double sigk = a.vp.sigma[k];
    double tau2 = 1.0;
    if (lane < DP) {
      double mu = 0.0, itau = 0.0;
      if (lane < D) {
        const double lam = a.vp.lambda[lane], ell = a.gp.ell[s * D + lane], dl = a.vp.delta[lane];
        tau2 = sigk * sigk * lam * lam + ell * ell + dl * dl;
        mu = a.vp.mu[k * D + lane];
        itau = 1.0 / sqrt(tau2);
      }
      s_mu[lane] = mu;
      s_itau[lane] = itau;
    }
    // sum_d log tau_d = 0.5 log prod_d tau_d^2: one logarithm (the product of <= 24 squared length scales cannot leave the
    // double range); xor-butterfly => every lane holds the same product
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tau2 *= __shfl_xor_sync(0xffffffffu, tau2, off);
    const double lnnf = a.gp.lnc[s] - 0.5 * log(tau2);  // lnnf_k = ln_sf2 + sum_lnell - sum(log(tau_k))  (:165)
    __syncwarp();

    double Ah = 0.0, Al = 0.0, Bh[DPH], Bl[DPH], Qh[DPH], Ql[DPH];
#pragma unroll
    for (int d = 0; d < DPH; ++d) Bh[d] = Bl[d] = Qh[d] = Ql[d] = 0.0;
    const double* __restrict__ alpha = a.wvec ? a.wvec + (static_cast<size_t>(s) * K + k) * N : a.gp.alpha + static_cast<size_t>(s) * N;
    // software-pipelined loads (when the registers allow it): the next iteration's coordinates are in flight while this one
    // is evaluated
    constexpr bool PREFETCH = true;
    double xn[PREFETCH ? DPH : 1], aln = 0.0;
    if (PREFETCH) {
      const int n = n0 + col;
      const bool live = n < n1;
#pragma unroll
      for (int d = 0; d < DPH; ++d) xn[PREFETCH ? d : 0] = (live && d0 + d < D) ? __ldg(X + static_cast<size_t>(d0 + d) * N + n) : 0.0;
      aln = live ? __ldg(alpha + n) : 0.0;
    }
    for (int base = n0; base < n1; base += COLS) {  // warp-uniform trip count: the lane pairs shuffle inside
      double x[DPH], dh[DPH], dl[DPH];
      double al_n;
      if (PREFETCH) {
#pragma unroll
        for (int d = 0; d < DPH; ++d) x[d] = xn[PREFETCH ? d : 0];
        al_n = aln;
        const int n = base + COLS + col;
        const bool live = n < n1;
#pragma unroll
        for (int d = 0; d < DPH; ++d) xn[PREFETCH ? d : 0] = (live && d0 + d < D) ? __ldg(X + static_cast<size_t>(d0 + d) * N + n) : 0.0;
        aln = live ? __ldg(alpha + n) : 0.0;
      } else {
        const int n = base + col;
        const bool live = n < n1;
#pragma unroll
        for (int d = 0; d < DPH; ++d) x[d] = (live && d0 + d < D) ? __ldg(X + static_cast<size_t>(d0 + d) * N + n) : 0.0;
        al_n = live ? __ldg(alpha + n) : 0.0;
      }
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
      glj_zeta(ssh, ssl, lnnf, al_n, s_tab, zh, zl);  // z_k(n)*alpha(n)  (:167-169); alpha = 0 beyond the segment
      if (half == 0) acc_add(Ah, Al, zh, zl);
      glj_accumulate<DPH>(dh, dl, zh, zl, Bh, Bl, Qh, Ql);
    }
    // ---- warp reduction in fixed order, two-word: value v of column `col` -> red[2v][col], red[2v+1][col] ----
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
    __syncwarp();
    for (int i = lane; i < nval; i += 32) {   // lane i sums value i over the columns, in column order
      const double* rh = red + (2 * i) * RS;
      const double* rl = red + (2 * i + 1) * RS;
      double h = rh[0], l = rl[0];
#pragma unroll 4
      for (int t = 1; t < COLS; ++t) acc_add(h, l, rh[t], rl[t]);
      double nh, nl;
      fast_two_sum(h, l, nh, nl);
      red[(2 * i) * RS] = nh;          // column 0 of the rows: the segment totals
      red[(2 * i + 1) * RS] = nl;
    }
    __syncwarp();
    // ---- segments of this pair: [w_first, w_last]; the last one to arrive combines them ----
    const int w_first = glj_owner(static_cast<long long>(pl) * N, T, a.W), w_last = glj_owner(static_cast<long long>(pl + 1) * N - 1, T, a.W);
    const int nseg = w_last - w_first + 1;
    if (nseg > 1) {
      double* mine = a.part + (static_cast<size_t>(pl) * a.maxseg + (w - w_first)) * nval2;
      for (int i = lane; i < nval2; i += 32) __stcg(mine + i, red[i * RS]);
      __threadfence();
      __syncwarp();
      unsigned tk = 0;
      if (lane == 0) {
        tk = atomicAdd(a.ticket + pl, 1u);
        if (tk == static_cast<unsigned>(nseg - 1)) a.ticket[pl] = 0;  // ready for the next launch (stream-ordered)
      }
      tk = __shfl_sync(0xffffffffu, tk, 0);
      if (tk != static_cast<unsigned>(nseg - 1)) continue;   // not the last segment: next pair of this range
      __threadfence();
      const double* all = a.part + static_cast<size_t>(pl) * a.maxseg * nval2;
      for (int i = lane; i < nval; i += 32) {
        double h = __ldcg(all + 2 * i), l = __ldcg(all + 2 * i + 1);
        for (int z = 1; z < nseg; ++z) acc_add(h, l, __ldcg(all + static_cast<size_t>(z) * nval2 + 2 * i), __ldcg(all + static_cast<size_t>(z) * nval2 + 2 * i + 1));
        double nh, nl;
        fast_two_sum(h, l, nh, nl);
        red[(2 * i) * RS] = nh;
        red[(2 * i + 1) * RS] = nl;
      }
      __syncwarp();
    }
    // ---- C_d = Q_d - A in two-word arithmetic, then everything rounds to one double; per-(s,k) epilogue, lane d owns d ----
    const double A_h = red[0], A_l = red[RS];
    const double Aval = A_h + A_l;
    double gs = 0.0, Iq = 0.0;
    double* o = a.out + static_cast<size_t>(pair) * a.ostride;
    const bool quad = a.meanfun == 4 && !a.raw;
    if (lane < D) {
      const int d = lane;
      const double Bd = red[(2 + 2 * d) * RS] + red[(3 + 2 * d) * RS];
      double ch, cl;
      dd_add(red[(2 + 2 * D + 2 * d) * RS], red[(3 + 2 * D + 2 * d) * RS], -A_h, -A_l, ch, cl);
      const double Cd = ch + cl;
      const double lam = a.vp.lambda[d], it = s_itau[d], dlt = a.vp.delta[d], m = s_mu[d];
      double gmu = -Bd * it;                                  // w(k)*dz_dmu*alpha / w(k)   (:206-208)
      double glam = sigk * sigk * lam * (Cd * it * it);       // (:248-249) / w(k)
      gs = (lam * it) * (lam * it) * Cd;                      // sum (lambda/tau)^2 (Delta^2-1) z alpha (:227-229)
      if (quad) {
        const double io2 = a.gp.iom2[s * D + d], xm = a.gp.xm[s * D + d];
        gmu -= io2 * (m - xm);                                // (:210)
        glam -= sigk * sigk * lam * io2;                      // (:252)
        Iq = -0.5 * io2 * (m * m + sigk * sigk * lam * lam - 2.0 * m * xm + xm * xm + dlt * dlt);  // nu_k (:172-174)
        gs -= io2 * lam * lam;                                // (:231)
      }
      o[2 + d] = gmu;
      o[2 + D + d] = glam;
    }
    // sums over d in ascending lane order (sequential adds through shuffles: same order as a serial loop)
    double gsum = 0.0, isum = 0.0;
    for (int d = 0; d < D; ++d) {
      gsum += __shfl_sync(0xffffffffu, gs, d);
      isum += __shfl_sync(0xffffffffu, Iq, d);
    }
    if (lane == 0) {
      o[0] = Aval + ((a.meanfun > 0 && !a.raw) ? a.gp.m0[s] : 0.0) + isum;  // I_k = z_k*alpha + m0 + nu_k   (:169-174)
      o[1] = sigk * gsum;
    }
    __syncwarp();  // red / s_mu are rewritten by the next pair
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
    // lanes stride over the flattened (local s, k) pairs: ~ s_count*K/32 INDEPENDENT loads per lane, four in flight per iteration
    // (the k-outer / s-inner loop made every lane wait for 2 x s_count dependent round trips to L2: 51 % of this kernel's samples)
    const double* o = out + static_cast<size_t>(s_begin) * K * ostride + 2 + D + d;
    const int npair = s_count * K;
    double a4[4] = {0.0, 0.0, 0.0, 0.0};
    int p = lane;
    for (; p + 96 < npair; p += 128) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int pq = p + 32 * q;
        a4[q] = fma(w[pq % K], o[static_cast<size_t>(pq) * ostride], a4[q]);
      }
    }
    for (; p < npair; p += 32) a4[0] = fma(w[p % K], o[static_cast<size_t>(p) * ostride], a4[0]);
    double acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
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
  KernelScope ks(c, "gplogjoint", st);
  kern<<<(a.W + 3) / 4, GLJ_THREADS, smem, st>>>(a);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

// number of ranges: one warp each; 12 warps per SM fill the register file (168 registers per thread); at least
// GLJ_MIN_POINTS points per range so that the per-segment reduction stays a small fraction of the sweep.
// Region 0 of the scratch buffers belongs to the step's own launch, region 1 to the weighted (variance-gradient) launch,
// which may be enqueued on another stream.
static int glj_dispatch(vbmc_b200_ctx* c, GljArgs& a, int region, cudaStream_t st) {
  const long long T = static_cast<long long>(a.s_count) * a.K * a.N;
  static const int min_pts = getenv("VBMC_B200_GLJ_MIN_POINTS") ? atoi(getenv("VBMC_B200_GLJ_MIN_POINTS")) : 512;
  static const int wps = getenv("VBMC_B200_GLJ_WARPS_PER_SM") ? atoi(getenv("VBMC_B200_GLJ_WARPS_PER_SM")) : 8;
  long long W = static_cast<long long>(wps) * c->num_sms;
  const long long wmax = (T + min_pts - 1) / (min_pts > 0 ? min_pts : 1);
  if (W > wmax) W = wmax;
  if (W < 1) W = 1;
  a.W = static_cast<int>(W);
  const long long per = (T + W - 1) / W;                       // points per range
  a.maxseg = static_cast<int>((a.N + per - 1) / per) + 1;       // a pair can straddle that many ranges
  const size_t npair_all = static_cast<size_t>(a.S) * a.K;
  const size_t nval2 = 2 * (1 + 2 * static_cast<size_t>(a.D));
  const size_t part_doubles = npair_all * static_cast<size_t>(a.maxseg) * nval2;
  if (c->glj_part.cap < 2 * part_doubles * sizeof(double)) {
    VB_CUDA(cudaStreamSynchronize(st));
    VB_TRY(c->glj_part.reserve(2 * part_doubles * sizeof(double) + (1 << 20)));
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
