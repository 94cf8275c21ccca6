// Small single-CTA kernels around the two hot kernels of a negelcbo step:
//   vp_unpack_kernel : theta -> vp.mu/sigma/lambda/eta/w          (misc/negelcbo_vbmc.m:32-48)
//   finalize_kernel  : reduced sums R -> H,dH (ent/entmc_vbmc.m:67,82-125), G,dG
//                      (misc/gplogjoint.m:352-413), soft-bound + weight penalties
//                      (misc/vpbndloss.m:9-71, utils/softbndloss.m:9-28, negelcbo_vbmc.m:136-164),
//                      F = -G - H + L, dF = -dG - dH + dL (negelcbo_vbmc.m:116-117).
// O(DK) work; they exist so that a step needs no host round trip between kernels.
#include <stdlib.h>

#include "common.cuh"

namespace vb {

struct UnpackArgs {
  int D, K, ntheta, have_theta, force_form, DPc, want_cblob;
  int opt[4];
  const double* theta;
  const unsigned long long* dyn_src;  // {seed, stream} of this step's draws behind theta in the staging buffer (or NULL)
  const double *base_mu, *base_sigma, *base_lambda, *base_w, *base_eta;
  VpDev vp;
  double* cn;  // [K] nf/sigma_k^D ; cn[K] = nf
  double* scratch;  // [K*D]
  // cost-weighted schedule of the FP64 entropy sweep (vp_unpack2_kernel only; null: none requested)
  int* plan;          // tstart[G + 1] | jlo[K] | jhi[K]
  int* plan_w;        // [K] scratch: estimated cost of one tile of component j
  int plan_tpc, plan_G, plan_c0, plan_crun;   // crun: fixed cost of starting a source component inside a range (table build, run result), in thirds of a scored component
  double plan_prune, plan_emax[3];   // pruning constant of the sweep; 10 / 50 / 90 % quantiles of the largest ||eps|| among a warp's 32 draws
};

// One CTA.  Everything the later phases re-read (mu, sigma, lambda, 1/(sigma lambda)^2) is kept in shared memory, so the
// kernel's critical path is three block-wide phases instead of a chain of dependent global round trips.
__global__ void __launch_bounds__(256) vp_unpack_kernel(const UnpackArgs a) {
  extern __shared__ double usm[];
  const int D = a.D, K = a.K, tid = threadIdx.x, nt = blockDim.x;
  double* s_mu = usm;                  // [K*D]
  double* s_isl2 = s_mu + K * D;       // [K*D]  1/(sigma_k lambda_d)^2
  double* s_sigma = s_isl2 + K * D;    // [K]
  double* s_eta = s_sigma + K;         // [K]
  double* s_lambda = s_eta + K;        // [D]
  __shared__ double part[256];
  __shared__ double s_es, s_nf;
  const bool ht = a.have_theta != 0;
  if (tid < 2 && a.dyn_src) a.vp.dyn_snap[tid] = a.dyn_src[tid];  // private copy for the ahead-of-time draw generator (api.cu)
  int idx = 0;
  const int o_mu = 0;
  if (ht && a.opt[0]) idx += D * K;
  const int o_sig = idx;
  if (ht && a.opt[1]) idx += K;
  const int o_lam = idx;
  const int o_eta = a.ntheta - K;
#pragma unroll 1
  for (int i = tid; i < D * K; i += nt) {
    const double m = (ht && a.opt[0]) ? a.theta[o_mu + i] : a.base_mu[i];
    s_mu[i] = m;
    a.vp.mu[i] = m;
  }
  double es = 0.0;
#pragma unroll 1
  for (int k = tid; k < K; k += nt) {
    double sg, ls;
    if (ht && a.opt[1]) {
      ls = a.theta[o_sig + k];
      sg = exp(ls);  // vp.sigma(1,:) = exp(theta(idx_start+(1:K)))  (:39-42)
    } else {
      sg = a.base_sigma[k];
      ls = log(sg);
    }
    a.vp.lnsigma[k] = ls;
    a.vp.sigma[k] = sg;
    s_sigma[k] = sg;
    const double et = (ht && a.opt[3]) ? a.theta[o_eta + k] : a.base_eta[k];
    a.vp.eta[k] = et;
    s_eta[k] = et;
    es += exp(et);  // (:46-47) no max-shift, like the reference
  }
#pragma unroll 1
  for (int d = tid; d < D; d += nt) {
    double lm, ll;
    if (ht && a.opt[2]) {
      ll = a.theta[o_lam + d];
      lm = exp(ll);  // (:43)
    } else {
      lm = a.base_lambda[d];
      ll = log(lm);
    }
    a.vp.lnlambda[d] = ll;
    a.vp.lambda[d] = lm;
    s_lambda[d] = lm;
  }
  part[tid] = es;
  __syncthreads();
#pragma unroll 1
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) part[tid] += part[tid + off];
    __syncthreads();
  }
  if (tid == 0) {
    s_es = part[0];
    double pl = 1.0;
#pragma unroll 1
    for (int d = 0; d < D; ++d) pl *= s_lambda[d];
    s_nf = 1.0 / pow(2.0 * 3.14159265358979323846, 0.5 * D) / pl;  // nf (entmc_vbmc.m:40)
    a.cn[K] = s_nf;
  }
#pragma unroll 1
  for (int i = tid; i < K * D; i += nt) {
    const double sl = s_sigma[i / D] * s_lambda[i % D];
    s_isl2[i] = 1.0 / (sl * sl);
  }
  __syncthreads();
  // ---- choose the entmc formulation for this step (see entmc.cu) ----
  // expanded form error ~ eps_mach * (||u_jk||^2 + r_jk^2 ||eps||^2); keep it below ~1e-10 absolute in d^2.
  // ||u_jk||^2 = sum_d (mu_jd - mu_kd)^2 * isl2_kd with isl2_kd = 1/(sigma_k lambda_d)^2
  {
    const double eemax = D + 12.0 * sqrt(2.0 * D) + 72.0;  // > 12 sigma bound on ||eps||^2
    double m = 0.0;
#pragma unroll 1
    for (int i = tid; i < K * K; i += nt) {
      const int j = i / K, k = i - j * K;
      double uu = 0.0;
#pragma unroll 1
      for (int d = 0; d < D; ++d) {
        const double dm = s_mu[j * D + d] - s_mu[k * D + d];
        uu = fma(dm * dm, s_isl2[k * D + d], uu);
      }
      const double r2 = s_sigma[j] * s_sigma[j] * s_isl2[k * D] * s_lambda[0] * s_lambda[0];
      const double v = fma(r2, eemax, uu);
      m = (v > m || !(v == v)) ? v : m;
    }
    part[tid] = m;
    __syncthreads();
#pragma unroll 1
    for (int off = 128; off > 0; off >>= 1) {
      if (tid < off) part[tid] = (part[tid + off] > part[tid] || !(part[tid + off] == part[tid + off])) ? part[tid + off] : part[tid];
      __syncthreads();
    }
    if (tid == 0) *a.vp.form_flag = a.force_form >= 0 ? a.force_form : ((part[0] <= 2.0e5) ? 2 : 1);
  }
  const int K2 = (K + 1) & ~1, DPc = a.DPc;
  double* b_mu = a.vp.cblob;              // [K2][DPc] means centred on their average (only differences matter)
  double* b_ck = b_mu + K2 * DPc;         // [K2]
  double* b_ak = b_ck + K2;               // [K2] ak_k / sigma_k
  double* b_il = b_ak + K2;               // [DPc] 1/lambda_d
#pragma unroll 1
  for (int k = tid; k < K2; k += nt) {
    double ck = 0.0, aks = 0.0;
    if (k < K) {
      const double w = (ht && a.opt[3]) ? exp(s_eta[k]) / s_es : a.base_w[k];
      a.vp.w[k] = w;
      const double sg = s_sigma[k];
      const double cn = s_nf / pow(sg, static_cast<double>(D));  // nf/sigma(k)^D  (:63)
      a.cn[k] = cn;
      ck = w * cn;
      a.vp.ck[k] = ck;
      a.vp.ak[k] = ck / sg;
      aks = ck / (sg * sg);
    }
    b_ck[k] = ck;
    b_ak[k] = aks;
  }
  if (a.want_cblob) {  // only the experimental separable entmc form reads the centred means
#pragma unroll 1
    for (int d = tid; d < DPc; d += nt) {
      double mean = 0.0;
      if (d < D) {
#pragma unroll 1
        for (int k = 0; k < K; ++k) mean += s_mu[k * D + d];
        mean /= K;
      }
      b_il[d] = d < D ? 1.0 / s_lambda[d] : 0.0;
#pragma unroll 1
      for (int k = 0; k < K2; ++k) b_mu[k * DPc + d] = (d < D && k < K) ? s_mu[k * D + d] - mean : 0.0;
    }
  }
}

// Multi-CTA theta -> vp (default; the one-CTA kernel above stays for VBMC_B200_UNPACK_V1=1).  grid = K CTAs of 128 threads.
// Every CTA recomputes the K + D exponentials it needs (sigma, eta -> softmax normaliser, lambda: a few hundred flops), CTA j
// then owns component j: its row of mu, sigma_j, w_j, c_j, a_j and row j of the form guard max_k(||u_jk||^2 + r_jk^2 eemax);
// the rows' maxima meet in one atomicMax on the bit pattern of a non-negative double, and the last CTA to arrive publishes the
// form flag.  Same formulas as vp_unpack_kernel (the softmax normaliser is summed over 128 instead of 256 threads: last-bit differences).
__global__ void __launch_bounds__(128) vp_unpack2_kernel(const UnpackArgs a) {
  extern __shared__ double usm[];
  const int D = a.D, K = a.K, tid = threadIdx.x, nt = blockDim.x, j = blockIdx.x;
  double* s_sigma = usm;             // [K]
  double* s_lambda = s_sigma + K;    // [D]
  double* s_muj = s_lambda + D;      // [D]
  double* s_lsig = s_muj + D;        // [K] log sigma_k
  double* s_ilam = s_lsig + K;       // [D] 1 / lambda_d (guard rows and schedule weights only: no parity-relevant value uses it)
  __shared__ double part[128];
  __shared__ double s_es, s_nf;
  __shared__ int s_cnt, s_last;
  const bool ht = a.have_theta != 0;
  if (tid == 0) { s_cnt = 0; s_last = 0; }
  if (j == 0 && tid < 2 && a.dyn_src) a.vp.dyn_snap[tid] = a.dyn_src[tid];
  int idx = 0;
  const int o_mu = 0;
  if (ht && a.opt[0]) idx += D * K;
  const int o_sig = idx;
  if (ht && a.opt[1]) idx += K;
  const int o_lam = idx;
  const int o_eta = a.ntheta - K;
  for (int d = tid; d < D; d += nt) {
    const double m = (ht && a.opt[0]) ? a.theta[o_mu + j * D + d] : a.base_mu[j * D + d];
    s_muj[d] = m;
    a.vp.mu[j * D + d] = m;
  }
  // softmax normaliser: the same left-to-right-per-thread + tree order in every CTA => the same es everywhere
  double es = 0.0;
  for (int k = tid; k < K; k += nt) {
    double sg, ls;
    if (ht && a.opt[1]) {
      ls = a.theta[o_sig + k];
      sg = exp(ls);  // vp.sigma(1,:) = exp(theta(idx_start+(1:K)))  (negelcbo_vbmc.m:39-42)
    } else {
      sg = a.base_sigma[k];
      ls = log(sg);
    }
    s_sigma[k] = sg;
    s_lsig[k] = ls;
    const double et = (ht && a.opt[3]) ? a.theta[o_eta + k] : a.base_eta[k];
    es += exp(et);  // (:46-47) no max-shift, like the reference
    if (k == j) {
      a.vp.lnsigma[k] = ls;
      a.vp.sigma[k] = sg;
      a.vp.eta[k] = et;
    }
  }
  for (int d = tid; d < D; d += nt) {
    double lm, ll;
    if (ht && a.opt[2]) {
      ll = a.theta[o_lam + d];
      lm = exp(ll);  // (:43)
    } else {
      lm = a.base_lambda[d];
      ll = log(lm);
    }
    s_lambda[d] = lm;
    s_ilam[d] = 1.0 / lm;
    if (j == 0) {
      a.vp.lnlambda[d] = ll;
      a.vp.lambda[d] = lm;
    }
  }
  part[tid] = es;
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if (tid < off) part[tid] += part[tid + off];
    __syncthreads();
  }
  if (tid == 0) {
    s_es = part[0];
    double pl = 1.0;
    for (int d = 0; d < D; ++d) pl *= s_lambda[d];
    s_nf = 1.0 / pow(2.0 * 3.14159265358979323846, 0.5 * D) / pl;  // nf (entmc_vbmc.m:40)
    if (j == 0) a.cn[K] = s_nf;
  }
  __syncthreads();
  if (tid == 0) {
    const double et = (ht && a.opt[3]) ? a.theta[o_eta + j] : a.base_eta[j];
    const double w = (ht && a.opt[3]) ? exp(et) / s_es : a.base_w[j];
    a.vp.w[j] = w;
    const double sg = s_sigma[j];
    const double cn = s_nf / pow(sg, static_cast<double>(D));  // nf/sigma(k)^D  (:63)
    a.cn[j] = cn;
    const double ck = w * cn;
    a.vp.ck[j] = ck;
    a.vp.ak[j] = ck / sg;
  }
  // ---- row j of the form guard (see vp_unpack_kernel) ----
  const double eemax = D + 12.0 * sqrt(2.0 * D) + 72.0;  // > 12 sigma bound on ||eps||^2
  double m = 0.0;
  for (int k = tid; k < K; k += nt) {
    double uu = 0.0;
    const double isg = 1.0 / s_sigma[k];
    for (int d = 0; d < D; ++d) {
      const double mk = (ht && a.opt[0]) ? a.theta[o_mu + k * D + d] : a.base_mu[k * D + d];
      const double u = (s_muj[d] - mk) * isg * s_ilam[d];
      uu = fma(u, u, uu);
    }
    const double r = s_sigma[j] * isg;
    const double v = fma(r * r, eemax, uu);
    m = (v > m || !(v == v)) ? v : m;
    if (a.plan) {
      // will component k survive the sweep's pruning test (entmc2.cu) for a typical warp of draws of component j?
      // log(ck_k / ck_j) = log(w_k / w_j) + D (log sigma_j - log sigma_k)
      const double lw = (ht && a.opt[3]) ? a.theta[o_eta + k] - a.theta[o_eta + j] : log(a.base_w[k] / a.base_w[j]);
      const double un = sqrt(uu), cst = a.plan_prune + lw + D * (s_lsig[j] - s_lsig[k]) + 0.5;
      int hits = 0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {   // the test depends on the warp's largest ||eps||: average over its distribution
        const double tt = un - r * a.plan_emax[q];
        const double bb = tt > 0.0 ? -0.5 * tt * tt : 0.0;
        const double lhs = bb + 0.5 * a.plan_emax[q] * a.plan_emax[q] + cst;
        hits += (!(lhs < 0.0) || a.plan_prune <= 0.0) ? 1 : 0;
      }
      if (hits) atomicAdd(&s_cnt, hits);
    }
  }
  part[tid] = m;
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if (tid < off) part[tid] = (part[tid + off] > part[tid] || !(part[tid + off] == part[tid + off])) ? part[tid + off] : part[tid];
    __syncthreads();
  }
  if (tid == 0) {
    if (a.plan) a.plan_w[j] = 3 * a.plan_c0 + (s_cnt < 1 ? 1 : s_cnt);   // in thirds of a scored component
    // non-negative doubles order like their bit patterns; a NaN (0x7ff8...) is larger than every finite value
    unsigned long long* gmax = reinterpret_cast<unsigned long long*>(a.vp.form_flag) + 1;
    unsigned* ticket = reinterpret_cast<unsigned*>(a.vp.form_flag) + 1;
    atomicMax(gmax, static_cast<unsigned long long>(__double_as_longlong(part[0] == part[0] ? fabs(part[0]) : part[0])));
    __threadfence();
    if (atomicAdd(ticket, 1u) == static_cast<unsigned>(K - 1)) {
      __threadfence();
      const double gm = __longlong_as_double(static_cast<long long>(atomicExch(gmax, 0ULL)));
      *a.vp.form_flag = a.force_form >= 0 ? a.force_form : ((gm <= 2.0e5) ? 2 : 1);
      *ticket = 0;
      s_last = 1;
    }
  }
  if (!a.plan) return;
  __syncthreads();
  if (!s_last) return;
  // ---- last CTA to arrive: cut the K * tpc tiles of the sweep into G contiguous ranges of equal estimated cost ----
  // (every tile of component j costs w_j; boundaries fall on tiles; tests/_sweep_plan.py restates this arithmetic in integers.
  // All quantities are integers below 2^53 carried as doubles -- exact, and an order of magnitude less code than 64-bit integer
  // division: this runs once per step in ONE CTA with a cold instruction cache, code size is what it pays for; loops not unrolled.)
  __shared__ double s_pref[257];   // s_pref[j] = w_0 + ... + w_{j-1}
  __shared__ int s_t[258];         // tstart
  if (tid < 32) {
    double run = 0.0;
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {   // lane: inclusive sums of its 8 consecutive components
      const int jj = 8 * tid + q;
      if (jj < K) {
        run += static_cast<double>(__ldcg(a.plan_w + jj));
        s_pref[jj + 1] = run;
      }
    }
    double inc = run;
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
      const double v = __shfl_up_sync(0xffffffffu, inc, off);
      if (tid >= off) inc += v;
    }
    const double base = inc - run;
    if (tid == 0) s_pref[0] = 0.0;
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
      const int jj = 8 * tid + q;
      if (jj < K) s_pref[jj + 1] += base;
    }
  }
  __syncthreads();
  // cumulative cost in front of component jj: C(jj) = tpc * s_pref[jj] + jj * crun -- every component also charges its start
  // (a CTA pays one table build per source component it touches; the one at the head of its range is common to all CTAs)
  const int G = a.plan_G, tpc = a.plan_tpc;
  const double crun = a.plan_crun, dtpc = tpc;
  const double Wtot = s_pref[K] * dtpc + K * crun;
#pragma unroll 1
  for (int b = tid; b <= G; b += nt) {
    const double target = floor(Wtot * b / G);
    int lo = 0, hi = K;   // largest jj in [0, K] with C(jj) <= target
#pragma unroll 1
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_pref[mid] * dtpc + mid * crun <= target) lo = mid; else hi = mid - 1;
    }
    int t = K * tpc;
    if (lo < K) {
      const double wj = s_pref[lo + 1] - s_pref[lo];
      const double off = target - (s_pref[lo] * dtpc + lo * crun) - crun;
      double q = off <= 0.0 ? 0.0 : floor((2.0 * off + wj) / (2.0 * wj));   // nearest tile boundary
      if (q > dtpc) q = dtpc;
      t = lo * tpc + static_cast<int>(q);
    }
    if (b == 0) t = 0;
    if (b == G) t = K * tpc;
    s_t[b] = t;
    a.plan[b] = t;
  }
  __syncthreads();
#pragma unroll 1
  for (int jj = tid; jj < K; jj += nt) {
    const int tlo = jj * tpc, thi = (jj + 1) * tpc;
    int lo = 0, hi = G - 1;   // smallest b in [0, G) with s_t[b + 1] > tlo
#pragma unroll 1
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_t[mid + 1] > tlo) hi = mid; else lo = mid + 1;
    }
    a.plan[G + 1 + jj] = lo;
    lo = 0; hi = G - 1;       // largest b in [0, G) with s_t[b] < thi
#pragma unroll 1
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_t[mid] < thi) lo = mid; else hi = mid - 1;
    }
    a.plan[G + 1 + K + jj] = lo;
  }
}

struct FinArgs {
  int D, K, S, Ns, ntheta_out;
  int gf[4];       // which gradient blocks are produced
  int jacobian;    // jacobian_flag
  int what;        // FIN_*
  int use_bnd, nbnd, opt[4];
  int stage_R;     // R fits in shared memory next to the work arrays
  double TolCon, WThresh, WPen;
  const double* R;
  double* Rw;      // == R, writable: receives the all-reduced sums when they do not fit in shared memory
  XchgDev xc;      // nranks > 1: all-reduce R across ranks over peer memory before it is used (common.cuh)
  int pushed;      // the reduction kernels already wrote this rank's R into every peer's inbox (xchg_push): only publish, wait, sum
  const double* lb;
  const double* ub;
  const double* cn;  // [K+1]
  VpDev vp;
  double* out;
};

__device__ __noinline__ double soft_pen(double x, double lb, double ub, double tol, double* dy) {
  // utils/softbndloss.m:12-27
  const double ell = (ub - lb) * tol;
  double y = 0.0;
  *dy = 0.0;
  if (x < lb) {
    const double t = (lb - x) / ell;
    y = 0.5 * t * t;
    *dy = (x - lb) / (ell * ell);
  } else if (x > ub) {
    const double t = (x - ub) / ell;
    y = 0.5 * t * t;
    *dy = (x - ub) / (ell * ell);
  }
  return y;
}

// deterministic block-wide sum (fixed order) of one value per thread; blockDim.x == 256.  Out of line and shuffle based:
// these one-CTA kernels run once per step with a cold instruction cache, so code size is what they pay for.
__device__ __noinline__ double block_sum256(double v, double* part) {
  const int tid = threadIdx.x;
#pragma unroll 1
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((tid & 31) == 0) part[tid >> 5] = v;
  __syncthreads();
  double t = part[0];
#pragma unroll 1
  for (int i = 1; i < 8; ++i) t += part[i];
  return t;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// One-shot all-reduce of this rank's partial sums R[0..total) over NVLink peer memory (layout: XchgDev, common.cuh):
// push R into slot [parity][my rank] of EVERY rank's inbox (plain stores to the IPC-mapped peer buffers), publish the step's
// sequence number in every rank's flag word (release, system scope), wait until all ranks' flags show it in the local
// buffer (acquire), then sum the nranks local slots in rank order into dst.  No kernel launch and no NCCL call between the
// partial sums and their use; every rank adds the same numbers in the same order => identical results everywhere.
__device__ __noinline__ void exchange_sum(const XchgDev& xc, const double* __restrict__ R, int total, double* dst, int pushed) {
  const int tid = threadIdx.x, nt = blockDim.x, nr = xc.nranks;
  unsigned long long* me = xc.peer[xc.rank];
  const unsigned long long seq = me[XCHG_SEQ] + 1;
  const int par = static_cast<int>(seq & 1);
  const size_t slot = (static_cast<size_t>(par) * nr + xc.rank) * xc.cap;
#pragma unroll 1
  for (int i0 = tid; i0 < (pushed ? 0 : total); i0 += 4 * nt) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (i0 + u * nt < total) ? R[i0 + u * nt] : 0.0;
#pragma unroll 1
    for (int r = 0; r < nr; ++r) {
      double* out = reinterpret_cast<double*>(xc.peer[r] + XCHG_HDR) + slot;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * nt < total) out[i0 + u * nt] = v[u];
    }
  }
  __threadfence_system();
  __syncthreads();
  bool late = false;
  if (tid < nr) {
    st_release_sys(xc.peer[tid] + par * XCHG_MAXR + xc.rank, seq);
    const long long t0 = clock64();
    while (ld_acquire_sys(me + par * XCHG_MAXR + tid) < seq) {
      if (clock64() - t0 > xc.timeout_cycles) {
        late = true;
        break;
      }
    }
  }
  late = __syncthreads_or(late) != 0;
  const double* in0 = reinterpret_cast<const double*>(me + XCHG_HDR) + static_cast<size_t>(par) * nr * xc.cap;
#pragma unroll 1
  for (int i0 = tid; i0 < total; i0 += 8 * nt) {
    double acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.0;
#pragma unroll 1
    for (int r = 0; r < nr; r += 4) {
      double x[4][8];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int u = 0; u < 8; ++u)
          x[q][u] = (r + q < nr && i0 + u * nt < total) ? __ldcg(in0 + static_cast<size_t>(r + q) * xc.cap + i0 + u * nt) : 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (r + q < nr) acc[u] += x[q][u];  // rank order
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i0 + u * nt < total) dst[i0 + u * nt] = late ? __longlong_as_double(0x7ff8000000000000LL) : acc[u];
  }
  if (tid == 0) {
    me[XCHG_SEQ] = seq;
    if (late) me[XCHG_ERR] = 1;
  }
}

__global__ void __launch_bounds__(256) finalize_kernel(const FinArgs a) {
  extern __shared__ double sm[];
  const int D = a.D, K = a.K, S = a.S, tid = threadIdx.x, nt = blockDim.x;
  RLayout rl;
  rl.init(D, K, S);
  OutLayout ol;
  ol.init(a.ntheta_out, S, K);
  double* out = a.out;
  double* wsm = sm;            // [K]   softmax(eta)
  double* gHw = wsm + K;       // [K]   entropy w-grad (before J_w)
  double* gGw = gHw + K;       // [K]   log-joint w-grad (before J_w)
  double* gPw = gGw + K;       // [K]   weight-penalty w-grad (before J_w)
  double* dls = gPw + K;       // [D*K] d(penalty)/d(lnscale)
  double* part = dls + D * K;  // [256]
  double* sc = part + 256;     // scalars: 0 es, 1 H, 2 G, 3 L, 4 Lw, 5 dotH, 6 dotG, 7 dotP
  // the step's inputs are staged in shared memory once (coalesced, all loads in flight together): the serial sums
  // below then run at shared-memory latency instead of one L2 round trip per term
  double* v_mu = sc + 16;          // [K*D]
  double* v_sigma = v_mu + K * D;  // [K]
  double* v_w = v_sigma + K;       // [K]
  double* v_eta = v_w + K;         // [K]
  double* v_lnsigma = v_eta + K;   // [K]
  double* v_cn = v_lnsigma + K;    // [K+1]
  double* v_lambda = v_cn + K + 1; // [D]
  double* v_lnlambda = v_lambda + D;  // [D]
  double* sR = v_lnlambda + D;     // [rl.total] when a.stage_R
#pragma unroll 1
  for (int i = tid; i < K * D; i += nt) v_mu[i] = a.vp.mu[i];
#pragma unroll 1
  for (int k = tid; k < K; k += nt) {
    v_sigma[k] = a.vp.sigma[k];
    v_w[k] = a.vp.w[k];
    v_eta[k] = a.vp.eta[k];
    v_lnsigma[k] = a.vp.lnsigma[k];
  }
#pragma unroll 1
  for (int k = tid; k < K + 1; k += nt) v_cn[k] = a.cn[k];
#pragma unroll 1
  for (int d = tid; d < D; d += nt) {
    v_lambda[d] = a.vp.lambda[d];
    v_lnlambda[d] = a.vp.lnlambda[d];
  }
  if (a.xc.nranks > 1) {
    exchange_sum(a.xc, a.R, rl.total, a.stage_R ? sR : a.Rw, a.pushed);
  } else if (a.stage_R) {
    // 8 independent loads in flight per thread: the staging is one L2 round trip per 2048 doubles, not one per 256
    int i = tid;
#pragma unroll 1
    for (; i + 7 * nt < rl.total; i += 8 * nt) {
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = a.R[i + u * nt];
#pragma unroll
      for (int u = 0; u < 8; ++u) sR[i + u * nt] = t[u];
    }
#pragma unroll 1
    for (; i < rl.total; i += nt) sR[i] = a.R[i];
  }
  const double* R = a.stage_R ? sR : a.R;
  __syncthreads();
  const bool doH = a.what != FIN_GPLOGJOINT && a.what != FIN_NEGELCBO_NOENT, doG = a.what != FIN_ENTMC;
  const double invNs = 1.0 / static_cast<double>(a.Ns > 0 ? a.Ns : 1);
  const double invS = 1.0 / static_cast<double>(S > 0 ? S : 1);

  // output offsets of the gradient blocks
  int o_mu = 0, o_sig = 0, o_lam = 0, o_w = 0, n = 0;
  o_mu = n; if (a.gf[0]) n += D * K;
  o_sig = n; if (a.gf[1]) n += K;
  o_lam = n; if (a.gf[2]) n += D;
  o_w = n; if (a.gf[3]) n += K;

  {
    double es = 0.0;
#pragma unroll 1
    for (int k = tid; k < K; k += nt) es += exp(v_eta[k]);
    es = block_sum256(es, part);
    if (tid == 0) {
      sc[0] = es;
      sc[1] = sc[2] = sc[3] = sc[4] = 0.0;
    }
  }
#pragma unroll 1
  for (int i = tid; i < 8; i += nt) out[i] = 0.0;
  __syncthreads();
#pragma unroll 1
  for (int k = tid; k < K; k += nt) {
    wsm[k] = exp(v_eta[k]) / sc[0];
    gPw[k] = 0.0;
  }
#pragma unroll 1
  for (int i = tid; i < 3 * a.ntheta_out; i += nt) out[ol.oDF + i] = 0.0;
  __syncthreads();

  // ------------------------------------------------------------------ entropy (entmc_vbmc.m)
  if (doH) {
    {
      double H = 0.0;
#pragma unroll 1
      for (int j = tid; j < K; j += nt) H -= v_w[j] * R[rl.oHs + j] * invNs;  // :67
      H = block_sum256(H, part);
      if (tid == 0) {
        sc[1] = H;
        out[ol.oH] = H;
      }
    }
    if (a.gf[0])
#pragma unroll 1
      for (int i = tid; i < D * K; i += nt) {
        const int j = i / D, d = i - j * D;
        out[ol.oDH + o_mu + i] = v_w[j] * R[rl.oM + i] * invNs / v_lambda[d];  // :82
      }
    if (a.gf[1])
#pragma unroll 1
      for (int j = tid; j < K; j += nt) {
        double acc = 0.0;
#pragma unroll 1
        for (int d = 0; d < D; ++d) acc += R[rl.oE + j * D + d];  // :87-88
        double g = v_w[j] * acc * invNs;
        if (a.jacobian) g *= v_sigma[j];  // :112-114
        out[ol.oDH + o_sig + j] = g;
      }
    if (a.gf[2])
#pragma unroll 1
      for (int d = tid; d < D; d += nt) {
        double acc = 0.0;
#pragma unroll 1
        for (int j = 0; j < K; ++j) acc += v_w[j] * v_sigma[j] * R[rl.oE + j * D + d];  // :93, :106-108
        double g = acc * invNs;
        if (!a.jacobian) g /= v_lambda[d];  // :116-118
        out[ol.oDH + o_lam + d] = g;
      }
    if (a.gf[3])
#pragma unroll 1
      for (int l = tid; l < K; l += nt)
        gHw[l] = -R[rl.oHs + l] * invNs - v_cn[l] * R[rl.oWc + l] * invNs;  // :97, :100  (Wc[l] = sum_j w_j W_jl, contracted upstream)
  }
  // ------------------------------------------------------------------ expected log joint
  if (doG) {
    {
      double G = 0.0;
#pragma unroll 1
      for (int i = tid; i < S * K; i += nt) G += v_w[i % K] * R[rl.oI + i];  // F(s) += w(k)*I_k  (:203)
      G = block_sum256(G, part) * invS;                                          // mean over s (:398-399)
      if (tid == 0) {
        sc[2] = G;
        out[ol.oG] = G;
      }
    }
#pragma unroll 1
    for (int i = tid; i < S * K; i += nt) out[ol.oIsk + i] = R[rl.oI + i];
#pragma unroll 1
    for (int s = tid; s < S; s += nt) {
      double Fs = 0.0;
#pragma unroll 1
      for (int k = 0; k < K; ++k) Fs += v_w[k] * R[rl.oI + s * K + k];
      out[ol.oFs + s] = Fs;
    }
    if (a.gf[0])
#pragma unroll 1
      for (int i = tid; i < D * K; i += nt) out[ol.oDG + o_mu + i] = v_w[i / D] * R[rl.oGmu + i] * invS;
    if (a.gf[1])
#pragma unroll 1
      for (int k = tid; k < K; k += nt) {
        double g = v_w[k] * R[rl.oGsig + k] * invS;
        if (a.jacobian) g *= v_sigma[k];  // :357-359
        out[ol.oDG + o_sig + k] = g;
      }
    if (a.gf[2])
#pragma unroll 1
      for (int d = tid; d < D; d += nt) {
        double g = R[rl.oGlam + d] * invS;  // sum_k w_k (...), contracted upstream  (gplogjoint.m:248-252)
        if (a.jacobian) g *= v_lambda[d];  // :361-363
        out[ol.oDG + o_lam + d] = g;
      }
    if (a.gf[3])
#pragma unroll 1
      for (int k = tid; k < K; k += nt) {
        double acc = 0.0;
#pragma unroll 1
        for (int s = 0; s < S; ++s) acc += R[rl.oI + s * K + k];
        gGw[k] = acc * invS;  // w_grad(k,s) = I_k  (:269-271), mean over s
      }
  }
  // ------------------------------------------------------------------ penalties (FIN_NEGELCBO only)
  const bool doP = (a.what == FIN_NEGELCBO || a.what == FIN_NEGELCBO_NOENT) && a.use_bnd && a.nbnd > 0;
  if (doP) {
    // theta_ext = [mu(:); lnscale(:); eta(:)]   (vpbndloss.m:34-38)
    int b_mu = 0, b_ls = 0, b_eta = 0, nb = 0;
    b_mu = nb; if (a.opt[0]) nb += D * K;
    b_ls = nb; if (a.opt[1] || a.opt[2]) nb += D * K;
    b_eta = nb; if (a.opt[3]) nb += K;
    double lacc = 0.0;
    if (a.opt[0])
#pragma unroll 1
      for (int i = tid; i < D * K; i += nt) {
        double dy;
        lacc += soft_pen(v_mu[i], a.lb[b_mu + i], a.ub[b_mu + i], a.TolCon, &dy);
        if (a.gf[0]) out[ol.oDF + o_mu + i] = dy;
      }
    if (a.opt[1] || a.opt[2])
#pragma unroll 1
      for (int i = tid; i < D * K; i += nt) {
        const int k = i / D, d = i - k * D;
        double dy;
        lacc += soft_pen(v_lnsigma[k] + v_lnlambda[d], a.lb[b_ls + i], a.ub[b_ls + i], a.TolCon, &dy);
        dls[i] = dy;
      }
    if (a.opt[3])
#pragma unroll 1
      for (int k = tid; k < K; k += nt) {
        double dy;
        lacc += soft_pen(v_eta[k], a.lb[b_eta + k], a.ub[b_eta + k], a.TolCon, &dy);
        if (a.gf[3]) out[ol.oDF + o_w + k] = dy;
      }
    {
      const double L = block_sum256(lacc, part);
      double Lw = 0.0;
      if (a.opt[3])  // negelcbo_vbmc.m:146-151
#pragma unroll 1
        for (int k = tid; k < K; k += nt) Lw += (v_w[k] < a.WThresh) ? v_w[k] : a.WThresh;
      Lw = block_sum256(Lw, part);
      if (tid == 0) {
        sc[3] = L;
        sc[4] = Lw * a.WPen;
      }
    }
    if (a.opt[1] && a.gf[1])
#pragma unroll 1
      for (int k = tid; k < K; k += nt) {
        double acc = 0.0;
#pragma unroll 1
        for (int d = 0; d < D; ++d) acc += dls[k * D + d];  // dsigma = sum(dlnscale,1)  (:52)
        out[ol.oDF + o_sig + k] = acc;
      }
    if (a.opt[2] && a.gf[2])
#pragma unroll 1
      for (int d = tid; d < D; d += nt) {
        double acc = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) acc += dls[k * D + d];  // dlambda = sum(dlnscale,2)  (:57)
        out[ol.oDF + o_lam + d] = acc;
      }
    if (a.opt[3] && a.gf[3])
#pragma unroll 1
      for (int k = tid; k < K; k += nt) gPw[k] = a.WPen * ((v_w[k] < a.WThresh) ? 1.0 : 0.0);  // :155
  }
  __syncthreads();
  // ------------------------------------------------------------------ softmax Jacobian J_w * g
  // J_w = diag(e/es) - e e'/es^2  =>  (J_w g)_i = wsm_i (g_i - sum_l wsm_l g_l)   (gplogjoint.m:366-368)
  if (a.gf[3]) {
    double dh = 0.0, dg = 0.0, dp = 0.0;
#pragma unroll 1
    for (int l = tid; l < K; l += nt) {
      if (doH) dh += wsm[l] * gHw[l];
      if (doG) dg += wsm[l] * gGw[l];
      dp += wsm[l] * gPw[l];
    }
    dh = block_sum256(dh, part);
    dg = block_sum256(dg, part);
    dp = block_sum256(dp, part);
    if (tid == 0) {
      sc[5] = dh; sc[6] = dg; sc[7] = dp;
    }
    __syncthreads();
#pragma unroll 1
    for (int k = tid; k < K; k += nt) {
      if (doH) out[ol.oDH + o_w + k] = a.jacobian ? wsm[k] * (gHw[k] - sc[5]) : gHw[k];
      if (doG) out[ol.oDG + o_w + k] = a.jacobian ? wsm[k] * (gGw[k] - sc[6]) : gGw[k];
      if (doP) out[ol.oDF + o_w + k] += wsm[k] * (gPw[k] - sc[7]);  // negelcbo_vbmc.m:156-160
    }
  }
  __syncthreads();
  // ------------------------------------------------------------------ F, dF
  if (a.what == FIN_NEGELCBO || a.what == FIN_NEGELCBO_NOENT) {
#pragma unroll 1
    for (int i = tid; i < a.ntheta_out; i += nt)
      out[ol.oDF + i] = -out[ol.oDG + i] - out[ol.oDH + i] + out[ol.oDF + i];  // dF = -dG - dH (+ dL)
    if (tid == 0) out[ol.oF] = -sc[2] - sc[1] + sc[3] + sc[4];               // F = -G - H (+ L)
  } else if (a.what == FIN_ENTMC) {
    if (tid == 0) out[ol.oF] = sc[1];
  } else {
    if (tid == 0) out[ol.oF] = sc[2];
  }
}

// ------------------------------------------------------------------------------------------------
int launch_vp_unpack(vbmc_b200_ctx* c, bool have_theta) {
  UnpackArgs a;
  a.D = c->D; a.K = c->K; a.ntheta = c->ntheta; a.have_theta = have_theta ? 1 : 0;
  a.force_form = c->entmc_form;
  a.DPc = c->vp_cblob_dp;
  for (int i = 0; i < 4; ++i) a.opt[i] = c->opt[i];
  a.theta = c->theta_dev.d();
  a.dyn_src = (have_theta && c->philox_dyn) ? reinterpret_cast<const unsigned long long*>(c->theta_dev.d() + c->ntheta) : nullptr;
  a.base_mu = c->base_mu; a.base_sigma = c->base_sigma; a.base_lambda = c->base_lambda;
  a.base_w = c->base_w; a.base_eta = c->base_eta;
  a.vp = c->vp;
  a.cn = c->vp.cn;
  a.scratch = c->vp.scratch;
  a.want_cblob = c->entmc_form == 0 ? 1 : 0;
  static const bool v1 = getenv("VBMC_B200_UNPACK_V1") && atoi(getenv("VBMC_B200_UNPACK_V1")) != 0;
  // cost-weighted sweep schedule requested by the step being enqueued (api.cu enqueue_step)
  a.plan = nullptr; a.plan_w = nullptr;
  a.plan_tpc = a.plan_G = a.plan_c0 = a.plan_crun = 0;
  a.plan_prune = 0.0;
  a.plan_emax[0] = a.plan_emax[1] = a.plan_emax[2] = 0.0;
  c->ent_plan_active = false;
  if (c->ent_plan_req && !v1 && c->K <= 256 && c->ent_plan_req_G <= 256) {
    const int G = c->ent_plan_req_G, K = c->K;
    VB_TRY(c->ent_plan.reserve(sizeof(int) * (512 + 3 * 256 + 8)));
    a.plan = reinterpret_cast<int*>(c->ent_plan.p);
    a.plan_w = a.plan + (G + 1 + 2 * K);
    a.plan_tpc = c->ent_plan_req_tpc; a.plan_G = G; a.plan_c0 = c->ent_balance_c0; a.plan_crun = 3 * c->ent_balance_crun;
    a.plan_prune = c->entmc_prune_c;
    // largest ||eps|| among a warp's 32 draws: P(max <= x) = F(x)^32, F the chi_D distribution; its 10 / 50 / 90 % quantiles are
    // the 0.9306 / 0.9786 / 0.99671 quantiles of chi^2_D (Wilson-Hilferty with z = 1.480, 2.026, 2.718)
    const double D = c->D, zq[3] = {1.480, 2.026, 2.718};
    for (int q = 0; q < 3; ++q) {
      const double wh = 1.0 - 2.0 / (9.0 * D) + zq[q] * sqrt(2.0 / (9.0 * D));
      a.plan_emax[q] = sqrt(D * wh * wh * wh);
    }
    c->ent_plan_active = true;
  }
  c->ent_plan_req = false;
  KernelScope ks(c, "vp_unpack", c->stream);
  if (!v1) {
    a.want_cblob = 0;
    vp_unpack2_kernel<<<c->K, 128, sizeof(double) * (2 * c->K + 3 * c->D), c->stream>>>(a);
    VB_CUDA(cudaGetLastError());
    return VBMC_B200_OK;
  }
  const size_t smem = sizeof(double) * (2 * static_cast<size_t>(c->D) * c->K + 2 * c->K + c->D);
  if (smem > 48 * 1024)
    VB_CUDA(cudaFuncSetAttribute(vp_unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  vp_unpack_kernel<<<1, 256, smem, c->stream>>>(a);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

// where the reduction kernels of a step should push their values (empty: single GPU, NCCL path, or a call that does not produce all of R)
XchgDev step_push_target(vbmc_b200_ctx* c, int S_layout, bool whole_step) {
  RLayout rl;
  rl.init(c->D, c->K, S_layout);
  static const bool off = getenv("VBMC_B200_XCHG_PUSH") && atoi(getenv("VBMC_B200_XCHG_PUSH")) == 0;
  if (off || !whole_step || c->nranks <= 1 || !c->p2p_ready || rl.total > c->xdev.cap) return XchgDev{};
  return c->xdev;
}

int launch_finalize(vbmc_b200_ctx* c, int Ns, int compute_grad, int use_bnd, int jacobian, int what, cudaStream_t st, bool pushed) {
  // compute_grad: bit mask of gradient blocks (bit i = grad_flags(i+1))
  FinArgs a;
  a.D = c->D; a.K = c->K; a.S = c->gp_ready ? c->gp.S : 0; a.Ns = Ns;
  if (what == FIN_ENTMC) a.S = 0;
  int n = 0;
  for (int i = 0; i < 4; ++i) a.gf[i] = (compute_grad >> i) & 1;
  if (a.gf[0]) n += c->D * c->K;
  if (a.gf[1]) n += c->K;
  if (a.gf[2]) n += c->D;
  if (a.gf[3]) n += c->K;
  a.ntheta_out = n;
  a.jacobian = jacobian;
  a.what = what;
  a.use_bnd = use_bnd;
  a.nbnd = c->nbnd;
  for (int i = 0; i < 4; ++i) a.opt[i] = c->opt[i];
  a.TolCon = c->TolCon; a.WThresh = c->WeightThreshold; a.WPen = c->WeightPenalty;
  a.R = c->R_dev.d();
  a.Rw = c->R_dev.d();
  a.xc = XchgDev{};
  a.pushed = pushed ? 1 : 0;
  a.lb = c->bnd.d();
  a.ub = c->bnd.d() + c->nbnd;
  a.cn = c->vp.cn;
  a.vp = c->vp;
  OutLayout ol;
  ol.init(n, a.S, c->K);
  VB_TRY(c->out_dev.reserve(sizeof(double) * ol.total));
  a.out = c->out_dev.d();
  RLayout rl;
  rl.init(c->D, c->K, a.S);
  size_t smem = sizeof(double) * (4 * c->K + static_cast<size_t>(c->D) * c->K + 256 + 16 +
                                  static_cast<size_t>(c->D) * c->K + 5 * c->K + 1 + 2 * c->D);
  a.stage_R = smem + sizeof(double) * rl.total <= c->smem_optin ? 1 : 0;
  if (a.stage_R) smem += sizeof(double) * rl.total;
  if (c->nranks > 1 && c->p2p_ready && rl.total <= c->xdev.cap) a.xc = c->xdev;  // else: allreduce_R (NCCL) ran before this launch
  if (smem > 48 * 1024)
    VB_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  KernelScope ks(c, "finalize", st);
  finalize_kernel<<<1, 256, smem, st>>>(a);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

}  // namespace vb
