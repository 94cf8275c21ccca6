// entmc_vbmc on sm_100a: Monte-Carlo entropy of the variational mixture and its
// reparameterisation gradient (reference: ent/entmc_vbmc.m:49-104).
//
// Work unit = antithetic PAIR (j, p): source component j, draw eps_p (D doubles) and its mirror
// -eps_p (entmc_vbmc.m:53-54).  One thread owns one pair and scores BOTH signs against all K
// components in one sweep (two independent dependency chains per thread); one warp owns 32
// consecutive pairs of one component, one CTA tile = nwarps*32 pairs of one component.  Persistent
// CTAs (one per SM) stride over the tiles.
//
// Per (pair, k):  z(+-)_d = u_jkd +- r_jk*eps_d,  u_jkd = (mu_jd-mu_kd)/(sigma_k lambda_d),
//                 r_jk = sigma_j/sigma_k            (== (xi-mu_k)/(sigma_k lambda), :55,:62)
//   e(+-)_k = exp(-0.5*||z(+-)||^2)                 (:63, direct exp, no max-shift, like the reference)
//   q(+-)  += ck_k*e(+-)_k,   ck_k = w_k*nf/sigma_k^D   (:63-64)
//   T(+-)_d = sum_k ak_k e(+-)_k z(+-)_d = A(+-)_d +- eps_d*B(+-),  ak = ck/sigma_k   (lsum*lambda, :77-79)
// Two formulations of ||z||^2, chosen on the device per step (flag written by vp_unpack_kernel):
//   EXPANDED : ||z(+-)||^2 = ||u||^2 + r^2||eps||^2 +- 2r(eps.u): one D-long dot product serves both
//              signs (the antithetic pair shares everything but the sign of the cross term);
//   DIRECT   : the reference's own subtraction-then-square; used when ||u||^2 is so large that the
//              expanded form's cancellation error (~eps_mach*||u||^2) could matter.
// The e(+-)_k of a warp's 32 pairs are staged in shared memory ([K][32] x {+,-}, XOR-swizzled) so
// that the w-gradient column sums W_jl = sum_s e_l(x_s)/q_s (:100) can be formed once q_s is known.
// eps tiles are staged global->shared with TMA bulk copies (cp.async.bulk + mbarrier), one tile
// ahead.  All reductions run in a fixed order => results are bit-reproducible run to run.
#include "entmc_shared.cuh"

namespace vb {

// exp(x) for x <= ~0: 2^(m + j/16) * p(r), 16-entry table in shared memory + degree-6 polynomial
// (|r| <= ln2/32, truncation 4e-16).  Arguments below -708 return 0 (denormals flushed; q always
// contains the source component's own term exp(-||eps||^2/2), so this is invisible in q, T, W).
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ t16) {
  const double L2E16 = 23.083120654223414;        // 16/ln2
  const double MAGIC = 6755399441055744.0;        // 1.5*2^52
  const double LN2_16_HI = 0.043321698783984175;  // ln2/16, low 18 bits zero
  const double LN2_16_LO = 1.0124068866660351e-12;
  const double kf = fma(x, L2E16, MAGIC);
  const int ki = __double2loint(kf);
  const double kd = kf - MAGIC;
  double r = fma(kd, -LN2_16_HI, x);
  r = fma(kd, -LN2_16_LO, r);
  const double r2 = r * r;
  const double a0 = 1.0 + r;
  const double a1 = fma(r, 1.0 / 6.0, 0.5);
  double a2 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  a2 = fma(r2, 1.0 / 720.0, a2);
  double p = fma(r2, a2, a1);
  p = fma(r2, p, a0);
  p *= t16[ki & 15];
  const int hi = __double2hiint(p) + ((ki >> 4) << 20);
  p = __hiloint2double(hi, __double2loint(p));
  return x < -708.0 ? 0.0 : p;
}

// four independent exponentials evaluated stage by stage (keeps four dependency chains in flight)
__device__ __forceinline__ void exp_neg4(const double (&x)[4], double (&y)[4], const double* __restrict__ t16) {
  const double L2E16 = 23.083120654223414, MAGIC = 6755399441055744.0;
  const double LN2_16_HI = 0.043321698783984175, LN2_16_LO = 1.0124068866660351e-12;
  double kf[4], r[4], r2[4], a2[4], p[4];
  int ki[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) kf[i] = fma(x[i], L2E16, MAGIC);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ki[i] = __double2loint(kf[i]);
    kf[i] -= MAGIC;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = fma(kf[i], -LN2_16_HI, x[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = fma(kf[i], -LN2_16_LO, r[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r2[i] = r[i] * r[i];
    a2[i] = fma(r[i], 1.0 / 120.0, 1.0 / 24.0);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a2[i] = fma(r2[i], 1.0 / 720.0, a2[i]);
    p[i] = fma(r[i], 1.0 / 6.0, 0.5);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    p[i] = fma(r2[i], a2[i], p[i]);
    a2[i] = 1.0 + r[i];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = fma(r2[i], p[i], a2[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] *= t16[ki[i] & 15];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int hi = __double2hiint(p[i]) + ((ki[i] >> 4) << 20);
    y[i] = x[i] < -708.0 ? 0.0 : __hiloint2double(hi, __double2loint(p[i]));
  }
}

__constant__ double c_t16[16] = {1.0, 1.0442737824274138, 1.0905077326652577, 1.1387886347566916, 1.189207115002721,
                                 1.241857812073484, 1.2968395546510096, 1.3542555469368927, 1.4142135623730951,
                                 1.4768261459394993, 1.5422108254079407, 1.6104903319492543, 1.681792830507429,
                                 1.7562521603732995, 1.8340080864093424, 1.9152065613971474};  // 2^(j/16)

template <int DP, int MAXW, bool EXPANDED>
__global__ void __launch_bounds__(MAXW * 32, 1) entmc_kernel(const EntmcArgs a) {
  if (*a.form_flag != (EXPANDED ? 2 : 1)) return;  // another kernel handles this step
  extern __shared__ __align__(16) unsigned char smem[];
  const int D = a.D, K = a.K;
  const int K2 = (K + 1) & ~1;  // tables are padded to an even number of components (dummy: ck = ak = 0)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nw = blockDim.x >> 5;

  double* tab_u = reinterpret_cast<double*>(smem + a.off_u);      // [K][DP]
  double4* tab_s = reinterpret_cast<double4*>(smem + a.off_s);    // [K] {r, -0.5||u||^2, ck, ak}
  double* t16 = reinterpret_cast<double*>(smem + a.off_t16);      // [16] 2^(j/16)
  float2* tab_m = reinterpret_cast<float2*>(smem + a.off_m);      // [K] {||u_jk|| (rounded down), prune_c + log(ck_k/ck_j) (rounded up)}
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + a.off_bar) + warp;
  unsigned char* wbase = smem + a.off_warp + static_cast<size_t>(warp) * a.warp_bytes;
  double* eps_s = reinterpret_cast<double*>(wbase + a.woff_eps);      // [32*D]
  unsigned char* klist = wbase + a.woff_klist;                        // [K2+2] indices of the components this warp scores
  double2* iq_s = reinterpret_cast<double2*>(wbase + a.woff_iq);      // [32] {1/q+, 1/q-} (optional)
  double2* stage = reinterpret_cast<double2*>(wbase + a.woff_stage);  // [K][32] {e+, e-}, swizzled
  double* wres = reinterpret_cast<double*>(wbase + a.woff_stage);     // [pstride]   (aliases stage)
  double* red = wres + ((a.pstride + 1) & ~1);                        // [1+2D][33]  (aliases stage)

  const bool needT = (a.need & (NEED_MU | NEED_E)) != 0;
  const bool needW = (a.need & NEED_W) != 0;

  if (lane == 0) mbar_init(bar, 1);
  if (tid < 16) t16[tid] = c_t16[tid];
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  uint32_t phase = 0;
  int tile = blockIdx.x;
  bool tma_pending = false;
  const int G = a.groups_per_tile, gpairs = nw * 32;  // a tile = G groups of (nw x 32) pairs of one component
  auto issue_eps = [&](int t, int g) -> bool {
    if (t >= a.ntiles) return false;
    const int j = t / a.tiles_per_comp, tt = t - j * a.tiles_per_comp;
    const int p0 = a.pair_begin + tt * a.pairs_per_tile + g * gpairs + warp * 32;
    int np = a.pair_end - p0;
    np = np < 0 ? 0 : (np > 32 ? 32 : np);
    if (np == 0) return false;
    const double* src = a.eps + (static_cast<size_t>(j) * a.half + p0) * D;
    return eps_stage(eps_s, src, np * D, bar, lane);
  };
  tma_pending = issue_eps(tile, 0);

  for (; tile < a.ntiles; tile += gridDim.x) {
    const int j = tile / a.tiles_per_comp, tt = tile - j * a.tiles_per_comp;

    // ---- per-component tables (shared by all warps of the CTA) ----
    __syncthreads();  // previous tile: tables and the wres/stage regions are free again
    {
      const double sj = a.sigma[j];
      for (int i = tid; i < (K2 + 1) * DP; i += blockDim.x) {  // row K2 (and K when K is odd): dummy component
        const int k = i / DP, d = i - k * DP;
        double u = 0.0;
        if (d < D && k < K) u = (a.mu[j * D + d] - a.mu[k * D + d]) / (a.sigma[k] * a.lambda[d]);
        tab_u[i] = u;
      }
      __syncthreads();
      for (int k = tid; k < K2 + 1; k += blockDim.x) {
        double uu = 0.0;
        for (int d = 0; d < D; ++d) uu = fma(tab_u[k * DP + d], tab_u[k * DP + d], uu);
        tab_s[k] = k < K ? make_double4(sj / a.sigma[k], -0.5 * uu, a.ck[k], a.ak[k]) : make_double4(0.0, 0.0, 0.0, 0.0);
        // pruning test operands in FP32, rounded so that the test can only err towards keeping a component
        if (k < K)
          tab_m[k] = make_float2(__double2float_rd(sqrt(uu) * 0.999999), __double2float_ru(a.prune_c + log(a.ck[k]) - log(a.ck[j]) + 0.5));
      }
    }
    __syncthreads();  // tables ready
    // per-lane running sums of this warp's tile result: lane i holds entries i, i+32, ... of [Hs | M[D] | E[D] | W[K]]
    double racc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
    const int p0 = a.pair_begin + tt * a.pairs_per_tile + g * gpairs + warp * 32;
    int np = a.pair_end - p0;
    np = np < 0 ? 0 : (np > 32 ? 32 : np);
    // ---- this thread's draw ----
    if (tma_pending) {
      mbar_wait(bar, phase);
      phase ^= 1;
    } else {
      __syncwarp();
    }
    double e[DP];
    const bool valid = lane < np;
#pragma unroll
    for (int d = 0; d < DP; ++d) e[d] = 0.0;
    if (valid) {
      if ((D & 1) == 0) {  // 16-byte loads: the 32 row reads (stride D doubles) stay free of bank conflicts
        const double2* row = reinterpret_cast<const double2*>(eps_s + lane * D);
#pragma unroll
        for (int d = 0; d < DP; d += 2)
          if (d < D) { const double2 v = row[d >> 1]; e[d] = v.x; e[d + 1] = v.y; }
      } else {
#pragma unroll
        for (int d = 0; d < DP; ++d)
          if (d < D) e[d] = eps_s[lane * D + d];
      }
    }
    __syncwarp();  // all lanes have consumed eps_s -> safe to refill it for the next group
    tma_pending = (g + 1 < G) ? issue_eps(tile, g + 1) : issue_eps(tile + gridDim.x, 0);

    if (np > 0) {
      double qp = 0.0, qm = 0.0, Bp = 0.0, Bm = 0.0;
      double Ap[DP], Am[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) Ap[d] = Am[d] = 0.0;
      double mhee = 0.0;  // -0.5*||eps||^2
#pragma unroll
      for (int d = 0; d < DP; ++d) mhee = fma(e[d], e[d], mhee);
      mhee *= -0.5;
      // ---- components that can matter for this warp's 32 pairs ----
      // For every pair of the warp ||eps|| <= emax, so both signs satisfy  -0.5||u_jk +- r eps||^2 <= -0.5 max(0,||u_jk|| - r emax)^2 =: b_k,
      // while q >= ck_j exp(-0.5 emax^2) (the source component's own term).  A component with
      //     log ck_k + b_k < log ck_j - 0.5 emax^2 - PRUNE_C
      // contributes less than exp(-PRUNE_C) = 2e-22 of q (and of T, W) for every pair of the warp: below FP64 round-off.  It is
      // left out (warp-uniform decision => no divergence); the reference's own exp() underflows to 0 for most of them.
      int nact = 0;
      {
        double mx = -2.0 * mhee;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        const float emax = __double2float_ru(sqrt(mx) * 1.000001);
        const float he2 = 0.5f * emax * emax;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int k = lane + 32 * rr;
          bool keep = false;
          if (32 * rr < K && k < K) {
            const float2 m = tab_m[k];
            const float t = m.x - __double2float_ru(tab_s[k].x) * emax;
            const float b = t > 0.0f ? -0.5f * t * t : 0.0f;
            keep = !(b + he2 + m.y < 0.0f) || a.prune_c <= 0.0;   // NaN keeps
          }
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          if (keep) klist[nact + __popc(bal & ((1u << lane) - 1u))] = static_cast<unsigned char>(k);  // compacted, ascending
          nact += __popc(bal);
        }
        if (lane == 0) klist[nact] = static_cast<unsigned char>(K2);  // dummy partner when the count is odd (ck = ak = 0)
        __syncwarp();
        if (a.prune_stats && lane == 0) {
          atomicAdd(a.prune_stats, static_cast<unsigned long long>(nact));
          atomicAdd(a.prune_stats + 1, static_cast<unsigned long long>(K));
        }
      }
      // two components per iteration => four independent exp chains (ka, kb) x (+, -) per thread
#pragma unroll 1
      for (int ia = 0; ia < nact; ia += 2) {
        const unsigned kk = *reinterpret_cast<const unsigned short*>(klist + ia);  // two byte indices, one broadcast load
        const int k = kk & 0xff, kb = kk >> 8;
        const double4 sa = tab_s[k], sb = tab_s[kb];  // {r, -0.5||u||^2, ck, ak}
        double ua[DP], ub[DP];
#pragma unroll
        for (int d = 0; d < DP; d += 2) {
          const double2 a2 = *reinterpret_cast<const double2*>(tab_u + k * DP + d);
          const double2 b2 = *reinterpret_cast<const double2*>(tab_u + kb * DP + d);
          ua[d] = a2.x; ua[d + 1] = a2.y;
          ub[d] = b2.x; ub[d + 1] = b2.y;
        }
        double x[4], ex[4];
        if (EXPANDED) {
          double ta0 = 0.0, ta1 = 0.0, tb0 = 0.0, tb1 = 0.0;
#pragma unroll
          for (int d = 0; d < DP; d += 2) {
            ta0 = fma(e[d], ua[d], ta0);
            tb0 = fma(e[d], ub[d], tb0);
            ta1 = fma(e[d + 1], ua[d + 1], ta1);
            tb1 = fma(e[d + 1], ub[d + 1], tb1);
          }
          const double rta = sa.x * (ta0 + ta1), rtb = sb.x * (tb0 + tb1);
          const double xba = fma(sa.x * sa.x, mhee, sa.y), xbb = fma(sb.x * sb.x, mhee, sb.y);
          x[0] = xba - rta; x[1] = xba + rta;
          x[2] = xbb - rtb; x[3] = xbb + rtb;
        } else {
          double da[4] = {0.0, 0.0, 0.0, 0.0}, db[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int d = 0; d < DP; d += 2) {
            const double zap0 = fma(sa.x, e[d], ua[d]), zam0 = fma(-sa.x, e[d], ua[d]);
            const double zbp0 = fma(sb.x, e[d], ub[d]), zbm0 = fma(-sb.x, e[d], ub[d]);
            const double zap1 = fma(sa.x, e[d + 1], ua[d + 1]), zam1 = fma(-sa.x, e[d + 1], ua[d + 1]);
            const double zbp1 = fma(sb.x, e[d + 1], ub[d + 1]), zbm1 = fma(-sb.x, e[d + 1], ub[d + 1]);
            da[0] = fma(zap0, zap0, da[0]); da[1] = fma(zam0, zam0, da[1]);
            db[0] = fma(zbp0, zbp0, db[0]); db[1] = fma(zbm0, zbm0, db[1]);
            da[2] = fma(zap1, zap1, da[2]); da[3] = fma(zam1, zam1, da[3]);
            db[2] = fma(zbp1, zbp1, db[2]); db[3] = fma(zbm1, zbm1, db[3]);
          }
          x[0] = -0.5 * (da[0] + da[2]); x[1] = -0.5 * (da[1] + da[3]);
          x[2] = -0.5 * (db[0] + db[2]); x[3] = -0.5 * (db[1] + db[3]);
        }
        exp_neg4(x, ex, t16);
        if (needW) {
          stage[k * 32 + (lane ^ (k & 7))] = make_double2(ex[0], ex[1]);
          if (kb < K2) stage[kb * 32 + (lane ^ (kb & 7))] = make_double2(ex[2], ex[3]);
        }
        qp = fma(sa.z, ex[0], qp);
        qm = fma(sa.z, ex[1], qm);
        qp = fma(sb.z, ex[2], qp);
        qm = fma(sb.z, ex[3], qm);
        if (needT) {
          const double tpa = sa.w * ex[0], tma = sa.w * ex[1], tpb = sb.w * ex[2], tmb = sb.w * ex[3];
          Bp = fma(tpa, sa.x, Bp);
          Bm = fma(tma, sa.x, Bm);
          Bp = fma(tpb, sb.x, Bp);
          Bm = fma(tmb, sb.x, Bm);
#pragma unroll
          for (int d = 0; d < DP; ++d) {
            Ap[d] = fma(tpa, ua[d], Ap[d]);
            Am[d] = fma(tma, ua[d], Am[d]);
          }
#pragma unroll
          for (int d = 0; d < DP; ++d) {
            Ap[d] = fma(tpb, ub[d], Ap[d]);
            Am[d] = fma(tmb, ub[d], Am[d]);
          }
        }
      }
      const double iqp = valid ? 1.0 / qp : 0.0;
      const double iqm = valid ? 1.0 / qm : 0.0;
      const double Hs = valid ? log(qp) + log(qm) : 0.0;
      if (needT) {
        // T+ = (A+ + eps*B+)/q+,  T- = (A- - eps*B-)/q-;  M_d = T+ + T-,  E_d = eps_d (T+ - T-)
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          const double tp = fma(e[d], Bp, Ap[d]) * iqp;
          const double tm = fma(-e[d], Bm, Am[d]) * iqm;
          Ap[d] = tp + tm;
          Am[d] = e[d] * (tp - tm);
        }
      }
      // ---- column sums W_l = sum_p e+_l/q+ + e-_l/q-  over the components this warp scored ----
      // lane i owns the rows klist[i + 32*rr]; rows rr and rr+1 share the 1/q broadcast of each pair; two partial sums per
      // row (even / odd pairs) keep two DFMA chains in flight
      double wacc[4] = {0.0, 0.0, 0.0, 0.0};
      int wrow[4] = {-1, -1, -1, -1};
      if (needW) {
        if (a.iq_in_smem) iq_s[lane] = make_double2(iqp, iqm);
        __syncwarp();
#pragma unroll
        for (int r2 = 0; r2 < 4; r2 += 2) {
          if (32 * r2 < nact) {  // warp-uniform
            const int ia = lane + 32 * r2, ib = ia + 32;
            const bool acta = ia < nact, actb = ib < nact;
            const int la = acta ? klist[ia] : 0, lb = actb ? klist[ib] : 0;
            const double2* rowa = stage + la * 32;
            const double2* rowb = stage + lb * 32;
            const int xa = la & 7, xb = lb & 7;
            const bool two = 32 * (r2 + 1) < nact;  // warp-uniform
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
            for (int p = 0; p < 32; p += 2) {
              double2 q0, q1;
              if (a.iq_in_smem) {
                q0 = iq_s[p];
                q1 = iq_s[p + 1];
              } else {
                q0 = make_double2(__shfl_sync(0xffffffffu, iqp, p), __shfl_sync(0xffffffffu, iqm, p));
                q1 = make_double2(__shfl_sync(0xffffffffu, iqp, p + 1), __shfl_sync(0xffffffffu, iqm, p + 1));
              }
              const double2 ea0 = rowa[p ^ xa], ea1 = rowa[(p + 1) ^ xa];
              a0 = fma(ea0.x, q0.x, a0);
              a1 = fma(ea1.x, q1.x, a1);
              a0 = fma(ea0.y, q0.y, a0);
              a1 = fma(ea1.y, q1.y, a1);
              if (two) {
                const double2 eb0 = rowb[p ^ xb], eb1 = rowb[(p + 1) ^ xb];
                b0 = fma(eb0.x, q0.x, b0);
                b1 = fma(eb1.x, q1.x, b1);
                b0 = fma(eb0.y, q0.y, b0);
                b1 = fma(eb1.y, q1.y, b1);
              }
            }
            if (acta) { wacc[r2] = a0 + a1; wrow[r2] = la; }
            if (actb) { wacc[r2 + 1] = b0 + b1; wrow[r2 + 1] = lb; }
          }
        }
        __syncwarp();  // stage is free: reuse it as reduction scratch
      }
      // ---- warp reduction in fixed order via transposed scratch red[i][33] ----
      red[0 * 33 + lane] = Hs;
      if (needT) {
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          if (d < D) {
            red[(1 + d) * 33 + lane] = Ap[d];
            red[(1 + D + d) * 33 + lane] = Am[d];
          }
        }
      }
      __syncwarp();
      const int nval = needT ? 1 + 2 * D : 1;
      for (int i = lane; i < nval; i += 32) {
        const double* rr = red + i * 33;
        double s = 0.0;
#pragma unroll 8
        for (int p = 0; p < 32; ++p) s += rr[p];
        wres[i] = s;
      }
      if (!needT)
        for (int i = 1 + lane; i < 1 + 2 * D; i += 32) wres[i] = 0.0;
      for (int l = lane; l < K; l += 32) wres[1 + 2 * D + l] = 0.0;  // components this warp skipped
      __syncwarp();
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
        if (wrow[rr] >= 0) wres[1 + 2 * D + wrow[rr]] = wacc[rr];
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 6; ++t)
        if (lane + 32 * t < a.pstride) racc[t] += wres[lane + 32 * t];
      __syncwarp();  // wres/red alias the stage planes the next group writes
    }
    }  // groups
#pragma unroll
    for (int t = 0; t < 6; ++t)
      if (lane + 32 * t < a.pstride) wres[lane + 32 * t] = racc[t];
    __syncthreads();  // every warp's wres is complete
    // ---- cross-warp sum (fixed order) -> tile partial ----
    for (int i = tid; i < a.pstride; i += blockDim.x) {
      double s = 0.0;
      for (int w = 0; w < nw; ++w)
        s += reinterpret_cast<const double*>(smem + a.off_warp + static_cast<size_t>(w) * a.warp_bytes + a.woff_stage)[i];
      a.partial[static_cast<size_t>(tile) * a.pstride + i] = s;
    }
  }
}

// Sum the tile partials of each component in tile order.  grid = K CTAs.
// R.Hs[j], R.M[j][d], R.E[j][d] (also pushed to the peers' inboxes, common.cuh xchg_push) and the un-contracted column sums
// Wfull[j][l] (scratch behind R); wc_contract_kernel then forms R.Wc[l] = sum_j w_j Wfull[j][l].
__global__ void entmc_reduce_kernel(const double* __restrict__ partial, int tiles_per_comp, int pstride, int D, int K,
                                    double* __restrict__ R, int oHs, int oM, int oE, int oWfull, XchgDev xc,
                                    const int* __restrict__ skip_if_expanded) {
  if (skip_if_expanded && *skip_if_expanded == 2) return;  // the second-generation sweep (entmc2.cu) produced this step's sums
  const int j = blockIdx.x;
  for (int i = threadIdx.x; i < pstride; i += blockDim.x) {
    double s = 0.0;
    const double* p = partial + static_cast<size_t>(j) * tiles_per_comp * pstride + i;
#pragma unroll 8
    for (int t = 0; t < tiles_per_comp; ++t) s += p[static_cast<size_t>(t) * pstride];
    if (i < 1 + 2 * D) {
      const int at = i == 0 ? oHs + j : (i < 1 + D ? oM + j * D + (i - 1) : oE + j * D + (i - 1 - D));
      R[at] = s;
      xchg_push(xc, at, s);
    } else {
      R[oWfull + j * K + (i - 1 - 2 * D)] = s;
    }
  }
  if (xc.peer) __threadfence_system();
}

// one warp per l: Wc[l] = sum_j w_j Wfull[j][l], lanes take j = lane, lane + 32, ...; fixed order
__global__ void __launch_bounds__(128) wc_contract_kernel(int K, const double* __restrict__ w, double* __restrict__ R, int oWfull, int oWc,
                                                           XchgDev xc, const int* __restrict__ skip_if_expanded) {
  if (skip_if_expanded && *skip_if_expanded == 2) return;
  const int l = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (l >= K) return;
  double acc = 0.0;
  for (int j = lane; j < K; j += 32) acc = fma(w[j], R[oWfull + j * K + l], acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    R[oWc + l] = acc;
    xchg_push(xc, oWc + l, acc);
  }
  if (xc.peer) __threadfence_system();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

static int pick_dp(int D) {
  static const int opts[] = {2, 4, 6, 8, 10, 12, 16, 20, 24};
  for (int o : opts)
    if (D <= o) return o;
  return -1;
}

int entmc_pick_dp(int D) { return pick_dp(D); }


static int make_plan(vbmc_b200_ctx* c, int Ns, EntmcPlan* pl) {
  const int D = c->D, K = c->K;
  const int half = Ns / 2;
  pl->DP = pick_dp(D);
  if (pl->DP < 0) VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:entmc: D=%d > 24 is not supported by this build", D);
  if (K > 128) VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:entmc: K=%d > 128 is not supported by this build", K);
  pl->maxw = 8;
  shard_range(half, c->nranks, c->rank, &pl->pair_begin, &pl->pair_end);
  pl->npairs_local = pl->pair_end - pl->pair_begin;
  const int DP = pl->DP;
  EntmcArgs& a = pl->a;
  memset(&a, 0, sizeof(a));
  a.D = D; a.K = K; a.half = half;
  a.pair_begin = pl->pair_begin; a.pair_end = pl->pair_end;
  a.pstride = 1 + 2 * D + K;
  // CTA-shared region
  int off = 0;
  const int K2 = (K + 1) & ~1;
  a.off_u = off; off += (K2 + 1) * DP * 8;   // + one dummy row (pads an odd number of active components)
  off = round_up(off, 32);
  a.off_s = off; off += (K2 + 1) * 32;
  a.off_t16 = off; off += 16 * 8;
  a.off_m = off; off += round_up((K2 + 1) * 8, 16);
  a.off_bar = off; off += 16 * 8;   // one mbarrier per warp (up to 16 warps)
  off = round_up(off, 16);
  a.off_warp = off;
  // per-warp region; the stage planes double as [wres | red] scratch after the column sums
  const bool f32 = c->precision == 32;
  const int ppw = 32;                                  // pairs per warp and group
  const int stage_bytes = K2 * 32 * (f32 ? 8 : 16);    // {e+, e-} per (component, pair)
  const int iq_bytes = f32 ? 32 * 8 : 32 * 16;
  const int eps_bytes = round_up(ppw * D * 8, 16);
  const int klist_bytes = round_up(K2 + 2, 16);
  const int scratch_bytes = (round_up(a.pstride, 2) + (1 + 2 * D) * 33) * 8;
  const int stage_alloc = round_up(stage_bytes > scratch_bytes ? stage_bytes : scratch_bytes, 16);
  const size_t avail = c->smem_optin;
  int best_nw = 0, best_iq = 0;
  for (int iq = 1; iq >= (f32 ? 1 : 0); --iq) {
    const int wb = eps_bytes + klist_bytes + (iq ? iq_bytes : 0) + stage_alloc;
    int nw_fit = static_cast<int>((avail - a.off_warp) / wb);
    if (nw_fit > pl->maxw) nw_fit = pl->maxw;
    if (nw_fit >= 4) nw_fit = nw_fit / 4 * 4;  // equal load on the 4 SM sub-partitions (5-7 warps measured no faster than 4)
    if (nw_fit > best_nw) {
      best_nw = nw_fit;
      best_iq = iq;
    }
  }
  if (best_nw < 1)
    VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:entmc: K=%d, D=%d needs more shared memory per warp than the %zu B available",
            K, D, avail - a.off_warp);
  a.iq_in_smem = best_iq;
  int w = 0;
  a.woff_eps = w; w += eps_bytes;
  a.woff_klist = w; w += klist_bytes;
  a.woff_iq = w; w += best_iq ? iq_bytes : 0;
  w = round_up(w, 16);
  a.woff_stage = w; w += stage_alloc;
  a.warp_bytes = w;
  int nw = best_nw;
  // small problems: prefer more, smaller tiles so that every SM gets work
  const long long total_pairs = static_cast<long long>(pl->npairs_local) * K;
  while (nw > 1 && total_pairs / (nw * ppw) < 2LL * c->num_sms) nw = (nw > 4) ? nw - 4 : nw - 1;
  pl->nw = nw;
  // FP32 sweep: several groups per tile amortise the table load and the block reduction, while keeping
  // >= ~10 tiles per SM so that the persistent CTAs stay balanced
  // several groups per tile amortise the table build, the barriers and the block reduction, while keeping >= ~10 tiles
  // per SM so that the persistent CTAs stay balanced
  int G = 1;
  {
    const long long groups = static_cast<long long>(K) * ((pl->npairs_local + nw * ppw - 1) / (nw * ppw));
    const long long g = groups / (10LL * c->num_sms);
    G = g < 1 ? 1 : (g > 8 ? 8 : static_cast<int>(g));
    if (const char* ge = getenv("VBMC_B200_ENTMC_GROUPS")) G = atoi(ge) > 0 ? atoi(ge) : G;
  }
  a.groups_per_tile = G;
  pl->pairs_per_tile = nw * ppw * G;
  pl->tiles_per_comp = (pl->npairs_local + pl->pairs_per_tile - 1) / pl->pairs_per_tile;
  pl->ntiles = pl->tiles_per_comp * K;
  pl->smem = a.off_warp + static_cast<size_t>(nw) * a.warp_bytes;
  a.tiles_per_comp = pl->tiles_per_comp;
  a.pairs_per_tile = pl->pairs_per_tile;
  a.ntiles = pl->ntiles;
  return VBMC_B200_OK;
}

int entmc_num_tiles(vbmc_b200_ctx* c, int Ns, int* tiles_per_comp, int* pairs_per_tile, int* nwarps, size_t* smem) {
  EntmcPlan pl;
  VB_TRY(make_plan(c, Ns, &pl));
  if (tiles_per_comp) *tiles_per_comp = pl.tiles_per_comp;
  if (pairs_per_tile) *pairs_per_tile = pl.pairs_per_tile;
  if (nwarps) *nwarps = pl.nw;
  if (smem) *smem = pl.smem;
  return VBMC_B200_OK;
}

template <int DP>
static int launch_one(vbmc_b200_ctx* c, const EntmcPlan& pl, cudaStream_t st, bool expanded_done) {
  auto kdir = entmc_kernel<DP, 8, false>;
  auto kexp = entmc_kernel<DP, 8, true>;
  VB_CUDA(cudaFuncSetAttribute(kdir, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c->smem_optin)));
  const int grid = pl.ntiles < c->num_sms ? pl.ntiles : c->num_sms;
  VB_CUDA(cudaFuncSetAttribute(kexp, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c->smem_optin)));
  if (!expanded_done) {
    KernelScope ks(c, "entmc", st);  // expanded form, tables in shared memory (default when the guard allows)
    kexp<<<grid, pl.nw * 32, pl.smem, st>>>(pl.a);
    VB_CUDA(cudaGetLastError());
  }
  {
    KernelScope ks(c, "entmc_direct", st);
    kdir<<<grid, pl.nw * 32, pl.smem, st>>>(pl.a);
    VB_CUDA(cudaGetLastError());
  }
  return VBMC_B200_OK;
}

bool entmc2_enabled(vbmc_b200_ctx* c);
int launch_entmc2(vbmc_b200_ctx* c, int Ns, int need_mask, cudaStream_t st, bool* handled);
int launch_entmc2_reduce(vbmc_b200_ctx* c, int Ns, int S_layout, cudaStream_t st, bool* handled);

int launch_entmc(vbmc_b200_ctx* c, int Ns, int need_mask, cudaStream_t st) {
  // FP64: the second-generation sweep (entmc2.cu, both formulations, K <= 256).  The kernels of this file serve the FP32 sweep
  // (entmc_f32.cu shares the plan and the tile partials) and VBMC_B200_ENTMC_V1=1 (first-generation FP64 kernel, A/B runs).
  if (entmc2_enabled(c)) {
    if (c->eps_f32)
      VB_FAIL(VBMC_B200_ESTATE, "entmc: the resident draws were generated in FP32 mode; upload or regenerate them for the FP64 sweep");
    bool v2 = false;
    VB_TRY(launch_entmc2(c, Ns, need_mask, st, &v2));
    if (v2) return VBMC_B200_OK;
    VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:entmc: D=%d, K=%d is not supported by this build (D <= 24, K <= 256)", c->D, c->K);
  }
  EntmcPlan pl;
  VB_TRY(make_plan(c, Ns, &pl));
  if (pl.ntiles == 0) return VBMC_B200_OK;
  VB_TRY(c->ent_partial.reserve(static_cast<size_t>(pl.ntiles) * pl.a.pstride * sizeof(double)));
  EntmcArgs& a = pl.a;
  a.need = need_mask;
  a.form_flag = c->vp.form_flag;
  a.eps = c->eps.d();
  a.mu = c->vp.mu;
  a.sigma = c->vp.sigma;
  a.lambda = c->vp.lambda;
  a.ck = c->vp.ck;
  a.ak = c->vp.ak;
  a.partial = c->ent_partial.d();
  a.eps_f32 = c->eps_f32 ? 1 : 0;
  a.stagger = 0;
  a.prune_c = c->entmc_prune_c;
  a.prune_stats = c->entmc_prune_stats_on ? reinterpret_cast<unsigned long long*>(c->entmc_prune_stats.p) : nullptr;
  if (c->precision == 32) return launch_entmc_f32(c, pl, st);
  if (c->eps_f32)
    VB_FAIL(VBMC_B200_ESTATE, "entmc: the resident draws were generated in FP32 mode; upload or regenerate them for the FP64 sweep");
  switch (pl.DP) {
    case 2: return launch_one<2>(c, pl, st, false);
    case 4: return launch_one<4>(c, pl, st, false);
    case 6: return launch_one<6>(c, pl, st, false);
    case 8: return launch_one<8>(c, pl, st, false);
    case 10: return launch_one<10>(c, pl, st, false);
    case 12: return launch_one<12>(c, pl, st, false);
    case 16: return launch_one<16>(c, pl, st, false);
    case 20: return launch_one<20>(c, pl, st, false);
    case 24: return launch_one<24>(c, pl, st, false);
  }
  VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:entmc: unsupported padded dimension %d", pl.DP);
}

int launch_entmc_reduce(vbmc_b200_ctx* c, int Ns, int S_layout, cudaStream_t st) {
  if (entmc2_enabled(c)) {
    bool v2 = false;
    VB_TRY(launch_entmc2_reduce(c, Ns, S_layout, st, &v2));
    return VBMC_B200_OK;
  }
  EntmcPlan pl;
  VB_TRY(make_plan(c, Ns, &pl));
  RLayout rl;
  rl.init(c->D, c->K, S_layout);
  double* R = c->R_dev.d();
  if (pl.ntiles == 0) return VBMC_B200_OK;  // R was zeroed at the start of the step
  const XchgDev xc = step_push_target(c, S_layout, S_layout > 0);
  KernelScope ks(c, "reduce", st);
  entmc_reduce_kernel<<<c->K, 128, 0, st>>>(c->ent_partial.d(), pl.tiles_per_comp, pl.a.pstride, c->D, c->K, R, rl.oHs, rl.oM, rl.oE,
                                            rl.oWfull, xc, nullptr);
  VB_CUDA(cudaGetLastError());
  c->launches++;
  wc_contract_kernel<<<(c->K + 3) / 4, 128, 0, st>>>(c->K, c->vp.w, R, rl.oWfull, rl.oWc, xc, nullptr);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

}  // namespace vb
