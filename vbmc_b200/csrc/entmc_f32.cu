// entmc_vbmc in single precision (BASELINE config 5: FP32 compute checked against FP64 at 1e-4 rel); reference
// ent/entmc_vbmc.m:49-104.  Same work decomposition, tile partials and reductions as entmc.cu; what changes:
//   * the K-sweep of one antithetic pair runs in FP32: tables, eps (converted from the FP64 draws on load),
//     q, A, B accumulators, and the exponential is MUFU ex2 (exponents are kept in log2 units);
//   * everything that sums over draws (warp / tile / component reductions, the all-reduce, the Jacobians and
//     the penalties) stays FP64, so the error does not grow with Ns;
//   * component masses ck = w_k*nf/sigma_k^D are scaled by 1/max(ck) inside the sweep (sigma^-D at D = 20
//     leaves the FP32 range); log q gets log(max ck) back in FP64 and N_k/q is scale free;
//   * per SOURCE component j the sweep uses the expanded form ||z||^2 = ||u||^2 + r^2||eps||^2 +- 2r(eps.u)
//     only if no target k can both matter (x > -30 for some 6-sigma draw) and carry intermediates larger than
//     F32_EXPANDED_MAX; otherwise the subtraction-then-square form of the reference (:55,:62) whose FP32
//     error is proportional to |u_d||z_d| instead of ||u||^2.
// Tried and dropped (round 1, c5 on B200): one SAMPLE per lane (16 pairs per warp, half the staging, 12 warps
// per SM instead of 4).  Issue slots went from 44% to 72% busy, but the kernel executes 37% more instructions
// (the dot product is no longer shared by the antithetic pair, table loads are amortised over half as many
// pairs): 8.1 ms vs 7.8 ms for this kernel.
// The expected log-joint (gplogjoint) is NOT run in FP32: z*alpha cancels 4-5 digits at VBMC's noise levels
// (|alpha| ~ 1e4), which would break the 1e-4 bound; it is <5% of a step.
#include <type_traits>

#include "entmc_shared.cuh"

namespace vb {

constexpr float F32_EXPANDED_MAX = 1500.0f;   // 2^-24 * 0.72 * 1500 = 6.4e-5 in the log2 exponent
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Per-step tables of the FP32 sweep, computed once in FP64 and rounded once (grid = K source components):
//   U[j][k][d] = (mu_jd - mu_kd)/(sigma_k lambda_d)                       (entmc_vbmc.m:55,:62 combined)
//   S[j][k]    = {r, y, h, g}, {ck', ak', 0, 0}   with r = sigma_j/sigma_k, y = -0.5*log2(e)*||u||^2,
//                h = -0.5*log2(e)*r^2, g = log2(e)*r, ck' = ck/max(ck), ak' = ak/max(ck)
//   direct[j]  = 1 when some target k needs the subtract-then-square form (see the header comment)
//   misc       = {max ck, log max ck, 1/max ck}
struct F32Tables {
  float* U;        // [K][K2*DP]
  float4* S;       // [K][2*K2]
  int* direct;     // [K]
  double* misc;    // [3]
};

__global__ void __launch_bounds__(256) entmc_f32_tables_kernel(const EntmcArgs a, const int DP, const F32Tables t) {
  __shared__ double s_u[128 * 24];  // u_jkd in FP64 (K <= 128, DP <= 24) for the row norms
  __shared__ double s_ic;
  __shared__ int s_direct;
  const int D = a.D, K = a.K, K2 = (K + 1) & ~1, j = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  if (tid < 32) {
    double m = 0.0;
    for (int k = lane; k < K; k += 32) m = fmax(m, a.ck[k]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane == 0) {
      if (!(m > 0.0)) m = 1.0;
      s_ic = 1.0 / m;
      s_direct = 0;
      if (j == 0) {
        t.misc[0] = m;
        t.misc[1] = log(m);
        t.misc[2] = 1.0 / m;
      }
    }
  }
  float* U = t.U + static_cast<size_t>(j) * K2 * DP;
  for (int i = tid; i < K2 * DP; i += 256) {
    const int k = i / DP, d = i - k * DP;
    double u = 0.0;
    if (d < D && k < K) u = (a.mu[j * D + d] - a.mu[k * D + d]) / (a.sigma[k] * a.lambda[d]);
    s_u[i] = u;
    U[i] = static_cast<float>(u);
  }
  __syncthreads();
  const double icmax = s_ic, sj = a.sigma[j];
  // 6-sigma bound on ||eps||: decides which formulation is accurate enough, never correctness
  const double em2 = D + 6.0 * sqrt(2.0 * D), em = sqrt(em2);
  float4* S = t.S + static_cast<size_t>(j) * 2 * K2;
  int want_direct = 0;
  for (int k = tid; k < K2; k += 256) {
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    if (k < K) {
      double uu = 0.0;
      for (int d = 0; d < D; ++d) uu = fma(s_u[k * DP + d], s_u[k * DP + d], uu);
      const double r = sj / a.sigma[k];
      s0 = make_float4(static_cast<float>(r), static_cast<float>(-0.5 * 1.4426950408889634 * uu),
                       static_cast<float>(-0.5 * 1.4426950408889634 * r * r), static_cast<float>(1.4426950408889634 * r));
      // .z/.w: operands of the warp-uniform pruning test (entmc.cu), rounded towards keeping the component
      const double pc = a.prune_c > 32.0 ? 32.0 : a.prune_c;   // exp(-32) = 1e-14 of q: far below FP32 resolution
      s1 = make_float4(static_cast<float>(a.ck[k] * icmax), static_cast<float>(a.ak[k] * icmax),
                       __double2float_rd(sqrt(uu) * 0.999999), __double2float_ru(pc + log(a.ck[k]) - log(a.ck[j]) + 0.5));
      const double gap = sqrt(uu) - r * em;   // smallest ||z|| any 6-sigma draw can reach
      const bool negligible = gap > 7.75;     // exp(-gap^2/2) < 1e-13
      const bool small = (uu + r * r * em2) <= F32_EXPANDED_MAX;
      want_direct |= (negligible || small) ? 0 : 1;
    }
    S[2 * k] = s0;
    S[2 * k + 1] = s1;
  }
  if (want_direct) atomicOr(&s_direct, 1);
  __syncthreads();
  if (tid == 0) t.direct[j] = s_direct;
}

template <int DP, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) entmc_f32_kernel(const EntmcArgs a, const F32Tables tabs) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int D = a.D, K = a.K;
  const int K2 = (K + 1) & ~1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nw = blockDim.x >> 5;

  float* tab_u = reinterpret_cast<float*>(smem + a.off_u);        // [K2][DP]
  float4* tab_s = reinterpret_cast<float4*>(smem + a.off_s);      // [K2][2]: {r, y, h, g}, {ck', ak', -, -}
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + a.off_bar) + warp;
  unsigned char* wbase = smem + a.off_warp + static_cast<size_t>(warp) * a.warp_bytes;
  double* eps_s = reinterpret_cast<double*>(wbase + a.woff_eps);      // [32*D] draws as stored in HBM: doubles, or
  float* eps_sf = reinterpret_cast<float*>(wbase + a.woff_eps);       //        floats when the device generator made them
  unsigned char* klist = wbase + a.woff_klist;                        // [K2+2] components this warp scores in the current group
  float2* iq_s = reinterpret_cast<float2*>(wbase + a.woff_iq);        // [32] {1/q+, 1/q-}
  float2* stage = reinterpret_cast<float2*>(wbase + a.woff_stage);    // [K2][32] {e+, e-}, swizzled
  double* wres = reinterpret_cast<double*>(wbase + a.woff_stage);     // [pstride]   (aliases stage)
  double* red = wres + ((a.pstride + 1) & ~1);                        // [1+2D][33]  (aliases stage)

  const bool needT = (a.need & (NEED_MU | NEED_E)) != 0;
  const bool needW = (a.need & NEED_W) != 0;

  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const double lcmax = tabs.misc[1], icmax = tabs.misc[2];

  // A tile = G consecutive groups of (nwarps*32) pairs of ONE source component: tables are loaded and the
  // block-level reduction runs once per tile; per group a warp sweeps 32 pairs and folds its results into
  // per-thread accumulators.
  const int G = a.groups_per_tile, group_pairs = nw * 32;
  uint32_t phase = 0;
  int tile = blockIdx.x;
  bool tma_pending = false;
  auto group_range = [&](int t, int g, int* jj, int* pp0) -> int {
    const int j = t / a.tiles_per_comp, tt = t - j * a.tiles_per_comp;
    const int p0 = a.pair_begin + (tt * G + g) * group_pairs + warp * 32;
    int np = a.pair_end - p0;
    *jj = j; *pp0 = p0;
    return np < 0 ? 0 : (np > 32 ? 32 : np);
  };
  // stage the first non-empty group at or after (t, g) in this CTA's tile sequence (a warp's groups of a tile are
  // ordered: once one is empty the rest of that tile is empty for this warp)
  auto issue_eps = [&](int t, int g) -> bool {
    for (;;) {
      if (g >= G) { t += gridDim.x; g = 0; }
      if (t >= a.ntiles) return false;
      int j, p0;
      const int np = group_range(t, g, &j, &p0);
      if (np > 0) {
        const size_t off = (static_cast<size_t>(j) * a.half + p0) * D;
        if (a.eps_f32) return eps_stage_f32(eps_sf, reinterpret_cast<const float*>(a.eps) + off, np * D, bar, lane);
        return eps_stage(eps_s, a.eps + off, np * D, bar, lane);
      }
      t += gridDim.x;
      g = 0;
    }
  };
  tma_pending = issue_eps(tile, 0);

  for (; tile < a.ntiles; tile += gridDim.x) {
    const int j = tile / a.tiles_per_comp;
    // ---- per-component tables: this step's FP32 tables of source j, global (L2) -> shared ----
    __syncthreads();  // previous tile: tables and the wres/stage regions are free again
    {
      const float4* gu = reinterpret_cast<const float4*>(tabs.U + static_cast<size_t>(j) * K2 * DP);
      float4* su = reinterpret_cast<float4*>(tab_u);
      for (int i = tid; i < K2 * DP / 4; i += blockDim.x) su[i] = __ldg(gu + i);
      const float4* gs = tabs.S + static_cast<size_t>(j) * 2 * K2;
      for (int i = tid; i < 2 * K2; i += blockDim.x) tab_s[i] = __ldg(gs + i);
      // row K2: dummy partner of an odd number of scored components (ck = ak = 0)
      for (int i = tid; i < DP; i += blockDim.x) tab_u[K2 * DP + i] = 0.f;
      if (tid < 2) tab_s[2 * K2 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int direct = tabs.direct[j];  // tile-uniform choice of the formulation
    __syncthreads();  // tables ready

    // per-thread accumulators over the groups of this tile (<= 8 groups: FP32 is enough, the FP64 sums start below)
    double accH = 0.0;
    float accM[DP], accE[DP], wacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < DP; ++d) accM[d] = accE[d] = 0.f;
    int np_total = 0;

#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      int jj, p0;
      const int np = group_range(tile, g, &jj, &p0);
      if (np == 0) break;  // groups are ordered: nothing follows for this warp in this tile
      np_total += np;
      // ---- this thread's draw ----
      if (tma_pending) {
        mbar_wait(bar, phase);
        phase ^= 1;
      } else {
        __syncwarp();
      }
      float e[DP];
      const bool valid = lane < np;
#pragma unroll
      for (int d = 0; d < DP; ++d) e[d] = 0.f;
      if (valid) {
        // rows of D values per pair: vector loads keep the 32 row reads (stride D) free of bank conflicts
        if (a.eps_f32) {
          if (DP % 4 == 0 && (D & 3) == 0) {
            const float4* row = reinterpret_cast<const float4*>(eps_sf + lane * D);
#pragma unroll
            for (int d = 0; d < DP; d += 4)
              if (d < D) { const float4 v = row[d >> 2]; e[d] = v.x; e[d + 1] = v.y; e[d + 2] = v.z; e[d + 3] = v.w; }
          } else if ((D & 1) == 0) {
            const float2* row = reinterpret_cast<const float2*>(eps_sf + lane * D);
#pragma unroll
            for (int d = 0; d < DP; d += 2)
              if (d < D) { const float2 v = row[d >> 1]; e[d] = v.x; e[d + 1] = v.y; }
          } else {
#pragma unroll
            for (int d = 0; d < DP; ++d)
              if (d < D) e[d] = eps_sf[lane * D + d];
          }
        } else if ((D & 1) == 0) {
          const double2* row = reinterpret_cast<const double2*>(eps_s + lane * D);
#pragma unroll
          for (int d = 0; d < DP; d += 2)
            if (d < D) { const double2 v = row[d >> 1]; e[d] = static_cast<float>(v.x); e[d + 1] = static_cast<float>(v.y); }
        } else {
#pragma unroll
          for (int d = 0; d < DP; ++d)
            if (d < D) e[d] = static_cast<float>(eps_s[lane * D + d]);
        }
      }
      __syncwarp();  // all lanes have consumed eps_s -> safe to refill it
      tma_pending = issue_eps(tile, g + 1);

      float qp = 0.f, qm = 0.f, Bp = 0.f, Bm = 0.f;
      float Ap[DP], Am[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) Ap[d] = Am[d] = 0.f;
      float ee = 0.f;
#pragma unroll
      for (int d = 0; d < DP; ++d) ee = fmaf(e[d], e[d], ee);
      // components that can matter for this warp's 32 pairs (same bound as the FP64 sweep, entmc.cu)
      unsigned long long kmask0 = 0ull, kmask1 = 0ull;
      int nact = 0;
      {
        float mx = ee;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        const float emax = sqrtf(mx) * 1.00001f, he2 = 0.5f * emax * emax;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int k = lane + 32 * rr;
          bool keep = false;
          if (32 * rr < K && k < K) {
            const float4 m = tab_s[2 * k + 1];
            const float t = m.z - tab_s[2 * k].x * 1.00001f * emax;
            const float b = t > 0.f ? -0.5f * t * t : 0.f;
            keep = !(b + he2 + m.w < 0.f) || a.prune_c <= 0.0;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          if (rr < 2) kmask0 |= static_cast<unsigned long long>(bal) << (32 * (rr & 1));
          else kmask1 |= static_cast<unsigned long long>(bal) << (32 * (rr & 1));
          if (keep) klist[nact + __popc(bal & ((1u << lane) - 1u))] = static_cast<unsigned char>(k);
          nact += __popc(bal);
        }
        if (lane == 0) klist[nact] = static_cast<unsigned char>(K2);
        __syncwarp();
        if (a.prune_stats && lane == 0) {
          atomicAdd(a.prune_stats, static_cast<unsigned long long>(nact));
          atomicAdd(a.prune_stats + 1, static_cast<unsigned long long>(K));
        }
      }

      // two components per iteration, both signs; the formulation is tile-uniform
      auto sweep = [&](auto form_tag) {
        constexpr bool EXPANDED = decltype(form_tag)::value;
#pragma unroll 1
        for (int ia = 0; ia < nact; ia += 2) {
          const unsigned kk = *reinterpret_cast<const unsigned short*>(klist + ia);
          const int k = kk & 0xff, kb = kk >> 8;
          const float4 sa0 = tab_s[2 * k], sa1 = tab_s[2 * k + 1], sb0 = tab_s[2 * kb], sb1 = tab_s[2 * kb + 1];
          float ua[DP], ub[DP];
          if (DP % 4 == 0) {
#pragma unroll
            for (int d = 0; d < DP; d += 4) {
              const float4 a4 = *reinterpret_cast<const float4*>(tab_u + k * DP + d);
              const float4 b4 = *reinterpret_cast<const float4*>(tab_u + kb * DP + d);
              ua[d] = a4.x; ua[d + 1] = a4.y; ua[d + 2] = a4.z; ua[d + 3] = a4.w;
              ub[d] = b4.x; ub[d + 1] = b4.y; ub[d + 2] = b4.z; ub[d + 3] = b4.w;
            }
          } else {
#pragma unroll
            for (int d = 0; d < DP; d += 2) {
              const float2 a2 = *reinterpret_cast<const float2*>(tab_u + k * DP + d);
              const float2 b2 = *reinterpret_cast<const float2*>(tab_u + kb * DP + d);
              ua[d] = a2.x; ua[d + 1] = a2.y;
              ub[d] = b2.x; ub[d + 1] = b2.y;
            }
          }
          float x0, x1, x2, x3;
          if (EXPANDED) {
            float ta0 = 0.f, ta1 = 0.f, tb0 = 0.f, tb1 = 0.f;
#pragma unroll
            for (int d = 0; d < DP; d += 2) {
              ta0 = fmaf(e[d], ua[d], ta0);
              tb0 = fmaf(e[d], ub[d], tb0);
              ta1 = fmaf(e[d + 1], ua[d + 1], ta1);
              tb1 = fmaf(e[d + 1], ub[d + 1], tb1);
            }
            const float ga = sa0.w * (ta0 + ta1), gb = sb0.w * (tb0 + tb1);
            const float xa = fmaf(sa0.z, ee, sa0.y), xb = fmaf(sb0.z, ee, sb0.y);
            x0 = xa - ga; x1 = xa + ga;
            x2 = xb - gb; x3 = xb + gb;
          } else {
            float dap = 0.f, dam = 0.f, dbp = 0.f, dbm = 0.f;
#pragma unroll
            for (int d = 0; d < DP; ++d) {
              const float zap = fmaf(sa0.x, e[d], ua[d]), zam = fmaf(-sa0.x, e[d], ua[d]);
              const float zbp = fmaf(sb0.x, e[d], ub[d]), zbm = fmaf(-sb0.x, e[d], ub[d]);
              dap = fmaf(zap, zap, dap); dam = fmaf(zam, zam, dam);
              dbp = fmaf(zbp, zbp, dbp); dbm = fmaf(zbm, zbm, dbm);
            }
            x0 = -0.5f * LOG2E * dap; x1 = -0.5f * LOG2E * dam;
            x2 = -0.5f * LOG2E * dbp; x3 = -0.5f * LOG2E * dbm;
          }
          const float e0 = ex2_approx(x0), e1 = ex2_approx(x1), e2 = ex2_approx(x2), e3 = ex2_approx(x3);
          if (needW) {
            stage[k * 32 + (lane ^ (k & 15))] = make_float2(e0, e1);
            if (kb < K2) stage[kb * 32 + (lane ^ (kb & 15))] = make_float2(e2, e3);
          }
          qp = fmaf(sa1.x, e0, qp);
          qm = fmaf(sa1.x, e1, qm);
          qp = fmaf(sb1.x, e2, qp);
          qm = fmaf(sb1.x, e3, qm);
          if (needT) {
            const float tpa = sa1.y * e0, tma = sa1.y * e1, tpb = sb1.y * e2, tmb = sb1.y * e3;
            Bp = fmaf(tpa, sa0.x, Bp);
            Bm = fmaf(tma, sa0.x, Bm);
            Bp = fmaf(tpb, sb0.x, Bp);
            Bm = fmaf(tmb, sb0.x, Bm);
#pragma unroll
            for (int d = 0; d < DP; ++d) {
              Ap[d] = fmaf(tpa, ua[d], Ap[d]);
              Am[d] = fmaf(tma, ua[d], Am[d]);
            }
#pragma unroll
            for (int d = 0; d < DP; ++d) {
              Ap[d] = fmaf(tpb, ub[d], Ap[d]);
              Am[d] = fmaf(tmb, ub[d], Am[d]);
            }
          }
        }
      };
      if (direct)
        sweep(std::false_type{});
      else
        sweep(std::true_type{});

      const float iqp = valid ? 1.0f / qp : 0.f;
      const float iqm = valid ? 1.0f / qm : 0.f;
      // log q = log q' + log(max ck); the second part is added once per tile below (FP64)
      if (valid) accH += static_cast<double>(logf(qp)) + static_cast<double>(logf(qm));
      if (needT) {
        // T+ = (A+ + eps*B+)/q+,  T- = (A- - eps*B-)/q-;  M_d += T+ + T-,  E_d += eps_d (T+ - T-)
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          const float tp = fmaf(e[d], Bp, Ap[d]) * iqp;
          const float tm = fmaf(-e[d], Bm, Am[d]) * iqm;
          accM[d] += tp + tm;
          accE[d] = fmaf(e[d], tp - tm, accE[d]);
        }
      }
      // ---- column sums W_l += sum_p e+_l/q+ + e-_l/q-  (lanes over l, p serial; FP32 over the 64 terms) ----
      if (needW) {
        iq_s[lane] = make_float2(iqp, iqm);
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int l = lane + 32 * rr;
          if (32 * rr < K && l < K && (((rr < 2 ? kmask0 : kmask1) >> (l & 63)) & 1ull)) {  // skipped rows hold stale values
            const float2* row = stage + l * 32;
            const int x = l & 15;
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
            for (int p = 0; p < 32; ++p) {
              const float2 ev = row[p ^ x];
              const float2 iq = iq_s[p];
              acc0 = fmaf(ev.x, iq.x, acc0);
              acc1 = fmaf(ev.y, iq.y, acc1);
            }
            wacc[rr] += acc0 + acc1;
          }
        }
        __syncwarp();  // stage and iq_s are free for the next group
      }
    }

    // ---- once per tile: warp reduction in fixed order via transposed FP64 scratch red[i][33] (aliases stage) ----
    if (np_total > 0) {
      red[0 * 33 + lane] = accH;
      if (needT) {
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          if (d < D) {
            red[(1 + d) * 33 + lane] = static_cast<double>(accM[d]);
            red[(1 + D + d) * 33 + lane] = static_cast<double>(accE[d]);
          }
        }
      }
      __syncwarp();
      const int nval = needT ? 1 + 2 * D : 1;
      for (int i = lane; i < nval; i += 32) {
        const double* rr = red + i * 33;
        double sacc = 0.0;
#pragma unroll 8
        for (int p = 0; p < 32; ++p) sacc += rr[p];
        if (i == 0) sacc += 2.0 * np_total * lcmax;
        wres[i] = sacc;
      }
      if (!needT)
        for (int i = 1 + lane; i < 1 + 2 * D; i += 32) wres[i] = 0.0;
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int l = lane + 32 * rr;
        if (l < K) wres[1 + 2 * D + l] = static_cast<double>(wacc[rr]) * icmax;
      }
    } else {
      for (int i = lane; i < a.pstride; i += 32) wres[i] = 0.0;
    }
    __syncthreads();  // every warp's wres is complete
    for (int i = tid; i < a.pstride; i += blockDim.x) {
      double sacc = 0.0;
      for (int w = 0; w < nw; ++w)
        sacc += reinterpret_cast<const double*>(smem + a.off_warp + static_cast<size_t>(w) * a.warp_bytes + a.woff_stage)[i];
      a.partial[static_cast<size_t>(tile) * a.pstride + i] = sacc;
    }
  }
}

template <int DP>
static int launch_f32(vbmc_b200_ctx* c, const EntmcPlan& pl, cudaStream_t st) {
  auto kern = entmc_f32_kernel<DP, 8>;
  VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c->smem_optin)));
  const int grid = pl.ntiles < c->num_sms ? pl.ntiles : c->num_sms;
  const int K = c->K, K2 = (K + 1) & ~1;
  const size_t nU = static_cast<size_t>(K) * K2 * DP * sizeof(float), nS = static_cast<size_t>(K) * 2 * K2 * sizeof(float4);
  VB_TRY(c->ent_tables.reserve(nU + nS + 32 + sizeof(int) * K));
  F32Tables t;
  unsigned char* base = static_cast<unsigned char*>(c->ent_tables.p);
  t.U = reinterpret_cast<float*>(base);
  t.S = reinterpret_cast<float4*>(base + nU);
  t.misc = reinterpret_cast<double*>(base + nU + nS);
  t.direct = reinterpret_cast<int*>(base + nU + nS + 32);
  {
    KernelScope ks(c, "entmc_f32_tables", st);
    entmc_f32_tables_kernel<<<K, 256, 0, st>>>(pl.a, DP, t);
    VB_CUDA(cudaGetLastError());
  }
  KernelScope ks(c, "entmc_f32", st);
  kern<<<grid, pl.nw * 32, pl.smem, st>>>(pl.a, t);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

int launch_entmc_f32(vbmc_b200_ctx* c, const EntmcPlan& pl, cudaStream_t st) {
  switch (pl.DP) {
    case 2: return launch_f32<2>(c, pl, st);
    case 4: return launch_f32<4>(c, pl, st);
    case 6: return launch_f32<6>(c, pl, st);
    case 8: return launch_f32<8>(c, pl, st);
    case 10: return launch_f32<10>(c, pl, st);
    case 12: return launch_f32<12>(c, pl, st);
    case 16: return launch_f32<16>(c, pl, st);
    case 20: return launch_f32<20>(c, pl, st);
    case 24: return launch_f32<24>(c, pl, st);
  }
  VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:entmc: unsupported padded dimension %d", pl.DP);
}

}  // namespace vb
