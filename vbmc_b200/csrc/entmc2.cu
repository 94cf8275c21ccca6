// entmc_vbmc on sm_100a, second generation of the FP64 sweep (reference: ent/entmc_vbmc.m:49-104).
//
// Same arithmetic as entmc.cu (thread = antithetic pair, warp = 32 pairs of one source component j, one D-long dot
// product serves both signs, table-based exponential, warp-uniform pruning) with the per-tile and per-group fixed costs
// taken out -- round 1 measured the component loop at 46 % of the kernel time, the rest went to table builds, block
// barriers, transposed reductions and the W column sums:
//
//  * CONTIGUOUS SCHEDULE.  The K * tpc CTA-tiles (a tile = one group of 32 pairs per warp, all of the same j) are cut
//    into gridDim.x contiguous ranges; a CTA sees at most 2-3 different j, so the component tables are built and the
//    warps' sums are combined 2-3 times per CTA instead of once per tile (10.8 times at c3).  No block barrier inside
//    a run: the warps stream through their groups independently.
//  * TABLES store v_jk = r_jk u_jk and r_jk^2, so the exponent is  -0.5||u||^2 + r^2 (-0.5||eps||^2) -+ eps.v  (two
//    multiplies fewer per pair-component) and the gradient sums use (a_k / r) e v.
//  * STAGE = thread-private.  e(+-) of (pair, position in the compacted component list) goes to stage[pos][lane]; after
//    the loop every lane weighs ITS OWN column by its 1/q(+-) and the 32 lanes are summed with a packed butterfly
//    (8 list positions per pass, 9 shuffles + 9 adds): no lane <-> row remapping, no second round when a warp keeps
//    33 components (66 % of the groups at c3), no shared-memory reads of other lanes' data.
//  * The per-group sums of log q, M_d, E_d go through the same packed butterfly into ONE register per lane and pass;
//    the W_jl sums are kept by owner lanes (k mod 32) in registers.  Nothing is written to shared memory per group
//    except the stage itself.
// Measured and NOT adopted (round 2): (a) software-pipelining the gradient sums of the previous component pair into the
// exponentials of the current one: 0.272 vs 0.246 ms at c3; (b) a lane-pair variant -- two lanes per antithetic pair, each
// owning half of the dimensions and one sign, 124 registers, 8-byte stage entries, 16 warps per CTA, parity-green: 0.306 vs
// 0.255 ms (the extra shuffles and 8-byte table loads cost more issue slots than the doubled warp count hides).
// Results are sums in a fixed order => bit-reproducible run to run.  DIRECT = true instantiates the reference's own
// subtract-then-square formulation, which the device guard (vp_unpack_kernel) selects when ||u||^2 is so large that the
// expanded form's cancellation error (~eps_mach * ||u||^2) could matter.  K <= 256 (vbmc.m:247: K <= N^(2/3) = 252 at N = 4000);
// the thread-private stage needs K * 512 B of shared memory per warp, so the warps per CTA shrink as K grows (8 up to K = 50,
// 4 at K = 100, 1 at K = 256).
#include "entmc_shared.cuh"

namespace vb {

struct Entmc2Args {
  int D, K, half;            // half = Ns/2 pairs per component
  int pair_begin, pair_end;  // this rank's shard of the pair axis (same range for every component)
  int gpc;                   // groups of 32 pairs per component
  int tpc;                   // CTA-tiles (nwarps groups) per component
  int ntiles;                // K * tpc
  int rmax;                  // partial slots per CTA (runs of equal j inside a CTA's tile range)
  const int* tstart;         // cost-weighted schedule: CTA b sweeps tiles [tstart[b], tstart[b+1]) (vp_unpack2_kernel); null: equal counts
  int need;                  // NEED_* mask
  int pstride;               // 1 + 2*D + K doubles per partial
  double prune_c;
  unsigned long long* prune_stats;
  const int* form_flag;
  const double* eps;     // [K][half][D]
  const double* mu;      // [K][D]
  const double* sigma;   // [K]
  const double* lambda;  // [D]
  const double* ck;      // [K]
  const double* ak;      // [K]
  double* partial;       // [gridDim.x][rmax][pstride]
  int off_v, off_s, off_t16, off_m, off_bar, off_warp, warp_bytes;
  int woff_eps, woff_klist, woff_stage;
  int tmem;               // 1: the stage lives in tensor memory, the warps' sums in lane-private shared-memory planes
  int woff_acc, woff_res; // (tmem) [pstride][32] lane-private accumulators, [pstride] run result
};

__constant__ double c2_t16[16] = {1.0, 1.0442737824274138, 1.0905077326652577, 1.1387886347566916, 1.189207115002721,
                                  1.241857812073484, 1.2968395546510096, 1.3542555469368927, 1.4142135623730951,
                                  1.4768261459394993, 1.5422108254079407, 1.6104903319492543, 1.681792830507429,
                                  1.7562521603732995, 1.8340080864093424, 1.9152065613971474};  // 2^(j/16)

// four independent exponentials exp(x), x <= ~0, evaluated stage by stage (four dependency chains in flight):
// 2^(m + j/16) * p(r), 16-entry table + degree-6 polynomial (|r| <= ln2/32, truncation 4e-16); below -708 -> 0
__device__ __forceinline__ void exp2_neg4(const double (&x)[4], double (&y)[4], const double* __restrict__ t16) {
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

// ---- tensor memory (TMEM, 256 KB per SM: 128 lanes x 512 columns x 32 bit) as a thread-private scratchpad ----
// A warp reaches the 32 lanes of its quadrant (warp % 4); with the 32x32b shape thread i owns lane base + i, so N consecutive
// columns are N private 32-bit words of that thread: 12-cycle loads, no bank conflicts, no shared-memory capacity.  The sweep
// parks its e(+-) there (4 columns per list position) and reads them back once 1/q is known (SASS: STTM / LDTM).
__device__ __forceinline__ void tmem_st4d(uint32_t taddr, double a0, double a1, double a2, double a3) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__double2loint(a0)),
               "r"(__double2hiint(a0)), "r"(__double2loint(a1)), "r"(__double2hiint(a1)), "r"(__double2loint(a2)), "r"(__double2hiint(a2)),
               "r"(__double2loint(a3)), "r"(__double2hiint(a3))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8d(uint32_t taddr, double (&v)[8]) {
  int r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __hiloint2double(r[2 * i + 1], r[2 * i]);
}

// Sum over the 32 lanes of 8 values per lane in 9 shuffles + 9 adds: every level halves the number of values a lane
// still carries.  Returns the total of value index  idx = 4*bit4(lane) + 2*bit3(lane) + bit2(lane)  (the same number on
// the 4 lanes that share those bits).  Additions commute, so partner lanes compute identical sums: deterministic.
__device__ __forceinline__ double reduce8(double (&v)[8], int lane) {
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double send = b4 ? v[i] : v[i + 4];
    const double keep = b4 ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = b3 ? v[i] : v[i + 2];
    const double keep = b3 ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const double send = b2 ? v[0] : v[1];
    const double keep = b2 ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}

template <int KW> struct KlistIndex { typedef unsigned char type; };
template <> struct KlistIndex<8> { typedef unsigned short type; };

template <int DP, int KW, bool DIRECT>
__device__ __forceinline__ void entmc2_body(const Entmc2Args& a, uint32_t tmem_base) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NV = 1 + 2 * DP;        // log q, M[DP], E[DP] (padded dimensions carry zeros)
  constexpr int NB = (NV + 7) / 8;      // packed-butterfly passes for them
  const int D = a.D, K = a.K;
  const int K2 = (K + 1) & ~1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nw = blockDim.x >> 5;

  double* tab_v = reinterpret_cast<double*>(smem + a.off_v);    // [K2+1][DP]  r_jk * u_jk
  double4* tab_s = reinterpret_cast<double4*>(smem + a.off_s);  // [K2+1] {r^2, -0.5||u||^2 (DIRECT: 1/r^2), ck, ak/r}
  double* t16 = reinterpret_cast<double*>(smem + a.off_t16);    // [16]
  float2* tab_m = reinterpret_cast<float2*>(smem + a.off_m);    // [K] {||u|| (rounded down), prune_c + log(ck_k/ck_j) (rounded up)}
  float* tab_r = reinterpret_cast<float*>(smem + a.off_m) + 2 * K;  // [K] r (rounded up)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + a.off_bar) + warp;
  unsigned char* wbase = smem + a.off_warp + static_cast<size_t>(warp) * a.warp_bytes;
  double* eps_s = reinterpret_cast<double*>(wbase + a.woff_eps);      // [32*D]
  using KI = typename KlistIndex<KW>::type;   // byte indices up to K = 128 (the shared-memory budget at K = 50 is counted in bytes)
  KI* klist = reinterpret_cast<KI*>(wbase + a.woff_klist);  // [K2+4]
  double2* stage = reinterpret_cast<double2*>(wbase + a.woff_stage);  // [K2][32] {e+, e-} by list position, thread-private columns
  double* wtmp = reinterpret_cast<double*>(wbase + a.woff_stage);     // [K] mailbox of the W totals / [pstride] run result (aliases stage)
  const bool TM = a.tmem != 0;
  double* accp = reinterpret_cast<double*>(wbase + a.woff_acc);       // (TM) [pstride][32]: lane-private running sums of the run
  double* wres = reinterpret_cast<double*>(wbase + (TM ? a.woff_res : a.woff_stage));   // [pstride] what the warp publishes
  // this warp's window of tensor memory: its quadrant's lanes; with more than 4 warps two warps share a quadrant, 256 columns each
  const uint32_t tbase = tmem_base + ((32u * (warp & 3)) << 16) + (blockDim.x > 128 ? 256u * (warp >> 2) : 0u);

  const bool needT = (a.need & (NEED_MU | NEED_E)) != 0;
  const bool needW = (a.need & NEED_W) != 0;

  if (lane == 0) mbar_init(bar, 1);
  if (tid < 16) t16[tid] = c2_t16[tid];
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  // this CTA's contiguous tile range: owner(t) = floor(t * G / T)
  const long long T = a.ntiles, G = gridDim.x, b = blockIdx.x;
  const int t0 = a.tstart ? a.tstart[b] : static_cast<int>((b * T + G - 1) / G);
  const int t1 = a.tstart ? a.tstart[b + 1] : static_cast<int>(((b + 1) * T + G - 1) / G);
  if (t0 >= t1) return;

  uint32_t phase = 0;
  bool tma_pending = false;
  auto issue_eps = [&](int t) -> bool {
    if (t >= t1) return false;
    const int j = t / a.tpc, g = (t - j * a.tpc) * nw + warp;
    if (g >= a.gpc) return false;
    const int p0 = a.pair_begin + g * 32;
    int np = a.pair_end - p0;
    np = np > 32 ? 32 : np;
    if (np <= 0) return false;
    const double* src = a.eps + (static_cast<size_t>(j) * a.half + p0) * D;
    return eps_stage(eps_s, src, np * D, bar, lane);
  };
  tma_pending = issue_eps(t0);

  int run = 0;
  for (int t = t0; t < t1; ++run) {
    const int j = t / a.tpc;
    const int t_end = min(t1, (j + 1) * a.tpc);
    // ---- component tables of source component j (shared by all warps) ----
    __syncthreads();  // previous run: tables and the run-result regions are free again
    {
      const double sj = a.sigma[j];
      for (int i = tid; i < (K2 + 1) * DP; i += blockDim.x) {  // rows >= K: dummy components
        const int k = i / DP, d = i - k * DP;
        double u = 0.0;
        if (d < D && k < K) u = (a.mu[j * D + d] - a.mu[k * D + d]) / (a.sigma[k] * a.lambda[d]);
        tab_v[i] = u;
      }
      __syncthreads();
      for (int k = tid; k < K2 + 1; k += blockDim.x) {
        double uu = 0.0;
        for (int d = 0; d < D; ++d) uu = fma(tab_v[k * DP + d], tab_v[k * DP + d], uu);
        if (k < K) {
          const double r = sj / a.sigma[k];
          tab_s[k] = make_double4(r * r, DIRECT ? 1.0 / (r * r) : -0.5 * uu, a.ck[k], a.ak[k] / r);
          // pruning test operands in FP32, rounded so that the test can only err towards keeping a component
          tab_m[k] = make_float2(__double2float_rd(sqrt(uu) * 0.999999), __double2float_ru(a.prune_c + log(a.ck[k]) - log(a.ck[j]) + 0.5));
          tab_r[k] = __double2float_ru(r);
        } else {
          tab_s[k] = make_double4(0.0, 0.0, 0.0, 0.0);
        }
      }
      __syncthreads();  // ||u||^2 has been read from the u rows: now turn them into v = r u
      for (int i = tid; i < K * DP; i += blockDim.x) tab_v[i] *= sj / a.sigma[i / DP];
    }
    __syncthreads();  // tables ready

    if (TM)
      for (int i = 0; i < a.pstride; ++i) accp[i * 32 + lane] = 0.0;
    double racc[NB];   // lane (lane & 3) == 0 with index idx: running total of value 8*bk + idx of [log q | M | E]
#pragma unroll
    for (int i = 0; i < NB; ++i) racc[i] = 0.0;
    double wreg[KW];   // lane l: running W_j,k for k = l + 32*m
#pragma unroll
    for (int i = 0; i < KW; ++i) wreg[i] = 0.0;
    const int idx8 = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);

#pragma unroll 1
    for (; t < t_end; ++t) {
      const int g = (t - j * a.tpc) * nw + warp;
      const int p0 = a.pair_begin + g * 32;
      int np = g < a.gpc ? a.pair_end - p0 : 0;
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
      __syncwarp();  // all lanes have consumed eps_s -> refill it for the next tile (which may belong to the next j)
      tma_pending = issue_eps(t + 1);
      if (np == 0) continue;

      double qp = 0.0, qm = 0.0, Bp = 0.0, Bm = 0.0;
      double Ap[DP], Am[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) Ap[d] = Am[d] = 0.0;
      double mhee = 0.0;  // -0.5*||eps||^2
#pragma unroll
      for (int d = 0; d < DP; ++d) mhee = fma(e[d], e[d], mhee);
      mhee *= -0.5;
      // ---- components that can matter for this warp's 32 pairs (see entmc.cu for the bound) ----
      int nact = 0;
      {
        double mx = -2.0 * mhee;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        const float emax = __double2float_ru(sqrt(mx) * 1.000001);
        const float he2 = 0.5f * emax * emax;
#pragma unroll
        for (int rr = 0; rr < KW; ++rr) {
          const int k = lane + 32 * rr;
          bool keep = false;
          if (32 * rr < K && k < K) {
            const float2 m = tab_m[k];
            const float tt = m.x - tab_r[k] * emax;
            const float bb = tt > 0.0f ? -0.5f * tt * tt : 0.0f;
            keep = !(bb + he2 + m.y < 0.0f) || a.prune_c <= 0.0;   // NaN keeps
          }
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          if (keep) klist[nact + __popc(bal & ((1u << lane) - 1u))] = static_cast<KI>(k);  // compacted, ascending
          nact += __popc(bal);
        }
        if (lane == 0) klist[nact] = static_cast<KI>(K2);  // dummy partner when the count is odd (ck = ak = 0)
        __syncwarp();
        if (a.prune_stats && lane == 0) {
          atomicAdd(a.prune_stats, static_cast<unsigned long long>(nact));
          atomicAdd(a.prune_stats + 1, static_cast<unsigned long long>(K));
        }
      }
      // ---- component loop: two components per iteration => four independent exp chains (ka, kb) x (+, -) ----
      // (a software-pipelined variant -- gradient sums of the previous pair interleaved with the exponentials of the current
      // one, rows re-read from the table -- was measured slower: 0.272 vs 0.246 ms at c3)
#pragma unroll 1
      for (int ia = 0; ia < nact; ia += 2) {
        int k, kb;   // two indices, one broadcast load
        if (sizeof(KI) == 1) {
          const unsigned kk = *reinterpret_cast<const unsigned short*>(klist + ia);
          k = kk & 0xff; kb = kk >> 8;
        } else {
          const unsigned kk = *reinterpret_cast<const unsigned*>(klist + ia);
          k = kk & 0xffff; kb = kk >> 16;
        }
        const double4 sa = tab_s[k], sb = tab_s[kb];  // {r^2, -0.5||u||^2 | 1/r^2, ck, ak/r}
        double va[DP], vb[DP];
#pragma unroll
        for (int d = 0; d < DP; d += 2) {
          const double2 a2 = *reinterpret_cast<const double2*>(tab_v + k * DP + d);
          const double2 b2 = *reinterpret_cast<const double2*>(tab_v + kb * DP + d);
          va[d] = a2.x; va[d + 1] = a2.y;
          vb[d] = b2.x; vb[d + 1] = b2.y;
        }
        double x[4], ex[4];
        if (!DIRECT) {
          double ta0 = 0.0, ta1 = 0.0, tb0 = 0.0, tb1 = 0.0;
#pragma unroll
          for (int d = 0; d < DP; d += 2) {
            ta0 = fma(e[d], va[d], ta0);
            tb0 = fma(e[d], vb[d], tb0);
            ta1 = fma(e[d + 1], va[d + 1], ta1);
            tb1 = fma(e[d + 1], vb[d + 1], tb1);
          }
          const double rta = ta0 + ta1, rtb = tb0 + tb1;                              // r (eps.u)
          const double xba = fma(sa.x, mhee, sa.y), xbb = fma(sb.x, mhee, sb.y);      // -0.5 (||u||^2 + r^2 ||eps||^2)
          x[0] = xba - rta; x[1] = xba + rta;   // -0.5||u + r eps||^2,  -0.5||u - r eps||^2
          x[2] = xbb - rtb; x[3] = xbb + rtb;
        } else {
          // the reference's own subtract-then-square (entmc_vbmc.m:62): r z(+-)_d = v_d +- r^2 eps_d, ||z||^2 = sum (.)^2 / r^2
          double da[4] = {0.0, 0.0, 0.0, 0.0}, db[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int d = 0; d < DP; d += 2) {
            const double zap0 = fma(sa.x, e[d], va[d]), zam0 = fma(-sa.x, e[d], va[d]);
            const double zbp0 = fma(sb.x, e[d], vb[d]), zbm0 = fma(-sb.x, e[d], vb[d]);
            const double zap1 = fma(sa.x, e[d + 1], va[d + 1]), zam1 = fma(-sa.x, e[d + 1], va[d + 1]);
            const double zbp1 = fma(sb.x, e[d + 1], vb[d + 1]), zbm1 = fma(-sb.x, e[d + 1], vb[d + 1]);
            da[0] = fma(zap0, zap0, da[0]); da[1] = fma(zam0, zam0, da[1]);
            db[0] = fma(zbp0, zbp0, db[0]); db[1] = fma(zbm0, zbm0, db[1]);
            da[2] = fma(zap1, zap1, da[2]); da[3] = fma(zam1, zam1, da[3]);
            db[2] = fma(zbp1, zbp1, db[2]); db[3] = fma(zbm1, zbm1, db[3]);
          }
          x[0] = -0.5 * sa.y * (da[0] + da[2]); x[1] = -0.5 * sa.y * (da[1] + da[3]);
          x[2] = -0.5 * sb.y * (db[0] + db[2]); x[3] = -0.5 * sb.y * (db[1] + db[3]);
        }
        exp2_neg4(x, ex, t16);
        if (needW) {
          if (TM) {
            tmem_st4d(tbase + 4 * ia, ex[0], ex[1], ex[2], ex[3]);
          } else {
            stage[ia * 32 + lane] = make_double2(ex[0], ex[1]);
            stage[(ia + 1) * 32 + lane] = make_double2(ex[2], ex[3]);   // row nact (odd count): dummy, never read back
          }
        }
        qp = fma(sa.z, ex[0], qp);
        qm = fma(sa.z, ex[1], qm);
        qp = fma(sb.z, ex[2], qp);
        qm = fma(sb.z, ex[3], qm);
        if (needT) {
          const double tpa = sa.w * ex[0], tma = sa.w * ex[1], tpb = sb.w * ex[2], tmb = sb.w * ex[3];   // (a_k / r) e
          Bp = fma(tpa, sa.x, Bp);   // a_k r e
          Bm = fma(tma, sa.x, Bm);
          Bp = fma(tpb, sb.x, Bp);
          Bm = fma(tmb, sb.x, Bm);
#pragma unroll
          for (int d = 0; d < DP; ++d) {
            Ap[d] = fma(tpa, va[d], Ap[d]);   // a_k e u_d
            Am[d] = fma(tma, va[d], Am[d]);
          }
#pragma unroll
          for (int d = 0; d < DP; ++d) {
            Ap[d] = fma(tpb, vb[d], Ap[d]);
            Am[d] = fma(tmb, vb[d], Am[d]);
          }
        }
      }
      const double iqp = valid ? 1.0 / qp : 0.0;
      const double iqm = valid ? 1.0 / qm : 0.0;
      const double Hs = valid ? log(qp) + log(qm) : 0.0;
      if (TM) {
        // ---- tensor-memory stage: every lane adds ITS OWN draws' terms to its private sums; the lanes meet once per run ----
        if (needW) {
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll 1
          for (int pos0 = 0; pos0 < nact; pos0 += 4) {
            double ee[8];
            tmem_ld8d(tbase + 4 * pos0, ee);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (pos0 + i < nact) {   // warp-uniform
                double* w = accp + (1 + 2 * D + klist[pos0 + i]) * 32 + lane;
                *w += fma(ee[2 * i], iqp, ee[2 * i + 1] * iqm);
              }
            }
          }
        }
        accp[lane] += Hs;
        if (needT) {
#pragma unroll
          for (int d = 0; d < DP; ++d) {
            if (d < D) {
              const double tp = fma(e[d], Bp, Ap[d]) * iqp;
              const double tm = fma(-e[d], Bm, Am[d]) * iqm;
              accp[(1 + d) * 32 + lane] += tp + tm;
              accp[(1 + D + d) * 32 + lane] += e[d] * (tp - tm);
            }
          }
        }
        __syncwarp();   // klist is rebuilt by the next group
        continue;
      }
      // ---- W_l += sum over the warp's pairs of e+_l/q+ + e-_l/q-: own column times own 1/q, then the packed butterfly ----
      if (needW) {
        __syncwarp();
        double wtot[(32 * KW + 7) / 8];   // block bk: lane (lane & 3) == 0 holds the total of list position 8*bk + idx8
#pragma unroll
        for (int bk = 0; bk < (32 * KW + 7) / 8; ++bk) {
          wtot[bk] = 0.0;
          if (8 * bk < nact) {  // warp-uniform
            double v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int pos = 8 * bk + i;
              const double2 ee = pos < nact ? stage[pos * 32 + lane] : make_double2(0.0, 0.0);
              v[i] = fma(ee.x, iqp, ee.y * iqm);
            }
            wtot[bk] = reduce8(v, lane);
          }
        }
        __syncwarp();  // every lane has read its stage column: the region becomes the mailbox (by component index)
        for (int l = lane; l < K; l += 32) wtmp[l] = 0.0;   // components this warp skipped
        __syncwarp();
        if ((lane & 3) == 0) {
#pragma unroll
          for (int bk = 0; bk < (32 * KW + 7) / 8; ++bk) {
            const int pos = 8 * bk + idx8;
            if (pos < nact) wtmp[klist[pos]] = wtot[bk];
          }
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < KW; ++m)
          if (lane + 32 * m < K) wreg[m] += wtmp[lane + 32 * m];
        __syncwarp();  // the next group's loop overwrites the stage
      }
      // ---- log q, M_d = T+ + T-, E_d = eps_d (T+ - T-) with T(+-) = (A(+-) +- eps B(+-)) / q(+-): packed butterfly into racc ----
      {
        double vals[8 * NB];
        vals[0] = Hs;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          double md = 0.0, ed = 0.0;
          if (needT) {
            const double tp = fma(e[d], Bp, Ap[d]) * iqp;
            const double tm = fma(-e[d], Bm, Am[d]) * iqm;
            md = tp + tm;
            ed = e[d] * (tp - tm);
          }
          vals[1 + d] = md;
          vals[1 + DP + d] = ed;
        }
#pragma unroll
        for (int i = NV; i < 8 * NB; ++i) vals[i] = 0.0;
#pragma unroll
        for (int bk = 0; bk < NB; ++bk) {
          if (bk == 0 || needT) {
            double v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = vals[8 * bk + i];
            racc[bk] += reduce8(v, lane);
          }
        }
      }
    }  // tiles of this run

    // ---- end of the run: warps publish [log q | M[D] | E[D] | W[K]], cross-warp sum in warp order -> partial slot ----
    __syncwarp();
    if (TM) {
#pragma unroll 1
      for (int i0 = 0; i0 < a.pstride; i0 += 8) {
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = i0 + i < a.pstride ? accp[(i0 + i) * 32 + lane] : 0.0;
        const double tot = reduce8(v, lane);
        if ((lane & 3) == 0 && i0 + idx8 < a.pstride) wres[i0 + idx8] = tot;
      }
    } else if ((lane & 3) == 0) {
#pragma unroll
      for (int bk = 0; bk < NB; ++bk) {
        const int i = 8 * bk + idx8;   // index in [log q | M[DP] | E[DP]]
        if (i == 0) wtmp[0] = racc[bk];
        else if (i < 1 + DP) { if (i - 1 < D) wtmp[1 + (i - 1)] = racc[bk]; }
        else if (i < NV) { if (i - 1 - DP < D) wtmp[1 + D + (i - 1 - DP)] = racc[bk]; }
      }
    }
    if (!TM) {
#pragma unroll
      for (int m = 0; m < KW; ++m)
        if (lane + 32 * m < K) wtmp[1 + 2 * D + lane + 32 * m] = wreg[m];
    }
    __syncthreads();
    double* slot = a.partial + (static_cast<size_t>(b) * a.rmax + run) * a.pstride;
    const int res_off = TM ? a.woff_res : a.woff_stage;
    for (int i = tid; i < a.pstride; i += blockDim.x) {
      double s = 0.0;
      for (int w = 0; w < nw; ++w) s += reinterpret_cast<const double*>(smem + a.off_warp + static_cast<size_t>(w) * a.warp_bytes + res_off)[i];
      slot[i] = s;
    }
  }
}

// one launch serves both formulations: the device guard (vp_unpack kernels) decides per step
template <int DP, int KW>
__global__ void __launch_bounds__(256, 1) entmc2_kernel(const Entmc2Args a) {
  __shared__ uint32_t s_tmem;
  if (a.tmem) {   // warp 0 allocates all 512 columns (one CTA per SM: nobody to share with), everybody reads the base address
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const uint32_t tb = a.tmem ? s_tmem : 0u;
  if (*a.form_flag == 1)
    entmc2_body<DP, KW, true>(a, tb);
  else
    entmc2_body<DP, KW, false>(a, tb);
  if (a.tmem) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
  }
}

// Run partials -> R (compact layout, common.cuh), every sum in a fixed order.
// Component j's tiles [j*tpc, (j+1)*tpc) belong to the CTAs owner(j*tpc) .. owner((j+1)*tpc - 1), owner(t) = floor(t*G/T);
// inside CTA b the run of component j has index j - first_j(b), first_j(b) = ceil(b*T/G) / tpc.
//   blocks [0, nb1): one THREAD per (j, i), i < 1 + 2D  ->  Hs[j], M[j][d], E[j][d]
//   blocks [nb1, ..): one WARP per l  ->  Wc[l] = sum_j w_j W_jl  (lanes take j = lane, lane + 32, ...; butterfly over the lanes)
// Each value is also written into every peer's exchange inbox when xc.peer is set (multi-GPU, common.cuh xchg_push).
struct Entmc2Red {
  const double* partial;
  const int* form_flag;
  int G, tpc, ntiles, rmax, pstride, D, K, nb1;
  const int* tstart;   // cost-weighted schedule (null: equal counts): tstart[G + 1], then jlo[K], jhi[K] = first / last CTA of component j
  const double* w;
  double* R;
  int oHs, oM, oE, oWc;
  XchgDev xc;
};

// `plan`: the schedule table staged in shared memory by the kernel (tstart[G + 1] | jlo[K] | jhi[K]); four partial loads in flight per
// iteration, added in CTA order (the table look-ups and the one-at-a-time loads were 74 % of the reduction kernel's samples)
__device__ __forceinline__ double entmc2_run_sum(const Entmc2Red& a, const int* plan, int j, int i) {
  if (plan) {
    const int b_lo = plan[a.G + 1 + j], b_hi = plan[a.G + 1 + a.K + j];
    double s = 0.0;
    for (int b = b_lo; b <= b_hi; b += 4) {
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int bq = b + q;
        v[q] = 0.0;
        if (bq <= b_hi) {
          const int t0 = plan[bq];
          if (t0 < plan[bq + 1])   // a CTA without tiles wrote nothing
            v[q] = a.partial[(static_cast<size_t>(bq) * a.rmax + (j - t0 / a.tpc)) * a.pstride + i];
        }
      }
      s += v[0];
      s += v[1];
      s += v[2];
      s += v[3];
    }
    return s;
  }
  const long long T = a.ntiles;
  const int b_lo = static_cast<int>((static_cast<long long>(j) * a.tpc * a.G) / T);
  const int b_hi = static_cast<int>(((static_cast<long long>(j + 1) * a.tpc - 1) * a.G) / T);
  double s = 0.0;
  for (int b = b_lo; b <= b_hi; ++b) {
    const int first_j = static_cast<int>((b * T + a.G - 1) / a.G) / a.tpc;
    s += a.partial[(static_cast<size_t>(b) * a.rmax + (j - first_j)) * a.pstride + i];
  }
  return s;
}

__global__ void __launch_bounds__(128) entmc2_reduce_kernel(const Entmc2Red a) {
  __shared__ int s_plan[256 + 1 + 2 * 256];
  const int D = a.D, K = a.K, nv = 1 + 2 * D;
  const int* plan = nullptr;
  if (a.tstart) {   // (G <= 256, K <= 256: launch_vp_unpack only builds the table within these bounds)
    for (int t = threadIdx.x; t < a.G + 1 + 2 * K; t += blockDim.x) s_plan[t] = a.tstart[t];
    plan = s_plan;
    __syncthreads();
  }
  if (static_cast<int>(blockIdx.x) < a.nb1) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= K * nv) return;
    const int j = o / nv, i = o - j * nv;
    const double s = entmc2_run_sum(a, plan, j, i);
    const int at = i == 0 ? a.oHs + j : (i < 1 + D ? a.oM + j * D + (i - 1) : a.oE + j * D + (i - 1 - D));
    a.R[at] = s;
    xchg_push(a.xc, at, s);
  } else {
    const int l = (blockIdx.x - a.nb1) * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (l >= K) return;
    double acc = 0.0;
    for (int j = lane; j < K; j += 32) acc = fma(a.w[j], entmc2_run_sum(a, plan, j, nv + l), acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      a.R[a.oWc + l] = acc;
      xchg_push(a.xc, a.oWc + l, acc);
    }
  }
  if (a.xc.peer) __threadfence_system();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline int round_up2(int x, int m) { return (x + m - 1) / m * m; }

struct Entmc2Plan {
  int DP, KW, nw, grid;
  size_t smem;
  Entmc2Args a;
};

// false: no plan (D > 24, K > 256, or not enough shared memory for even one warp)
static bool make_plan2(vbmc_b200_ctx* c, int Ns, Entmc2Plan* pl) {
  const int D = c->D, K = c->K;
  const int half = Ns / 2;
  pl->DP = entmc_pick_dp(D);
  if (pl->DP < 0 || K > 256) return false;
  pl->KW = K <= 64 ? 2 : (K <= 128 ? 4 : 8);
  Entmc2Args& a = pl->a;
  memset(&a, 0, sizeof(a));
  a.D = D; a.K = K; a.half = half;
  shard_range(half, c->nranks, c->rank, &a.pair_begin, &a.pair_end);
  const int npairs = a.pair_end - a.pair_begin;
  a.pstride = 1 + 2 * D + K;
  const int DP = pl->DP, K2 = (K + 1) & ~1;
  int off = 0;
  a.off_v = off; off += (K2 + 1) * DP * 8;
  off = round_up2(off, 32);
  a.off_s = off; off += (K2 + 1) * 32;
  a.off_t16 = off; off += 16 * 8;
  a.off_m = off; off += round_up2(K * 12, 16);
  a.off_bar = off; off += 8 * 8;
  off = round_up2(off, 16);
  a.off_warp = off;
  const int eps_bytes = round_up2(32 * D * 8, 16);
  const int klist_bytes = round_up2((pl->KW == 8 ? 2 : 1) * (K2 + 4), 16);
  int stage_bytes = K2 * 32 * 16;
  const int res_bytes = round_up2(a.pstride * 8, 16);
  if (stage_bytes < res_bytes) stage_bytes = res_bytes;
  // VBMC_B200_ENTMC_TMEM=1: stage in tensor memory (K <= 128): 4 columns per list position in the warp's window (256 columns with
  // 8 warps, 512 with <= 4); shared memory then holds lane-private sums [pstride][32] instead of the stage, so the per-group
  // butterflies (9 SHFL + 9 DADD per 8 values) become one LDS + DADD + STS per value and the lanes meet once per run.
  // Measured on B200 (same box, same build): c3 sweep 0.289 ms vs 0.275 ms, c4 1.054 vs 1.009 ms -- 5 % SLOWER: the
  // read-modify-write chain through shared memory costs more issue slots than the packed butterfly it replaces.  Bit-parity
  // with the binary128 truth holds (tests/test_gpu_parity.py runs it), so it stays as an opt-in, off by default.
  static const bool tmem_env = getenv("VBMC_B200_ENTMC_TMEM") && atoi(getenv("VBMC_B200_ENTMC_TMEM")) != 0;
  a.tmem = (tmem_env && 4 * K2 <= 512) ? 1 : 0;
  int w = 0;
  a.woff_eps = w; w += eps_bytes;
  a.woff_klist = w; w += klist_bytes;
  if (a.tmem) {
    a.woff_acc = w; w += a.pstride * 32 * 8;
    a.woff_res = w; w += res_bytes;
    a.woff_stage = a.woff_res;
  } else {
    a.woff_stage = w; w += stage_bytes;
  }
  a.warp_bytes = w;
  int nw = static_cast<int>((c->smem_optin - a.off_warp - 64) / w);
  if (nw > 8) nw = 8;
  if (a.tmem && 4 * K2 > 256 && nw > 4) nw = 4;   // a warp needs more than half of its quadrant's columns
  if (nw >= 4) nw = nw / 4 * 4;   // equal load on the 4 SM sub-partitions
  if (nw < 1) return false;
  a.gpc = (npairs + 31) / 32;
  // small problems: fewer warps per CTA so that every SM gets a tile
  while (nw > 1 && static_cast<long long>(K) * ((a.gpc + nw - 1) / nw) < 2LL * c->num_sms) nw = (nw > 4) ? nw - 4 : nw - 1;
  pl->nw = nw;
  a.tpc = (a.gpc + nw - 1) / nw;
  a.ntiles = a.tpc * K;
  pl->grid = a.ntiles < c->num_sms ? a.ntiles : c->num_sms;
  if (pl->grid > 0) {
    int per_cta = (a.ntiles + pl->grid - 1) / pl->grid;
    // cost-weighted ranges: a tile weighs between c0 and c0 + K, so a range holds at most (c0 + K) / c0 times the average count
    // (+ the share of the per-component start costs, which only shortens ranges but is counted in the total)
    if (c->ent_plan_active)
      per_cta = static_cast<int>((static_cast<long long>(per_cta) * (c->ent_balance_c0 + K) + c->ent_balance_c0 - 1) / c->ent_balance_c0) + 3 +
                static_cast<int>((static_cast<long long>(K) * c->ent_balance_crun) / (static_cast<long long>(pl->grid) * c->ent_balance_c0));
    a.rmax = (per_cta + a.tpc - 1) / a.tpc + 1;
    if (a.rmax > K + 1) a.rmax = K + 1;
  }
  pl->smem = a.off_warp + static_cast<size_t>(nw) * a.warp_bytes;
  return true;
}

bool entmc2_enabled(vbmc_b200_ctx* c);
// Shape of the sweep's schedule for the step being enqueued (api.cu asks before launching vp_unpack2_kernel, which builds the
// cost-weighted tile ranges); false: this step's sweep does not take the FP64 second-generation kernel.
bool entmc2_balance_params(vbmc_b200_ctx* c, int Ns, int* tpc, int* G) {
  if (!c->ent_balance || !entmc2_enabled(c)) return false;
  Entmc2Plan pl;
  // below two tiles per CTA a range cannot follow the weights (boundaries fall on tiles): equal counts are at least as good there
  if (!make_plan2(c, Ns, &pl) || pl.a.ntiles == 0 || pl.grid < 2 || pl.a.ntiles < 2 * pl.grid) return false;
  *tpc = pl.a.tpc;
  *G = pl.grid;
  return true;
}

bool entmc2_enabled(vbmc_b200_ctx* c) {
  static const bool off = getenv("VBMC_B200_ENTMC_V1") && atoi(getenv("VBMC_B200_ENTMC_V1")) != 0;
  // D > 16: the 4*DP accumulator registers spill in both generations; the first-generation kernel measured 5 % faster
  // at c5 (D = 20, K = 100: 16.8 vs 17.7 ms), so it keeps those shapes while it can (K <= 128)
  // (VBMC_B200_ENTMC_V2_D20=1 sends them to this kernel as well: A/B runs of the cost-weighted schedule)
  static const bool d20 = getenv("VBMC_B200_ENTMC_V2_D20") && atoi(getenv("VBMC_B200_ENTMC_V2_D20")) != 0;
  if (!off && !d20 && c->precision == 64 && entmc_pick_dp(c->D) >= 20 && c->K <= 128) return false;
  return !off && c->precision == 64;
}

template <int DP, int KW>
static int launch2(vbmc_b200_ctx* c, const Entmc2Plan& pl, cudaStream_t st) {
  auto kern = entmc2_kernel<DP, KW>;
  VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c->smem_optin) - 64));   // 16 B static
  KernelScope ks(c, "entmc", st);
  kern<<<pl.grid, pl.nw * 32, pl.smem, st>>>(pl.a);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

// expanded-form sweep of the step; returns VBMC_B200_EUNSUPPORTED (without an error message) when the shape is not covered
int launch_entmc2(vbmc_b200_ctx* c, int Ns, int need_mask, cudaStream_t st, bool* handled) {
  Entmc2Plan pl;
  *handled = false;
  if (!make_plan2(c, Ns, &pl)) return VBMC_B200_OK;
  *handled = true;
  if (pl.a.ntiles == 0) return VBMC_B200_OK;
  VB_TRY(c->ent_partial2.reserve(static_cast<size_t>(pl.grid) * pl.a.rmax * pl.a.pstride * sizeof(double)));
  Entmc2Args& a = pl.a;
  a.need = need_mask;
  a.form_flag = c->vp.form_flag;
  a.eps = c->eps.d();
  a.mu = c->vp.mu; a.sigma = c->vp.sigma; a.lambda = c->vp.lambda; a.ck = c->vp.ck; a.ak = c->vp.ak;
  a.partial = c->ent_partial2.d();
  a.tstart = c->ent_plan_active ? reinterpret_cast<const int*>(c->ent_plan.p) : nullptr;
  a.prune_c = c->entmc_prune_c;
  a.prune_stats = c->entmc_prune_stats_on ? reinterpret_cast<unsigned long long*>(c->entmc_prune_stats.p) : nullptr;
#define VB_E2(dp)                                                                                                       \
  case dp:                                                                                                              \
    return pl.KW == 2 ? launch2<dp, 2>(c, pl, st) : (pl.KW == 4 ? launch2<dp, 4>(c, pl, st) : launch2<dp, 8>(c, pl, st));
  switch (pl.DP) {
    VB_E2(2) VB_E2(4) VB_E2(6) VB_E2(8) VB_E2(10) VB_E2(12) VB_E2(16) VB_E2(20) VB_E2(24)
  }
#undef VB_E2
  VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:entmc: unsupported padded dimension %d", pl.DP);
}

int launch_entmc2_reduce(vbmc_b200_ctx* c, int Ns, int S_layout, cudaStream_t st, bool* handled) {
  Entmc2Plan pl;
  *handled = false;
  if (!make_plan2(c, Ns, &pl)) return VBMC_B200_OK;
  *handled = true;
  if (pl.a.ntiles == 0) return VBMC_B200_OK;
  RLayout rl;
  rl.init(c->D, c->K, S_layout);
  Entmc2Red r;
  r.partial = c->ent_partial2.d();
  r.tstart = c->ent_plan_active ? reinterpret_cast<const int*>(c->ent_plan.p) : nullptr;
  r.form_flag = c->vp.form_flag;
  r.G = pl.grid; r.tpc = pl.a.tpc; r.ntiles = pl.a.ntiles; r.rmax = pl.a.rmax; r.pstride = pl.a.pstride;
  r.D = c->D; r.K = c->K;
  r.nb1 = (c->K * (1 + 2 * c->D) + 127) / 128;
  r.w = c->vp.w;
  r.R = c->R_dev.d();
  r.oHs = rl.oHs; r.oM = rl.oM; r.oE = rl.oE; r.oWc = rl.oWc;
  r.xc = step_push_target(c, S_layout, S_layout > 0);
  KernelScope ks(c, "reduce", st);
  entmc2_reduce_kernel<<<r.nb1 + (c->K + 3) / 4, 128, 0, st>>>(r);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

}  // namespace vb
