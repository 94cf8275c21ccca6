// entmc_vbmc on sm_100a: Monte-Carlo entropy of the variational mixture and its
// reparameterisation gradient (reference: ent/entmc_vbmc.m:49-104).
//
// Work unit = antithetic PAIR (j, p): source component j, draw eps_p (D doubles) and its mirror
// -eps_p (entmc_vbmc.m:53-54).  One thread owns one pair and scores it against all K components
// in two sequential passes (sign = +1, -1); one warp owns 32 consecutive pairs of one component,
// one CTA tile = nwarps*32 pairs of one component.  Persistent CTAs (one per SM) stride over tiles.
//
// Per (pair, sign, k):  z_d = u_jkd + sign*r_jk*eps_d,  u_jkd = (mu_jd-mu_kd)/(sigma_k lambda_d),
//                       r_jk = sigma_j/sigma_k         (== (xi-mu_k)/(sigma_k lambda), :55,:62)
//   e_k  = exp(-0.5*sum_d z_d^2)                        (:63, direct exp like the reference)
//   q   += ck_k*e_k,    ck_k = w_k*nf/sigma_k^D         (:63-64)
//   T_d += (ck_k/sigma_k)*e_k*z_d                       (lsum_d*lambda_d, :77-79)
// The e_k of a warp's 32 samples are staged in shared memory (XOR-swizzled [K][32] plane) so that
// the w-gradient column sums  W_jl = sum_s e_l(x_s)/q_s  (:100) can be formed after q_s is known.
// eps tiles are staged global->shared with TMA bulk copies (cp.async.bulk + mbarrier), one
// prefetch ahead.  All reductions run in a fixed order => results are bit-reproducible.
#include "common.cuh"

namespace vb {

struct EntmcArgs {
  int D, K, half;          // half = Ns/2 pairs per component
  int pair_begin, pair_end;  // this rank's shard of the pair axis (same range for every component)
  int tiles_per_comp, pairs_per_tile, ntiles;
  int need;                // NEED_* mask
  int pstride;             // 1 + 2*D + K doubles per tile partial
  const double* eps;       // [K][half][D]
  const double* mu;        // [K][D]
  const double* sigma;     // [K]
  const double* lambda;    // [D]
  const double* ck;        // [K]  w_k*nf/sigma_k^D
  const double* ak;        // [K]  ck_k/sigma_k
  double* partial;         // [ntiles][pstride]
  // shared-memory carve-up (byte offsets, computed on the host)
  int off_u, off_r, off_ck, off_ak, off_bar, off_warp, warp_bytes;
  int woff_eps, woff_iq, woff_wsum, woff_res, woff_stage;  // offsets inside a warp region
};

// ---- PTX helpers: mbarrier + TMA bulk copy (cp.async.bulk => SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Stage the eps chunk of one warp (npairs*D doubles, contiguous in global memory) into shared
// memory.  TMA bulk copy when 16-byte aligned, plain coalesced loads otherwise (odd D tails).
// Returns true when the chunk was issued through TMA (consumer must wait on the mbarrier).
__device__ __forceinline__ bool eps_stage(double* dst, const double* src, int ndbl, uint64_t* bar, int lane) {
  const uint32_t bytes = static_cast<uint32_t>(ndbl) * 8u;
  const bool tma_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15u) == 0) && bytes > 0;
  if (tma_ok) {
    if (lane == 0) {
      mbar_expect_tx(bar, bytes);
      tma_bulk_g2s(dst, src, bytes, bar);
    }
  } else {
    for (int i = lane; i < ndbl; i += 32) dst[i] = __ldg(src + i);
  }
  return tma_ok;
}

template <int DP, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) entmc_kernel(const EntmcArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int D = a.D, K = a.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nw = blockDim.x >> 5;

  double* tab_u = reinterpret_cast<double*>(smem + a.off_u);    // [K][DP]
  double* tab_r = reinterpret_cast<double*>(smem + a.off_r);    // [K]
  double* tab_ck = reinterpret_cast<double*>(smem + a.off_ck);  // [K]
  double* tab_ak = reinterpret_cast<double*>(smem + a.off_ak);  // [K]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + a.off_bar) + warp;
  unsigned char* wbase = smem + a.off_warp + static_cast<size_t>(warp) * a.warp_bytes;
  double* eps_s = reinterpret_cast<double*>(wbase + a.woff_eps);    // [32*D]
  double* iq_s = reinterpret_cast<double*>(wbase + a.woff_iq);      // [32]
  double* wsum = reinterpret_cast<double*>(wbase + a.woff_wsum);    // [K]
  double* wres = reinterpret_cast<double*>(wbase + a.woff_res);     // [pstride]
  double* stage = reinterpret_cast<double*>(wbase + a.woff_stage);  // [K][32] (aliased by red[][33])

  const bool needT = (a.need & (NEED_MU | NEED_E)) != 0;
  const bool needW = (a.need & NEED_W) != 0;

  if (lane == 0) mbar_init(bar, 1);
  for (int k = tid; k < K; k += blockDim.x) {
    tab_ck[k] = a.ck[k];
    tab_ak[k] = a.ak[k];
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  uint32_t phase = 0;
  // prefetch eps of the first tile
  int tile = blockIdx.x;
  bool tma_pending = false;
  auto issue_eps = [&](int t) -> bool {
    if (t >= a.ntiles) return false;
    const int j = t / a.tiles_per_comp, tt = t - j * a.tiles_per_comp;
    const int p0 = a.pair_begin + tt * a.pairs_per_tile + warp * 32;
    int np = a.pair_end - p0;
    np = np < 0 ? 0 : (np > 32 ? 32 : np);
    if (np == 0) return false;
    const double* src = a.eps + (static_cast<size_t>(j) * a.half + p0) * D;
    return eps_stage(eps_s, src, np * D, bar, lane);
  };
  tma_pending = issue_eps(tile);

  for (; tile < a.ntiles; tile += gridDim.x) {
    const int j = tile / a.tiles_per_comp, tt = tile - j * a.tiles_per_comp;
    const int p0 = a.pair_begin + tt * a.pairs_per_tile + warp * 32;
    int np = a.pair_end - p0;
    np = np < 0 ? 0 : (np > 32 ? 32 : np);

    // ---- per-component tables (shared by all warps of the CTA) ----
    __syncthreads();  // previous tile finished with the tables and with wres
    {
      const double sj = a.sigma[j];
      for (int i = tid; i < K * DP; i += blockDim.x) {
        const int k = i / DP, d = i - k * DP;
        double u = 0.0;
        if (d < D) u = (a.mu[j * D + d] - a.mu[k * D + d]) / (a.sigma[k] * a.lambda[d]);
        tab_u[i] = u;
      }
      for (int k = tid; k < K; k += blockDim.x) tab_r[k] = sj / a.sigma[k];
    }
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
    for (int d = 0; d < DP; ++d) e[d] = (valid && d < D) ? eps_s[lane * D + d] : 0.0;
    __syncwarp();  // all lanes have consumed eps_s -> safe to refill it for the next tile
    tma_pending = issue_eps(tile + gridDim.x);
    __syncthreads();  // tables ready

    double Hs = 0.0;
    double Tkeep[DP];  // (A_d + eps_d*B)/q of the +pass
#pragma unroll
    for (int d = 0; d < DP; ++d) Tkeep[d] = 0.0;
    if (needW)
      for (int l = lane; l < K; l += 32) wsum[l] = 0.0;

    if (np > 0) {
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const double sgn = pass == 0 ? 1.0 : -1.0;
        double q = 0.0, B = 0.0;
        double A[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) A[d] = 0.0;
        const int sw = lane;  // swizzled column = lane ^ (k & 15)
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
          const double r = sgn * tab_r[k];
          const double* uk = tab_u + k * DP;
          double d2a = 0.0, d2b = 0.0;
#pragma unroll
          for (int d = 0; d < DP; d += 2) {
            const double2 u2 = *reinterpret_cast<const double2*>(uk + d);
            const double z0 = fma(r, e[d], u2.x);
            const double z1 = fma(r, e[d + 1], u2.y);
            d2a = fma(z0, z0, d2a);
            d2b = fma(z1, z1, d2b);
          }
          const double ex = exp(-0.5 * (d2a + d2b));
          if (needW) stage[k * 32 + (sw ^ (k & 15))] = ex;
          q = fma(tab_ck[k], ex, q);
          if (needT) {
            const double t = tab_ak[k] * ex;
            B = fma(t, r, B);
#pragma unroll
            for (int d = 0; d < DP; d += 2) {
              const double2 u2 = *reinterpret_cast<const double2*>(uk + d);
              A[d] = fma(t, u2.x, A[d]);
              A[d + 1] = fma(t, u2.y, A[d + 1]);
            }
          }
        }
        const double iq = valid ? 1.0 / q : 0.0;
        Hs += valid ? log(q) : 0.0;
        if (needT) {
          if (pass == 0) {
#pragma unroll
            for (int d = 0; d < DP; ++d) Tkeep[d] = fma(e[d], B, A[d]) * iq;
          } else {
            // fold both signs:  M_d = T+ + T-,  E_d = eps_d (T+ - T-)   (eps of the mirror is -eps)
#pragma unroll
            for (int d = 0; d < DP; ++d) {
              const double tm = fma(e[d], B, A[d]) * iq;
              const double tp = Tkeep[d];
              Tkeep[d] = tp + tm;
              A[d] = e[d] * (tp - tm);
            }
            // A[] now holds E_d; stash into the reduction scratch below
          }
        }
        if (needW) {
          iq_s[lane] = iq;
          __syncwarp();
          for (int l = lane; l < K; l += 32) {
            const double* row = stage + l * 32;
            const int x = l & 15;
            double acc = 0.0;
#pragma unroll 8
            for (int p = 0; p < 32; ++p) acc = fma(row[p ^ x], iq_s[p], acc);
            wsum[l] += acc;
          }
          __syncwarp();  // stage / iq_s free for the next pass
        }
        if (pass == 1) {
          // ---- warp reduction in fixed order via transposed scratch red[i][33] (aliases stage) ----
          double* red = stage;
          red[0 * 33 + lane] = Hs;
          if (needT) {
#pragma unroll
            for (int d = 0; d < DP; ++d) {
              if (d < D) {
                red[(1 + d) * 33 + lane] = Tkeep[d];
                red[(1 + D + d) * 33 + lane] = A[d];
              }
            }
          }
          __syncwarp();
          const int nval = needT ? 1 + 2 * D : 1;
          for (int i = lane; i < nval; i += 32) {
            const double* rr = red + i * 33;
            double s = 0.0;
            for (int p = 0; p < 32; ++p) s += rr[p];
            wres[i] = s;
          }
          if (!needT)
            for (int i = 1 + lane; i < 1 + 2 * D; i += 32) wres[i] = 0.0;
          for (int l = lane; l < K; l += 32) wres[1 + 2 * D + l] = needW ? wsum[l] : 0.0;
          __syncwarp();
        }
      }
    } else {
      for (int i = lane; i < a.pstride; i += 32) wres[i] = 0.0;
    }
    __syncthreads();  // every warp's wres is complete
    // ---- cross-warp sum (fixed order) -> tile partial ----
    for (int i = tid; i < a.pstride; i += blockDim.x) {
      double s = 0.0;
      for (int w = 0; w < nw; ++w)
        s += reinterpret_cast<const double*>(smem + a.off_warp + static_cast<size_t>(w) * a.warp_bytes + a.woff_res)[i];
      a.partial[static_cast<size_t>(tile) * a.pstride + i] = s;
    }
  }
}

// Sum the tile partials of each component in tile order.  grid = K CTAs.
// R.Hs[j], R.M[j][d], R.E[j][d], Wj[j][l] (Wj goes to R.oWc region sized K*K)
__global__ void entmc_reduce_kernel(const double* __restrict__ partial, int tiles_per_comp, int pstride, int D, int K,
                                    double* __restrict__ Hs, double* __restrict__ M, double* __restrict__ E,
                                    double* __restrict__ Wj) {
  const int j = blockIdx.x;
  for (int i = threadIdx.x; i < pstride; i += blockDim.x) {
    double s = 0.0;
    const double* p = partial + static_cast<size_t>(j) * tiles_per_comp * pstride + i;
    for (int t = 0; t < tiles_per_comp; ++t) s += p[static_cast<size_t>(t) * pstride];
    if (i == 0)
      Hs[j] = s;
    else if (i < 1 + D)
      M[j * D + (i - 1)] = s;
    else if (i < 1 + 2 * D)
      E[j * D + (i - 1 - D)] = s;
    else
      Wj[j * K + (i - 1 - 2 * D)] = s;
  }
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

struct EntmcPlan {
  int DP, maxw, nw, pairs_per_tile, tiles_per_comp, ntiles, npairs_local, pair_begin, pair_end;
  size_t smem;
  EntmcArgs a;
};

static int make_plan(vbmc_b200_ctx* c, int Ns, EntmcPlan* pl) {
  const int D = c->D, K = c->K;
  const int half = Ns / 2;
  pl->DP = pick_dp(D);
  if (pl->DP < 0) VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:entmc: D=%d > 24 is not supported by this build", D);
  pl->maxw = pl->DP <= 12 ? 12 : 8;
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
  a.off_u = off; off += K * DP * 8;
  a.off_r = off; off += K * 8;
  a.off_ck = off; off += K * 8;
  a.off_ak = off; off += K * 8;
  a.off_bar = off; off += 16 * 8;
  off = round_up(off, 16);
  a.off_warp = off;
  // per-warp region
  int w = 0;
  a.woff_eps = w; w += round_up(32 * D * 8, 16);
  a.woff_iq = w; w += 32 * 8;
  a.woff_wsum = w; w += round_up(K * 8, 16);
  a.woff_res = w; w += round_up(a.pstride * 8, 16);
  a.woff_stage = w;
  const int stage_dbl = (K * 32 > (1 + 2 * D) * 33) ? K * 32 : (1 + 2 * D) * 33;
  w += round_up(stage_dbl * 8, 16);
  a.warp_bytes = w;
  const size_t avail = c->smem_optin;
  int nw_fit = static_cast<int>((avail - a.off_warp) / a.warp_bytes);
  if (nw_fit < 1)
    VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:entmc: K=%d, D=%d needs %d B shared memory per warp (> %zu available)", K,
            D, a.warp_bytes, avail - a.off_warp);
  int nw = nw_fit < pl->maxw ? nw_fit : pl->maxw;
  if (nw >= 4) nw = nw / 4 * 4;  // equal load on the 4 SM sub-partitions
  // small problems: prefer more, smaller tiles so that every SM gets work
  const long long total_pairs = static_cast<long long>(pl->npairs_local) * K;
  while (nw > 1 && total_pairs / (nw * 32) < 2LL * c->num_sms) nw = (nw > 4) ? nw - 4 : nw - 1;
  pl->nw = nw;
  pl->pairs_per_tile = nw * 32;
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

template <int DP, int MAXW>
static int launch_one(vbmc_b200_ctx* c, const EntmcPlan& pl, cudaStream_t st) {
  auto kern = entmc_kernel<DP, MAXW>;
  VB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(c->smem_optin)));
  const int grid = pl.ntiles < c->num_sms ? pl.ntiles : c->num_sms;
  KernelScope ks(c, "entmc", st);
  kern<<<grid, pl.nw * 32, pl.smem, st>>>(pl.a);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

int launch_entmc(vbmc_b200_ctx* c, int Ns, int need_mask, cudaStream_t st) {
  EntmcPlan pl;
  VB_TRY(make_plan(c, Ns, &pl));
  if (pl.ntiles == 0) return VBMC_B200_OK;
  VB_TRY(c->ent_partial.reserve(static_cast<size_t>(pl.ntiles) * pl.a.pstride * sizeof(double)));
  EntmcArgs& a = pl.a;
  a.need = need_mask;
  a.eps = c->eps.d();
  a.mu = c->vp.mu;
  a.sigma = c->vp.sigma;
  a.lambda = c->vp.lambda;
  a.ck = c->vp.ck;
  a.ak = c->vp.ak;
  a.partial = c->ent_partial.d();
  switch (pl.DP) {
    case 2: return launch_one<2, 12>(c, pl, st);
    case 4: return launch_one<4, 12>(c, pl, st);
    case 6: return launch_one<6, 12>(c, pl, st);
    case 8: return launch_one<8, 12>(c, pl, st);
    case 10: return launch_one<10, 12>(c, pl, st);
    case 12: return launch_one<12, 12>(c, pl, st);
    case 16: return launch_one<16, 8>(c, pl, st);
    case 20: return launch_one<20, 8>(c, pl, st);
    case 24: return launch_one<24, 8>(c, pl, st);
  }
  VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:entmc: unsupported padded dimension %d", pl.DP);
}

int launch_entmc_reduce(vbmc_b200_ctx* c, int Ns, int S_layout, cudaStream_t st) {
  EntmcPlan pl;
  VB_TRY(make_plan(c, Ns, &pl));
  RLayout rl;
  rl.init(c->D, c->K, S_layout);
  double* R = c->R_dev.d();
  if (pl.ntiles == 0) return VBMC_B200_OK;  // R was zeroed at the start of the step
  KernelScope ks(c, "reduce", st);
  entmc_reduce_kernel<<<c->K, 128, 0, st>>>(c->ent_partial.d(), pl.tiles_per_comp, pl.a.pstride, c->D, c->K,
                                            R + rl.oHs, R + rl.oM, R + rl.oE, R + rl.oWc);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

}  // namespace vb
