// Pieces shared by the FP64 (entmc.cu) and FP32 (entmc_f32.cu) entropy kernels: launch arguments,
// mbarrier/TMA helpers and the eps staging routine.
#pragma once
#include "common.cuh"

namespace vb {

struct EntmcArgs {
  int D, K, half;            // half = Ns/2 pairs per component
  int pair_begin, pair_end;  // this rank's shard of the pair axis (same range for every component)
  int tiles_per_comp, pairs_per_tile, ntiles;
  int groups_per_tile;       // groups of (nwarps*32) pairs swept per tile (pairs_per_tile = groups*nwarps*32)
  int need;                  // NEED_* mask
  int pstride;               // 1 + 2*D + K doubles per tile partial
  int iq_in_smem;            // 1: 1/q broadcast through shared memory, 0: through shuffles
  int eps_f32;               // 1: eps holds floats (FP32 mode, device generator), 0: doubles
  int stagger;               // unused (kept for layout stability of experiments)
  double prune_c;            // components below exp(-prune_c) of q for a whole warp are skipped (<= 0: keep all)
  unsigned long long* prune_stats;  // optional {kept, total} (warp, component) counters
  const int* form_flag;      // device flag: 2 -> expanded form (default), 1 -> direct, 0 -> separable (experimental)
  int c_mu, c_ck, c_akis, c_ilam;  // offsets (doubles) inside the __constant__ blob c_ent
  const double* eps;         // [K][half][D]
  const double* mu;          // [K][D]
  const double* sigma;       // [K]
  const double* lambda;      // [D]
  const double* ck;          // [K]  w_k*nf/sigma_k^D
  const double* ak;          // [K]  ck_k/sigma_k
  double* partial;           // [ntiles][pstride]
  // shared-memory carve-up (byte offsets, computed on the host)
  int off_u, off_s, off_t16, off_m, off_bar, off_warp, warp_bytes;
  int woff_eps, woff_iq, woff_stage, woff_klist;  // offsets inside a warp region
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

// same for a chunk of `nflt` floats (FP32 mode, draws from the device generator)
__device__ __forceinline__ bool eps_stage_f32(float* dst, const float* src, int nflt, uint64_t* bar, int lane) {
  const uint32_t bytes = static_cast<uint32_t>(nflt) * 4u;
  const bool tma_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15u) == 0) && bytes > 0;
  if (tma_ok) {
    if (lane == 0) {
      mbar_expect_tx(bar, bytes);
      tma_bulk_g2s(dst, src, bytes, bar);
    }
  } else {
    for (int i = lane; i < nflt; i += 32) dst[i] = __ldg(src + i);
  }
  return tma_ok;
}

struct EntmcPlan {
  int DP, maxw, nw, pairs_per_tile, tiles_per_comp, ntiles, npairs_local, pair_begin, pair_end;
  size_t smem;
  EntmcArgs a;
};

// FP32 variant of the sweep (entmc_f32.cu); same tile partials as the FP64 kernels
int launch_entmc_f32(vbmc_b200_ctx* c, const EntmcPlan& pl, cudaStream_t st);

}  // namespace vb
