// Blocked forward substitution V = R' \ Z with many right-hand sides on the FP64 tensor path (gplite_pred.m:100-101,
// gplite_post.m:228-230; the solves behind gplite_pred and the rank-one update).
//
// Round 1 solved 8 right-hand sides per CTA and streamed the WHOLE factor through every CTA: 16 MB x T/8 CTAs x S samples
// (41 GB at Nstar = 1024, c3) -- bandwidth bound at 2.5 TFLOP/s, and a single column (rank-one update) kept one SM per
// sample busy for 6 ms.  Here the right-hand sides are a padded matrix Zp[S][Tp][Np] (Np, Tp multiples of 64) and the
// substitution is blocked by 64 rows, right-looking:
//   for b = 0 .. nb-1:   Z_b <- R_bb^-T Z_b                 trsm_diag_kernel   (64 x 64 tiles, registers + shuffles)
//                        Z_j <- Z_j - R_bj' Z_b,  j > b     trsm_update_kernel (DMMA m8n8k4, cp.async operands)
// Every R tile is read once per 64 right-hand sides, all (j, column-tile, sample) work items of a step run in parallel.
// The factors are read in place from the resident buffer (leading dimension ld >= Np, unit diagonal in the padding rows:
// garbage in padding rows / columns only ever reaches padding rows of Z, which are dropped).
#include <stdlib.h>

#include "common.cuh"

namespace vb {

constexpr int TT = 64;

struct TrsmArgs {
  int Np, ld, Tp, S;
  const double* L;      // [S][ld][ld] upper factors, column-major
  size_t Lstride;
  double* Z;            // [S][Tp][Np]
  const int* isfac;     // [S] or null: samples whose L is not a Cholesky factor are skipped
};

__device__ __forceinline__ void t_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void t_cp16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src)
               : "memory");
}

// grid (Tp/64, S).  Thread (c = tid/4, q = tid%4) owns rows q, q+4, ... of right-hand side c of its 64-column tile.
__global__ void __launch_bounds__(256) trsm_diag_kernel(const TrsmArgs a, int b) {
  __shared__ double R[TT][TT + 1];  // R[c][r] = R_bb(r, c)
  __shared__ double ird[TT];
  const int s = blockIdx.y, tid = threadIdx.x;
  if (a.isfac && !a.isfac[s]) return;
  const double* Ls = a.L + static_cast<size_t>(s) * a.Lstride;
  const int b0 = b * TT;
  for (int i = tid; i < TT * TT; i += 256) {
    const int c = i >> 6, r = i & 63;
    R[c][r] = (r <= c) ? Ls[static_cast<size_t>(b0 + c) * a.ld + b0 + r] : 0.0;
  }
  __syncthreads();
  if (tid < TT) ird[tid] = 1.0 / R[tid][tid];
  __syncthreads();
  const int c = tid >> 2, q = tid & 3;
  double* col = a.Z + (static_cast<size_t>(s) * a.Tp + blockIdx.x * TT + c) * a.Np + b0;
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = col[q + 4 * i];
  const unsigned lane = tid & 31;
#pragma unroll
  for (int p = 0; p < TT; ++p) {
    double xp = x[p >> 2] * ird[p];
    xp = __shfl_sync(0xffffffffu, xp, (lane & ~3u) | (p & 3));
    if (q == (p & 3)) x[p >> 2] = xp;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = q + 4 * i;
      if (r > p) x[i] = fma(-R[r][p], xp, x[i]);   // z_r -= R(p, r) v_p
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) col[q + 4 * i] = x[i];
}

// copy a 64(k) x 64(cols) block (column-major, leading dim ld) into dst[c*68 + k]
__device__ __forceinline__ void t_load_block(double* dst, const double* src, size_t ld, int tid) {
#pragma unroll
  for (int it = 0; it < 64 * 32 / 256; ++it) {
    const int idx = tid + it * 256;
    const int c = idx >> 5, ch = idx & 31;
    t_cp16(dst + c * 68 + ch * 2, src + static_cast<size_t>(c) * ld + ch * 2);
  }
}

// grid (nj * Tp/64, S): work item = (row block j = b+1+jj, column tile tt).  C(j rows, tt cols) -= R(b rows, j cols)' Z(b rows, tt cols).
__global__ void __launch_bounds__(256) trsm_update_kernel(const TrsmArgs a, int b) {
  extern __shared__ __align__(16) double tsm[];
  double* PA = tsm;             // PA[m*68 + k] = R(b0+k, j0+m)
  double* PB = tsm + TT * 68;   // PB[n*68 + k] = Z(b0+k, t0+n)
  const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.isfac && !a.isfac[s]) return;
  const int ntile = a.Tp / TT;
  const int jj = blockIdx.x / ntile, tt = blockIdx.x - jj * ntile;
  const int b0 = b * TT, j0 = (b + 1 + jj) * TT, t0 = tt * TT;
  const double* Ls = a.L + static_cast<size_t>(s) * a.Lstride;
  double* Zs = a.Z + static_cast<size_t>(s) * a.Tp * a.Np;
  t_load_block(PA, Ls + static_cast<size_t>(j0) * a.ld + b0, a.ld, tid);
  t_load_block(PB, Zs + static_cast<size_t>(t0) * a.Np + b0, a.Np, tid);
  asm volatile("cp.async.commit_group;" ::: "memory");
  const int wm = warp >> 1, wn = warp & 1;
  const int g4 = lane >> 2, t4 = lane & 3;
  double cold[2][4][2];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int m = j0 + wm * 16 + x * 8 + g4;
      const int n = t0 + wn * 32 + y * 8 + 2 * t4;
      const double* c0 = Zs + static_cast<size_t>(n) * a.Np + m;
      cold[x][y][0] = c0[0];
      cold[x][y][1] = c0[a.Np];
    }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  double acc[2][4][2];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
#pragma unroll 4
  for (int k = 0; k < TT; k += 4) {
    double af[2], bf[4];
#pragma unroll
    for (int x = 0; x < 2; ++x) af[x] = PA[(wm * 16 + x * 8 + g4) * 68 + k + t4];
#pragma unroll
    for (int y = 0; y < 4; ++y) bf[y] = PB[(wn * 32 + y * 8 + g4) * 68 + k + t4];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) t_dmma(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
  }
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int m = j0 + wm * 16 + x * 8 + g4;
      const int n = t0 + wn * 32 + y * 8 + 2 * t4;
      double* c0 = Zs + static_cast<size_t>(n) * a.Np + m;
      c0[0] = cold[x][y][0] - acc[x][y][0];
      c0[a.Np] = cold[x][y][1] - acc[x][y][1];
    }
}

// ---- backward substitution W = R \ Z (same blocking, bottom-up):  Z_b <- R_bb^-1 Z_b,  Z_i <- Z_i - R_ib Z_b for i < b ----
__global__ void __launch_bounds__(256) trsm_bdiag_kernel(const TrsmArgs a, int b) {
  __shared__ double R[TT][TT + 1];  // R[c][r] = R_bb(r, c), r <= c
  __shared__ double ird[TT];
  const int s = blockIdx.y, tid = threadIdx.x;
  if (a.isfac && !a.isfac[s]) return;
  const double* Ls = a.L + static_cast<size_t>(s) * a.Lstride;
  const int b0 = b * TT;
  for (int i = tid; i < TT * TT; i += 256) {
    const int c = i >> 6, r = i & 63;
    R[c][r] = (r <= c) ? Ls[static_cast<size_t>(b0 + c) * a.ld + b0 + r] : 0.0;
  }
  __syncthreads();
  if (tid < TT) ird[tid] = 1.0 / R[tid][tid];
  __syncthreads();
  const int c = tid >> 2, q = tid & 3;
  double* col = a.Z + (static_cast<size_t>(s) * a.Tp + blockIdx.x * TT + c) * a.Np + b0;
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = col[q + 4 * i];
  const unsigned lane = tid & 31;
#pragma unroll
  for (int p = TT - 1; p >= 0; --p) {
    double xp = x[p >> 2] * ird[p];
    xp = __shfl_sync(0xffffffffu, xp, (lane & ~3u) | (p & 3));
    if (q == (p & 3)) x[p >> 2] = xp;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = q + 4 * i;
      if (r < p) x[i] = fma(-R[p][r], xp, x[i]);   // z_r -= R(r, p) w_p
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) col[q + 4 * i] = x[i];
}

// grid (b * Tp/64, S): work item = (row block i < b, column tile tt).  C(i rows, tt cols) -= R(i rows, b cols) Z(b rows, tt cols).
__global__ void __launch_bounds__(256) trsm_bupdate_kernel(const TrsmArgs a, int b) {
  extern __shared__ __align__(16) double tsm[];
  double* PA = tsm;             // PA[k*68 + m] = R(i0+m, b0+k)   (column b0+k of R is contiguous over m)
  double* PB = tsm + TT * 68;   // PB[n*68 + k] = Z(b0+k, t0+n)
  const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.isfac && !a.isfac[s]) return;
  const int ntile = a.Tp / TT;
  const int ii = blockIdx.x / ntile, tt = blockIdx.x - ii * ntile;
  const int b0 = b * TT, i0 = ii * TT, t0 = tt * TT;
  const double* Ls = a.L + static_cast<size_t>(s) * a.Lstride;
  double* Zs = a.Z + static_cast<size_t>(s) * a.Tp * a.Np;
  t_load_block(PA, Ls + static_cast<size_t>(b0) * a.ld + i0, a.ld, tid);
  t_load_block(PB, Zs + static_cast<size_t>(t0) * a.Np + b0, a.Np, tid);
  asm volatile("cp.async.commit_group;" ::: "memory");
  const int wm = warp >> 1, wn = warp & 1;
  const int g4 = lane >> 2, t4 = lane & 3;
  double cold[2][4][2];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int m = i0 + wm * 16 + x * 8 + g4;
      const int n = t0 + wn * 32 + y * 8 + 2 * t4;
      const double* c0 = Zs + static_cast<size_t>(n) * a.Np + m;
      cold[x][y][0] = c0[0];
      cold[x][y][1] = c0[a.Np];
    }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  double acc[2][4][2];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
#pragma unroll 4
  for (int k = 0; k < TT; k += 4) {
    double af[2], bf[4];
#pragma unroll
    for (int x = 0; x < 2; ++x) af[x] = PA[(k + t4) * 68 + wm * 16 + x * 8 + g4];
#pragma unroll
    for (int y = 0; y < 4; ++y) bf[y] = PB[(wn * 32 + y * 8 + g4) * 68 + k + t4];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) t_dmma(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
  }
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int m = i0 + wm * 16 + x * 8 + g4;
      const int n = t0 + wn * 32 + y * 8 + 2 * t4;
      double* c0 = Zs + static_cast<size_t>(n) * a.Np + m;
      c0[0] = cold[x][y][0] - acc[x][y][0];
      c0[a.Np] = cold[x][y][1] - acc[x][y][1];
    }
}

// Z[S][T][N] <-> Zp[S][Tp][Np] (zero padding)
__global__ void trsm_pack_kernel(const double* __restrict__ Z, double* __restrict__ Zp, int N, int T, int Np, int Tp, int unpack) {
  const int s = blockIdx.z, t = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Np) return;
  double* zp = Zp + (static_cast<size_t>(s) * Tp + t) * Np + n;
  if (unpack) {
    if (t < T && n < N) const_cast<double*>(Z)[(static_cast<size_t>(s) * T + t) * N + n] = *zp;
  } else {
    *zp = (t < T && n < N) ? Z[(static_cast<size_t>(s) * T + t) * N + n] : 0.0;
  }
}

// V = R' \ Z (forward) or W = R \ Z (backward) for the resident factors, Z[S][T][N] in place.  Returns false (nothing done) when
// the blocked path does not apply.
bool run_trsm_blocked(vbmc_b200_ctx* c, int T, double* Z, const int* isfac_dev, cudaStream_t st, int* rc, bool backward) {
  *rc = VBMC_B200_OK;
  const int N = c->gp.N, S = c->gp.S, ld = c->gpLd;
  const int Np = (N + TT - 1) / TT * TT, Tp = (T + TT - 1) / TT * TT;
  if (ld < Np) return false;   // factors without padding (should not happen: gp_attach / gp_post both pad)
  const size_t zp = static_cast<size_t>(S) * Tp * Np;
  if (c->trsmWork.reserve(zp * sizeof(double)) != VBMC_B200_OK) {
    *rc = VBMC_B200_ECUDA;
    return true;
  }
  TrsmArgs a;
  a.Np = Np; a.ld = ld; a.Tp = Tp; a.S = S;
  a.L = c->gpL.d();
  a.Lstride = static_cast<size_t>(ld) * ld;
  a.Z = c->trsmWork.d();
  a.isfac = isfac_dev;
  KernelScope ks(c, "pred_trsm", st);
  const dim3 pg((Np + 255) / 256, Tp, S);
  trsm_pack_kernel<<<pg, 256, 0, st>>>(Z, a.Z, N, T, Np, Tp, 0);
  const int nb = Np / TT, ntile = Tp / TT;
  const int USM = 2 * TT * 68 * static_cast<int>(sizeof(double));
  if (cudaFuncSetAttribute(trsm_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, USM) != cudaSuccess ||
      cudaFuncSetAttribute(trsm_bupdate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, USM) != cudaSuccess) {
    set_error("trsm: cannot reserve %d bytes of shared memory", USM);
    *rc = VBMC_B200_ECUDA;
    return true;
  }
  // the 2*nb dependent launches of a direction are replayed as one CUDA graph while shapes and buffers are unchanged (a single
  // right-hand side -- the rank-one update -- is otherwise bound by launch latency: 130 launches of a few microseconds each)
  auto body = [&]() {
    if (!backward) {
      for (int b = 0; b < nb; ++b) {
        trsm_diag_kernel<<<dim3(ntile, S), 256, 0, st>>>(a, b);
        const int nj = nb - b - 1;
        if (nj > 0) trsm_update_kernel<<<dim3(nj * ntile, S), 256, USM, st>>>(a, b);
      }
    } else {
      for (int b = nb - 1; b >= 0; --b) {
        trsm_bdiag_kernel<<<dim3(ntile, S), 256, 0, st>>>(a, b);
        if (b > 0) trsm_bupdate_kernel<<<dim3(b * ntile, S), 256, USM, st>>>(a, b);
      }
    }
  };
  c->launches += 2 * nb;
  const int dir = backward ? 1 : 0;
  const std::vector<long long> key = {Np, ld, Tp, S, reinterpret_cast<long long>(a.L), reinterpret_cast<long long>(a.Z),
                                      reinterpret_cast<long long>(a.isfac)};
  bool replayed = false;
  if (c->graphs_enabled && !c->profiling) {
    if (!(c->trsm_graph[dir] && key == c->trsm_key[dir])) {
      if (c->trsm_graph[dir]) { cudaGraphExecDestroy(c->trsm_graph[dir]); c->trsm_graph[dir] = nullptr; }
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        body();
        if (cudaStreamEndCapture(st, &graph) == cudaSuccess && graph && cudaGraphInstantiate(&c->trsm_graph[dir], graph, 0) == cudaSuccess)
          c->trsm_key[dir] = key;
        else
          c->trsm_graph[dir] = nullptr;
        if (graph) cudaGraphDestroy(graph);
      }
      cudaGetLastError();
    }
    if (c->trsm_graph[dir]) replayed = cudaGraphLaunch(c->trsm_graph[dir], st) == cudaSuccess;
  }
  if (!replayed) body();
  trsm_pack_kernel<<<pg, 256, 0, st>>>(Z, a.Z, N, T, Np, Tp, 1);
  if (cudaGetLastError() != cudaSuccess) {
    set_error("trsm: kernel launch failed");
    *rc = VBMC_B200_ECUDA;
  }
  return true;
}

// ---- ONE right-hand side per sample (the rank-one update, gplite_post.m:191,228-230; gplite_pred at a single point) ----------
// The blocked sweep above is 2 nb dependent launches of (almost) empty grids for a single column: ~1 ms per direction at N = 2000.
// Here one launch does the whole sweep: P CTAs per sample (P * S <= number of SMs, all resident), block J of the vector belongs
// to CTA J mod P, which is the only one that ever writes z_J.  Per 64-block b, in sweep order:
//   owner:   stage R_bb in shared memory, warp 0 substitutes through it (lane owns entries lane, lane+32; the pivot value is
//            broadcast by shuffle, reciprocals of the diagonal are taken off the chain), store x_b in place, raise flag[s][b];
//   others:  the factor entries of the step are loaded BEFORE the flag is awaited (their addresses do not depend on x);
//   all:     z_J -= R(b, J)' x_b (forward) / R(J, b) x_b (backward) for the owned blocks not yet solved, four per pass.
// `epoch` tells launches apart (flags are never reset).  A peer that does not arrive within ~2^22 polls poisons the sample's
// vector with NaN and raises *err instead of hanging the GPU.
__device__ __forceinline__ int t_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void t_st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <bool BACKWARD>
__global__ void __launch_bounds__(1024) trsv1_kernel(int N, int ld, const double* __restrict__ L, size_t Lstride, double* Z,
                                                     const int* isfac, int* flags, int nbflag, int epoch, int* err) {
  extern __shared__ __align__(16) double vsm[];
  double (*Rk)[TT + 1] = reinterpret_cast<double (*)[TT + 1]>(vsm);   // Rk[c][r] = R_bb(r, c), r <= c; identity outside the N x N factor
  double (*part)[TT] = reinterpret_cast<double (*)[TT]>(vsm + TT * (TT + 1));   // [16][64]
  double* xk = vsm + TT * (TT + 1) + 16 * TT;                           // [64]
  double* zown = xk + TT;                                               // [ceil(nbN / P)][64]: the blocks of z this CTA owns
  __shared__ int timed_out;
  const int s = blockIdx.y, me = blockIdx.x, P = gridDim.x, tid = threadIdx.x, lane = tid & 31;
  if (isfac && !isfac[s]) return;
  const int nbN = (N + TT - 1) / TT;
  const double* Ls = L + static_cast<size_t>(s) * Lstride;
  double* z = Z + static_cast<size_t>(s) * N;
  int* fl = flags + static_cast<size_t>(s) * nbflag;
  const int blk = tid >> 8, sub = (tid >> 6) & 3, r = tid & 63;
  if (tid == 0) timed_out = 0;
  for (int i = tid; me + P * (i >> 6) < nbN; i += 1024) {
    const int g = (me + P * (i >> 6)) * TT + (i & 63);
    zown[i] = g < N ? z[g] : 0.0;
  }
  auto stage_diag = [&](int b) {
    const int b0 = b * TT;
    for (int i = tid; i < TT * TT; i += 1024) {
      const int c = i >> 6, rr = i & 63;
      Rk[c][rr] = (b0 + c < N && b0 + rr < N && rr <= c) ? Ls[static_cast<size_t>(b0 + c) * ld + b0 + rr] : (c == rr ? 1.0 : 0.0);
    }
  };
  int staged = -1;   // block whose diagonal tile sits in Rk
  __syncthreads();
  for (int step = 0; step < nbN; ++step) {
    const int b = BACKWARD ? nbN - 1 - step : step;
    const int bnext = BACKWARD ? b - 1 : b + 1;
    const int b0 = b * TT;
    // first owned block that still needs this step's update; the others follow at distance P
    int Jfirst;
    if (BACKWARD) {
      Jfirst = b == 0 ? -1 : b - 1 - ((b - 1 - me) % P + P) % P;           // largest J < b, J == me (mod P); negative: none
    } else {
      Jfirst = b + 1 + ((me - (b + 1)) % P + P) % P;                        // smallest J > b, J == me (mod P)
      if (Jfirst >= nbN) Jfirst = -1;
    }
    auto block_of = [&](int J0, int q) -> int {   // q-th block of a pass that starts at J0; -1: none
      if (J0 < 0) return -1;
      const int J = BACKWARD ? J0 - q * P : J0 + q * P;
      return (J < 0 || J >= nbN) ? -1 : J;
    };
    // thread (blk, sub, r): backward -> row r of block J, columns 16 sub .. 16 sub + 15 of block b  (coalesced over r);
    //                       forward  -> column r of block J, rows 16 sub .. 16 sub + 15 of block b    (128 contiguous bytes)
    auto load16 = [&](int J, double* rv) {
      if (BACKWARD) {
        const double* col = Ls + static_cast<size_t>(b0 + 16 * sub) * ld + J * TT + r;
#pragma unroll
        for (int q = 0; q < 16; ++q) rv[q] = __ldcs(col + static_cast<size_t>(q) * ld);
      } else {
        const double2* col = reinterpret_cast<const double2*>(Ls + static_cast<size_t>(J * TT + r) * ld + b0 + 16 * sub);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const double2 v = __ldcs(col + q);
          rv[2 * q] = v.x;
          rv[2 * q + 1] = v.y;
        }
      }
    };
    double rv[16];
    const int Jmine = block_of(Jfirst, blk);
    if (Jmine >= 0) load16(Jmine, rv);
    if (b % P == me) {
      if (staged != b) {
        stage_diag(b);
        staged = b;
        __syncthreads();
      }
      if (tid < 32) {
        const double d0 = 1.0 / Rk[lane][lane], d1 = 1.0 / Rk[lane + 32][lane + 32];
        const double* zb = zown + (b / P) * TT;
        double x0 = zb[lane], x1 = zb[lane + 32];
        if (BACKWARD) {
#pragma unroll 8
          for (int p = TT - 1; p >= 32; --p) {
            const double xp = __shfl_sync(0xffffffffu, x1 * d1, p - 32);
            if (lane == p - 32) x1 = xp;
            if (lane + 32 < p) x1 = fma(-Rk[p][lane + 32], xp, x1);
            x0 = fma(-Rk[p][lane], xp, x0);
          }
#pragma unroll 8
          for (int p = 31; p >= 0; --p) {
            const double xp = __shfl_sync(0xffffffffu, x0 * d0, p);
            if (lane == p) x0 = xp;
            if (lane < p) x0 = fma(-Rk[p][lane], xp, x0);
          }
        } else {
#pragma unroll 8
          for (int p = 0; p < 32; ++p) {
            const double xp = __shfl_sync(0xffffffffu, x0 * d0, p);
            if (lane == p) x0 = xp;
            if (lane > p) x0 = fma(-Rk[lane][p], xp, x0);
            x1 = fma(-Rk[lane + 32][p], xp, x1);
          }
#pragma unroll 8
          for (int p = 32; p < TT; ++p) {
            const double xp = __shfl_sync(0xffffffffu, x1 * d1, p - 32);
            if (lane == p - 32) x1 = xp;
            if (lane + 32 > p) x1 = fma(-Rk[lane + 32][p], xp, x1);
          }
        }
        xk[lane] = x0;
        xk[lane + 32] = x1;
        if (b0 + lane < N) z[b0 + lane] = x0;
        if (b0 + lane + 32 < N) z[b0 + lane + 32] = x1;
        __threadfence();
      }
      __syncthreads();
      if (tid == 0 && P > 1) t_st_release(fl + b, epoch);
    } else {
      // my turn comes next: fetch my diagonal tile while the current owner is still solving
      if (bnext >= 0 && bnext < nbN && bnext % P == me) {
        stage_diag(bnext);
        staged = bnext;
      }
      if (tid == 0) {
        int spins = 0;
        while (t_ld_acquire(fl + b) != epoch) {
          if (++spins > (1 << 22)) { timed_out = 1; break; }
        }
      }
      __syncthreads();
      if (timed_out) {
        if (tid == 0) {
          atomicExch(err, epoch);
          z[0] = __longlong_as_double(0x7ff8000000000000LL);
        }
        return;
      }
      if (tid < TT) xk[tid] = (b0 + tid < N) ? __ldcg(z + b0 + tid) : 0.0;
      __syncthreads();
    }
    // ---- update of the owned, not yet solved blocks, four per pass ----
    for (int J0 = Jfirst; J0 >= 0 && J0 < nbN; J0 += BACKWARD ? -4 * P : 4 * P) {
      const int J = block_of(J0, blk);
      if (J0 != Jfirst && J >= 0) load16(J, rv);
      double a0 = 0.0, a1 = 0.0;
      if (J >= 0) {
#pragma unroll
        for (int q = 0; q < 16; q += 2) {
          a0 = fma(rv[q], xk[16 * sub + q], a0);
          a1 = fma(rv[q + 1], xk[16 * sub + q + 1], a1);
        }
      }
      part[4 * blk + sub][r] = a0 + a1;
      __syncthreads();
      if (tid < 256) {
        const int J2 = block_of(J0, tid >> 6);
        if (J2 >= 0 && J2 * TT + r < N) {
          const int q4 = 4 * (tid >> 6);
          zown[(J2 / P) * TT + r] -= (part[q4][r] + part[q4 + 1][r]) + (part[q4 + 2][r] + part[q4 + 3][r]);
        }
      }
      __syncthreads();
    }
  }
}

// Single right-hand side per sample, Z[S][N] in place.  Returns false (nothing done) when this path does not apply.
bool run_trsv1(vbmc_b200_ctx* c, double* Z, const int* isfac_dev, cudaStream_t st, int* rc, bool backward) {
  *rc = VBMC_B200_OK;
  static const bool off = getenv("VBMC_B200_TRSV1") && atoi(getenv("VBMC_B200_TRSV1")) == 0;
  const int N = c->gp.N, S = c->gp.S, ld = c->gpLd;
  const int nbN = (N + TT - 1) / TT;
  if (off || ld < nbN * TT || (ld & 1) || S > c->num_sms) return false;
  const int nbflag = ld / TT + 1;
  const size_t need = sizeof(int) * (static_cast<size_t>(S) * nbflag + 1);
  if (c->trsvFlags.cap < need) {   // zeroed when (re)allocated; launches are told apart by an epoch; last word: error flag
    if (c->trsvFlags.reserve(need + sizeof(int) * 8 * S) != VBMC_B200_OK || cudaMemsetAsync(c->trsvFlags.p, 0, c->trsvFlags.cap, st) != cudaSuccess) {
      *rc = VBMC_B200_ECUDA;
      return true;
    }
  }
  int* flags = reinterpret_cast<int*>(c->trsvFlags.p);
  int* err = flags + c->trsvFlags.cap / sizeof(int) - 1;
  int P = c->num_sms / S;
  P = P < 1 ? 1 : (P > nbN ? nbN : P);
  if (P > 8) P = 8;
  const int smem = static_cast<int>(sizeof(double)) * (TT * (TT + 1) + 16 * TT + TT + (nbN + P - 1) / P * TT);
  if (smem > static_cast<int>(c->smem_optin) - 1024) return false;   // N beyond ~400 000 per CTA: the blocked sweep takes it
  if (cudaFuncSetAttribute(trsv1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
      cudaFuncSetAttribute(trsv1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  KernelScope ks(c, "pred_trsm", st);
  const int epoch = ++c->bsolve_epoch;
  if (backward)
    trsv1_kernel<true><<<dim3(P, S), 1024, smem, st>>>(N, ld, c->gpL.d(), static_cast<size_t>(ld) * ld, Z, isfac_dev, flags, nbflag, epoch, err);
  else
    trsv1_kernel<false><<<dim3(P, S), 1024, smem, st>>>(N, ld, c->gpL.d(), static_cast<size_t>(ld) * ld, Z, isfac_dev, flags, nbflag, epoch, err);
  if (cudaGetLastError() != cudaSuccess) {
    set_error("trsv1: kernel launch failed");
    *rc = VBMC_B200_ECUDA;
  }
  c->trsv1_pending = true;
  return true;
}

// After the caller's stream synchronisation: did a CTA of a single-column sweep give up waiting for a peer?
int trsv1_check(vbmc_b200_ctx* c) {
  if (!c->trsv1_pending) return VBMC_B200_OK;
  c->trsv1_pending = false;
  int h = 0;
  const int* err = reinterpret_cast<const int*>(c->trsvFlags.p) + c->trsvFlags.cap / sizeof(int) - 1;
  VB_CUDA(cudaMemcpy(&h, err, sizeof(int), cudaMemcpyDeviceToHost));
  if (h != 0) {
    cudaMemset(const_cast<int*>(err), 0, sizeof(int));
    VB_FAIL(VBMC_B200_ECUDA, "vbmc_b200:backsolve: a CTA of the single-column triangular sweep did not arrive (epoch %d)", h);
  }
  return VBMC_B200_OK;
}

// identity on the padding diagonal of factors attached from the host (rows/columns N .. Np-1)
__global__ void pad_identity_kernel(double* L, int N, int Np, size_t stride) {
  double* Ls = L + blockIdx.x * stride;
  for (int i = N + threadIdx.x; i < Np; i += blockDim.x) Ls[static_cast<size_t>(i) * Np + i] = 1.0;
}
int pad_identity(double* L, int N, int Np, int S, cudaStream_t st) {
  pad_identity_kernel<<<S, 64, 0, st>>>(L, N, Np, static_cast<size_t>(Np) * Np);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

}  // namespace vb
