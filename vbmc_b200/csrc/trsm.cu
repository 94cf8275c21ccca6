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
