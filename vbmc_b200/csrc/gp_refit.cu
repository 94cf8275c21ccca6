// GP-surrogate refit on sm_100a: gplite_post / gplite_nlZ -> gplite_core
// (reference: gplite/private/gplite_core.m:33-102,193,226-261,278-285; gplite/gplite_post.m:94-172;
//  gplite/gplite_nlZ.m:27-66; gplite/private/sq_dist.m; gplite_meanfun.m cases 0,1,4; gplite_noisefun.m:176-210).
//
// All S hyper-parameter samples are processed as ONE batch (grid.z / grid.y = sample): the S Cholesky
// factorisations are independent (replicas), batching them is what fills the 148 SMs.
//
// Per sample the augmented matrix  M = [ A  b ],  A = K/(sl) + diag(sn2/sn2div)  (Lchol branch, :69-82)
//                                                  or K + mult*diag(sn2)          (low-noise branch, :86-95)
// b = y - m, lives in an Np x Np column-major buffer (Np = multiple of 64 >= N+1).  A right-looking
// blocked upper Cholesky (R'R = A, block 64) runs over it; because b is carried as column N the
// forward solve z = R'\b comes for free.  The trailing update C_IJ -= P_I' P_J is the N^3/3 dense
// contraction and runs on the FP64 tensor path (mma.sync.m8n8k4.f64 = DMMA; tcgen05 has no f64 kind).
// A failed factorisation (non-positive pivot) is reported per sample so that the host can apply the
// reference's "sn2_mult *= 10, retry (<= 10x)" rule (:78-81,92-95) to exactly the failing samples.
#include <math.h>

#include <algorithm>

#include <stdlib.h>

#include "common.cuh"

namespace vb {

constexpr int TB = 64;        // block size of the factorisation
constexpr int TLD = TB + 4;   // shared-memory leading dimension (bank-conflict-free DMMA fragment loads)

struct GpBatch {
  int N, D, Np, S, Nhyp, Ncov, Nnoise, Nmean, meanfun;
  int nf0, nf1, nf2;      // noisefun
  const double* X;        // [D][N]
  const double* y;        // [N]
  const double* s2;       // [N] or null
  const double* hyp;      // [S][Nhyp]
  double* sn2;            // [S][N]
  double* mvec;           // [S][N]
  double* M;              // [S][Np*Np]
  const int* active;      // [nact] sample indices handled by this launch
  const double* scale;    // [S] 1/(sn2div*mult)   (Lchol) or 1 (low noise)
  const double* dscale;   // [S] 1/sn2div          (Lchol) or mult (low noise)
  int* info;              // [S] 0 ok, >0 first failing pivot (1-based)
  double* dscratch;       // [S][Np/64][64*64] Z_kk = R_kk^-T of every diagonal block (gp_potf2i_kernel; read by the row-panel and back-substitution kernels)
};

// ---- per-point noise variance and mean (gplite_noisefun.m:176-210, gplite_meanfun.m cases 0,1,4) ----
__global__ void gp_prep_kernel(const GpBatch g) {
  const int s = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.N) return;
  const double* h = g.hyp + static_cast<size_t>(s) * g.Nhyp;
  const double* hn = h + g.Ncov;
  int idx = 0;
  double sn2 = 2.220446049250313e-16;  // eps
  if (g.nf0 == 1) sn2 = exp(2.0 * hn[idx++]);
  if (g.nf1 == 1) sn2 += g.s2[n];
  else if (g.nf1 == 2) sn2 += exp(hn[idx++]) * g.s2[n];
  if (g.nf2 == 1) {
    const double zz = fmax(0.0, hn[idx] - g.y[n]);
    sn2 += exp(2.0 * hn[idx + 1]) * zz * zz;
  }
  g.sn2[static_cast<size_t>(s) * g.N + n] = sn2;
  const double* hm = h + g.Ncov + g.Nnoise;
  double m = 0.0;
  if (g.meanfun == 1) m = hm[0];
  if (g.meanfun == 4) {
    double z2 = 0.0;
    for (int d = 0; d < g.D; ++d) {
      const double z = (g.X[static_cast<size_t>(d) * g.N + n] - hm[1 + d]) / exp(hm[1 + g.D + d]);
      z2 = fma(z, z, z2);
    }
    m = hm[0] - 0.5 * z2;
  }
  g.mvec[static_cast<size_t>(s) * g.N + n] = m;
}

// min over n of sn2[s][n]  (Lchol = min(sn2) >= 1e-6, gplite_core.m:67)
__global__ void gp_minsn2_kernel(const double* sn2, int N, double* out) {
  __shared__ double part[256];
  const int s = blockIdx.x;
  double m = INFINITY;
  for (int n = threadIdx.x; n < N; n += blockDim.x) m = fmin(m, sn2[static_cast<size_t>(s) * N + n]);
  part[threadIdx.x] = m;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) part[threadIdx.x] = fmin(part[threadIdx.x], part[threadIdx.x + off]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[s] = part[0];
}

// ---- SE-ARD Gram, upper 64x64 tiles, + rhs column N, + identity padding -------------------------
// grid (ntiles_upper, nact), 256 threads, each thread 4x4 entries (rows ty + 16 b, columns tx + 16 a).
__global__ void __launch_bounds__(256) gp_gram_kernel(const GpBatch g) {
  extern __shared__ double sm[];
  const int s = g.active[blockIdx.y];
  const int nb = g.Np / TB;
  // decode upper tile index -> (bi, bj), bi <= bj
  int t = blockIdx.x, bi = 0;
  while (t >= nb - bi) { t -= nb - bi; ++bi; }
  const int bj = bi + t;
  const int D = g.D, N = g.N, Np = g.Np;
  double* xi = sm;             // [D][64] scaled coordinates of the row block
  double* xj = sm + D * TB;    // [D][64]
  const double* h = g.hyp + static_cast<size_t>(s) * g.Nhyp;
  for (int i = threadIdx.x; i < D * TB; i += 256) {
    const int d = i / TB, r = i - d * TB;
    const double il = exp(-h[d]);
    const int ri = bi * TB + r, rj = bj * TB + r;
    xi[i] = ri < N ? g.X[static_cast<size_t>(d) * N + ri] * il : 0.0;
    xj[i] = rj < N ? g.X[static_cast<size_t>(d) * N + rj] * il : 0.0;
  }
  __syncthreads();
  const double sf2 = exp(2.0 * h[D]) * g.scale[s];
  const double dsc = g.dscale[s];
  // rows vary fastest across the threads of a warp: a store instruction covers 16 consecutive rows (128 contiguous bytes) of two
  // columns (with columns fastest it touched 16 columns x 2 rows: 16 half-used sectors per instruction; the kernel ran at 1 TB/s)
  const int tx = threadIdx.x >> 4, ty = threadIdx.x & 15;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  // squared distances of the thread's 4x4 patch, dimension by dimension: 8 shared-memory reads per 16 entries (the entry-by-entry
  // loop read 2 per FMA and was bound by them); per entry the same operations in the same order
  double sq[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) sq[a][b] = 0.0;
  for (int d = 0; d < D; ++d) {
    double xr[4], xc[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) xr[b] = xi[d * TB + ty + 16 * b];
#pragma unroll
    for (int a = 0; a < 4; ++a) xc[a] = xj[d * TB + tx + 16 * a];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double df = xr[b] - xc[a];
        sq[a][b] = fma(df, df, sq[a][b]);
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int c = tx + 16 * a;  // column inside tile
    const int gj = bj * TB + c;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = ty + 16 * b;
      const int gi = bi * TB + r;
      double v;
      if (gi < N && gj < N) {
        v = sf2 * exp(-0.5 * sq[a][b]);                              // K_mat (:55-56), scaled
        if (gi == gj) v += dsc * g.sn2[static_cast<size_t>(s) * N + gi];
      } else if (gj == N && gi < N) {
        v = g.y[gi] - g.mvec[static_cast<size_t>(s) * N + gi];      // rhs column b = y - m
      } else {
        v = (gi == gj) ? 1.0 : 0.0;                                   // padding
      }
      Ms[static_cast<size_t>(gj) * Np + gi] = v;
    }
  }
}

// ---- panel step kb, part 1: factor the 64x64 diagonal block (one CTA per sample) -----------------------
// Right-looking elimination in shared memory, 16x16 threads x 4x4 entries, two barriers per pivot.
// Rows >= N (rhs / padding rows) are unit rows.  A non-positive pivot is reported in info[s].
__global__ void __launch_bounds__(256) gp_potf2_kernel(const GpBatch g, int kb) {
  __shared__ double A[TB][TB + 1];  // A[c][r]: column c, row r
  __shared__ double lrow[TB];       // scaled pivot row
  __shared__ int bad;
  const int s = g.active[blockIdx.x];
  const int Np = g.Np, N = g.N, tid = threadIdx.x;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  const int k0 = kb * TB;
  if (tid == 0) bad = 0;
  for (int i = tid; i < TB * TB; i += 256) {
    const int c = i >> 6, r = i & 63;
    A[c][r] = Ms[static_cast<size_t>(k0 + c) * Np + k0 + r];
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;  // thread owns columns tx+16a, rows ty+16b
  for (int p = 0; p < TB; ++p) {
    const bool unit = (k0 + p) >= N;
    const double d = A[p][p];
    if (tid == 0 && !unit && !(d > 0.0) && bad == 0) bad = k0 + p + 1;
    const double isq = unit ? 0.0 : rsqrt(d);
    if (tid < TB) {
      const int j = tid;
      double v = 0.0;
      if (j >= p) v = unit ? (j == p ? 1.0 : 0.0) : A[j][p] * isq;   // R(p, j)
      lrow[j] = v;
    }
    __syncthreads();
    if (tid < TB) A[tid][p] = lrow[tid];                              // store row p of R in place
    if (!unit) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int j = tx + 16 * a;
        const double lj = lrow[j];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = ty + 16 * b;
          if (i > p && i <= j) A[j][i] = fma(-lrow[i], lj, A[j][i]);
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0 && bad != 0) atomicCAS(&g.info[s], 0, bad);
  for (int i = tid; i < TB * TB; i += 256) {
    const int c = i >> 6, r = i & 63;
    if (r <= c) Ms[static_cast<size_t>(k0 + c) * Np + k0 + r] = A[c][r];
  }
}

// ---- panel step kb, part 2: R_kJ = R_kk^-T A_kJ for every block J to the right ---------------------------
// grid (nb-kb-1, nact).  Thread (c = tid/4, q = tid%4) owns rows q, q+4, ... of column c in registers; the
// four owners of a column sit in one warp, so the forward substitution needs no block barrier.
__global__ void __launch_bounds__(256) gp_trsm_kernel(const GpBatch g, int kb) {
  __shared__ double R[TB][TB + 1];  // R[c][r] = R_kk(r, c)
  __shared__ double ird[TB];        // 1 / R(p, p)
  const int s = g.active[blockIdx.y];
  const int Np = g.Np, tid = threadIdx.x;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  const int k0 = kb * TB;
  const int jb = kb + 1 + blockIdx.x;
  for (int i = tid; i < TB * TB; i += 256) {
    const int c = i >> 6, r = i & 63;
    R[c][r] = (r <= c) ? Ms[static_cast<size_t>(k0 + c) * Np + k0 + r] : 0.0;
  }
  __syncthreads();
  if (tid < TB) ird[tid] = 1.0 / R[tid][tid];
  __syncthreads();
  const int c = tid >> 2, q = tid & 3;
  double* col = Ms + static_cast<size_t>(jb * TB + c) * Np + k0;
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = col[q + 4 * i];
  const unsigned lane = tid & 31;
#pragma unroll
  for (int p = 0; p < TB; ++p) {
    // owner of row p: q == p % 4, register index p / 4
    double xp = x[p >> 2] * ird[p];
    xp = __shfl_sync(0xffffffffu, xp, (lane & ~3u) | (p & 3));
    if (q == (p & 3)) x[p >> 2] = xp;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = q + 4 * i;
      if (r > p) x[i] = fma(-R[r][p], xp, x[i]);   // b_r -= R(p, r) x_p
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) col[q + 4 * i] = x[i];
}

// ---- trailing update on the FP64 tensor path:  C_IJ -= P_I' P_J  (I <= J, blocks right of kb) ----
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// One CTA owns a strip of up to UPD_CHUNK tiles (I fixed, J = J0..J0+len-1): P_I stays in shared memory,
// the P_J blocks stream through a 2-stage cp.async ring, and the C tile is prefetched into registers before
// the DMMA loop, so that global-memory latency overlaps the tensor work.
// grid (nwork, nact), 256 threads = 8 warps; warp (wm, wn) owns a 16 x 32 patch of the 64x64 tile.
constexpr int UPD_CHUNK = 8;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NKEEP>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NKEEP) : "memory"); }

// copy a KD(k) x 64(cols) block (column-major, leading dim Np) into dst[c*(KD+4) + k]
template <int KD>
__device__ __forceinline__ void load_pblock_async(double* dst, const double* src, int Np, int tid) {
  constexpr int CH = KD / 2;  // 16-byte chunks per column
#pragma unroll
  for (int it = 0; it < 64 * CH / 256; ++it) {
    const int idx = tid + it * 256;
    const int c = idx / CH, ch = idx - c * CH;
    cp_async16(dst + c * (KD + 4) + ch * 2, src + static_cast<size_t>(c) * Np + ch * 2);
  }
}

// Trailing update with a panel of depth KD rows starting at row krow0:  C_IJ -= P_I' P_J for the block rows
// I = Ifirst .. Ifirst+Icount-1 and J >= I.  KD = 64: strip update inside a 128-wide panel; KD = 128: the big
// trailing update — one read-modify-write of C per 128 panel rows doubles the flop/byte (8 instead of 4), which
// is what lifts the kernel off the HBM roof (B200: DMMA peak 37 TFLOP/s needs > 5.7 flop/B).
template <int KD>
__global__ void __launch_bounds__(256) gp_update_kernel(const GpBatch g, int krow0, int Ifirst, int Icount, int chunk, int nitems) {
  extern __shared__ __align__(16) double usm[];
  constexpr int LD = KD + 4;
  double* PI = usm;              // PI[m*LD + k] = P_I(k, m)
  double* PJ0 = usm + TB * LD;   // two stages
  double* PJ1 = usm + 2 * TB * LD;
  const int s = g.active[blockIdx.y];
  const int Np = g.Np, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = Np / TB;
  // a CTA takes the work items blockIdx.x, blockIdx.x + gridDim.x, ...: with gridDim.x * gridDim.y below the SM count the
  // kernel leaves SMs free for the panel kernels that run beside it (look-ahead, refit_core)
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
  // decode work item -> (ii, chunk): row I = Ifirst+ii has ceil((nb-I)/UPD_CHUNK) chunks
  int t = item, ii = 0;
  for (;;) {
    const int nch = (nb - (Ifirst + ii) + chunk - 1) / chunk;
    if (t < nch) break;
    t -= nch;
    ++ii;
  }
  (void)Icount;
  const int I = Ifirst + ii;
  const int J0 = I + t * chunk;
  int len = nb - J0;
  len = len > chunk ? chunk : len;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  const double* prow = Ms + krow0;  // row offset of the panel
  load_pblock_async<KD>(PI, prow + static_cast<size_t>(I * TB) * Np, Np, tid);
  load_pblock_async<KD>(PJ0, prow + static_cast<size_t>(J0 * TB) * Np, Np, tid);
  cp_async_commit();
  const int wm = warp >> 1, wn = warp & 1;
  const int g4 = lane >> 2, t4 = lane & 3;
  for (int jj = 0; jj < len; ++jj) {
    const int J = J0 + jj;
    double* PJ = (jj & 1) ? PJ1 : PJ0;
    if (jj + 1 < len) load_pblock_async<KD>((jj & 1) ? PJ0 : PJ1, prow + static_cast<size_t>((J + 1) * TB) * Np, Np, tid);
    cp_async_commit();
    double cold[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int m = I * TB + wm * 16 + a * 8 + g4;
        const int n = J * TB + wn * 32 + b * 8 + 2 * t4;
        const double* c0 = Ms + static_cast<size_t>(n) * Np + m;
        cold[a][b][0] = c0[0];
        cold[a][b][1] = c0[Np];
      }
    cp_async_wait<1>();
    __syncthreads();
    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
#pragma unroll 4
    for (int k = 0; k < KD; k += 4) {
      double af[2], bf[4];
#pragma unroll
      for (int a = 0; a < 2; ++a) af[a] = PI[(wm * 16 + a * 8 + g4) * LD + k + t4];
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = PJ[(wn * 32 + b * 8 + g4) * LD + k + t4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int m = I * TB + wm * 16 + a * 8 + g4;
        const int n = J * TB + wn * 32 + b * 8 + 2 * t4;
        double* c0 = Ms + static_cast<size_t>(n) * Np + m;
        c0[0] = cold[a][b][0] - acc[a][b][0];
        c0[Np] = cold[a][b][1] - acc[a][b][1];
      }
    __syncthreads();  // all warps done with PJ before the ring slot is refilled
  }
  }  // work items
}

// ---- panel step kb, second generation (round 2) -------------------------------------------------------------------------------
// The panel is the latency-bound part of the factorisation (64 dependent pivots per step, 32 steps at N = 2000).
// (1) gp_potf2i_kernel: the 64x64 diagonal block lives in REGISTERS (thread (tx, ty) owns rows ty+16b, columns tx+16a); per pivot
//     the 16 owners of row p publish it raw, ONE barrier, every thread scales what it needs by rsqrt(d) itself and updates its own
//     patch -- no shared-memory read-modify-write, no second barrier, block rows/columns that cannot be touched any more are
//     skipped at compile time.  The same elimination runs on an identity right-hand side, so Z = R_kk^-T (lower triangular) falls
//     out of the loop with it (R'Z = I) and goes to the scratch buffer.  Same operations in the same order as gp_potf2_kernel:
//     the factor is bit-identical.  (Measured slower: the owners scaling the row BEFORE the barrier, one rsqrt per pivot instead
//     of one per thread -- 1.04 vs 0.91 ms over the 32 steps at c3: the loads of the row no longer overlap the rsqrt.)
// (2) gp_trsmg_kernel (below): the row panel R_kJ = R_kk^-T A_kJ as block substitution on the FP64 tensor path instead of 64
//     dependent scalar steps.
__global__ void __launch_bounds__(256) gp_potf2i_kernel(const GpBatch g, int kb) {
  __shared__ double A[TB][TB + 1];       // A[c][r]: column c, row r; rows of R are stored in place as they are finished
  __shared__ double rowbuf[2][2 * TB];   // raw pivot row, by pivot parity: [0, 64) A(p, j), [64, 128) E(p, c)
  const int s = g.active[blockIdx.x];
  const int Np = g.Np, N = g.N, tid = threadIdx.x;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  double* Z = g.dscratch + (static_cast<size_t>(s) * (Np / TB) + kb) * TB * TB;   // Z[p*64 + c] = R_kk^-T (p, c)
  const int k0 = kb * TB;
  for (int i = tid; i < TB * TB; i += 256) {
    const int c = i >> 6, r = i & 63;
    A[c][r] = Ms[static_cast<size_t>(k0 + c) * Np + k0 + r];
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  double a[4][4], e[4][4];   // a[ja][ib] = A(ty + 16 ib, tx + 16 ja), e[ca][ib] = E(ty + 16 ib, tx + 16 ca), E = I at the start
#pragma unroll
  for (int ja = 0; ja < 4; ++ja)
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) {
      a[ja][ib] = A[tx + 16 * ja][ty + 16 * ib];
      e[ja][ib] = (tx + 16 * ja == ty + 16 * ib) ? 1.0 : 0.0;
    }
  int bad = 0;
#pragma unroll
  for (int pb = 0; pb < 4; ++pb) {
#pragma unroll 1
    for (int pt = 0; pt < 16; ++pt) {
      const int p = 16 * pb + pt, par = p & 1;
      if (ty == pt) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          rowbuf[par][tx + 16 * q] = a[q][pb];
          rowbuf[par][TB + tx + 16 * q] = e[q][pb];
        }
      }
      __syncthreads();   // (the buffer of the other parity is free again: everybody passed the previous barrier after reading it)
      const bool unit = (k0 + p) >= N;
      const double d = rowbuf[par][p];
      if (!unit && !(d > 0.0) && bad == 0) bad = k0 + p + 1;
      const double isq = unit ? 0.0 : rsqrt(d);
      double lj[4], li[4], zc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        lj[q] = rowbuf[par][tx + 16 * q] * isq;
        li[q] = rowbuf[par][ty + 16 * q] * isq;
        zc[q] = rowbuf[par][TB + tx + 16 * q] * isq;
      }
      if (!unit) {
#pragma unroll
        for (int ib = pb; ib < 4; ++ib) {
          if (ty + 16 * ib > p) {
#pragma unroll
            for (int ja = ib; ja < 4; ++ja) a[ja][ib] = fma(-li[ib], lj[ja], a[ja][ib]);
#pragma unroll
            for (int ca = 0; ca <= pb; ++ca) e[ca][ib] = fma(-li[ib], zc[ca], e[ca][ib]);
          }
        }
      }
      if (ty == pt) {   // off the critical path: row p of R and of Z = R^-T
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = tx + 16 * q;
          A[j][p] = unit ? (j == p ? 1.0 : 0.0) : (j >= p ? lj[q] : 0.0);   // R(p, j)
          Z[p * TB + j] = unit ? (j == p ? 1.0 : 0.0) : (j <= p ? zc[q] : 0.0);
        }
      }
    }
  }
  if (tid == 0 && bad != 0) atomicCAS(&g.info[s], 0, bad);   // every thread saw the same pivots: `bad` is the first failure
  __syncthreads();
  for (int i = tid; i < TB * TB; i += 256) {
    const int c = i >> 6, r = i & 63;
    if (r <= c) Ms[static_cast<size_t>(k0 + c) * Np + k0 + r] = A[c][r];
  }
}

// Row panel R_kJ = R_kk^-T A_kJ for the block J = kb + 1 + blockIdx.x: block substitution in four 16-row stages,
//   T_i = B_i - sum_{p < 16 i} R(p, .)' X(p, .)      (DMMA, depth 16 i)
//   X_i = Z_ii T_i,   Z_ii = inv(R_ii)' = the i-th 16x16 diagonal block of Z      (DMMA, depth 16)
// so that only the 16x16 diagonal blocks are applied as explicit inverses (their condition number, not the 64x64 block's, enters
// the error: measured backward error |R'R - A|/|A| of the whole factor 1e-15 like scalar substitution, against 1-2e-14 with the
// full 64x64 inverse).  A warp owns 8 columns of the block through all four stages: no block barrier after the loads.
// grid (nb-kb-1, nact), 256 threads; 71.6 KB of shared memory: 3 CTAs per SM.
constexpr int TG_LDR = 52, TG_LDX = TB + 4, TG_LDZ = 20;
constexpr int TRSMG_SMEM_BYTES = (TB * TG_LDR + TB * TG_LDX + 4 * 16 * TG_LDZ) * 8;
__global__ void __launch_bounds__(256) gp_trsmg_kernel(const GpBatch g, int kb) {
  extern __shared__ __align__(16) double tsm[];
  double* RA = tsm;                  // RA[r*52 + p] = R_kk(p, r), p < 48
  double* XS = RA + TB * TG_LDR;     // XS[c*68 + p] = B(p, c), solved in place
  double* ZD = XS + TB * TG_LDX;     // ZD[i][m*20 + k] = Z(16 i + m, 16 i + k)
  const int s = g.active[blockIdx.y];
  const int Np = g.Np, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  const int k0 = kb * TB, jb = kb + 1 + blockIdx.x;
  double* Bg = Ms + static_cast<size_t>(jb * TB) * Np + k0;
  load_pblock_async<48>(RA, Ms + static_cast<size_t>(k0) * Np + k0, Np, tid);
  load_pblock_async<TB>(XS, Bg, Np, tid);
  {
    const double* Zg = g.dscratch + (static_cast<size_t>(s) * (Np / TB) + kb) * TB * TB;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + 256 * it, i = idx >> 7, m = (idx >> 3) & 15, ch = idx & 7;
      cp_async16(ZD + i * 16 * TG_LDZ + m * TG_LDZ + 2 * ch, Zg + (16 * i + m) * TB + 16 * i + 2 * ch);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int g4 = lane >> 2, t4 = lane & 3, c0 = warp * 8;
  const double* xb = XS + (c0 + g4) * TG_LDX;   // B-fragment column of this lane
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double acc[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a) acc[a][0] = acc[a][1] = 0.0;
#pragma unroll
    for (int k = 0; k < 16 * i; k += 4) {
      const double bf = xb[k + t4];
#pragma unroll
      for (int a = 0; a < 2; ++a) dmma_m8n8k4(acc[a][0], acc[a][1], RA[(16 * i + 8 * a + g4) * TG_LDR + k + t4], bf);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double* t = XS + (c0 + 2 * t4 + j) * TG_LDX + 16 * i + 8 * a + g4;
        *t -= acc[a][j];
      }
    __syncwarp();
    double x[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a) x[a][0] = x[a][1] = 0.0;
#pragma unroll
    for (int k = 0; k < 16; k += 4) {
      const double bf = xb[16 * i + k + t4];
#pragma unroll
      for (int a = 0; a < 2; ++a) dmma_m8n8k4(x[a][0], x[a][1], ZD[i * 16 * TG_LDZ + (8 * a + g4) * TG_LDZ + k + t4], bf);
    }
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int j = 0; j < 2; ++j) XS[(c0 + 2 * t4 + j) * TG_LDX + 16 * i + 8 * a + g4] = x[a][j];
    __syncwarp();
  }
#pragma unroll
  for (int cc = 0; cc < 8; ++cc)
#pragma unroll
    for (int h = 0; h < 2; ++h) Bg[static_cast<size_t>(c0 + cc) * Np + lane + 32 * h] = XS[(c0 + cc) * TG_LDX + lane + 32 * h];
}

// number of CTAs (work items) of gp_update_kernel for block rows Ifirst .. Ifirst+Icount-1
static int update_work_items(int nb, int Ifirst, int Icount, int chunk) {
  int n = 0;
  for (int ii = 0; ii < Icount; ++ii) n += (nb - (Ifirst + ii) + chunk - 1) / chunk;
  return n;
}
// tiles per CTA: as long as possible (P_I reuse, pipelined P_J) while keeping >= 2 CTAs per SM in flight
static int update_chunk(int nb, int Ifirst, int Icount, int nact, int num_sms) {
  long long tiles = 0;
  for (int ii = 0; ii < Icount; ++ii) tiles += nb - (Ifirst + ii);
  tiles *= nact;
  long long ch = tiles / (2LL * num_sms);
  return ch < 1 ? 1 : (ch > UPD_CHUNK ? UPD_CHUNK : static_cast<int>(ch));
}

// ---- back substitution R x = z (z = column N of the factored buffer), blocked like the factorisation ----
// nlZ ingredients: sum(log(diag R)) and z'z  (gplite_core.m:193); one CTA per sample.
__global__ void __launch_bounds__(256) gp_nlzparts_kernel(const GpBatch g, double* logdet, double* zz) {
  __shared__ double part[256];
  const int s = blockIdx.x;
  const int Np = g.Np, N = g.N, tid = threadIdx.x;
  const double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  const double* z = Ms + static_cast<size_t>(N) * Np;
  double ld = 0.0, q = 0.0;
  for (int i = tid; i < N; i += 256) {
    ld += log(Ms[static_cast<size_t>(i) * Np + i]);
    q = fma(z[i], z[i], q);
  }
  part[tid] = ld;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) part[tid] += part[tid + off];
    __syncthreads();
  }
  if (tid == 0) logdet[s] = part[0];
  __syncthreads();
  part[tid] = q;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) part[tid] += part[tid + off];
    __syncthreads();
  }
  if (tid == 0) zz[s] = part[0];
}

// solve the 64x64 diagonal block kb for x_k (in place in column N); one warp per sample, lane owns rows
// lane and lane+32; the pivot value is broadcast by shuffle.
__global__ void __launch_bounds__(32) gp_bsolve_diag_kernel(const GpBatch g, int kb) {
  __shared__ double Rk[TB][TB + 1];  // Rk[c][r]
  const int s = blockIdx.x;
  const int Np = g.Np, N = g.N, lane = threadIdx.x;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  double* z = Ms + static_cast<size_t>(N) * Np;
  const int k0 = kb * TB;
  for (int i = lane; i < TB * TB; i += 32) {
    const int c = i >> 6, r = i & 63;
    Rk[c][r] = (k0 + c < N && k0 + r < N && r <= c) ? Ms[static_cast<size_t>(k0 + c) * Np + k0 + r] : (c == r ? 1.0 : 0.0);
  }
  double x0 = (k0 + lane < N) ? z[k0 + lane] : 0.0;
  double x1 = (k0 + lane + 32 < N) ? z[k0 + lane + 32] : 0.0;
  __syncwarp();
  for (int p = TB - 1; p >= 0; --p) {
    double xp = (p >= 32 ? x1 : x0) / Rk[p][p];
    xp = __shfl_sync(0xffffffffu, xp, p & 31);
    if (p >= 32) {
      if (lane == (p & 31)) x1 = xp;
      if (lane + 32 < p) x1 = fma(-Rk[p][lane + 32], xp, x1);
      x0 = fma(-Rk[p][lane], xp, x0);
    } else {
      if (lane == p) x0 = xp;
      if (lane < p) x0 = fma(-Rk[p][lane], xp, x0);
    }
  }
  if (k0 + lane < N) z[k0 + lane] = x0;
  if (k0 + lane + 32 < N) z[k0 + lane + 32] = x1;
}

// z_i -= sum_{j<64} R(i, k0+j) x_{k0+j}   for i < k0;  grid (ceil(k0/256), S), coalesced over i
__global__ void __launch_bounds__(256) gp_bsolve_update_kernel(const GpBatch g, int kb) {
  __shared__ double xk[TB];
  const int s = blockIdx.y;
  const int Np = g.Np, N = g.N, tid = threadIdx.x;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  double* z = Ms + static_cast<size_t>(N) * Np;
  const int k0 = kb * TB;
  if (tid < TB) xk[tid] = (k0 + tid < N) ? z[k0 + tid] : 0.0;
  __syncthreads();
  const int i = blockIdx.x * 256 + tid;
  if (i >= k0) return;
  double acc0 = 0.0, acc1 = 0.0;
  const double* col = Ms + static_cast<size_t>(k0) * Np + i;
#pragma unroll 8
  for (int j = 0; j < TB; j += 2) {
    acc0 = fma(col[static_cast<size_t>(j) * Np], xk[j], acc0);
    acc1 = fma(col[static_cast<size_t>(j + 1) * Np], xk[j + 1], acc1);
  }
  z[i] -= acc0 + acc1;
}

// Whole back substitution of one sample in ONE CTA of 1024 threads (replaces the 2-kernels-per-block chain):
// per 64-block: warp 0 solves the diagonal block (lane owns rows lane, lane+32; pivot broadcast by shuffle) while
// the block was staged by all threads; then every thread updates rows i < k0 (coalesced over i).
__global__ void __launch_bounds__(1024) gp_bsolve_kernel(const GpBatch g, const double* ascale, double* alpha) {
  __shared__ double Rk[TB][TB + 1];
  __shared__ double xk[TB];
  const int s = blockIdx.x;
  const int Np = g.Np, N = g.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  double* z = Ms + static_cast<size_t>(N) * Np;
  for (int kb = (N + TB - 1) / TB - 1; kb >= 0; --kb) {
    const int k0 = kb * TB;
    for (int i = tid; i < TB * TB; i += 1024) {
      const int c = i >> 6, r = i & 63;
      Rk[c][r] = (k0 + c < N && k0 + r < N && r <= c) ? Ms[static_cast<size_t>(k0 + c) * Np + k0 + r] : (c == r ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp == 0) {
      double x0 = (k0 + lane < N) ? z[k0 + lane] : 0.0;
      double x1 = (k0 + lane + 32 < N) ? z[k0 + lane + 32] : 0.0;
      for (int p = TB - 1; p >= 0; --p) {
        double xp = (p >= 32 ? x1 : x0) / Rk[p][p];
        xp = __shfl_sync(0xffffffffu, xp, p & 31);
        if (p >= 32) {
          if (lane == (p & 31)) x1 = xp;
          if (lane + 32 < p) x1 = fma(-Rk[p][lane + 32], xp, x1);
          x0 = fma(-Rk[p][lane], xp, x0);
        } else {
          if (lane == p) x0 = xp;
          if (lane < p) x0 = fma(-Rk[p][lane], xp, x0);
        }
      }
      xk[lane] = x0;
      xk[lane + 32] = x1;
      if (k0 + lane < N) z[k0 + lane] = x0;
      if (k0 + lane + 32 < N) z[k0 + lane + 32] = x1;
    }
    __syncthreads();
    for (int i = tid; i < k0; i += 1024) {
      double a0 = 0.0, a1 = 0.0;
      const double* col = Ms + static_cast<size_t>(k0) * Np + i;
#pragma unroll 8
      for (int j = 0; j < TB; j += 2) {
        a0 = fma(col[static_cast<size_t>(j) * Np], xk[j], a0);
        a1 = fma(col[static_cast<size_t>(j + 1) * Np], xk[j + 1], a1);
      }
      z[i] -= a0 + a1;
    }
    __syncthreads();
  }
  const double sc = ascale[s];
  for (int i = tid; i < N; i += 1024) alpha[static_cast<size_t>(s) * N + i] = z[i] * sc;
}

// Second generation: the diagonal blocks are applied through Z_kk = R_kk^-T kept by gp_potf2i_kernel (x_k = Z_kk' y_k: a 64x64
// matrix-vector product over all 1024 threads instead of 64 dependent divide-and-update steps in one warp), and the update of
// the rows above keeps four independent sums and 16 loads in flight per thread (one SM streams the 16 MB of a sample's factor).
__global__ void __launch_bounds__(1024) gp_bsolve2_kernel(const GpBatch g, const double* ascale, double* alpha) {
  __shared__ double part[16][TB];
  __shared__ double xk[TB], yk[TB];
  const int s = blockIdx.x;
  const int Np = g.Np, N = g.N, tid = threadIdx.x;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  double* z = Ms + static_cast<size_t>(N) * Np;
  const int r = tid & 63, pg = tid >> 6;
  for (int kb = (N + TB - 1) / TB - 1; kb >= 0; --kb) {
    const int k0 = kb * TB;
    const double* Zg = g.dscratch + (static_cast<size_t>(s) * (Np / TB) + kb) * TB * TB;   // Z(p, r) at [p*64 + r], zero for r > p
    if (tid < TB) yk[tid] = (k0 + tid < N) ? z[k0 + tid] : 0.0;
    double zr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) zr[q] = Zg[(4 * pg + q) * TB + r];
    __syncthreads();
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc = fma(zr[q], yk[4 * pg + q], acc);
    part[pg][r] = acc;
    __syncthreads();
    if (tid < TB) {
      double x = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) x += part[q][tid];
      xk[tid] = x;
      if (k0 + tid < N) z[k0 + tid] = x;
    }
    __syncthreads();
    for (int i = tid; i < k0; i += 1024) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      const double* col = Ms + static_cast<size_t>(k0) * Np + i;
#pragma unroll 4
      for (int j = 0; j < TB; j += 4) {
        a0 = fma(col[static_cast<size_t>(j) * Np], xk[j], a0);
        a1 = fma(col[static_cast<size_t>(j + 1) * Np], xk[j + 1], a1);
        a2 = fma(col[static_cast<size_t>(j + 2) * Np], xk[j + 2], a2);
        a3 = fma(col[static_cast<size_t>(j + 3) * Np], xk[j + 3], a3);
      }
      z[i] -= (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
  }
  const double sc = ascale[s];
  for (int i = tid; i < N; i += 1024) alpha[static_cast<size_t>(s) * N + i] = z[i] * sc;
}

// Third generation: P CTAs per sample (P * S <= number of SMs: all of them resident at once).  Block row J of the solution belongs
// to CTA J mod P, which is the only one that ever writes z_J; the owner of block kb turns it into x_kb = Z_kb' z_kb, stores it in
// place and raises flag[s][kb] (release); every CTA then subtracts R(J, kb) x_kb from the blocks J < kb it owns.  The factor's
// entries of a step are loaded BEFORE the flag is awaited (their addresses do not depend on x), so the chain per step is
// flag -> x -> FMAs -> the next owner's 64x64 product.  `epoch` distinguishes launches (the flags are never reset).
// A peer that does not arrive within ~2^22 polls (seconds) is reported through info[s] = -7 instead of hanging the GPU.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void __launch_bounds__(1024) gp_bsolve3_kernel(const GpBatch g, const double* ascale, double* alpha, int* flags, int epoch) {
  __shared__ double part[16][TB];   // partial sums: [16 row groups][64] (solve) or [4 blocks][4 column groups][64] (update)
  __shared__ double xk[TB], yk[TB];
  __shared__ int timed_out;
  const int s = blockIdx.y, me = blockIdx.x, P = gridDim.x;
  const int Np = g.Np, N = g.N, tid = threadIdx.x;
  const int nbp = Np / TB, nbN = (N + TB - 1) / TB;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  double* z = Ms + static_cast<size_t>(N) * Np;
  int* fl = flags + static_cast<size_t>(s) * nbp;
  const int r = tid & 63;
  if (tid == 0) timed_out = 0;
  __syncthreads();
  for (int kb = nbN - 1; kb >= 0; --kb) {
    const int k0 = kb * TB;
    // prefetch this step's factor entries for the first (up to) four own blocks: thread (blk, cg, r) takes 16 columns of one row
    const int blk = tid >> 8, cg = (tid >> 6) & 3;
    int Jfirst = kb - 1 - ((kb - 1 - me) % P + P) % P;   // largest J < kb with J == me (mod P); negative: none
    if (kb == 0) Jfirst = -1;
    const int Jmine = Jfirst - blk * P;
    double rv[16];
    if (Jmine >= 0) {
      const double* col = Ms + static_cast<size_t>(k0 + 16 * cg) * Np + Jmine * TB + r;
#pragma unroll
      for (int q = 0; q < 16; ++q) rv[q] = __ldcs(col + static_cast<size_t>(q) * Np);
    }
    if (kb % P == me) {
      // ---- my block: x = Z' y (Z(p, c) at [p*64 + c], zero for c > p), thread (pg, r) sums four rows p ----
      const int pg = tid >> 6;
      const double* Zg = g.dscratch + (static_cast<size_t>(s) * nbp + kb) * TB * TB;
      double zr[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) zr[q] = Zg[(4 * pg + q) * TB + r];
      if (tid < TB) yk[tid] = (k0 + tid < N) ? z[k0 + tid] : 0.0;
      __syncthreads();
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) acc = fma(zr[q], yk[4 * pg + q], acc);
      part[pg][r] = acc;
      __syncthreads();
      if (tid < TB) {
        double x = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) x += part[q][tid];
        xk[tid] = x;
        if (k0 + tid < N) {
          z[k0 + tid] = x;
          alpha[static_cast<size_t>(s) * N + k0 + tid] = x * ascale[s];
        }
        __threadfence();
      }
      __syncthreads();
      if (tid == 0 && P > 1) st_release_gpu(fl + kb, epoch);
    } else {
      if (tid == 0) {
        int spins = 0;
        while (ld_acquire_gpu(fl + kb) != epoch) {
          if (++spins > (1 << 22)) { timed_out = 1; break; }
        }
      }
      __syncthreads();
      if (timed_out) {
        if (tid == 0) atomicExch(&g.info[s], -7);
        return;
      }
      if (tid < TB) xk[tid] = (k0 + tid < N) ? __ldcg(z + k0 + tid) : 0.0;
      __syncthreads();
    }
    // ---- z_J -= R(J, kb) x_kb for my blocks J < kb, four at a time ----
    for (int J0 = Jfirst; J0 >= 0; J0 -= 4 * P) {
      const int J = J0 - blk * P;
      if (J0 != Jfirst && J >= 0) {
        const double* col = Ms + static_cast<size_t>(k0 + 16 * cg) * Np + J * TB + r;
#pragma unroll
        for (int q = 0; q < 16; ++q) rv[q] = __ldcs(col + static_cast<size_t>(q) * Np);
      }
      double a0 = 0.0, a1 = 0.0;
      if (J >= 0) {
#pragma unroll
        for (int q = 0; q < 16; q += 2) {
          a0 = fma(rv[q], xk[16 * cg + q], a0);
          a1 = fma(rv[q + 1], xk[16 * cg + q + 1], a1);
        }
      }
      part[4 * blk + cg][r] = a0 + a1;
      __syncthreads();
      if (tid < 256) {
        const int b2 = tid >> 6, J2 = J0 - b2 * P;
        if (J2 >= 0) z[J2 * TB + r] -= (part[4 * b2][r] + part[4 * b2 + 1][r]) + (part[4 * b2 + 2][r] + part[4 * b2 + 3][r]);
      }
      __syncthreads();
    }
  }
}

__global__ void gp_alpha_kernel(const GpBatch g, const double* ascale, double* alpha) {
  const int s = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.N) return;
  const double* z = g.M + static_cast<size_t>(s) * g.Np * g.Np + static_cast<size_t>(g.N) * g.Np;
  alpha[static_cast<size_t>(s) * g.N + i] = z[i] * ascale[s];
}

// ---- inverse tiles on the tensor path: inv(A) = X'X with X = R^-T (lower), fused epilogues --------------------
//   mode 0 (gplite_core.m:226-234): Q = inv(A)/sl - alpha*alpha';  S_i = sum sum Q.*K_mat.*sq_dist(X(:,i)'/ell_i),
//           S_D = sum sum Q.*K_mat, diag(Q)   — inv(A) is never stored;
//   mode 1 (gplite_core.m:96-99): pL = -inv(A) written out (low-noise posterior).
// grid = upper 64x64 tiles (I <= J); X(c,a) at Xinv[a*N + c].
struct InvArgs {
  int N, D, mode;
  const double* Xinv;
  const double* Xc;     // [D][N] training inputs
  const double* hyp;    // [Nhyp]
  const double* alpha;  // [N]
  double inv_sl;
  double* partial;      // [ntiles][D+1]
  double* diagQ;        // [N]
  double* outL;         // [N][N] (mode 1)
};

__global__ void __launch_bounds__(256) gp_invtile_kernel(const InvArgs g) {
  extern __shared__ __align__(16) double ism[];
  double* PI = ism;
  double* PJ = ism + TB * TLD;
  double* xi = PJ + TB * TLD;      // [D][64]
  double* xj = xi + g.D * TB;      // [D][64]
  double* red = xj + g.D * TB;     // [(D+1)][256]
  const int N = g.N, D = g.D, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = (N + TB - 1) / TB;
  int t = blockIdx.x, I = 0;
  while (t >= nb - I) { t -= nb - I; ++I; }
  const int J = I + t;
  const int wm = warp >> 1, wn = warp & 1, g4 = lane >> 2, t4 = lane & 3;
  double acc[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  for (int cb = J; cb < nb; ++cb) {
    __syncthreads();
    for (int i = tid; i < TB * TB; i += 256) {
      const int col = i >> 6, k = i & 63;
      const int c = cb * TB + k, a = I * TB + col, b = J * TB + col;
      PI[col * TLD + k] = (c < N && a < N) ? g.Xinv[static_cast<size_t>(a) * N + c] : 0.0;
      PJ[col * TLD + k] = (c < N && b < N) ? g.Xinv[static_cast<size_t>(b) * N + c] : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < TB; k += 4) {
      double af[2], bf[4];
#pragma unroll
      for (int a = 0; a < 2; ++a) af[a] = PI[(wm * 16 + a * 8 + g4) * TLD + k + t4];
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = PJ[(wn * 32 + b * 8 + g4) * TLD + k + t4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
  if (g.mode == 1) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int m = I * TB + wm * 16 + a * 8 + g4, n = J * TB + wn * 32 + b * 8 + 2 * t4 + e;
          if (m < N && n < N) {
            g.outL[static_cast<size_t>(n) * N + m] = -acc[a][b][e];
            g.outL[static_cast<size_t>(m) * N + n] = -acc[a][b][e];
          }
        }
    return;
  }
  // ---- mode 0: fused Hadamard reductions ----
  __syncthreads();
  for (int i = tid; i < D * TB; i += 256) {
    const int d = i / TB, r = i - d * TB;
    const double il = exp(-g.hyp[d]);
    const int ri = I * TB + r, rj = J * TB + r;
    xi[i] = ri < N ? g.Xc[static_cast<size_t>(d) * N + ri] * il : 0.0;
    xj[i] = rj < N ? g.Xc[static_cast<size_t>(d) * N + rj] * il : 0.0;
  }
  __syncthreads();
  const double sf2 = exp(2.0 * g.hyp[D]);
  const double wgt = (I == J) ? 1.0 : 2.0;  // the strictly upper tiles stand for their mirror images as well
  double sums[25];
#pragma unroll
  for (int d = 0; d < 25; ++d) sums[d] = 0.0;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int rm = wm * 16 + a * 8 + g4, rn = wn * 32 + b * 8 + 2 * t4 + e;
        const int m = I * TB + rm, n = J * TB + rn;
        if (m < N && n < N) {
          const double Q = acc[a][b][e] * g.inv_sl - g.alpha[m] * g.alpha[n];   // (:226)
          double sq = 0.0;
          for (int d = 0; d < D; ++d) {
            const double df = xi[d * TB + rm] - xj[d * TB + rn];
            sq = fma(df, df, sq);
          }
          const double QK = wgt * Q * sf2 * exp(-0.5 * sq);
          sums[D] += QK;                                                          // (:234)
          for (int d = 0; d < D; ++d) {
            const double df = xi[d * TB + rm] - xj[d * TB + rn];
            sums[d] = fma(QK, df * df, sums[d]);                                  // (:230-232)
          }
          if (m == n) g.diagQ[m] = Q;
        }
      }
  for (int d = 0; d <= D; ++d) red[d * 256 + tid] = sums[d];
  __syncthreads();
  for (int d = warp; d <= D; d += 8) {
    double v = 0.0;
    for (int i = lane; i < 256; i += 32) v += red[d * 256 + i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) g.partial[static_cast<size_t>(blockIdx.x) * (D + 1) + d] = v;
  }
}

// copy the N x N factor out of the padded buffer, zeroing the strictly lower part (MATLAB's chol output)
__global__ void gp_extract_kernel(const double* M, int Np, int N, double* out, int negate_full) {
  const int s = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= N) return;
  const double v = M[static_cast<size_t>(s) * Np * Np + static_cast<size_t>(j) * Np + i];
  out[static_cast<size_t>(s) * N * N + static_cast<size_t>(j) * N + i] = negate_full ? v : (i <= j ? v : 0.0);
}

}  // namespace vb

using namespace vb;

namespace {

struct RefitResult {
  std::vector<double> minsn2, mult, sl, logdet, zz;
  std::vector<int> Lchol;
};

// Runs prep + Gram + batched Cholesky (+ retries) + back substitution for all S samples of `gd`.
// Leaves alpha in c->gpAlpha ([S][N]) and the factors in c->gpL (padded [S][Np][Np]).
int refit_core(vbmc_b200_ctx* c, const vbmc_b200_gp_desc* gd, int Ncov, int Nnoise, int Nmean, RefitResult* rr) {
  const int N = gd->N, D = gd->D, S = gd->S;
  const int Np = (N + 1 + TB - 1) / TB * TB;
  cudaStream_t st = c->stream;
  VB_TRY(c->gpX.reserve(sizeof(double) * N * D));
  VB_TRY(c->gpY.reserve(sizeof(double) * N));
  VB_TRY(c->gpHyp.reserve(sizeof(double) * S * gd->Nhyp));
  VB_TRY(c->gpAlpha.reserve(sizeof(double) * static_cast<size_t>(S) * N));
  VB_TRY(c->gpL.reserve(sizeof(double) * static_cast<size_t>(S) * Np * Np));
  // work: sn2[S][N] mvec[S][N] scale[S] dscale[S] ascale[S] minsn2[S] logdet[S] zz[S] | info[S] active[S]
  const size_t nwork = 2 * static_cast<size_t>(S) * N + 6 * S + static_cast<size_t>(S) * (Np / TB) * TB * TB;
  VB_TRY(c->gpWork.reserve(sizeof(double) * nwork + sizeof(int) * 2 * S + 64));
  if (c->gpFlags.cap < sizeof(int) * static_cast<size_t>(S) * (Np / TB)) {   // zeroed when (re)allocated; launches are told apart by an epoch
    VB_TRY(c->gpFlags.reserve(sizeof(int) * static_cast<size_t>(S) * (Np / TB)));
    VB_CUDA(cudaMemsetAsync(c->gpFlags.p, 0, c->gpFlags.cap, st));
  }
  int* d_flags = c->gpFlags.i();
  VB_CUDA(cudaMemcpyAsync(c->gpX.p, gd->X, sizeof(double) * N * D, cudaMemcpyHostToDevice, st));
  VB_CUDA(cudaMemcpyAsync(c->gpY.p, gd->y, sizeof(double) * N, cudaMemcpyHostToDevice, st));
  VB_CUDA(cudaMemcpyAsync(c->gpHyp.p, gd->hyp, sizeof(double) * S * gd->Nhyp, cudaMemcpyHostToDevice, st));
  if (gd->s2) {
    VB_TRY(c->gpS2.reserve(sizeof(double) * N));
    VB_CUDA(cudaMemcpyAsync(c->gpS2.p, gd->s2, sizeof(double) * N, cudaMemcpyHostToDevice, st));
  }
  double* w = c->gpWork.d();
  GpBatch g;
  g.N = N; g.D = D; g.Np = Np; g.S = S; g.Nhyp = gd->Nhyp; g.Ncov = Ncov; g.Nnoise = Nnoise; g.Nmean = Nmean;
  g.meanfun = gd->meanfun;
  g.nf0 = gd->noisefun[0]; g.nf1 = gd->noisefun[1]; g.nf2 = gd->noisefun[2];
  g.X = c->gpX.d(); g.y = c->gpY.d(); g.s2 = gd->s2 ? c->gpS2.d() : nullptr; g.hyp = c->gpHyp.d();
  g.sn2 = w; g.mvec = w + static_cast<size_t>(S) * N;
  double* d_scale = g.mvec + static_cast<size_t>(S) * N;
  double* d_dscale = d_scale + S;
  double* d_ascale = d_dscale + S;
  double* d_minsn2 = d_ascale + S;
  double* d_logdet = d_minsn2 + S;
  double* d_zz = d_logdet + S;
  double* d_dscratch = d_zz + S;
  int* d_info = reinterpret_cast<int*>(d_dscratch + static_cast<size_t>(S) * (Np / TB) * TB * TB);
  int* d_active = d_info + S;
  g.dscratch = d_dscratch;
  g.M = c->gpL.d(); g.scale = d_scale; g.dscale = d_dscale; g.info = d_info; g.active = d_active;
  {
    dim3 grid((N + 255) / 256, S);
    KernelScope ks(c, "gp_prep", st);
    gp_prep_kernel<<<grid, 256, 0, st>>>(g);
  }
  {
    KernelScope ks(c, "gp_prep", st);
    gp_minsn2_kernel<<<S, 256, 0, st>>>(g.sn2, N, d_minsn2);
  }
  VB_CUDA(cudaGetLastError());
  rr->minsn2.assign(S, 0.0);
  VB_CUDA(cudaMemcpyAsync(rr->minsn2.data(), d_minsn2, sizeof(double) * S, cudaMemcpyDeviceToHost, st));
  VB_CUDA(cudaStreamSynchronize(st));
  const bool scalar_noise = gd->noisefun[1] == 0 && gd->noisefun[2] == 0;
  (void)scalar_noise;  // scalar and per-point noise share one code path: sn2div = min(sn2) == sn2 when scalar
  rr->mult.assign(S, 1.0);
  rr->sl.assign(S, 1.0);
  rr->Lchol.assign(S, 1);
  for (int s = 0; s < S; ++s) rr->Lchol[s] = rr->minsn2[s] >= 1e-6 ? 1 : 0;  // gplite_core.m:67
  std::vector<int> active(S);
  for (int s = 0; s < S; ++s) active[s] = s;
  std::vector<double> h_scale(S), h_dscale(S), h_ascale(S);
  std::vector<int> h_info(S);
  const int nb = Np / TB;
  const int UPDATE_SMEM64 = 3 * TB * (64 + 4) * sizeof(double), UPDATE_SMEM128 = 3 * TB * (128 + 4) * sizeof(double);
  VB_CUDA(cudaFuncSetAttribute(gp_update_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, UPDATE_SMEM64));
  VB_CUDA(cudaFuncSetAttribute(gp_update_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, UPDATE_SMEM128));
  const int TRSMG_SMEM = TRSMG_SMEM_BYTES;
  VB_CUDA(cudaFuncSetAttribute(gp_trsmg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSMG_SMEM));
  static const bool panel_v1 = getenv("VBMC_B200_REFIT_PANEL_V1") && atoi(getenv("VBMC_B200_REFIT_PANEL_V1")) != 0;
  static const bool trsm_v1 = getenv("VBMC_B200_REFIT_TRSM_V1") && atoi(getenv("VBMC_B200_REFIT_TRSM_V1")) != 0;
  static const bool bsolve_v2 = getenv("VBMC_B200_REFIT_BSOLVE_V2") && atoi(getenv("VBMC_B200_REFIT_BSOLVE_V2")) != 0;
  for (int attempt = 0; attempt < 10 && !active.empty(); ++attempt) {
    for (int s : active) {
      if (rr->Lchol[s]) {
        rr->sl[s] = rr->minsn2[s] * rr->mult[s];     // sl = sn2div*sn2_mult (:83)
        h_scale[s] = 1.0 / rr->sl[s];                // K_mat/(sn2div*sn2_mult)
        h_dscale[s] = 1.0 / rr->minsn2[s];           // diag(sn2/sn2div)
      } else {
        rr->sl[s] = 1.0;
        h_scale[s] = 1.0;
        h_dscale[s] = rr->mult[s];                   // K_mat + sn2_mult*diag(sn2) (:93)
      }
      h_ascale[s] = 1.0 / rr->sl[s];
    }
    const int nact = static_cast<int>(active.size());
    VB_CUDA(cudaMemcpyAsync(d_scale, h_scale.data(), sizeof(double) * S, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemcpyAsync(d_dscale, h_dscale.data(), sizeof(double) * S, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemcpyAsync(d_ascale, h_ascale.data(), sizeof(double) * S, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemcpyAsync(d_active, active.data(), sizeof(int) * nact, cudaMemcpyHostToDevice, st));
    VB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int) * S, st));
    auto factor_body = [&]() -> int {
      {
        dim3 grid(nb * (nb + 1) / 2, nact);
        KernelScope ks(c, "gram", st);
        gp_gram_kernel<<<grid, 256, sizeof(double) * 2 * D * TB, st>>>(g);
      }
      // 128-wide panels: two 64-steps share one big trailing update of depth 128.
      // LOOK-AHEAD: the panel work (diagonal-block factorisation on `nact` CTAs, row-panel solves) is latency bound and leaves
      // the GPU nearly idle.  After panel kb is done, only the two block rows of the NEXT panel are updated on the panel stream;
      // the rest of the trailing update runs on a second stream while the next panel is factored (they touch disjoint rows:
      // the update reads panel rows kb, kb+1 and writes rows >= kb+4; the next panel reads and writes rows kb+2, kb+3).
      static const bool la_env = !(getenv("VBMC_B200_REFIT_LOOKAHEAD") && atoi(getenv("VBMC_B200_REFIT_LOOKAHEAD")) == 0);
      // (measured: 6.40 -> 5.96 ms at c3, S=20 x N=2000; at N=4000 the update dominates and giving up SMs costs more than the
      // hidden panels return: 52.1 -> 54.4 ms, so it is used up to Np = 2560 only)
      const bool lookahead = la_env && !c->profiling && Np <= 2560;   // per-kernel event timing needs the serial order
      cudaStream_t side = c->stream2;
      auto panel = [&](int kb) -> int {
        for (int h = 0; h < 2 && kb + h < nb; ++h) {
          const int k = kb + h, nr = nb - k - 1;
          {  // (a fused potf2+trsm kernel was measured slower in round 1: 2.6 vs 2.2 ms at c3)
            KernelScope ks(c, "potrf_potf2", st);
            if (panel_v1)
              gp_potf2_kernel<<<nact, 256, 0, st>>>(g, k);
            else   // register-resident factorisation that also yields R_kk^-T
              gp_potf2i_kernel<<<nact, 256, 0, st>>>(g, k);
          }
          if (nr > 0) {
            dim3 grid(nr, nact);
            KernelScope ks(c, "potrf_trsm", st);
            if (panel_v1 || trsm_v1)
              gp_trsm_kernel<<<grid, 256, 0, st>>>(g, k);
            else   // the row panel as DMMA products with R_kk^-T
              gp_trsmg_kernel<<<grid, 256, TRSMG_SMEM, st>>>(g, k);
          }
          if (h == 0 && nr > 0) {  // update block row kb+1 only (needed by the second half of the panel)
            const int ch = update_chunk(nb, kb + 1, 1, nact, c->num_sms);
            dim3 grid(update_work_items(nb, kb + 1, 1, ch), nact);
            KernelScope ks(c, "potrf_update", st);
            gp_update_kernel<64><<<grid, 256, UPDATE_SMEM64, st>>>(g, kb * TB, kb + 1, 1, ch, grid.x);
          }
        }
        return VBMC_B200_OK;
      };
      // rows Ifirst.. with panel rows kb, kb+1; max_ctas > 0: persistent CTAs on at most that many SMs
      auto update128 = [&](int kb, int Ifirst, int Icount, cudaStream_t sx, int max_ctas) -> int {
        const int ch = update_chunk(nb, Ifirst, Icount, nact, c->num_sms);
        const int items = update_work_items(nb, Ifirst, Icount, ch);
        int gx = items;
        if (max_ctas > 0 && gx * nact > max_ctas) gx = max_ctas / nact > 0 ? max_ctas / nact : 1;
        dim3 grid(gx, nact);
        KernelScope ks(c, "potrf_update", sx);
        gp_update_kernel<128><<<grid, 256, UPDATE_SMEM128, sx>>>(g, kb * TB, Ifirst, Icount, ch, items);
        return VBMC_B200_OK;
      };
      // side-stream budget: num_sms - reserve persistent update CTAs.  0 = one per SM, the panel kernels' CTAs (35 and 72 KB of
      // shared memory) co-reside with them; measured at c3: reserve 28 -> 5.73 ms, 0 -> 5.30, -92 (240 CTAs) -> 5.69, no look-ahead 5.66
      static const int reserve = getenv("VBMC_B200_REFIT_PANEL_SMS") ? atoi(getenv("VBMC_B200_REFIT_PANEL_SMS")) : 0;
      bool side_busy = false;
      for (int kb = 0; kb < nb; kb += 2) {
        VB_TRY(panel(kb));
        const int nrest = nb - kb - 2;   // block rows behind the panel
        if (nrest <= 0) break;
        if (!lookahead) {
          VB_TRY(update128(kb, kb + 2, nrest, st, 0));
          continue;
        }
        if (side_busy) VB_CUDA(cudaStreamWaitEvent(st, c->ev_la_side, 0));   // the previous trailing update wrote these rows
        side_busy = false;
        const int la = nrest < 2 ? nrest : 2;
        VB_TRY(update128(kb, kb + 2, la, st, 0));
        if (nrest > la) {
          VB_CUDA(cudaEventRecord(c->ev_la_main, st));
          VB_CUDA(cudaStreamWaitEvent(side, c->ev_la_main, 0));
          VB_TRY(update128(kb, kb + 2 + la, nrest - la, side, c->num_sms - reserve));
          VB_CUDA(cudaEventRecord(c->ev_la_side, side));
          side_busy = true;
        }
      }
      if (side_busy) VB_CUDA(cudaStreamWaitEvent(st, c->ev_la_side, 0));
      return VBMC_B200_OK;
    };
    // ~100 dependent launches: replay them as one CUDA graph while shapes and buffers are unchanged
    std::vector<long long> key = {N, D, Np, S, nact, gd->Nhyp, gd->meanfun, reinterpret_cast<long long>(c->gpL.p),
                                  reinterpret_cast<long long>(c->gpWork.p), reinterpret_cast<long long>(c->gpX.p),
                                  reinterpret_cast<long long>(c->gpY.p), reinterpret_cast<long long>(c->gpHyp.p),
                                  reinterpret_cast<long long>(c->gpS2.p), gd->noisefun[0], gd->noisefun[1], gd->noisefun[2]};
    if (c->graphs_enabled && !c->profiling) {
      if (!(c->refit_graph && key == c->refit_key)) {
        if (c->refit_graph) { cudaGraphExecDestroy(c->refit_graph); c->refit_graph = nullptr; }
        cudaGraph_t graph = nullptr;
        const long long l0 = c->launches;
        VB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = factor_body();
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc == VBMC_B200_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&c->refit_graph, graph, 0) == cudaSuccess) {
          c->refit_key = key;
          c->refit_launches = c->launches - l0;
          c->launches = l0;
        } else {
          c->refit_graph = nullptr;
          cudaGetLastError();
        }
        if (graph) cudaGraphDestroy(graph);
      }
      if (c->refit_graph) {
        VB_CUDA(cudaGraphLaunch(c->refit_graph, st));
        c->launches += c->refit_launches;
      } else {
        VB_TRY(factor_body());
      }
    } else {
      VB_TRY(factor_body());
    }
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaMemcpyAsync(h_info.data(), d_info, sizeof(int) * S, cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    std::vector<int> next;
    for (int s : active)
      if (h_info[s] != 0) {
        rr->mult[s] *= 10.0;  // if p > 0; sn2_mult = sn2_mult*10 (:80)
        next.push_back(s);
      }
    if (attempt == 9 && !next.empty())
      VB_FAIL(VBMC_B200_EREFERENCE, "vbmc_b200:CholFailed: Cholesky failed for %d sample(s) after 10 jitter retries",
              static_cast<int>(next.size()));
    active.swap(next);
  }
  // back substitution + nlZ ingredients for all samples
  {
    KernelScope ks(c, "trsv", st);
    gp_nlzparts_kernel<<<S, 256, 0, st>>>(g, d_logdet, d_zz);
  }
  {
    KernelScope ks(c, "trsv", st);
    if (panel_v1)
      gp_bsolve_kernel<<<S, 1024, 0, st>>>(g, d_ascale, c->gpAlpha.d());
    else if (bsolve_v2)   // the diagonal blocks through the inverses the panel kernels left behind, one CTA per sample
      gp_bsolve2_kernel<<<S, 1024, 0, st>>>(g, d_ascale, c->gpAlpha.d());
    else {   // ... and P CTAs per sample that hand the solved blocks to each other through flags
      const int nbN = (N + TB - 1) / TB;
      int P = c->num_sms / S;
      P = P < 1 ? 1 : (P > nbN ? nbN : P);
      if (P > 8) P = 8;
      gp_bsolve3_kernel<<<dim3(P, S), 1024, 0, st>>>(g, d_ascale, c->gpAlpha.d(), d_flags, ++c->bsolve_epoch);
    }
  }
  VB_CUDA(cudaGetLastError());
  rr->logdet.assign(S, 0.0);
  rr->zz.assign(S, 0.0);
  VB_CUDA(cudaMemcpyAsync(rr->logdet.data(), d_logdet, sizeof(double) * S, cudaMemcpyDeviceToHost, st));
  VB_CUDA(cudaMemcpyAsync(rr->zz.data(), d_zz, sizeof(double) * S, cudaMemcpyDeviceToHost, st));
  VB_CUDA(cudaMemcpyAsync(h_info.data(), d_info, sizeof(int) * S, cudaMemcpyDeviceToHost, st));
  VB_CUDA(cudaStreamSynchronize(st));
  for (int s = 0; s < S; ++s)
    if (h_info[s] == -7) VB_FAIL(VBMC_B200_ECUDA, "vbmc_b200:backsolve: a CTA of the multi-CTA back substitution did not arrive (sample %d)", s);
  return VBMC_B200_OK;
}

// inv(A) tiles of sample `s` of the last refit: mode 0 -> (S_0..S_D, diagQ) on the host, mode 1 -> -inv(A) into dev_out
int run_invtiles_ld(vbmc_b200_ctx* c, int N, int D, int Nhyp, int Np, int s, double sl, int mode, std::vector<double>* sums,
                    std::vector<double>* diagQ, double* dev_out);
int run_invtiles(vbmc_b200_ctx* c, const vbmc_b200_gp_desc* gd, int s, double sl, int mode, std::vector<double>* sums,
                 std::vector<double>* diagQ, double* dev_out) {
  return run_invtiles_ld(c, gd->N, gd->D, gd->Nhyp, (gd->N + 1 + TB - 1) / TB * TB, s, sl, mode, sums, diagQ, dev_out);
}
// the same for a resident posterior of N points whose factors have leading dimension Np (grown by rank-one updates)
int run_invtiles_ld(vbmc_b200_ctx* c, int N, int D, int Nhyp, int Np, int s, double sl, int mode, std::vector<double>* sums,
                    std::vector<double>* diagQ, double* dev_out) {
  const int nb = (N + TB - 1) / TB, ntiles = nb * (nb + 1) / 2;
  if (D > 24) VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:nlz_grad: D=%d > 24 is not supported by this build", D);
  vb::DevBuf& wk = c->varWork;
  const size_t nX = static_cast<size_t>(N) * N;
  VB_TRY(wk.reserve(sizeof(double) * (nX + static_cast<size_t>(ntiles) * (D + 1) + N)));
  double* X = wk.d();
  double* partial = X + nX;
  double* dq = partial + static_cast<size_t>(ntiles) * (D + 1);
  const double* R = c->gpL.d() + static_cast<size_t>(s) * Np * Np;
  VB_TRY(vb::run_factor_inverse(c, N, Np, R, X));
  InvArgs a;
  a.N = N; a.D = D; a.mode = mode;
  a.Xinv = X; a.Xc = c->gpX.d(); a.hyp = c->gpHyp.d() + static_cast<size_t>(s) * Nhyp;
  a.alpha = c->gpAlpha.d() + static_cast<size_t>(s) * N;
  a.inv_sl = 1.0 / sl;
  a.partial = partial; a.diagQ = dq; a.outL = dev_out;
  const size_t smem = sizeof(double) * (2 * TB * TLD + 2 * static_cast<size_t>(D) * TB + static_cast<size_t>(D + 1) * 256);
  VB_CUDA(cudaFuncSetAttribute(gp_invtile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  {
    KernelScope ks(c, "potri_tiles", c->stream);
    gp_invtile_kernel<<<ntiles, 256, smem, c->stream>>>(a);
  }
  VB_CUDA(cudaGetLastError());
  if (mode == 0) {
    std::vector<double> hp(static_cast<size_t>(ntiles) * (D + 1));
    diagQ->assign(N, 0.0);
    VB_CUDA(cudaMemcpyAsync(hp.data(), partial, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost, c->stream));
    VB_CUDA(cudaMemcpyAsync(diagQ->data(), dq, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
    VB_CUDA(cudaStreamSynchronize(c->stream));
    sums->assign(D + 1, 0.0);
    for (int t = 0; t < ntiles; ++t)
      for (int d = 0; d <= D; ++d) (*sums)[d] += hp[static_cast<size_t>(t) * (D + 1) + d];
  }
  return VBMC_B200_OK;
}

int check_desc(const vbmc_b200_gp_desc* g, int* Ncov, int* Nnoise, int* Nmean, const char* who) {
  if (!g || !g->X || !g->hyp || !g->y) VB_FAIL(VBMC_B200_EINVAL, "%s: X, y and hyp are required", who);
  if (g->N <= 0 || g->D <= 0 || g->S <= 0) VB_FAIL(VBMC_B200_EINVAL, "%s: N, D, S must be positive", who);
  if (g->covfun != 1) VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:UnsupportedCovFun: only covfun=1 (SE-ARD) is in scope");
  *Ncov = g->D + 1;
  *Nnoise = (g->noisefun[0] == 1) + (g->noisefun[1] == 2) + 2 * (g->noisefun[2] == 1);
  switch (g->meanfun) {
    case 0: *Nmean = 0; break;
    case 1: *Nmean = 1; break;
    case 4: *Nmean = 1 + 2 * g->D; break;
    default: VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:UnsupportedMeanFun: this build supports meanfun 0, 1, 4 (got %d)", g->meanfun);
  }
  if ((g->noisefun[1] == 1 || g->noisefun[1] == 2) && !g->s2) VB_FAIL(VBMC_B200_EINVAL, "%s: noisefun(2) > 0 needs s2", who);
  return VBMC_B200_OK;
}

}  // namespace

extern "C" {

int vbmc_b200_gp_post(vbmc_b200_ctx* c, const vbmc_b200_gp_desc* gd, double* alpha, double* L, double* sW1,
                      double* sn2_mult, int* Lchol) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  c->gp_tag = 0;  // the resident posterior is about to change: a caller's fingerprint of it no longer holds
  int Ncov, Nnoise, Nmean;
  VB_TRY(check_desc(gd, &Ncov, &Nnoise, &Nmean, "gp_post"));
  if (gd->Nhyp != Ncov + Nnoise + Nmean)
    VB_FAIL(VBMC_B200_EREFERENCE,
            "gplite_post:dimmismatch: Number of hyperparameters mismatched with GP model specification (Nhyp=%d, expected %d).",
            gd->Nhyp, Ncov + Nnoise + Nmean);
  VB_CUDA(cudaSetDevice(c->device));
  c->gp_ready = false;
  RefitResult rr;
  VB_TRY(refit_core(c, gd, Ncov, Nnoise, Nmean, &rr));
  const int N = gd->N, S = gd->S;
  const int Np = (N + 1 + TB - 1) / TB * TB;
  std::vector<double> h_sw(S);
  for (int s = 0; s < S; ++s) h_sw[s] = 1.0 / sqrt(rr.minsn2[s] * rr.mult[s]);  // post.sW (:281)
  if (alpha) VB_CUDA(cudaMemcpyAsync(alpha, c->gpAlpha.p, sizeof(double) * S * N, cudaMemcpyDeviceToHost, c->stream));
  if (L) {
    VB_TRY(c->gpWork.reserve(c->gpWork.cap));  // keep
    vb::DevBuf tmp;
    VB_TRY(tmp.reserve(sizeof(double) * static_cast<size_t>(S) * N * N));
    dim3 grid((N + 255) / 256, N, S);
    {
      KernelScope ks(c, "extract", c->stream);
      gp_extract_kernel<<<grid, 256, 0, c->stream>>>(c->gpL.d(), Np, N, tmp.d(), 0);
    }
    // low-noise samples: post.L = -inv(K + sn2_mult*diag(sn2))  (gplite_core.m:96-99)
    for (int s = 0; s < S; ++s)
      if (!rr.Lchol[s]) VB_TRY(run_invtiles(c, gd, s, 1.0, 1, nullptr, nullptr, tmp.d() + static_cast<size_t>(s) * N * N));
    VB_CUDA(cudaMemcpyAsync(L, tmp.p, sizeof(double) * static_cast<size_t>(S) * N * N, cudaMemcpyDeviceToHost, c->stream));
    VB_CUDA(cudaStreamSynchronize(c->stream));
    tmp.release();
  }
  VB_CUDA(cudaStreamSynchronize(c->stream));
  if (sW1) memcpy(sW1, h_sw.data(), sizeof(double) * S);
  if (sn2_mult) memcpy(sn2_mult, rr.mult.data(), sizeof(double) * S);
  if (Lchol) memcpy(Lchol, rr.Lchol.data(), sizeof(int) * S);
  // the refit posterior becomes the attached GP
  c->gpLchol = rr.Lchol;
  c->gpLfactor.assign(S, 1);  // the device keeps R with R'R = K + sn2_mult*diag(sn2) for the low-noise samples too
  c->gpSn2mult = rr.mult;
  c->gpHypHost.assign(gd->hyp, gd->hyp + static_cast<size_t>(S) * gd->Nhyp);
  for (int i = 0; i < 3; ++i) c->gp_noisefun[i] = gd->noisefun[i];
  c->gpHasL = true;
  c->gpLd = Np;
  c->gp.N = N; c->gp.D = gd->D; c->gp.S = S; c->gp.Nhyp = gd->Nhyp;
  c->gp.Ncov = Ncov; c->gp.Nnoise = Nnoise; c->gp.Nmean = Nmean; c->gp.meanfun = gd->meanfun;
  c->gp.X = c->gpX.d(); c->gp.hyp = c->gpHyp.d(); c->gp.alpha = c->gpAlpha.d();
  VB_TRY(vb::gp_upload_derived(c, gd, Ncov, Nnoise, h_sw.data(), rr.Lchol.data()));
  c->gp_ready = true;
  return VBMC_B200_OK;
}

// gp.post(s).L of the resident posterior in the reference's representation (gplite_core.m:67-100): the upper Cholesky factor
// (Lchol) or -inv(K + sn2_mult*diag(sn2)) (low noise) -- for a low-noise sample whose factor lives on the device the inverse
// is built here (R^-T on the tensor path, then X'X tiles).  L_out: N x N, column-major.
int vbmc_b200_gp_get_factor(vbmc_b200_ctx* c, int s, double* L_out) {
  if (!c || !L_out) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  if (!c->gp_ready || !c->gpHasL) VB_FAIL(VBMC_B200_ESTATE, "gp_get_L: no posterior with factors is resident (gp_attach with L, or gp_post)");
  const int N = c->gp.N, S = c->gp.S, ld = c->gpLd;
  if (s < 0 || s >= S) VB_FAIL(VBMC_B200_EINVAL, "gp_get_L: sample %d of %d", s, S);
  VB_CUDA(cudaSetDevice(c->device));
  vb::DevBuf tmp;
  VB_TRY(tmp.reserve(sizeof(double) * static_cast<size_t>(N) * N));
  if (c->gpLfactor[s] && !c->gpLchol[s]) {
    VB_TRY(run_invtiles_ld(c, N, c->gp.D, c->gp.Nhyp, ld, s, 1.0, 1, nullptr, nullptr, tmp.d()));
  } else {
    KernelScope ks(c, "extract", c->stream);
    gp_extract_kernel<<<dim3((N + 255) / 256, N, 1), 256, 0, c->stream>>>(c->gpL.d() + static_cast<size_t>(s) * ld * ld, ld, N, tmp.d(),
                                                                          c->gpLfactor[s] ? 0 : 1);
    VB_CUDA(cudaGetLastError());
  }
  VB_CUDA(cudaMemcpyAsync(L_out, tmp.p, sizeof(double) * static_cast<size_t>(N) * N, cudaMemcpyDeviceToHost, c->stream));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  tmp.release();
  return VBMC_B200_OK;
}

// nlZ of sample s from the factorisation results, minus the log hyper-prior (gplite_core.m:193, gplite_nlZ.m:57-66,
// gplite_hypprior.m:24-58; O(Nhyp) on the host)
static double nlz_value(const vbmc_b200_gp_desc* gd, const vbmc_b200_hprior* hprior, const RefitResult& rr, int s) {
  const int N = gd->N;
  // nlZ = (y-m)'*alpha/2 + sum(log(diag(L))) + N*log(2*pi*sl)/2   (:193);  (y-m)'alpha = z'z/sl
  double v = 0.5 * rr.zz[s] / rr.sl[s] + rr.logdet[s] + 0.5 * N * log(2.0 * 3.14159265358979323846 * rr.sl[s]);
  if (hprior && hprior->mu && hprior->sigma) {
    const double* hyp = gd->hyp + static_cast<size_t>(s) * gd->Nhyp;
    double lp = 0.0;
    for (int i = 0; i < gd->Nhyp; ++i) {
      const double mu = hprior->mu[i], sg = fabs(hprior->sigma[i]);
      const double df = hprior->df ? hprior->df[i] : 7.0;
      if (!isfinite(mu) || !isfinite(sg)) continue;  // uniform
      const double z2 = ((hyp[i] - mu) / sg) * ((hyp[i] - mu) / sg);
      if (df == 0.0 || !isfinite(df))
        lp -= 0.5 * (log(2.0 * 3.14159265358979323846 * sg * sg) + z2);
      else if (df > 0.0)
        lp += lgamma(0.5 * (df + 1.0)) - lgamma(0.5 * df) - 0.5 * log(3.14159265358979323846 * df) - log(sg) -
              0.5 * (df + 1.0) * log1p(z2 / df);
    }
    v -= lp;
  }
  return v;
}

int vbmc_b200_gp_nlz(vbmc_b200_ctx* c, const vbmc_b200_gp_desc* gd, const vbmc_b200_hprior* hprior, double* nlZ,
                     double* dnlZ) {
  if (!c || !nlZ) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  c->gp_tag = 0;  // the resident posterior is about to change: a caller's fingerprint of it no longer holds
  int Ncov, Nnoise, Nmean;
  VB_TRY(check_desc(gd, &Ncov, &Nnoise, &Nmean, "gp_nlz"));
  if (gd->Nhyp != Ncov + Nnoise + Nmean)
    VB_FAIL(VBMC_B200_EREFERENCE,
            "gplite_nlZ:dimmismatch: Number of hyperparameters mismatched with dimension of training inputs.");
  if (gd->S != 1)
    VB_FAIL(VBMC_B200_EREFERENCE,
            "gplite_nlZ:NoSampling: Computation of the log marginal likelihood is available only for one-sample "
            "hyperparameter inputs.");
  VB_CUDA(cudaSetDevice(c->device));
  c->gp_ready = false;  // the GP buffers are reused
  RefitResult rr;
  VB_TRY(refit_core(c, gd, Ncov, Nnoise, Nmean, &rr));
  const int N = gd->N, D = gd->D;
  const double v = nlz_value(gd, hprior, rr, 0);
  *nlZ = v;
  if (!dnlZ) return VBMC_B200_OK;
  // ---- gradient (gplite_core.m:226-261) ----
  std::vector<double> sums, dq, al(N);
  VB_TRY(run_invtiles(c, gd, 0, rr.sl[0], 0, &sums, &dq, nullptr));
  VB_CUDA(cudaMemcpyAsync(al.data(), c->gpAlpha.p, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  const double* h = gd->hyp;
  for (int i = 0; i < gd->Nhyp; ++i) dnlZ[i] = 0.0;
  for (int i = 0; i < D; ++i) dnlZ[i] = 0.5 * sums[i];   // sum(sum(Q.*K_mat.*sq_dist(X(:,i)'/ell(i))))/2
  dnlZ[D] = sums[D];                                      // sum(sum(Q.*(2*K_mat)))/2
  // noise hyper-parameters: 0.5*sn2_mult*sum(dsn2(:,i).*diag(Q))  (:242-251; the scalar case is the same sum)
  {
    const double* hn = h + Ncov;
    int idx = 0;
    const double mult = rr.mult[0];
    if (gd->noisefun[0] == 1) {
      const double d1 = 2.0 * exp(2.0 * hn[idx]);
      double acc = 0.0;
      for (int n = 0; n < N; ++n) acc += dq[n];
      dnlZ[Ncov + idx] = 0.5 * mult * d1 * acc;
      ++idx;
    }
    if (gd->noisefun[1] == 2) {
      double acc = 0.0;
      for (int n = 0; n < N; ++n) acc += exp(hn[idx]) * gd->s2[n] * dq[n];
      dnlZ[Ncov + idx] = 0.5 * mult * acc;
      ++idx;
    }
    if (gd->noisefun[2] == 1) {
      const double yth = hn[idx], w2 = exp(2.0 * hn[idx + 1]);
      double a0 = 0.0, a1 = 0.0;
      for (int n = 0; n < N; ++n) {
        const double zz = fmax(0.0, yth - gd->y[n]);
        a0 += 2.0 * w2 * (yth - gd->y[n]) * (zz > 0.0 ? 1.0 : 0.0) * dq[n];
        a1 += 2.0 * w2 * zz * zz * dq[n];
      }
      dnlZ[Ncov + idx] = 0.5 * mult * a0;
      dnlZ[Ncov + idx + 1] = 0.5 * mult * a1;
    }
  }
  // mean function: -dm'*alpha  (:254-261; gplite_meanfun.m cases 1 and 4)
  {
    const double* hm = h + Ncov + Nnoise;
    double* g = dnlZ + Ncov + Nnoise;
    if (gd->meanfun == 1 || gd->meanfun == 4) {
      double acc = 0.0;
      for (int n = 0; n < N; ++n) acc += al[n];
      g[0] = -acc;
    }
    if (gd->meanfun == 4)
      for (int d = 0; d < D; ++d) {
        const double xm = hm[1 + d], om = exp(hm[1 + D + d]);
        double a1 = 0.0, a2 = 0.0;
        for (int n = 0; n < N; ++n) {
          const double df = gd->X[static_cast<size_t>(d) * N + n] - xm;
          a1 += df / (om * om) * al[n];
          a2 += (df / om) * (df / om) * al[n];
        }
        g[1 + d] = -a1;
        g[1 + D + d] = -a2;
      }
  }
  if (hprior && hprior->mu && hprior->sigma) {  // dnlZ = dnlZ - dP
    for (int i = 0; i < gd->Nhyp; ++i) {
      const double mu = hprior->mu[i], sg = fabs(hprior->sigma[i]);
      const double df = hprior->df ? hprior->df[i] : 7.0;
      if (!isfinite(mu) || !isfinite(sg)) continue;
      const double dlt = gd->hyp[i] - mu, z2 = (dlt / sg) * (dlt / sg);
      double dlp = 0.0;
      if (df == 0.0 || !isfinite(df)) dlp = -dlt / (sg * sg);
      else if (df > 0.0) dlp = -(df + 1.0) / df / (1.0 + z2 / df) * dlt / (sg * sg);
      dnlZ[i] -= dlp;
    }
  }
  return VBMC_B200_OK;
}

// nlZ(s) = gplite_nlZ(hyp(:,s), gp, hprior) for S hyper-parameter vectors in ONE batched factorisation (value only).
// The reference evaluates them one call at a time: the space-filling design of gplite_train.m:200-204 (fminfill ->
// gpoptimize_fun -> gplite_nlZ, Ninit vectors) and every slice-sampler step (gplite_train.m:318-330).
int vbmc_b200_gp_nlz_batch(vbmc_b200_ctx* c, const vbmc_b200_gp_desc* gd, const vbmc_b200_hprior* hprior, double* nlZ) {
  if (!c || !nlZ) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  c->gp_tag = 0;  // the resident posterior is about to change: a caller's fingerprint of it no longer holds
  int Ncov, Nnoise, Nmean;
  VB_TRY(check_desc(gd, &Ncov, &Nnoise, &Nmean, "gp_nlz_batch"));
  if (gd->Nhyp != Ncov + Nnoise + Nmean)
    VB_FAIL(VBMC_B200_EREFERENCE,
            "gplite_nlZ:dimmismatch: Number of hyperparameters mismatched with dimension of training inputs.");
  VB_CUDA(cudaSetDevice(c->device));
  c->gp_ready = false;  // the GP buffers are reused
  RefitResult rr;
  VB_TRY(refit_core(c, gd, Ncov, Nnoise, Nmean, &rr));
  for (int s = 0; s < gd->S; ++s) nlZ[s] = nlz_value(gd, hprior, rr, s);
  return VBMC_B200_OK;
}

}  // extern "C"
