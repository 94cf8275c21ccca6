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
// grid (ntiles_upper, nact), 256 threads, each thread 4x4 entries.
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
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
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
        double sq = 0.0;
        for (int d = 0; d < D; ++d) {
          const double df = xi[d * TB + r] - xj[d * TB + c];
          sq = fma(df, df, sq);
        }
        v = sf2 * exp(-0.5 * sq);                                    // K_mat (:55-56), scaled
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

// ---- panel step kb: factor the diagonal block, then R_kJ = R_kk^-T A_kJ for the blocks to the right -----
// grid (max(1, nb-kb-1), nact).  Every CTA refactors the 64x64 diagonal block (cheap) so that no
// inter-CTA dependency exists inside the launch; CTA x==0 writes it back.
__global__ void __launch_bounds__(256) gp_panel_kernel(const GpBatch g, int kb) {
  extern __shared__ __align__(16) double psm[];
  typedef double (*Blk)[TB + 1];
  Blk A = reinterpret_cast<Blk>(psm);                            // A[c][r]: column c, row r (upper part), unscaled
  Blk R = reinterpret_cast<Blk>(psm + TB * (TB + 1));            // factor
  Blk Bx = reinterpret_cast<Blk>(psm + 2 * TB * (TB + 1));       // right-hand block, Bx[c][r]
  __shared__ int bad;
  const int s = g.active[blockIdx.y];
  const int Np = g.Np, N = g.N, tid = threadIdx.x;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  const int k0 = kb * TB;
  if (tid == 0) bad = 0;
  for (int i = tid; i < TB * TB; i += 256) {
    const int c = i / TB, r = i - c * TB;
    A[c][r] = Ms[static_cast<size_t>(k0 + c) * Np + k0 + r];
    R[c][r] = 0.0;
  }
  const int nJ = Np / TB - kb - 1;
  const int jb = kb + 1 + blockIdx.x;
  const bool have_rhs = blockIdx.x < nJ;
  if (have_rhs)
    for (int i = tid; i < TB * TB; i += 256) {
      const int c = i / TB, r = i - c * TB;
      Bx[c][r] = Ms[static_cast<size_t>(jb * TB + c) * Np + k0 + r];
    }
  __syncthreads();
  // right-looking elimination, one barrier per pivot: rows >= N (rhs/padding rows) are unit rows
  for (int p = 0; p < TB; ++p) {
    const bool unit = (k0 + p) >= N;
    const double d = A[p][p];
    if (!unit && !(d > 0.0) && tid == 0 && bad == 0) bad = k0 + p + 1;
    const double id = unit ? 0.0 : 1.0 / d;
    const double isq = unit ? 0.0 : rsqrt(d);
    // R[p][j] = a_pj / sqrt(d)
    if (tid < TB) {
      const int j = tid;
      if (j >= p) R[j][p] = unit ? (j == p ? 1.0 : 0.0) : A[j][p] * isq;
    }
    // trailing update of the diagonal block: A[i][j] -= a_pi a_pj / d,  p < i <= j
    const int rem = TB - 1 - p;
    for (int e = tid; e < rem * rem; e += 256) {
      const int ii = e / rem, jj = e - ii * rem;
      if (ii <= jj) {
        const int i = p + 1 + ii, j = p + 1 + jj;
        A[j][i] -= A[i][p] * A[j][p] * id;
      }
    }
    __syncthreads();
  }
  if (tid == 0 && bad != 0) atomicCAS(&g.info[s], 0, bad);
  if (blockIdx.x == 0)
    for (int i = tid; i < TB * TB; i += 256) {
      const int c = i / TB, r = i - c * TB;
      if (r <= c) Ms[static_cast<size_t>(k0 + c) * Np + k0 + r] = R[c][r];
    }
  if (!have_rhs) return;
  // forward substitution R' X = B, row by row; one barrier per row
  for (int p = 0; p < TB; ++p) {
    const double ir = 1.0 / R[p][p];
    if (tid < TB) Bx[tid][p] *= ir;  // x_p for column tid
    __syncthreads();
    const int rem = TB - 1 - p;
    for (int e = tid; e < rem * TB; e += 256) {
      const int rr = e / TB, c = e - rr * TB;
      const int r = p + 1 + rr;
      Bx[c][r] -= R[r][p] * Bx[c][p];
    }
    __syncthreads();
  }
  for (int i = tid; i < TB * TB; i += 256) {
    const int c = i / TB, r = i - c * TB;
    Ms[static_cast<size_t>(jb * TB + c) * Np + k0 + r] = Bx[c][r];
  }
}

// ---- trailing update on the FP64 tensor path:  C_IJ -= P_I' P_J  (I <= J, blocks right of kb) ----
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// grid (ntile_pairs, nact), 256 threads = 8 warps; warp (wm, wn) owns a 16 x 32 patch of the 64x64 tile.
__global__ void __launch_bounds__(256) gp_update_kernel(const GpBatch g, int kb) {
  extern __shared__ __align__(16) double usm[];
  double* PI = usm;             // PI[m*TLD + k] = P_I(k, m)
  double* PJ = usm + TB * TLD;
  const int s = g.active[blockIdx.y];
  const int Np = g.Np, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nr = Np / TB - kb - 1;  // blocks to the right of kb
  int t = blockIdx.x, ii = 0;
  while (t >= nr - ii) { t -= nr - ii; ++ii; }
  const int I = kb + 1 + ii, J = I + t;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  const int k0 = kb * TB;
  for (int i = tid; i < TB * TB; i += 256) {
    const int c = i / TB, k = i - c * TB;
    PI[c * TLD + k] = Ms[static_cast<size_t>(I * TB + c) * Np + k0 + k];
    PJ[c * TLD + k] = Ms[static_cast<size_t>(J * TB + c) * Np + k0 + k];
  }
  __syncthreads();
  const int wm = warp >> 1, wn = warp & 1;
  const int g4 = lane >> 2, t4 = lane & 3;
  double acc[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
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
  // C(m, n) at column-major (I*64+m) + (J*64+n)*Np ; thread holds rows g4, cols 2*t4+{0,1} of each 8x8 tile
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int m = I * TB + wm * 16 + a * 8 + g4;
      const int n = J * TB + wn * 32 + b * 8 + 2 * t4;
      double* c0 = Ms + static_cast<size_t>(n) * Np + m;
      c0[0] -= acc[a][b][0];
      c0[Np] -= acc[a][b][1];
    }
}

// ---- back substitution R x = z (z = column N of the factored buffer), alpha = x * ascale; one CTA per sample.
// Also returns sum(log(diag R)) and z'z (nlZ ingredients, gplite_core.m:193).
__global__ void __launch_bounds__(256) gp_backsolve_kernel(const GpBatch g, const double* ascale, double* alpha,
                                                           double* logdet, double* zz) {
  __shared__ double Rk[TB][TB + 1];
  __shared__ double xk[TB];
  __shared__ double part[256];
  const int s = g.active[blockIdx.x];
  const int Np = g.Np, N = g.N, tid = threadIdx.x;
  double* Ms = g.M + static_cast<size_t>(s) * Np * Np;
  double* z = Ms + static_cast<size_t>(N) * Np;  // column N (overwritten by the solution)
  // nlZ ingredients
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
  __syncthreads();
  const int nbN = (N + TB - 1) / TB;
  for (int kb = nbN - 1; kb >= 0; --kb) {
    const int k0 = kb * TB;
    for (int i = tid; i < TB * TB; i += 256) {
      const int c = i / TB, r = i - c * TB;
      Rk[c][r] = (k0 + c < N && k0 + r < N) ? Ms[static_cast<size_t>(k0 + c) * Np + k0 + r] : (c == r ? 1.0 : 0.0);
    }
    if (tid < TB) xk[tid] = (k0 + tid < N) ? z[k0 + tid] : 0.0;
    __syncthreads();
    // 64x64 upper back substitution by one warp-sized group of threads, column oriented
    for (int p = TB - 1; p >= 0; --p) {
      if (tid == 0) xk[p] = xk[p] / Rk[p][p];
      __syncthreads();
      if (tid < p) xk[tid] -= Rk[p][tid] * xk[p];
      __syncthreads();
    }
    if (tid < TB && k0 + tid < N) z[k0 + tid] = xk[tid];
    // z_i -= sum_{j in block} R(i, j) x_j   for i < k0   (coalesced over i)
    for (int i = tid; i < k0; i += 256) {
      double acc = 0.0;
#pragma unroll 8
      for (int j = 0; j < TB; ++j) acc = fma(Ms[static_cast<size_t>(k0 + j) * Np + i], xk[j], acc);
      z[i] -= acc;
    }
    __syncthreads();
  }
  const double sc = ascale[s];
  for (int i = tid; i < N; i += 256) alpha[static_cast<size_t>(s) * N + i] = z[i] * sc;
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
  const size_t nwork = 2 * static_cast<size_t>(S) * N + 6 * S;
  VB_TRY(c->gpWork.reserve(sizeof(double) * nwork + sizeof(int) * 2 * S + 64));
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
  int* d_info = reinterpret_cast<int*>(d_zz + S);
  int* d_active = d_info + S;
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
  const int PANEL_SMEM = 3 * TB * (TB + 1) * sizeof(double), UPDATE_SMEM = 2 * TB * TLD * sizeof(double);
  VB_CUDA(cudaFuncSetAttribute(gp_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM));
  VB_CUDA(cudaFuncSetAttribute(gp_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UPDATE_SMEM));
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
    {
      dim3 grid(nb * (nb + 1) / 2, nact);
      KernelScope ks(c, "gram", st);
      gp_gram_kernel<<<grid, 256, sizeof(double) * 2 * D * TB, st>>>(g);
    }
    for (int kb = 0; kb < nb; ++kb) {
      const int nr = nb - kb - 1;
      {
        dim3 grid(nr > 0 ? nr : 1, nact);
        KernelScope ks(c, "potrf_panel", st);
        gp_panel_kernel<<<grid, 256, PANEL_SMEM, st>>>(g, kb);
      }
      if (nr > 0) {
        dim3 grid(nr * (nr + 1) / 2, nact);
        KernelScope ks(c, "potrf_update", st);
        gp_update_kernel<<<grid, 256, UPDATE_SMEM, st>>>(g, kb);
      }
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
  for (int s = 0; s < S; ++s) h_info[s] = s;
  VB_CUDA(cudaMemcpyAsync(d_active, h_info.data(), sizeof(int) * S, cudaMemcpyHostToDevice, st));
  {
    KernelScope ks(c, "trsv", st);
    gp_backsolve_kernel<<<S, 256, 0, st>>>(g, d_ascale, c->gpAlpha.d(), d_logdet, d_zz);
  }
  VB_CUDA(cudaGetLastError());
  rr->logdet.assign(S, 0.0);
  rr->zz.assign(S, 0.0);
  VB_CUDA(cudaMemcpyAsync(rr->logdet.data(), d_logdet, sizeof(double) * S, cudaMemcpyDeviceToHost, st));
  VB_CUDA(cudaMemcpyAsync(rr->zz.data(), d_zz, sizeof(double) * S, cudaMemcpyDeviceToHost, st));
  VB_CUDA(cudaStreamSynchronize(st));
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
  for (int s = 0; s < S; ++s)
    if (!rr.Lchol[s])
      VB_FAIL(VBMC_B200_EUNSUPPORTED,
              "vbmc_b200:NotYet: min(sn2) < 1e-6 selects the explicit-inverse posterior (gplite_core.m:86-100), not built yet");
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
  c->gpSn2mult = rr.mult;
  c->gpHasL = true;
  c->gpLd = Np;
  c->gp.N = N; c->gp.D = gd->D; c->gp.S = S; c->gp.Nhyp = gd->Nhyp;
  c->gp.Ncov = Ncov; c->gp.Nnoise = Nnoise; c->gp.Nmean = Nmean; c->gp.meanfun = gd->meanfun;
  c->gp.X = c->gpX.d(); c->gp.hyp = c->gpHyp.d(); c->gp.alpha = c->gpAlpha.d();
  VB_TRY(vb::gp_upload_derived(c, gd, Ncov, Nnoise, h_sw.data()));
  c->gp_ready = true;
  return VBMC_B200_OK;
}

int vbmc_b200_gp_nlz(vbmc_b200_ctx* c, const vbmc_b200_gp_desc* gd, const vbmc_b200_hprior* hprior, double* nlZ,
                     double* dnlZ) {
  if (!c || !nlZ) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  int Ncov, Nnoise, Nmean;
  VB_TRY(check_desc(gd, &Ncov, &Nnoise, &Nmean, "gp_nlz"));
  if (gd->Nhyp != Ncov + Nnoise + Nmean)
    VB_FAIL(VBMC_B200_EREFERENCE,
            "gplite_nlZ:dimmismatch: Number of hyperparameters mismatched with dimension of training inputs.");
  if (gd->S != 1)
    VB_FAIL(VBMC_B200_EREFERENCE,
            "gplite_nlZ:NoSampling: Computation of the log marginal likelihood is available only for one-sample "
            "hyperparameter inputs.");
  if (dnlZ)
    VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:NotYet: gplite_nlZ gradient (gplite_core.m:226-261) is not built yet");
  VB_CUDA(cudaSetDevice(c->device));
  c->gp_ready = false;  // the GP buffers are reused
  RefitResult rr;
  VB_TRY(refit_core(c, gd, Ncov, Nnoise, Nmean, &rr));
  if (!rr.Lchol[0])
    VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:NotYet: low-noise branch (gplite_core.m:86-100) is not built yet");
  const int N = gd->N;
  // nlZ = (y-m)'*alpha/2 + sum(log(diag(L))) + N*log(2*pi*sl)/2   (:193);  (y-m)'alpha = z'z/sl
  double v = 0.5 * rr.zz[0] / rr.sl[0] + rr.logdet[0] + 0.5 * N * log(2.0 * 3.14159265358979323846 * rr.sl[0]);
  if (hprior && hprior->mu && hprior->sigma) {  // gplite_hypprior.m:24-58 (host: O(Nhyp))
    double lp = 0.0;
    for (int i = 0; i < gd->Nhyp; ++i) {
      const double mu = hprior->mu[i], sg = fabs(hprior->sigma[i]);
      const double df = hprior->df ? hprior->df[i] : 7.0;
      if (!isfinite(mu) || !isfinite(sg)) continue;  // uniform
      const double z2 = ((gd->hyp[i] - mu) / sg) * ((gd->hyp[i] - mu) / sg);
      if (df == 0.0 || !isfinite(df))
        lp -= 0.5 * (log(2.0 * 3.14159265358979323846 * sg * sg) + z2);
      else if (df > 0.0)
        lp += lgamma(0.5 * (df + 1.0)) - lgamma(0.5 * df) - 0.5 * log(3.14159265358979323846 * df) - log(sg) -
              0.5 * (df + 1.0) * log1p(z2 / df);
    }
    v -= lp;
  }
  *nlZ = v;
  return VBMC_B200_OK;
}

}  // extern "C"
