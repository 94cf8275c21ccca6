// Variance of the expected log-joint (reference: misc/gplogjoint.m:273-339, 349-350, 398-407).
//
//   J_jk(s) = exp(lnnf_jk - 0.5*sum(delta_jk.^2)) - z_k * (L\(L'\z_j')) / sn2_eff           (:320-324)
// With V = L'\Z (Z = [z_1 .. z_K], N x K) the quadratic forms are Gram entries: z_k K^-1 z_j = V_j.V_k / sn2_eff,
// so the K(K+1)/2 pairs of triangular solves per sample of the reference collapse to ONE forward substitution
// with K right-hand sides + a K x K Gram per sample.
//   compute_var == 2 (diagonal, :273-304) needs only G_kk; compute_var == 1 the full matrix.
// Kernels: glj_z_kernel (Z), var_fwd_kernel (V = R'\Z, R upper factor, 8 RHS per CTA resident in shared
// memory, factor columns streamed coalesced from L2), var_gram_kernel (V'V), var_final_kernel (J_sjk, varF_s).
#include <math.h>

#include <stdlib.h>

#include "common.cuh"

namespace vb {

struct VarArgs {
  int N, D, K, S, ld;
  size_t Lstride;     // doubles between the factors of consecutive samples
  const double* L;    // upper factors R (R'R = K/sl + ...), column-major, leading dimension ld
  GpDev gp;
  VpDev vp;
  double* Z;          // [S][K][N]  z, then V in place
  double* G;          // [S][K][K]  V_j . V_k
  double* J;          // [S][K][K]  J_sjk (j fastest == MATLAB (s,j,k) after host transpose)
  double* varFs;      // [S]
  int full;           // 1: full matrix, 0: diagonal only
  int unit_rhs;       // 1: right-hand sides are the identity columns k0.. (inverse of the factor), Z is output only
  const int* isfac;   // [S] or null: 1 = L is the Cholesky factor (Lchol), 0 = L is -inv(K + diag) (gplite_core.m:96-99)
  double* W;          // [S][K][N]  K^-1 z_k of the samples with isfac == 0
};

// z_k(n) = exp(lnnf_k - 0.5*sum_d ((mu_kd - X_nd)/tau_kd)^2)   (gplogjoint.m:164-167); grid (K, S)
__global__ void __launch_bounds__(256) glj_z_kernel(const VarArgs a) {
  __shared__ double s_mu[32], s_itau[32], s_lnnf;
  const int k = blockIdx.x, s = blockIdx.y, tid = threadIdx.x, D = a.D, N = a.N;
  const double sigk = a.vp.sigma[k];
  if (tid < D) {
    const double lam = a.vp.lambda[tid], ell = a.gp.ell[s * D + tid], dl = a.vp.delta[tid];
    s_mu[tid] = a.vp.mu[k * D + tid];
    s_itau[tid] = 1.0 / sqrt(sigk * sigk * lam * lam + ell * ell + dl * dl);
  }
  __syncthreads();
  if (tid == 0) {
    double slt = 0.0;
    for (int d = 0; d < D; ++d) slt -= log(s_itau[d]);
    s_lnnf = a.gp.lnc[s] - slt;
  }
  __syncthreads();
  double* z = a.Z + (static_cast<size_t>(s) * a.K + k) * N;
  for (int n = tid; n < N; n += 256) {
    double ss = 0.0;
    for (int d = 0; d < D; ++d) {
      const double dl = (s_mu[d] - a.gp.X[static_cast<size_t>(d) * N + n]) * s_itau[d];
      ss = fma(dl, dl, ss);
    }
    z[n] = exp(s_lnnf - 0.5 * ss);
  }
}

// V = R' \ Z for up to VR right-hand sides per CTA (kept in shared memory), blocked by 64 rows.
// grid (ceil(K/VR), S), 256 threads = 8 warps.  VR = 8 unless N is too large for 8 resident columns.
template <int VR>
__global__ void __launch_bounds__(256) var_fwd_kernel(const VarArgs a) {
  extern __shared__ __align__(16) double vsm[];
  const int N = a.N, ld = a.ld, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.y, k0 = blockIdx.x * VR;
  if (a.isfac && !a.isfac[s]) return;
  const int nrhs = (a.K - k0) < VR ? (a.K - k0) : VR;
  double* v = vsm;                 // [VR][N]
  double* Rb = vsm + VR * N;       // [64][65]  Rb[c][r] = R(b0+r, b0+c)
  double* pd = Rb + 64 * 65;       // [64][VR]  partial dots of the current block
  const double* R = a.L + static_cast<size_t>(s) * a.Lstride;
  double* Zs = a.Z + (static_cast<size_t>(s) * a.K + k0) * N;
  if (a.unit_rhs) {
    for (int i = tid; i < VR * N; i += 256) {
      const int cc = i / N, r = i - cc * N;
      v[i] = (cc < nrhs && r == k0 + cc) ? 1.0 : 0.0;
    }
  } else {
    for (int i = tid; i < nrhs * N; i += 256) v[i] = Zs[i];
    for (int i = nrhs * N + tid; i < VR * N; i += 256) v[i] = 0.0;
  }
  __syncthreads();
  const int nb = (N + 63) / 64;
  const int bfirst = a.unit_rhs ? k0 / 64 : 0;   // e_a has no entries above row a: the solution is zero there
  const int jfirst = bfirst * 64;
  for (int b = bfirst; b < nb; ++b) {
    const int b0 = b * 64;
    // (1) pd[i][c] = sum_{j<b0} R(j, b0+i) v_c[j]   — column b0+i of R is contiguous over j
    for (int i = warp; i < 64; i += 8) {
      double acc[VR];
#pragma unroll
      for (int c = 0; c < VR; ++c) acc[c] = 0.0;
      if (b0 + i < N) {
        const double* col = R + static_cast<size_t>(b0 + i) * ld;
        for (int j = jfirst + lane; j < b0; j += 32) {
          const double r = col[j];
#pragma unroll
          for (int c = 0; c < VR; ++c) acc[c] = fma(r, v[c * N + j], acc[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < VR; ++c) {
        double x = acc[c];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if (lane == 0) pd[i * VR + c] = x;
      }
    }
    // diagonal block
    for (int i = tid; i < 64 * 64; i += 256) {
      const int c = i >> 6, r = i & 63;
      Rb[c * 65 + r] = (b0 + c < N && b0 + r < N && r <= c) ? R[static_cast<size_t>(b0 + c) * ld + b0 + r] : (c == r ? 1.0 : 0.0);
    }
    __syncthreads();
    // (2) triangular solve inside the block: warp c handles right-hand side c, lane owns rows lane, lane+32
    if (warp < nrhs) {
      double* vc = v + warp * N;
      double x0 = (b0 + lane < N) ? vc[b0 + lane] - pd[lane * VR + warp] : 0.0;
      double x1 = (b0 + lane + 32 < N) ? vc[b0 + lane + 32] - pd[(lane + 32) * VR + warp] : 0.0;
      for (int p = 0; p < 64; ++p) {
        // v_p = (rhs_p - sum_{j<p} R(j,p) v_j) / R(p,p); the running rhs already holds the subtraction
        double xp = (p < 32 ? x0 : x1) / Rb[p * 65 + p];
        xp = __shfl_sync(0xffffffffu, xp, p & 31);
        if (p < 32) {
          if (lane == p) x0 = xp;
          if (lane > p) x0 = fma(-Rb[lane * 65 + p], xp, x0);      // row lane > p: R(p, lane)
          x1 = fma(-Rb[(lane + 32) * 65 + p], xp, x1);
        } else {
          if (lane == (p & 31)) x1 = xp;
          if (lane + 32 > p) x1 = fma(-Rb[(lane + 32) * 65 + p], xp, x1);
        }
      }
      if (b0 + lane < N) vc[b0 + lane] = x0;
      if (b0 + lane + 32 < N) vc[b0 + lane + 32] = x1;
    }
    __syncthreads();
  }
  for (int i = tid; i < nrhs * N; i += 256) Zs[i] = v[i];
}

// W = R \ V (backward substitution, upper factor) for up to VR right-hand sides per CTA, in place in Z.
// After W: K^-1 z_k = W_k / sn2_eff (gplogjoint.m:276-277).  grid (ceil(K/VR), S), 256 threads.
template <int VR>
__global__ void __launch_bounds__(256) var_bwd_kernel(const VarArgs a) {
  extern __shared__ __align__(16) double vsm[];
  const int N = a.N, ld = a.ld, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.y, k0 = blockIdx.x * VR;
  if (a.isfac && !a.isfac[s]) return;
  const int nrhs = (a.K - k0) < VR ? (a.K - k0) : VR;
  double* v = vsm;             // [VR][N]
  double* Rb = vsm + VR * N;   // [64][65]
  const double* R = a.L + static_cast<size_t>(s) * a.Lstride;
  double* Zs = a.Z + (static_cast<size_t>(s) * a.K + k0) * N;
  for (int i = tid; i < nrhs * N; i += 256) v[i] = Zs[i];
  for (int i = nrhs * N + tid; i < VR * N; i += 256) v[i] = 0.0;
  __syncthreads();
  const int nb = (N + 63) / 64;
  for (int b = nb - 1; b >= 0; --b) {
    const int b0 = b * 64;
    for (int i = tid; i < 64 * 64; i += 256) {
      const int c = i >> 6, r = i & 63;
      Rb[c * 65 + r] = (b0 + c < N && b0 + r < N && r <= c) ? R[static_cast<size_t>(b0 + c) * ld + b0 + r] : (c == r ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp < nrhs) {  // upper back substitution inside the block; lane owns rows lane, lane+32
      double* vc = v + warp * N;
      double x0 = (b0 + lane < N) ? vc[b0 + lane] : 0.0;
      double x1 = (b0 + lane + 32 < N) ? vc[b0 + lane + 32] : 0.0;
      for (int p = 63; p >= 0; --p) {
        double xp = (p >= 32 ? x1 : x0) / Rb[p * 65 + p];
        xp = __shfl_sync(0xffffffffu, xp, p & 31);
        if (p >= 32) {
          if (lane == (p & 31)) x1 = xp;
          if (lane + 32 < p) x1 = fma(-Rb[p * 65 + lane + 32], xp, x1);   // R(lane+32, p)
          x0 = fma(-Rb[p * 65 + lane], xp, x0);
        } else {
          if (lane == p) x0 = xp;
          if (lane < p) x0 = fma(-Rb[p * 65 + lane], xp, x0);
        }
      }
      if (b0 + lane < N) vc[b0 + lane] = x0;
      if (b0 + lane + 32 < N) vc[b0 + lane + 32] = x1;
    }
    __syncthreads();
    // v_i -= sum_{j in block} R(i, b0+j) w_j  for i < b0   (coalesced over i)
    for (int i = tid; i < b0; i += 256) {
      double acc[VR];
#pragma unroll
      for (int c = 0; c < VR; ++c) acc[c] = 0.0;
      const int jn = (N - b0) < 64 ? (N - b0) : 64;
      for (int j = 0; j < jn; ++j) {
        const double r = R[static_cast<size_t>(b0 + j) * ld + i];
#pragma unroll
        for (int c = 0; c < VR; ++c) acc[c] = fma(r, v[c * N + b0 + j], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < VR; ++c) v[c * N + i] -= acc[c];
    }
    __syncthreads();
  }
  for (int i = tid; i < nrhs * N; i += 256) Zs[i] = v[i];
}

// Low-noise posterior handed over by gp_attach: post.L = -inv(K + diag) (gplite_core.m:96-99), so
// K^-1 z_k = -L z_k (gplogjoint.m:279, 325).  W_k(i) = -sum_j L(j,i) z_k(j): column i of the symmetric L is
// contiguous over j.  grid (ceil(K/VR), S); samples with a Cholesky factor are skipped.
template <int VR>
__global__ void __launch_bounds__(256) var_symv_kernel(const VarArgs a) {
  extern __shared__ __align__(16) double vsm[];
  const int N = a.N, ld = a.ld, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.y, k0 = blockIdx.x * VR;
  if (!a.isfac || a.isfac[s]) return;
  const int nrhs = (a.K - k0) < VR ? (a.K - k0) : VR;
  double* v = vsm;  // [VR][N]
  const double* Lm = a.L + static_cast<size_t>(s) * a.Lstride;
  const double* Zs = a.Z + (static_cast<size_t>(s) * a.K + k0) * N;
  double* Ws = a.W + (static_cast<size_t>(s) * a.K + k0) * N;
  for (int i = tid; i < nrhs * N; i += 256) v[i] = Zs[i];
  for (int i = nrhs * N + tid; i < VR * N; i += 256) v[i] = 0.0;
  __syncthreads();
  for (int i = warp; i < N; i += 8) {
    double acc[VR];
#pragma unroll
    for (int c = 0; c < VR; ++c) acc[c] = 0.0;
    const double* col = Lm + static_cast<size_t>(i) * ld;
    for (int j = lane; j < N; j += 32) {
      const double r = col[j];
#pragma unroll
      for (int c = 0; c < VR; ++c) acc[c] = fma(r, v[c * N + j], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < VR; ++c) {
      double x = acc[c];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
      if (lane == 0 && c < nrhs) Ws[static_cast<size_t>(c) * N + i] = -x;
    }
  }
}

// G[s][j][k] = V_j . V_k  (k <= j, mirrored); grid (K, S): CTA (j, s) loops over k <= j.
// Samples without a factor: G[s][j][k] = z_j . W_k = z_j K^-1 z_k directly.
__global__ void __launch_bounds__(256) var_gram_kernel(const VarArgs a) {
  extern __shared__ __align__(16) double gsm[];  // V_j [N]
  __shared__ double part[8];
  const int j = blockIdx.x, s = blockIdx.y, N = a.N, K = a.K, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Vs = a.Z + static_cast<size_t>(s) * K * N;
  const double* Bs = (a.isfac && !a.isfac[s]) ? a.W + static_cast<size_t>(s) * K * N : Vs;
  for (int i = tid; i < N; i += 256) gsm[i] = Vs[static_cast<size_t>(j) * N + i];
  __syncthreads();
  const int kbeg = a.full ? 0 : j;
  for (int k = kbeg; k <= j; ++k) {
    double acc = 0.0;
    const double* vk = Bs + static_cast<size_t>(k) * N;
    for (int i = tid; i < N; i += 256) acc = fma(gsm[i], vk[i], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += part[w];
      a.G[(static_cast<size_t>(s) * K + j) * K + k] = t;
      a.G[(static_cast<size_t>(s) * K + k) * K + j] = t;
    }
    __syncthreads();
  }
}

// J_sjk and varF(s)  (gplogjoint.m:284-287, 309-336, 350); grid (S), 256 threads
__global__ void __launch_bounds__(256) var_final_kernel(const VarArgs a) {
  __shared__ double part[256];
  const int s = blockIdx.x, K = a.K, D = a.D, tid = threadIdx.x;
  const double EPS = 2.220446049250313e-16;
  const double sn2eff = a.gp.sn2eff[s];
  double acc = 0.0;
  for (int i = tid; i < K * K; i += 256) {
    const int j = i / K, k = i - j * K;
    double Jv = 0.0;
    if (j <= k && (a.full || j == k)) {
      // tau_jk = sqrt((sigma_j^2 + sigma_k^2)*lambda^2 + ell^2 + 2*delta^2)   (:314 ; j == k gives tau_kk :274)
      double slt = 0.0, dd = 0.0;
      for (int d = 0; d < D; ++d) {
        const double lam = a.vp.lambda[d], ell = a.gp.ell[s * D + d], dl = a.vp.delta[d];
        const double t2 = (a.vp.sigma[j] * a.vp.sigma[j] + a.vp.sigma[k] * a.vp.sigma[k]) * lam * lam + ell * ell + 2.0 * dl * dl;
        slt += 0.5 * log(t2);
        const double dm = a.vp.mu[j * D + d] - a.vp.mu[k * D + d];
        dd += dm * dm / t2;
      }
      Jv = exp(a.gp.lnc[s] - slt - 0.5 * dd) - a.G[(static_cast<size_t>(s) * K + j) * K + k] / sn2eff;
      if (j == k)
        acc += a.vp.w[k] * a.vp.w[k] * fmax(EPS, Jv);     // :283 / :330
      else
        acc += 2.0 * a.vp.w[j] * a.vp.w[k] * Jv;           // :333
    }
    a.J[(static_cast<size_t>(s) * K + j) * K + k] = Jv;
  }
  __syncthreads();
  // mirror the upper part (J_sjk(s,k,j) = J_jk, :334)
  for (int i = tid; i < K * K; i += 256) {
    const int j = i / K, k = i - j * K;
    if (j > k) a.J[(static_cast<size_t>(s) * K + j) * K + k] = a.full ? a.J[(static_cast<size_t>(s) * K + k) * K + j] : 0.0;
  }
  part[tid] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) part[tid] += part[tid + off];
    __syncthreads();
  }
  if (tid == 0) a.varFs[s] = fmax(part[0], EPS);  // varF = max(varF,eps) (:350)
}

enum { VK_FWD = 0, VK_BWD = 1, VK_SYMV = 2 };

template <int VR>
static int launch_vk(int which, const VarArgs& a, int ncols, int S, size_t smem, cudaStream_t st) {
  dim3 grid((ncols + VR - 1) / VR, S);
  switch (which) {
    case VK_FWD:
      VB_CUDA(cudaFuncSetAttribute(var_fwd_kernel<VR>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      var_fwd_kernel<VR><<<grid, 256, smem, st>>>(a);
      break;
    case VK_BWD:
      VB_CUDA(cudaFuncSetAttribute(var_bwd_kernel<VR>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      var_bwd_kernel<VR><<<grid, 256, smem, st>>>(a);
      break;
    default:
      VB_CUDA(cudaFuncSetAttribute(var_symv_kernel<VR>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      var_symv_kernel<VR><<<grid, 256, smem, st>>>(a);
      break;
  }
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

// one of the column-block kernels with the widest block that fits; `extra` = doubles of shared memory besides the columns
static int launch_var_kernel(vbmc_b200_ctx* c, int which, const VarArgs& a, int ncols, int S, size_t extra_fixed, size_t extra_per_col,
                             cudaStream_t st, const char* what) {
  int vr = 0;
  for (int t = 8; t >= 1; t >>= 1)
    if (sizeof(double) * (static_cast<size_t>(t) * a.N + extra_fixed + extra_per_col * t) <= c->smem_optin) { vr = t; break; }
  if (!vr)
    VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:%s: N=%d does not fit one solution column in shared memory (%zu B)", what, a.N, c->smem_optin);
  const size_t smem = sizeof(double) * (static_cast<size_t>(vr) * a.N + extra_fixed + extra_per_col * vr);
  switch (vr) {
    case 8: return launch_vk<8>(which, a, ncols, S, smem, st);
    case 4: return launch_vk<4>(which, a, ncols, S, smem, st);
    case 2: return launch_vk<2>(which, a, ncols, S, smem, st);
    default: return launch_vk<1>(which, a, ncols, S, smem, st);
  }
}

// Runs the variance pipeline for all S samples; results to host: varFs[S], J[S][K][K] (optional).
int run_variance(vbmc_b200_ctx* c, int compute_var, std::vector<double>* varFs, std::vector<double>* J, std::vector<double>* vgrad) {
  if (!c->gpHasL)
    VB_FAIL(VBMC_B200_ESTATE, "gplogjoint variance needs the factors gp.post(s).L on the device (gp_attach with L, or gp_post)");
  const int N = c->gp.N, K = c->K, S = c->gp.S;
  bool any_inv = false;
  for (int s = 0; s < S; ++s) any_inv = any_inv || !c->gpLfactor[s];
  VarArgs a;
  a.N = N; a.D = c->D; a.K = K; a.S = S; a.ld = c->gpLd;
  a.Lstride = static_cast<size_t>(c->gpLd) * c->gpLd;
  a.L = c->gpL.d();
  a.gp = c->gp; a.vp = c->vp;
  a.full = compute_var == 1 ? 1 : 0;
  a.unit_rhs = 0;
  const size_t nz = static_cast<size_t>(S) * K * N, ng = static_cast<size_t>(S) * K * K;
  const size_t nvg = vgrad ? static_cast<size_t>(S) * K * (2 + 2 * c->D) : 0;
  const size_t nw = any_inv ? nz : 0, nflag = any_inv ? (static_cast<size_t>(S) + 1) / 2 : 0;  // ints packed behind the doubles
  VB_TRY(c->varWork.reserve(sizeof(double) * (nz + 2 * ng + S + nvg + nw + nflag)));
  a.Z = c->varWork.d(); a.G = a.Z + nz; a.J = a.G + ng; a.varFs = a.J + ng;
  double* vg = a.varFs + S;
  a.W = any_inv ? vg + nvg : nullptr;
  a.isfac = nullptr;
  cudaStream_t st = c->stream;
  if (any_inv) {
    int* flags = reinterpret_cast<int*>(a.W + nw);
    VB_CUDA(cudaMemcpyAsync(flags, c->gpLfactor.data(), sizeof(int) * S, cudaMemcpyHostToDevice, st));
    a.isfac = flags;
  }
  {
    dim3 grid(K, S);
    KernelScope ks(c, "var_z", st);
    glj_z_kernel<<<grid, 256, 0, st>>>(a);
  }
  {
    KernelScope ks(c, "var_trsm", st);
    VB_TRY(launch_var_kernel(c, VK_FWD, a, K, S, 64 * 65, 64, st, "variance"));
  }
  if (any_inv) {
    KernelScope ks(c, "var_symv", st);
    VB_TRY(launch_var_kernel(c, VK_SYMV, a, K, S, 0, 0, st, "variance"));
  }
  {
    const size_t smem = sizeof(double) * N;
    if (smem > 48 * 1024) VB_CUDA(cudaFuncSetAttribute(var_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    dim3 grid(K, S);
    KernelScope ks(c, "var_gram", st);
    var_gram_kernel<<<grid, 256, smem, st>>>(a);
  }
  {
    KernelScope ks(c, "var_final", st);
    var_final_kernel<<<S, 256, 0, st>>>(a);
  }
  VB_CUDA(cudaGetLastError());
  varFs->assign(S, 0.0);
  VB_CUDA(cudaMemcpyAsync(varFs->data(), a.varFs, sizeof(double) * S, cudaMemcpyDeviceToHost, st));
  if (J) {
    J->assign(ng, 0.0);
    VB_CUDA(cudaMemcpyAsync(J->data(), a.J, sizeof(double) * ng, cudaMemcpyDeviceToHost, st));
  }
  if (vgrad) {
    // K^-1 z_k = R \ V_k / sn2_eff (or W_k for an inverse-form posterior), then the derivative contractions
    // dz_d(.) * K^-1 z_k  (gplogjoint.m:289-299)
    {
      KernelScope ks(c, "var_trsm", st);
      VB_TRY(launch_var_kernel(c, VK_BWD, a, K, S, 64 * 65, 0, st, "variance"));
    }
    for (int s = 0; any_inv && s < S; ++s)
      if (!c->gpLfactor[s])
        VB_CUDA(cudaMemcpyAsync(a.Z + static_cast<size_t>(s) * K * N, a.W + static_cast<size_t>(s) * K * N, sizeof(double) * K * N,
                                cudaMemcpyDeviceToDevice, st));
    VB_TRY(launch_gplogjoint_weighted(c, a.Z, vg, st));
    vgrad->assign(nvg, 0.0);
    VB_CUDA(cudaMemcpyAsync(vgrad->data(), vg, sizeof(double) * nvg, cudaMemcpyDeviceToHost, st));
  }
  VB_CUDA(cudaStreamSynchronize(st));
  return VBMC_B200_OK;
}


// V = R' \ Z in place for `ncols` right-hand sides per sample (Z: [S][ncols][N]) against the resident factors; samples whose
// resident matrix is -inv(K+Sigma) instead of a factor get W = K^-1 Z (gplite_pred.m:96-102).
int run_rhs_solve(vbmc_b200_ctx* c, int ncols, double* Z, double* W, const int* isfac_dev, cudaStream_t st) {
  const int S = c->gp.S;
  bool any_inv = false;
  for (int s = 0; s < S; ++s) any_inv = any_inv || !c->gpLfactor[s];
  VarArgs a;
  memset(&a, 0, sizeof(a));
  a.N = c->gp.N; a.D = c->gp.D; a.K = ncols; a.S = S; a.ld = c->gpLd;
  a.Lstride = static_cast<size_t>(c->gpLd) * c->gpLd;
  a.L = c->gpL.d();
  a.gp = c->gp; a.vp = c->vp;
  a.Z = Z; a.W = W; a.isfac = isfac_dev; a.unit_rhs = 0;
  static const bool blocked_off = getenv("VBMC_B200_TRSM_BLOCKED") && atoi(getenv("VBMC_B200_TRSM_BLOCKED")) == 0;
  int rc = VBMC_B200_OK;
  if (!blocked_off && ncols == 1 && run_trsv1(c, Z, isfac_dev, st, &rc, false)) {
    VB_TRY(rc);   // one column per sample: the whole sweep in one launch (trsm.cu)
  } else if (!blocked_off && run_trsm_blocked(c, ncols, Z, isfac_dev, st, &rc)) {
    VB_TRY(rc);   // blocked DMMA forward substitution (trsm.cu)
  } else {
    KernelScope ks(c, "pred_trsm", st);
    VB_TRY(launch_var_kernel(c, VK_FWD, a, ncols, S, 64 * 65, 64, st, "gplite_pred"));
  }
  if (any_inv) {
    KernelScope ks(c, "pred_symv", st);
    VB_TRY(launch_var_kernel(c, VK_SYMV, a, ncols, S, 0, 0, st, "gplite_pred"));
  }
  return VBMC_B200_OK;
}

// W = R \ Z in place (backward substitution with the resident upper factors), `ncols` right-hand sides per sample
int run_rhs_backsolve(vbmc_b200_ctx* c, int ncols, double* Z, cudaStream_t st, const int* isfac_dev) {
  VarArgs a;
  memset(&a, 0, sizeof(a));
  a.N = c->gp.N; a.D = c->gp.D; a.K = ncols; a.S = c->gp.S; a.ld = c->gpLd;
  a.Lstride = static_cast<size_t>(c->gpLd) * c->gpLd;
  a.L = c->gpL.d();
  a.gp = c->gp; a.vp = c->vp;
  a.Z = Z; a.isfac = isfac_dev;
  static const bool blocked_off = getenv("VBMC_B200_TRSM_BLOCKED") && atoi(getenv("VBMC_B200_TRSM_BLOCKED")) == 0;
  int rc = VBMC_B200_OK;
  if (!blocked_off && ncols == 1 && run_trsv1(c, Z, isfac_dev, st, &rc, true)) return rc;
  if (!blocked_off && run_trsm_blocked(c, ncols, Z, isfac_dev, st, &rc, true)) return rc;
  KernelScope ks(c, "pred_trsm", st);
  return launch_var_kernel(c, VK_BWD, a, ncols, c->gp.S, 64 * 65, 0, st, "gplite_post");
}

// X = R^-T (lower triangular) of sample `s`, written column-major with leading dimension N into `out`.
int run_factor_inverse(vbmc_b200_ctx* c, int N, int ld, const double* R, double* out) {
  VarArgs a;
  memset(&a, 0, sizeof(a));
  a.N = N; a.D = 1; a.K = N; a.S = 1; a.ld = ld;
  a.Lstride = 0;
  a.L = R;
  a.Z = out;
  a.unit_rhs = 1;
  KernelScope ks(c, "trtri", c->stream);
  VB_TRY(launch_var_kernel(c, VK_FWD, a, N, 1, 64 * 65, 64, c->stream, "inverse"));
  return VBMC_B200_OK;
}

}  // namespace vb
