// fminadam on the device (reference: utils/fminadam.m:20-102).  The reference calls negelcbo_vbmc once per
// iteration through a function handle and updates x on the host; here the iterate, the Adam moments, the
// iterate/value history (xtab, ftab) and the termination statistics live in HBM, so that a whole stochastic
// optimisation is one library call with no per-iteration host<->device traffic.
//   adam_step_kernel  : m, v, bias correction, step-size schedule, update, clamp, history   (:51-63)
//   adam_check_kernel : slope of the last 20 values (polyfit degree 1) + random-walk distance (:65-83)
//   adam_final_kernel : x = mean of the last 20 iterates, f = mean of the last 20 values     (:95-96)
// O(nvars) work in one CTA each; they only exist to keep the loop on the device.
#include "common.cuh"

namespace vb {

__global__ void __launch_bounds__(256) adam_step_kernel(const AdamArgs a) {
  const int tid = threadIdx.x, nt = blockDim.x, n = a.n;
  const int iter = *a.it + 1;  // 1-based iteration number of this update
  const double beta1 = 0.9, beta2 = 0.999;
  const double c1 = 1.0 - pow(beta1, static_cast<double>(iter));  // 1-beta1^iter (:53)
  const double c2 = 1.0 - pow(beta2, static_cast<double>(iter));
  const double stepsize = a.step_min + (a.step_max - a.step_min) * exp(-static_cast<double>(iter) / a.decay);  // :56-57
  const double fudge = 1.4901161193847656e-08;  // sqrt(eps) (:21)
  double* xcol = a.xtab + static_cast<size_t>(iter - 1) * n;
  for (int i = tid; i < n; i += nt) {
    const double g = a.grad[i];
    const double m = beta1 * a.m[i] + (1.0 - beta1) * g;        // :51
    const double v = beta2 * a.v[i] + (1.0 - beta2) * (g * g);  // :52
    a.m[i] = m;
    a.v[i] = v;
    const double mhat = m / c1, vhat = v / c2;
    double x = a.x[i] - stepsize * mhat / (sqrt(vhat) + fudge);  // :59
    x = fmin(fmax(x, a.lb[i]), a.ub[i]);                         // :60
    a.x[i] = x;
    xcol[i] = x;                                                 // :63
  }
  if (tid == 0) a.ftab[iter - 1] = *a.fval;                      // :48
  __syncthreads();
  if (tid == 0) {
    *a.it = iter;
    if (a.dyn) a.dyn[1] += 1;  // next iteration draws from the next Philox stream (fresh randn per call, entmc_vbmc.m:53)
  }
}

// blockDim.x == 256; stats: [0] stop flag, [1] dx, [2] slope, [3] slope_err, [4] slope_err_max
__global__ void __launch_bounds__(256) adam_check_kernel(const AdamArgs a, int iter, double TolFun) {
  __shared__ double part[256];
  const int tid = threadIdx.x, nt = blockDim.x, n = a.n, B = 20;
  // random-walk distance between the means of the last two mini-batches (:77)
  double acc = 0.0;
  for (int i = tid; i < n; i += nt) {
    double s1 = 0.0, s0 = 0.0;
    for (int t = 0; t < B; ++t) {
      s1 += a.xtab[static_cast<size_t>(iter - B + t) * n + i];
      s0 += a.xtab[static_cast<size_t>(iter - 2 * B + t) * n + i];
    }
    const double d = s1 / B - s0 / B;
    acc += d * d / B;
  }
  part[tid] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) part[tid] += part[tid + off];
    __syncthreads();
  }
  if (tid == 0) {
    const double dx = sqrt(part[0]);
    // polyfit(xxp, y, 1) with xxp = -(B-1)/2 .. (B-1)/2 (sum xxp = 0): slope = sum(xxp.*y)/sum(xxp.^2),
    // intercept = mean(y); covariance A = inv(V'V)*normr^2/df, A(1,1) = normr^2/df/sum(xxp.^2)  (:68-73)
    const double* y = a.ftab + (iter - B);
    double sxy = 0.0, sy = 0.0, sxx = 0.0;
    for (int t = 0; t < B; ++t) {
      const double xx = t - 0.5 * (B - 1);
      sxy += xx * y[t];
      sy += y[t];
      sxx += xx * xx;
    }
    const double slope = sxy / sxx, icpt = sy / B;
    double r2 = 0.0;
    for (int t = 0; t < B; ++t) {
      const double xx = t - 0.5 * (B - 1);
      const double r = y[t] - (slope * xx + icpt);
      r2 += r * r;
    }
    const double A11 = r2 / (B - 2) / sxx;
    const double TolX = 0.001, TolX_max = 0.1, TolFun_max = TolFun * 100.0;  // :25-27
    const double slope_err = sqrt(A11 + TolFun * TolFun);
    const double slope_err_max = sqrt(A11 + TolFun_max * TolFun_max);
    const bool stop = (dx < TolX && fabs(slope) < slope_err_max) || (fabs(slope) < slope_err && dx < TolX_max);  // :80
    a.stats[0] = stop ? 1.0 : 0.0;
    a.stats[1] = dx;
    a.stats[2] = slope;
    a.stats[3] = slope_err;
    a.stats[4] = slope_err_max;
  }
}

// xout[i] = mean(xtab(i, iter-19:iter)), stats[5] = mean(ftab(iter-19:iter))   (:95-96)
__global__ void __launch_bounds__(256) adam_final_kernel(const AdamArgs a, int iter) {
  const int tid = threadIdx.x, nt = blockDim.x, n = a.n, B = 20;
  for (int i = tid; i < n; i += nt) {
    double s = 0.0;
    for (int t = 0; t < B; ++t) s += a.xtab[static_cast<size_t>(iter - B + t) * n + i];
    a.xout[i] = s / B;
  }
  if (tid == 0) {
    double s = 0.0;
    for (int t = 0; t < B; ++t) s += a.ftab[iter - B + t];
    a.stats[5] = s / B;
  }
}

// F += corr[0], dF += corr[1..n]: the variance penalty of negelcbo_vbmc.m:119-130 folded into the step's outputs before the
// Adam update reads them (beta ~= 0 only)
__global__ void adam_penalty_kernel(int n, const double* corr, double* fval, double* grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *fval += corr[0];
  if (i < n) grad[i] += corr[1 + i];
}
int launch_adam_penalty(vbmc_b200_ctx* c, const AdamArgs& a, const double* corr, cudaStream_t st) {
  KernelScope ks(c, "adam", st);
  adam_penalty_kernel<<<(a.n + 255) / 256, 256, 0, st>>>(a.n, corr, const_cast<double*>(a.fval), const_cast<double*>(a.grad));
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

int launch_adam_step(vbmc_b200_ctx* c, const AdamArgs& a, cudaStream_t st) {
  KernelScope ks(c, "adam", st);
  adam_step_kernel<<<1, 256, 0, st>>>(a);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}
int launch_adam_check(vbmc_b200_ctx* c, const AdamArgs& a, int iter, double TolFun, cudaStream_t st) {
  KernelScope ks(c, "adam_check", st);
  adam_check_kernel<<<1, 256, 0, st>>>(a, iter, TolFun);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}
int launch_adam_final(vbmc_b200_ctx* c, const AdamArgs& a, int iter, cudaStream_t st) {
  KernelScope ks(c, "adam_final", st);
  adam_final_kernel<<<1, 256, 0, st>>>(a, iter);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

}  // namespace vb
