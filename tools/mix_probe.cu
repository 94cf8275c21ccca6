// Micro-probe: can the FP64 FMA pipe (DFMA) and the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64) run concurrently on
// sm_100?  Three kernels per warp count: DFMA only, DMMA only, both interleaved in one instruction stream.
// If t(both) ~ max(t_dfma, t_dmma) the pipes are independent; if ~ t_dfma + t_dmma they share issue/datapath.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NF, int NM>
__global__ void k(double* out, int iters, double a, double b) {
  double x[NF > 0 ? NF : 1], c0[NM > 0 ? NM : 1], c1[NM > 0 ? NM : 1];
#pragma unroll
  for (int i = 0; i < NF; ++i) x[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < NM; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
    // 4 rounds: NF DFMA (independent chains) + NM/4... keep it simple: per iteration 4*NF DFMA and NM DMMA, interleaved
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < NF; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
      for (int i = r; i < NM; i += 4) dmma(c0[i], c1[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < NM; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NF, int NM>
float run(int warps, int sms, int iters) {
  double* out;
  cudaMalloc(&out, sizeof(double) * 1024 * sms);
  k<NF, NM><<<sms, warps * 32>>>(out, iters, 0.999999, 1e-9);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<NF, NM><<<sms, warps * 32>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaFree(out);
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, iters = 20000;
  const int ws[] = {4, 8, 16, 32};
  for (int w : ws) {
    // per iteration and warp: 32 DFMA (8 chains x 4 rounds) and 4 DMMA
    const float tf = run<8, 0>(w, sms, iters), tm = run<0, 4>(w, sms, iters), tb = run<8, 4>(w, sms, iters);
    const float tf2 = run<8, 0>(w, sms, iters), tm8 = run<0, 8>(w, sms, iters), tb8 = run<8, 8>(w, sms, iters);
    printf("warps/SM=%2d | 32 DFMA: %.3f ms  4 DMMA: %.3f ms  both: %.3f ms (sum %.3f, max %.3f) | 32 DFMA: %.3f  8 DMMA: %.3f  both: %.3f (sum %.3f, max %.3f)\n",
           w, tf, tm, tb, tf + tm, tf > tm ? tf : tm, tf2, tm8, tb8, tf2 + tm8, tf2 > tm8 ? tf2 : tm8);
  }
  return 0;
}
