// Micro-probe: DMMA (mma.sync.m8n8k4.f64) throughput vs warps/SM and independent accumulators.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
  double c0[CH], c1[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma(c0[i], c1[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void run(int warps, int sms) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, sizeof(double) * 2048 * sms); cudaMalloc(&cyc, 8);
  const int iters = 10000;
  k<CH><<<sms, warps * 32>>>(out, iters, 1e-3, 1e-3, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<CH><<<sms, warps * 32>>>(out, iters, 1e-3, 1e-3, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double fma_per_clk_sm = 256.0 * warps * iters * CH / (double)h;
  printf("DMMA warps/SM=%2d acc=%d : %.1f cyc per DMMA per warp, %.1f FMA/clk/SM, %.2f TFLOP/s\n", warps, CH,
         (double)h / ((double)iters * CH), fma_per_clk_sm, 2.0 * 256.0 * warps * iters * CH * sms / (ms * 1e-3) / 1e12);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  run<1>(1, sms); run<2>(1, sms); run<4>(1, sms); run<8>(1, sms);
  run<1>(4, sms); run<2>(4, sms); run<4>(4, sms); run<8>(4, sms);
  run<2>(8, sms); run<4>(8, sms); run<8>(8, sms);
  run<4>(16, sms); run<8>(16, sms); run<8>(32, sms);
  return 0;
}
