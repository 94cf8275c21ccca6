// Micro-probe: DFMA dependent-issue latency and pipe throughput vs (#warps per SM, #independent chains).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
  double x[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) x[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void run(int warps_per_sm, int sms) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, sizeof(double) * 2048 * sms); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  k<CH><<<sms, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-9, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<CH><<<sms, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-9, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per_instr = (double)h / ((double)iters * CH);            // cycles per DFMA issued by one warp
  double dfma_per_clk_sm = (double)warps_per_sm * 32 * iters * CH / (double)h;
  printf("warps/SM=%2d chains=%d : %.2f cyc per DFMA per warp, %.1f DFMA lanes/clk/SM, %.2f TFLOP/s\n", warps_per_sm, CH,
         per_instr, dfma_per_clk_sm, 2.0 * warps_per_sm * 32.0 * iters * CH * sms / (ms * 1e-3) / 1e12);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  run<1>(1, sms); run<2>(1, sms); run<4>(1, sms); run<8>(1, sms); run<16>(1, sms);
  run<1>(4, sms); run<2>(4, sms); run<4>(4, sms); run<8>(4, sms);
  run<1>(8, sms); run<2>(8, sms); run<4>(8, sms); run<8>(8, sms);
  run<1>(16, sms); run<2>(16, sms); run<4>(16, sms);
  run<1>(32, sms); run<2>(32, sms); run<8>(32, sms);
  return 0;
}
