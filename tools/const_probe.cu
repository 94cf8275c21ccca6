// Probe: DFMA throughput when one operand is a warp-uniform, loop-indexed __constant__ table entry
// (ULDC/constant-bank path) vs the same table read with broadcast LDS from shared memory.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double ctab[64 * 10];
template <int MODE>
__global__ void k(double* out, int iters, int K, long long* cyc) {
  __shared__ double stab[64 * 10];
  for (int i = threadIdx.x; i < 640; i += blockDim.x) stab[i] = ctab[i];
  __syncthreads();
  double a[10], b[10], e[10];
#pragma unroll
  for (int d = 0; d < 10; ++d) { a[d] = 0; b[d] = 0; e[d] = threadIdx.x * 1e-3 + d; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int kk = 0; kk < K; ++kk) {
      double t0a = 0, t1a = 0;
#pragma unroll
      for (int d = 0; d < 10; d += 2) {
        const double u0 = MODE ? stab[kk * 10 + d] : ctab[kk * 10 + d];
        const double u1 = MODE ? stab[kk * 10 + d + 1] : ctab[kk * 10 + d + 1];
        t0a = fma(e[d], u0, t0a); t1a = fma(e[d + 1], u1, t1a);
      }
      const double tp = t0a + t1a, tm = t0a - t1a;
#pragma unroll
      for (int d = 0; d < 10; ++d) {
        const double u = MODE ? stab[kk * 10 + d] : ctab[kk * 10 + d];
        a[d] = fma(tp, u, a[d]); b[d] = fma(tm, u, b[d]);
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int d = 0; d < 10; ++d) s += a[d] + b[d];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(int warps, int sms) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, sizeof(double) * 2048 * sms); cudaMalloc(&cyc, 8);
  const int iters = 200, K = 50;
  k<MODE><<<sms, warps * 32>>>(out, iters, K, cyc);
  k<MODE><<<sms, warps * 32>>>(out, iters, K, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double dfma = 32.0 * iters * K;  // 10 + 2 + 20 per (it,kk) per thread
  printf("%s warps/SM=%2d : %.1f DFMA lanes/clk/SM\n", MODE ? "smem " : "const", warps, warps * 32.0 * dfma / (double)h);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  double h[640]; for (int i = 0; i < 640; ++i) h[i] = 1e-3 * i;
  cudaMemcpyToSymbol(ctab, h, sizeof(h));
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  for (int w : {4, 8, 16}) { run<0>(w, sms); run<1>(w, sms); }
  return 0;
}
