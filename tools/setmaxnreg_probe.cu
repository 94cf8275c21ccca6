// Round-2 probe: can the entropy sweep's CTA carry a third, register-poor warp group that works in
// the sweep's idle issue slots?  The sweep runs 8 warps at 238 registers with 227 KB of dynamic shared memory, so no other
// kernel can co-reside with it; a 12-warp CTA compiled for 168 registers whose two "consumer" groups raise their budget with
// setmaxnreg.inc (232) while a "producer" group lowers its own (setmaxnreg.dec, 40) fits the 64 K register file exactly.
// Measures: (1) consumer-only time (producer idle), (2) producer-only time, (3) both together — if (3) ~ max(1,2) the
// producer's work (draw generation: integer Philox rounds + a few FP64 operations) hides in the consumer's issue gaps.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/setmaxnreg_probe tools/setmaxnreg_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint4 philox_round10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// mode bit 0: consumers work, bit 1: producers work
__global__ void __launch_bounds__(384, 1) probe_kernel(double* out, int iters, int mode, double a, double b) {
  extern __shared__ double smem[];
  const int wg = threadIdx.x >> 7;  // warp group 0,1: consumers; 2: producer
  if (wg < 2) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    if (mode & 1) {
      // a sweep-like instruction stream: 2 warps per scheduler, groups of dependent DFMAs with limited ILP and a
      // shared-memory broadcast load per 8 FMAs
      double x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
      for (int it = 0; it < iters; ++it) {
        const double t = smem[it & 255];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], b, x[(i + 1) & 7]);
      }
      double s = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += x[i];
      out[blockIdx.x * 384 + threadIdx.x] = s;
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (mode & 2) {
      // generator-like stream: one Philox block + a ziggurat-like fast path (2 FP64 ops) per iteration
      double acc = 0;
      uint4 c = make_uint4(threadIdx.x, blockIdx.x, 0, 0);
      for (int it = 0; it < iters / 4; ++it) {
        c.z = it;
        const uint4 r = philox_round10(c, make_uint2(1, 2));
        const double u = __longlong_as_double(0x4330000000000000LL | ((static_cast<long long>(r.y) << 20) ^ r.x)) - 4503599627370496.0;
        acc = fma(u, 1e-16, acc) + static_cast<double>(r.z & 1023u);
      }
      out[blockIdx.x * 384 + threadIdx.x] = acc;
    }
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = 220 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 384);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 200000;
  for (int mode = 1; mode <= 3; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      probe_kernel<<<sms, 384, smem>>>(out, iters, mode, 0.999999, 1e-9);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    printf("mode %d (%s): %.3f ms  [%s]\n", mode, mode == 1 ? "consumers only" : mode == 2 ? "producer only" : "both", best,
           cudaGetErrorString(err));
  }
  return 0;
}
