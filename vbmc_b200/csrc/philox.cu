// Device-side source of the entropy draws: counter-based Philox4x32-10 (Salmon et al., SC'11)
// + Box-Muller.  Replaces MATLAB's global randn stream (ent/entmc_vbmc.m:53), which cannot be
// reproduced outside MATLAB; draws depend only on (seed, stream, element index), never on the
// number of GPUs or on the sharding, so any rank can regenerate any slice.
#include <stdlib.h>

#include "common.cuh"

namespace vb {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  // 53 random bits -> (0,1), never 0 or 1
  const double x = static_cast<double>(hi >> 5) * 67108864.0 + static_cast<double>(lo >> 6);
  return (x + 0.5) * (1.0 / 9007199254740992.0);
}

struct PhiloxArgs {
  int D, K, half, pair_begin, pair_end;
  uint64_t seed, stream;
  uint64_t stream_add;  // added to the stream id (1 when generating the NEXT step's draws ahead of time)
  const uint64_t* dyn;  // optional device pointer to {seed, stream} (graph replay: values change, the node does not)
  double* eps;  // [K][half][D]  (floats in FP32 mode)
};

// one thread per Philox counter = two consecutive elements (2c, 2c+1) of the flat eps array
__global__ void philox_normal_kernel(const PhiloxArgs a) {
  const int j = blockIdx.y;
  const long long e_begin = (static_cast<long long>(j) * a.half + a.pair_begin) * a.D;
  const long long e_end = (static_cast<long long>(j) * a.half + a.pair_end) * a.D;
  const long long c_begin = e_begin >> 1, c_end = (e_end + 1) >> 1;
  const uint64_t seed = a.dyn ? a.dyn[0] : a.seed, strm = (a.dyn ? a.dyn[1] : a.stream) + a.stream_add;
  for (long long c = c_begin + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; c < c_end;
       c += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 ctr = make_uint4(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32),
                                 static_cast<uint32_t>(strm), static_cast<uint32_t>(strm >> 32));
    const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const uint4 r = philox4x32_10(ctr, key);
    const double u1 = u53(r.x, r.y), u2 = u53(r.z, r.w);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    const long long e0 = 2 * c, e1 = 2 * c + 1;
    if (e0 >= e_begin && e0 < e_end) a.eps[e0] = rad * cs;
    if (e1 >= e_begin && e1 < e_end) a.eps[e1] = rad * sn;
  }
}

// FP32 mode: one Philox counter = FOUR consecutive elements (4c .. 4c+3); each 32-bit word gives one uniform
// u = (x + 0.5) * 2^-32 and two words one Box-Muller pair in single precision (|z| <= 6.7).  Stored as float.
__global__ void philox_normal_f32_kernel(const PhiloxArgs a) {
  const int j = blockIdx.y;
  float* out = reinterpret_cast<float*>(a.eps);
  const long long e_begin = (static_cast<long long>(j) * a.half + a.pair_begin) * a.D;
  const long long e_end = (static_cast<long long>(j) * a.half + a.pair_end) * a.D;
  const long long c_begin = e_begin >> 2, c_end = (e_end + 3) >> 2;
  const uint64_t seed = a.dyn ? a.dyn[0] : a.seed, strm = (a.dyn ? a.dyn[1] : a.stream) + a.stream_add;
  for (long long c = c_begin + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; c < c_end;
       c += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 ctr = make_uint4(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32),
                                 static_cast<uint32_t>(strm) ^ 0x32323232u, static_cast<uint32_t>(strm >> 32));
    const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const uint4 r = philox4x32_10(ctr, key);
    const float S = 2.3283064365386963e-10f, H = 1.1641532182693481e-10f;  // 2^-32, 2^-33
    const float u1 = fmaf(static_cast<float>(r.x), S, H), u2 = fmaf(static_cast<float>(r.y), S, H);
    const float u3 = fmaf(static_cast<float>(r.z), S, H), u4 = fmaf(static_cast<float>(r.w), S, H);
    const float ra = sqrtf(-2.0f * logf(u1)), rb = sqrtf(-2.0f * logf(u3));
    float sa, ca, sb, cb;
    sincospif(2.0f * u2, &sa, &ca);
    sincospif(2.0f * u4, &sb, &cb);
    const float z[4] = {ra * ca, ra * sa, rb * cb, rb * sb};
    const long long e0 = 4 * c;
    if (e0 >= e_begin && e0 + 3 < e_end) {
      *reinterpret_cast<float4*>(out + e0) = make_float4(z[0], z[1], z[2], z[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (e0 + i >= e_begin && e0 + i < e_end) out[e0 + i] = z[i];
    }
  }
}

// raw generator output for the known-answer test (tests/test_philox.py)
__global__ void philox_raw_kernel(uint4 ctr, uint2 key, uint32_t* out) {
  const uint4 r = philox4x32_10(ctr, key);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

int launch_philox(vbmc_b200_ctx* c, int D, int K, int Ns, uint64_t seed, uint64_t stream_id, cudaStream_t st, const uint64_t* dyn,
                  uint64_t stream_add) {
  PhiloxArgs a;
  a.D = D; a.K = K; a.half = Ns / 2;
  shard_range(a.half, c->nranks, c->rank, &a.pair_begin, &a.pair_end);
  a.seed = seed; a.stream = stream_id;
  a.dyn = dyn;
  a.stream_add = stream_add;
  a.eps = c->eps.d();
  if (a.pair_end <= a.pair_begin) return VBMC_B200_OK;
  const bool f32 = c->precision == 32;
  c->eps_f32 = f32;
  const int per_ctr = f32 ? 4 : 2;
  const long long per_comp = (static_cast<long long>(a.pair_end - a.pair_begin) * D + per_ctr) / per_ctr;
  int bx = static_cast<int>((per_comp + 255) / 256);
  // resident blocks per SM are capped so that the gplogjoint branch (own stream) can co-run instead of queueing behind the generator
  static const int per_sm = getenv("VBMC_B200_PHILOX_BLOCKS_PER_SM") ? atoi(getenv("VBMC_B200_PHILOX_BLOCKS_PER_SM")) : 4;
  const int cap = (c->num_sms * per_sm + K - 1) / K;
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  dim3 grid(bx, K);
  KernelScope ks(c, "philox", st);
  if (f32)
    philox_normal_f32_kernel<<<grid, 256, 0, st>>>(a);
  else
    philox_normal_kernel<<<grid, 256, 0, st>>>(a);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

}  // namespace vb

extern "C" int vbmc_b200_philox_raw(vbmc_b200_ctx* ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  if (!ctx || !ctr || !key || !out) VB_FAIL(VBMC_B200_EINVAL, "vbmc_b200_philox_raw: null argument");
  VB_CUDA(cudaSetDevice(ctx->device));
  uint32_t* d = nullptr;
  VB_CUDA(cudaMalloc(&d, 16));
  {
    vb::KernelScope ks(ctx, "philox_raw", ctx->stream);
    vb::philox_raw_kernel<<<1, 1, 0, ctx->stream>>>(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]), d);
  }
  VB_CUDA(cudaMemcpyAsync(out, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
  VB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d);
  return VBMC_B200_OK;
}
