// Device-side source of the entropy draws: counter-based Philox4x32-10 (Salmon et al., SC'11)
// feeding a 1024-strip ziggurat for the standard normal (Marsaglia & Tsang 2000, in Doornik's ZIGNOR
// formulation with a floating-point uniform) — the family of algorithm MATLAB's own randn uses.
// Replaces MATLAB's global randn stream (ent/entmc_vbmc.m:53), which cannot be reproduced outside
// MATLAB; draws depend only on (seed, stream, element index), never on the number of GPUs or on
// the sharding, so any rank can regenerate any slice.  tests/test_devgen.py pins the output element
// by element against a NumPy restatement and checks its distribution.
//
// Why a ziggurat: a step is FP64-pipe bound as a whole, and Box-Muller spends ~50
// FP64 instructions per draw (log, sqrt, sincospi) where the ziggurat's accepted attempt costs one
// table look-up, one compare and one multiply.  1024 strips (99.7 % accepted at once) rather than
// the usual 128/256 because a warp pays for the slow path whenever ANY of its 32 lanes takes it:
// 256 strips were measured slower than Box-Muller on B200 (64 % of the warp iterations diverged),
// 1024 strips keep 90 % of them on the fast path.  FP32 mode keeps Box-Muller on the MUFU unit.
#include <math.h>
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

// ---- ziggurat (tables built on the host by philox_init_tables, staged in shared memory) ----
constexpr int ZIG_C = 1024;
constexpr double ZIG_R = 4.038849846109504522714;     // right edge of the base strip
constexpr double ZIG_V = 0.001226324646353088072885;  // area of every strip

// top 52 bits of w as an exact double in [0, 2^52)
__device__ __forceinline__ double bits52(uint64_t w) {
  return __longlong_as_double(static_cast<long long>(0x4330000000000000ULL | (w >> 12))) - 4503599627370496.0;
}
__device__ __forceinline__ double u_sym(uint64_t w) { return fma(bits52(w), 4.440892098500626e-16, 2.220446049250313e-16 - 1.0); }  // (m+.5)2^-51 - 1
__device__ __forceinline__ double u_01(uint64_t w) { return fma(bits52(w), 2.220446049250313e-16, 1.1102230246251565e-16); }        // (m+.5)2^-52

// key of the extra Philox calls a rejected attempt needs (same counter, perturbed key)
__device__ __forceinline__ uint2 retry_key(uint2 key, uint32_t round, int which) {
  return make_uint2(key.x ^ (0xA5A5A5A5u + 0x9E3779B9u * round), key.y ^ (which ? 0xC2B2AE35u : 0x27D4EB2Fu));
}

// One standard normal from the 64-bit word w of element `which` (0/1) of counter ctr.  zx[0..C]: strip edges (zx[0] the
// virtual width of the base strip, zx[1] = R, decreasing to zx[C] = 0); zr[i] = zx[i+1]/zx[i]; zf[i] = exp(-zx[i]^2/2).
__device__ __noinline__ double zig_slow(uint64_t w, const uint4 ctr, const uint2 key, const int which, const double* __restrict__ zx,
                                        const double* __restrict__ zr, const double* __restrict__ zf) {
  uint32_t round = 0;
  for (;;) {
    const int i = static_cast<int>(w & (ZIG_C - 1));
    const double u = u_sym(w);
    if (fabs(u) < zr[i]) return u * zx[i];
    uint4 q = philox4x32_10(ctr, retry_key(key, ++round, which));
    uint64_t wa = (static_cast<uint64_t>(q.y) << 32) | q.x, wb = (static_cast<uint64_t>(q.w) << 32) | q.z;
    if (i == 0) {  // base strip: the tail beyond R (Marsaglia 1964)
      for (;;) {
        const double xx = -log(u_01(wa)) / ZIG_R, yy = -log(u_01(wb));
        if (yy + yy >= xx * xx) return u < 0.0 ? -(ZIG_R + xx) : ZIG_R + xx;
        q = philox4x32_10(ctr, retry_key(key, ++round, which));
        wa = (static_cast<uint64_t>(q.y) << 32) | q.x;
        wb = (static_cast<uint64_t>(q.w) << 32) | q.z;
      }
    }
    const double xv = u * zx[i];
    if (zf[i] + u_01(wb) * (zf[i + 1] - zf[i]) < exp(-0.5 * xv * xv)) return xv;  // uniform height inside the strip vs the density
    w = wa;  // rejected: next attempt with a fresh (strip, u) word
  }
}

struct PhiloxArgs {
  int D, K, half, pair_begin, pair_end;
  uint64_t seed, stream;
  uint64_t stream_add;  // added to the stream id (1 when generating the NEXT step's draws ahead of time)
  const uint64_t* dyn;  // optional device pointer to {seed, stream} (graph replay: values change, the node does not)
  double* eps;  // [K][half][D]  (floats in FP32 mode)
  const double* ztab;  // ziggurat tables: zx[C+1] | zr[C] | zf[C+1]
};

// one thread per Philox counter = two consecutive elements (2c, 2c+1) of the flat eps array
__global__ void __launch_bounds__(256) philox_normal_kernel(const PhiloxArgs a) {
  __shared__ double ztab[3 * ZIG_C + 2];
  for (int i = threadIdx.x; i < 3 * ZIG_C + 2; i += blockDim.x) ztab[i] = __ldg(a.ztab + i);
  const double* zx = ztab;
  const double* zr = ztab + ZIG_C + 1;
  const double* zf = ztab + 2 * ZIG_C + 1;
  __syncthreads();
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
    const uint64_t w0 = (static_cast<uint64_t>(r.y) << 32) | r.x, w1 = (static_cast<uint64_t>(r.w) << 32) | r.z;
    // accepted on the first attempt (98.8 %): one look-up, one compare, one multiply; everything else out of line
    const int i0 = static_cast<int>(w0 & (ZIG_C - 1)), i1 = static_cast<int>(w1 & (ZIG_C - 1));
    const double u0 = u_sym(w0), u1 = u_sym(w1);
    double z0 = u0 * zx[i0], z1 = u1 * zx[i1];
    if (!(fabs(u0) < zr[i0])) z0 = zig_slow(w0, ctr, key, 0, zx, zr, zf);
    if (!(fabs(u1) < zr[i1])) z1 = zig_slow(w1, ctr, key, 1, zx, zr, zf);
    const long long e0 = 2 * c, e1 = 2 * c + 1;
    if (e0 >= e_begin && e0 < e_end) a.eps[e0] = z0;
    if (e1 >= e_begin && e1 < e_end) a.eps[e1] = z1;
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

// ziggurat tables, computed once per context on the host (tests/test_devgen.py checks the strips' equal areas)
int philox_init_tables(vbmc_b200_ctx* c) {
  static double t[3 * ZIG_C + 2];
  double* x = t;
  double* r = t + ZIG_C + 1;
  double* fz = t + 2 * ZIG_C + 1;
  double f = exp(-0.5 * ZIG_R * ZIG_R);
  x[0] = ZIG_V / f;
  x[1] = ZIG_R;
  for (int i = 2; i < ZIG_C; ++i) {
    x[i] = sqrt(-2.0 * log(ZIG_V / x[i - 1] + f));
    f = exp(-0.5 * x[i] * x[i]);
  }
  x[ZIG_C] = 0.0;
  for (int i = 0; i < ZIG_C; ++i) r[i] = x[i + 1] / x[i];
  for (int i = 0; i <= ZIG_C; ++i) fz[i] = exp(-0.5 * x[i] * x[i]);
  VB_TRY(c->zigTab.reserve(sizeof(t)));
  VB_CUDA(cudaMemcpy(c->zigTab.p, t, sizeof(t), cudaMemcpyHostToDevice));
  return VBMC_B200_OK;
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
  a.ztab = c->zigTab.d();
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
