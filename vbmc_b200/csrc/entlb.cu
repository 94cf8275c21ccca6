// entlb_vbmc on the device: the deterministic entropy lower bound negelcbo_vbmc uses when Ns == 0
// (ent/entlb_vbmc.m:1-147, negelcbo_vbmc.m:102-109).  Four tiny launches, one thread per output element of each
// stage; the element functions live in entlb_math.cuh and are the ones tests/test_entlb_host.py checks on the CPU.
#include "common.cuh"
#include "entlb_math.cuh"

namespace vb {

template <int STAGE>
__global__ void __launch_bounds__(128) entlb_stage_kernel(const EntlbArgs a, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) entlb_stage_elem(a, STAGE, idx);
}

template <int STAGE>
static int launch_stage(vbmc_b200_ctx* c, const EntlbArgs& a, cudaStream_t st) {
  const int n = entlb_stage_size(a, STAGE);
  if (n <= 0) return VBMC_B200_OK;
  KernelScope ks(c, "entlb", st);
  entlb_stage_kernel<STAGE><<<(n + 127) / 128, 128, 0, st>>>(a, n);
  VB_CUDA(cudaGetLastError());
  return VBMC_B200_OK;
}

// H_lb and dH_lb of the vp currently unpacked on the device (c->vp, written by vp_unpack_kernel), to the host.
// gmask: bit i = grad_flags(i+1).  dH (host, may be NULL) receives the requested blocks packed like entmc's dH.
int run_entlb(vbmc_b200_ctx* c, int gmask, int jacobian, double* H, double* dH) {
  const int D = c->D, K = c->K;
  EntlbArgs a;
  a.D = D; a.K = K; a.jacobian = jacobian ? 1 : 0;
  for (int i = 0; i < 4; ++i) a.gf[i] = (gmask >> i) & 1;
  const int n = entlb_ngrad(a);
  const size_t nwork = static_cast<size_t>(K) * K + 2 * static_cast<size_t>(K) + 1 + static_cast<size_t>(D) * K + 2 * K + D;
  VB_TRY(c->entlbWork.reserve(sizeof(double) * nwork));
  double* p = c->entlbWork.d();
  a.gamma = p; p += static_cast<size_t>(K) * K;
  a.gsum = p; p += K;
  a.wraw = p; p += K;
  a.out = p;
  a.mu = c->vp.mu; a.sigma = c->vp.sigma; a.lambda = c->vp.lambda; a.w = c->vp.w; a.eta = c->vp.eta;
  VB_CUDA(cudaMemsetAsync(a.out, 0, sizeof(double) * (1 + static_cast<size_t>(n)), c->stream));
  VB_TRY(launch_stage<0>(c, a, c->stream));
  VB_TRY(launch_stage<1>(c, a, c->stream));
  VB_TRY(launch_stage<2>(c, a, c->stream));
  VB_TRY(launch_stage<3>(c, a, c->stream));
  std::vector<double> h(1 + static_cast<size_t>(n));
  VB_CUDA(cudaMemcpyAsync(h.data(), a.out, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, c->stream));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  if (H) *H = h[0];
  if (dH && n) memcpy(dH, h.data() + 1, sizeof(double) * n);
  return VBMC_B200_OK;
}

}  // namespace vb

// [H,dH] = entlb_vbmc(vp,grad_flags,jacobian_flag) for the vp last set on the context (no theta unpacking)
extern "C" int vbmc_b200_entlb(vbmc_b200_ctx* c, const int grad_flags[4], int jacobian_flag, double* H, double* dH) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  if (!c->vp_ready) VB_FAIL(VBMC_B200_ESTATE, "entlb: call vbmc_b200_vp_set first");
  VB_CUDA(cudaSetDevice(c->device));
  int gmask = 0;
  if (grad_flags && dH)
    for (int i = 0; i < 4; ++i)
      if (grad_flags[i]) gmask |= 1 << i;
  VB_TRY(vb::launch_vp_unpack(c, false));
  return vb::run_entlb(c, gmask, jacobian_flag, H, dH);
}
