// C ABI of libvbmc_b200 (see include/vbmc_b200.h): context, GP/VP residency, one negelcbo step.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace vb {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void shard_range(int total, int nranks, int rank, int* begin, int* end) {
  // contiguous balanced split: the first (total % nranks) ranks get one extra unit
  const int base = total / nranks, rem = total % nranks;
  const int b = rank * base + (rank < rem ? rank : rem);
  *begin = b;
  *end = b + base + (rank < rem ? 1 : 0);
}

KernelScope::KernelScope(vbmc_b200_ctx* ctx, const char* nm, cudaStream_t s) : c(ctx), name(nm), st(s) {
  c->launches++;
  if (c->profiling) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
  }
}
KernelScope::~KernelScope() {
  if (c->profiling && e0) {
    cudaEventRecord(e1, st);
    c->prof_events.push_back(e0);
    c->prof_events.push_back(e1);
    c->prof[name].n += 1;
    c->prof_names.push_back(name);
  }
}
int profile_collect(vbmc_b200_ctx* c) {
  VB_CUDA(cudaDeviceSynchronize());
  for (size_t i = 0; i + 1 < c->prof_events.size(); i += 2) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->prof_events[i], c->prof_events[i + 1]);
    c->prof[c->prof_names[i / 2]].ms += ms;
    cudaEventDestroy(c->prof_events[i]);
    cudaEventDestroy(c->prof_events[i + 1]);
  }
  c->prof_events.clear();
  c->prof_names.clear();
  return VBMC_B200_OK;
}

static int ensure_pinned(double** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return VBMC_B200_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *cap = 0;
  VB_CUDA(cudaMallocHost(reinterpret_cast<void**>(p), bytes));
  *cap = bytes;
  return VBMC_B200_OK;
}

static int grad_mask_len(const vbmc_b200_ctx* c, int mask) {
  int n = 0;
  if (mask & 1) n += c->D * c->K;
  if (mask & 2) n += c->K;
  if (mask & 4) n += c->D;
  if (mask & 8) n += c->K;
  return n;
}

}  // namespace vb

using namespace vb;

extern "C" {

int vbmc_b200_version(void) { return VBMC_B200_VERSION; }
const char* vbmc_b200_last_error(void) { return vb::g_err; }

int vbmc_b200_create(vbmc_b200_ctx** out, int device) {
  if (!out) VB_FAIL(VBMC_B200_EINVAL, "vbmc_b200_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    VB_FAIL(VBMC_B200_ENODEV, "vbmc_b200:nodevice: no CUDA device (%s); this library has no CPU fallback",
            e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) VB_FAIL(VBMC_B200_EINVAL, "vbmc_b200_create: device %d out of range [0,%d)", device, ndev);
  cudaDeviceProp prop;
  VB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    VB_FAIL(VBMC_B200_ENODEV, "vbmc_b200:nodevice: device %d (%s) is sm_%d%d; this library is built for sm_100a only",
            device, prop.name, prop.major, prop.minor);
  VB_CUDA(cudaSetDevice(device));
  vbmc_b200_ctx* c = new vbmc_b200_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  // The step's own kernels outrank the ahead-of-time draw generator (stream3): the generator fills the whole GPU for 40-150 us
  // right behind the sweep, and without priorities the small reduction / finalize kernels of the step queue behind its CTAs.
  {
    int prio_lo = 0, prio_hi = 0;   // numerically lower = higher priority
    VB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    static const bool flat = getenv("VBMC_B200_STREAM_PRIORITIES") && atoi(getenv("VBMC_B200_STREAM_PRIORITIES")) == 0;
    VB_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, flat ? prio_lo : prio_hi));
    VB_CUDA(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, flat ? prio_lo : prio_hi));
    VB_CUDA(cudaStreamCreateWithPriority(&c->stream3, cudaStreamNonBlocking, prio_lo));
  }
  VB_CUDA(cudaEventCreateWithFlags(&c->ev_fork0, cudaEventDisableTiming));
  VB_CUDA(cudaEventCreateWithFlags(&c->ev_philox, cudaEventDisableTiming));
  VB_CUDA(cudaEventCreateWithFlags(&c->ev_glj, cudaEventDisableTiming));
  VB_CUDA(cudaEventCreateWithFlags(&c->ev_la_main, cudaEventDisableTiming));
  VB_CUDA(cudaEventCreateWithFlags(&c->ev_la_side, cudaEventDisableTiming));
  VB_CUDA(cudaEventCreateWithFlags(&c->ev_trail_fork, cudaEventDisableTiming));
  VB_CUDA(cudaEventCreateWithFlags(&c->ev_trail, cudaEventDisableTiming));
  if (const char* pf = getenv("VBMC_B200_PREFETCH")) c->prefetch_enabled = strcmp(pf, "0") != 0;
  if (const char* gf = getenv("VBMC_B200_GLJ_FIRST")) c->glj_first = strcmp(gf, "0") != 0;
  VB_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  VB_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  VB_CUDA(cudaEventCreate(&c->ev_t0));
  VB_CUDA(cudaEventCreate(&c->ev_t1));
  {
    const int rc = vb::philox_init_tables(c);
    if (rc != VBMC_B200_OK) {
      vbmc_b200_destroy(c);
      return rc;
    }
  }
  if (const char* g = getenv("VBMC_B200_GRAPHS")) c->graphs_enabled = strcmp(g, "0") != 0;
  if (const char* pc = getenv("VBMC_B200_ENTMC_PRUNE")) c->entmc_prune_c = atof(pc);
  if (const char* pb = getenv("VBMC_B200_ENTMC_BALANCE")) c->ent_balance = atoi(pb) != 0;
  if (const char* pr = getenv("VBMC_B200_ENTMC_CRUN")) c->ent_balance_crun = atoi(pr) < 0 ? 0 : (atoi(pr) > 4096 ? 4096 : atoi(pr));
  if (const char* p0 = getenv("VBMC_B200_ENTMC_C0")) c->ent_balance_c0 = atoi(p0) < 1 ? 1 : (atoi(p0) > 4096 ? 4096 : atoi(p0));
  if (const char* f = getenv("VBMC_B200_ENTMC_FORM")) {
    if (!strcmp(f, "separable")) c->entmc_form = 0;
    if (!strcmp(f, "direct")) c->entmc_form = 1;
    if (!strcmp(f, "expanded")) c->entmc_form = 2;
  }
  *out = c;
  return VBMC_B200_OK;
}

int vbmc_b200_destroy(vbmc_b200_ctx* c) {
  if (!c) return VBMC_B200_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  vb::comm_destroy(c);
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  if (c->refit_graph) cudaGraphExecDestroy(c->refit_graph);
  if (c->adam_graph) cudaGraphExecDestroy(c->adam_graph);
  for (int i = 0; i < 2; ++i)
    if (c->trsm_graph[i]) cudaGraphExecDestroy(c->trsm_graph[i]);
  vb::DevBuf* bufs[] = {&c->gpX, &c->gpHyp, &c->gpAlpha, &c->gpDerived, &c->gpL, &c->gpY, &c->gpS2, &c->gpWork, &c->gpFlags, &c->trsvFlags,
                        &c->vpBase, &c->vpCur, &c->bnd, &c->eps, &c->theta_dev, &c->out_dev, &c->R_dev,
                        &c->ent_partial, &c->ent_partial2, &c->ent_plan, &c->glj_out, &c->flush, &c->varWork, &c->adamState, &c->adamXtab, &c->ent_tables, &c->entmc_prune_stats, &c->glj_part, &c->glj_ticket, &c->predWork, &c->trsmWork, &c->gpXalt, &c->gpAlphaAlt, &c->zigTab, &c->entlbWork};
  for (auto* b : bufs) b->release();
  if (c->theta_pinned) cudaFreeHost(c->theta_pinned);
  if (c->out_pinned) cudaFreeHost(c->out_pinned);
  for (auto ev : c->prof_events) cudaEventDestroy(ev);
  cudaEventDestroy(c->ev_fork);
  cudaEventDestroy(c->ev_join);
  cudaEventDestroy(c->ev_t0);
  cudaEventDestroy(c->ev_t1);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->stream2);
  cudaStreamDestroy(c->stream3);
  cudaEventDestroy(c->ev_fork0);
  cudaEventDestroy(c->ev_philox);
  cudaEventDestroy(c->ev_glj);
  cudaEventDestroy(c->ev_la_main);
  cudaEventDestroy(c->ev_la_side);
  cudaEventDestroy(c->ev_trail_fork);
  cudaEventDestroy(c->ev_trail);
  delete c;
  return VBMC_B200_OK;
}

int vbmc_b200_gp_tag_set(vbmc_b200_ctx* c, unsigned long long tag) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  c->gp_tag = c->gp_ready ? tag : 0;
  return VBMC_B200_OK;
}
int vbmc_b200_gp_tag_get(vbmc_b200_ctx* c, unsigned long long* tag) {
  if (!c || !tag) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  *tag = c->gp_ready ? c->gp_tag : 0;
  return VBMC_B200_OK;
}

int vbmc_b200_gp_shape(vbmc_b200_ctx* c, int* N, int* D, int* S) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  if (N) *N = c->gp_ready ? c->gp.N : 0;
  if (D) *D = c->gp_ready ? c->gp.D : 0;
  if (S) *S = c->gp_ready ? c->gp.S : 0;
  return VBMC_B200_OK;
}

// process-wide contexts, one per device (see include/vbmc_b200.h)
static vbmc_b200_ctx* g_shared[64] = {nullptr};
int vbmc_b200_shared(vbmc_b200_ctx** out, int device) {
  if (!out) VB_FAIL(VBMC_B200_EINVAL, "vbmc_b200_shared: out is NULL");
  *out = nullptr;
  if (device < 0 || device >= 64) VB_FAIL(VBMC_B200_EINVAL, "vbmc_b200_shared: device %d out of range", device);
  if (!g_shared[device]) VB_TRY(vbmc_b200_create(&g_shared[device], device));
  *out = g_shared[device];
  return VBMC_B200_OK;
}
int vbmc_b200_shared_release(void) {
  for (int d = 0; d < 64; ++d) {
    if (g_shared[d]) vbmc_b200_destroy(g_shared[d]);
    g_shared[d] = nullptr;
  }
  return VBMC_B200_OK;
}

int vbmc_b200_sync(vbmc_b200_ctx* c) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  VB_CUDA(cudaSetDevice(c->device));
  VB_CUDA(cudaStreamSynchronize(c->stream3));
  VB_CUDA(cudaStreamSynchronize(c->stream2));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  return VBMC_B200_OK;
}

int vbmc_b200_set_precision(vbmc_b200_ctx* c, int bits) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  if (bits != 32 && bits != 64) VB_FAIL(VBMC_B200_EINVAL, "vbmc_b200_set_precision: bits must be 32 or 64 (got %d)", bits);
  c->precision = bits;
  return VBMC_B200_OK;
}

int vbmc_b200_entmc_prune(vbmc_b200_ctx* c, double log_threshold) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  if (!(log_threshold >= 0.0)) VB_FAIL(VBMC_B200_EINVAL, "vbmc_b200_entmc_prune: threshold must be >= 0 (0 disables pruning)");
  c->entmc_prune_c = log_threshold;
  return VBMC_B200_OK;
}

int vbmc_b200_entmc_balance(vbmc_b200_ctx* c, int on, int c0) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  c->ent_balance = on != 0;
  if (c0 > 0) c->ent_balance_c0 = c0 > 4096 ? 4096 : c0;
  return VBMC_B200_OK;
}

int vbmc_b200_entmc_plan_get(vbmc_b200_ctx* c, int* out, int cap, int* G, int* tpc) {
  if (!c || !out || !G || !tpc) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  *G = 0;
  *tpc = 0;
  if (!c->ent_plan_active || !c->ent_plan.p) return VBMC_B200_OK;
  const int n = c->ent_plan_req_G + 1 + 2 * c->K;
  if (cap < n) VB_FAIL(VBMC_B200_EINVAL, "vbmc_b200_entmc_plan_get: out holds %d ints, %d needed", cap, n);
  VB_CUDA(cudaSetDevice(c->device));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  VB_CUDA(cudaMemcpy(out, c->ent_plan.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
  *G = c->ent_plan_req_G;
  *tpc = c->ent_plan_req_tpc;
  return VBMC_B200_OK;
}

int vbmc_b200_entmc_prune_stats(vbmc_b200_ctx* c, int enable, unsigned long long* kept, unsigned long long* total) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  VB_CUDA(cudaSetDevice(c->device));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  unsigned long long h[2] = {0ull, 0ull};
  if (c->entmc_prune_stats.p) VB_CUDA(cudaMemcpy(h, c->entmc_prune_stats.p, sizeof(h), cudaMemcpyDeviceToHost));
  if (kept) *kept = h[0];
  if (total) *total = h[1];
  VB_TRY(c->entmc_prune_stats.reserve(sizeof(h)));
  VB_CUDA(cudaMemset(c->entmc_prune_stats.p, 0, sizeof(h)));
  c->entmc_prune_stats_on = enable != 0;
  return VBMC_B200_OK;
}

int vbmc_b200_launch_count(vbmc_b200_ctx* c, long long* count) {
  if (!c || !count) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  *count = c->launches;
  return VBMC_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// GP
// ---------------------------------------------------------------------------------------------
static int gp_check_desc(const vbmc_b200_gp_desc* g, int* Ncov, int* Nnoise, int* Nmean) {
  if (!g || !g->X || !g->hyp) VB_FAIL(VBMC_B200_EINVAL, "gp descriptor: X and hyp are required");
  if (g->N <= 0 || g->D <= 0 || g->S <= 0) VB_FAIL(VBMC_B200_EINVAL, "gp descriptor: N, D, S must be positive");
  if (g->covfun != 1)
    VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:UnsupportedCovFun: only covfun=1 (SE-ARD) is in scope (got %d)", g->covfun);
  *Ncov = g->D + 1;  // gplite_covfun('info'): SE-ARD
  int nn = 0;
  if (g->noisefun[0] == 1) nn += 1;
  if (g->noisefun[1] == 2) nn += 1;
  if (g->noisefun[2] == 1) nn += 2;
  *Nnoise = nn;
  switch (g->meanfun) {
    case 0: *Nmean = 0; break;
    case 1: *Nmean = 1; break;
    case 4: *Nmean = 1 + 2 * g->D; break;
    default:
      VB_FAIL(VBMC_B200_EREFERENCE,
              "gplogjoint:UnsupportedMeanFun: this build supports meanfun 0 (zero), 1 (const), 4 (negquad); got %d",
              g->meanfun);
  }
  if (g->Nhyp != *Ncov + *Nnoise + *Nmean)
    VB_FAIL(VBMC_B200_EREFERENCE,
            "gplite_post:dimmismatch: Number of hyperparameters mismatched with GP model specification (Nhyp=%d, expected %d).",
            g->Nhyp, *Ncov + *Nnoise + *Nmean);
  return VBMC_B200_OK;
}

}  // extern "C"
namespace vb {
// derived per-sample constants used by gplogjoint (gplogjoint.m:99-122)
int gp_upload_derived(vbmc_b200_ctx* c, const vbmc_b200_gp_desc* g, int Ncov, int Nnoise, const double* sW1, const int* Lchol) {
  const int D = g->D, S = g->S;
  std::vector<double> h(static_cast<size_t>(S) * (3 * D + 3));
  double* ell = h.data();
  double* lnc = ell + S * D;
  double* m0 = lnc + S;
  double* xm = m0 + S;
  double* iom2 = xm + S * D;
  double* sn2eff = iom2 + S * D;
  for (int s = 0; s < S; ++s) {
    const double* hyp = g->hyp + static_cast<size_t>(s) * g->Nhyp;
    double sum_lnell = 0.0;
    for (int d = 0; d < D; ++d) {
      ell[s * D + d] = exp(hyp[d]);
      sum_lnell += hyp[d];
    }
    lnc[s] = 2.0 * hyp[D] + sum_lnell;  // ln_sf2 + sum_lnell
    m0[s] = g->meanfun > 0 ? hyp[Ncov + Nnoise] : 0.0;
    for (int d = 0; d < D; ++d) {
      if (g->meanfun == 4) {
        xm[s * D + d] = hyp[Ncov + Nnoise + 1 + d];
        const double om = exp(hyp[Ncov + Nnoise + D + 1 + d]);
        iom2[s * D + d] = 1.0 / (om * om);
      } else {
        xm[s * D + d] = 0.0;
        iom2[s * D + d] = 0.0;
      }
    }
    // low-noise posterior (Lchol == 0): K^-1 is used unscaled (gplogjoint.m:279), i.e. sn2_eff plays no role
    sn2eff[s] = (sW1 && !(Lchol && !Lchol[s])) ? 1.0 / (sW1[s] * sW1[s]) : 1.0;
  }
  c->gpSn2effHost.assign(sn2eff, sn2eff + S);
  VB_TRY(c->gpDerived.reserve(h.size() * sizeof(double)));
  VB_CUDA(cudaMemcpyAsync(c->gpDerived.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  double* base = c->gpDerived.d();
  c->gp.ell = base;
  c->gp.lnc = base + S * D;
  c->gp.m0 = base + S * D + S;
  c->gp.xm = base + S * D + 2 * S;
  c->gp.iom2 = base + 2 * S * D + 2 * S;
  c->gp.sn2eff = base + 3 * S * D + 2 * S;
  return VBMC_B200_OK;
}
}  // namespace vb
extern "C" {

int vbmc_b200_gp_attach(vbmc_b200_ctx* c, const vbmc_b200_gp_desc* g, const double* alpha, const double* sW1,
                        const int* Lchol, const double* L) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  c->gp_tag = 0;  // the resident posterior is about to change: a caller's fingerprint of it no longer holds
  int Ncov, Nnoise, Nmean;
  VB_TRY(gp_check_desc(g, &Ncov, &Nnoise, &Nmean));
  if (!alpha) VB_FAIL(VBMC_B200_EINVAL, "gp_attach: alpha is required");
  VB_CUDA(cudaSetDevice(c->device));
  c->gp_ready = false;
  const size_t N = g->N, D = g->D, S = g->S;
  VB_TRY(c->gpX.reserve(N * D * sizeof(double)));
  VB_TRY(c->gpHyp.reserve(S * g->Nhyp * sizeof(double)));
  VB_TRY(c->gpAlpha.reserve(S * N * sizeof(double)));
  VB_CUDA(cudaMemcpyAsync(c->gpX.p, g->X, N * D * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  VB_CUDA(cudaMemcpyAsync(c->gpHyp.p, g->hyp, S * g->Nhyp * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  VB_CUDA(cudaMemcpyAsync(c->gpAlpha.p, alpha, S * N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->gpHasL = false;
  c->gpLd = g->N;
  if (L) {
    // same layout as the factors gp_post leaves behind: leading dimension Np = 64*ceil((N+1)/64), unit diagonal in the padding
    // (the blocked solves of trsm.cu read whole 64 x 64 tiles in place)
    const size_t Np = (N + 1 + 63) / 64 * 64;
    VB_TRY(c->gpL.reserve(S * Np * Np * sizeof(double)));
    VB_CUDA(cudaMemsetAsync(c->gpL.p, 0, S * Np * Np * sizeof(double), c->stream));
    for (size_t s = 0; s < S; ++s)
      VB_CUDA(cudaMemcpy2DAsync(c->gpL.d() + s * Np * Np, Np * sizeof(double), L + s * N * N, N * sizeof(double), N * sizeof(double), N,
                                cudaMemcpyHostToDevice, c->stream));
    VB_TRY(vb::pad_identity(c->gpL.d(), static_cast<int>(N), static_cast<int>(Np), static_cast<int>(S), c->stream));
    c->gpLd = static_cast<int>(Np);
    c->gpHasL = true;
  }
  c->gpLchol.assign(S, 1);
  if (Lchol)
    for (size_t s = 0; s < S; ++s) c->gpLchol[s] = Lchol[s];
  c->gpLfactor = c->gpLchol;
  c->gpSn2mult.assign(S, 1.0);
  c->gpHypHost.assign(g->hyp, g->hyp + S * g->Nhyp);
  for (int i = 0; i < 3; ++i) c->gp_noisefun[i] = g->noisefun[i];
  c->gp.N = g->N; c->gp.D = g->D; c->gp.S = g->S; c->gp.Nhyp = g->Nhyp;
  c->gp.Ncov = Ncov; c->gp.Nnoise = Nnoise; c->gp.Nmean = Nmean; c->gp.meanfun = g->meanfun;
  c->gp.X = c->gpX.d();
  c->gp.hyp = c->gpHyp.d();
  c->gp.alpha = c->gpAlpha.d();
  VB_TRY(gp_upload_derived(c, g, Ncov, Nnoise, sW1, Lchol));
  c->gp_ready = true;
  return VBMC_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// VP / thetabnd
// ---------------------------------------------------------------------------------------------
int vbmc_b200_vp_set(vbmc_b200_ctx* c, const vbmc_b200_vp_desc* v) {
  if (!c || !v) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  if (v->D <= 0 || v->K <= 0 || !v->mu || !v->sigma || !v->lambda || !v->w)
    VB_FAIL(VBMC_B200_EINVAL, "vp descriptor: D, K > 0 and mu, sigma, lambda, w are required");
  VB_CUDA(cudaSetDevice(c->device));
  const int D = v->D, K = v->K;
  c->vp_ready = false;
  c->D = D; c->K = K;
  c->opt[0] = v->optimize_mu != 0; c->opt[1] = v->optimize_sigma != 0;
  c->opt[2] = v->optimize_lambda != 0; c->opt[3] = v->optimize_weights != 0;
  c->ntheta = (c->opt[0] ? D * K : 0) + (c->opt[1] ? K : 0) + (c->opt[2] ? D : 0) + (c->opt[3] ? K : 0);
  // base: mu[DK] sigma[K] lambda[D] w[K] eta[K] ; cur: mu[DK] sigma[K] lambda[D] w[K] eta[K] lnsigma[K]
  //       lnlambda[D] delta[D] ck[K] ak[K] cn[K+1]
  const size_t nbase = static_cast<size_t>(D) * K + 3 * K + D;
  std::vector<double> h(nbase);
  double* p = h.data();
  memcpy(p, v->mu, sizeof(double) * D * K); p += D * K;
  memcpy(p, v->sigma, sizeof(double) * K); p += K;
  memcpy(p, v->lambda, sizeof(double) * D); p += D;
  memcpy(p, v->w, sizeof(double) * K); p += K;
  for (int k = 0; k < K; ++k) p[k] = v->eta ? v->eta[k] : log(v->w[k]);
  VB_TRY(c->vpBase.reserve(nbase * sizeof(double)));
  VB_CUDA(cudaMemcpyAsync(c->vpBase.p, h.data(), nbase * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  double* b = c->vpBase.d();
  c->base_mu = b; b += D * K;
  c->base_sigma = b; b += K;
  c->base_lambda = b; b += D;
  c->base_w = b; b += K;
  c->base_eta = b;
  const int DPc = vb::entmc_pick_dp(D) > 0 ? vb::entmc_pick_dp(D) : D + (D & 1);
  const int K2c = (K + 1) & ~1;
  c->vp_cblob_dp = DPc;
  c->vp_cblob_len = K2c * DPc + 2 * K2c + DPc;
  const size_t ncur = 2 * static_cast<size_t>(D) * K + 6 * K + 3 * D + 1 + K + 2 + 2 + c->vp_cblob_len;
  VB_TRY(c->vpCur.reserve(ncur * sizeof(double)));
  double* q = c->vpCur.d();
  c->vp.D = D; c->vp.K = K;
  c->vp.mu = q; q += D * K;
  c->vp.sigma = q; q += K;
  c->vp.lambda = q; q += D;
  c->vp.w = q; q += K;
  c->vp.eta = q; q += K;
  c->vp.lnsigma = q; q += K;
  c->vp.lnlambda = q; q += D;
  c->vp.delta = q; q += D;
  c->vp.ck = q; q += K;
  c->vp.ak = q; q += K;
  c->vp.cn = q; q += K + 1;
  c->vp.form_flag = reinterpret_cast<int*>(q); q += 2;
  c->vp.dyn_snap = reinterpret_cast<unsigned long long*>(q); q += 2;
  c->vp.scratch = q; q += D * K;
  c->vp.cblob = q; q += c->vp_cblob_len;
  std::vector<double> dl(D, 0.0);
  if (v->delta)
    for (int d = 0; d < D; ++d) dl[d] = v->delta[d];
  VB_CUDA(cudaMemcpyAsync(c->vp.delta, dl.data(), sizeof(double) * D, cudaMemcpyHostToDevice, c->stream));
  VB_CUDA(cudaMemsetAsync(c->vp.form_flag, 0, 16, c->stream));   // {form flag, arrival ticket, guard maximum} of vp_unpack2_kernel
  VB_CUDA(cudaStreamSynchronize(c->stream));
  c->vp_ready = true;
  c->eps_ready = c->eps_ready && c->epsD == D && c->epsK == K;
  return VBMC_B200_OK;
}

int vbmc_b200_thetabnd_set(vbmc_b200_ctx* c, int n, const double* lb, const double* ub, double TolCon,
                           double WeightThreshold, double WeightPenalty) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  VB_CUDA(cudaSetDevice(c->device));
  if (n <= 0) {
    c->nbnd = 0;
    return VBMC_B200_OK;
  }
  if (!lb || !ub) VB_FAIL(VBMC_B200_EINVAL, "thetabnd_set: lb/ub required");
  VB_TRY(c->bnd.reserve(sizeof(double) * 2 * n));
  VB_CUDA(cudaMemcpyAsync(c->bnd.d(), lb, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  VB_CUDA(cudaMemcpyAsync(c->bnd.d() + n, ub, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  c->nbnd = n;
  c->TolCon = TolCon;
  c->WeightThreshold = WeightThreshold;
  c->WeightPenalty = WeightPenalty;
  return VBMC_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// eps
// ---------------------------------------------------------------------------------------------
static int eps_reserve(vbmc_b200_ctx* c, int D, int K, int Ns) {
  if (D <= 0 || K <= 0 || Ns <= 0 || (Ns & 1)) VB_FAIL(VBMC_B200_EINVAL, "eps: D, K > 0 and Ns even > 0 required");
  const size_t bytes = sizeof(double) * static_cast<size_t>(D) * K * (Ns / 2);
  if (bytes > c->eps.cap || c->epsD != D || c->epsK != K || c->epsNs != Ns) c->eps_key.valid = false;  // contents are lost / reinterpreted
  VB_TRY(c->eps.reserve(bytes));
  c->epsD = D; c->epsK = K; c->epsNs = Ns;
  return VBMC_B200_OK;
}

static void eps_key_set(vbmc_b200_ctx* c, uint64_t seed, uint64_t stream, int Ns) {
  c->eps_key.valid = true;
  c->eps_key.seed = seed; c->eps_key.stream = stream;
  c->eps_key.D = c->D; c->eps_key.K = c->K; c->eps_key.Ns = Ns;
  c->eps_key.f32 = c->precision == 32;
}
static bool eps_key_matches(const vbmc_b200_ctx* c, uint64_t seed, uint64_t stream, int Ns) {
  const auto& k = c->eps_key;
  return k.valid && c->eps_ready && k.seed == seed && k.stream == stream && k.D == c->D && k.K == c->K && k.Ns == Ns &&
         k.f32 == (c->precision == 32) && c->epsD == c->D && c->epsK == c->K && c->epsNs == Ns;
}

int vbmc_b200_eps_upload(vbmc_b200_ctx* c, int D, int K, int Ns, const double* eps) {
  if (!c || !eps) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  VB_CUDA(cudaSetDevice(c->device));
  Ns = (Ns + 1) / 2 * 2;
  VB_TRY(eps_reserve(c, D, K, Ns));
  VB_CUDA(cudaMemcpyAsync(c->eps.p, eps, sizeof(double) * static_cast<size_t>(D) * K * (Ns / 2), cudaMemcpyHostToDevice,
                          c->stream));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  c->eps_ready = true;
  c->eps_f32 = false;
  c->eps_key.valid = false;
  return VBMC_B200_OK;
}

int vbmc_b200_eps_philox(vbmc_b200_ctx* c, int D, int K, int Ns, uint64_t seed, uint64_t stream, double* eps_out) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  VB_CUDA(cudaSetDevice(c->device));
  Ns = (Ns + 1) / 2 * 2;
  VB_TRY(eps_reserve(c, D, K, Ns));
  c->eps_key.valid = false;  // (D, K) may differ from the resident vp: never treated as a step's draws
  if (c->nranks > 1 && eps_out)  // a read-back must see every element, not only this rank's shard
    VB_CUDA(cudaMemsetAsync(c->eps.p, 0, sizeof(double) * static_cast<size_t>(D) * K * (Ns / 2), c->stream));
  VB_TRY(launch_philox(c, D, K, Ns, seed, stream, c->stream));
  const size_t n = static_cast<size_t>(D) * K * (Ns / 2);
  if (eps_out && !c->eps_f32)
    VB_CUDA(cudaMemcpyAsync(eps_out, c->eps.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  if (eps_out && c->eps_f32) {  // FP32 mode: the draws are floats on the device; hand them out widened (exactly)
    std::vector<float> tmp(n);
    VB_CUDA(cudaMemcpy(tmp.data(), c->eps.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) eps_out[i] = static_cast<double>(tmp[i]);
  }
  c->eps_ready = true;
  return VBMC_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// one step
// ---------------------------------------------------------------------------------------------
static int prepare_eps(vbmc_b200_ctx* c, int Ns, int mode, const double* eps, uint64_t seed, uint64_t stream_id) {
  const int D = c->D, K = c->K;
  const size_t n = static_cast<size_t>(D) * K * (Ns / 2);
  switch (mode) {
    case VBMC_B200_EPS_HOST:
      if (!eps) VB_FAIL(VBMC_B200_EINVAL, "eps_mode EPS_HOST requires the eps pointer");
      VB_TRY(eps_reserve(c, D, K, Ns));
      VB_CUDA(cudaMemcpyAsync(c->eps.p, eps, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
      c->eps_ready = true;
      c->eps_f32 = false;
      c->eps_key.valid = false;
      return VBMC_B200_OK;
    case VBMC_B200_EPS_RESIDENT:
      if (!c->eps_ready || c->epsD != D || c->epsK != K || c->epsNs != Ns)
        VB_FAIL(VBMC_B200_ESTATE, "eps_mode EPS_RESIDENT: no resident draws of shape D=%d K=%d Ns=%d (have %d %d %d)", D,
                K, Ns, c->epsD, c->epsK, c->epsNs);
      return VBMC_B200_OK;
    case VBMC_B200_EPS_PHILOX:  // generated inside enqueue_step, after the gplogjoint branch has been forked
      VB_TRY(eps_reserve(c, D, K, Ns));
      c->philox_pending = true;
      c->philox_seed = seed;
      c->philox_stream = stream_id;
      c->eps_ready = true;
      eps_key_set(c, seed, stream_id, Ns);  // what the buffer holds once the step has been enqueued
      return VBMC_B200_OK;
  }
  VB_FAIL(VBMC_B200_EINVAL, "unknown eps_mode %d", mode);
}

// multi-GPU steps whose partial sums go through NCCL (peer mapping unavailable or R larger than the exchange slots)
static bool step_uses_nccl(const vbmc_b200_ctx* c) {
  if (c->nranks <= 1) return false;
  RLayout rl;
  rl.init(c->D, c->K, c->gp_ready ? c->gp.S : 0);
  return !(c->p2p_ready && rl.total <= c->xdev.cap);
}

// enqueue the device part of one evaluation; `what` selects negelcbo / entmc / gplogjoint
static int enqueue_step(vbmc_b200_ctx* c, int Ns, int gmask, int use_bnd, int jacobian, int what, bool have_theta) {
  const bool doH = what != FIN_GPLOGJOINT && what != FIN_NEGELCBO_NOENT, doG = what != FIN_ENTMC;
  RLayout rl;
  rl.init(c->D, c->K, (doG || c->gp_ready) ? c->gp.S : 0);
  if (what == FIN_ENTMC) rl.init(c->D, c->K, 0);
  VB_TRY(c->R_dev.reserve(sizeof(double) * rl.total_all));
  VB_CUDA(cudaMemsetAsync(c->R_dev.p, 0, sizeof(double) * rl.total, c->stream));
  // multi-GPU over peer memory: when the step produces every entry of R (entropy AND log-joint), the reduction kernels write
  // their values straight into the peers' inboxes and finalize_kernel only publishes, waits and sums
  const bool pushed = doH && doG && step_push_target(c, rl.S, true).peer != nullptr;
  // the draws do not depend on theta: generate them on a third stream while vp_unpack (and then gplogjoint) run
  const bool philox_now = doH && c->philox_pending;
  const bool trail = doH && c->philox_trail;
  c->philox_trail = false;
  const uint64_t* dyn_src = c->philox_dyn ? reinterpret_cast<const uint64_t*>(c->theta_dev.d() + c->ntheta) : nullptr;
  if (philox_now) {
    c->philox_pending = false;
    VB_CUDA(cudaEventRecord(c->ev_fork0, c->stream));
    VB_CUDA(cudaStreamWaitEvent(c->stream3, c->ev_fork0, 0));
    VB_TRY(launch_philox(c, c->D, c->K, Ns, c->philox_seed, c->philox_stream, c->stream3, dyn_src));
    VB_CUDA(cudaEventRecord(c->ev_philox, c->stream3));
  }
  // cost-weighted schedule of the FP64 sweep: vp_unpack2_kernel builds the tile ranges of this step from theta (entmc2.cu)
  c->ent_plan_req = doH && entmc2_balance_params(c, Ns, &c->ent_plan_req_tpc, &c->ent_plan_req_G);
  VB_TRY(launch_vp_unpack(c, have_theta));
  if (doG) {
    VB_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    VB_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    VB_TRY(launch_gplogjoint(c, 0, c->stream2));
    VB_CUDA(cudaEventRecord(c->ev_glj, c->stream2));
    VB_TRY(launch_glj_reduce(c, c->stream2, doH && doG));
    VB_CUDA(cudaEventRecord(c->ev_join, c->stream2));
  }
  if (doH) {
    if (philox_now) VB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_philox, 0));
    // The entropy sweep's persistent CTAs take a whole SM each (registers + shared memory), so the log-joint kernel cannot
    // co-reside with them.  Default: both are released together once the unpack kernel is done; the log-joint CTAs (enqueued
    // first, two per SM, equal length) go first and the sweep's CTA of an SM starts as soon as THAT SM is free -- no grid-wide
    // dependency between the two kernels (measured with the round's final kernels, same box, bench.py: value 2281 -> 2307,
    // e2e 2131 -> 2176, resident loop 2373 -> 2443 steps/s against the explicit order).  VBMC_B200_GLJ_FIRST=1 makes the sweep
    // wait for the whole log-joint kernel (the mid-round default, better with that round's slower sweep tail).
    if (doG && !philox_now && (c->glj_first || c->profiling)) {   // per-kernel event timing needs the serial order
      VB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_glj, 0));  // the kernel itself; its small reduction may trail behind
    }
    int need = 0;
    if (gmask & 1) need |= NEED_MU;
    if (gmask & (2 | 4)) need |= NEED_E;
    if (gmask & 8) need |= NEED_W;
    VB_TRY(launch_entmc(c, Ns, need, c->stream));
    if (trail) {
      // the sweep has consumed this step's draws: generate the next step's (stream + 1) on stream3, in the shadow of the
      // single-CTA tail (reduce, all-reduce, finalize, Adam update).  With a device-side key the generator reads the
      // snapshot vp_unpack_kernel took at the start of THIS step (adam_step_kernel advances the original concurrently).
      VB_CUDA(cudaEventRecord(c->ev_trail_fork, c->stream));
      VB_CUDA(cudaStreamWaitEvent(c->stream3, c->ev_trail_fork, 0));
      VB_TRY(launch_philox(c, c->D, c->K, Ns, c->philox_seed, c->philox_stream, c->stream3,
                           dyn_src ? reinterpret_cast<const uint64_t*>(c->vp.dyn_snap) : nullptr, 1));
      VB_CUDA(cudaEventRecord(c->ev_trail, c->stream3));
      c->trail_join_pending = true;
    }
    VB_TRY(launch_entmc_reduce(c, Ns, rl.S, c->stream));
  }
  if (doG) VB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  // multi-GPU: the partial sums are all-reduced inside finalize_kernel over NVLink peer memory (finalize.cu exchange_sum);
  // NCCL only when the peer mapping is unavailable or R exceeds the exchange slots
  if (c->nranks > 1 && !(c->p2p_ready && rl.total <= c->xdev.cap)) VB_TRY(allreduce_R(c, rl.total, c->stream));
  VB_TRY(launch_finalize(c, Ns, gmask, use_bnd, jacobian, what, c->stream, pushed));
  return VBMC_B200_OK;
}

// make `stream` wait for a pending ahead-of-time draw generation (no-op when none)
static int join_trail(vbmc_b200_ctx* c) {
  if (!c->trail_join_pending) return VBMC_B200_OK;
  c->trail_join_pending = false;
  VB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_trail, 0));
  return VBMC_B200_OK;
}

static int fetch_out(vbmc_b200_ctx* c, int nout_theta, int S) {
  OutLayout ol;
  ol.init(nout_theta, S, c->K);
  VB_TRY(ensure_pinned(&c->out_pinned, &c->out_pinned_cap, sizeof(double) * ol.total));
  VB_CUDA(cudaMemcpyAsync(c->out_pinned, c->out_dev.p, sizeof(double) * ol.total, cudaMemcpyDeviceToHost, c->stream));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  return VBMC_B200_OK;
}

static int negelcbo_validate(vbmc_b200_ctx* c, const vbmc_b200_negelcbo_args* a, double* beta, int* Ns, int* gmask) {
  if (!c || !a) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  if (!c->vp_ready) VB_FAIL(VBMC_B200_ESTATE, "negelcbo: call vbmc_b200_vp_set first");
  if (!c->gp_ready) VB_FAIL(VBMC_B200_ESTATE, "negelcbo: call vbmc_b200_gp_attach or vbmc_b200_gp_post first");
  if (c->gp.D != c->D) VB_FAIL(VBMC_B200_EINVAL, "negelcbo: vp.D=%d but gp.D=%d", c->D, c->gp.D);
  if (a->ntheta != c->ntheta)
    VB_FAIL(VBMC_B200_EINVAL, "negelcbo: numel(theta)=%d but the optimize_* flags imply %d", a->ntheta, c->ntheta);
  if (!a->theta) VB_FAIL(VBMC_B200_EINVAL, "negelcbo: theta is required");
  double b = a->beta;
  if (!isfinite(b)) b = 0.0;  // negelcbo_vbmc.m:15
  *beta = b;
  if (a->compute_grad && b != 0.0 && a->compute_var != 2)
    VB_FAIL(VBMC_B200_EREFERENCE,
            "negelcbo_vbmc:vargrad: Computation of the gradient of ELBO with full variance not supported.");
  if (a->separate_K && a->compute_grad)
    VB_FAIL(VBMC_B200_EREFERENCE,
            "negelcbo_vbmc:separateK: Computing the gradient of variational parameters and requesting per-component "
            "results at the same time.");
  if (c->opt[3] && !c->opt[0] && !c->opt[1] && !c->opt[2])
    VB_FAIL(VBMC_B200_EUNSUPPORTED,
            "vbmc_b200:OutOfScope: weights-only optimisation (gplogjoint_weights.m) is outside this build (SURVEY.md 2 #5)");
  if (a->Ns <= 0)
    VB_FAIL(VBMC_B200_EINVAL, "negelcbo: Ns must be positive here (Ns == 0, the deterministic entropy bound entlb_vbmc, is handled by "
                              "vbmc_b200_negelcbo only; the device fminadam loop and the resident loop need Monte-Carlo draws)");
  if (a->compute_var != 0 && a->compute_var != 1 && a->compute_var != 2)
    VB_FAIL(VBMC_B200_EINVAL, "negelcbo: compute_var must be 0, 1 (full) or 2 (diagonal)");

  if (a->use_thetabnd && c->nbnd > 0) {
    const int expect = (c->opt[0] ? c->D * c->K : 0) + ((c->opt[1] || c->opt[2]) ? c->D * c->K : 0) + (c->opt[3] ? c->K : 0);
    if (expect != c->nbnd)
      VB_FAIL(VBMC_B200_EINVAL, "negelcbo: thetabnd has %d entries, expected %d (vpbounds.m:34-46)", c->nbnd, expect);
  }
  *Ns = (a->Ns + 1) / 2 * 2;  // entmc_vbmc.m:45
  int m = 0;
  if (a->compute_grad)
    for (int i = 0; i < 4; ++i)
      if (c->opt[i]) m |= 1 << i;
  *gmask = m;
  return VBMC_B200_OK;
}

// Average over hyper-parameter samples (gplogjoint.m:398-407): varF, varss from F(s), varF(s)
static void combine_variance(const double* Fs, const std::vector<double>& vF, int S, double* varF, double* varss) {
  if (S > 1) {
    double Fbar = 0.0, vm = 0.0;
    for (int s = 0; s < S; ++s) { Fbar += Fs[s]; vm += vF[s]; }
    Fbar /= S; vm /= S;
    double ss = 0.0, sv = 0.0;
    for (int s = 0; s < S; ++s) { ss += (Fs[s] - Fbar) * (Fs[s] - Fbar); sv += (vF[s] - vm) * (vF[s] - vm); }
    const double varFss = ss / (S - 1);          // :402
    *varss = varFss + sqrt(sv / (S - 1));       // varFss + std(varF) (:403)
    *varF = vm + varFss;                         // :404
  } else {
    *varF = vF[0];
    *varss = 0.0;
  }
}

static void scatter_J(const std::vector<double>& J, int S, int K, double* out) {
  // device J[s][j][k] -> MATLAB J_sjk(s,j,k) column-major
  for (int s = 0; s < S; ++s)
    for (int j = 0; j < K; ++j)
      for (int k = 0; k < K; ++k) out[s + static_cast<size_t>(j) * S + static_cast<size_t>(k) * S * K] = J[(static_cast<size_t>(s) * K + j) * K + k];
}

static void scatter_negelcbo(vbmc_b200_ctx* c, const vbmc_b200_negelcbo_args* a, int nth) {
  OutLayout ol;
  ol.init(nth, c->gp.S, c->K);
  const double* o = c->out_pinned;
  if (a->F) *a->F = o[ol.oF];
  if (a->G) *a->G = o[ol.oG];
  if (a->H) *a->H = o[ol.oH];
  if (a->varF) *a->varF = 0.0;
  if (a->varGss) *a->varGss = 0.0;
  if (a->varG) *a->varG = 0.0;
  if (a->varH) *a->varH = 0.0;
  if (a->dF && nth) memcpy(a->dF, o + ol.oDF, sizeof(double) * nth);
  if (a->dH && nth) memcpy(a->dH, o + ol.oDH, sizeof(double) * nth);
  if (a->I_sk) {  // S x K column-major
    const int S = c->gp.S, K = c->K;
    for (int s = 0; s < S; ++s)
      for (int k = 0; k < K; ++k) a->I_sk[s + static_cast<size_t>(k) * S] = o[ol.oIsk + s * K + k];
  }
}

// Everything a captured step depends on: shapes, flags and the addresses of every buffer a kernel node reads or
// writes.  A graph is replayed only while this signature is unchanged.
static std::vector<long long> step_signature(vbmc_b200_ctx* c, int Ns, int gmask, int use_thetabnd, int philox) {
  std::vector<long long> key = {Ns, gmask, use_thetabnd, philox, c->D, c->K, c->gp.S, c->gp.N, c->ntheta, c->nbnd, c->entmc_form, c->precision,
                                c->opt[0] + 2 * c->opt[1] + 4 * c->opt[2] + 8 * c->opt[3], c->gp.meanfun,
                                reinterpret_cast<long long>(c->theta_dev.p), reinterpret_cast<long long>(c->out_dev.p),
                                reinterpret_cast<long long>(c->R_dev.p), reinterpret_cast<long long>(c->eps.p),
                                reinterpret_cast<long long>(c->ent_partial.p), reinterpret_cast<long long>(c->ent_partial2.p), reinterpret_cast<long long>(c->glj_out.p), reinterpret_cast<long long>(c->glj_part.p), reinterpret_cast<long long>(c->glj_ticket.p),
                                reinterpret_cast<long long>(c->ent_tables.p),
                                reinterpret_cast<long long>(c->vpCur.p), reinterpret_cast<long long>(c->vpBase.p),
                                reinterpret_cast<long long>(c->gpAlpha.p), reinterpret_cast<long long>(c->gpX.p),
                                reinterpret_cast<long long>(c->gpDerived.p), reinterpret_cast<long long>(c->bnd.p),
                                reinterpret_cast<long long>(c->theta_pinned), reinterpret_cast<long long>(c->out_pinned)};
  long long bits[4];
  memcpy(&bits[0], &c->TolCon, 8); memcpy(&bits[1], &c->WeightThreshold, 8); memcpy(&bits[2], &c->WeightPenalty, 8);
  memcpy(&bits[3], &c->entmc_prune_c, 8);
  key.insert(key.end(), bits, bits + 4);
  key.push_back(c->entmc_prune_stats_on ? reinterpret_cast<long long>(c->entmc_prune_stats.p) : 0);
  key.push_back(c->nranks);
  key.push_back(c->rank);
  key.push_back(c->p2p_ready ? reinterpret_cast<long long>(c->xchg_peers.p) : 0);
  key.push_back(c->glj_first);
  key.push_back(c->ent_balance ? c->ent_balance_c0 + 8192LL * c->ent_balance_crun : 0);
  return key;
}

// One negelcbo evaluation on the device: H2D {theta, seed, stream}, the kernels, D2H of the output block.
// Single rank, no profiling: the whole sequence is captured once into a CUDA graph and replayed while the step
// signature (shapes, flags, buffer addresses) is unchanged — 10 launches + 2 copies become one graph launch.
static int step_with_graph_impl(vbmc_b200_ctx* c, const vbmc_b200_negelcbo_args* a, int Ns, int gmask, int nth, bool sync, bool hit,
                                bool trail) {
  const size_t nstage = static_cast<size_t>(c->ntheta) + 2;
  VB_TRY(c->theta_dev.reserve(sizeof(double) * nstage));
  VB_TRY(ensure_pinned(&c->theta_pinned, &c->theta_pinned_cap, sizeof(double) * nstage));
  memcpy(c->theta_pinned, a->theta, sizeof(double) * c->ntheta);
  uint64_t dyn[2] = {a->seed, a->stream};
  memcpy(c->theta_pinned + c->ntheta, dyn, sizeof(dyn));
  OutLayout ol;
  ol.init(nth, c->gp.S, c->K);
  VB_TRY(ensure_pinned(&c->out_pinned, &c->out_pinned_cap, sizeof(double) * ol.total));
  const bool philox = a->eps_mode == VBMC_B200_EPS_PHILOX;
  // generator mode: skip the generation when the buffer already holds this step's draws (generated ahead of time by the
  // previous step); generate the next step's draws ahead of time once the caller has been seen to advance the stream by one
  const bool lead = philox && !hit;
  auto body = [&]() -> int {
    VB_CUDA(cudaMemcpyAsync(c->theta_dev.p, c->theta_pinned, sizeof(double) * nstage, cudaMemcpyHostToDevice, c->stream));
    c->philox_dyn = philox;
    if (lead) VB_TRY(prepare_eps(c, Ns, VBMC_B200_EPS_PHILOX, nullptr, a->seed, a->stream));
    c->philox_seed = a->seed;
    c->philox_stream = a->stream;
    c->philox_trail = trail;
    const int rc = enqueue_step(c, Ns, gmask, a->use_thetabnd, 1, FIN_NEGELCBO, true);
    c->philox_dyn = false;
    c->philox_trail = false;
    VB_TRY(rc);
    VB_CUDA(cudaMemcpyAsync(c->out_pinned, c->out_dev.p, sizeof(double) * ol.total, cudaMemcpyDeviceToHost, c->stream));
    VB_TRY(join_trail(c));  // after the read-back has been enqueued: the copy does not wait for the generator
    return VBMC_B200_OK;
  };

  if (!philox) VB_TRY(prepare_eps(c, Ns, a->eps_mode, a->eps, a->seed, a->stream));  // host / resident draws: outside the graph
  const bool use_graph = c->graphs_enabled && !c->profiling && !step_uses_nccl(c);  // no NCCL node inside a captured step
  if (!use_graph) {
    VB_TRY(body());
    VB_CUDA(cudaStreamSynchronize(c->stream));
    return VBMC_B200_OK;
  }
  auto make_key = [&]() { return step_signature(c, Ns, gmask, a->use_thetabnd, (philox ? 1 : 0) | (lead ? 2 : 0) | (trail ? 4 : 0)); };
  const std::vector<long long> key = make_key();
  if (c->graph_exec && key == c->graph_key) {
    VB_CUDA(cudaGraphLaunch(c->graph_exec, c->stream));
    c->launches += c->graph_launches;
  } else if (key == c->warm_key) {
    // second call with this signature: every buffer is allocated, capture the sequence
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    const long long l0 = c->launches;
    cudaGraph_t graph = nullptr;
    VB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body();
    cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
    if (rc != VBMC_B200_OK || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      c->graphs_enabled = false;  // fall back to direct launches for the rest of the session
      VB_TRY(body());
      VB_CUDA(cudaStreamSynchronize(c->stream));
      return VBMC_B200_OK;
    }
    c->graph_launches = c->launches - l0;
    ce = cudaGraphInstantiate(&c->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
      c->graph_exec = nullptr;
      c->graphs_enabled = false;
      cudaGetLastError();
      VB_TRY(body());
      VB_CUDA(cudaStreamSynchronize(c->stream));
      return VBMC_B200_OK;
    }
    c->graph_key = key;
    VB_CUDA(cudaGraphLaunch(c->graph_exec, c->stream));
  } else {
    VB_TRY(body());   // first call with this signature: allocations happen here
    // the key must describe the buffers AFTER this call (they may have been allocated by it)
    c->warm_key = make_key();
  }
  if (philox) c->eps_f32 = c->precision == 32;  // a replayed graph does not pass through launch_philox
  if (sync) VB_CUDA(cudaStreamSynchronize(c->stream));
  return VBMC_B200_OK;
}

// a NaN objective on a multi-GPU step: tell a peer that never arrived (exchange_sum timed out) from a genuine NaN
static int check_exchange(vbmc_b200_ctx* c) {
  if (c->nranks <= 1 || !c->p2p_ready) return VBMC_B200_OK;
  unsigned long long err = 0;
  VB_CUDA(cudaMemcpy(&err, static_cast<unsigned long long*>(c->xchg.p) + XCHG_ERR, sizeof(err), cudaMemcpyDeviceToHost));
  if (err)
    VB_FAIL(VBMC_B200_ENCCL, "vbmc_b200:exchange: a peer rank did not deliver its partial sums in time (peer-memory all-reduce, VBMC_B200_P2P_TIMEOUT_S); re-create the communicator");
  return VBMC_B200_OK;
}

static int step_with_graph(vbmc_b200_ctx* c, const vbmc_b200_negelcbo_args* a, int Ns, int gmask, int nth, bool sync = true) {
  const bool philox = a->eps_mode == VBMC_B200_EPS_PHILOX;
  const bool hit = philox && eps_key_matches(c, a->seed, a->stream, Ns);
  // ahead-of-time generation only for a caller that is seen to stream (previous call had stream - 1): a repeated key — e.g. the
  // token draws of Ns == 0 calls — must not trigger it
  const bool trail = philox && c->prefetch_enabled && c->have_last_key && c->last_seed == a->seed && c->last_stream + 1 == a->stream;
  const int rc = step_with_graph_impl(c, a, Ns, gmask, nth, sync, hit, trail);
  if (rc != VBMC_B200_OK) {  // whatever was enqueued, the buffer is not trusted any more
    c->eps_key.valid = false;
    c->have_last_key = false;
    c->philox_pending = false;
    c->trail_join_pending = false;
    return rc;
  }
  if (philox) {
    c->have_last_key = true;
    c->last_seed = a->seed;
    c->last_stream = a->stream;
    eps_key_set(c, a->seed, a->stream + (trail ? 1 : 0), Ns);   // what the buffer holds now (a replayed graph bypasses prepare_eps)
  }
  if (sync && c->nranks > 1 && c->out_pinned && c->out_pinned[0] != c->out_pinned[0]) VB_TRY(check_exchange(c));
  return VBMC_B200_OK;
}

// Gradient of the (diagonal) log-joint variance, averaged over hyper-parameter samples
// (gplogjoint.m:289-303 per sample, :370-385 Jacobians, :407-410 average).  O(S K D) host work on small
// device read-backs; the N-long contractions were done on the device (run_variance with vgrad).
static int assemble_vargrad(vbmc_b200_ctx* c, int gmask, int jacobian, const std::vector<double>& vF, const std::vector<double>& J,
                            const std::vector<double>& vg, const double* Fs, std::vector<double>* dvarF) {
  const int D = c->D, K = c->K, S = c->gp.S, os = 2 + 2 * D;
  std::vector<double> vpc(static_cast<size_t>(D) * K + 3 * K + D + 2 * K + 2 * D), der(static_cast<size_t>(S) * (3 * D + 3)),
      go(static_cast<size_t>(S) * K * os);
  // the step filled only this rank's shard of the per-sample results: every rank needs all of them here
  if (c->nranks > 1) VB_TRY(launch_gplogjoint(c, 1, c->stream));
  VB_CUDA(cudaMemcpyAsync(vpc.data(), c->vp.mu, sizeof(double) * vpc.size(), cudaMemcpyDeviceToHost, c->stream));
  VB_CUDA(cudaMemcpyAsync(der.data(), c->gpDerived.p, sizeof(double) * der.size(), cudaMemcpyDeviceToHost, c->stream));
  VB_CUDA(cudaMemcpyAsync(go.data(), c->glj_out.p, sizeof(double) * go.size(), cudaMemcpyDeviceToHost, c->stream));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  const double* sigma = vpc.data() + D * K;
  const double* lambda = sigma + K;
  const double* w = lambda + D;
  const double* eta = w + K;
  const double* delta = eta + K + K + D;  // after lnsigma[K], lnlambda[D]
  const double* ell = der.data();
  const double* lnc = ell + S * D;
  const double* sn2eff = lnc + S + S + 2 * S * D;
  const double EPS = 2.220446049250313e-16;
  int n = 0, o_mu = 0, o_sig = 0, o_lam = 0, o_w = 0;
  o_mu = n; if (gmask & 1) n += D * K;
  o_sig = n; if (gmask & 2) n += K;
  o_lam = n; if (gmask & 4) n += D;
  o_w = n; if (gmask & 8) n += K;
  std::vector<double> wsm(K);
  {
    double es = 0.0;
    for (int k = 0; k < K; ++k) es += exp(eta[k]);
    for (int k = 0; k < K; ++k) wsm[k] = exp(eta[k]) / es;
  }
  auto jw = [&](std::vector<double>& g) {  // softmax Jacobian
    double dot = 0.0;
    for (int k = 0; k < K; ++k) dot += wsm[k] * g[k];
    for (int k = 0; k < K; ++k) g[k] = wsm[k] * (g[k] - dot);
  };
  std::vector<double> dv(static_cast<size_t>(n) * S, 0.0), dFs(static_cast<size_t>(n) * S, 0.0), gw(K);
  for (int s = 0; s < S; ++s) {
    double* dvs = dv.data() + static_cast<size_t>(s) * n;
    double* dfs = dFs.data() + static_cast<size_t>(s) * n;
    for (int k = 0; k < K; ++k) {
      const double* g = vg.data() + (static_cast<size_t>(s) * K + k) * os;   // raw contractions with R\V_k
      const double* f = go.data() + (static_cast<size_t>(s) * K + k) * os;   // [I, gsig, gmu, glam] of F(s)
      double slt = 0.0, s_l2t2 = 0.0;
      std::vector<double> t2(D);
      for (int d = 0; d < D; ++d) {
        t2[d] = 2.0 * sigma[k] * sigma[k] * lambda[d] * lambda[d] + ell[s * D + d] * ell[s * D + d] + 2.0 * delta[d] * delta[d];
        slt += 0.5 * log(t2[d]);
        s_l2t2 += lambda[d] * lambda[d] / t2[d];
      }
      const double nf_kk = exp(lnc[s] - slt);                                   // :275
      const double w2 = w[k] * w[k], ise = 1.0 / sn2eff[s];
      const double Jkk = J[(static_cast<size_t>(s) * K + k) * K + k];
      for (int d = 0; d < D; ++d) {
        if (gmask & 1) {
          dvs[o_mu + k * D + d] = -w2 * 2.0 * g[2 + d] * ise;                   // :290
          dfs[o_mu + k * D + d] = w[k] * f[2 + d];
        }
        if (gmask & 4) {
          dvs[o_lam + d] -= 2.0 * w2 * (sigma[k] * sigma[k] * nf_kk * lambda[d] / t2[d] + g[2 + D + d] * ise);  // :298
          dfs[o_lam + d] += w[k] * f[2 + D + d];
        }
      }
      if (gmask & 2) {
        dvs[o_sig + k] = -2.0 * w2 * (sigma[k] * nf_kk * s_l2t2 + g[1] * ise);  // :294
        dfs[o_sig + k] = w[k] * f[1];
        if (jacobian) { dvs[o_sig + k] *= sigma[k]; dfs[o_sig + k] *= sigma[k]; }
      }
      if (gmask & 8) {
        dvs[o_w + k] = 2.0 * w[k] * fmax(EPS, Jkk);                             // :302
        dfs[o_w + k] = f[0];
      }
    }
    if ((gmask & 4) && jacobian)
      for (int d = 0; d < D; ++d) { dvs[o_lam + d] *= lambda[d]; dfs[o_lam + d] *= lambda[d]; }
    if ((gmask & 8) && jacobian) {
      for (int k = 0; k < K; ++k) gw[k] = dvs[o_w + k];
      jw(gw);
      for (int k = 0; k < K; ++k) dvs[o_w + k] = gw[k];
      for (int k = 0; k < K; ++k) gw[k] = dfs[o_w + k];
      jw(gw);
      for (int k = 0; k < K; ++k) dfs[o_w + k] = gw[k];
    }
  }
  dvarF->assign(n, 0.0);
  if (S > 1) {  // dvv = 2*sum(F.*dF,2)/(Ns-1) - 2*Fbar.*sum(dF,2)/(Ns-1);  dvarF = sum(dvarF,2)/Ns + dvv  (:407-410)
    double Fbar = 0.0;
    for (int s = 0; s < S; ++s) Fbar += Fs[s];
    Fbar /= S;
    for (int i = 0; i < n; ++i) {
      double sv = 0.0, sfd = 0.0, sd = 0.0;
      for (int s = 0; s < S; ++s) {
        sv += dv[static_cast<size_t>(s) * n + i];
        sfd += Fs[s] * dFs[static_cast<size_t>(s) * n + i];
        sd += dFs[static_cast<size_t>(s) * n + i];
      }
      (*dvarF)[i] = sv / S + 2.0 * sfd / (S - 1) - 2.0 * Fbar * sd / (S - 1);
    }
  } else {
    for (int i = 0; i < n; ++i) (*dvarF)[i] = dv[i];
  }
  (void)vF;
  return VBMC_B200_OK;
}

int vbmc_b200_negelcbo(vbmc_b200_ctx* c, const vbmc_b200_negelcbo_args* a_in) {
  if (!c || !a_in) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  // Ns == 0: negelcbo_vbmc.m:102-109 replaces the Monte-Carlo entropy by the deterministic lower bound entlb_vbmc (what
  // vpsieve_vbmc.m:76 evaluates for every candidate).  The step runs WITHOUT the entropy sweep (no draws are generated, the
  // resident draw buffer and its bookkeeping are untouched): F0 = -G + penalties, dF0 likewise; then
  //     F = F0 - H_lb,   dF = dF0 - dH_lb        (F = -G - H + penalties, negelcbo_vbmc.m:116-117)
  const bool lower_bound = a_in->Ns == 0;
  vbmc_b200_negelcbo_args a_lb = *a_in;
  if (lower_bound) a_lb.Ns = 2;  // passes the Ns > 0 validation below; no draw is ever made
  const vbmc_b200_negelcbo_args* a = lower_bound ? &a_lb : a_in;
  double beta;
  int Ns, gmask;
  VB_TRY(negelcbo_validate(c, a, &beta, &Ns, &gmask));
  VB_CUDA(cudaSetDevice(c->device));
  const int nth = grad_mask_len(c, gmask);
  if (lower_bound) {
    const size_t nstage = static_cast<size_t>(c->ntheta) + 2;
    VB_TRY(c->theta_dev.reserve(sizeof(double) * nstage));
    VB_TRY(ensure_pinned(&c->theta_pinned, &c->theta_pinned_cap, sizeof(double) * nstage));
    memcpy(c->theta_pinned, a->theta, sizeof(double) * c->ntheta);
    memset(c->theta_pinned + c->ntheta, 0, 2 * sizeof(double));
    VB_CUDA(cudaMemcpyAsync(c->theta_dev.p, c->theta_pinned, sizeof(double) * nstage, cudaMemcpyHostToDevice, c->stream));
    VB_TRY(enqueue_step(c, Ns, gmask, a->use_thetabnd, 1, FIN_NEGELCBO_NOENT, true));
    VB_TRY(fetch_out(c, nth, c->gp.S));
    if (c->nranks > 1 && c->out_pinned[0] != c->out_pinned[0]) VB_TRY(check_exchange(c));
  } else {
    VB_TRY(step_with_graph(c, a, Ns, gmask, nth));
  }
  scatter_negelcbo(c, a, nth);
  if (lower_bound) {
    std::vector<double> dHlb(nth > 0 ? nth : 1);
    double Hlb = 0.0;
    VB_TRY(run_entlb(c, gmask, 1, &Hlb, nth ? dHlb.data() : nullptr));
    if (a->F) *a->F -= Hlb;
    if (a->H) *a->H = Hlb;
    for (int i = 0; i < nth; ++i) {
      if (a->dF) a->dF[i] -= dHlb[i];
      if (a->dH) a->dH[i] = dHlb[i];
    }
  }
  if (a->compute_var) {
    // varG (and J_sjk) from the factors; F = F + beta*sqrt(varF)   (negelcbo_vbmc.m:119-130)
    std::vector<double> vF, J, vg;
    const bool need_vgrad = a->compute_grad && beta != 0.0;  // then compute_var == 2 (validated above)
    VB_TRY(run_variance(c, a->compute_var, &vF, ((a->separate_K && a->J_sjk) || need_vgrad) ? &J : nullptr, need_vgrad ? &vg : nullptr));
    OutLayout ol;
    ol.init(nth, c->gp.S, c->K);
    double varG, varGss;
    combine_variance(c->out_pinned + ol.oFs, vF, c->gp.S, &varG, &varGss);
    if (need_vgrad && a->dF) {  // dF = dF + 0.5*beta*dvarG/sqrt(varF)   (negelcbo_vbmc.m:128-130)
      std::vector<double> dvar;
      VB_TRY(assemble_vargrad(c, gmask, 1, vF, J, vg, c->out_pinned + ol.oFs, &dvar));
      for (int i = 0; i < nth; ++i) a->dF[i] += 0.5 * beta * dvar[i] / sqrt(varG);
    }
    if (a->varG) *a->varG = varG;
    if (a->varGss) *a->varGss = varGss;
    if (a->varH) *a->varH = 0.0;
    if (a->varF) *a->varF = varG;
    if (beta != 0.0 && a->F) *a->F += beta * sqrt(varG);
    if (a->separate_K && a->J_sjk) scatter_J(J, c->gp.S, c->K, a->J_sjk);
  }
  return VBMC_B200_OK;
}

int vbmc_b200_negelcbo_resident_loop(vbmc_b200_ctx* c, const vbmc_b200_negelcbo_args* a, int steps, float* ms_total) {
  double beta;
  int Ns, gmask;
  VB_TRY(negelcbo_validate(c, a, &beta, &Ns, &gmask));
  if (steps <= 0 || !ms_total) VB_FAIL(VBMC_B200_EINVAL, "resident_loop: steps > 0 and ms_total required");
  if (a->compute_var) VB_FAIL(VBMC_B200_EINVAL, "resident_loop: compute_var must be 0");
  VB_CUDA(cudaSetDevice(c->device));
  vbmc_b200_negelcbo_args b = *a;
  if (a->eps_mode == VBMC_B200_EPS_HOST) {  // upload once, then reuse the resident draws
    VB_TRY(prepare_eps(c, Ns, VBMC_B200_EPS_HOST, a->eps, 0, 0));
    b.eps_mode = VBMC_B200_EPS_RESIDENT;
  }
  const int nth = grad_mask_len(c, gmask);
  VB_CUDA(cudaStreamSynchronize(c->stream));
  VB_CUDA(cudaEventRecord(c->ev_t0, c->stream));
  for (int i = 0; i < steps; ++i) {
    b.stream = a->stream + i;
    VB_TRY(step_with_graph(c, &b, Ns, gmask, nth, /*sync=*/i + 1 < steps));
  }
  VB_CUDA(cudaEventRecord(c->ev_t1, c->stream));
  VB_CUDA(cudaEventSynchronize(c->ev_t1));
  VB_CUDA(cudaEventElapsedTime(ms_total, c->ev_t0, c->ev_t1));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  scatter_negelcbo(c, a, nth);
  return VBMC_B200_OK;
}

// fminadam(@(theta_) negelcbo_vbmc(theta_, ...), x0, LB, UB, TolFun, MaxIter, master_stepsize) with the loop on the
// device (utils/fminadam.m:20-102; kernels in adam.cu).  Iteration 1 runs as direct launches (buffers get allocated),
// iteration 2 is captured into a CUDA graph, every later iteration is one graph launch; the host syncs only at the
// mini-batch ends where the reference tests for termination.
static int fminadam_impl(vbmc_b200_ctx* c, const vbmc_b200_fminadam_args* f);
int vbmc_b200_fminadam(vbmc_b200_ctx* c, const vbmc_b200_fminadam_args* f) {
  if (!c || !f) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  const int rc = fminadam_impl(c, f);
  if (rc != VBMC_B200_OK) {  // the draw buffer's bookkeeping is only trusted after a completed loop
    c->eps_key.valid = false;
    c->have_last_key = false;
    c->philox_pending = false;
    c->philox_trail = false;
    c->trail_join_pending = false;
  }
  return rc;
}
static int fminadam_impl(vbmc_b200_ctx* c, const vbmc_b200_fminadam_args* f) {
  if (!f->x0) VB_FAIL(VBMC_B200_EINVAL, "fminadam: x0 is required");
  vbmc_b200_negelcbo_args na;
  memset(&na, 0, sizeof(na));
  na.theta = f->x0; na.ntheta = f->nvars; na.beta = f->beta; na.Ns = f->Ns;
  na.compute_grad = 1; na.compute_var = f->compute_var; na.separate_K = 0; na.use_thetabnd = f->use_thetabnd;
  na.eps_mode = f->eps_mode; na.eps = f->eps; na.seed = f->seed; na.stream = f->stream;
  double beta;
  int Ns, gmask;
  VB_TRY(negelcbo_validate(c, &na, &beta, &Ns, &gmask));
  // beta ~= 0 (ELCBOWeight, negelcbo_vbmc.m:119-130): every iteration also runs the variance path (triangular solves + Gram on the
  // device, O(S K^2) assembly on the host) and folds beta*sqrt(varG) and its gradient into F, dF before the Adam update: the loop
  // is then driven from the host, one synchronisation per iteration, instead of replaying a graph.
  const bool penalised = beta != 0.0;
  const double TolFun = (f->TolFun > 0.0) ? f->TolFun : 0.001;        // fminadam.m:6 (NaN compares false)
  const int MaxIter = f->MaxIter > 0 ? f->MaxIter : 10000;            // :7
  const int B = 20;                                                   // batchsize (:24)
  if (MaxIter < B) VB_FAIL(VBMC_B200_EINVAL, "fminadam: MaxIter=%d < 20 (the reference indexes xtab(:,iter-19:iter))", MaxIter);
  VB_CUDA(cudaSetDevice(c->device));
  const int n = c->ntheta;
  const int nth = grad_mask_len(c, gmask);  // == n: every optimised block has a gradient
  const size_t nstage = static_cast<size_t>(n) + 2;
  VB_TRY(c->theta_dev.reserve(sizeof(double) * nstage));
  VB_TRY(ensure_pinned(&c->theta_pinned, &c->theta_pinned_cap, sizeof(double) * nstage));
  OutLayout ol;
  ol.init(nth, c->gp.S, c->K);
  VB_TRY(c->out_dev.reserve(sizeof(double) * ol.total));
  VB_TRY(ensure_pinned(&c->out_pinned, &c->out_pinned_cap, sizeof(double) * (ol.total > 16 ? ol.total : 16)));
  // state: m | v | lb | ub | xout | stats[8] | it (8 bytes) | ftab[MaxIter] | corr[1 + n]
  const size_t nstate = 5 * static_cast<size_t>(n) + 8 + 1 + MaxIter + 1 + n;
  VB_TRY(c->adamState.reserve(sizeof(double) * nstate));
  VB_TRY(c->adamXtab.reserve(sizeof(double) * static_cast<size_t>(n) * MaxIter));
  double* sb = c->adamState.d();
  AdamArgs aa;
  aa.n = n;
  aa.step_max = (f->stepsize_max > 0.0) ? f->stepsize_max : 0.1;      // :11-18
  aa.step_min = (f->stepsize_min > 0.0) ? f->stepsize_min : 0.001;
  aa.decay = (f->stepsize_decay > 0.0) ? f->stepsize_decay : 200.0;
  aa.x = c->theta_dev.d();
  aa.grad = c->out_dev.d() + ol.oDF;
  aa.fval = c->out_dev.d() + ol.oF;
  aa.m = sb; aa.v = sb + n;
  double* lb = sb + 2 * static_cast<size_t>(n);
  double* ub = sb + 3 * static_cast<size_t>(n);
  aa.lb = lb; aa.ub = ub;
  aa.xout = sb + 4 * static_cast<size_t>(n);
  aa.stats = sb + 5 * static_cast<size_t>(n);
  aa.it = reinterpret_cast<int*>(aa.stats + 8);
  aa.ftab = aa.stats + 9;
  aa.xtab = c->adamXtab.d();
  double* d_corr = aa.ftab + MaxIter;
  std::vector<double> corr_h(penalised ? 1 + static_cast<size_t>(n) : 0);
  const bool philox = f->eps_mode == VBMC_B200_EPS_PHILOX;
  aa.dyn = philox ? reinterpret_cast<unsigned long long*>(c->theta_dev.d() + n) : nullptr;

  // ---- initial state ----
  memcpy(c->theta_pinned, f->x0, sizeof(double) * n);
  uint64_t dyn[2] = {f->seed, f->stream};
  memcpy(c->theta_pinned + n, dyn, sizeof(dyn));
  VB_CUDA(cudaMemcpyAsync(c->theta_dev.p, c->theta_pinned, sizeof(double) * nstage, cudaMemcpyHostToDevice, c->stream));
  VB_CUDA(cudaMemsetAsync(sb, 0, sizeof(double) * (2 * static_cast<size_t>(n)), c->stream));          // m = v = 0 (:38)
  VB_CUDA(cudaMemsetAsync(aa.stats, 0, sizeof(double) * 9, c->stream));                               // stats, it = 0
  {
    std::vector<double> bnd(2 * static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) {
      bnd[i] = f->LB ? f->LB[i] : -INFINITY;                                                          // :35-36
      bnd[n + i] = f->UB ? f->UB[i] : INFINITY;
    }
    VB_CUDA(cudaMemcpyAsync(lb, bnd.data(), sizeof(double) * 2 * n, cudaMemcpyHostToDevice, c->stream));
    VB_CUDA(cudaStreamSynchronize(c->stream));
  }
  int eps_mode = f->eps_mode;
  if (eps_mode == VBMC_B200_EPS_HOST) {  // parity mode: the same draws every iteration, uploaded once
    VB_TRY(prepare_eps(c, Ns, VBMC_B200_EPS_HOST, f->eps, 0, 0));
    eps_mode = VBMC_B200_EPS_RESIDENT;
  }
  if (!philox) VB_TRY(prepare_eps(c, Ns, eps_mode, nullptr, 0, 0));

  // generator mode: iteration i + 1's draws are generated ahead of time in the tail of iteration i (the loop advances the stream
  // by one per iteration), so only iteration 1 may have to generate its own — unless an earlier call left them behind
  const bool hit0 = philox && eps_key_matches(c, f->seed, f->stream, Ns);
  const bool trail = philox && c->prefetch_enabled;
  c->eps_key.valid = false;  // set again when the loop has completed
  c->have_last_key = false;
  auto body = [&](bool lead) -> int {
    c->philox_dyn = philox;
    if (lead) VB_TRY(prepare_eps(c, Ns, VBMC_B200_EPS_PHILOX, nullptr, f->seed, f->stream));
    c->philox_seed = f->seed;
    c->philox_stream = f->stream;
    c->philox_trail = trail;
    const int rc = enqueue_step(c, Ns, gmask, f->use_thetabnd, 1, FIN_NEGELCBO, true);
    c->philox_dyn = false;
    c->philox_trail = false;
    VB_TRY(rc);
    if (penalised) {
      VB_TRY(fetch_out(c, nth, c->gp.S));   // Fs of this iterate (and a synchronisation: vp.cur is this iterate's)
      std::vector<double> vF, J, vg, dvar;
      VB_TRY(run_variance(c, na.compute_var, &vF, &J, &vg));
      double varG, varGss;
      combine_variance(c->out_pinned + ol.oFs, vF, c->gp.S, &varG, &varGss);
      VB_TRY(assemble_vargrad(c, gmask, 1, vF, J, vg, c->out_pinned + ol.oFs, &dvar));
      corr_h[0] = beta * sqrt(varG);                                                    // F = F + beta*sqrt(varF)   (:125)
      for (int i = 0; i < n; ++i) corr_h[1 + i] = 0.5 * beta * dvar[i] / sqrt(varG);    // dF = dF + 0.5*beta*dvarF/sqrt(varF)  (:128-130)
      VB_CUDA(cudaMemcpyAsync(d_corr, corr_h.data(), sizeof(double) * corr_h.size(), cudaMemcpyHostToDevice, c->stream));
      VB_TRY(launch_adam_penalty(c, aa, d_corr, c->stream));
    }
    VB_TRY(launch_adam_step(c, aa, c->stream));
    VB_TRY(join_trail(c));
    return VBMC_B200_OK;
  };
  const bool lead_first = philox && !hit0, lead_later = philox && !trail;
  auto make_key = [&]() {
    std::vector<long long> key = step_signature(c, Ns, gmask, f->use_thetabnd, (philox ? 1 : 0) | (lead_later ? 2 : 0) | (trail ? 4 : 0));
    long long bits[3];
    memcpy(&bits[0], &aa.step_max, 8); memcpy(&bits[1], &aa.step_min, 8); memcpy(&bits[2], &aa.decay, 8);
    key.insert(key.end(), bits, bits + 3);
    key.push_back(reinterpret_cast<long long>(c->adamState.p));
    key.push_back(reinterpret_cast<long long>(c->adamXtab.p));
    key.push_back(MaxIter);
    return key;
  };
  bool use_graph = c->graphs_enabled && !c->profiling && !step_uses_nccl(c) && !penalised;
  double* stats_h = c->out_pinned;  // the pinned output mirror doubles as the read-back slot of the termination test
  int it = 0;
  while (it < MaxIter) {
    ++it;
    if (!use_graph || it == 1) {
      VB_TRY(body(it == 1 ? lead_first : lead_later));
    } else {
      if (it == 2) {
        const std::vector<long long> key = make_key();
        if (!(c->adam_graph && key == c->adam_key)) {
          if (c->adam_graph) { cudaGraphExecDestroy(c->adam_graph); c->adam_graph = nullptr; }
          const long long l0 = c->launches;
          cudaGraph_t graph = nullptr;
          VB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
          const int rc = body(lead_later);
          cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
          if (rc == VBMC_B200_OK && ce == cudaSuccess && graph) {
            c->adam_graph_launches = c->launches - l0;
            c->launches = l0;
            ce = cudaGraphInstantiate(&c->adam_graph, graph, 0);
            if (ce != cudaSuccess) c->adam_graph = nullptr;
          }
          if (graph) cudaGraphDestroy(graph);
          if (!c->adam_graph) {  // capture unavailable: direct launches for the rest of this call
            cudaGetLastError();
            use_graph = false;
            VB_TRY(body(lead_later));
            continue;
          }
          c->adam_key = key;
        }
      }
      VB_CUDA(cudaGraphLaunch(c->adam_graph, c->stream));
      c->launches += c->adam_graph_launches;
    }
    if (it % B == 0 && it >= 2 * B) {  // isMinibatchEnd && iter >= MinIter (:65)
      VB_TRY(launch_adam_check(c, aa, it, TolFun, c->stream));
      VB_CUDA(cudaMemcpyAsync(stats_h, aa.stats, sizeof(double) * 8, cudaMemcpyDeviceToHost, c->stream));
      VB_CUDA(cudaStreamSynchronize(c->stream));
      if (stats_h[0] != 0.0) break;    // :80-82
    }
  }
  VB_TRY(launch_adam_final(c, aa, it, c->stream));
  VB_CUDA(cudaMemcpyAsync(stats_h, aa.stats, sizeof(double) * 8, cudaMemcpyDeviceToHost, c->stream));
  if (f->x) VB_CUDA(cudaMemcpyAsync(f->x, aa.xout, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  if (f->ftab) VB_CUDA(cudaMemcpyAsync(f->ftab, aa.ftab, sizeof(double) * it, cudaMemcpyDeviceToHost, c->stream));
  if (f->xtab)
    VB_CUDA(cudaMemcpyAsync(f->xtab, aa.xtab, sizeof(double) * static_cast<size_t>(n) * it, cudaMemcpyDeviceToHost, c->stream));
  VB_CUDA(cudaStreamSynchronize(c->stream));
  if (f->f) *f->f = stats_h[5];
  if (f->iter) *f->iter = it;
  if (f->stats)
    for (int i = 0; i < 5; ++i) f->stats[i] = stats_h[i];
  if (philox) {
    c->eps_f32 = c->precision == 32;
    // the buffer holds the draws of stream + it (generated ahead of time) or of the last iteration, stream + it - 1
    eps_key_set(c, f->seed, f->stream + static_cast<uint64_t>(it) - (trail ? 0 : 1), Ns);
    c->have_last_key = true;
    c->last_seed = f->seed;
    c->last_stream = f->stream + static_cast<uint64_t>(it) - 1;
  }
  return VBMC_B200_OK;
}

int vbmc_b200_entmc(vbmc_b200_ctx* c, int Ns, const int grad_flags[4], int jacobian_flag, int eps_mode, const double* eps,
                    uint64_t seed, uint64_t stream, double* H, double* dH) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  if (!c->vp_ready) VB_FAIL(VBMC_B200_ESTATE, "entmc: call vbmc_b200_vp_set first");
  if (Ns <= 0) VB_FAIL(VBMC_B200_EINVAL, "entmc: Ns must be positive");
  VB_CUDA(cudaSetDevice(c->device));
  Ns = (Ns + 1) / 2 * 2;
  int gmask = 0;
  if (grad_flags && dH)
    for (int i = 0; i < 4; ++i)
      if (grad_flags[i]) gmask |= 1 << i;
  VB_TRY(prepare_eps(c, Ns, eps_mode, eps, seed, stream));
  VB_TRY(enqueue_step(c, Ns, gmask, 0, jacobian_flag, FIN_ENTMC, false));
  const int nth = grad_mask_len(c, gmask);
  VB_TRY(fetch_out(c, nth, 0));
  OutLayout ol;
  ol.init(nth, 0, c->K);
  if (H) *H = c->out_pinned[ol.oH];
  if (dH && nth) memcpy(dH, c->out_pinned + ol.oDH, sizeof(double) * nth);
  return VBMC_B200_OK;
}

int vbmc_b200_gplogjoint(vbmc_b200_ctx* c, const int grad_flags[4], int avg_flag, int jacobian_flag, int compute_var,
                         double* F, double* dF, double* varF, double* dvarF, double* varss, double* I_sk, double* J_sjk) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  if (!c->vp_ready) VB_FAIL(VBMC_B200_ESTATE, "gplogjoint: call vbmc_b200_vp_set first");
  if (!c->gp_ready) VB_FAIL(VBMC_B200_ESTATE, "gplogjoint: call vbmc_b200_gp_attach or vbmc_b200_gp_post first");
  if (c->gp.D != c->D) VB_FAIL(VBMC_B200_EINVAL, "gplogjoint: vp.D=%d but gp.D=%d", c->D, c->gp.D);
  if (dvarF && compute_var && compute_var != 2)
    VB_FAIL(VBMC_B200_EREFERENCE,
            "gplogjoint:FullVarianceGradient: Computation of gradient of log joint variance is currently available only "
            "for diagonal approximation of the variance.");
  if (compute_var != 0 && compute_var != 1 && compute_var != 2) VB_FAIL(VBMC_B200_EINVAL, "gplogjoint: compute_var must be 0, 1 or 2");
  if (!avg_flag && c->gp.S > 1)
    VB_FAIL(VBMC_B200_EUNSUPPORTED, "vbmc_b200:OutOfScope: avg_flag=0 with several hyper-parameter samples");
  VB_CUDA(cudaSetDevice(c->device));
  int gmask = 0;
  if (grad_flags && dF)
    for (int i = 0; i < 4; ++i)
      if (grad_flags[i]) gmask |= 1 << i;
  VB_TRY(enqueue_step(c, 2, gmask, 0, jacobian_flag, FIN_GPLOGJOINT, false));
  const int nth = grad_mask_len(c, gmask);
  VB_TRY(fetch_out(c, nth, c->gp.S));
  OutLayout ol;
  ol.init(nth, c->gp.S, c->K);
  const double* o = c->out_pinned;
  if (F) *F = o[ol.oG];
  if (dF && nth) memcpy(dF, o + ol.oDG, sizeof(double) * nth);
  if (varss) *varss = 0.0;
  if (I_sk) {
    const int S = c->gp.S, K = c->K;
    for (int s = 0; s < S; ++s)
      for (int k = 0; k < K; ++k) I_sk[s + static_cast<size_t>(k) * S] = o[ol.oIsk + s * K + k];
  }
  if (compute_var) {
    std::vector<double> vF, J, vg;
    const bool need_vgrad = dvarF != nullptr && gmask != 0;
    VB_TRY(run_variance(c, compute_var, &vF, (J_sjk || need_vgrad) ? &J : nullptr, need_vgrad ? &vg : nullptr));
    double vG, vss;
    combine_variance(c->out_pinned + ol.oFs, vF, c->gp.S, &vG, &vss);
    if (need_vgrad) {
      std::vector<double> dvar;
      VB_TRY(assemble_vargrad(c, gmask, jacobian_flag, vF, J, vg, c->out_pinned + ol.oFs, &dvar));
      memcpy(dvarF, dvar.data(), sizeof(double) * nth);
    }
    if (varF) *varF = vG;
    if (varss) *varss = vss;
    if (J_sjk) scatter_J(J, c->gp.S, c->K, J_sjk);
  }
  return VBMC_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// profiling hooks
// ---------------------------------------------------------------------------------------------
int vbmc_b200_profile_enable(vbmc_b200_ctx* c, int on) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  VB_CUDA(cudaSetDevice(c->device));
  if (!on && c->profiling) VB_TRY(profile_collect(c));
  c->profiling = on != 0;
  return VBMC_B200_OK;
}
int vbmc_b200_profile_get(vbmc_b200_ctx* c, const char* name, double* ms_sum, long long* launches) {
  if (!c || !name) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  VB_CUDA(cudaSetDevice(c->device));
  VB_TRY(profile_collect(c));
  auto it = c->prof.find(name);
  if (ms_sum) *ms_sum = it == c->prof.end() ? 0.0 : it->second.ms;
  if (launches) *launches = it == c->prof.end() ? 0 : it->second.n;
  return VBMC_B200_OK;
}
int vbmc_b200_profile_reset(vbmc_b200_ctx* c) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  VB_CUDA(cudaSetDevice(c->device));
  VB_TRY(profile_collect(c));
  c->prof.clear();
  return VBMC_B200_OK;
}

}  // extern "C"
