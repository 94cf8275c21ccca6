// Multi-GPU exchange (one NCCL all-reduce of the partial-sum vector R per step)
// and measurement helpers (FP64 FMA peak, HBM copy bandwidth, L2 flush).
#include <dlfcn.h>

#include "common.cuh"

namespace vb {

// NCCL is loaded lazily (dlopen) so that single-GPU use has no NCCL dependency.  Inside a
// torch process the already-loaded libnccl.so.2 (same SONAME) is the one that resolves.
typedef int (*fn_getuid)(void*);
typedef int (*fn_initrank)(ncclComm**, int, /*ncclUniqueId by value*/ struct UidBlob, int);
struct UidBlob {
  char b[128];
};
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, ncclComm*, cudaStream_t);
typedef int (*fn_destroy)(ncclComm*);
typedef const char* (*fn_errstr)(int);

static struct {
  void* h = nullptr;
  fn_getuid getuid = nullptr;
  fn_initrank initrank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
} g_nccl;

static int nccl_load() {
  if (g_nccl.h) return VBMC_B200_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) VB_FAIL(VBMC_B200_ENCCL, "vbmc_b200:nccl: cannot dlopen libnccl.so.2 (%s)", dlerror());
  g_nccl.getuid = reinterpret_cast<fn_getuid>(dlsym(g_nccl.h, "ncclGetUniqueId"));
  g_nccl.initrank = reinterpret_cast<fn_initrank>(dlsym(g_nccl.h, "ncclCommInitRank"));
  g_nccl.allreduce = reinterpret_cast<fn_allreduce>(dlsym(g_nccl.h, "ncclAllReduce"));
  g_nccl.destroy = reinterpret_cast<fn_destroy>(dlsym(g_nccl.h, "ncclCommDestroy"));
  g_nccl.errstr = reinterpret_cast<fn_errstr>(dlsym(g_nccl.h, "ncclGetErrorString"));
  if (!g_nccl.getuid || !g_nccl.initrank || !g_nccl.allreduce || !g_nccl.destroy)
    VB_FAIL(VBMC_B200_ENCCL, "vbmc_b200:nccl: libnccl lacks a required symbol");
  return VBMC_B200_OK;
}

#define VB_NCCL(expr)                                                                                  \
  do {                                                                                                 \
    int _r = (expr);                                                                                   \
    if (_r != 0) {                                                                                     \
      vb::set_error("vbmc_b200:nccl: %s failed: %s", #expr, g_nccl.errstr ? g_nccl.errstr(_r) : "?"); \
      return VBMC_B200_ENCCL;                                                                          \
    }                                                                                                  \
  } while (0)

int allreduce_R(vbmc_b200_ctx* c, int count, cudaStream_t st) {
  if (c->nranks <= 1) return VBMC_B200_OK;
  if (!c->comm) VB_FAIL(VBMC_B200_ESTATE, "allreduce: communicator not initialised");
  c->launches++;  // NCCL's kernel
  // ncclDouble = 8, ncclSum = 0
  VB_NCCL(g_nccl.allreduce(c->R_dev.p, c->R_dev.p, static_cast<size_t>(count), 8, 0, c->comm, st));
  return VBMC_B200_OK;
}

// ---- peer-memory exchange set-up (see XchgDev in common.cuh) ----
static void p2p_teardown(vbmc_b200_ctx* c) {
  for (int r = 0; r < XCHG_MAXR; ++r) {
    if (c->xchg_mapped[r]) cudaIpcCloseMemHandle(c->xchg_mapped[r]);
    c->xchg_mapped[r] = nullptr;
  }
  c->xchg.release();
  c->xchg_peers.release();
  c->p2p_ready = false;
  c->xdev = XchgDev{};
  cudaGetLastError();
}

// gather `words` int32 per rank through the communicator (sum of disjoint slots); used for the IPC handles and to agree on success
static int gather_words(vbmc_b200_ctx* c, const int* mine, int words, std::vector<int>* all) {
  const int n = c->nranks * words;
  DevBuf d;
  VB_TRY(d.reserve(sizeof(int) * n));
  std::vector<int> h(n, 0);
  memcpy(h.data() + c->rank * words, mine, sizeof(int) * words);
  cudaError_t e = cudaMemcpyAsync(d.p, h.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream);
  int rc = 0;
  if (e == cudaSuccess) rc = g_nccl.allreduce(d.p, d.p, static_cast<size_t>(n), /*ncclInt32*/ 2, /*ncclSum*/ 0, c->comm, c->stream);
  if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(h.data(), d.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(c->stream);
  d.release();
  if (rc != 0) VB_FAIL(VBMC_B200_ENCCL, "vbmc_b200:nccl: handle exchange failed: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
  if (e != cudaSuccess) VB_FAIL(VBMC_B200_ECUDA, "CUDA error during the handle exchange: %s", cudaGetErrorString(e));
  *all = h;
  return VBMC_B200_OK;
}

// Every rank allocates its exchange buffer, publishes its IPC handle and maps the others'.  Ranks agree (second gather) on
// whether ALL mappings succeeded; if not, every rank keeps the NCCL all-reduce.  VBMC_B200_P2P=0 skips the attempt.
static int p2p_setup(vbmc_b200_ctx* c) {
  if (const char* e = getenv("VBMC_B200_P2P"))
    if (!strcmp(e, "0")) return VBMC_B200_OK;
  if (c->nranks > XCHG_MAXR) return VBMC_B200_OK;
  const int cap = 40960;  // doubles per slot: covers R for K <= 128, D <= 24, S <= 100 (larger steps use NCCL)
  const size_t bytes = sizeof(double) * (XCHG_HDR + 2 * static_cast<size_t>(c->nranks) * cap);
  int ok = 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (c->xchg.reserve(bytes) != VBMC_B200_OK) ok = 0;
  if (ok && cudaMemset(c->xchg.p, 0, bytes) != cudaSuccess) ok = 0;
  if (ok && cudaDeviceSynchronize() != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, c->xchg.p) != cudaSuccess) ok = 0;
  cudaGetLastError();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  int words[17];
  memcpy(words, &mine, 64);
  words[16] = ok;
  std::vector<int> all;
  VB_TRY(gather_words(c, words, 17, &all));
  std::vector<unsigned long long*> peers(c->nranks, nullptr);
  for (int r = 0; r < c->nranks && ok; ++r) {
    if (!all[r * 17 + 16]) { ok = 0; break; }
    if (r == c->rank) { peers[r] = static_cast<unsigned long long*>(c->xchg.p); continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, &all[r * 17], 64);
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    c->xchg_mapped[r] = ptr;
    peers[r] = static_cast<unsigned long long*>(ptr);
  }
  if (ok && c->xchg_peers.reserve(sizeof(void*) * XCHG_MAXR) != VBMC_B200_OK) ok = 0;
  if (ok && cudaMemcpy(c->xchg_peers.p, peers.data(), sizeof(void*) * c->nranks, cudaMemcpyHostToDevice) != cudaSuccess) ok = 0;
  int okw = ok;
  VB_TRY(gather_words(c, &okw, 1, &all));  // also a barrier: nobody pushes before every buffer is zeroed and mapped
  for (int r = 0; r < c->nranks; ++r) ok = ok && all[r];
  if (!ok) {
    p2p_teardown(c);
    return VBMC_B200_OK;
  }
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
  c->xdev.nranks = c->nranks;
  c->xdev.rank = c->rank;
  c->xdev.cap = cap;
  c->xdev.peer = reinterpret_cast<unsigned long long* const*>(c->xchg_peers.p);
  // How long a rank waits inside finalize_kernel for its peers' partial sums.  Ranks are driven by independent host
  // processes, so a peer can legitimately be seconds behind (NCCL itself waits for ever); the bound only exists so that a
  // dead peer surfaces as vbmc_b200:exchange instead of a kernel that never ends.  VBMC_B200_P2P_TIMEOUT_S, default 120.
  double tmo_s = 120.0;
  if (const char* e = getenv("VBMC_B200_P2P_TIMEOUT_S")) tmo_s = atof(e) > 0.0 ? atof(e) : tmo_s;
  c->xdev.timeout_cycles = static_cast<long long>(static_cast<double>(khz > 0 ? khz : 2000000) * 1000.0 * tmo_s);
  c->p2p_ready = true;
  return VBMC_B200_OK;
}

void comm_destroy(vbmc_b200_ctx* c) {
  p2p_teardown(c);
  if (c->comm && g_nccl.destroy) g_nccl.destroy(c->comm);
  c->comm = nullptr;
  c->nranks = 1;
  c->rank = 0;
  c->eps_key.valid = false;  // the shard of the pair axis this rank generates changes with the communicator
}

// ---- FP64 FMA peak: 8 independent DFMA chains per thread ----
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace vb

using namespace vb;

extern "C" {

int vbmc_b200_comm_unique_id(void* id128) {
  if (!id128) VB_FAIL(VBMC_B200_EINVAL, "null id buffer");
  VB_TRY(nccl_load());
  VB_NCCL(g_nccl.getuid(id128));
  return VBMC_B200_OK;
}

int vbmc_b200_comm_init(vbmc_b200_ctx* c, int nranks, int rank, const void* id128) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  if (nranks < 1 || rank < 0 || rank >= nranks) VB_FAIL(VBMC_B200_EINVAL, "comm_init: bad nranks/rank %d/%d", nranks, rank);
  VB_CUDA(cudaSetDevice(c->device));
  comm_destroy(c);
  if (nranks == 1) return VBMC_B200_OK;
  if (!id128) VB_FAIL(VBMC_B200_EINVAL, "comm_init: unique id required for nranks > 1");
  VB_TRY(nccl_load());
  UidBlob uid;
  memcpy(uid.b, id128, 128);
  VB_NCCL(g_nccl.initrank(&c->comm, nranks, uid, rank));
  c->nranks = nranks;
  c->rank = rank;
  c->eps_key.valid = false;
  VB_TRY(p2p_setup(c));
  return VBMC_B200_OK;
}

int vbmc_b200_shard_range(int total, int nranks, int rank, int* begin, int* end) {
  if (total < 0 || nranks < 1 || rank < 0 || rank >= nranks || !begin || !end) VB_FAIL(VBMC_B200_EINVAL, "shard_range: bad arguments");
  shard_range(total, nranks, rank, begin, end);
  return VBMC_B200_OK;
}

int vbmc_b200_comm_p2p(vbmc_b200_ctx* c, int* peer_memory) {
  if (!c || !peer_memory) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  *peer_memory = (c->nranks > 1 && c->p2p_ready) ? 1 : 0;
  return VBMC_B200_OK;
}

int vbmc_b200_comm_info(vbmc_b200_ctx* c, int* nranks, int* rank) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  if (nranks) *nranks = c->nranks;
  if (rank) *rank = c->rank;
  return VBMC_B200_OK;
}

int vbmc_b200_measure_fp64_peak(vbmc_b200_ctx* c, double* tflops) {
  if (!c || !tflops) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  VB_CUDA(cudaSetDevice(c->device));
  const int blocks = c->num_sms * 8, threads = 256, iters = 1 << 15;
  DevBuf buf;
  VB_TRY(buf.reserve(sizeof(double) * blocks * threads));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    VB_CUDA(cudaEventRecord(c->ev_t0, c->stream));
    {
      KernelScope ks(c, "dfma_peak", c->stream);
      dfma_peak_kernel<<<blocks, threads, 0, c->stream>>>(buf.d(), iters, 0.999999, 1e-9);
    }
    VB_CUDA(cudaEventRecord(c->ev_t1, c->stream));
    VB_CUDA(cudaEventSynchronize(c->ev_t1));
    float ms = 0.f;
    VB_CUDA(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
    const double tf = 2.0 * 8.0 * iters * static_cast<double>(blocks) * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  buf.release();
  *tflops = best;
  return VBMC_B200_OK;
}

int vbmc_b200_measure_hbm_copy(vbmc_b200_ctx* c, double* gbs) {
  if (!c || !gbs) VB_FAIL(VBMC_B200_EINVAL, "null argument");
  VB_CUDA(cudaSetDevice(c->device));
  const size_t bytes = 1ull << 30;
  DevBuf a, b;
  VB_TRY(a.reserve(bytes));
  VB_TRY(b.reserve(bytes));
  VB_CUDA(cudaMemsetAsync(a.p, 1, bytes, c->stream));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    VB_CUDA(cudaEventRecord(c->ev_t0, c->stream));
    VB_CUDA(cudaMemcpyAsync(b.p, a.p, bytes, cudaMemcpyDeviceToDevice, c->stream));
    VB_CUDA(cudaEventRecord(c->ev_t1, c->stream));
    VB_CUDA(cudaEventSynchronize(c->ev_t1));
    float ms = 0.f;
    VB_CUDA(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
    const double g = 2.0 * bytes / (ms * 1e-3) / 1e9;
    if (rep > 0 && g > best) best = g;
  }
  a.release();
  b.release();
  *gbs = best;
  return VBMC_B200_OK;
}

int vbmc_b200_flush_l2(vbmc_b200_ctx* c) {
  if (!c) VB_FAIL(VBMC_B200_EINVAL, "null context");
  VB_CUDA(cudaSetDevice(c->device));
  const size_t bytes = 256ull << 20;  // > 126 MB L2
  VB_TRY(c->flush.reserve(bytes));
  VB_CUDA(cudaMemsetAsync(c->flush.p, 0, bytes, c->stream));
  return VBMC_B200_OK;
}

}  // extern "C"
