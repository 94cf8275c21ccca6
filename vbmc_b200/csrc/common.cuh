// Shared declarations of libvbmc_b200 (sm_100a only).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/vbmc_b200.h"

namespace vb {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define VB_FAIL(code, ...)      \
  do {                          \
    vb::set_error(__VA_ARGS__); \
    return (code);              \
  } while (0)
#define VB_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      vb::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                    cudaGetErrorString(_e));                                               \
      return VBMC_B200_ECUDA;                                                              \
    }                                                                                      \
  } while (0)
#define VB_TRY(expr)                     \
  do {                                   \
    int _rc = (expr);                    \
    if (_rc != VBMC_B200_OK) return _rc; \
  } while (0)

// ------------------------------------------------------------------------------------------------
// growable device buffer of doubles / bytes
// ------------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return VBMC_B200_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      return VBMC_B200_ECUDA;
    }
    cap = bytes;
    return VBMC_B200_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  double* d() const { return static_cast<double*>(p); }
  int* i() const { return static_cast<int*>(p); }
};

// ------------------------------------------------------------------------------------------------
// device-side views passed to kernels
// ------------------------------------------------------------------------------------------------
struct VpDev {  // unpacked variational posterior (filled by vp_unpack_kernel each step)
  int D, K;
  double* mu;        // [K][D]  (== MATLAB D x K column-major)
  double* sigma;     // [K]
  double* lambda;    // [D]
  double* w;         // [K]
  double* eta;       // [K]
  double* lnsigma;   // [K]
  double* lnlambda;  // [D]
  double* delta;     // [D] (zeros when vp.delta empty)
  double* ck;        // [K]  w_k*nf/sigma_k^D      (entmc_vbmc.m:40,63)
  double* ak;        // [K]  ck_k/sigma_k
  double* cn;        // [K+1] nf/sigma_k^D, cn[K] = nf
  double* scratch;   // [K*D] work array of vp_unpack_kernel
  double* cblob;     // [K2*DP | K2 | K2 | DP] centred mu, ck, ak/sigma, 1/lambda: source of entmc's constant-bank tables
  int* form_flag;    // entmc formulation of this step: 0 expanded, 1 direct (set by vp_unpack_kernel)
  unsigned long long* dyn_snap;  // [2] {seed, stream} of the step being evaluated (copied by vp_unpack_kernel from the theta staging
                                 // buffer): the ahead-of-time draw generator reads it while adam_step_kernel may already advance the original
};

struct GpDev {  // attached GP posterior
  int N, D, S, Nhyp, Ncov, Nnoise, Nmean, meanfun;
  const double* X;      // [D][N] (N x D column-major)
  const double* hyp;    // [S][Nhyp]
  const double* alpha;  // [S][N]
  const double* ell;    // [S][D]  exp(hyp(1:D))
  const double* lnc;    // [S]     2*hyp(D+1) + sum(hyp(1:D))
  const double* m0;     // [S]
  const double* xm;     // [S][D]
  const double* iom2;   // [S][D]  1/omega^2
  const double* sn2eff; // [S]     1/sW(1)^2
};

// partial-sum vector R that is all-reduced across ranks.  Compact: the K x K matrix of column sums W_jl and
// the K x D lambda-gradient pieces are contracted with the weights BEFORE the exchange (every rank knows w), so the exchanged
// vector is O(KD + SK) -- 2660 doubles at c3 (it was 5110), 10 320 at c5 (it was 22 320).
struct RLayout {
  int D, K, S;
  int oHs, oM, oE, oWc, oI, oGmu, oGsig, oGlam, total;
  int oWfull;     // scratch behind the exchanged part: un-contracted W_jl [K][K] (first-generation / FP32 sweeps only)
  int total_all;
  __host__ __device__ void init(int D_, int K_, int S_) {
    D = D_; K = K_; S = S_;
    oHs = 0;
    oM = oHs + K;
    oE = oM + K * D;
    oWc = oE + K * D;
    oI = oWc + K;        // Wc[l] = sum_j w_j W_jl
    oGmu = oI + S * K;
    oGsig = oGmu + K * D;
    oGlam = oGsig + K;
    total = oGlam + D;   // Glam[d] = sum_k w_k sum_s glam[s][k][d]
    oWfull = total;
    total_all = total + K * K;
  }
};

// output block written by finalize_kernel (device), copied to the host in one memcpy
struct OutLayout {
  int ntheta, S, K;
  int oF, oG, oH, oVarF, oVarGss, oVarG, oVarH, oDF, oDH, oDG, oIsk, oFs, total;
  __host__ __device__ void init(int ntheta_, int S_, int K_) {
    ntheta = ntheta_; S = S_; K = K_;
    oF = 0; oG = 1; oH = 2; oVarF = 3; oVarGss = 4; oVarG = 5; oVarH = 6;
    oDF = 8;
    oDH = oDF + ntheta;
    oDG = oDH + ntheta;
    oIsk = oDG + ntheta;
    oFs = oIsk + S * K;   // F(s) = sum_k w_k I_sk per hyper-parameter sample (gplogjoint.m:203)
    total = oFs + S;
  }
};

// device-resident fminadam state (adam.cu; utils/fminadam.m:20-102)
struct AdamArgs {
  int n;               // nvars
  double step_max, step_min, decay;
  double* x;           // [n] current iterate == the theta staging buffer the step kernels read
  const double* grad;  // [n] dF of the step that just ran (output block)
  const double* fval;  // F of that step
  double *m, *v;       // [n] Adam moments
  const double *lb, *ub;  // [n]
  double* xtab;        // [MaxIter][n] iterate history (MATLAB xtab(:,iter))
  double* ftab;        // [MaxIter]
  double* xout;        // [n] mean of the last 20 iterates
  double* stats;       // [8] stop flag, dx, slope, slope_err, slope_err_max, f
  int* it;             // iterations done so far
  unsigned long long* dyn;  // {seed, stream} of the device draw generator (stream advanced every iteration) or NULL
};

// One-shot all-reduce of R over NVLink peer memory, fused into finalize_kernel (misc.cu sets it up, finalize.cu uses it).
// Every rank owns one exchange buffer and maps every peer's through CUDA IPC:
//   words [0, 32)   : arrival flags  [parity][rank]   (sequence number of the step whose partial sums have landed)
//   word  32        : sequence number of the last completed exchange (same on every rank)
//   word  33        : error flag (a peer did not arrive in time)
//   words [64, ...) : inbox [parity][rank][cap] doubles
// A step pushes its R into slot [seq & 1][my rank] of every rank's inbox, raises its flag there, waits for all flags in
// its own buffer, and sums the nranks slots in rank order (=> bit-identical results on all ranks).
enum { XCHG_MAXR = 16, XCHG_SEQ = 32, XCHG_ERR = 33, XCHG_HDR = 64 };
struct XchgDev {
  int nranks = 1, rank = 0;
  int cap = 0;                         // doubles per (parity, rank) slot
  unsigned long long* const* peer = nullptr;  // device array [nranks]: base of every rank's exchange buffer
  long long timeout_cycles = 0;        // give up waiting for a peer after this many SM cycles
};

#ifdef __CUDACC__
// Producer side of the peer-memory all-reduce (finalize.cu): a reduction kernel writes each value of R it produces straight
// into slot [parity][my rank] of EVERY rank's inbox, so the single-CTA finalize_kernel only has to publish the flags, wait and
// sum -- the 8 x 21 KB push is spread over all the threads that produce R instead of one CTA.  No-op when xc.peer is null.
__device__ __forceinline__ void xchg_push(const XchgDev& xc, int idx, double v) {
  if (xc.peer == nullptr) return;
  const unsigned long long* me = xc.peer[xc.rank];
  const unsigned long long seq = me[XCHG_SEQ] + 1;   // finalize_kernel of the previous step stored it (stream order)
  const size_t at = (static_cast<size_t>(seq & 1) * xc.nranks + xc.rank) * xc.cap + idx;
  for (int r = 0; r < xc.nranks; ++r) reinterpret_cast<double*>(xc.peer[r] + XCHG_HDR)[at] = v;
}
#endif

struct Prof {
  double ms = 0;
  long long n = 0;
};

}  // namespace vb

// ------------------------------------------------------------------------------------------------
// the context
// ------------------------------------------------------------------------------------------------
struct ncclComm;
struct vbmc_b200_ctx {
  int device = 0;
  int num_sms = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // gplogjoint branch
  cudaStream_t stream3 = nullptr;  // draw generation (independent of theta)
  cudaEvent_t ev_fork0 = nullptr, ev_philox = nullptr;
  cudaEvent_t ev_glj = nullptr;  // log-joint kernel done (before its reduction)
  cudaEvent_t ev_la_main = nullptr, ev_la_side = nullptr;  // refit look-ahead: panel stream <-> trailing-update stream (gp_refit.cu)
  cudaEvent_t ev_trail_fork = nullptr, ev_trail = nullptr;  // ahead-of-time draw generation (forked after the entropy sweep)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  long long launches = 0;
  int precision = 64;  // 64: everything FP64; 32: the entropy sweep in FP32 (vbmc_b200_set_precision)
  double entmc_prune_c = 50.0;  // entmc: skip components that contribute < exp(-c) of q to a whole warp (VBMC_B200_ENTMC_PRUNE, 0 = off)
  bool entmc_prune_stats_on = false;
  vb::DevBuf entmc_prune_stats;  // {kept, total} counters (bench)
  int entmc_form = -1;  // -1 auto (device guard), 0 force expanded, 1 force direct (VBMC_B200_ENTMC_FORM)

  // multi-GPU
  int nranks = 1, rank = 0;
  ncclComm* comm = nullptr;
  // peer-memory exchange (fused all-reduce): ready when every rank mapped every other rank's buffer
  bool p2p_ready = false;
  vb::DevBuf xchg, xchg_peers;
  void* xchg_mapped[16] = {nullptr};   // cudaIpcOpenMemHandle results (own slot unused)
  vb::XchgDev xdev{};

  // GP
  bool gp_ready = false;
  unsigned long long gp_tag = 0;  // opaque caller fingerprint of the resident posterior (vbmc_b200_gp_tag_*); 0 = none
  vb::GpDev gp{};
  vb::DevBuf gpX, gpHyp, gpAlpha, gpDerived, gpL, gpY, gpS2, gpWork;
  vb::DevBuf gpFlags;      // [S][Np/64] hand-over flags of the multi-CTA back substitution (gp_bsolve3_kernel)
  int bsolve_epoch = 0;    // value a raised flag carries in the current launch
  vb::DevBuf trsvFlags;    // [S][ld/64 + 1] flags of the single-column sweeps (trsv1_kernel) + one error word at the end
  bool trsv1_pending = false;  // a single-column sweep ran since the last trsv1_check
  std::vector<int> gpLchol;
  std::vector<int> gpLfactor;  // per sample: device L is a Cholesky factor (1) or -inv(K + diag) handed over by the host (0)
  std::vector<double> gpSn2mult;
  std::vector<double> gpSn2effHost;  // 1/sW(1)^2 per sample (host copy of gp.sn2eff)
  std::vector<double> gpHypHost;  // host copy of gp.post(s).hyp ([S][Nhyp]) for O(S) host epilogues
  int gp_noisefun[3] = {1, 0, 0};
  vb::DevBuf predWork;  // gplite_pred: test points, cross-kernel columns, results
  vb::DevBuf trsmWork;  // blocked forward substitution: padded right-hand sides (trsm.cu)
  cudaGraphExec_t trsm_graph[2] = {nullptr, nullptr};   // forward / backward sweep of trsm.cu
  std::vector<long long> trsm_key[2];
  vb::DevBuf gpXalt, gpAlphaAlt;   // ping-pong partners of gpX / gpAlpha for the rank-one update (no allocation per update)
  bool gpHasL = false;
  int gpLd = 0;  // leading dimension of the factors in gpL (N when attached from the host, Np after gp_post)

  // VP
  bool vp_ready = false;
  int D = 0, K = 0;
  int opt[4] = {0, 0, 0, 0};
  int ntheta = 0;
  int vp_cblob_len = 0, vp_cblob_dp = 0;
  vb::DevBuf vpBase;  // base vp as set by vp_set: mu, sigma, lambda, w, eta, delta
  vb::DevBuf vpCur;   // unpacked vp of the current step
  vb::VpDev vp{};
  double* base_mu = nullptr; double* base_sigma = nullptr; double* base_lambda = nullptr;
  double* base_w = nullptr; double* base_eta = nullptr;

  // thetabnd
  int nbnd = 0;
  vb::DevBuf bnd;  // lb[n], ub[n]
  double TolCon = 0, WeightThreshold = 0, WeightPenalty = 0;

  // eps
  vb::DevBuf eps;
  vb::DevBuf zigTab;  // ziggurat tables of the draw generator (philox.cu)
  int epsD = 0, epsK = 0, epsNs = 0;
  bool eps_ready = false;
  bool eps_f32 = false;  // the resident draws are floats (generated by the device generator in FP32 mode)
  bool philox_pending = false;
  bool philox_dyn = false;  // read {seed, stream} from theta_dev[ntheta..] (set while building / replaying a step graph)
  uint64_t philox_seed = 0, philox_stream = 0;
  // Draws of the device generator are a pure function of (seed, stream, shape), so the NEXT step's draws (stream + 1) are generated
  // right after the entropy sweep has consumed the current ones, in the shadow of the single-CTA tail of the step (reduce, finalize,
  // Adam update).  `eps_key` says which draws the buffer holds; a step whose (seed, stream) matches skips its own generation.
  bool prefetch_enabled = true;      // VBMC_B200_PREFETCH=0 turns the ahead-of-time generation off
  bool glj_first = false;            // true (VBMC_B200_GLJ_FIRST=1): the sweep waits for the whole log-joint kernel; false: both are released together after the unpack kernel
  bool philox_trail = false;         // request: generate (seed, stream + 1) after the sweep of the step being enqueued
  bool trail_join_pending = false;   // stream3 carries an ahead-of-time generation that `stream` has not joined yet
  struct EpsKey {
    bool valid = false;
    uint64_t seed = 0, stream = 0;
    int D = 0, K = 0, Ns = 0;
    bool f32 = false;
  } eps_key;
  bool have_last_key = false;        // (seed, stream) of the previous generator-mode step: a +1 successor switches prefetching on
  uint64_t last_seed = 0, last_stream = 0;

  // step buffers
  vb::DevBuf theta_dev, out_dev, R_dev, ent_partial, ent_partial2, glj_out, glj_part, glj_ticket;
  vb::DevBuf ent_tables;  // FP32 sweep: per-step tables (entmc_f32.cu)
  // Cost-weighted schedule of the FP64 sweep (entmc2.cu): vp_unpack2_kernel estimates, per source component j, how many
  // components survive the pruning test and cuts the K * tpc CTA-tiles into G ranges of equal estimated COST instead of equal
  // count.  ent_plan: int tstart[G + 1] | jlo[K] | jhi[K].  ent_plan_req_*: request of the step being enqueued (consumed by
  // launch_vp_unpack); ent_plan_active: the sweep and its reduction of this step read the table.
  vb::DevBuf ent_plan;
  bool ent_balance = true;          // VBMC_B200_ENTMC_BALANCE=0: equal-count ranges (round-2 schedule)
  int ent_balance_c0 = 16;          // fixed cost of a tile in units of one scored component (VBMC_B200_ENTMC_C0)
  int ent_balance_crun = 0;         // fixed cost of a source component inside a range, same units (VBMC_B200_ENTMC_CRUN; 40 would follow from the per-SM
                                    // active cycles of the ncu capture at c3, but measured no gain: sweep 0.2539 / 0.2525 / 0.2520 ms at 0 / 40 / 80, c4 0.900 / 0.907)
  bool ent_plan_req = false;
  int ent_plan_req_tpc = 0, ent_plan_req_G = 0;
  bool ent_plan_active = false;
  double* theta_pinned = nullptr;
  double* out_pinned = nullptr;
  size_t theta_pinned_cap = 0, out_pinned_cap = 0;
  vb::DevBuf flush;  // L2 flush scratch
  vb::DevBuf varWork;  // variance path: Z/V, Gram, J, varF(s)
  vb::DevBuf entlbWork;  // entlb_vbmc: gamma[K][K], gsum, wraw, out

  // CUDA graph of one negelcbo step (single rank): replayed while the step signature is unchanged
  bool graphs_enabled = true;
  std::vector<long long> graph_key, warm_key;
  cudaGraphExec_t graph_exec = nullptr;
  long long graph_launches = 0;
  std::vector<long long> adam_key;    // one fminadam iteration (step kernels + Adam update, no host copies)
  cudaGraphExec_t adam_graph = nullptr;
  long long adam_graph_launches = 0;
  vb::DevBuf adamState, adamXtab;
  std::vector<long long> refit_key;   // same for the batched Cholesky of gp_post / gp_nlz
  cudaGraphExec_t refit_graph = nullptr;
  long long refit_launches = 0;

  // profiling
  bool profiling = false;
  std::map<std::string, vb::Prof> prof;
  std::vector<cudaEvent_t> prof_events;  // pending (start, stop) pairs
  std::vector<std::string> prof_names;   // kernel name of each pending pair
};

namespace vb {

// launch bookkeeping: every kernel launch goes through KernelScope so that launches are counted
// and (when profiling) bracketed by CUDA events on the launch stream.
struct KernelScope {
  vbmc_b200_ctx* c;
  const char* name;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  KernelScope(vbmc_b200_ctx* ctx, const char* nm, cudaStream_t s);
  ~KernelScope();
};
int profile_collect(vbmc_b200_ctx* c);
int gp_upload_derived(vbmc_b200_ctx* c, const vbmc_b200_gp_desc* g, int Ncov, int Nnoise, const double* sW1, const int* Lchol);  // sync + fold pending event pairs into c->prof

// ---- step pieces (each defined in its own .cu) ----
int launch_vp_unpack(vbmc_b200_ctx* c, bool have_theta);
int launch_entmc(vbmc_b200_ctx* c, int Ns, int need_mask, cudaStream_t st);
bool entmc2_balance_params(vbmc_b200_ctx* c, int Ns, int* tpc, int* G);
int launch_entmc_reduce(vbmc_b200_ctx* c, int Ns, int S_layout, cudaStream_t st);
int launch_gplogjoint(vbmc_b200_ctx* c, int all_samples, cudaStream_t st);
int launch_glj_reduce(vbmc_b200_ctx* c, cudaStream_t st, bool whole_step);
XchgDev step_push_target(vbmc_b200_ctx* c, int S_layout, bool whole_step);
int launch_gplogjoint_weighted(vbmc_b200_ctx* c, const double* wvec, double* out, cudaStream_t st);
int launch_finalize(vbmc_b200_ctx* c, int Ns, int compute_grad, int use_bnd, int jacobian, int what, cudaStream_t st, bool pushed = false);
int philox_init_tables(vbmc_b200_ctx* c);
int launch_philox(vbmc_b200_ctx* c, int D, int K, int Ns, uint64_t seed, uint64_t stream_id, cudaStream_t st,
                  const uint64_t* dyn = nullptr, uint64_t stream_add = 0);
int allreduce_R(vbmc_b200_ctx* c, int count, cudaStream_t st);
int run_entlb(vbmc_b200_ctx* c, int gmask, int jacobian, double* H, double* dH);
int launch_adam_step(vbmc_b200_ctx* c, const AdamArgs& a, cudaStream_t st);
int launch_adam_penalty(vbmc_b200_ctx* c, const AdamArgs& a, const double* corr, cudaStream_t st);
int launch_adam_check(vbmc_b200_ctx* c, const AdamArgs& a, int iter, double TolFun, cudaStream_t st);
int launch_adam_final(vbmc_b200_ctx* c, const AdamArgs& a, int iter, cudaStream_t st);
void comm_destroy(vbmc_b200_ctx* c);
int run_variance(vbmc_b200_ctx* c, int compute_var, std::vector<double>* varFs, std::vector<double>* J, std::vector<double>* vgrad = nullptr);
int run_factor_inverse(vbmc_b200_ctx* c, int N, int ld, const double* R, double* out);
int run_rhs_solve(vbmc_b200_ctx* c, int ncols, double* Z, double* W, const int* isfac_dev, cudaStream_t st);
int run_rhs_backsolve(vbmc_b200_ctx* c, int ncols, double* Z, cudaStream_t st, const int* isfac_dev = nullptr);
bool run_trsm_blocked(vbmc_b200_ctx* c, int T, double* Z, const int* isfac_dev, cudaStream_t st, int* rc, bool backward = false);
bool run_trsv1(vbmc_b200_ctx* c, double* Z, const int* isfac_dev, cudaStream_t st, int* rc, bool backward);
int trsv1_check(vbmc_b200_ctx* c);
int pad_identity(double* L, int N, int Np, int S, cudaStream_t st);
int entmc_num_tiles(vbmc_b200_ctx* c, int Ns, int* tiles_per_comp, int* pairs_per_tile, int* nwarps, size_t* smem);
void shard_range(int total, int nranks, int rank, int* begin, int* end);
int entmc_pick_dp(int D);

enum { NEED_MU = 1, NEED_E = 2, NEED_W = 4 };
enum { FIN_NEGELCBO = 0, FIN_ENTMC = 1, FIN_GPLOGJOINT = 2, FIN_NEGELCBO_NOENT = 3 };  // NOENT: F = -G + penalties (the caller adds the entropy bound)

}  // namespace vb
