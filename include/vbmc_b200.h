/*
 * vbmc_b200.h — C ABI of libvbmc_b200.so: the B200 (sm_100a) implementation of the VBMC
 * variational-optimisation hot path (negelcbo_vbmc + gradient, gplogjoint, entmc_vbmc,
 * vpbndloss) and of the GP-surrogate refit (gplite_post / gplite_nlZ -> gplite_core).
 *
 * The reference (acerbilab/vbmc) is pure MATLAB and has no FFI; the only replaceable seam is
 * the MATLAB function call itself (a MEX file shadows a same-named .m).  Every entry point
 * below therefore names the MATLAB function (reference file:line) whose body it replaces, and
 * INTEGRATION.md shows the MEX gateway that binds it.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; all arrays are HOST pointers unless the
 *     name ends in _dev; all matrices are MATLAB column-major FP64 (double).
 *   - every function returns 0 on success or a VBMC_B200_E* code; vbmc_b200_last_error()
 *     returns a thread-local message that starts with the MATLAB error identifier the
 *     reference would raise (e.g. "negelcbo_vbmc:vargrad: ...").
 *   - a context owns ONE GPU (one process per GPU, or several contexts in one process);
 *     calls on one context are serialised by the caller (MATLAB's interpreter thread).
 *   - there is NO CPU fallback: without a usable sm_100 device vbmc_b200_create() fails.
 */
#ifndef VBMC_B200_H
#define VBMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VBMC_B200_VERSION 100 /* 0.1.0 */

enum {
  VBMC_B200_OK = 0,
  VBMC_B200_EINVAL = 1,      /* bad argument / dimension mismatch                           */
  VBMC_B200_ECUDA = 2,       /* CUDA runtime / driver error                                 */
  VBMC_B200_ENODEV = 3,      /* no usable sm_100 device                                     */
  VBMC_B200_ESTATE = 4,      /* call order: e.g. negelcbo before gp_attach / vp_set         */
  VBMC_B200_EUNSUPPORTED = 5,/* reference feature outside this build's scope */
  VBMC_B200_EREFERENCE = 6,  /* the reference itself would raise error(...): id in message  */
  VBMC_B200_ENCCL = 7        /* NCCL error                                                  */
};

typedef struct vbmc_b200_ctx vbmc_b200_ctx;

/* ---------------------------------------------------------------------------------------
 * context
 * ------------------------------------------------------------------------------------- */
int vbmc_b200_version(void);
const char* vbmc_b200_last_error(void);
/* device: CUDA ordinal; fails with ENODEV when the device is not compute capability 10.x. */
int vbmc_b200_create(vbmc_b200_ctx** out, int device);
int vbmc_b200_destroy(vbmc_b200_ctx* ctx);
/* Process-wide context of `device`, created on first use and owned by the library (do not destroy it; released by
 * vbmc_b200_shared_release or at process exit).  Every MEX gateway of mex/ is its own shared object but links this one
 * library, so this is how negelcbo_vbmc, gplite_post, gplite_pred ... see the same resident GP posterior and draws. */
int vbmc_b200_shared(vbmc_b200_ctx** out, int device);
int vbmc_b200_shared_release(void);
int vbmc_b200_sync(vbmc_b200_ctx* ctx);
/* number of kernels this library launched on the context since creation (bench: gpu_launches) */
int vbmc_b200_launch_count(vbmc_b200_ctx* ctx, long long* count);
/* Arithmetic of the Monte-Carlo entropy sweep (ent/entmc_vbmc.m:49-104): bits = 64 (default; 1e-10 rel against
 * the reference) or 32 (BASELINE config 5; 1e-4 rel).  With 32 the per-draw scoring of the K components runs
 * in FP32 with hardware exp2; every sum over draws, the expected log-joint, Jacobians and penalties stay FP64.
 * The reference has no such switch (MATLAB doubles throughout): a caller that never calls this gets FP64. */
int vbmc_b200_set_precision(vbmc_b200_ctx* ctx, int bits);

/* Entropy sweep: a mixture component whose term is below exp(-log_threshold) of q(x) for ALL 32 draws a warp scores
 * (bound from ||u_jk||, sigma_j/sigma_k and the warp's largest ||eps||; see csrc/entmc.cu) is skipped for that warp.
 * Default 50 (2e-22 of q: below FP64 round-off; the reference's exp() underflows to exactly 0 for most such terms);
 * 0 scores every component like the reference loop (ent/entmc_vbmc.m:60-65).  Also VBMC_B200_ENTMC_PRUNE in the environment. */
int vbmc_b200_entmc_prune(vbmc_b200_ctx* ctx, double log_threshold);
/* counters of (warp, component) blocks scored / candidate since the last call; enable != 0 keeps counting (bench) */
int vbmc_b200_entmc_prune_stats(vbmc_b200_ctx* ctx, int enable, unsigned long long* kept, unsigned long long* total);
/* Entropy sweep schedule (FP64 sweep, K <= 256, D < 20): the K*tpc CTA-tiles of a step (tile = one group of 32 antithetic pairs
 * per warp, all of one source component j) are cut into one contiguous range per SM.  on != 0 (default): ranges of equal estimated
 * COST -- per source component the unpack kernel counts the components that survive the pruning test above for a typical warp
 * of draws, a tile of component j weighs c0 + that count and every component charges a fixed start cost (its table build;
 * VBMC_B200_ENTMC_CRUN, default 0: measured no gain); on == 0, or fewer than two tiles per CTA: ranges of equal tile count.  c0 <= 0 keeps the current
 * value (default 16).  Results are sums in a fixed order for a given theta either way; the two schedules differ by round-off.
 * Also VBMC_B200_ENTMC_BALANCE / VBMC_B200_ENTMC_C0 in the environment.  No counterpart in the reference (ent/entmc_vbmc.m:60-100
 * is one vectorised loop over components). */
int vbmc_b200_entmc_balance(vbmc_b200_ctx* ctx, int on, int c0);
/* the schedule of the LAST step that ran the balanced sweep: tstart[G + 1] tile boundaries, then first / last CTA of every
 * component (jlo[K], jhi[K]); out holds at least G + 1 + 2 K ints (cap).  G == 0: the last step used equal counts. */
int vbmc_b200_entmc_plan_get(vbmc_b200_ctx* ctx, int* out, int cap, int* G, int* tpc);

/* ---------------------------------------------------------------------------------------
 * multi-GPU: one context per rank, one all-reduce of the partial sums per negelcbo step.
 * The MC pair axis of entmc_vbmc and the hyper-parameter-sample axis of gplogjoint are
 * sharded over ranks; results are replicated (bit-identical) on every rank after the all-reduce.
 * comm_init builds an NCCL communicator (rendezvous, fallback) and, over it, exchanges CUDA IPC handles so that
 * every rank maps every other rank's exchange buffer: the per-step all-reduce then runs INSIDE the step's last
 * kernel over NVLink peer memory (push to all inboxes, flag, wait, sum in rank order) — no NCCL call and no extra
 * launch per step, and the whole multi-GPU step is CUDA-graph capturable.  VBMC_B200_P2P=0 keeps NCCL.
 * ------------------------------------------------------------------------------------- */
#define VBMC_B200_UNIQUE_ID_BYTES 128
int vbmc_b200_comm_unique_id(void* id128);                                  /* rank 0, then broadcast out of band */
int vbmc_b200_comm_init(vbmc_b200_ctx* ctx, int nranks, int rank, const void* id128);
int vbmc_b200_comm_info(vbmc_b200_ctx* ctx, int* nranks, int* rank);
/* *peer_memory = 1 when the per-step all-reduce runs over mapped peer memory, 0 when it goes through NCCL */
int vbmc_b200_comm_p2p(vbmc_b200_ctx* ctx, int* peer_memory);
/* contiguous balanced split of `total` units (MC pairs per component, hyper-parameter samples) over ranks */
int vbmc_b200_shard_range(int total, int nranks, int rank, int* begin, int* end);

/* ---------------------------------------------------------------------------------------
 * GP struct (reference type: gplite/gplite_post.m:94-151; consumed by misc/gplogjoint.m:32-45)
 * ------------------------------------------------------------------------------------- */
typedef struct vbmc_b200_gp_desc {
  int N, D, S;        /* training points, dimension, hyper-parameter samples numel(gp.post)      */
  int Nhyp;           /* rows of hyp = Ncov + Nnoise + Nmean                                      */
  int covfun;         /* gp.covfun(1); only 1 (SE-ARD) (gplite_core.m:52)                         */
  int meanfun;        /* gp.meanfun; 0 zero, 1 const, 4 negquad (gplite_meanfun.m cases 0,1,4)    */
  int noisefun[3];    /* gp.noisefun, gplite_noisefun.m:176-210                                   */
  const double* X;    /* N x D  (gp.X)                                                            */
  const double* y;    /* N      (gp.y)                 — gp_post / gp_nlz only                    */
  const double* s2;   /* N or NULL (gp.s2)             — gp_post / gp_nlz only                    */
  const double* hyp;  /* Nhyp x S (gp.post(s).hyp)                                                */
} vbmc_b200_gp_desc;

/* Adopt a posterior MATLAB already computed (needed when only negelcbo_vbmc is shadowed).
 * alpha: N x S (gp.post(s).alpha); sW1: S values gp.post(s).sW(1); Lchol: S flags;
 * L: N x N x S upper factors (gp.post(s).L) or NULL when variances are never requested. */
int vbmc_b200_gp_attach(vbmc_b200_ctx* ctx, const vbmc_b200_gp_desc* gp, const double* alpha,
                        const double* sW1, const int* Lchol, const double* L);

/* Caller's fingerprint of the resident posterior.  The MEX gateways hash the data addresses of every gp.post(s).alpha (MATLAB's
 * copy-on-write keeps them stable until a posterior is recomputed), S, N, D, Nhyp, covfun/meanfun/noisefun and the CONTENT of
 * every hyper-parameter vector plus first/last alpha values (mex/vbmc_b200_mex_common.h); the address of post(1).alpha alone is
 * NOT a fingerprint -- `gp1 = gp; gp1.post = gp.post(1)` (activeimportancesampling_vbmc.m:170-171) shares it with another S.
 * The library only stores the tag and resets it to 0 whenever the resident posterior changes (gp_attach, gp_post,
 * gp_post_update1, gp_nlz*), so that several gateways sharing one context agree on whether `gp` has to be attached again. */
int vbmc_b200_gp_tag_set(vbmc_b200_ctx* ctx, unsigned long long tag);
int vbmc_b200_gp_tag_get(vbmc_b200_ctx* ctx, unsigned long long* tag);
/* Shape of the resident posterior (all 0 when none): callers that size output arrays from their own struct must check it
 * matches before a compute call -- the library always computes with the RESIDENT N, D, S. */
int vbmc_b200_gp_shape(vbmc_b200_ctx* ctx, int* N, int* D, int* S);

/* gplite_post(hyp,X,y,covfun,meanfun,noisefun,s2)  — gplite/gplite_post.m:94-172, i.e. S x
 * gplite_core(hyp,gp,0,0) (gplite/private/gplite_core.m:33-102,278-285): SE-ARD Gram, Cholesky
 * with the x10 jitter retry, alpha.  The posterior stays resident on the device (it becomes the
 * attached GP); any non-NULL output is also copied to the host:
 *   alpha N x S, L N x N x S (upper; -inv(K+Sigma) when Lchol==0), sW1 S, sn2_mult S, Lchol S. */
int vbmc_b200_gp_post(vbmc_b200_ctx* ctx, const vbmc_b200_gp_desc* gp, double* alpha, double* L,
                      double* sW1, double* sn2_mult, int* Lchol);

/* [ymu,ys2,fmu,fs2,lp] = gplite_pred(gp,Xstar,ystar,s2star,ssflag) — gplite/gplite_pred.m:1-163 for the attached posterior
 * (SE-ARD covfun, meanfun 0/1/4, no integrated mean, no output warping: the VBMC defaults).
 * Xstar: Nstar x D column-major; ystar, s2star: Nstar or NULL.  want_var == 0 is nargout == 1 (means only).
 * Outputs (any may be NULL): Nstar x S column-major when ssflag != 0 or S == 1, else Nstar (averaged over the
 * samples, :153-163); lp is always Nstar x S (the reference never averages it) and needs ystar and want_var. */
int vbmc_b200_gp_pred(vbmc_b200_ctx* ctx, int Nstar, const double* Xstar, const double* ystar, const double* s2star,
                      int ssflag, int want_var, double* ymu, double* ys2, double* fmu, double* fs2, double* lp);
/* gp.post(s).sn2_mult of an attached posterior (S values; 1 after vbmc_b200_gp_attach, set by vbmc_b200_gp_post):
 * gplite_pred.m:119 scales the test-point noise with it. */
int vbmc_b200_gp_set_sn2_mult(vbmc_b200_ctx* ctx, const double* sn2_mult);

/* gp = gplite_post(gp,xstar,ystar,[],[],[],[],1) — the rank-one update of gplite/gplite_post.m:50-92,173-251 applied to the
 * resident posterior (constant or output-dependent noise, no s2: with s2 the reference itself refits).
 * xstar: D values.  Optional outputs: alpha (N+1) x S, sW_new S, and Lcol (N+1) x S = the last column of the new
 * gp.post(s).L for the Cholesky samples (:226-233, L_new = [L c; 0 d]).  Low-noise samples (Lchol == 0, :234-238) change
 * every entry of L = -inv(K + diag): fetch those with vbmc_b200_gp_get_factor; their Lcol is the new column of whatever the device
 * keeps (the inverse handed over by gp_attach, or the unscaled factor a device refit left).
 * The context's GP then has N+1 training points (private/activesample_vbmc.m:483 calls this once per acquired point). */
int vbmc_b200_gp_post_update1(vbmc_b200_ctx* ctx, const double* xstar, double ystar, double* alpha, double* Lcol,
                              double* sW_new);
/* gp.post(s).L of the resident posterior as the reference stores it (gplite/private/gplite_core.m:67-100): the upper Cholesky
 * factor, or -inv(K + sn2_mult*diag(sn2)) for a low-noise sample.  L: N x N column-major, N = the resident point count. */
int vbmc_b200_gp_get_factor(vbmc_b200_ctx* ctx, int s, double* L);

/* Hyper-prior of gplite_nlZ (gplite/gplite_hypprior.m:18-58); arrays of length Nhyp. */
typedef struct vbmc_b200_hprior {
  const double* mu;
  const double* sigma;
  const double* df;   /* NULL => 7 for every hyper-parameter (gplite_hypprior.m:33-35) */
} vbmc_b200_hprior;

/* [nlZ,dnlZ] = gplite_nlZ(hyp,gp,hprior) — gplite/gplite_nlZ.m:27-66 -> gplite_core(hyp,gp,1,grad)
 * (gplite_core.m:193,226-261).  gp->S must be 1 and gp->hyp one column; dnlZ (Nhyp) may be NULL
 * (no gradient, like nargout==1).  hprior may be NULL. */
int vbmc_b200_gp_nlz(vbmc_b200_ctx* ctx, const vbmc_b200_gp_desc* gp, const vbmc_b200_hprior* hprior,
                     double* nlZ, double* dnlZ);

/* nlZ[s] = gplite_nlZ(hyp(:,s),gp,hprior), s = 1..gp->S, value only, as ONE batched Gram + Cholesky:
 * what gplite_train.m evaluates one call at a time for the fminfill design (:200-204) and the slice sampler (:318-330). */
int vbmc_b200_gp_nlz_batch(vbmc_b200_ctx* ctx, const vbmc_b200_gp_desc* gp, const vbmc_b200_hprior* hprior, double* nlZ);

/* ---------------------------------------------------------------------------------------
 * VP struct (reference type: misc/setupvars_vbmc.m:78-99)
 * ------------------------------------------------------------------------------------- */
typedef struct vbmc_b200_vp_desc {
  int D, K;
  const double* mu;      /* D x K                                                        */
  const double* sigma;   /* K                                                            */
  const double* lambda;  /* D                                                            */
  const double* w;       /* K                                                            */
  const double* eta;     /* K or NULL (then log(w)); vp.eta is what J_w uses             */
  const double* delta;   /* D, or NULL for vp.delta empty/0 (gplogjoint.m:86-90)         */
  int optimize_mu, optimize_sigma, optimize_lambda, optimize_weights;
} vbmc_b200_vp_desc;

int vbmc_b200_vp_set(vbmc_b200_ctx* ctx, const vbmc_b200_vp_desc* vp);

/* thetabnd struct of misc/vpbounds.m:32-52; pass n == 0 to clear (thetabnd == []). */
int vbmc_b200_thetabnd_set(vbmc_b200_ctx* ctx, int n, const double* lb, const double* ub, double TolCon,
                           double WeightThreshold, double WeightPenalty);

/* ---------------------------------------------------------------------------------------
 * entropy draws.  The reference draws epsilon = randn(D,1,Ns/2) per component from MATLAB's
 * global stream (ent/entmc_vbmc.m:53).  Two sources:
 *   parity mode : the caller supplies the draws (host buffer, D x Ns/2 x K column-major);
 *   device mode : counter-based Philox4x32-10 feeding a 1024-strip ziggurat (FP32 mode: Box-Muller), keyed by
 *                 (seed, call counter, element index) — independent of the number of GPUs and of the sharding.
 *                 A caller whose call counter advances by one per call gets the next call's draws generated
 *                 ahead of time, in the tail of the current call.
 * ------------------------------------------------------------------------------------- */
enum { VBMC_B200_EPS_HOST = 0, VBMC_B200_EPS_RESIDENT = 1, VBMC_B200_EPS_PHILOX = 2 };
/* copy draws to the device once; later calls may use VBMC_B200_EPS_RESIDENT */
int vbmc_b200_eps_upload(vbmc_b200_ctx* ctx, int D, int K, int Ns, const double* eps);
/* fill the resident buffer with Philox draws and (optionally, eps_out != NULL) read them back */
int vbmc_b200_eps_philox(vbmc_b200_ctx* ctx, int D, int K, int Ns, uint64_t seed, uint64_t stream,
                         double* eps_out);

/* ---------------------------------------------------------------------------------------
 * [F,dF,G,H,varF,dH,varGss,varG,varH,I_sk,J_sjk] =
 *     negelcbo_vbmc(theta,beta,vp,gp,Ns,compute_grad,compute_var,altent_flag,thetabnd,entropy_alpha)
 * misc/negelcbo_vbmc.m:1-164 -> misc/gplogjoint.m:1-413, ent/entmc_vbmc.m:1-125,
 * misc/vpbndloss.m:1-72, utils/softbndloss.m:9-28.
 * The MEX gateway resolves MATLAB's nargin/nargout defaults (negelcbo_vbmc.m:9-17) and fills
 * this struct; vp / gp / thetabnd are the ones last set on the context.
 * ------------------------------------------------------------------------------------- */
typedef struct vbmc_b200_negelcbo_args {
  /* inputs */
  const double* theta;   /* ntheta (host)                                                 */
  int ntheta;
  double beta;           /* 0 or non-finite => 0 (negelcbo_vbmc.m:15)                      */
  int Ns;                /* MC draws per component, made even like entmc_vbmc.m:45; 0 = deterministic entropy
                            lower bound entlb_vbmc instead of the Monte-Carlo estimate (negelcbo_vbmc.m:102-109) */
  int compute_grad;      /* 0/1                                                            */
  int compute_var;       /* 0 none, 1 full, 2 diagonal (gplogjoint.m:273,306)              */
  int separate_K;        /* nargout > 9: fill I_sk (and J_sjk when compute_var)            */
  int use_thetabnd;      /* 0 => thetabnd == [] for this call (negelcbo_vbmc.m:136)        */
  int eps_mode;          /* VBMC_B200_EPS_*                                                */
  const double* eps;     /* EPS_HOST: D x Ns/2 x K draws                                   */
  uint64_t seed, stream; /* EPS_PHILOX: key and call counter                               */
  /* outputs (host; NULL = not wanted) */
  double* F;             /* 1      */
  double* dF;            /* ntheta */
  double* G;             /* 1      */
  double* H;             /* 1      */
  double* varF;          /* 1      */
  double* dH;            /* ntheta */
  double* varGss;        /* 1      */
  double* varG;          /* 1      */
  double* varH;          /* 1      */
  double* I_sk;          /* S x K  */
  double* J_sjk;         /* S x K x K */
} vbmc_b200_negelcbo_args;

int vbmc_b200_negelcbo(vbmc_b200_ctx* ctx, const vbmc_b200_negelcbo_args* args);

/* ---------------------------------------------------------------------------------------
 * [x,f,xtab,ftab,iter] = fminadam(fun,x0,LB,UB,TolFun,MaxIter,master_stepsize) — utils/fminadam.m:1-102,
 * with fun = @(theta_) negelcbo_vbmc(theta_,beta,vp0,gp,Ns,1,compute_var,altent,thetabnd,entropy_alpha), the closure
 * misc/vpoptimize_vbmc.m:71 builds and passes at :127.  The iterate, the Adam moments and
 * the histories stay on the device: one call = one whole stochastic optimisation, no per-iteration host copies;
 * the host only reads the termination flag every 20 iterations (:65-83).
 * vp / gp / thetabnd are the ones last set on the context.  With beta ~= 0 (compute_var must be 2, negelcbo_vbmc.m:19-20; the
 * factors gp.post(s).L must be resident) every iteration also runs the variance path and its O(S K^2) host assembly
 * (negelcbo_vbmc.m:119-130): the loop is then host-driven with one synchronisation per iteration instead of a replayed graph.
 * ------------------------------------------------------------------------------------- */
typedef struct vbmc_b200_fminadam_args {
  /* inputs */
  const double* x0;       /* nvars (= numel(theta0), vpoptimize_vbmc.m:58-63)                          */
  int nvars;
  const double* LB;       /* nvars or NULL (= -Inf, fminadam.m:35)                                     */
  const double* UB;       /* nvars or NULL (= +Inf, fminadam.m:36)                                     */
  double TolFun;          /* <= 0 or NaN => 0.001 (fminadam.m:6)                                       */
  int MaxIter;            /* <= 0 => 10000 (fminadam.m:7); must be >= 20 (the reference indexes iter-19) */
  double stepsize_max, stepsize_min, stepsize_decay; /* <= 0 or NaN => 0.1, 0.001, 200 (fminadam.m:11-18) */
  double beta;            /* objective arguments, as vbmc_b200_negelcbo_args                            */
  int Ns;
  int compute_var;        /* no effect on F, dF when beta == 0; must be 2 (diagonal) when beta ~= 0     */
  int use_thetabnd;
  int eps_mode;           /* EPS_HOST / EPS_RESIDENT: the same draws every iteration (parity mode);
                             EPS_PHILOX: iteration i (0-based) draws from Philox stream `stream + i`     */
  const double* eps;
  uint64_t seed, stream;
  /* outputs (host; NULL = not wanted) */
  double* x;              /* nvars: mean of the last 20 iterates (fminadam.m:95)                        */
  double* f;              /* mean of the last 20 objective values (fminadam.m:96)                       */
  double* xtab;           /* nvars x MaxIter column-major; the first *iter columns are written (:63,98) */
  double* ftab;           /* MaxIter; the first *iter entries are written (:48,99)                      */
  int* iter;              /* iterations executed                                                       */
  double* stats;          /* optional [5]: stop flag, dx, slope, slope_err, slope_err_max of the last test */
} vbmc_b200_fminadam_args;

int vbmc_b200_fminadam(vbmc_b200_ctx* ctx, const vbmc_b200_fminadam_args* args);

/* [H,dH] = entmc_vbmc(vp,Ns,grad_flags,jacobian_flag) — ent/entmc_vbmc.m:1-125.
 * Uses the vp last set (no theta unpacking).  grad_flags[4]; dH length = D*K*gf0 + K*gf1 + D*gf2 + K*gf3. */
int vbmc_b200_entmc(vbmc_b200_ctx* ctx, int Ns, const int grad_flags[4], int jacobian_flag, int eps_mode,
                    const double* eps, uint64_t seed, uint64_t stream, double* H, double* dH);

/* [H,dH] = entlb_vbmc(vp,grad_flags,jacobian_flag) — ent/entlb_vbmc.m:1-147, the deterministic entropy lower bound that
 * negelcbo_vbmc uses instead of entmc_vbmc when Ns == 0 (negelcbo_vbmc.m:102-109; misc/vpsieve_vbmc.m:76 scores every
 * candidate posterior that way).  Uses the vp last set; dH layout as vbmc_b200_entmc.  vbmc_b200_negelcbo handles
 * Ns == 0 itself (same outputs as for Ns > 0, H and dH being the bound and its gradient). */
int vbmc_b200_entlb(vbmc_b200_ctx* ctx, const int grad_flags[4], int jacobian_flag, double* H, double* dH);

/* [F,dF,varF,dvarF,varss,I_sk,J_sjk] = gplogjoint(vp,gp,grad_flags,avg_flag,jacobian_flag,compute_var,separate_K)
 * misc/gplogjoint.m:1-413.  avg_flag must be 1 when S > 1.  dvarF: compute_var == 2 only
 * (else the reference raises gplogjoint:FullVarianceGradient). */
int vbmc_b200_gplogjoint(vbmc_b200_ctx* ctx, const int grad_flags[4], int avg_flag, int jacobian_flag,
                         int compute_var, double* F, double* dF, double* varF, double* dvarF, double* varss,
                         double* I_sk, double* J_sjk);

/* ---------------------------------------------------------------------------------------
 * benchmark / profiling hooks (not part of the reference surface)
 * ------------------------------------------------------------------------------------- */
/* Run `steps` consecutive negelcbo evaluations entirely on the device (theta, eps already
 * resident; no host copies) and return the CUDA-event time of the whole region in ms. */
int vbmc_b200_negelcbo_resident_loop(vbmc_b200_ctx* ctx, const vbmc_b200_negelcbo_args* args, int steps,
                                     float* ms_total);
/* per-kernel CUDA-event timing: when enabled each kernel of a step is bracketed by events on
 * the launch stream (graph replay is bypassed).  names: "entmc","gplogjoint","vp_unpack",
 * "reduce","finalize","philox","gram","potrf",...  returns accumulated ms and launch count. */
int vbmc_b200_profile_enable(vbmc_b200_ctx* ctx, int on);
int vbmc_b200_profile_get(vbmc_b200_ctx* ctx, const char* name, double* ms_sum, long long* launches);
int vbmc_b200_profile_reset(vbmc_b200_ctx* ctx);
/* measured FP64 FMA peak of this device (DFMA micro-benchmark, TFLOP/s) and copy bandwidth (GB/s) */
int vbmc_b200_measure_fp64_peak(vbmc_b200_ctx* ctx, double* tflops);
int vbmc_b200_measure_hbm_copy(vbmc_b200_ctx* ctx, double* gbs);
/* raw Philox4x32-10 block for the known-answer test (Random123 vectors) */
int vbmc_b200_philox_raw(vbmc_b200_ctx* ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* write >126 MB to flush L2 between timed iterations */
int vbmc_b200_flush_l2(vbmc_b200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VBMC_B200_H */
