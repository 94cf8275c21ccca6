"""CPU oracle for the VBMC variational-optimisation hot path (TEST INFRASTRUCTURE).

This module is a plain NumPy/SciPy FP64 restatement of the reference MATLAB code
of acerbilab/vbmc for the path named in BASELINE.json:

    negelcbo_vbmc -> gplogjoint, entmc_vbmc, vpbndloss/softbndloss
    gplite_post / gplite_nlZ -> gplite_core (sq_dist, SE-ARD, meanfun 0/1/4, noisefun)

It is *test infrastructure only*: nothing under ``vbmc_b200/`` imports it; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may.

PARITY UNPINNED: the reference ships no golden vectors for this path and neither
MATLAB nor Octave exists in the build container, so the
oracle cannot be checked against reference outputs.  It is pinned instead by the
analytic identities in ``tests/test_oracle_identities.py`` (K=1 closed-form
entropy, gplite_quad cross-formulation, finite differences, Cholesky residuals,
multivariate-normal log-density) and by a three-way agreement NumPy <-> C <-> CUDA.

Conventions: MATLAB structs are dicts.  ``vp``: D, K, mu (D,K), sigma (K,),
lambda (D,), w (K,), eta (K,), delta (None|scalar|(D,)), optimize_mu/sigma/lambda/
weights.  ``gp``: X (N,D), y (N,), s2 (None|(N,)), Ncov, Nnoise, Nmean, covfun,
meanfun, noisefun, meanfun_extras, intmeanfun, post = list of dicts
{hyp, alpha, sW, L, sn2_mult, Lchol}.  The reference draws ``epsilon`` from
MATLAB's global ``randn`` stream (ent/entmc_vbmc.m:53); here it is an explicit
input of shape (K, Ns/2, D) (C order == MATLAB D x Ns/2 x K column-major, i.e. the
stream order j outer, sample, d fastest).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla
from scipy.special import gammaln

EPS = np.finfo(np.float64).eps


class OracleError(Exception):
    """Mirrors MATLAB error('id:tag', msg); ``identifier`` carries the id."""

    def __init__(self, identifier: str, msg: str):
        super().__init__(f"{identifier}: {msg}")
        self.identifier = identifier


# --------------------------------------------------------------------------------------
# theta <-> vp helpers
# --------------------------------------------------------------------------------------
def rescale_params(vp: dict, theta=None) -> dict:
    """misc/rescale_params.m:6-40."""
    vp = dict(vp)
    D = vp["D"]
    if theta is not None:
        K = vp["K"]
        theta = np.asarray(theta, dtype=np.float64).ravel()
        idx = 0
        if vp["optimize_mu"]:
            vp["mu"] = theta[: D * K].reshape(K, D).T.copy()
            idx = D * K
        if vp["optimize_sigma"]:
            vp["sigma"] = np.exp(theta[idx : idx + K])
            idx += K
        if vp["optimize_lambda"]:
            vp["lambda"] = np.exp(theta[idx : idx + D])
        if vp["optimize_weights"]:
            eta = theta[-K:].copy()
            eta = eta - eta.max()
            vp["w"] = np.exp(eta)
    lam = np.asarray(vp["lambda"], dtype=np.float64).ravel()
    nl = math.sqrt(np.sum(lam**2) / D)
    vp["lambda"] = lam / nl
    vp["sigma"] = np.asarray(vp["sigma"], dtype=np.float64).ravel() * nl
    if vp["optimize_weights"]:
        w = np.asarray(vp["w"], dtype=np.float64).ravel()
        vp["w"] = w / w.sum()
        vp.pop("eta", None)
    vp.pop("mode", None)
    return vp


def get_vptheta(vp: dict):
    """misc/get_vptheta.m:17-21 (flags taken from vp)."""
    vp = rescale_params(vp)
    parts = []
    if vp["optimize_mu"]:
        parts.append(np.asarray(vp["mu"]).T.ravel())  # mu(:) column-major, d fastest
    if vp["optimize_sigma"]:
        parts.append(np.log(vp["sigma"]))
    if vp["optimize_lambda"]:
        parts.append(np.log(vp["lambda"]))
    if vp["optimize_weights"]:
        parts.append(np.log(vp["w"]))
    return np.concatenate(parts), vp


def vpbounds(vp: dict, gp: dict, options: dict, K=None):
    """misc/vpbounds.m:8-52.  options: TolLength, TolWeight, TolConLoss, WeightPenalty."""
    vp = dict(vp)
    if K is None:
        K = vp["K"]
    D = vp["D"]
    b = dict(vp.get("bounds") or {})
    if not b:
        b["mu_lb"] = np.full(D, np.inf)
        b["mu_ub"] = np.full(D, -np.inf)
        b["lnscale_lb"] = np.full(D, np.inf)
        b["lnscale_ub"] = np.full(D, -np.inf)
    X = gp["X"]
    b["mu_lb"] = np.minimum(X.min(axis=0), b["mu_lb"])
    b["mu_ub"] = np.maximum(X.max(axis=0), b["mu_ub"])
    lnrange = np.log(X.max(axis=0) - X.min(axis=0))
    b["lnscale_lb"] = np.minimum(b["lnscale_lb"], lnrange + math.log(options["TolLength"]))
    b["lnscale_ub"] = np.maximum(b["lnscale_ub"], lnrange)
    if vp["optimize_weights"]:
        b["eta_lb"] = math.log(0.5 * options["TolWeight"])
        b["eta_ub"] = 0.0
    vp["bounds"] = b
    lb, ub = [], []
    if vp["optimize_mu"]:
        lb.append(np.tile(b["mu_lb"], K))
        ub.append(np.tile(b["mu_ub"], K))
    if vp["optimize_sigma"] or vp["optimize_lambda"]:
        lb.append(np.tile(b["lnscale_lb"], K))
        ub.append(np.tile(b["lnscale_ub"], K))
    if vp["optimize_weights"]:
        lb.append(np.full(K, b["eta_lb"]))
        ub.append(np.full(K, b["eta_ub"]))
    thetabnd = {"lb": np.concatenate(lb), "ub": np.concatenate(ub), "TolCon": options["TolConLoss"]}
    if vp["optimize_weights"]:
        thetabnd["WeightThreshold"] = max(1.0 / (4 * K), options["TolWeight"])
        thetabnd["WeightPenalty"] = options["WeightPenalty"]
    return vp, thetabnd


def softbndloss(x, slb, sub, TolCon=1e-3, compute_grad=True):
    """utils/softbndloss.m:9-28."""
    x = np.asarray(x, dtype=np.float64)
    ell = (sub - slb) * TolCon
    y = 0.0
    dy = np.zeros_like(x)
    idx = x < slb
    if idx.any():
        y += 0.5 * np.sum(((slb[idx] - x[idx]) / ell[idx]) ** 2)
        if compute_grad:
            dy[idx] = (x[idx] - slb[idx]) / ell[idx] ** 2
    idx = x > sub
    if idx.any():
        y += 0.5 * np.sum(((x[idx] - sub[idx]) / ell[idx]) ** 2)
        if compute_grad:
            dy[idx] = (x[idx] - sub[idx]) / ell[idx] ** 2
    return y, dy


def vpbndloss(theta, vp, thetabnd, TolCon, compute_grad=True):
    """misc/vpbndloss.m:9-71."""
    K, D = vp["K"], vp["D"]
    theta = np.asarray(theta, dtype=np.float64).ravel()
    if vp["optimize_mu"]:
        mu = theta[: K * D]
        idx = K * D
    else:
        mu = np.asarray(vp["mu"]).T.ravel()
        idx = 0
    if vp["optimize_sigma"]:
        lnsigma = theta[idx : idx + K]
        idx += K
    else:
        lnsigma = np.log(np.asarray(vp["sigma"]).ravel())
    if vp["optimize_lambda"]:
        lnlambda = theta[idx : idx + D]
    else:
        lnlambda = np.log(np.asarray(vp["lambda"]).ravel())
    eta = theta[-K:] if vp["optimize_weights"] else np.zeros(0)
    lnscale = lnsigma[None, :] + lnlambda[:, None]  # (D,K)
    ext = []
    if vp["optimize_mu"]:
        ext.append(mu)
    if vp["optimize_sigma"] or vp["optimize_lambda"]:
        ext.append(lnscale.T.ravel())  # lnscale(:) column-major (d fastest)
    if vp["optimize_weights"]:
        ext.append(eta)
    theta_ext = np.concatenate(ext)
    L, dLext = softbndloss(theta_ext, thetabnd["lb"], thetabnd["ub"], TolCon, compute_grad)
    if not compute_grad:
        return L, None
    out = []
    idx = 0
    if vp["optimize_mu"]:
        out.append(dLext[: D * K])
        idx = D * K
    if vp["optimize_sigma"] or vp["optimize_lambda"]:
        dlnscale = dLext[idx : idx + D * K].reshape(K, D).T  # (D,K)
        if vp["optimize_sigma"]:
            out.append(dlnscale.sum(axis=0))
        if vp["optimize_lambda"]:
            out.append(dlnscale.sum(axis=1))
    if vp["optimize_weights"]:
        out.append(dLext[-K:])
    return L, np.concatenate(out)


def _softmax_jacobian(eta):
    """J_w = -exp(eta)' * exp(eta)/sum^2 + diag(exp(eta)/sum)  (gplogjoint.m:366-367)."""
    e = np.exp(np.asarray(eta, dtype=np.float64).ravel())
    es = e.sum()
    return -np.outer(e, e / es**2) + np.diag(e / es)


# --------------------------------------------------------------------------------------
# entropy
# --------------------------------------------------------------------------------------
def entmc_vbmc(vp, Ns, grad_flags=True, jacobian_flag=True, epsilon=None, nargout=2):
    """ent/entmc_vbmc.m:16-125, vectorised like the .m (same D x Ns x K temporaries).

    ``epsilon``: (K, Ns/2, D) normal draws replacing randn(D,1,Ns/2) per component j.
    Returns (H, dH).
    """
    if nargout < 2:
        grad_flags = False
    if np.isscalar(grad_flags):
        grad_flags = [bool(grad_flags)] * 4
    gf = [bool(g) for g in grad_flags]
    D, K = vp["D"], vp["K"]
    mu = np.asarray(vp["mu"], dtype=np.float64).reshape(D, K)
    sigma = np.asarray(vp["sigma"], dtype=np.float64).ravel()
    lam = np.asarray(vp["lambda"], dtype=np.float64).ravel()
    w = np.asarray(vp["w"], dtype=np.float64).ravel()

    mu_grad = np.zeros((D, K)) if gf[0] else np.zeros((0,))
    sigma_grad = np.zeros(K) if gf[1] else np.zeros((0,))
    lambda_grad = np.zeros(D) if gf[2] else np.zeros((0,))
    w_grad = np.zeros(K) if gf[3] else np.zeros((0,))

    sigmalambda = lam[:, None, None] * sigma[None, None, :]  # (D,1,K)
    nconst = 1.0 / (2 * math.pi) ** (D / 2) / np.prod(lam)
    nf = nconst

    Ns = int(math.ceil(Ns / 2) * 2)
    half = Ns // 2
    epsilon = np.asarray(epsilon, dtype=np.float64)
    if epsilon.shape != (K, half, D):
        raise ValueError(f"epsilon must have shape (K,Ns/2,D)=({K},{half},{D}), got {epsilon.shape}")
    H = 0.0
    for j in range(K):
        eps_half = epsilon[j].T  # (D, Ns/2)
        eps = np.concatenate([eps_half, -eps_half], axis=1)  # (D,Ns) antithetic, :53-54
        xi = eps * lam[:, None] * sigma[j] + mu[:, j : j + 1]  # (D,Ns)   :55
        Xs = xi.T  # (Ns,D)
        ys = np.zeros(Ns)
        for k in range(K):  # :60-65
            d2 = np.sum(((Xs - mu[:, k][None, :]) / (sigma[k] * lam[None, :])) ** 2, axis=1)
            nn = w[k] * nf / sigma[k] ** D * np.exp(-0.5 * d2)
            ys = ys + nn
        with np.errstate(divide="ignore"):
            H = H - w[j] * np.sum(np.log(ys)) / Ns  # :67
        if any(gf):
            diff = xi[:, :, None] - mu[:, None, :]  # (D,Ns,K)
            norm_jl = (nconst / sigma**D)[None, :] * np.exp(
                -0.5 * np.sum((diff / sigmalambda) ** 2, axis=0)
            )  # (Ns,K)   :72
            q_j = np.sum(w[None, :] * norm_jl, axis=1)  # (Ns,)   :73
            lsum = np.sum((diff / sigmalambda**2) * (norm_jl * w[None, :])[None, :, :], axis=2)  # (D,Ns) :77-79
            with np.errstate(divide="ignore", invalid="ignore"):
                if gf[0]:
                    mu_grad[:, j] = w[j] * np.sum(lsum / q_j[None, :], axis=1) / Ns  # :82
                if gf[1]:
                    isum = np.sum(lsum * (eps * lam[:, None]), axis=0)  # :87
                    sigma_grad[j] = w[j] * np.sum(isum / q_j) / Ns  # :88
                if gf[2]:
                    lambda_grad = lambda_grad + np.sum(lsum * (w[j] * sigma[j] * eps / q_j[None, :]), axis=1) / Ns  # :93
                if gf[3]:
                    w_grad[j] = w_grad[j] - np.sum(np.log(q_j)) / Ns  # :97
                    w_grad = w_grad - w[j] * np.sum(norm_jl / q_j[:, None], axis=0) / Ns  # :100
    if gf[2]:
        lambda_grad = lambda_grad * lam  # :106-108
    dH = None
    if nargout > 1:
        if jacobian_flag and gf[1]:
            sigma_grad = sigma_grad * sigma
        if (not jacobian_flag) and gf[2]:
            lambda_grad = lambda_grad / lam
        if jacobian_flag and gf[3]:
            w_grad = _softmax_jacobian(vp["eta"]) @ w_grad
        dH = np.concatenate([mu_grad.T.ravel(), sigma_grad.ravel(), lambda_grad.ravel(), w_grad.ravel()])
    return H, dH


def entlb_vbmc(vp, grad_flags=None, jacobian_flag=True, nargout=2):
    """[H,dH] = entlb_vbmc(vp,grad_flags,jacobian_flag) — ent/entlb_vbmc.m:1-147, the deterministic entropy lower bound of
    Gershman et al. (2012) that negelcbo_vbmc uses when Ns == 0 (negelcbo_vbmc.m:102-109): vpsieve_vbmc.m:76 scores every
    candidate with it (NSentFast = 0, vbmc.m:216), and the whole optimisation does when K == 1 or EntropySwitch is on
    (vpsieve_vbmc.m:29-33).  Vectorised like the .m code (K x K temporaries); the BigK branch (:48-66) is dead (BigK = Inf)."""
    if nargout < 2:  # :8-9
        gf = [False] * 4
    elif grad_flags is None:  # :10-11
        gf = [True] * 4
    else:
        gf = list(np.atleast_1d(grad_flags).astype(bool))
        if len(gf) == 1:  # :13
            gf = gf * 4
    D, K = vp["D"], vp["K"]
    mu = np.asarray(vp["mu"], dtype=np.float64).reshape(D, K)
    sigma = np.asarray(vp["sigma"], dtype=np.float64).ravel()
    lam = np.asarray(vp["lambda"], dtype=np.float64).ravel()
    w = np.asarray(vp["w"], dtype=np.float64).ravel()
    mu_grad = np.zeros((D, K)) if gf[0] else np.zeros((0,))
    sigma_grad = np.zeros(K) if gf[1] else np.zeros((0,))
    lambda_grad = np.zeros(D) if gf[2] else np.zeros((0,))
    w_grad = np.zeros(K) if gf[3] else np.zeros((0,))
    if K == 1:  # :32-47 exact entropy of one Gaussian
        H = 0.5 * D * (1 + math.log(2 * math.pi)) + D * np.sum(np.log(sigma)) + np.sum(np.log(lam))
        if gf[1]:
            sigma_grad[:] = D / sigma
        if gf[2]:
            lambda_grad[:] = 1.0  # :42-43 (already the log-lambda gradient)
        if gf[3]:
            w_grad = np.zeros(1)  # :46
    else:  # :68-135
        # index convention of the .m code: dim 2 = j (row of gammasum), dim 3 = k
        sumsigma2 = sigma[:, None] ** 2 + sigma[None, :] ** 2  # (j,k)  :76
        sumsigma = np.sqrt(sumsigma2)
        nconst = 1.0 / (2 * math.pi) ** (D / 2) / np.prod(lam)  # :79
        dmu_jk = mu[:, :, None] - mu[:, None, :]  # (D,j,k) = mu_j - mu_k   (bsxfun(@minus, mu, mu_3), :81)
        d2 = np.sum((dmu_jk / (sumsigma[None, :, :] * lam[:, None, None])) ** 2, axis=0)  # (j,k)  :81
        gamma = nconst / sumsigma**D * np.exp(-0.5 * d2)  # (j,k)  :82
        gammasum = np.sum(w[:, None] * gamma, axis=0)  # (k,): sum over dim 2 (j) of w(1,j).*gamma(1,j,k)  :83 (gamma is symmetric)
        H = -np.sum(w * np.log(gammasum))  # :85
        if any(gf):
            # :90 divides gamma(1,j,k) by gammasum(1,1,k) (gammasum is 1x1xK): element (j,k) -> gamma(j,k)/gammasum(k)
            gammafrac = gamma / gammasum[None, :]
            wgammafrac = w[None, :] * gammafrac  # w_3(k) * gamma(j,k)/gammasum(k)  :91
            if gf[0]:
                dmu = (mu[:, None, :] - mu[:, :, None]) / (sumsigma2[None, :, :] * (lam**2)[:, None, None])  # (D,j,k) = (mu_k-mu_j)/..  :94
            if gf[1]:
                dsigma = -D / sumsigma2 + 1.0 / sumsigma2**2 * np.sum((dmu_jk / lam[:, None, None]) ** 2, axis=0)  # (j,k)  :97
            for j in range(K):  # :101-115
                if gf[0]:
                    m1 = np.sum(wgammafrac[j, :][None, :] * dmu[:, j, :], axis=1)  # :104
                    m2 = np.sum(dmu[:, j, :] * (gamma[j, :] * w)[None, :], axis=1) / gammasum[j]  # :105
                    mu_grad[:, j] = -w[j] * (m1 + m2)  # :106
                if gf[1]:
                    s1 = np.sum(wgammafrac[j, :] * dsigma[j, :])  # :111
                    s2 = np.sum(dsigma[j, :] * gamma[j, :] * w) / gammasum[j]  # :112
                    sigma_grad[j] = -w[j] * sigma[j] * (s1 + s2)  # :113
            if gf[2]:
                dmu2 = (mu[:, None, :] - mu[:, :, None]) ** 2 / (sumsigma2[None, :, :] * (lam**2)[:, None, None])  # (D,j,k)  :118
                inner = np.sum((dmu2 - 1.0) * (gamma * w[:, None])[None, :, :], axis=1)  # sum over j of (.)*gamma(j,k)*w(j) -> (D,k)  :120
                lambda_grad[:] = -np.sum(w[None, :] * inner / gammasum[None, :], axis=1)  # :119-121
            if gf[3]:
                w_grad[:] = -np.log(gammasum) - np.sum(wgammafrac, axis=1)  # :126  sum over dim 3 (k), result indexed by j
    dH = None
    if nargout > 1:  # :137-145
        if jacobian_flag and gf[1]:
            sigma_grad = sigma_grad * sigma
        if (not jacobian_flag) and gf[2]:
            lambda_grad = lambda_grad / lam
        if jacobian_flag and gf[3]:
            w_grad = _softmax_jacobian(vp["eta"]) @ np.atleast_1d(w_grad)
        dH = np.concatenate([mu_grad.T.ravel(), np.ravel(sigma_grad), np.ravel(lambda_grad), np.ravel(w_grad)])
    return H, dH


def entlb_K1(vp):
    """ent/entlb_vbmc.m:32-34: exact entropy of a single component (known answer for entmc)."""
    D = vp["D"]
    return 0.5 * D * (1 + math.log(2 * math.pi)) + D * np.sum(np.log(vp["sigma"])) + np.sum(np.log(vp["lambda"]))


# --------------------------------------------------------------------------------------
# expected log joint
# --------------------------------------------------------------------------------------
_SUPPORTED_MEANFUN_REF = (0, 1, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22)  # gplogjoint.m:47
_SUPPORTED_MEANFUN_BUILD = (0, 1, 4)  # scope of this build


def _solve_Kinv(L, Lchol, sn2_eff, B):
    """K^{-1} B as the reference does: (L\\(L'\\B))/sn2_eff or -L*B  (gplogjoint.m:276-280)."""
    if Lchol:
        t = sla.solve_triangular(L, B, trans="T", lower=False)
        return sla.solve_triangular(L, t, lower=False) / sn2_eff
    return -L @ B


def gplogjoint(vp, gp, grad_flags=None, avg_flag=True, jacobian_flag=True, compute_var=None,
               separate_K=None, nargout=2):
    """misc/gplogjoint.m:1-413 for meanfun in {0,1,4}.

    Returns (F, dF, varF, dvarF, varss, I_sk, J_sjk); entries not requested via
    ``nargout`` are None (mirrors MATLAB nargout-dependent defaults, :9-22).
    """
    if separate_K is None:
        separate_K = nargout > 5
    if compute_var is None:
        compute_var = nargout > 2
    compute_var = int(compute_var)
    if nargout < 2:
        grad_flags = False
    elif grad_flags is None:
        grad_flags = True
    if np.isscalar(grad_flags):
        grad_flags = [bool(grad_flags)] * 4
    gf = [bool(g) for g in grad_flags]
    compute_vargrad = nargout > 3 and compute_var and any(gf)
    if compute_vargrad and compute_var != 2:
        raise OracleError("gplogjoint:FullVarianceGradient",
                          "Computation of gradient of log joint variance is currently available only for diagonal approximation of the variance.")
    D, K = vp["D"], vp["K"]
    X = np.asarray(gp["X"], dtype=np.float64)
    N = X.shape[0]
    mu = np.asarray(vp["mu"], dtype=np.float64).reshape(D, K)
    sigma = np.asarray(vp["sigma"], dtype=np.float64).ravel()
    lam = np.asarray(vp["lambda"], dtype=np.float64).ravel()
    w = np.asarray(vp["w"], dtype=np.float64).ravel()
    Ncov, Nnoise = gp["Ncov"], gp["Nnoise"]
    Ns = len(gp["post"])
    meanfun = gp["meanfun"]
    if meanfun not in _SUPPORTED_MEANFUN_REF or meanfun not in _SUPPORTED_MEANFUN_BUILD:
        raise OracleError("gplogjoint:UnsupportedMeanFun",
                          "Log joint computation currently only supports zero, constant, negative quadratic mean functions in this build.")
    quadratic_meanfun = meanfun == 4

    F = np.zeros(Ns)
    mu_grad = np.zeros((D, K, Ns)) if gf[0] else None
    sigma_grad = np.zeros((K, Ns)) if gf[1] else None
    lambda_grad = np.zeros((D, Ns)) if gf[2] else None
    w_grad = np.zeros((K, Ns)) if gf[3] else None
    varF = np.zeros(Ns) if compute_var else None
    if compute_vargrad:
        mu_vargrad = np.zeros((D, K, Ns)) if gf[0] else None
        sigma_vargrad = np.zeros((K, Ns)) if gf[1] else None
        lambda_vargrad = np.zeros((D, Ns)) if gf[2] else None
        w_vargrad = np.zeros((K, Ns)) if gf[3] else None
    I_sk = np.zeros((Ns, K)) if separate_K else None
    J_sjk = np.zeros((Ns, K, K)) if (separate_K and compute_var) else None

    delta = vp.get("delta")
    if delta is None or (hasattr(delta, "__len__") and len(delta) == 0):
        delta = 0.0
    delta = np.asarray(delta, dtype=np.float64) * np.ones(D)

    Xt = mu[:, None, :] - X.T[:, :, None]  # (D,N,K)   :92-95

    for s in range(Ns):
        post = gp["post"][s]
        hyp = np.asarray(post["hyp"], dtype=np.float64).ravel()
        ell = np.exp(hyp[:D])
        ln_sf2 = 2 * hyp[D]
        sum_lnell = np.sum(hyp[:D])
        m0 = hyp[Ncov + Nnoise] if meanfun > 0 else 0.0
        if quadratic_meanfun:
            xm = hyp[Ncov + Nnoise + 1 : Ncov + Nnoise + 1 + D]
            omega = np.exp(hyp[Ncov + Nnoise + D + 1 : Ncov + Nnoise + 2 * D + 1])
        alpha = np.asarray(post["alpha"], dtype=np.float64).ravel()
        L = post.get("L")
        Lchol = post.get("Lchol", True)
        sn2_eff = 1.0 / np.asarray(post["sW"]).ravel()[0] ** 2

        for k in range(K):
            tau_k = np.sqrt(sigma[k] ** 2 * lam**2 + ell**2 + delta**2)
            lnnf_k = ln_sf2 + sum_lnell - np.sum(np.log(tau_k))
            delta_k = Xt[:, :, k] / tau_k[:, None]  # (D,N)
            z_k = np.exp(lnnf_k - 0.5 * np.sum(delta_k**2, axis=0))  # (N,)
            I_k = z_k @ alpha + m0
            if quadratic_meanfun:
                nu_k = -0.5 * np.sum(1.0 / omega**2 * (mu[:, k] ** 2 + sigma[k] ** 2 * lam**2
                                                     - 2 * mu[:, k] * xm + xm**2 + delta**2))
                I_k = I_k + nu_k
            F[s] += w[k] * I_k
            if separate_K:
                I_sk[s, k] = I_k
            if gf[0]:
                dz_dmu = -(delta_k / tau_k[:, None]) * z_k[None, :]
                mu_grad[:, k, s] = w[k] * (dz_dmu @ alpha)
                if quadratic_meanfun:
                    mu_grad[:, k, s] -= w[k] / omega**2 * (mu[:, k] - xm)
            if gf[1]:
                dz_dsigma = np.sum((lam / tau_k)[:, None] ** 2 * (delta_k**2 - 1), axis=0) * (sigma[k] * z_k)
                sigma_grad[k, s] = w[k] * (dz_dsigma @ alpha)
                if quadratic_meanfun:
                    sigma_grad[k, s] -= w[k] * sigma[k] * np.sum(1.0 / omega**2 * lam**2)
            if gf[2]:
                dz_dlambda = ((sigma[k] / tau_k) ** 2)[:, None] * (delta_k**2 - 1) * (lam[:, None] * z_k[None, :])
                lambda_grad[:, s] += w[k] * (dz_dlambda @ alpha)
                if quadratic_meanfun:
                    lambda_grad[:, s] -= w[k] * sigma[k] ** 2 / omega**2 * lam
            if gf[3]:
                w_grad[k, s] = I_k

            if compute_var == 2:  # :273-304
                tau_kk = np.sqrt(2 * sigma[k] ** 2 * lam**2 + ell**2 + 2 * delta**2)
                nf_kk = np.exp(ln_sf2 + sum_lnell - np.sum(np.log(tau_kk)))
                invKzk = _solve_Kinv(L, Lchol, sn2_eff, z_k)
                J_kk = nf_kk - z_k @ invKzk
                varF[s] += w[k] ** 2 * max(EPS, J_kk)
                if separate_K:
                    J_sjk[s, k, k] = J_kk
                if compute_vargrad:
                    if gf[0]:
                        mu_vargrad[:, k, s] = -w[k] ** 2 * (2 * dz_dmu @ invKzk)
                    if gf[1]:
                        sigma_vargrad[k, s] = -2 * w[k] ** 2 * (sigma[k] * nf_kk * np.sum(lam**2 / tau_kk**2) + dz_dsigma @ invKzk)
                    if gf[2]:
                        lambda_vargrad[:, s] -= 2 * w[k] ** 2 * (sigma[k] ** 2 * nf_kk * lam / tau_kk**2 + dz_dlambda @ invKzk)
                    if gf[3]:
                        w_vargrad[k, s] = 2 * w[k] * max(EPS, J_kk)
            elif compute_var:  # :306-339
                for j in range(k + 1):
                    tau_j = np.sqrt(sigma[j] ** 2 * lam**2 + ell**2 + delta**2)
                    lnnf_j = ln_sf2 + sum_lnell - np.sum(np.log(tau_j))
                    delta_j = (mu[:, j][:, None] - X.T) / tau_j[:, None]
                    z_j = np.exp(lnnf_j - 0.5 * np.sum(delta_j**2, axis=0))
                    tau_jk = np.sqrt((sigma[j] ** 2 + sigma[k] ** 2) * lam**2 + ell**2 + 2 * delta**2)
                    lnnf_jk = ln_sf2 + sum_lnell - np.sum(np.log(tau_jk))
                    delta_jk = (mu[:, j] - mu[:, k]) / tau_jk
                    J_jk = math.exp(lnnf_jk - 0.5 * np.sum(delta_jk**2)) - z_k @ _solve_Kinv(L, Lchol, sn2_eff, z_j)
                    if j == k:
                        varF[s] += w[k] ** 2 * max(EPS, J_jk)
                        if separate_K:
                            J_sjk[s, k, k] = J_jk
                    else:
                        varF[s] += 2 * w[j] * w[k] * J_jk
                        if separate_K:
                            J_sjk[s, j, k] = J_jk
                            J_sjk[s, k, j] = J_jk

    if compute_var:
        varF = np.maximum(varF, EPS)  # :350

    dF = None
    J_w = None
    if any(gf):
        parts = []
        if gf[0]:
            parts.append(mu_grad.transpose(1, 0, 2).reshape(D * K, Ns))  # reshape(mu_grad,[D*K,Ns])
        if gf[1]:
            parts.append(sigma_grad * sigma[:, None] if jacobian_flag else sigma_grad)
        if gf[2]:
            parts.append(lambda_grad * lam[:, None] if jacobian_flag else lambda_grad)
        if gf[3]:
            if jacobian_flag:
                J_w = _softmax_jacobian(vp["eta"])
                parts.append(J_w @ w_grad)
            else:
                parts.append(w_grad)
        dF = np.concatenate(parts, axis=0)  # (#theta, Ns)

    dvarF = None
    if compute_vargrad:
        parts = []
        if gf[0]:
            parts.append(mu_vargrad.transpose(1, 0, 2).reshape(D * K, Ns))
        if gf[1]:
            parts.append(sigma_vargrad * sigma[:, None] if jacobian_flag else sigma_vargrad)
        if gf[2]:
            parts.append(lambda_vargrad * lam[:, None] if jacobian_flag else lambda_vargrad)
        if gf[3]:
            parts.append(J_w @ w_vargrad if jacobian_flag else w_vargrad)
        dvarF = np.concatenate(parts, axis=0)

    varss = 0.0
    if Ns > 1 and avg_flag:  # :398-413
        Fbar = np.sum(F) / Ns
        if compute_var:
            varFss = np.sum((F - Fbar) ** 2) / (Ns - 1)
            varss = varFss + np.std(varF, ddof=1)
            varF = np.sum(varF) / Ns + varFss
        if compute_vargrad:
            dvv = 2 * np.sum(F[None, :] * dF, axis=1) / (Ns - 1) - 2 * Fbar * np.sum(dF, axis=1) / (Ns - 1)
            dvarF = np.sum(dvarF, axis=1) / Ns + dvv
        F = Fbar
        if any(gf):
            dF = np.sum(dF, axis=1) / Ns
    else:
        if Ns == 1:
            F = F[0]
            if dF is not None:
                dF = dF[:, 0]
            if compute_var:
                varF = varF[0]
            if dvarF is not None:
                dvarF = dvarF[:, 0]
    return F, dF, varF, dvarF, varss, I_sk, J_sjk


def gplite_quad(gp, mu, sigma, ssflag=False):
    """gplite/gplite_quad.m:38-107 (mean only): independent formulation of I_sk, used as KAT."""
    X = gp["X"]
    N, D = X.shape
    S = len(gp["post"])
    Ncov, Nnoise = gp["Ncov"], gp["Nnoise"]
    mu = np.atleast_2d(mu)
    Nstar = mu.shape[0]
    sigma = np.atleast_2d(sigma)
    if sigma.shape[0] == 1:
        sigma = np.tile(sigma, (Nstar, 1))
    F = np.zeros((Nstar, S))
    for s in range(S):
        hyp = gp["post"][s]["hyp"]
        ell = np.exp(hyp[:D])[None, :]
        ln_sf2 = 2 * hyp[D]
        sum_lnell = np.sum(hyp[:D])
        m0 = hyp[Ncov + Nnoise] if gp["meanfun"] > 0 else 0.0
        alpha = gp["post"][s]["alpha"]
        tau = np.sqrt(sigma**2 + ell**2)
        lnnf = ln_sf2 + sum_lnell - np.sum(np.log(tau), axis=1)
        sumdelta2 = np.zeros((Nstar, N))
        for i in range(D):
            sumdelta2 += ((mu[:, i : i + 1] - X[:, i][None, :]) / tau[:, i : i + 1]) ** 2
        z = np.exp(lnnf[:, None] - 0.5 * sumdelta2)
        F[:, s] = z @ alpha + m0
        if gp["meanfun"] == 4:
            xm = hyp[Ncov + Nnoise + 1 : Ncov + Nnoise + 1 + D][None, :]
            omega = np.exp(hyp[Ncov + Nnoise + D + 1 : Ncov + Nnoise + 2 * D + 1])[None, :]
            nu_k = -0.5 * np.sum(1.0 / omega**2 * (mu**2 + sigma**2 - 2 * mu * xm + xm**2), axis=1)
            F[:, s] += nu_k
    if S > 1 and not ssflag:
        return F.sum(axis=1) / S
    return F


# --------------------------------------------------------------------------------------
# negative ELCBO
# --------------------------------------------------------------------------------------
def negelcbo_vbmc(theta, beta, vp, gp, Ns=0, compute_grad=None, compute_var=None, altent_flag=False,
                  thetabnd=None, entropy_alpha=0, epsilon=None, nargout=2):
    """misc/negelcbo_vbmc.m:1-164 (MC entropy branch and full-parameter gplogjoint branch).

    Returns (F, dF, G, H, varF, dH, varGss, varG, varH, I_sk, J_sjk).
    """
    if compute_grad is None:
        compute_grad = nargout > 1
    if beta is None or not np.isfinite(beta):
        beta = 0.0
    if compute_var is None:
        compute_var = (beta != 0) or nargout > 4
    compute_var = int(compute_var)
    separate_K = nargout > 9
    if compute_grad and beta != 0 and compute_var != 2:
        raise OracleError("negelcbo_vbmc:vargrad",
                          "Computation of the gradient of ELBO with full variance not supported.")
    vp = dict(vp)
    D, K = vp["D"], vp["K"]
    theta = np.asarray(theta, dtype=np.float64).ravel()
    idx = 0
    if vp["optimize_mu"]:
        vp["mu"] = theta[: D * K].reshape(K, D).T.copy()
        idx = D * K
    if vp["optimize_sigma"]:
        vp["sigma"] = np.exp(theta[idx : idx + K])
        idx += K
    if vp["optimize_lambda"]:
        vp["lambda"] = np.exp(theta[idx : idx + D])
    if vp["optimize_weights"]:
        vp["eta"] = theta[-K:].copy()
        ww = np.exp(vp["eta"])
        vp["w"] = ww / ww.sum()
    gflags = [bool(compute_grad) and bool(vp[f]) for f in
              ("optimize_mu", "optimize_sigma", "optimize_lambda", "optimize_weights")]
    onlyweights = vp["optimize_weights"] and not (vp["optimize_mu"] or vp["optimize_sigma"] or vp["optimize_lambda"])
    if onlyweights:
        raise OracleError("vbmc_b200:OutOfScope", "gplogjoint_weights fast path is out of scope (SURVEY.md 2 #5)")
    dG = dvarG = None
    I_sk = J_sjk = None
    if separate_K:
        if compute_grad:
            raise OracleError("negelcbo_vbmc:separateKgrad",
                              "Computing the gradient of variational parameters and requesting per-component results at the same time.")
        if compute_var:
            G, _, varG, _, varGss, I_sk, J_sjk = gplogjoint(vp, gp, gflags, True, True, compute_var, nargout=7)
        else:
            G, dG, _, _, _, I_sk, _ = gplogjoint(vp, gp, gflags, True, True, 0, nargout=6)
            varGss, varG = 0.0, 0.0
    else:
        if compute_var:
            if compute_grad:
                G, dG, varG, dvarG, varGss, _, _ = gplogjoint(vp, gp, gflags, True, True, compute_var, nargout=5)
            else:
                G, _, varG, _, varGss, _, _ = gplogjoint(vp, gp, gflags, True, True, compute_var, nargout=5, separate_K=False)
        else:
            G, dG, _, _, _, _, _ = gplogjoint(vp, gp, gflags, True, True, 0, nargout=2)
            varGss, varG = 0.0, 0.0
    if Ns > 0:
        H, dH = entmc_vbmc(vp, Ns, gflags, True, epsilon=epsilon, nargout=2)
    else:
        H, dH = entlb_vbmc(vp, gflags, True, nargout=2)  # deterministic lower bound (negelcbo_vbmc.m:107-109)
    F = -G - H
    dF = (-dG - dH) if compute_grad else None
    varH = 0.0
    varF = (varG + varH) if compute_var else 0.0
    if beta != 0:
        F = F + beta * math.sqrt(varF)
        if compute_grad:
            dF = dF + 0.5 * beta * dvarG / math.sqrt(varF)
    if thetabnd is not None:
        L, dL = vpbndloss(theta, vp, thetabnd, thetabnd["TolCon"], compute_grad)
        if compute_grad:
            dF = dF + dL
        F = F + L
        if vp["optimize_weights"]:
            Thresh = thetabnd["WeightThreshold"]
            w = vp["w"]
            L = np.sum(w * (w < Thresh) + Thresh * (w >= Thresh)) * thetabnd["WeightPenalty"]
            F = F + L
            if compute_grad:
                wg = thetabnd["WeightPenalty"] * (w < Thresh).astype(np.float64)
                wg = _softmax_jacobian(vp["eta"]) @ wg
                dL = np.zeros_like(dF)
                dL[-K:] = wg
                dF = dF + dL
    return F, dF, G, H, varF, dH, varGss, varG, varH, I_sk, J_sjk


# --------------------------------------------------------------------------------------
# gplite: SE-ARD GP surrogate
# --------------------------------------------------------------------------------------
def sq_dist(a, b=None):
    """gplite/private/sq_dist.m:14-50.  a: (D,n), b: (D,m) -> (n,m)."""
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[1]
    if b is None:
        mu = a.mean(axis=1, keepdims=True)
        a = a - mu
        b = a
    else:
        b = np.asarray(b, dtype=np.float64)
        m = b.shape[1]
        mu = (m / (n + m)) * b.mean(axis=1, keepdims=True) + (n / (n + m)) * a.mean(axis=1, keepdims=True)
        a = a - mu
        b = b - mu
    C = np.sum(a * a, axis=0)[:, None] + (np.sum(b * b, axis=0)[None, :] - 2 * a.T @ b)
    return np.maximum(C, 0)


def gplite_covfun_info(D, covfun=1):
    """gplite_covfun('info'): SE-ARD has D+1 hyperparameters."""
    if covfun not in (1, [1], (1,)):
        raise OracleError("vbmc_b200:OutOfScope", "only covfun=1 (SE-ARD) is in scope")
    return D + 1


def gplite_meanfun_info(D, meanfun):
    """Number of mean-function hyper-parameters, gplite_meanfun.m:57-71 (cases 0,1,4)."""
    if meanfun == 0:
        return 0
    if meanfun == 1:
        return 1
    if meanfun == 4:
        return 1 + 2 * D
    raise OracleError("vbmc_b200:OutOfScope", f"meanfun {meanfun} out of scope (only 0,1,4)")


def gplite_noisefun_info(noisefun):
    """Number of noise hyper-parameters (gplite_noisefun.m info branch)."""
    n = 0
    if noisefun[0] == 1:
        n += 1
    if noisefun[1] == 2:
        n += 1
    if noisefun[2] == 1:
        n += 2
    return n


def gplite_meanfun(hyp, X, meanfun, compute_grad=False):
    """gplite/gplite_meanfun.m:400-436 for cases 0, 1, 4."""
    N, D = X.shape
    hyp = np.asarray(hyp, dtype=np.float64).ravel()
    dm = None
    if meanfun == 0:
        m = np.zeros(N)
        if compute_grad:
            dm = np.zeros((N, 0))
    elif meanfun == 1:
        m = hyp[0] * np.ones(N)
        if compute_grad:
            dm = np.ones((N, 1))
    elif meanfun == 4:
        m0 = hyp[0]
        xm = hyp[1 : 1 + D][None, :]
        omega = np.exp(hyp[D + 1 : 2 * D + 1])[None, :]
        z2 = ((X - xm) / omega) ** 2
        m = m0 - 0.5 * np.sum(z2, axis=1)
        if compute_grad:
            dm = np.zeros((N, 2 * D + 1))
            dm[:, 0] = 1
            dm[:, 1 : D + 1] = (X - xm) / omega**2
            dm[:, D + 1 : 2 * D + 1] = z2
    else:
        raise OracleError("vbmc_b200:OutOfScope", f"meanfun {meanfun} out of scope (only 0,1,4)")
    return (m, dm) if compute_grad else m


def gplite_noisefun(hyp, X, noisefun, y=None, s2=None, compute_grad=False):
    """gplite/gplite_noisefun.m:153-210."""
    N = X.shape[0]
    hyp = np.asarray(hyp, dtype=np.float64).ravel()
    Nnoise = gplite_noisefun_info(noisefun)
    dsn2 = None
    if compute_grad:
        dsn2 = np.zeros((N, Nnoise)) if any(n > 0 for n in noisefun[1:]) else np.zeros((1, Nnoise))
    idx = 0
    if noisefun[0] == 0:
        sn2 = EPS
    else:
        sn2 = math.exp(2 * hyp[idx])
        if compute_grad:
            dsn2[:, idx] = 2 * sn2
        idx += 1
    if noisefun[1] == 1:
        sn2 = sn2 + np.asarray(s2, dtype=np.float64).ravel()
    elif noisefun[1] == 2:
        s2 = np.asarray(s2, dtype=np.float64).ravel()
        sn2 = sn2 + math.exp(hyp[idx]) * s2
        if compute_grad:
            dsn2[:, idx] = math.exp(hyp[idx]) * s2
        idx += 1
    if noisefun[2] == 1:
        if y is not None:
            ythresh = hyp[idx]
            w2 = math.exp(2 * hyp[idx + 1])
            zz = np.maximum(0, ythresh - y)
            sn2 = sn2 + w2 * zz**2
            if compute_grad:
                dsn2[:, idx] = 2 * w2 * (ythresh - y) * (zz > 0)
                dsn2[:, idx + 1] = 2 * w2 * zz**2
        idx += 2
    return (sn2, dsn2) if compute_grad else sn2


def _chol_upper(A):
    """MATLAB [L,p] = chol(A): upper factor, p>0 when not positive definite."""
    try:
        return sla.cholesky(A, lower=False, check_finite=False), 0
    except np.linalg.LinAlgError:
        return None, 1


def gplite_core(hyp, gp, compute_nlZ=False, compute_nlZ_grad=False, want_post=True):
    """gplite/private/gplite_core.m:6-11,33-102,193,226-261,278-285 (no intmeanfun / outwarp).

    Returns (nlZ, dnlZ, post, K_mat, Q).
    """
    X = np.asarray(gp["X"], dtype=np.float64)
    N, D = X.shape
    Ncov, Nnoise, Nmean = gp["Ncov"], gp["Nnoise"], gp["Nmean"]
    hyp = np.asarray(hyp, dtype=np.float64).ravel()
    y = np.asarray(gp["y"], dtype=np.float64).ravel()
    s2 = gp.get("s2")
    hyp_noise = hyp[Ncov : Ncov + Nnoise]
    if compute_nlZ_grad:
        sn2, dsn2 = gplite_noisefun(hyp_noise, X, gp["noisefun"], y, s2, True)
    else:
        sn2 = gplite_noisefun(hyp_noise, X, gp["noisefun"], y, s2)
    sn2_mult = 1.0
    hyp_mean = hyp[Ncov + Nnoise : Ncov + Nnoise + Nmean]
    if compute_nlZ_grad:
        m, dm = gplite_meanfun(hyp_mean, X, gp["meanfun"], True)
    else:
        m = gplite_meanfun(hyp_mean, X, gp["meanfun"])
    ell = np.exp(hyp[:D])
    sf2 = math.exp(2 * hyp[D])
    K_mat = sq_dist(X.T / ell[:, None])
    K_mat = sf2 * np.exp(-K_mat / 2)

    scalar_sn2 = np.isscalar(sn2) or np.ndim(sn2) == 0
    Lchol = bool(np.min(sn2) >= 1e-6)
    if Lchol:
        if scalar_sn2:
            sn2div = float(sn2)
            sn2_mat = np.eye(N)
        else:
            sn2div = float(np.min(sn2))
            sn2_mat = np.diag(sn2 / sn2div)
        for _ in range(10):
            L, p = _chol_upper(K_mat / (sn2div * sn2_mult) + sn2_mat)
            if p > 0:
                sn2_mult *= 10
            else:
                break
        sl = sn2div * sn2_mult
        pL = L
    else:
        sn2_mat = (float(sn2) * np.eye(N)) if scalar_sn2 else np.diag(sn2)
        for _ in range(10):
            L, p = _chol_upper(K_mat + sn2_mult * sn2_mat)
            if p > 0:
                sn2_mult *= 10
            else:
                break
        sl = 1.0
        pL = -sla.solve_triangular(L, sla.solve_triangular(L, np.eye(N), trans="T", lower=False), lower=False)
    if L is None:
        raise OracleError("vbmc_b200:CholFailed", "Cholesky failed after 10 jitter retries (MATLAB would error on L\\)")
    alpha = sla.solve_triangular(L, sla.solve_triangular(L, y - m, trans="T", lower=False), lower=False) / sl

    nlZ = dnlZ = Q = None
    if compute_nlZ:
        Nhyp = hyp.shape[0]
        nlZ = (y - m) @ alpha / 2 + np.sum(np.log(np.diag(L))) + N * math.log(2 * math.pi * sl) / 2
        if compute_nlZ_grad:
            dnlZ = np.zeros(Nhyp)
            Q = sla.solve_triangular(L, sla.solve_triangular(L, np.eye(N), trans="T", lower=False), lower=False) / sl \
                - np.outer(alpha, alpha)
            for i in range(D):
                K_temp = K_mat * sq_dist(X[:, i][None, :] / ell[i])
                dnlZ[i] = np.sum(Q * K_temp) / 2
            dnlZ[D] = np.sum(Q * (2 * K_mat)) / 2
            if scalar_sn2:
                trQ = np.trace(Q)
                for i in range(Nnoise):
                    dnlZ[Ncov + i] = 0.5 * sn2_mult * dsn2[0, i] * trQ
            else:
                dgQ = np.diag(Q)
                for i in range(Nnoise):
                    dnlZ[Ncov + i] = 0.5 * sn2_mult * np.sum(dsn2[:, i] * dgQ)
            if Nmean > 0:
                dnlZ[Ncov + Nnoise : Ncov + Nnoise + Nmean] = -dm.T @ alpha
    post = None
    if want_post:
        post = {
            "hyp": hyp.copy(),
            "alpha": alpha,
            "sW": np.ones(N) / math.sqrt(np.min(sn2) * sn2_mult),
            "L": pL,
            "sn2_mult": sn2_mult,
            "Lchol": Lchol,
        }
    return nlZ, dnlZ, post, K_mat, Q


def gplite_hypprior(hyp, hprior, compute_grad=False):
    """gplite/gplite_hypprior.m:18-58."""
    hyp = np.asarray(hyp, dtype=np.float64).ravel()
    Nhyp = hyp.shape[0]
    mu = np.asarray(hprior["mu"], dtype=np.float64).ravel()
    sigma = np.abs(np.asarray(hprior["sigma"], dtype=np.float64).ravel())
    df = hprior.get("df")
    df = 7 * np.ones(Nhyp) if df is None or len(df) == 0 else np.asarray(df, dtype=np.float64).ravel()
    uidx = ~np.isfinite(mu) | ~np.isfinite(sigma)
    gidx = ~uidx & ((df == 0) | ~np.isfinite(df)) & np.isfinite(sigma)
    tidx = ~uidx & (df > 0) & np.isfinite(df)
    z2 = np.zeros(Nhyp)
    gt = gidx | tidx
    z2[gt] = ((hyp[gt] - mu[gt]) / sigma[gt]) ** 2
    lp = 0.0
    dlp = np.zeros(Nhyp)
    if gidx.any():
        lp -= 0.5 * np.sum(np.log(2 * math.pi * sigma[gidx] ** 2) + z2[gidx])
        dlp[gidx] = -(hyp[gidx] - mu[gidx]) / sigma[gidx] ** 2
    if tidx.any():
        lp += np.sum(gammaln(0.5 * (df[tidx] + 1)) - gammaln(0.5 * df[tidx]) - 0.5 * np.log(math.pi * df[tidx])
                     - np.log(sigma[tidx]) - 0.5 * (df[tidx] + 1) * np.log1p(z2[tidx] / df[tidx]))
        dlp[tidx] = -(df[tidx] + 1) / df[tidx] / (1 + z2[tidx] / df[tidx]) * (hyp[tidx] - mu[tidx]) / sigma[tidx] ** 2
    return (lp, dlp) if compute_grad else lp


def gplite_nlZ(hyp, gp, hprior=None, nargout=1):
    """gplite/gplite_nlZ.m:27-66.  Returns (nlZ, dnlZ, post, K_mat, Q)."""
    hyp = np.asarray(hyp, dtype=np.float64)
    if hyp.ndim == 1:
        hyp = hyp[:, None]
    Nhyp, Ns = hyp.shape
    compute_grad = nargout > 1
    if Nhyp != gp["Ncov"] + gp["Nnoise"] + gp["Nmean"]:
        raise OracleError("gplite_nlZ:dimmismatch",
                          "Number of hyperparameters mismatched with dimension of training inputs.")
    if compute_grad and Ns > 1:
        raise OracleError("gplite_nlZ:NoSampling",
                          "Computation of the log marginal likelihood is available only for one-sample hyperparameter inputs.")
    nlZ, dnlZ, post, K_mat, Q = gplite_core(hyp[:, 0], gp, True, compute_grad, want_post=nargout > 2)
    if hprior is not None:
        if compute_grad:
            P, dP = gplite_hypprior(hyp[:, 0], hprior, True)
            nlZ = nlZ - P
            dnlZ = dnlZ - dP
        else:
            nlZ = nlZ - gplite_hypprior(hyp[:, 0], hprior)
    return nlZ, dnlZ, post, K_mat, Q


def gplite_post(hyp, X, y, covfun=None, meanfun=None, noisefun=None, s2=None):
    """gplite/gplite_post.m:94-172 (create + full refit; rank-1 branch :173-251 is a 'next' row)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    N, D = X.shape
    hyp = np.asarray(hyp, dtype=np.float64)
    if hyp.ndim == 1:
        hyp = hyp[:, None]
    Nhyp, S = hyp.shape
    if covfun is None:
        covfun = 1
    if meanfun is None:
        meanfun = 1
    if noisefun is None:
        noisefun = [1, 0, 0] if s2 is None else [1, 1, 0]
    gp = {
        "X": X, "y": y, "s2": None if s2 is None else np.asarray(s2, dtype=np.float64).ravel(),
        "Ncov": gplite_covfun_info(D, covfun), "covfun": covfun,
        "Nnoise": gplite_noisefun_info(noisefun), "noisefun": list(noisefun),
        "Nmean": gplite_meanfun_info(D, meanfun), "meanfun": meanfun, "meanfun_extras": None,
        "intmeanfun": 0,
    }
    if Nhyp != gp["Ncov"] + gp["Nnoise"] + gp["Nmean"]:
        raise OracleError("gplite_post:dimmismatch",
                          "Number of hyperparameters mismatched with GP model specification.")
    gp["post"] = []
    for s in range(S):
        _, _, post, _, _ = gplite_core(hyp[:, s], gp, False, False)
        gp["post"].append(post)
    return gp


def gplite_pred_mean(gp, Xstar):
    """gplite/gplite_pred.m:52-87 (posterior mean only), used to sanity check alpha."""
    X = gp["X"]
    N, D = X.shape
    out = np.zeros((Xstar.shape[0], len(gp["post"])))
    for s, post in enumerate(gp["post"]):
        hyp = post["hyp"]
        ell = np.exp(hyp[:D])
        sf2 = math.exp(2 * hyp[D])
        Ks = sf2 * np.exp(-sq_dist(X.T / ell[:, None], Xstar.T / ell[:, None]) / 2)
        mstar = gplite_meanfun(hyp[gp["Ncov"] + gp["Nnoise"] :], Xstar, gp["meanfun"])
        out[:, s] = mstar + Ks.T @ post["alpha"]
    return out


# --------------------------------------------------------------------------------------
# fminadam update (the definition of a "grad-step", utils/fminadam.m:42-60)
# --------------------------------------------------------------------------------------
class AdamState:
    """State for the modified Adam of utils/fminadam.m:20-60 (update only, no termination)."""

    def __init__(self, nvars, step_max=0.1, step_min=0.001, decay=200.0):
        self.m = np.zeros(nvars)
        self.v = np.zeros(nvars)
        self.iter = 0
        self.step_max, self.step_min, self.decay = step_max, step_min, decay
        self.beta1, self.beta2 = 0.9, 0.999
        self.fudge = math.sqrt(EPS)

    def update(self, x, grad):
        self.iter += 1
        it = self.iter
        self.m = self.beta1 * self.m + (1 - self.beta1) * grad
        self.v = self.beta2 * self.v + (1 - self.beta2) * grad**2
        mhat = self.m / (1 - self.beta1**it)
        vhat = self.v / (1 - self.beta2**it)
        stepsize = self.step_min + (self.step_max - self.step_min) * math.exp(-it / self.decay)
        return x - stepsize * mhat / (np.sqrt(vhat) + self.fudge)


def fminadam(fun, x0, LB=None, UB=None, TolFun=None, MaxIter=None, master_stepsize=None):
    """[x,f,xtab,ftab,iter] = fminadam(fun,x0,LB,UB,TolFun,MaxIter,master_stepsize), utils/fminadam.m:1-102.

    ``fun(x) -> (f, grad)``.  Returns x (mean of the last 20 iterates, :95), f (mean of the last 20 values, :96),
    xtab (nvars, iter), ftab (iter,), iter.  The slope test follows polyfit's own algebra (QR of the Vandermonde
    matrix, S.R / S.df / S.normr, :69-73).
    """
    TolFun = 0.001 if TolFun is None else TolFun            # :6
    MaxIter = 10000 if MaxIter is None else int(MaxIter)    # :7
    ms = {"max": 0.1, "min": 0.001, "decay": 200.0}         # :11-18
    for k, v in (master_stepsize or {}).items():
        if v is not None:
            ms[k] = v
    fudge = math.sqrt(EPS)                                  # :21
    beta1, beta2, batchsize = 0.9, 0.999, 20                # :22-24
    TolX, TolX_max, TolFun_max = 0.001, 0.1, TolFun * 100   # :25-27
    MinIter = batchsize * 2                                 # :29
    x = np.asarray(x0, dtype=np.float64).ravel().copy()
    nvars = x.size
    LB = np.full(nvars, -np.inf) if LB is None else np.asarray(LB, dtype=np.float64).ravel()
    UB = np.full(nvars, np.inf) if UB is None else np.asarray(UB, dtype=np.float64).ravel()
    m = np.zeros(nvars)
    v = np.zeros(nvars)
    xtab = np.zeros((nvars, MaxIter))
    ftab = np.full(MaxIter, np.nan)
    it = 0
    for it in range(1, MaxIter + 1):
        f, grad = fun(x)                                    # :48
        ftab[it - 1] = f
        grad = np.asarray(grad, dtype=np.float64).ravel()
        m = beta1 * m + (1 - beta1) * grad                  # :51
        v = beta2 * v + (1 - beta2) * grad**2               # :52
        mhat = m / (1 - beta1**it)
        vhat = v / (1 - beta2**it)
        stepsize = ms["min"] + (ms["max"] - ms["min"]) * math.exp(-it / ms["decay"])   # :56-57
        x = x - stepsize * mhat / (np.sqrt(vhat) + fudge)   # :59
        x = np.minimum(np.maximum(x, LB), UB)               # :60
        xtab[:, it - 1] = x                                 # :63
        if it % batchsize == 0 and it >= MinIter:           # :65
            xxp = np.linspace(-(batchsize - 1) / 2, (batchsize - 1) / 2, batchsize)
            y = ftab[it - batchsize:it]
            V = np.stack([xxp, np.ones(batchsize)], axis=1)  # polyfit(x,y,1): Vandermonde, QR, p = R\(Q'y)
            Q, R = np.linalg.qr(V)
            p = np.linalg.solve(R, Q.T @ y)
            normr = np.linalg.norm(y - V @ p)
            df = batchsize - 2
            Rinv = np.linalg.inv(R)
            A = (Rinv @ Rinv.T) * normr**2 / df             # :71
            slope = p[0]
            slope_err = math.sqrt(A[0, 0] + TolFun**2)      # :72
            slope_err_max = math.sqrt(A[0, 0] + TolFun_max**2)
            cur = xtab[:, it - batchsize:it].mean(axis=1)
            prev = xtab[:, it - 2 * batchsize:it - batchsize].mean(axis=1)
            dx = math.sqrt(np.sum((cur - prev) ** 2 / batchsize))   # :77
            if (dx < TolX and abs(slope) < slope_err_max) or (abs(slope) < slope_err and dx < TolX_max):   # :80
                break
    x = xtab[:, it - batchsize:it].mean(axis=1)             # :95
    f = float(np.mean(ftab[it - batchsize:it]))             # :96
    return x, f, xtab[:, :it].copy(), ftab[:it].copy(), it


def gplite_pred(gp, Xstar, ystar=None, s2star=None, ssflag=False, nowarpflag=False, nargout=2):
    """[ymu,ys2,fmu,fs2,lp] = gplite_pred(gp,Xstar,ystar,s2star,ssflag,nowarpflag), gplite/gplite_pred.m:1-163
    (SE-ARD covfun, no integrated mean function, no output warping)."""
    X = gp["X"]
    N, D = X.shape
    Ns = len(gp["post"])
    Xstar = np.atleast_2d(np.asarray(Xstar, dtype=np.float64))
    Nstar = Xstar.shape[0]
    if ystar is not None and np.size(ystar) != Nstar:
        raise OracleError("gplite_pred:ydimmismatch", "YSTAR should be empty or a column vector of NSTAR observations.")
    if s2star is not None and np.size(s2star) != Nstar:
        raise OracleError("gplite_pred:s2dimmismatch", "S2STAR should be empty or a column vector of NSTAR estimated variances.")
    fmu = np.zeros((Nstar, Ns))
    ymu = np.zeros((Nstar, Ns))
    fs2 = np.zeros((Nstar, Ns))
    ys2 = np.zeros((Nstar, Ns))
    lp = np.zeros((Nstar, Ns)) if (ystar is not None and nargout > 4) else None
    Ncov, Nnoise, Nmean = gp["Ncov"], gp["Nnoise"], gp["Nmean"]
    for s, post in enumerate(gp["post"]):
        hyp = np.asarray(post["hyp"], dtype=np.float64).ravel()
        alpha, L, Lchol = post["alpha"], post["L"], post["Lchol"]
        sW = np.asarray(post["sW"], dtype=np.float64).ravel()
        sn2_mult = post.get("sn2_mult", 1.0)
        sn2_star = gplite_noisefun(hyp[Ncov:Ncov + Nnoise], Xstar, gp["noisefun"], ystar, s2star)   # :60-61
        if isinstance(sn2_star, tuple):
            sn2_star = sn2_star[0]
        mstar = gplite_meanfun(hyp[Ncov + Nnoise:Ncov + Nnoise + Nmean], Xstar, gp["meanfun"])      # :64-65
        if isinstance(mstar, tuple):
            mstar = mstar[0]
        ell = np.exp(hyp[:D])
        sf2 = math.exp(2 * hyp[D])
        Ks = sf2 * np.exp(-sq_dist(X.T / ell[:, None], Xstar.T / ell[:, None]) / 2)                 # :68-71
        kss = sf2 * np.ones(Nstar)                                                                    # :72
        fmu[:, s] = mstar + Ks.T @ alpha                                                              # :80
        ymu[:, s] = fmu[:, s]                                                                         # :92
        if nargout > 1:
            if Lchol:
                V = sla.solve_triangular(L, sW[:, None] * Ks, trans="T", lower=False)                 # :97
                fs2[:, s] = kss - np.sum(V * V, axis=0)                                               # :98
            else:
                fs2[:, s] = kss + np.sum(Ks * (L @ Ks), axis=0)                                       # :100-101
            fs2[:, s] = np.maximum(fs2[:, s], 0)                                                      # :118
            ys2[:, s] = fs2[:, s] + sn2_star * sn2_mult                                               # :119
            if lp is not None:
                lp[:, s] = -0.5 * (np.ravel(ystar) - ymu[:, s]) ** 2 / ys2[:, s] - 0.5 * np.log(2 * math.pi * ys2[:, s])   # :124
    if Ns > 1 and not ssflag:                                                                         # :153-163
        fbar = fmu.sum(axis=1) / Ns
        ybar = ymu.sum(axis=1) / Ns
        if nargout > 1:
            vf = ((fmu - fbar[:, None]) ** 2).sum(axis=1) / (Ns - 1)
            fs2 = fs2.sum(axis=1) / Ns + vf
            vy = ((ymu - ybar[:, None]) ** 2).sum(axis=1) / (Ns - 1)
            ys2 = ys2.sum(axis=1) / Ns + vy
        fmu, ymu = fbar, ybar
    elif Ns == 1:
        pass
    out = (ymu, ys2 if nargout > 1 else None, fmu, fs2 if nargout > 1 else None, lp)
    return out[:max(1, nargout)]


def gplite_post_update1(gp, xstar, ystar, s2star=None):
    """gp = gplite_post(gp,xstar,ystar,[],[],[],s2star,1): rank-one update, gplite/gplite_post.m:50-92,173-251.
    With s2star the reference falls back to the standard update with the enlarged training set (:78-91)."""
    xstar = np.atleast_2d(np.asarray(xstar, dtype=np.float64))
    if xstar.shape[0] > 1:
        raise OracleError("gplite_post:NotRankOne", "GPLITE_POST with this input format only supports rank-one updates.")
    if gp is None:
        raise OracleError("gplite_post:NoGP", "GPLITE_POST can perform rank-one update only with an existing GP struct.")
    ystar = float(np.ravel(ystar)[0])
    D = gp["X"].shape[1]
    if s2star is not None or gp.get("intmeanfun", 0):
        hyp = np.stack([p["hyp"] for p in gp["post"]], axis=1)
        s2 = None if s2star is None else np.append(np.ravel(gp["s2"]), np.ravel(s2star))
        return gplite_post(hyp, np.vstack([gp["X"], xstar]), np.append(gp["y"], ystar), gp["covfun"], gp["meanfun"], gp["noisefun"], s2)
    Ncov, Nnoise = gp["Ncov"], gp["Nnoise"]
    mstar, vstar = gplite_pred(gp, xstar, np.array([ystar]), None, True, True, nargout=2)     # :191
    new = dict(gp)
    new["post"] = []
    for s, post in enumerate(gp["post"]):
        hyp = np.asarray(post["hyp"], dtype=np.float64).ravel()
        sn2 = gplite_noisefun(hyp[Ncov:Ncov + Nnoise], xstar, gp["noisefun"], np.array([ystar]), None)   # :207-208
        if isinstance(sn2, tuple):
            sn2 = sn2[0]
        sn2 = float(np.ravel(sn2)[0])
        sn2_eff = sn2 * post["sn2_mult"]                                                      # :209
        ell = np.exp(hyp[:D])
        sf2 = math.exp(2 * hyp[D])
        K = sf2
        Ks = sf2 * np.exp(-sq_dist(gp["X"].T / ell[:, None], xstar.T / ell[:, None]) / 2)   # :215-217 (N x 1)
        L, Lchol = post["L"], post["Lchol"]
        if Lchol:                                                                             # :227-233
            c = sla.solve_triangular(L, Ks, trans="T", lower=False) / sn2_eff
            alpha_update = sla.solve_triangular(L, sla.solve_triangular(L, Ks, trans="T", lower=False), lower=False) / sn2_eff
            n = L.shape[0]
            Lnew = np.zeros((n + 1, n + 1))
            Lnew[:n, :n] = L
            Lnew[:n, n] = c[:, 0]
            Lnew[n, n] = math.sqrt(1 + K / sn2_eff - float(c[:, 0] @ c[:, 0]))
        else:                                                                                 # :234-238
            alpha_update = -L @ Ks
            v = -alpha_update / vstar[0, s]
            n = L.shape[0]
            Lnew = np.zeros((n + 1, n + 1))
            Lnew[:n, :n] = L + v @ alpha_update.T
            Lnew[:n, n] = -v[:, 0]
            Lnew[n, :n] = -v[:, 0]
            Lnew[n, n] = -1 / vstar[0, s]
        p = dict(post)
        p["L"] = Lnew
        p["sW"] = np.append(np.ravel(post["sW"]), 1 / math.sqrt(sn2_eff))                     # :242
        p["alpha"] = np.append(post["alpha"], 0.0) + (mstar[0, s] - ystar) / vstar[0, s] * np.append(alpha_update[:, 0], -1.0)   # :245-247
        new["post"].append(p)
    new["X"] = np.vstack([gp["X"], xstar])                                                    # :251-253
    new["y"] = np.append(gp["y"], ystar)
    return new
