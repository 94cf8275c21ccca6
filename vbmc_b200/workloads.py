"""Synthetic workloads of the shapes BASELINE.json names.

Inputs only: targets (Rosenbrock / "lumpy" mixture), training sets, GP hyper-parameter samples,
variational posteriors and entropy draws, all from fixed NumPy ``Philox`` seeds.  The GP posterior
itself (alpha, L) is NOT computed here: callers pass ``gp_post`` — the product's
``vbmc_b200.gplite_post`` (bench, GPU tests) or the oracle's (CPU tests).

    c1  D=2  N=50   K=2   Ns=100     S=8    rosenbrock_test.m plumbing case (CPU)
    c2  D=6  N=400  K=20  Ns=4096    S=8    Rosenbrock
    c3  D=10 N=2000 K=50  Ns=32768   S=20   lumpy   <- the configuration BASELINE's metric is quoted on
    c4  D=10 N=2000 K=50  Ns=131072  S=20   lumpy, MC shard across 2/4/8 GPUs
    c5  D=20 N=4000 K=100 Ns=262144  S=40   lumpy + noise (noisefun [1 1 0]); FP32 target, later round
"""
from __future__ import annotations

import math

import numpy as np

CONFIGS = {
    "c1": dict(D=2, N=50, K=2, Ns=100, S=8, target="rosenbrock", noisy=False),
    "c2": dict(D=6, N=400, K=20, Ns=4096, S=8, target="rosenbrock", noisy=False),
    "c3": dict(D=10, N=2000, K=50, Ns=32768, S=20, target="lumpy", noisy=False),
    "c4": dict(D=10, N=2000, K=50, Ns=131072, S=20, target="lumpy", noisy=False),
    "c5": dict(D=20, N=4000, K=100, Ns=262144, S=40, target="lumpy", noisy=True),
}

# vbmc.m defaults that define thetabnd: TolLength (:252), TolWeight (:303), TolConLoss (:300), WeightPenalty (:205)
VP_OPTIONS = dict(TolLength=1e-6, TolWeight=1e-2, TolConLoss=0.01, WeightPenalty=0.1)


def _rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def rosenbrock_logpost(X):
    """rosenbrock_test.m:7 log-likelihood + N(0,3^2 I) prior (vbmc_examples.m:43-50)."""
    X = np.atleast_2d(X)
    ll = -np.sum((X[:, :-1] ** 2 - X[:, 1:]) ** 2 + (X[:, :-1] - 1) ** 2 / 100.0, axis=1)
    lp = -0.5 * np.sum(X**2, axis=1) / 9.0 - 0.5 * X.shape[1] * math.log(2 * math.pi * 9.0)
    return ll + lp


class Lumpy:
    """y = log sum_i pi_i N(x; m_i, s_i^2 I), 12 bumps (the README figure's 'lumpy' posterior is not
    defined in the reference; this is the definition of SURVEY.md §8d, seed 300)."""

    def __init__(self, D, seed=300, nbumps=12):
        r = _rng(seed)
        self.D = D
        self.m = r.uniform(-3, 3, size=(nbumps, D))
        self.s = r.uniform(0.5, 1.5, size=nbumps)
        self.pi = r.dirichlet(np.ones(nbumps))

    def logpdf(self, X):
        X = np.atleast_2d(X)
        d2 = ((X[:, None, :] - self.m[None, :, :]) ** 2).sum(axis=2) / self.s[None, :] ** 2
        la = np.log(self.pi)[None, :] - 0.5 * d2 - self.D * np.log(self.s)[None, :] - 0.5 * self.D * math.log(2 * math.pi)
        mx = la.max(axis=1, keepdims=True)
        return (mx + np.log(np.exp(la - mx).sum(axis=1, keepdims=True))).ravel()

    def sample(self, n, rng, widen=1.5):
        idx = rng.choice(len(self.pi), size=n, p=self.pi)
        return self.m[idx] + widen * self.s[idx, None] * rng.standard_normal((n, self.D))


def make_training_set(cfg, seed=101):
    D, N = cfg["D"], cfg["N"]
    r = _rng(seed)
    if cfg["target"] == "rosenbrock":
        X = 1.5 * r.standard_normal((N, D))
        y = rosenbrock_logpost(X)
    else:
        t = Lumpy(D)
        X = t.sample(N, r)
        y = t.logpdf(X)
    s2 = None
    if cfg.get("noisy"):
        y = y + r.standard_normal(N)
        s2 = np.ones(N)
    return X, y, s2


def make_hyp_samples(cfg, X, y, seed=102, log_sn=None):
    """Nhyp x S hyper-parameter samples: [log ell (D); log sf; log sn; m0; xm (D); log omega (D)]
    (SE-ARD covfun 1, noisefun [1 0 0] / [1 1 0], negquad meanfun 4 — setupvars_vbmc.m:276-281, vbmc.m:245)."""
    D, S = cfg["D"], cfg["S"]
    r = _rng(seed)
    rng_d = X.max(axis=0) - X.min(axis=0)
    hyp = np.zeros((3 * D + 3, S))
    for s in range(S):
        hyp[:D, s] = np.log(0.5 * rng_d) + 0.2 * r.standard_normal(D)
        hyp[D, s] = math.log(np.std(y)) + 0.1 * r.standard_normal()
        hyp[D + 1, s] = 0.5 * math.log(1e-5) + 0.1 * abs(r.standard_normal())  # >= TolGPNoise (vbmc.m:307)
        if log_sn is not None:
            hyp[D + 1, s] = log_sn + 0.1 * abs(r.standard_normal())
        hyp[D + 2, s] = np.max(y) + 0.1 * r.standard_normal()
        hyp[D + 3 : 2 * D + 3, s] = X.mean(axis=0) + 0.05 * r.standard_normal(D)
        hyp[2 * D + 3 :, s] = np.log(2 * X.std(axis=0)) + 0.05 * r.standard_normal(D)
    return hyp


def make_vp(cfg, X, y, seed=103, sigma_scale=0.35):
    """K components centred on the K best training points (+ jitter), log-normal sigma, lambda = 1
    rescaled like rescale_params.m, softmax weights; all optimize_* flags on."""
    D, K = cfg["D"], cfg["K"]
    r = _rng(seed)
    order = np.argsort(-y)
    mu = X[order[:K]].T + 0.1 * r.standard_normal((D, K))
    sigma = sigma_scale * np.exp(0.3 * r.standard_normal(K))
    lam = np.exp(0.2 * r.standard_normal(D))
    nl = math.sqrt(np.sum(lam**2) / D)
    lam, sigma = lam / nl, sigma * nl
    eta = 0.3 * r.standard_normal(K)
    w = np.exp(eta - eta.max())
    w /= w.sum()
    return {"D": D, "K": K, "mu": mu, "sigma": sigma, "lambda": lam, "w": w, "eta": np.log(w), "delta": None,
            "optimize_mu": True, "optimize_sigma": True, "optimize_lambda": True, "optimize_weights": True}


def make_epsilon(cfg, seed=104, Ns=None):
    Ns = cfg["Ns"] if Ns is None else Ns
    Ns = int(math.ceil(Ns / 2) * 2)
    return _rng(seed).standard_normal((cfg["K"], Ns // 2, cfg["D"]))


def theta_of(vp):
    """[mu(:); log sigma; log lambda; eta] (negelcbo_vbmc.m:32-48 layout)."""
    return np.concatenate([np.asarray(vp["mu"]).T.ravel(), np.log(vp["sigma"]), np.log(vp["lambda"]), np.asarray(vp["eta"])])


def build(name_or_cfg, gp_post, *, Ns=None, seeds=(101, 102, 103, 104), with_eps=True, overrides=None):
    """Return dict(cfg, X, y, s2, hyp, gp, vp, theta, epsilon).  ``gp_post(hyp,X,y,covfun,meanfun,noisefun,s2)``
    computes the GP posterior (product or oracle implementation)."""
    cfg = dict(CONFIGS[name_or_cfg]) if isinstance(name_or_cfg, str) else dict(name_or_cfg)
    if overrides:
        cfg.update(overrides)
    if Ns is not None:
        cfg["Ns"] = Ns
    X, y, s2 = make_training_set(cfg, seeds[0])
    hyp = make_hyp_samples(cfg, X, y, seeds[1], log_sn=cfg.get("log_sn"))
    noisefun = [1, 1, 0] if s2 is not None else [1, 0, 0]
    gp = gp_post(hyp, X, y, 1, 4, noisefun, s2)
    vp = make_vp(cfg, X, y, seeds[2])
    out = dict(cfg=cfg, X=X, y=y, s2=s2, hyp=hyp, gp=gp, vp=vp, theta=theta_of(vp))
    out["epsilon"] = make_epsilon(cfg, seeds[3]) if with_eps else None
    return out
