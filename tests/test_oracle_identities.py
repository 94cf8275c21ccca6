"""Pins for the CPU oracle: the reference has no golden vectors for this path and
MATLAB cannot run here ("parity unpinned"), so the oracle is checked against analytic identities
and independent formulations inside the reference itself."""
import math

import numpy as np
import pytest
import scipy.stats

from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads


def small_workload(D=3, N=30, K=4, S=3, Ns=40, seed=0, target="rosenbrock", log_sn=None):
    cfg = dict(D=D, N=N, K=K, S=S, Ns=Ns, target=target, noisy=False, log_sn=log_sn)
    return workloads.build(cfg, orc.gplite_post, seeds=(seed + 1, seed + 2, seed + 3, seed + 4))


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


# ---- entmc ------------------------------------------------------------------------------------------
def test_entmc_K1_closed_form():
    """K=1: H_mc = D/2 log 2pi + D log sigma + sum log lambda + mean||eps||^2/2 exactly (entmc_vbmc.m:60-67),
    mu_grad == 0 exactly by antithetic symmetry."""
    rng = np.random.default_rng(1)
    D, Ns = 5, 200
    vp = dict(D=D, K=1, mu=rng.standard_normal((D, 1)), sigma=np.array([0.7]), w=np.array([1.0]), eta=np.array([0.0]),
              optimize_mu=True, optimize_sigma=True, optimize_lambda=True, optimize_weights=True)
    vp["lambda"] = np.exp(0.3 * rng.standard_normal(D))
    eps = rng.standard_normal((1, Ns // 2, D))
    H, dH = orc.entmc_vbmc(vp, Ns, True, True, epsilon=eps)
    expect = 0.5 * D * math.log(2 * math.pi) + D * math.log(0.7) + np.sum(np.log(vp["lambda"])) + 0.5 * np.mean(eps**2) * D
    assert abs(H - expect) < 1e-12 * abs(expect)
    assert np.max(np.abs(dH[:D])) < 1e-13  # mu_grad
    # sigma_grad (log-sigma Jacobian applied): d/dlog(sigma) of the estimator = mean ||eps||^2 ... check vs closed form
    # lsum/q = (xi-mu)/(sigma lambda)^2 = eps/(sigma lambda); isum = ||eps||^2/sigma; times sigma => mean ||eps||^2
    assert abs(dH[D] - np.mean(np.sum(eps**2, axis=2))) < 1e-12 * abs(dH[D])
    # converges to the exact entropy (entlb_vbmc.m:34)
    big = rng.standard_normal((1, 200000, D))
    Hbig, _ = orc.entmc_vbmc(vp, 400000, False, True, epsilon=big, nargout=1)
    assert abs(Hbig - orc.entlb_K1(vp)) < 2e-2


def test_entmc_weight_gradient_matches_direct_sum():
    """dH wrt w before the softmax Jacobian: -mean log q_j - sum_j w_j mean N_l/q (entmc_vbmc.m:97-100);
    and J_w rows sum to zero => the eta-gradient sums to ~0."""
    w = small_workload(D=2, K=3, Ns=60)
    vp, eps = w["vp"], w["epsilon"]
    H, dH = orc.entmc_vbmc(vp, 60, [0, 0, 0, 1], True, epsilon=eps)
    assert abs(np.sum(dH)) < 1e-12 * np.max(np.abs(dH))


def test_entmc_matches_loop_restatement():
    """The vectorised oracle equals a scalar triple loop written straight from the formulas."""
    w = small_workload(D=2, K=3, Ns=20)
    vp, eps = w["vp"], w["epsilon"]
    D, K, Ns = 2, 3, 20
    H, dH = orc.entmc_vbmc(vp, Ns, True, False, epsilon=eps)  # jacobian_flag = False
    mu, sg, lam, ww = vp["mu"], vp["sigma"], vp["lambda"], vp["w"]
    nf = 1 / (2 * math.pi) ** (D / 2) / np.prod(lam)
    Hl, mug, sgg, lmg, wg = 0.0, np.zeros((D, K)), np.zeros(K), np.zeros(D), np.zeros(K)
    for j in range(K):
        for s in range(Ns):
            e = eps[j, s % (Ns // 2)] * (1 if s < Ns // 2 else -1)
            x = mu[:, j] + sg[j] * lam * e
            Nl = np.array([nf / sg[k] ** D * math.exp(-0.5 * np.sum(((x - mu[:, k]) / (sg[k] * lam)) ** 2)) for k in range(K)])
            q = np.sum(ww * Nl)
            Hl -= ww[j] * math.log(q) / Ns
            lsum = sum(ww[k] * Nl[k] * (x - mu[:, k]) / (sg[k] * lam) ** 2 for k in range(K))
            mug[:, j] += ww[j] * lsum / q / Ns
            sgg[j] += ww[j] * np.sum(lsum * e * lam) / q / Ns
            lmg += ww[j] * sg[j] * lsum * e / q / Ns
            wg[j] -= math.log(q) / Ns
            wg -= ww[j] * Nl / q / Ns
    # jacobian_flag=False: sigma raw, lambda_grad*lambda/lambda = raw accumulation, w raw
    ref = np.concatenate([mug.T.ravel(), sgg, lmg, wg])
    assert abs(H - Hl) < 1e-13 * abs(Hl)
    assert rel(dH, ref) < 1e-12


# ---- gplogjoint -------------------------------------------------------------------------------------
def test_gplogjoint_Isk_equals_gplite_quad():
    """I_sk == gplite_quad(gp, mu_k', (sigma_k lambda)', 1): an independent formulation inside the reference
    (gplite/gplite_quad.m:70-82)."""
    w = small_workload(D=3, K=4, S=3)
    vp, gp = w["vp"], w["gp"]
    out = orc.gplogjoint(vp, gp, False, True, True, 0, nargout=6)
    I_sk = out[5]
    Fq = orc.gplite_quad(gp, vp["mu"].T, (vp["sigma"][:, None] * vp["lambda"][None, :]), ssflag=True)  # (K,S)
    # two summation orders of z*alpha with |alpha| ~ 1e4 (sn2 = 1e-5): agreement is limited by cancellation
    assert rel(I_sk, Fq.T) < 1e-10
    assert abs(out[0] - np.mean(I_sk @ vp["w"])) < 1e-12 * abs(out[0])


def _fd(fun, theta, h=1e-6):
    g = np.zeros_like(theta)
    for i in range(theta.size):
        tp, tm = theta.copy(), theta.copy()
        tp[i] += h
        tm[i] -= h
        g[i] = (fun(tp) - fun(tm)) / (2 * h)
    return g


def test_gplogjoint_gradient_finite_differences():
    """Central FD of G(theta) vs the analytic dG, through negelcbo's theta unpacking (exact for G)."""
    # log_sn = log 0.1: with sn2 = 1e-5 alpha ~ 1e4 and cancellation noise in G (~1e-11 rel) swamps the FD quotient
    w = small_workload(D=2, K=3, S=2, N=25, log_sn=math.log(0.1))
    vp, gp, theta = w["vp"], w["gp"], w["theta"]

    def G_of(t):
        v = dict(vp)
        D, K = v["D"], v["K"]
        v["mu"] = t[: D * K].reshape(K, D).T
        v["sigma"] = np.exp(t[D * K : D * K + K])
        v["lambda"] = np.exp(t[D * K + K : D * K + K + D])
        v["eta"] = t[-K:]
        e = np.exp(v["eta"])
        v["w"] = e / e.sum()
        return orc.gplogjoint(v, gp, False, True, True, 0, nargout=1)[0]

    v = dict(vp)
    v["eta"] = theta[-vp["K"]:]
    G, dG = orc.gplogjoint(v, gp, True, True, True, 0, nargout=2)[:2]
    fd = _fd(G_of, theta, h=1e-5)
    assert rel(dG, fd) < 2e-6


def test_negelcbo_penalty_gradient_and_eta_sum():
    """FD check of the soft-bound + weight penalties (exact, deterministic part of F) and of the assembled F with
    the entropy term held fixed via fixed epsilon for the parts that are exact derivatives (mu through G + L)."""
    w = small_workload(D=2, K=3, S=2, N=25, Ns=30)
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"].copy(), w["epsilon"]
    _, tb = orc.vpbounds(vp, gp, workloads.VP_OPTIONS)
    theta[0] = tb["ub"][0] + 0.3           # violate a mu bound
    theta[2 * 3] = tb["lb"][2 * 3 + 0] - 25  # tiny sigma: violates lnscale lower bounds
    theta[-1] = 0.5                        # eta above its upper bound 0
    L, dL = orc.vpbndloss(theta, vp, tb, tb["TolCon"])
    fd = _fd(lambda t: orc.vpbndloss(t, vp, tb, tb["TolCon"], False)[0], theta, h=1e-5)
    assert L > 0 and rel(dL, fd) < 1e-6


def test_negelcbo_error_ids():
    w = small_workload(D=2, K=2, S=2, N=20, Ns=10)
    with pytest.raises(orc.OracleError) as ei:
        orc.negelcbo_vbmc(w["theta"], 1.0, w["vp"], w["gp"], 10, 1, 1, epsilon=w["epsilon"])
    assert ei.value.identifier == "negelcbo_vbmc:vargrad"
    gp = dict(w["gp"], meanfun=7)
    with pytest.raises(orc.OracleError) as ei:
        orc.gplogjoint(w["vp"], gp)
    assert ei.value.identifier == "gplogjoint:UnsupportedMeanFun"


def test_gplogjoint_variance_paths_consistent():
    """Full variance (compute_var=1) diagonal terms equal the diagonal approximation's J_kk (gplogjoint.m:273-339)."""
    w = small_workload(D=2, K=3, S=2, N=25)
    vp, gp = w["vp"], w["gp"]
    full = orc.gplogjoint(vp, gp, False, True, True, 1, nargout=7)
    diag = orc.gplogjoint(vp, gp, False, True, True, 2, nargout=7)
    Jf, Jd = full[6], diag[6]
    for k in range(vp["K"]):
        assert rel(Jf[:, k, k], Jd[:, k, k]) < 1e-9
    assert full[2] > 0 and diag[2] > 0


# ---- gplite -----------------------------------------------------------------------------------------
def test_sq_dist_matches_direct():
    rng = np.random.default_rng(3)
    a = rng.standard_normal((4, 17)) * 3 + 10
    C = orc.sq_dist(a)
    direct = ((a[:, :, None] - a[:, None, :]) ** 2).sum(axis=0)
    assert np.max(np.abs(C - direct)) < 1e-11
    assert np.all(np.diag(C) == 0) or np.max(np.abs(np.diag(C))) < 1e-12
    b = rng.standard_normal((4, 5))
    assert np.max(np.abs(orc.sq_dist(a, b) - ((a[:, :, None] - b[:, None, :]) ** 2).sum(axis=0))) < 1e-11


@pytest.mark.parametrize("meanfun,noisy", [(4, False), (1, False), (0, False), (4, True)])
def test_gplite_core_identities(meanfun, noisy):
    """L'L == K/sl + diag, (K+Sigma) alpha == y-m, nlZ == -log N(y; m, K+Sigma) (scipy), dnlZ vs FD
    (what gplite_test.m:70-75 checks with an external gradest)."""
    rng = np.random.default_rng(5)
    N, D = 25, 2
    X = rng.standard_normal((N, D))
    y = -np.sum(X**2, axis=1) + 0.1 * rng.standard_normal(N)
    s2 = 0.05 * np.ones(N) * (1 + rng.random(N)) if noisy else None
    noisefun = [1, 1, 0] if noisy else [1, 0, 0]
    Nmean = {0: 0, 1: 1, 4: 1 + 2 * D}[meanfun]
    hyp = np.concatenate([np.log([0.8, 1.2]), [math.log(1.5)], [math.log(0.1)],
                          {0: [], 1: [0.3], 4: [0.3, 0.1, -0.2, math.log(1.5), math.log(2.0)]}[meanfun]])
    assert hyp.size == D + 1 + 1 + Nmean
    gp = orc.gplite_post(hyp, X, y, 1, meanfun, noisefun, s2)
    post = gp["post"][0]
    nlZ, dnlZ, _, K_mat, _ = orc.gplite_core(hyp, gp, True, True)
    sn2 = orc.gplite_noisefun(hyp[D + 1 : D + 2], X, noisefun, y, s2)
    Sigma = np.diag(np.broadcast_to(sn2, (N,)))
    sl = (float(np.min(sn2))) * post["sn2_mult"]
    A = K_mat / sl + Sigma / float(np.min(sn2))
    assert post["Lchol"]
    assert np.max(np.abs(post["L"].T @ post["L"] - A)) < 1e-10 * np.max(np.abs(A))
    m = orc.gplite_meanfun(hyp[D + 2 :], X, meanfun)
    assert np.max(np.abs((K_mat + Sigma) @ post["alpha"] - (y - m))) < 1e-8
    ref = -scipy.stats.multivariate_normal(mean=m, cov=K_mat + Sigma).logpdf(y)
    assert abs(nlZ - ref) < 1e-9 * abs(ref)
    fd = _fd(lambda h: orc.gplite_core(h, gp, True, False)[0], hyp, h=1e-6)
    assert rel(dnlZ, fd) < 2e-6
    # posterior mean interpolates the data up to the noise level
    pm = orc.gplite_pred_mean(gp, X)[:, 0]
    assert np.max(np.abs(pm - y)) < 1.0


def test_gplite_core_low_noise_branch_and_retry():
    """min(sn2) < 1e-6 takes the explicit-inverse branch: post.L == -inv(K+Sigma) (gplite_core.m:86-100)."""
    rng = np.random.default_rng(6)
    N, D = 12, 2
    X = rng.standard_normal((N, D)) * 2
    y = rng.standard_normal(N)
    hyp = np.array([0.0, 0.0, 0.0, math.log(3e-4), 0.1])
    gp = orc.gplite_post(hyp, X, y, 1, 1, [1, 0, 0], None)
    post = gp["post"][0]
    assert not post["Lchol"]
    K = np.exp(-0.5 * orc.sq_dist(X.T))
    A = K + post["sn2_mult"] * math.exp(2 * hyp[3]) * np.eye(N)
    assert np.max(np.abs(post["L"] @ A + np.eye(N))) < 1e-6


def test_hypprior_gaussian_and_student():
    hyp = np.array([0.2, -0.5, 1.0, 0.3])
    hp = dict(mu=np.array([0.0, 0.0, np.nan, 0.5]), sigma=np.array([1.0, 2.0, 1.0, np.inf]), df=np.array([0.0, 3.0, 3.0, 3.0]))
    lp, dlp = orc.gplite_hypprior(hyp, hp, True)
    expect = scipy.stats.norm(0, 1).logpdf(0.2) + scipy.stats.t(3, loc=0, scale=2).logpdf(-0.5)
    assert abs(lp - expect) < 1e-12
    fd = _fd(lambda h: orc.gplite_hypprior(h, hp), hyp)
    assert np.max(np.abs(dlp - fd)) < 1e-8


def test_rescale_and_theta_roundtrip():
    w = small_workload()
    theta, vp = orc.get_vptheta(w["vp"])
    assert abs(np.mean(vp["lambda"] ** 2) - 1) < 1e-14 and abs(vp["w"].sum() - 1) < 1e-14
    vp2 = orc.rescale_params(dict(vp), theta)
    assert rel(vp2["mu"], vp["mu"]) < 1e-15 and rel(vp2["sigma"], vp["sigma"]) < 1e-14
