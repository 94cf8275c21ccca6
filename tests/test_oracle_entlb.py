"""Oracle restatement of ent/entlb_vbmc.m (deterministic entropy lower bound; negelcbo_vbmc with Ns == 0,
negelcbo_vbmc.m:102-109 — what misc/vpsieve_vbmc.m:76 evaluates for every candidate).  Pins: the K = 1 closed form, an
independent scalar-loop evaluation of the bound, central finite differences (the bound is deterministic, so FD is exact
up to truncation), Jensen's inequality against a large Monte-Carlo entropy, and the Jacobian identities."""
import math

import numpy as np
import pytest

from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads


def mk_vp(D, K, seed):
    cfg = dict(D=D, N=30, K=K, S=1, Ns=2, target="rosenbrock", noisy=False)
    w = workloads.build(cfg, orc.gplite_post, seeds=(seed, seed + 1, seed + 2, seed + 3))
    return w["vp"], w["gp"], w["theta"]


def scalar_entlb(vp):
    """H_lb = -sum_k w_k log sum_j w_j N(mu_k; mu_j, (sigma_j^2 + sigma_k^2) diag(lambda^2))   (Gershman et al. 2012)."""
    D, K = vp["D"], vp["K"]
    mu, sig, lam, w = np.asarray(vp["mu"]).reshape(D, K), np.ravel(vp["sigma"]), np.ravel(vp["lambda"]), np.ravel(vp["w"])
    H = 0.0
    for k in range(K):
        acc = 0.0
        for j in range(K):
            s2 = sig[j] ** 2 + sig[k] ** 2
            q = sum((mu[d, k] - mu[d, j]) ** 2 / (s2 * lam[d] ** 2) for d in range(D))
            acc += w[j] * math.exp(-0.5 * q) / ((2 * math.pi) ** (D / 2) * s2 ** (D / 2) * np.prod(lam))
        H -= w[k] * math.log(acc)
    return H


@pytest.mark.parametrize("D,K", [(2, 2), (3, 5), (6, 9)])
def test_entlb_value_matches_scalar_definition(D, K):
    vp, _, _ = mk_vp(D, K, 11)
    H, _ = orc.entlb_vbmc(vp, nargout=1)
    assert abs(H - scalar_entlb(vp)) < 1e-12 * max(1.0, abs(H))


def test_entlb_single_component_is_exact():
    vp, _, _ = mk_vp(4, 1, 21)
    H, dH = orc.entlb_vbmc(vp, [1, 1, 1, 1], True)
    D = 4
    assert abs(H - orc.entlb_K1(vp)) < 1e-14
    # d/dmu = 0; d/dlog(sigma) = D (entlb_vbmc.m:38-39 times sigma, :139-141); d/dlog(lambda) = 1 (:42-43); d/deta = 0
    assert np.allclose(dH, np.concatenate([np.zeros(D), [D], np.ones(D), [0.0]]), atol=1e-14)


@pytest.mark.parametrize("D,K", [(2, 3), (4, 6)])
def test_entlb_gradient_matches_finite_differences(D, K):
    """dH wrt theta = [mu(:); log sigma; log lambda; eta] with the Jacobians on (the parameterisation negelcbo uses)."""
    vp, _, theta = mk_vp(D, K, 31)
    vp = dict(vp, optimize_weights=True)
    theta = workloads.theta_of(vp)

    def unpack(t):
        v = dict(vp)
        v["mu"] = t[:D * K].reshape(K, D).T.copy()
        v["sigma"] = np.exp(t[D * K:D * K + K])
        v["lambda"] = np.exp(t[D * K + K:D * K + K + D])
        v["eta"] = t[-K:].copy()
        e = np.exp(v["eta"])
        v["w"] = e / e.sum()
        return v

    H, dH = orc.entlb_vbmc(unpack(theta), [1, 1, 1, 1], True)
    fd = np.zeros_like(theta)
    for i in range(theta.size):
        h = 1e-6
        tp, tm = theta.copy(), theta.copy()
        tp[i] += h
        tm[i] -= h
        fd[i] = (orc.entlb_vbmc(unpack(tp), nargout=1)[0] - orc.entlb_vbmc(unpack(tm), nargout=1)[0]) / (2 * h)
    assert np.max(np.abs(dH - fd)) < 2e-7 * max(1.0, np.max(np.abs(fd)))
    assert abs(np.sum(dH[-K:])) < 1e-12          # J_w rows sum to zero
    # without the Jacobians: sigma block divided by sigma, lambda block divided by lambda (:138-145)
    _, dH0 = orc.entlb_vbmc(unpack(theta), [1, 1, 1, 0], False)
    v = unpack(theta)
    assert np.allclose(dH0[D * K:D * K + K] * v["sigma"], dH[D * K:D * K + K], rtol=1e-12)
    assert np.allclose(dH0[D * K + K:D * K + K + D] * v["lambda"], dH[D * K + K:D * K + K + D], rtol=1e-12)


def test_entlb_is_a_lower_bound_of_the_mc_entropy():
    vp, _, _ = mk_vp(3, 4, 41)
    Hlb, _ = orc.entlb_vbmc(vp, nargout=1)
    eps = np.random.default_rng(0).standard_normal((4, 20000, 3))
    Hmc, _ = orc.entmc_vbmc(vp, 40000, [0, 0, 0, 0], True, epsilon=eps, nargout=1)
    assert Hlb <= Hmc + 0.02


def test_negelcbo_with_Ns_zero_uses_the_bound():
    vp, gp, theta = mk_vp(3, 4, 51)
    _, tb = orc.vpbounds(vp, gp, workloads.VP_OPTIONS)
    F0, dF0, G0, H0, _, dH0 = orc.negelcbo_vbmc(theta, 0.0, vp, gp, 0, 1, 0, 0, tb, 0, nargout=6)[:6]
    eps = workloads.make_epsilon(dict(D=3, K=4, Ns=64))
    F1, dF1, G1, H1, _, dH1 = orc.negelcbo_vbmc(theta, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)[:6]
    assert G0 == G1                                   # the entropy term is the only difference
    assert abs((F0 + H0) - (F1 + H1)) < 1e-12 * max(1.0, abs(F1))
    assert np.allclose(dF0 + dH0, dF1 + dH1, rtol=0, atol=1e-11 * max(1.0, np.max(np.abs(dF1))))
