"""SURVEY.md 8(c) pin 4: the entropy gradient of entmc_vbmc is a reparameterisation ESTIMATOR (it drops the zero-mean score
term), so finite differences of H at fixed draws do not reproduce it; it is checked instead against a 40-digit (mpmath)
evaluation of the same estimator formulas (ent/entmc_vbmc.m:55-125), written as scalar loops from the maths, not from the
oracle's vectorised code.  The expected log-joint and its gradient (misc/gplogjoint.m:164-252, meanfun 4) get the same treatment."""
import math

import mpmath as mp
import numpy as np

from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads

mp.mp.dps = 40


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def problem(D, K, N, S, Ns, seed):
    cfg = dict(D=D, N=N, K=K, S=S, Ns=Ns, target="rosenbrock", noisy=False)
    return workloads.build(cfg, orc.gplite_post, seeds=(seed, seed + 1, seed + 2, seed + 3))


def softmax_jacobian_apply(eta, g):
    e = [mp.e ** x for x in eta]
    es = sum(e)
    dot = sum(ei * gi for ei, gi in zip(e, g))
    return [ei / es * gi - ei / es ** 2 * dot for ei, gi in zip(e, g)]


def test_entmc_estimator_in_40_digits():
    D, K, Ns = 2, 3, 8
    w = problem(D, K, 30, 1, Ns, 71)
    vp, eps = w["vp"], w["epsilon"]
    mu = [[mp.mpf(float(vp["mu"][d, k])) for k in range(K)] for d in range(D)]
    sig = [mp.mpf(float(x)) for x in np.ravel(vp["sigma"])]
    lam = [mp.mpf(float(x)) for x in np.ravel(vp["lambda"])]
    wt = [mp.mpf(float(x)) for x in np.ravel(vp["w"])]
    eta = [mp.mpf(float(x)) for x in np.ravel(vp["eta"])]
    nf = 1 / (2 * mp.pi) ** (mp.mpf(D) / 2) / mp.fprod(lam)
    H = mp.mpf(0)
    gmu = [[mp.mpf(0)] * K for _ in range(D)]
    gsig = [mp.mpf(0)] * K
    glam = [mp.mpf(0)] * D
    gw = [mp.mpf(0)] * K
    for j in range(K):
        draws = [[mp.mpf(float(eps[j, s, d])) for d in range(D)] for s in range(Ns // 2)]
        draws = draws + [[-e for e in row] for row in draws]                               # antithetic (:53-54)
        for e in draws:
            x = [mu[d][j] + sig[j] * lam[d] * e[d] for d in range(D)]                        # :55
            Nk = [nf / sig[k] ** D * mp.e ** (-sum(((x[d] - mu[d][k]) / (sig[k] * lam[d])) ** 2 for d in range(D)) / 2) for k in range(K)]
            q = sum(wt[k] * Nk[k] for k in range(K))                                         # :60-65
            H -= wt[j] * mp.log(q) / Ns                                                      # :67
            lsum = [sum(wt[k] * Nk[k] * (x[d] - mu[d][k]) / (sig[k] * lam[d]) ** 2 for k in range(K)) for d in range(D)]   # :77-79
            for d in range(D):
                gmu[d][j] += wt[j] * lsum[d] / q / Ns                                         # :82
                glam[d] += wt[j] * sig[j] * lsum[d] * e[d] / q / Ns                           # :93
            gsig[j] += wt[j] * sum(lsum[d] * e[d] * lam[d] for d in range(D)) / q / Ns        # :87-88
            gw[j] -= mp.log(q) / Ns                                                          # :97
            for l in range(K):
                gw[l] -= wt[j] * Nk[l] / q / Ns                                               # :100
    glam = [glam[d] * lam[d] for d in range(D)]                                              # :106-108
    gsig = [gsig[k] * sig[k] for k in range(K)]                                              # :112-114 (jacobian)
    gw = softmax_jacobian_apply(eta, gw)                                                     # :120-124
    ref = [float(gmu[d][k]) for k in range(K) for d in range(D)] + [float(v) for v in gsig + glam + gw]
    Ho, dHo = orc.entmc_vbmc(vp, Ns, [1, 1, 1, 1], True, epsilon=eps, nargout=2)
    assert abs(Ho - float(H)) < 1e-14 * max(1.0, abs(float(H)))
    assert rel(dHo, ref) < 1e-13


def test_gplogjoint_in_40_digits():
    D, K, N = 2, 2, 12
    w = problem(D, K, N, 1, 4, 81)
    vp, gp = w["vp"], w["gp"]
    post = gp["post"][0]
    h = [mp.mpf(float(x)) for x in post["hyp"]]
    X = [[mp.mpf(float(v)) for v in row] for row in gp["X"]]
    alpha = [mp.mpf(float(v)) for v in post["alpha"]]
    mu = [[mp.mpf(float(vp["mu"][d, k])) for k in range(K)] for d in range(D)]
    sig = [mp.mpf(float(x)) for x in np.ravel(vp["sigma"])]
    lam = [mp.mpf(float(x)) for x in np.ravel(vp["lambda"])]
    wt = [mp.mpf(float(x)) for x in np.ravel(vp["w"])]
    eta = [mp.mpf(float(x)) for x in np.ravel(vp["eta"])]
    ell = [mp.e ** h[d] for d in range(D)]
    ln_sf2 = 2 * h[D]
    # hyp layout: [log ell (D); log sf; log sn (1); m0; xm (D); log omega (D)]  (gplite_meanfun case 4)
    m0 = h[D + 2]
    xm = h[D + 3:D + 3 + D]
    om = [mp.e ** v for v in h[D + 3 + D:D + 3 + 2 * D]]
    G = mp.mpf(0)
    gmu = [[mp.mpf(0)] * K for _ in range(D)]
    gsig = [mp.mpf(0)] * K
    glam = [mp.mpf(0)] * D
    I = [mp.mpf(0)] * K
    for k in range(K):
        tau = [mp.sqrt(sig[k] ** 2 * lam[d] ** 2 + ell[d] ** 2) for d in range(D)]          # :164
        lnnf = ln_sf2 + sum(h[:D]) - sum(mp.log(t) for t in tau)                             # :165
        za = []
        for n in range(N):
            dl = [(mu[d][k] - X[n][d]) / tau[d] for d in range(D)]
            za.append((mp.e ** (lnnf - sum(v * v for v in dl) / 2) * alpha[n], dl))          # :166-169
        I[k] = sum(z for z, _ in za) + m0 - sum((mu[d][k] ** 2 + sig[k] ** 2 * lam[d] ** 2 - 2 * mu[d][k] * xm[d] + xm[d] ** 2) / om[d] ** 2
                                                for d in range(D)) / 2                         # :169-174
        G += wt[k] * I[k]
        for d in range(D):
            Bd = sum(z * dl[d] for z, dl in za) / tau[d]
            Cd = sum(z * (dl[d] ** 2 - 1) for z, dl in za)
            gmu[d][k] = wt[k] * (-Bd - (mu[d][k] - xm[d]) / om[d] ** 2)                       # :206-210
            gsig[k] += wt[k] * sig[k] * ((lam[d] / tau[d]) ** 2 * Cd - lam[d] ** 2 / om[d] ** 2)   # :227-231
            glam[d] += wt[k] * sig[k] ** 2 * lam[d] * (Cd / tau[d] ** 2 - 1 / om[d] ** 2)     # :248-252
    gsig = [gsig[k] * sig[k] for k in range(K)]                                              # :357-359
    glam = [glam[d] * lam[d] for d in range(D)]                                              # :361-363
    gw = softmax_jacobian_apply(eta, I)                                                      # :269-271, :365-369
    ref = [float(gmu[d][k]) for k in range(K) for d in range(D)] + [float(v) for v in gsig + glam + gw]
    Go, dGo = orc.gplogjoint(vp, gp, [1, 1, 1, 1], True, True, 0, nargout=2)[:2]
    assert abs(Go - float(G)) < 1e-12 * max(1.0, abs(float(G)))
    assert rel(dGo, ref) < 1e-11


def test_gplite_core_in_40_digits():
    """nlZ = 1/2 r'(K + sn2 I)^-1 r + 1/2 log det(K + sn2 I) + N/2 log 2 pi and alpha = (K + sn2 I)^-1 r with the SE-ARD kernel and
    the negative-quadratic mean (gplite/private/gplite_core.m:33-102,193), evaluated with mpmath matrices."""
    D, N = 2, 9
    w = problem(D, 2, N, 1, 4, 91)
    gp = w["gp"]
    X, y, hyp = gp["X"], gp["y"], np.asarray(gp["post"][0]["hyp"], dtype=float)
    h = [mp.mpf(float(v)) for v in hyp]
    ell = [mp.e ** h[d] for d in range(D)]
    sf2, sn2 = mp.e ** (2 * h[D]), mp.e ** (2 * h[D + 1])
    m0, xm, om = h[D + 2], h[D + 3:D + 3 + D], [mp.e ** v for v in h[D + 3 + D:D + 3 + 2 * D]]
    Xm = [[mp.mpf(float(v)) for v in row] for row in X]
    K = mp.matrix(N, N)
    for i in range(N):
        for j in range(N):
            K[i, j] = sf2 * mp.e ** (-sum(((Xm[i][d] - Xm[j][d]) / ell[d]) ** 2 for d in range(D)) / 2) + (sn2 if i == j else 0)
    r = mp.matrix([mp.mpf(float(y[i])) - (m0 - sum(((Xm[i][d] - xm[d]) / om[d]) ** 2 for d in range(D)) / 2) for i in range(N)])
    alpha = mp.lu_solve(K, r)
    nlZ = (r.T * alpha)[0] / 2 + mp.log(mp.det(K)) / 2 + N * mp.log(2 * mp.pi) / 2
    nlZo = orc.gplite_nlZ(hyp, gp, None, nargout=1)[0]
    assert abs(nlZo - float(nlZ)) < 1e-9 * max(1.0, abs(float(nlZ)))
    ao = np.asarray(gp["post"][0]["alpha"], dtype=float)
    ref = np.array([float(v) for v in alpha])
    assert np.max(np.abs(ao - ref)) < 1e-7 * np.max(np.abs(ref))      # cond(K + sn2 I) ~ sf2/sn2
