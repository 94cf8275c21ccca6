"""Size-independent properties of the hot path that hold for the reference's formulas whatever the implementation:
(1) relabelling the mixture components permutes dF and leaves F, G, H unchanged (every sum over k is symmetric);
(2) the entropy estimator uses each draw with both signs (ent/entmc_vbmc.m:53-54), so feeding -eps changes nothing.
CPU: the NumPy oracle and the C port.  GPU: tests/test_zz_golden_gpu.py runs the same checks through the CUDA path."""
import numpy as np
import pytest

from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def permuted_problem(D=4, K=6, N=60, S=3, Ns=128, seed=5):
    cfg = dict(D=D, N=N, K=K, S=S, Ns=Ns, target="rosenbrock", noisy=False)
    w = workloads.build(cfg, orc.gplite_post, seeds=(seed, seed + 1, seed + 2, seed + 3))
    vp, theta, eps = w["vp"], w["theta"], w["epsilon"]
    perm = np.random.Generator(np.random.Philox(seed)).permutation(K)
    vp2 = dict(vp)
    vp2["mu"] = np.asarray(vp["mu"]).reshape(D, K)[:, perm].copy()
    for f in ("sigma", "w", "eta"):
        vp2[f] = np.ravel(vp[f])[perm].copy()
    vp2.pop("bounds", None)
    theta2 = workloads.theta_of(vp2)
    eps2 = eps[perm].copy()
    # index map of theta = [mu(:) (d fastest within a component); log sigma (K); log lambda (D); eta (K)]
    idx = np.concatenate([(perm[:, None] * D + np.arange(D)[None, :]).ravel(), D * K + perm, D * K + K + np.arange(D), D * K + K + D + perm])
    return w, vp2, theta2, eps2, idx


def check_permutation(negelcbo, vpbounds, tol=1e-12):
    w, vp2, theta2, eps2, idx = permuted_problem()
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    assert np.array_equal(theta[idx], theta2)
    _, tb = vpbounds(vp, gp, workloads.VP_OPTIONS)
    _, tb2 = vpbounds(vp2, gp, workloads.VP_OPTIONS)
    a = negelcbo(theta, 0.0, vp, gp, 128, 1, 0, 0, tb, 0, epsilon=eps, nargout=4)
    b = negelcbo(theta2, 0.0, vp2, gp, 128, 1, 0, 0, tb2, 0, epsilon=eps2, nargout=4)
    assert rel(b[0], a[0]) < tol and rel(b[2], a[2]) < tol and rel(b[3], a[3]) < tol
    assert rel(b[1], np.asarray(a[1])[idx]) < 1e-10


def check_antithetic(negelcbo, vpbounds, tol=1e-13):
    w, *_ = permuted_problem(seed=9)
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    _, tb = vpbounds(vp, gp, workloads.VP_OPTIONS)
    a = negelcbo(theta, 0.0, vp, gp, 128, 1, 0, 0, tb, 0, epsilon=eps, nargout=4)
    b = negelcbo(theta, 0.0, vp, gp, 128, 1, 0, 0, tb, 0, epsilon=-eps, nargout=4)
    assert rel(b[3], a[3]) < tol and rel(b[1], a[1]) < 1e-10 and rel(b[0], a[0]) < tol


def test_oracle_component_relabelling():
    check_permutation(orc.negelcbo_vbmc, orc.vpbounds)


def test_oracle_antithetic_draws():
    check_antithetic(orc.negelcbo_vbmc, orc.vpbounds)
