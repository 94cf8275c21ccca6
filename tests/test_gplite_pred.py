"""gplite_pred (gplite/gplite_pred.m:1-163): oracle identities (CPU) and CUDA-vs-oracle parity through the C ABI (GPU)."""
import math

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads

TOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def _mk(D, N, S, seed=0, target="rosenbrock", noisy=False, **kw):
    cfg = dict(D=D, N=N, K=2, S=S, Ns=4, target=target, noisy=noisy, **kw)
    return workloads.build(cfg, orc.gplite_post, seeds=(seed + 1, seed + 2, seed + 3, seed + 4), with_eps=False)


# ---------------------------------------------------------------------------------------------- oracle (CPU)
def test_oracle_pred_matches_dense_gaussian_conditioning():
    """fmu, fs2 against the textbook formulas with an explicit (K + Sigma)^-1 (independent of the factor algebra)."""
    w = _mk(3, 35, 2)
    gp = w["gp"]
    X, y = gp["X"], gp["y"]
    Xs = np.random.default_rng(5).standard_normal((6, 3))
    ymu, ys2, fmu, fs2 = orc.gplite_pred(gp, Xs, ssflag=True, nargout=4)
    for s, post in enumerate(gp["post"]):
        h = post["hyp"]
        ell, sf2, sn2 = np.exp(h[:3]), math.exp(2 * h[3]), math.exp(2 * h[4])
        k = lambda A, B: sf2 * np.exp(-0.5 * (((A[:, None, :] - B[None, :, :]) / ell) ** 2).sum(-1))
        m = lambda A: h[5] - 0.5 * (((A - h[6:9]) / np.exp(h[9:12])) ** 2).sum(-1)
        Kn = k(X, X) + sn2 * post["sn2_mult"] * np.eye(len(y))
        Ks = k(X, Xs)
        assert rel(fmu[:, s], m(Xs) + Ks.T @ np.linalg.solve(Kn, y - m(X))) < 1e-8
        assert rel(fs2[:, s], sf2 - np.sum(Ks * np.linalg.solve(Kn, Ks), axis=0)) < 1e-6
        assert rel(ys2[:, s], fs2[:, s] + sn2 * post["sn2_mult"]) < 1e-12


def test_oracle_pred_sample_average_and_lp():
    w = _mk(2, 30, 5)
    gp = w["gp"]
    Xs = np.random.default_rng(6).standard_normal((4, 2))
    ys = np.random.default_rng(7).standard_normal(4)
    ymu_s, ys2_s, fmu_s, fs2_s, lp = orc.gplite_pred(gp, Xs, ys, None, True, nargout=5)
    ymu, ys2, fmu, fs2 = orc.gplite_pred(gp, Xs, nargout=4)
    assert rel(fmu, fmu_s.mean(1)) < 1e-14
    assert rel(fs2, fs2_s.mean(1) + fmu_s.var(1, ddof=1)) < 1e-13           # :157-158
    assert rel(ys2, ys2_s.mean(1) + ymu_s.var(1, ddof=1)) < 1e-13
    assert rel(lp, -0.5 * (ys[:, None] - ymu_s) ** 2 / ys2_s - 0.5 * np.log(2 * math.pi * ys2_s)) < 1e-14
    with pytest.raises(orc.OracleError):
        orc.gplite_pred(gp, Xs, ystar=np.zeros(3))


# ---------------------------------------------------------------------------------------------- CUDA (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [dict(D=2, N=50, S=8), dict(D=1, N=20, S=1), dict(D=5, N=130, S=3), dict(D=10, N=300, S=4, target="lumpy")],
                         ids=lambda s: "D{D}N{N}S{S}".format(**s))
@pytest.mark.parametrize("Nstar", [1, 37])
def test_pred_matches_oracle(gpu_ctx, shape, Nstar):
    import vbmc_b200
    w = _mk(**shape)
    gp = w["gp"]
    D = shape["D"]
    r = np.random.default_rng(11)
    Xs = np.concatenate([gp["X"][:Nstar // 2] + 0.05 * r.standard_normal((Nstar // 2, D)), 1.5 * r.standard_normal((Nstar - Nstar // 2, D))])
    ys = r.standard_normal(Nstar)
    for ssflag in (False, True):
        got = vbmc_b200.gplite_pred(gp, Xs, ys, None, ssflag, nargout=5)
        ref = orc.gplite_pred(gp, Xs, ys, None, ssflag, nargout=5)
        for g, o, name in zip(got, ref, ("ymu", "ys2", "fmu", "fs2", "lp")):
            assert g.shape == o.shape, name
            # fs2 = kss - sum(V.^2) cancels: compare on the scale of kss (the reference has the same cancellation)
            scale = np.max(np.abs(o)) if name not in ("fs2", "ys2") else max(np.max(np.abs(o)), math.exp(2 * gp["post"][0]["hyp"][D]))
            # lp divides by ys2 = fs2 + sn2 with sn2 ~ 1e-5: the round-off of the cancellation in fs2 (~1e-16 kss, in the
            # reference as well) is amplified by kss/ys2 ~ 1e5..1e7
            assert np.max(np.abs(g - o)) / scale < (1e-6 if name == "lp" else TOL), name
    (ymu,) = vbmc_b200.gplite_pred(gp, Xs, nargout=1)
    assert rel(ymu, orc.gplite_pred(gp, Xs, nargout=1)[0]) < TOL


@pytest.mark.gpu
def test_pred_after_device_refit_and_noisy_gp(gpu_ctx):
    """Posterior computed by vbmc_b200.gplite_post (factors stay on the device, leading dimension Np) + user noise s2star."""
    import vbmc_b200
    w = _mk(4, 90, 3, noisy=True)
    X, y, s2, hyp = w["X"], w["y"], w["s2"], w["hyp"]
    gp_dev = vbmc_b200.gplite_post(hyp, X, y, 1, 4, [1, 1, 0], s2, want_L=False)
    gp_ref = orc.gplite_post(hyp, X, y, 1, 4, [1, 1, 0], s2)
    r = np.random.default_rng(3)
    Xs, s2s = r.standard_normal((25, 4)), 0.5 + r.random(25)
    got = vbmc_b200.gplite_pred(gp_dev, Xs, None, s2s, True, nargout=4)
    ref = orc.gplite_pred(gp_ref, Xs, None, s2s, True, nargout=4)
    for g, o in zip(got, ref):
        assert rel(g, o) < 1e-9


@pytest.mark.gpu
def test_pred_low_noise_inverse_form_posterior(gpu_ctx):
    """Lchol == 0: post.L = -inv(K+Sigma) handed over by the host (gplite_pred.m:100-102)."""
    import vbmc_b200
    w = _mk(2, 40, 2, log_sn=0.5 * math.log(1e-7))
    gp = w["gp"]
    assert not any(p["Lchol"] for p in gp["post"])
    Xs = np.random.default_rng(4).standard_normal((9, 2))
    got = vbmc_b200.gplite_pred(gp, Xs, None, None, True, nargout=4)
    ref = orc.gplite_pred(gp, Xs, None, None, True, nargout=4)
    assert rel(got[2], ref[2]) < 1e-8
    sf2 = math.exp(2 * gp["post"][0]["hyp"][2])
    # sn2 = 1e-7 makes K + Sigma numerically singular (cond ~ 1e10): kss + sum(Ks.*(L*Ks)) with an explicit inverse carries
    # round-off of order cond * eps * kss in the reference too (half of its values clamp to 0)
    assert np.max(np.abs(got[3] - ref[3])) / sf2 < 1e-5


@pytest.mark.gpu
def test_pred_errors(gpu_ctx):
    import vbmc_b200
    w = _mk(2, 30, 2)
    with pytest.raises(vbmc_b200.VbmcB200Error) as e:
        vbmc_b200.gplite_pred(w["gp"], np.zeros((3, 2)), ystar=np.zeros(4))
    assert e.value.identifier == "gplite_pred:ydimmismatch"
    with pytest.raises(vbmc_b200.VbmcB200Error) as e:
        vbmc_b200.gplite_pred(w["gp"], np.zeros((3, 2)), s2star=np.zeros(2))
    assert e.value.identifier == "gplite_pred:s2dimmismatch"


# ---------------------------------------------------------------------------------------------- rank-one update
def test_oracle_rank1_equals_full_refit():
    """gplite_test.m:87-105: a sequence of rank-one updates reproduces the posterior of the full refit."""
    w = _mk(3, 40, 2, log_sn=math.log(0.05))
    X, y, hyp = w["X"], w["y"], w["hyp"]
    full = orc.gplite_post(hyp, X, y, 1, 4, [1, 0, 0], None)
    gp1 = orc.gplite_post(hyp, X[:32], y[:32], 1, 4, [1, 0, 0], None)
    for i in range(32, 40):
        gp1 = orc.gplite_post_update1(gp1, X[i], y[i])
    for s in range(2):
        assert rel(gp1["post"][s]["alpha"], full["post"][s]["alpha"]) < 1e-7
        assert rel(gp1["post"][s]["L"], full["post"][s]["L"]) < 1e-9
        assert rel(gp1["post"][s]["sW"], full["post"][s]["sW"]) < 1e-13
    with pytest.raises(orc.OracleError):
        orc.gplite_post_update1(gp1, X[:2], y[0])


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [True, False], ids=["after_gp_post", "attached_from_host"])
def test_rank1_update_matches_oracle(gpu_ctx, resident):
    """Eight successive updates on the device (factor buffer grows past its leading dimension when attached from the host)."""
    import vbmc_b200
    w = _mk(4, 70, 3, log_sn=math.log(0.05))
    X, y, hyp = w["X"], w["y"], w["hyp"]
    n0 = 62   # the factor of a 62-point GP is padded to 64 columns: updates 63, 64 fit, 65 forces the buffer to grow
    ref = orc.gplite_post(hyp, X[:n0], y[:n0], 1, 4, [1, 0, 0], None)
    gp = vbmc_b200.gplite_post(hyp, X[:n0], y[:n0], 1, 4, [1, 0, 0], None, want_L=True) if resident else ref
    for i in range(n0, 70):
        gp = vbmc_b200.gplite_post_update1(gp, X[i], y[i])
        ref = orc.gplite_post_update1(ref, X[i], y[i])
        for s in range(3):
            assert rel(gp["post"][s]["alpha"], ref["post"][s]["alpha"]) < 1e-8
            assert rel(gp["post"][s]["L"], ref["post"][s]["L"]) < 1e-10
            assert rel(gp["post"][s]["sW"], ref["post"][s]["sW"]) < 1e-13
    assert gp["X"].shape == (70, 4) and gp["y"].shape == (70,)
    # the updated posterior is the one resident on the device: predictions and a negelcbo step use it
    Xs = np.random.default_rng(2).standard_normal((5, 4))
    got = vbmc_b200.gplite_pred(gp, Xs, None, None, True, nargout=4)
    exp = orc.gplite_pred(ref, Xs, None, None, True, nargout=4)
    sf2 = math.exp(2 * hyp[4, 0])
    assert rel(got[2], exp[2]) < 1e-8 and np.max(np.abs(got[3] - exp[3])) / sf2 < 1e-9
    full = orc.gplite_post(hyp, X, y, 1, 4, [1, 0, 0], None)
    assert rel(gp["post"][0]["alpha"], full["post"][0]["alpha"]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [True, False], ids=["device_factor", "inverse_from_host"])
def test_rank1_update_of_a_low_noise_posterior(gpu_ctx, resident):
    """gplite_post.m:234-238: sample 0 has sn2 < 1e-6, so post.L = -inv(K + sn2*I) and the update touches every entry of it; the
    other samples take the Cholesky branch in the same call.  After a device refit the low-noise sample lives as an unscaled
    factor (bordered-factor update, inverse rebuilt on request); attached from the host it lives as the inverse itself."""
    import vbmc_b200
    w = _mk(4, 70, 3, log_sn=math.log(0.05))
    X, y, hyp = w["X"], w["y"], w["hyp"].copy()
    hyp[5, 0] = math.log(3e-4)
    n0 = 62
    ref = orc.gplite_post(hyp, X[:n0], y[:n0], 1, 4, [1, 0, 0], None)
    assert [p["Lchol"] for p in ref["post"]] == [False, True, True]
    gp = vbmc_b200.gplite_post(hyp, X[:n0], y[:n0], 1, 4, [1, 0, 0], None, want_L=True) if resident else ref
    assert [p["Lchol"] for p in gp["post"]] == [False, True, True] and gp["post"][0]["sn2_mult"] == ref["post"][0]["sn2_mult"]
    ell, sf2, sn2 = np.exp(hyp[:4, 0]), math.exp(2 * hyp[4, 0]), math.exp(2 * hyp[5, 0])
    for i in range(n0, 70):
        gp = vbmc_b200.gplite_post_update1(gp, X[i], y[i])
        ref = orc.gplite_post_update1(ref, X[i], y[i])
        n = i + 1
        A = sf2 * np.exp(-0.5 * orc.sq_dist((X[:n] / ell).T)) + sn2 * np.eye(n)
        bound = 50 * np.linalg.cond(A) * np.finfo(float).eps
        a, b = gp["post"][0], ref["post"][0]
        assert a["L"].shape == (n, n) and not a["Lchol"]
        # it IS the negated inverse of the bordered matrix (the inverse-form recursion accumulates round-off in the oracle too)
        assert np.max(np.abs(a["L"] @ A + np.eye(n))) < max(bound, 2 * np.max(np.abs(b["L"] @ A + np.eye(n))))
        assert rel(a["L"], b["L"]) < bound and rel(a["alpha"], b["alpha"]) < bound
        assert rel(a["sW"], b["sW"]) < 1e-13
        for s in (1, 2):
            assert rel(gp["post"][s]["alpha"], ref["post"][s]["alpha"]) < 1e-8
            assert rel(gp["post"][s]["L"], ref["post"][s]["L"]) < 1e-10
    Xs = np.random.default_rng(2).standard_normal((5, 4))
    got = vbmc_b200.gplite_pred(gp, Xs, None, None, True, nargout=4)
    exp = orc.gplite_pred(ref, Xs, None, None, True, nargout=4)
    assert rel(got[2], exp[2]) < 1e-6 and np.max(np.abs(got[3] - exp[3])) / sf2 < 1e-6


@pytest.mark.gpu
def test_rank1_update_with_s2_falls_back_to_refit(gpu_ctx):
    import vbmc_b200
    w = _mk(3, 40, 2, noisy=True)
    X, y, s2, hyp = w["X"], w["y"], w["s2"], w["hyp"]
    gp = vbmc_b200.gplite_post(hyp, X[:39], y[:39], 1, 4, [1, 1, 0], s2[:39])
    gp = vbmc_b200.gplite_post_update1(gp, X[39], y[39], s2[39])
    ref = orc.gplite_post(hyp, X, y, 1, 4, [1, 1, 0], s2)
    assert gp["X"].shape == (40, 3) and rel(gp["post"][1]["alpha"], ref["post"][1]["alpha"]) < 1e-8
    with pytest.raises(vbmc_b200.VbmcB200Error) as e:
        vbmc_b200.gplite_post_update1(gp, X[:2], y[0])
    assert e.value.identifier == "gplite_post:NotRankOne"
