"""fminadam (utils/fminadam.m:1-102): the oracle restatement on closed-form objectives (CPU), and the device-resident
loop vbmc_b200_fminadam against the oracle loop driving the oracle's negelcbo_vbmc (GPU, through the C ABI)."""
import math

import numpy as np
import pytest

from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


# ---------------------------------------------------------------------------------------------- oracle (CPU)
def test_oracle_fminadam_update_matches_adamstate():
    """The loop's update (:51-60) is the AdamState update used by the step harness."""
    rng = np.random.default_rng(0)
    A = rng.standard_normal((6, 6))
    A = A @ A.T + np.eye(6)
    fun = lambda x: (0.5 * x @ A @ x, A @ x)
    x0 = rng.standard_normal(6)
    x, f, xtab, ftab, it = orc.fminadam(fun, x0, MaxIter=30)
    st = orc.AdamState(6)
    xs = x0.copy()
    for i in range(30):
        xs = st.update(xs, A @ xs)
        assert np.allclose(xtab[:, i], xs, rtol=0, atol=1e-15)
    assert it == 30 and xtab.shape == (6, 30) and ftab.shape == (30,)
    assert np.allclose(x, xtab[:, 10:30].mean(axis=1)) and math.isclose(f, ftab[10:30].mean())


def test_oracle_fminadam_terminates_on_flat_objective():
    """Constant objective, zero gradient: slope = 0 and dx = 0 at the first test (iter 40) -> break (:80)."""
    fun = lambda x: (3.0, np.zeros_like(x))
    x, f, xtab, ftab, it = orc.fminadam(fun, np.ones(4), MaxIter=500)
    assert it == 40 and f == 3.0 and np.array_equal(x, np.ones(4))


def test_oracle_fminadam_bounds_and_stepsize():
    fun = lambda x: (float(np.sum(x)), np.ones_like(x))   # pushes x down forever
    x, f, xtab, ftab, it = orc.fminadam(fun, np.zeros(3), LB=-0.25 * np.ones(3), UB=np.ones(3), MaxIter=100,
                                         master_stepsize={"max": 0.05, "min": 0.01, "decay": 50})
    assert np.all(xtab >= -0.25) and np.allclose(xtab[:, -1], -0.25)
    # first update: mhat = g, vhat = g^2 -> x1 = -stepsize(1) * 1/(1+sqrt(eps))
    s1 = 0.01 + 0.04 * math.exp(-1 / 50)
    assert math.isclose(xtab[0, 0], -s1 / (1 + math.sqrt(np.finfo(float).eps)), rel_tol=1e-14)


def test_oracle_fminadam_slope_statistics_match_polyfit():
    """The normal-equation form used on the device equals polyfit's QR form (:69-73)."""
    rng = np.random.default_rng(3)
    y = 2.0 - 0.03 * np.arange(20) + 0.1 * rng.standard_normal(20)
    xx = np.linspace(-9.5, 9.5, 20)
    p, cov = np.polyfit(xx, y, 1, cov="unscaled")
    r = y - np.polyval(p, xx)
    A11 = cov[0, 0] * (r @ r) / 18
    slope = (xx @ y) / (xx @ xx)
    A11_dev = (np.sum((y - (slope * xx + y.mean())) ** 2) / 18) / (xx @ xx)
    assert math.isclose(slope, p[0], rel_tol=1e-12) and math.isclose(A11, A11_dev, rel_tol=1e-10)
    assert math.isclose(xx @ xx, 665.0)


# ---------------------------------------------------------------------------------------------- device loop (GPU)
def _mk(D, N, K, S, Ns, seed=0, target="rosenbrock"):
    cfg = dict(D=D, N=N, K=K, S=S, Ns=Ns, target=target, noisy=False)
    return workloads.build(cfg, orc.gplite_post, seeds=(seed + 1, seed + 2, seed + 3, seed + 4))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,maxiter", [(dict(D=2, N=50, K=2, S=4, Ns=100), 60), (dict(D=4, N=60, K=6, S=3, Ns=64), 45),
                                           (dict(D=6, N=80, K=9, S=2, Ns=128), 100)],
                         ids=["c1", "D4K6", "D6K9"])
def test_device_fminadam_matches_oracle_loop(gpu_ctx, shape, maxiter):
    import vbmc_b200
    w = _mk(**shape)
    vp, gp, theta0, eps, Ns = w["vp"], w["gp"], w["theta"], w["epsilon"], shape["Ns"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    ms = {"max": 0.02, "min": 0.001, "decay": 200}
    fun = lambda t: orc.negelcbo_vbmc(t, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=2)[:2]
    xo, fo, xtabo, ftabo, ito = orc.fminadam(fun, theta0, None, None, 0.001, maxiter, ms)
    x, f, xtab, ftab, it = vbmc_b200.fminadam_negelcbo(theta0, 0.0, vp, gp, Ns, 0, tb, None, None, 0.001, maxiter, ms, epsilon=eps)
    assert it == ito and xtab.shape == (it, theta0.size) and ftab.shape == (it,)
    # the iterates feed back into themselves: round-off grows with the iteration count; 1e-8 after <=100 iterations
    assert rel(ftab, ftabo) < 1e-8 and rel(xtab, xtabo.T) < 1e-8
    assert rel(x, xo) < 1e-8 and rel(f, fo) < 1e-8
    # the first evaluation is at the 1e-10 parity bar of the single step; later ones inherit Adam's own sensitivity:
    # mhat/(sqrt(vhat)+sqrt(eps)) amplifies the absolute round-off of a near-zero gradient component by 1/sqrt(eps)
    assert rel(ftab[0], ftabo[0]) < 1e-10 and rel(xtab[:3], xtabo.T[:3]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("beta", [1.5, -0.7])
def test_device_fminadam_with_variance_penalty(gpu_ctx, beta):
    """beta ~= 0 (ELCBOWeight, negelcbo_vbmc.m:119-130): F + beta*sqrt(varF) with the diagonal variance and its gradient, every
    iteration; the loop is host-driven then (one synchronisation per iteration), all N-long work stays on the device."""
    import vbmc_b200
    shape, maxiter = dict(D=4, N=60, K=6, S=3, Ns=64), 45
    w = _mk(**shape)
    vp, gp, theta0, eps, Ns = w["vp"], w["gp"], w["theta"], w["epsilon"], shape["Ns"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    ms = {"max": 0.02, "min": 0.001, "decay": 200}
    fun = lambda t: orc.negelcbo_vbmc(t, beta, vp, gp, Ns, 1, 2, 0, tb, 0, epsilon=eps, nargout=2)[:2]
    xo, fo, xtabo, ftabo, ito = orc.fminadam(fun, theta0, None, None, 0.001, maxiter, ms)
    x, f, xtab, ftab, it = vbmc_b200.fminadam_negelcbo(theta0, beta, vp, gp, Ns, 2, tb, None, None, 0.001, maxiter, ms, epsilon=eps)
    assert it == ito
    assert rel(ftab[0], ftabo[0]) < 1e-10 and rel(xtab[:3], xtabo.T[:3]) < 1e-9
    assert rel(ftab, ftabo) < 1e-8 and rel(xtab, xtabo.T) < 1e-8 and rel(x, xo) < 1e-8 and rel(f, fo) < 1e-8
    # and it is not the unpenalised path
    f0 = orc.negelcbo_vbmc(theta0, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=2)[0]
    assert abs(ftab[0] - f0) > 1e-6 * abs(f0)


@pytest.mark.gpu
def test_device_fminadam_terminates_like_reference(gpu_ctx):
    """Large TolFun: the slope test passes at iteration 40 in both implementations (fminadam.m:65-83)."""
    import vbmc_b200
    w = _mk(D=3, N=40, K=4, S=2, Ns=64)
    vp, gp, theta0, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    ms = {"max": 0.002, "min": 0.001, "decay": 200}
    fun = lambda t: orc.negelcbo_vbmc(t, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, epsilon=eps, nargout=2)[:2]
    xo, fo, xtabo, ftabo, ito = orc.fminadam(fun, theta0, None, None, 10.0, 200, ms)
    x, f, xtab, ftab, it = vbmc_b200.fminadam_negelcbo(theta0, 0.0, vp, gp, 64, 0, tb, None, None, 10.0, 200, ms, epsilon=eps)
    assert ito == 40 and it == 40
    assert rel(x, xo) < 1e-9 and rel(f, fo) < 1e-9
    st = vbmc_b200.fminadam_negelcbo.last_stats
    assert st["stop"] == 1.0 and st["dx"] < 0.1 and abs(st["slope"]) < st["slope_err"]


@pytest.mark.gpu
def test_device_fminadam_bounds_and_philox_streams(gpu_ctx):
    """LB/UB clamp (:60) and fresh draws per iteration: iteration i uses Philox stream `stream + i`."""
    import vbmc_b200
    w = _mk(D=3, N=40, K=4, S=2, Ns=64)
    vp, gp, theta0 = w["vp"], w["gp"], w["theta"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    lb, ub = theta0 - 0.01, theta0 + 0.01
    seed, stream, maxiter = 99, 1000, 25
    draws = [gpu_ctx.eps_philox(3, 4, 64, seed, stream + i, readback=True) for i in range(maxiter)]
    calls = {"i": 0}

    def fun(t):
        e = draws[calls["i"]]
        calls["i"] += 1
        return orc.negelcbo_vbmc(t, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, epsilon=e, nargout=2)[:2]

    xo, fo, xtabo, ftabo, ito = orc.fminadam(fun, theta0, lb, ub, None, maxiter, None)
    x, f, xtab, ftab, it = vbmc_b200.fminadam_negelcbo(theta0, 0.0, vp, gp, 64, 0, tb, lb, ub, None, maxiter, None, rng=(seed, stream))
    assert it == ito == maxiter
    assert np.all(xtab >= lb - 1e-15) and np.all(xtab <= ub + 1e-15)
    assert rel(ftab, ftabo) < 1e-9 and rel(xtab, xtabo.T) < 1e-9 and rel(x, xo) < 1e-9


@pytest.mark.gpu
def test_device_fminadam_errors(gpu_ctx):
    import vbmc_b200
    w = _mk(D=2, N=30, K=2, S=2, Ns=32)
    vp, gp, theta0, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    with pytest.raises(vbmc_b200.VbmcB200Error) as e:   # negelcbo_vbmc.m:19-20: the gradient of the full variance does not exist
        vbmc_b200.fminadam_negelcbo(theta0, 1.5, vp, gp, 32, 1, None, epsilon=eps)
    assert e.value.identifier == "negelcbo_vbmc:vargrad"
    with pytest.raises(vbmc_b200.VbmcB200Error):
        vbmc_b200.fminadam_negelcbo(theta0, 0.0, vp, gp, 32, 0, None, MaxIter=10, epsilon=eps)
    with pytest.raises(vbmc_b200.VbmcB200Error):
        vbmc_b200.fminadam_negelcbo(theta0[:-1], 0.0, vp, gp, 32, 0, None, epsilon=eps)
