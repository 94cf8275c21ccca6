"""Device entlb_vbmc and negelcbo_vbmc with Ns == 0 (ent/entlb_vbmc.m; negelcbo_vbmc.m:102-109; the call of
misc/vpsieve_vbmc.m:76) against the oracle restatement.  (File name: runs last under `pytest -x`.)"""
import numpy as np
import pytest

from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def mk(D, N, K, S, seed=0, **kw):
    cfg = dict(D=D, N=N, K=K, S=S, Ns=2, target="rosenbrock", noisy=False, **kw)
    return workloads.build(cfg, orc.gplite_post, seeds=(seed + 1, seed + 2, seed + 3, seed + 4))


@pytest.mark.parametrize("D,K", [(1, 1), (4, 1), (2, 2), (3, 5), (6, 20), (10, 50), (7, 128)])
@pytest.mark.parametrize("jac", [True, False])
def test_entlb_matches_oracle(gpu_ctx, D, K, jac):
    import vbmc_b200
    vp = mk(D, max(30, K + 5), K, 1)["vp"]
    H, dH = vbmc_b200.entlb_vbmc(vp, [1, 1, 1, 1], jac)
    Ho, dHo = orc.entlb_vbmc(vp, [1, 1, 1, 1], jac)
    assert rel(H, Ho) < 1e-11 and dH.shape == dHo.shape and rel(dH, dHo) < TOL
    (H1,) = vbmc_b200.entlb_vbmc(vp, nargout=1)
    assert H1 == H


def test_entlb_grad_flag_subsets(gpu_ctx):
    import vbmc_b200
    vp = mk(3, 30, 4, 1)["vp"]
    for gf in [(1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1), (1, 0, 1, 1), (0, 0, 0, 0)]:
        H, dH = vbmc_b200.entlb_vbmc(vp, list(gf), True)
        Ho, dHo = orc.entlb_vbmc(vp, list(gf), True)
        assert rel(H, Ho) < 1e-11 and dH.shape == dHo.shape
        assert dH.size == 0 or rel(dH, dHo) < TOL


@pytest.mark.parametrize("shape", [dict(D=2, N=50, K=2, S=8), dict(D=3, N=33, K=5, S=2), dict(D=6, N=120, K=20, S=4),
                                   dict(D=4, N=40, K=1, S=2)])
def test_negelcbo_with_Ns_zero_matches_oracle(gpu_ctx, shape):
    """The sieve's call pattern (no gradient) and the deterministic-entropy optimisation's (with gradient)."""
    import vbmc_b200
    w = mk(**shape)
    vp, gp, theta = w["vp"], w["gp"], w["theta"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    th = theta.copy()
    th[0] = tb["ub"][0] + 0.1
    F, dF, G, H, varF, dH = vbmc_b200.negelcbo_vbmc(th, 0.0, vp, gp, 0, 1, 0, 0, tb, 0, nargout=6)
    Fo, dFo, Go, Ho, _, dHo = orc.negelcbo_vbmc(th, 0.0, vp, gp, 0, 1, 0, 0, tb, 0, nargout=6)[:6]
    assert rel(G, Go) < TOL and rel(H, Ho) < TOL and rel(dH, dHo) < TOL
    assert rel(F, Fo) < TOL and rel(dF, dFo) < TOL
    F1, _, _, _, varF1 = vbmc_b200.negelcbo_vbmc(th, 0.0, vp, gp, 0, 0, 0, 0, tb, 0, nargout=5)   # vpsieve_vbmc.m:76
    assert rel(F1, Fo) < TOL and varF1 == 0.0
    # an ordinary Monte-Carlo step afterwards is unaffected by the token draws of the Ns == 0 call
    eps = workloads.make_epsilon(dict(D=shape["D"], K=shape["K"], Ns=64))
    a = vbmc_b200.negelcbo_vbmc(th, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, epsilon=eps, nargout=2)
    b = orc.negelcbo_vbmc(th, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, epsilon=eps, nargout=2)
    assert rel(a[0], b[0]) < TOL and rel(a[1], b[1]) < TOL


def test_Ns_zero_call_leaves_the_resident_draws_alone(gpu_ctx):
    """a vpsieve-style Ns == 0 evaluation between two steps on RESIDENT draws (uploaded once) or between two steps of a
    streaming generator-mode caller must not disturb either: it makes no draws at all."""
    import vbmc_b200
    w = mk(D=3, N=40, K=4, S=2)
    vp, gp, theta = w["vp"], w["gp"], w["theta"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    eps = workloads.make_epsilon(dict(D=3, K=4, Ns=64))
    ref = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, epsilon=eps, nargout=2)
    gpu_ctx.eps_upload(eps)
    a = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, epsilon="resident", nargout=2)
    vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 0, 0, 0, 0, tb, 0, nargout=1)
    b = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, epsilon="resident", nargout=2)   # ESTATE in round 1
    assert a[0] == ref[0] == b[0] and np.array_equal(a[1], ref[1]) and np.array_equal(b[1], ref[1])
    # streaming caller (stream advancing by one: ahead-of-time draws) with sieve calls in between
    plain = [vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, rng=(31, 200 + i), nargout=2) for i in range(5)]
    mixed = []
    for i in range(5):
        mixed.append(vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, rng=(31, 200 + i), nargout=2))
        vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 0, 0, 0, 0, tb, 0, nargout=1)
    for p, m in zip(plain, mixed):
        assert p[0] == m[0] and np.array_equal(p[1], m[1])
