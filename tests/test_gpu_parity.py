"""CUDA path vs the CPU oracle through the C ABI (host NumPy buffers in, host buffers out).
FP64 tolerance (BASELINE.json north_star): 1e-10 relative on F, dF, G, H, dH (max-norm relative)."""
import math

import numpy as np
import pytest

from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads

from _truth import errs_vs_truth, truth_negelcbo

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def mk(D, N, K, S, Ns, seed=0, target="rosenbrock", **kw):
    cfg = dict(D=D, N=N, K=K, S=S, Ns=Ns, target=target, noisy=False, **kw)
    return workloads.build(cfg, orc.gplite_post, seeds=(seed + 1, seed + 2, seed + 3, seed + 4))


SHAPES = [
    dict(D=2, N=50, K=2, S=8, Ns=100),          # c1 (rosenbrock_test plumbing case)
    dict(D=1, N=20, K=1, S=1, Ns=2),            # smallest legal problem
    dict(D=3, N=33, K=5, S=2, Ns=37),           # odd D, odd Ns (made even), ragged tiles
    dict(D=5, N=100, K=7, S=3, Ns=1000),
    dict(D=6, N=400, K=20, S=8, Ns=4096),       # c2
    dict(D=9, N=64, K=33, S=2, Ns=130),         # K > 32: second round of the column sums
    dict(D=10, N=300, K=50, S=4, Ns=512, target="lumpy"),   # c3 shape, reduced N/Ns/S
    dict(D=20, N=128, K=12, S=2, Ns=96, target="lumpy"),    # c5 dimension
    dict(D=20, N=160, K=100, S=2, Ns=128, target="lumpy"),  # c5 dimension and component count (K=100)
    dict(D=7, N=140, K=128, S=1, Ns=66),
    dict(D=4, N=200, K=160, S=2, Ns=64),                    # K = N^(2/3) at N = 2000 (vbmc.m:247): five list rounds of 32
    dict(D=3, N=260, K=252, S=1, Ns=34),                    # K = N^(2/3) at N = 4000; one warp per CTA (the stage takes 126 KB)
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "D{D}N{N}K{K}S{S}Ns{Ns}".format(**s))
def test_negelcbo_matches_oracle(gpu_ctx, shape):
    import vbmc_b200
    w = mk(**shape)
    vp, gp, theta, eps, Ns = w["vp"], w["gp"], w["theta"], w["epsilon"], shape["Ns"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    got = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
    ref = orc.negelcbo_vbmc(theta, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
    F, dF, G, H, varF, dH = got
    Fo, dFo, Go, Ho, _, dHo = ref[:6]
    assert rel(H, Ho) < TOL and rel(dH, dHo) < TOL
    # log-joint side: the gate is against the binary128 evaluation (tests/test_truth128.py explains why); the FP64 NumPy oracle
    # must agree with the CUDA path to within ITS OWN distance from that truth
    t = truth_negelcbo(vp, gp, theta, Ns, eps, tb)
    e = errs_vs_truth(dict(F=F, dF=dF, G=G, H=H, dH=dH), t)
    eo = errs_vs_truth(dict(F=Fo, dF=dFo, G=Go, H=Ho, dH=dHo), t)
    assert max(e.values()) < TOL, e
    assert rel(G, Go) < TOL + 2 * eo["G"] and rel(F, Fo) < TOL + 2 * eo["F"] and rel(dF, dFo) < TOL + 2 * eo["dF"], (e, eo)
    assert varF == 0.0


def test_negelcbo_no_grad_and_no_bounds(gpu_ctx):
    import vbmc_b200
    w = mk(D=4, N=60, K=6, S=3, Ns=200)
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    (F,) = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 200, epsilon=eps, nargout=1)
    Fo = orc.negelcbo_vbmc(theta, 0.0, vp, gp, 200, epsilon=eps, nargout=1)[0]
    assert rel(F, Fo) < TOL
    F4 = vbmc_b200.negelcbo_vbmc(theta, float("nan"), vp, gp, 200, 0, 0, epsilon=eps, nargout=4)
    assert F4[1] is None and rel(F4[0], Fo) < TOL  # beta NaN -> 0 (negelcbo_vbmc.m:15)


def test_soft_bound_and_weight_penalties(gpu_ctx):
    """theta outside the soft bounds + small weights below WeightThreshold (negelcbo_vbmc.m:136-164)."""
    import vbmc_b200
    w = mk(D=3, N=40, K=6, S=2, Ns=64)
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"].copy(), w["epsilon"]
    D, K = 3, 6
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    theta[0] = tb["ub"][0] + 0.4
    theta[5] = tb["lb"][5] - 0.2
    theta[D * K + 1] = -20.0            # log sigma far below the lnscale lower bounds
    theta[D * K + K] = 3.0              # log lambda large: above lnscale upper bound for some k
    theta[-1] = 0.7                     # eta above 0
    theta[-2] = -9.0                    # tiny weight -> w < WeightThreshold
    got = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, epsilon=eps, nargout=2)
    ref = orc.negelcbo_vbmc(theta, 0.0, vp, gp, 64, 1, 0, 0, tb, 0, epsilon=eps, nargout=2)
    assert rel(got[0], ref[0]) < TOL and rel(got[1], ref[1]) < TOL
    # the penalty must be active in this case
    nob = orc.negelcbo_vbmc(theta, 0.0, vp, gp, 64, 1, 0, 0, None, 0, epsilon=eps, nargout=2)
    assert abs(ref[0] - nob[0]) > 1.0


@pytest.mark.parametrize("flags", [(1, 1, 1, 1), (1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1), (1, 1, 0, 0), (1, 0, 1, 1)])
def test_optimize_flag_subsets(gpu_ctx, flags):
    """theta layout depends on vp.optimize_* (negelcbo_vbmc.m:32-48, vpbndloss.m:9-38)."""
    import vbmc_b200
    if flags == (0, 0, 0, 1):
        pytest.skip("weights-only path = gplogjoint_weights.m, out of scope (SURVEY.md 2 #5)")
    w = mk(D=3, N=40, K=4, S=2, Ns=50)
    vp, gp, eps = dict(w["vp"]), w["gp"], w["epsilon"]
    for f, v in zip(("optimize_mu", "optimize_sigma", "optimize_lambda", "optimize_weights"), flags):
        vp[f] = bool(v)
    parts = []
    if flags[0]: parts.append(vp["mu"].T.ravel() + 0.05)
    if flags[1]: parts.append(np.log(vp["sigma"]) - 0.1)
    if flags[2]: parts.append(np.log(vp["lambda"]) + 0.02)
    if flags[3]: parts.append(np.asarray(vp["eta"]) + 0.1)
    theta = np.concatenate(parts)
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    got = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 50, 1, 0, 0, tb, 0, epsilon=eps, nargout=2)
    ref = orc.negelcbo_vbmc(theta, 0.0, vp, gp, 50, 1, 0, 0, tb, 0, epsilon=eps, nargout=2)
    assert got[1].shape == ref[1].shape
    assert rel(got[0], ref[0]) < TOL and rel(got[1], ref[1]) < TOL


@pytest.mark.parametrize("meanfun", [0, 1, 4])
def test_gplogjoint_meanfuns_and_Isk(gpu_ctx, meanfun):
    import vbmc_b200
    rng = np.random.default_rng(7)
    D, N, K, S = 3, 45, 5, 3
    X = rng.standard_normal((N, D)) * 1.5
    y = workloads.rosenbrock_logpost(X)
    Nmean = {0: 0, 1: 1, 4: 1 + 2 * D}[meanfun]
    hyp = np.zeros((D + 2 + Nmean, S))
    for s in range(S):
        hyp[:D, s] = np.log(1.0 + 0.3 * rng.random(D))
        hyp[D, s] = math.log(np.std(y))
        hyp[D + 1, s] = math.log(0.05)
        if meanfun >= 1:
            hyp[D + 2, s] = np.max(y)
        if meanfun == 4:
            hyp[D + 3 : 2 * D + 3, s] = X.mean(axis=0) + 0.1 * rng.standard_normal(D)
            hyp[2 * D + 3 :, s] = np.log(2 * X.std(axis=0))
    gp = orc.gplite_post(hyp, X, y, 1, meanfun, [1, 0, 0], None)
    vp = workloads.make_vp(dict(D=D, K=K), X, y, seed=11)
    vp["delta"] = 0.05 * np.ones(D) if meanfun == 4 else None
    got = vbmc_b200.gplogjoint(vp, gp, True, True, True, 0, nargout=6)
    ref = orc.gplogjoint(vp, gp, True, True, True, 0, nargout=6)
    assert rel(got[0], ref[0]) < TOL and rel(got[1], ref[1]) < TOL
    assert got[5].shape == (S, K) and rel(got[5], ref[5]) < TOL
    # jacobian_flag = 0
    got = vbmc_b200.gplogjoint(vp, gp, [1, 1, 1, 1], True, False, 0, nargout=2)
    ref = orc.gplogjoint(vp, gp, [1, 1, 1, 1], True, False, 0, nargout=2)
    assert rel(got[1], ref[1]) < TOL


@pytest.mark.parametrize("gf", [(1, 1, 1, 1), (1, 0, 0, 0), (0, 1, 1, 0), (0, 0, 0, 1), (0, 0, 0, 0)])
@pytest.mark.parametrize("jac", [True, False])
def test_entmc_grad_flags(gpu_ctx, gf, jac):
    import vbmc_b200
    w = mk(D=4, N=30, K=6, S=1, Ns=300)
    vp, eps = w["vp"], w["epsilon"]
    H, dH = vbmc_b200.entmc_vbmc(vp, 300, list(gf), jac, epsilon=eps)
    Ho, dHo = orc.entmc_vbmc(vp, 300, list(gf), jac, epsilon=eps)
    assert rel(H, Ho) < TOL
    assert dH.shape == dHo.shape
    if dHo.size:
        assert rel(dH, dHo) < TOL


def test_entmc_K1_known_answer(gpu_ctx):
    """K=1 closed form (entmc_vbmc.m:60-67 / entlb_vbmc.m:34) evaluated by the CUDA path itself."""
    import vbmc_b200
    rng = np.random.default_rng(1)
    D, Ns = 6, 4096
    vp = dict(D=D, K=1, mu=rng.standard_normal((D, 1)), sigma=np.array([0.3]), w=np.array([1.0]), eta=np.array([0.0]),
              optimize_mu=True, optimize_sigma=True, optimize_lambda=True, optimize_weights=True, delta=None)
    vp["lambda"] = np.exp(0.3 * rng.standard_normal(D))
    eps = rng.standard_normal((1, Ns // 2, D))
    H, dH = vbmc_b200.entmc_vbmc(vp, Ns, True, True, epsilon=eps)
    expect = 0.5 * D * math.log(2 * math.pi) + D * math.log(0.3) + np.sum(np.log(vp["lambda"])) + 0.5 * D * np.mean(eps**2)
    assert abs(H - expect) < 1e-12 * abs(expect)
    assert np.max(np.abs(dH[:D])) < 1e-12
    assert abs(dH[D] - np.mean(np.sum(eps**2, axis=2))) < 1e-11 * abs(dH[D])


def test_exp_underflow_matches_reference_semantics(gpu_ctx):
    """Far-apart tiny components: cross terms underflow to exactly 0 as in MATLAB's direct exp (entmc_vbmc.m:63)."""
    import vbmc_b200
    rng = np.random.default_rng(2)
    D, K, Ns = 3, 4, 64
    vp = dict(D=D, K=K, mu=50.0 * rng.standard_normal((D, K)), sigma=0.01 * np.ones(K), w=np.ones(K) / K,
              eta=np.log(np.ones(K) / K), optimize_mu=True, optimize_sigma=True, optimize_lambda=True,
              optimize_weights=True, delta=None)
    vp["lambda"] = np.ones(D)
    eps = rng.standard_normal((K, Ns // 2, D))
    H, dH = vbmc_b200.entmc_vbmc(vp, Ns, True, True, epsilon=eps)
    Ho, dHo = orc.entmc_vbmc(vp, Ns, True, True, epsilon=eps)
    assert np.isfinite(H) and rel(H, Ho) < TOL and rel(dH, dHo) < TOL


@pytest.mark.parametrize("switch", [("VBMC_B200_ENTMC_FORM", "direct"), ("VBMC_B200_ENTMC_TMEM", "1")], ids=["direct", "tmem_stage"])
def test_direct_formulation_forced_matches_oracle(switch):
    """VBMC_B200_ENTMC_FORM=direct runs the subtract-then-square instantiation (normally only selected by the device guard for
    huge ||u||^2) on ordinary shapes; VBMC_B200_ENTMC_TMEM=1 the opt-in variant that parks e(+-) in tensor memory.  The switches
    are read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np, vbmc_b200
from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads
rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1e-300, np.max(np.abs(b))))
for shape in (dict(D=2, N=50, K=2, S=8, Ns=100), dict(D=5, N=60, K=7, S=2, Ns=200), dict(D=10, N=80, K=50, S=2, Ns=256), dict(D=20, N=60, K=12, S=1, Ns=64),
              dict(D=3, N=200, K=140, S=1, Ns=34)):
    cfg = dict(shape, target="rosenbrock", noisy=False)
    w = workloads.build(cfg, orc.gplite_post)
    vp, eps = w["vp"], w["epsilon"]
    H, dH = vbmc_b200.entmc_vbmc(vp, shape["Ns"], True, True, epsilon=eps)
    Ho, dHo = orc.entmc_vbmc(vp, shape["Ns"], True, True, epsilon=eps)
    assert rel(H, Ho) < 1e-10 and rel(dH, dHo) < 1e-10, (shape, rel(H, Ho), rel(dH, dHo))
print("OK")
'''
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env[switch[0]] = switch[1]
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_device_rng_equals_parity_mode_on_dumped_draws(gpu_ctx):
    """Device Philox mode == parity mode fed with the draws it dumped; draws are standard normal."""
    import vbmc_b200
    w = mk(D=5, N=40, K=8, S=2, Ns=2048)
    vp, gp, theta = w["vp"], w["gp"], w["theta"]
    eps = gpu_ctx.eps_philox(5, 8, 2048, seed=1234, stream=7, readback=True)
    assert abs(eps.mean()) < 0.02 and abs(eps.std() - 1) < 0.02 and abs(np.mean(eps**3)) < 0.05
    assert abs(np.mean(eps**4) - 3) < 0.15
    a = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 2048, 1, 0, rng=(1234, 7), nargout=2)
    b = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 2048, 1, 0, epsilon=eps, nargout=2)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    c = orc.negelcbo_vbmc(theta, 0.0, vp, gp, 2048, 1, 0, epsilon=eps, nargout=2)
    assert rel(a[0], c[0]) < TOL and rel(a[1], c[1]) < TOL
    # a different stream gives different draws
    eps2 = gpu_ctx.eps_philox(5, 8, 2048, seed=1234, stream=8, readback=True)
    assert not np.array_equal(eps, eps2)


def test_ahead_of_time_draws_are_invisible(gpu_ctx):
    """Generator mode produces the next stream's draws ahead of time once the caller is seen to advance the stream by one
    (csrc/api.cu step_with_graph): whatever the call pattern — streaming (direct launches, graph capture, graph replay),
    a repeated stream, jumps forwards and backwards — every call must equal parity mode fed with that stream's dumped draws."""
    import vbmc_b200
    w = mk(D=5, N=40, K=8, S=2, Ns=256)
    vp, gp, theta = w["vp"], w["gp"], w["theta"]
    streams = [7, 8, 9, 10, 11, 12, 13, 13, 40, 41, 42, 43, 5, 6, 6, 7]
    ref = {}
    for t in sorted(set(streams)):
        eps = gpu_ctx.eps_philox(5, 8, 256, seed=77, stream=t, readback=True)
        ref[t] = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 256, 1, 0, epsilon=eps, nargout=2)
    for i, t in enumerate(streams):
        th = theta + 1e-3 * i          # the parameters move while the draws are being prefetched
        a = vbmc_b200.negelcbo_vbmc(th, 0.0, vp, gp, 256, 1, 0, rng=(77, t), nargout=2)
        eps = gpu_ctx.eps_philox(5, 8, 256, seed=77, stream=t, readback=True) if i % 5 == 4 else None   # also disturbs the buffer
        b = vbmc_b200.negelcbo_vbmc(th, 0.0, vp, gp, 256, 1, 0, epsilon=eps, nargout=2) if eps is not None else None
        if b is not None:
            assert a[0] == b[0] and np.array_equal(a[1], b[1]), (i, t)
        if i == 0:
            assert a[0] == ref[t][0] and np.array_equal(a[1], ref[t][1])
    # same sequence, fixed theta: every call equals its parity-mode reference bit for bit
    for t in streams:
        a = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 256, 1, 0, rng=(77, t), nargout=2)
        assert a[0] == ref[t][0] and np.array_equal(a[1], ref[t][1]), t
    # a different seed with the stream the buffer was prefetched for is a miss
    eps = gpu_ctx.eps_philox(5, 8, 256, seed=78, stream=8, readback=True)
    a = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 256, 1, 0, rng=(78, 8), nargout=2)
    b = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 256, 1, 0, epsilon=eps, nargout=2)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


def test_philox_known_answers(gpu_ctx):
    """Random123 known-answer vectors for philox4x32-10."""
    import ctypes as C
    from vbmc_b200 import _lib
    lib = _lib.load()

    def run(ctr, key):
        c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        _lib.check(lib.vbmc_b200_philox_raw(gpu_ctx.handle, c, k, o))
        return list(o)

    assert run([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert run([f, f, f, f], [f, f]) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert run([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_reference_error_identifiers(gpu_ctx):
    import vbmc_b200
    w = mk(D=2, N=20, K=2, S=2, Ns=10)
    with pytest.raises(vbmc_b200.VbmcB200Error) as ei:
        vbmc_b200.negelcbo_vbmc(w["theta"], 1.0, w["vp"], w["gp"], 10, 1, 1, epsilon=w["epsilon"])
    assert ei.value.identifier == "negelcbo_vbmc:vargrad"
    gp = dict(w["gp"], meanfun=7)
    with pytest.raises(vbmc_b200.VbmcB200Error) as ei:
        vbmc_b200.gplogjoint(w["vp"], gp)
    assert ei.value.identifier == "gplogjoint:UnsupportedMeanFun"
    with pytest.raises(vbmc_b200.VbmcB200Error):
        vbmc_b200.negelcbo_vbmc(w["theta"][:-1], 0.0, w["vp"], w["gp"], 10, epsilon=w["epsilon"])


def test_full_size_properties_c3(gpu_ctx):
    """c3 shape at full K, Ns (oracle too slow there): size-independent properties.
    (1) linearity in the draws: H over [eps_a; eps_b] == mean of H over eps_a and eps_b;
    (2) the eta-gradient of G+H sums to ~0 (J_w rows sum to zero);
    (3) run-to-run bit reproducibility."""
    import vbmc_b200
    w = mk(D=10, N=256, K=50, S=2, Ns=32768, target="lumpy")
    vp = w["vp"]
    rng = np.random.Generator(np.random.Philox(9))
    ea = rng.standard_normal((50, 8192, 10))
    eb = rng.standard_normal((50, 8192, 10))
    Ha, dHa = vbmc_b200.entmc_vbmc(vp, 16384, True, True, epsilon=ea)
    Hb, dHb = vbmc_b200.entmc_vbmc(vp, 16384, True, True, epsilon=eb)
    Hab, dHab = vbmc_b200.entmc_vbmc(vp, 32768, True, True, epsilon=np.concatenate([ea, eb], axis=1))
    assert abs(Hab - 0.5 * (Ha + Hb)) < 1e-12 * abs(Hab)
    assert rel(dHab, 0.5 * (dHa + dHb)) < 1e-11
    assert abs(np.sum(dHab[-50:])) < 1e-10 * np.max(np.abs(dHab[-50:]))
    H2, dH2 = vbmc_b200.entmc_vbmc(vp, 32768, True, True, epsilon=np.concatenate([ea, eb], axis=1))
    assert H2 == Hab and np.array_equal(dH2, dHab)


# ---- variance of the expected log-joint (gplogjoint.m:273-339) and the eval_fullelcbo call pattern ----
def _var_problem(K=5, S=3, N=70, D=3):
    import math as _m
    return mk(D=D, N=N, K=K, S=S, Ns=64, log_sn=_m.log(0.1))


@pytest.mark.parametrize("compute_var", [1, 2])
@pytest.mark.parametrize("source", ["attach", "refit"])
def test_gplogjoint_variance(gpu_ctx, compute_var, source):
    """varF, varss, J_sjk.  J = prior term - z K^-1 z suffers cancellation (the GP is confident near its data),
    so the comparison is relative to max|J| with a looser bound than the 1e-10 used for F and dF."""
    import vbmc_b200
    w = _var_problem()
    vp, gp = w["vp"], w["gp"]
    if source == "refit":   # factors computed by the GPU refit and left resident
        gp = vbmc_b200.gplite_post(w["hyp"], w["X"], w["y"], 1, 4, [1, 0, 0], None)
    got = vbmc_b200.gplogjoint(vp, gp, False, True, True, compute_var, nargout=7)
    ref = orc.gplogjoint(vp, w["gp"], False, True, True, compute_var, nargout=7)
    assert rel(got[0], ref[0]) < 1e-9
    assert got[6].shape == ref[6].shape == (3, 5, 5)
    assert rel(got[6], ref[6]) < 1e-7
    assert rel(got[2], ref[2]) < 1e-7 and rel(got[4], ref[4]) < 1e-7
    assert got[2] > 0


def test_eval_fullelcbo_call_pattern(gpu_ctx):
    """vpoptimize_vbmc.m:288-289: negelcbo_vbmc(theta,0,vp,gp,NSentFineK,0,1,...) with 11 outputs."""
    import vbmc_b200
    w = _var_problem(K=6, S=4)
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    got = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 64, 0, 1, epsilon=eps, nargout=11)
    ref = orc.negelcbo_vbmc(theta, 0.0, vp, gp, 64, 0, 1, epsilon=eps, nargout=11)
    F, dF, G, H, varF, dH, varGss, varG, varH, I_sk, J_sjk = got
    assert dF is None and dH is None and varH == 0.0
    assert rel(F, ref[0]) < 1e-9 and rel(G, ref[2]) < 1e-9 and rel(H, ref[3]) < TOL
    assert rel(varF, ref[4]) < 1e-7 and rel(varGss, ref[6]) < 1e-7 and rel(varG, ref[7]) < 1e-7
    assert rel(I_sk, ref[9]) < 1e-9 and rel(J_sjk, ref[10]) < 1e-7
    # beta ~= 0 without gradient: F = -G - H + beta*sqrt(varF)  (negelcbo_vbmc.m:126)
    gb = vbmc_b200.negelcbo_vbmc(theta, 1.5, vp, gp, 64, 0, 2, epsilon=eps, nargout=5)
    rb = orc.negelcbo_vbmc(theta, 1.5, vp, gp, 64, 0, 2, epsilon=eps, nargout=5)
    assert rel(gb[0], rb[0]) < 1e-8 and rel(gb[4], rb[4]) < 1e-7


@pytest.mark.parametrize("S", [1, 3])
def test_variance_gradient_elcbo(gpu_ctx, S):
    """beta ~= 0 with gradient: dF += 0.5*beta*dvarG/sqrt(varF), diagonal variance only (negelcbo_vbmc.m:21-24,128-130;
    gplogjoint.m:289-303,370-385,407-410)."""
    import vbmc_b200
    w = _var_problem(K=4, S=S, N=60, D=3)
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    got = vbmc_b200.negelcbo_vbmc(theta, 0.7, vp, gp, 64, 1, 2, 0, tb, 0, epsilon=eps, nargout=5)
    ref = orc.negelcbo_vbmc(theta, 0.7, vp, gp, 64, 1, 2, 0, tb, 0, epsilon=eps, nargout=5)
    assert rel(got[0], ref[0]) < 1e-8 and rel(got[4], ref[4]) < 1e-7
    assert rel(got[1], ref[1]) < 1e-7
    # gplogjoint's own 4th output
    v = dict(vp)
    g = vbmc_b200.gplogjoint(v, gp, True, True, True, 2, nargout=5)
    r = orc.gplogjoint(v, gp, True, True, True, 2, nargout=5)
    assert rel(g[3], r[3]) < 1e-6 and rel(g[2], r[2]) < 1e-7 and rel(g[4], r[4]) < 1e-7
    with pytest.raises(vbmc_b200.VbmcB200Error) as ei:
        vbmc_b200.gplogjoint(v, gp, True, True, True, 1, nargout=4)
    assert ei.value.identifier == "gplogjoint:FullVarianceGradient"


def _low_noise_problem(K=5, S=3, N=70, D=3):
    """GP whose noise variance is below 1e-6: gplite_core takes the branch post.L = -inv(K + diag), Lchol = 0
    (gplite_core.m:86-100).  Short length scales keep K + diag well conditioned so J_sjk is not round-off noise."""
    w = _var_problem(K=K, S=S, N=N, D=D)
    hyp = w["hyp"].copy()
    hyp[:D] -= 1.5
    hyp[D + 1] = 0.5 * math.log(2e-7)
    gp = orc.gplite_post(hyp, w["X"], w["y"], 1, 4, [1, 0, 0], None)
    assert not any(p["Lchol"] for p in gp["post"])
    return w, hyp, gp


@pytest.mark.parametrize("compute_var", [1, 2])
@pytest.mark.parametrize("source", ["attach", "refit"])
def test_gplogjoint_variance_low_noise_posterior(gpu_ctx, compute_var, source):
    """Lchol == 0: K^-1 z_k = -L z_k with L = -inv (gplogjoint.m:279, 325).  'attach' multiplies by the matrix the
    host hands over; 'refit' keeps the Cholesky factor of K + sn2_mult*diag(sn2) resident and solves with it."""
    import vbmc_b200
    w, hyp, gp_ref = _low_noise_problem()
    gp = gp_ref if source == "attach" else vbmc_b200.gplite_post(hyp, w["X"], w["y"], 1, 4, [1, 0, 0], None)
    got = vbmc_b200.gplogjoint(w["vp"], gp, False, True, True, compute_var, nargout=7)
    ref = orc.gplogjoint(w["vp"], gp_ref, False, True, True, compute_var, nargout=7)
    assert rel(got[0], ref[0]) < 1e-9
    assert rel(got[6], ref[6]) < 1e-8
    assert rel(got[2], ref[2]) < 1e-8 and rel(got[4], ref[4]) < 1e-8


def test_variance_gradient_low_noise_posterior(gpu_ctx):
    import vbmc_b200
    w, hyp, gp = _low_noise_problem(K=4, S=2, N=60)
    g = vbmc_b200.gplogjoint(w["vp"], gp, True, True, True, 2, nargout=5)
    r = orc.gplogjoint(w["vp"], gp, True, True, True, 2, nargout=5)
    assert rel(g[1], r[1]) < 1e-9 and rel(g[2], r[2]) < 1e-8 and rel(g[3], r[3]) < 1e-7


def test_variance_more_points_than_eight_columns_fit(gpu_ctx):
    """N = 3600: only 4 solution columns of V = R'\\Z fit in shared memory (the kernel is instantiated for 8/4/2/1)."""
    import vbmc_b200
    w = mk(D=2, N=3600, K=3, S=1, Ns=8, log_sn=math.log(0.3))
    got = vbmc_b200.gplogjoint(w["vp"], w["gp"], True, True, True, 2, nargout=7)
    ref = orc.gplogjoint(w["vp"], w["gp"], True, True, True, 2, nargout=7)
    # the value itself against the binary128 evaluation (the FP64 oracle is ~1e-9 away from it at N = 3600)
    t = truth_negelcbo(w["vp"], w["gp"], w["theta"], 2, np.zeros((3, 1, 2)), None, compute_grad=False)
    assert rel(got[0], t["G"]) < TOL, (rel(got[0], t["G"]), rel(ref[0], t["G"]))
    # J = prior term (O(1)) - z K^-1 z cancels four digits here, on top of the conditioning of a 3600-point Gram matrix
    assert rel(got[2], ref[2]) < 1e-5 and rel(got[6], ref[6]) < 1e-5
    assert rel(got[3], ref[3]) < 1e-4


def test_entmc_pruning_is_invisible(gpu_ctx):
    """Components that a whole warp of draws cannot see (below exp(-50) of q) are skipped; the result must equal the
    un-pruned sweep to round-off and the oracle to the parity bar, and the counters must show that pruning happened."""
    import vbmc_b200
    w = mk(D=10, N=120, K=50, S=2, Ns=1024, target="lumpy")
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    vp = dict(vp)
    vp["sigma"] = vp["sigma"] * 0.5          # well separated components: most cross terms underflow
    theta = workloads.theta_of(vp)
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    ref = orc.negelcbo_vbmc(theta, 0.0, vp, gp, 1024, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
    try:
        gpu_ctx.entmc_prune(0.0)
        gpu_ctx.entmc_prune_stats(True)
        full = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 1024, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
        kept0, tot0 = gpu_ctx.entmc_prune_stats(True)
        gpu_ctx.entmc_prune(50.0)
        pruned = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 1024, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
        kept1, tot1 = gpu_ctx.entmc_prune_stats(False)
    finally:
        gpu_ctx.entmc_prune(50.0)
        gpu_ctx.entmc_prune_stats(False)
    assert kept0 == tot0 and tot1 == tot0 and kept1 < 0.9 * tot1
    for i in (0, 1, 3, 5):   # F, dF, H, dH
        assert rel(pruned[i], full[i]) < 1e-14
        assert rel(pruned[i], ref[i]) < TOL


def test_entmc_pruning_overlapping_components_keeps_everything(gpu_ctx):
    """Heavily overlapping components: nothing may be skipped."""
    import vbmc_b200
    w = mk(D=3, N=40, K=9, S=2, Ns=256)
    vp = dict(w["vp"])
    vp["mu"] = 0.05 * vp["mu"]
    try:
        gpu_ctx.entmc_prune_stats(True)
        H, dH = vbmc_b200.entmc_vbmc(vp, 256, epsilon=w["epsilon"])
        kept, tot = gpu_ctx.entmc_prune_stats(False)
    finally:
        gpu_ctx.entmc_prune_stats(False)
    Ho, dHo = orc.entmc_vbmc(vp, 256, epsilon=w["epsilon"])
    assert kept == tot and rel(H, Ho) < TOL and rel(dH, dHo) < TOL
