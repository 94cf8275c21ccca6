"""FP32 mode of the entropy sweep (BASELINE config 5: FP32 compute, FP64 oracle, 1e-4 relative).
Same draws, same C ABI; only vbmc_b200_set_precision(ctx, 32) differs from the FP64 parity tests."""
import math

import numpy as np
import pytest

from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads

pytestmark = pytest.mark.gpu
TOL32 = 1e-4


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def mk(D, N, K, S, Ns, seed=0, target="rosenbrock", **kw):
    cfg = dict(D=D, N=N, K=K, S=S, Ns=Ns, target=target, noisy=False, **kw)
    return workloads.build(cfg, orc.gplite_post, seeds=(seed + 1, seed + 2, seed + 3, seed + 4))


@pytest.fixture()
def ctx32(gpu_ctx):
    gpu_ctx.set_precision(32)
    yield gpu_ctx
    gpu_ctx.set_precision(64)


SHAPES = [
    dict(D=2, N=50, K=2, S=8, Ns=100),
    dict(D=1, N=20, K=1, S=1, Ns=2),
    dict(D=3, N=33, K=5, S=2, Ns=37),
    dict(D=6, N=400, K=20, S=8, Ns=4096),                       # c2
    dict(D=10, N=300, K=50, S=4, Ns=2048, target="lumpy"),      # c3 shape, reduced N/Ns/S
    dict(D=20, N=160, K=100, S=2, Ns=512, target="lumpy"),      # c5 dimension and component count
    dict(D=7, N=140, K=128, S=1, Ns=66),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "D{D}N{N}K{K}S{S}Ns{Ns}".format(**s))
def test_negelcbo_fp32_within_1e4(ctx32, shape):
    import vbmc_b200
    w = mk(**shape)
    vp, gp, theta, eps, Ns = w["vp"], w["gp"], w["theta"], w["epsilon"], shape["Ns"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    got = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
    ref = orc.negelcbo_vbmc(theta, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
    F, dF, G, H, _, dH = got
    Fo, dFo, Go, Ho, _, dHo = ref[:6]
    assert rel(H, Ho) < TOL32 and rel(dH, dHo) < TOL32
    assert rel(G, Go) < 1e-10                      # the expected log-joint stays FP64
    assert rel(F, Fo) < TOL32 and rel(dF, dFo) < TOL32
    # and it really is the FP32 sweep that ran: agreement with FP64 is not at round-off level
    if shape["K"] > 1 and shape["Ns"] > 64:
        assert rel(dH, dHo) > 1e-9


def _spread_vp(w, spread):
    """Components with very different widths: sigma_j/sigma_k up to exp(spread) -> the expanded form would carry
    intermediates of 1e4..1e6, so the kernel must fall back to subtract-then-square for the wide sources."""
    vp = dict(w["vp"])
    K = vp["mu"].shape[1]
    sig = np.exp(np.linspace(-spread / 2, spread / 2, K))[None, :] * float(np.mean(vp["sigma"]))
    vp["sigma"] = sig
    theta = w["theta"].copy()
    D = vp["mu"].shape[0]
    theta[D * K:D * K + K] = np.log(sig.ravel())
    return vp, theta


@pytest.mark.parametrize("spread", [3.0, 6.0])
def test_fp32_wide_and_narrow_components(ctx32, spread):
    import vbmc_b200
    w = mk(D=5, N=80, K=12, S=2, Ns=512)
    vp, theta = _spread_vp(w, spread)
    got = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, w["gp"], 512, 1, 0, epsilon=w["epsilon"], nargout=6)
    ref = orc.negelcbo_vbmc(theta, 0.0, vp, w["gp"], 512, 1, 0, epsilon=w["epsilon"], nargout=6)
    assert rel(got[3], ref[3]) < TOL32 and rel(got[5], ref[5]) < TOL32
    assert rel(got[0], ref[0]) < TOL32 and rel(got[1], ref[1]) < TOL32


def test_fp32_large_dimension_small_masses(ctx32):
    """D = 20 with wide components: w*nf/sigma^D ~ 1e-40 is below the FP32 range; the sweep scales by max ck."""
    import vbmc_b200
    w = mk(D=20, N=64, K=6, S=1, Ns=256, target="lumpy")
    vp = dict(w["vp"])
    vp["sigma"] = np.full_like(vp["sigma"], 40.0)
    theta = w["theta"].copy()
    theta[20 * 6:20 * 6 + 6] = math.log(40.0)
    H, dH = vbmc_b200.entmc_vbmc(vp, 256, epsilon=w["epsilon"])
    Ho, dHo = orc.entmc_vbmc(vp, 256, epsilon=w["epsilon"])
    assert np.isfinite(H) and rel(H, Ho) < TOL32 and rel(dH, dHo) < TOL32


def test_entmc_fp32_K1_closed_form(ctx32):
    """K = 1: H = D/2 log(2 pi) + D log(sigma) + sum log(lambda) + mean ||eps||^2 / 2 (entmc_vbmc.m:60-67)."""
    import vbmc_b200
    w = mk(D=4, N=30, K=1, S=1, Ns=2048)
    vp, eps = w["vp"], w["epsilon"]
    H, dH = vbmc_b200.entmc_vbmc(vp, 2048, epsilon=eps)
    D = 4
    e = np.asarray(eps).reshape(-1, D)
    expect = 0.5 * D * math.log(2 * math.pi) + D * math.log(float(vp["sigma"].ravel()[0])) + np.sum(np.log(vp["lambda"])) \
        + 0.5 * np.mean(np.sum(e * e, axis=1))
    assert abs(H - expect) / abs(expect) < 1e-6
    assert np.max(np.abs(dH[:D])) < 1e-6       # antithetic pairs cancel the mu-gradient


def test_precision_switch_restores_fp64(gpu_ctx):
    import vbmc_b200
    w = mk(D=3, N=40, K=4, S=2, Ns=256)
    ref = orc.negelcbo_vbmc(w["theta"], 0.0, w["vp"], w["gp"], 256, 1, 0, epsilon=w["epsilon"], nargout=4)
    gpu_ctx.set_precision(32)
    a = vbmc_b200.negelcbo_vbmc(w["theta"], 0.0, w["vp"], w["gp"], 256, 1, 0, epsilon=w["epsilon"], nargout=4)
    gpu_ctx.set_precision(64)
    b = vbmc_b200.negelcbo_vbmc(w["theta"], 0.0, w["vp"], w["gp"], 256, 1, 0, epsilon=w["epsilon"], nargout=4)
    assert rel(b[1], ref[1]) < 1e-10 and rel(b[0], ref[0]) < 1e-10
    assert rel(a[1], ref[1]) < TOL32
    with pytest.raises(vbmc_b200.VbmcB200Error):
        gpu_ctx.set_precision(16)


def test_fp32_device_generator_equals_parity_mode_on_dumped_draws(ctx32):
    """FP32 mode generates single-precision Box-Muller draws (4 per Philox counter) and keeps them as floats;
    dumped (widened exactly) and fed back through the parity path they must give the identical result."""
    import vbmc_b200
    w = mk(D=5, N=40, K=8, S=2, Ns=4096)
    vp, gp, theta = w["vp"], w["gp"], w["theta"]
    eps = ctx32.eps_philox(5, 8, 4096, seed=99, stream=3, readback=True)
    assert np.array_equal(eps, eps.astype(np.float32).astype(np.float64))       # floats on the device
    assert abs(eps.mean()) < 0.02 and abs(eps.std() - 1) < 0.02 and abs(np.mean(eps**3)) < 0.05
    assert abs(np.mean(eps**4) - 3) < 0.15 and np.max(np.abs(eps)) < 6.8
    a = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 4096, 1, 0, rng=(99, 3), nargout=2)
    b = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, 4096, 1, 0, epsilon=eps, nargout=2)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    c = orc.negelcbo_vbmc(theta, 0.0, vp, gp, 4096, 1, 0, epsilon=eps, nargout=2)
    assert rel(a[0], c[0]) < TOL32 and rel(a[1], c[1]) < TOL32
    # odd D*half: chunks that are not 16-byte aligned take the plain-load path
    w3 = mk(D=3, N=30, K=3, S=1, Ns=74)
    e3 = ctx32.eps_philox(3, 3, 74, seed=5, stream=1, readback=True)
    a3 = vbmc_b200.negelcbo_vbmc(w3["theta"], 0.0, w3["vp"], w3["gp"], 74, 1, 0, rng=(5, 1), nargout=2)
    b3 = vbmc_b200.negelcbo_vbmc(w3["theta"], 0.0, w3["vp"], w3["gp"], 74, 1, 0, epsilon=e3, nargout=2)
    assert a3[0] == b3[0] and np.array_equal(a3[1], b3[1])


def test_fp64_sweep_refuses_single_precision_resident_draws(gpu_ctx):
    import vbmc_b200
    w = mk(D=3, N=30, K=3, S=1, Ns=64)
    gpu_ctx.set_precision(32)
    try:
        vbmc_b200.negelcbo_vbmc(w["theta"], 0.0, w["vp"], w["gp"], 64, 1, 0, rng=(1, 1), nargout=2)
    finally:
        gpu_ctx.set_precision(64)
    with pytest.raises(vbmc_b200.VbmcB200Error):
        vbmc_b200.negelcbo_vbmc(w["theta"], 0.0, w["vp"], w["gp"], 64, 1, 0, epsilon="resident", nargout=2)
    # regenerating in FP64 mode makes them usable again
    vbmc_b200.negelcbo_vbmc(w["theta"], 0.0, w["vp"], w["gp"], 64, 1, 0, rng=(1, 1), nargout=2)
    vbmc_b200.negelcbo_vbmc(w["theta"], 0.0, w["vp"], w["gp"], 64, 1, 0, epsilon="resident", nargout=2)
