"""vbmc_b200/csrc/dd_math.cuh on the CPU: the two-word exponential and the term sums of the expected-log-joint kernel
(misc/gplogjoint.m:164-252) against 60-digit mpmath evaluations of the same formulas on the same double inputs."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import mpmath as mp
import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
mp.mp.dps = 60
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("ddm") / "libdd_math_host.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-Wall", "-Wextra", "-shared", "-fPIC",
                           "-o", str(so), str(ROOT / "tests" / "host_harness" / "dd_math_host.cpp")])
    L = C.CDLL(str(so))
    L.dd_exp_host.argtypes = [C.c_double, C.c_double, dp, dp]
    L.dd_glj_sums_host.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, dp]
    return L


def test_exp_dd_relative_error(lib):
    r = np.random.default_rng(5)
    # (below ~-660 the low word is subnormal: such terms are < 1e-286 and irrelevant to any sum)
    args = np.concatenate([r.uniform(-650, 60, 4000), r.uniform(-1, 1, 2000), r.uniform(-1e-3, 1e-3, 500), [0.0, 700.0, -1e-300]])
    worst = 0.0
    for a in args:
        al = float(r.uniform(-1, 1)) * abs(a) * 2.0 ** -53
        eh, el = C.c_double(), C.c_double()
        lib.dd_exp_host(float(a), al, C.byref(eh), C.byref(el))
        truth = mp.e ** (mp.mpf(float(a)) + mp.mpf(al))
        err = abs((mp.mpf(eh.value) + mp.mpf(el.value)) / truth - 1)
        worst = max(worst, float(err))
    assert worst < 2e-23, worst
    eh, el = C.c_double(), C.c_double()
    lib.dd_exp_host(-800.0, 0.0, C.byref(eh), C.byref(el))
    assert eh.value == 0.0 and el.value == 0.0
    lib.dd_exp_host(float("nan"), 0.0, C.byref(eh), C.byref(el))
    assert np.isnan(eh.value)


@pytest.mark.parametrize("N,D,amp", [(40, 2, 1e4), (200, 3, 1e4), (300, 10, 3e5), (64, 24, 1e3)])
def test_term_sums_against_mpmath(lib, N, D, amp):
    r = np.random.default_rng(N + D)
    X = r.standard_normal((N, D)) * 1.5
    mu = r.standard_normal(D) * 0.5
    tau = np.exp(0.3 * r.standard_normal(D))
    itau = 1.0 / tau
    # alternating-sign weights of large magnitude: the sums cancel to ~1e-6 of sum |zeta|
    alpha = amp * r.standard_normal(N) * np.where(np.arange(N) % 2, 1.0, -1.0)
    lnnf = 0.37
    out = np.zeros(2 + 4 * D)
    Xc = np.ascontiguousarray(X.T)
    lib.dd_glj_sums_host(N, D, mu.ctypes.data_as(dp), itau.ctypes.data_as(dp), Xc.ctypes.data_as(dp), alpha.ctypes.data_as(dp), lnnf,
                         out.ctypes.data_as(dp))
    A = mp.mpf(0)
    B = [mp.mpf(0)] * D
    Q = [mp.mpf(0)] * D
    absA = mp.mpf(0)
    for n in range(N):
        dl = [(mp.mpf(float(mu[d])) - mp.mpf(float(X[n, d]))) * mp.mpf(float(itau[d])) for d in range(D)]
        z = mp.e ** (mp.mpf(lnnf) - sum(v * v for v in dl) / 2) * mp.mpf(float(alpha[n]))
        A += z
        absA += abs(z)
        B = [B[d] + z * dl[d] for d in range(D)]
        Q = [Q[d] + z * dl[d] ** 2 for d in range(D)]
    got_A = mp.mpf(out[0]) + mp.mpf(out[1])
    assert abs(got_A - A) < 1e-21 * absA
    for d in range(D):
        assert abs(mp.mpf(out[2 + d]) + mp.mpf(out[2 + D + d]) - B[d]) < 1e-21 * absA * 10
        assert abs(mp.mpf(out[2 + 2 * D + d]) + mp.mpf(out[2 + 3 * D + d]) - Q[d]) < 1e-21 * absA * 100
