"""The "multi" variant of the expected-log-joint contraction (vbmc_b200/csrc/glj_multi.cuh; off by default, written without
GPU time left) run on the CPU through the thread-per-CUDA-thread shim (tests/host_harness/cuda_shim.h): every chunk partial
it writes, summed over the chunks, must equal the plain sums  A = sum zeta_n,  B_d = sum zeta_n Delta_dn,
C_d = sum zeta_n (Delta_dn^2 - 1)  of misc/gplogjoint.m:164-252 for every (s, k) — ragged N, K not a multiple of the group
size, a shard of the samples, padded dimensions."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("gljm") / "libglj_multi_host.so"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-Wall", "-Wextra", "-Wno-unknown-pragmas", "-Wl,-Bsymbolic", "-shared", "-fPIC", "-o", str(so),
                           str(ROOT / "tests" / "host_harness" / "glj_multi_host.cpp")])
    lib = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)
    lib.glj_multi_host.argtypes = [C.c_int] * 8 + [dp] * 9
    lib.glj_multi_host.restype = C.c_int
    return lib


def reference_sums(X, alpha, ell, lnc, mu, sigma, lam, delta, s_list):
    N, D = X.shape
    K = mu.shape[0]
    out = np.zeros((len(s_list), K, 1 + 2 * D))
    for i, s in enumerate(s_list):
        for k in range(K):
            tau = np.sqrt(sigma[k] ** 2 * lam ** 2 + ell[s] ** 2 + delta ** 2)
            lnnf = lnc[s] - np.sum(np.log(tau))
            dl = (mu[k][None, :] - X) / tau[None, :]
            z = np.exp(lnnf - 0.5 * np.sum(dl * dl, axis=1)) * alpha[s]
            out[i, k, 0] = z.sum()
            out[i, k, 1:1 + D] = (z[:, None] * dl).sum(axis=0)
            out[i, k, 1 + D:] = (z[:, None] * (dl * dl - 1)).sum(axis=0)
    return out


@pytest.mark.parametrize("N,D,K,S,s_begin,s_count,kg,DP,P", [
    (700, 3, 7, 3, 0, 3, 3, 4, 4),       # ragged last chunk (700 = 512 + 188), K = 2 groups of 3 + 1, D padded to 4
    (130, 2, 5, 2, 1, 1, 10, 2, 4),      # one chunk mostly empty, group larger than K, a shard that starts at sample 1
    (1100, 10, 12, 2, 0, 2, 5, 10, 4),   # c3's dimension
    (300, 4, 6, 1, 0, 1, 4, 4, 2),       # two points per thread
])
def test_multi_variant_matches_plain_sums(harness, N, D, K, S, s_begin, s_count, kg, DP, P):
    r = np.random.Generator(np.random.Philox(N + D))
    X = r.standard_normal((N, D))
    alpha = r.standard_normal((S, N))
    ell = np.exp(0.3 * r.standard_normal((S, D)))
    lnc = r.standard_normal(S)
    mu = r.standard_normal((K, D))
    sigma = np.exp(0.2 * r.standard_normal(K))
    lam = np.exp(0.2 * r.standard_normal(D))
    delta = np.zeros(D)
    V = 1 + 2 * D
    nchunks = (N + P * 128 - 1) // (P * 128)
    part = np.full(nchunks * V * s_count * K, np.nan)
    Xc = np.ascontiguousarray(X.T)           # [D][N]
    d = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(C.c_double))
    keep = [np.ascontiguousarray(v) for v in (alpha, ell, lnc, mu, sigma, lam, delta)]
    got = harness.glj_multi_host(N, D, K, s_begin, s_count, kg, DP, P, d(Xc), *[d(v) for v in keep], d(part))
    assert got == nchunks
    assert not np.isnan(part).any()           # every (chunk, value, pair) slot was written exactly by its owner
    sums = part.reshape(nchunks, V, s_count * K).sum(axis=0)            # [value][pair], pair = sl*K + k
    ref = reference_sums(X, alpha, ell, lnc, mu, sigma, lam, delta, list(range(s_begin, s_begin + s_count)))
    for sl in range(s_count):
        for k in range(K):
            a, b = sums[:, sl * K + k], ref[sl, k]
            assert np.max(np.abs(a - b)) <= 1e-12 * max(1.0, np.max(np.abs(b))), (sl, k)
