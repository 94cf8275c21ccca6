"""The element functions behind the device entlb_vbmc (vbmc_b200/csrc/entlb_math.cuh) compiled for the HOST and checked
element by element against the NumPy restatement of ent/entlb_vbmc.m (oracle.entlb_vbmc).  The CUDA kernels execute the
same functions one element per thread (csrc/entlb.cu), so this pins their arithmetic and indexing without a GPU."""
import ctypes as C
import itertools
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("entlb") / "libentlb_host.so"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-Wl,-Bsymbolic", "-shared", "-fPIC", "-o", str(so),
                           str(ROOT / "tests" / "host_harness" / "entlb_host.cpp")])
    lib = C.CDLL(str(so))
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.entlb_host.argtypes = [C.c_int, C.c_int, ip, C.c_int, dp, dp, dp, dp, dp, dp]
    lib.entlb_host.restype = C.c_int
    return lib


def run(lib, vp, gf, jac):
    D, K = vp["D"], vp["K"]
    f = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    mu = np.ascontiguousarray(np.asarray(vp["mu"], dtype=np.float64).reshape(D, K).T)   # [K][D]
    sigma, lam, w, eta = f(vp["sigma"]), f(vp["lambda"]), f(vp["w"]), f(vp["eta"])
    out = np.full(1 + D * K + 2 * K + D, np.nan)
    g = (C.c_int * 4)(*[int(bool(v)) for v in gf])
    d = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
    n = lib.entlb_host(D, K, g, int(jac), d(mu), d(sigma), d(lam), d(w), d(eta), d(out))
    return out[0], out[1:1 + n]


def mk_vp(D, K, seed):
    cfg = dict(D=D, N=max(30, K + 5), K=K, S=1, Ns=2, target="rosenbrock", noisy=False)
    return workloads.build(cfg, orc.gplite_post, seeds=(seed, seed + 1, seed + 2, seed + 3))["vp"]


@pytest.mark.parametrize("D,K", [(1, 1), (4, 1), (2, 2), (3, 5), (6, 20), (10, 50), (20, 12)])
def test_host_build_matches_oracle_all_blocks(harness, D, K):
    vp = mk_vp(D, K, 7)
    for jac in (True, False):
        H, dH = run(harness, vp, [1, 1, 1, 1], jac)
        Ho, dHo = orc.entlb_vbmc(vp, [1, 1, 1, 1], jac)
        assert abs(H - Ho) <= 1e-13 * max(1.0, abs(Ho))
        assert dH.shape == dHo.shape
        assert np.max(np.abs(dH - dHo)) <= 1e-12 * max(1.0, np.max(np.abs(dHo)))


def test_host_build_grad_flag_subsets(harness):
    vp = mk_vp(3, 4, 9)
    for gf in itertools.product([0, 1], repeat=4):
        H, dH = run(harness, vp, gf, True)
        Ho, dHo = orc.entlb_vbmc(vp, list(gf), True)
        assert abs(H - Ho) <= 1e-13 * max(1.0, abs(Ho))
        assert dH.shape == dHo.shape and (dH.size == 0 or np.max(np.abs(dH - dHo)) <= 1e-12 * max(1.0, np.max(np.abs(dHo))))
