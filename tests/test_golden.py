"""Committed vectors (tests/golden/*.npz, made by tests/golden/make_golden.py — read its header for what they are):
the NumPy oracle, the C/OpenMP port (here) and the CUDA path (tests/test_zz_golden_gpu.py) must all reproduce them; the K = 1 file also carries closed-form
known answers derived from the reference's formulas (ent/entmc_vbmc.m:60-67,82-88)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import cport
from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
NAMES = sorted(make_golden.CASES)


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def load(name):
    """Fixture + the problem rebuilt from its seeds; the rebuilt inputs must be the stored ones."""
    g = np.load(os.path.join(HERE, name + ".npz"))
    shape, w = make_golden.build(name)
    eps = w["epsilon"]
    fp = np.array([eps.size, eps.sum(), np.sum(eps * eps), *eps.ravel()[:8], *eps.ravel()[-8:]])
    assert np.allclose(fp, g["eps_fingerprint"], rtol=1e-13, atol=0)
    if "epsilon" in g.files:
        assert np.array_equal(g["epsilon"], eps)
    assert np.array_equal(g["X"], w["X"]) and np.array_equal(g["hyp"], w["hyp"]) and np.allclose(g["y"], w["y"], rtol=1e-14)
    alpha = np.stack([p["alpha"] for p in w["gp"]["post"]], axis=1)
    assert rel(alpha, g["alpha"]) < 1e-6          # same problem; LAPACK builds / CPU kernels differ by cond(K) * eps
    # the posterior weights are an INPUT of the path under test: use the committed ones, so that the comparison does not
    # depend on which BLAS kernels the box at hand selects
    for s_, p_ in enumerate(w["gp"]["post"]):
        p_["alpha"] = g["alpha"][:, s_].copy()
    return g, shape, w


@pytest.mark.parametrize("name", NAMES)
def test_numpy_oracle_reproduces_golden(name):
    g, shape, w = load(name)
    vp, gp, eps, Ns = w["vp"], w["gp"], w["epsilon"], shape["Ns"]
    _, tb = orc.vpbounds(vp, gp, workloads.VP_OPTIONS)
    F, dF, G, H, _, dH = orc.negelcbo_vbmc(g["theta"], 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)[:6]
    assert rel(F, g["F"]) < 1e-10 and rel(dF, g["dF"]) < 1e-10 and rel(G, g["G"]) < 1e-10
    assert rel(H, g["H"]) < 1e-13 and rel(dH, g["dH"]) < 1e-12
    I_sk = orc.negelcbo_vbmc(g["theta"], 0.0, vp, gp, Ns, 0, 1, 0, tb, 0, epsilon=eps, nargout=11)[9]
    assert rel(I_sk, g["I_sk"]) < 1e-10


@pytest.mark.parametrize("name", NAMES)
def test_c_port_reproduces_golden(name):
    g, shape, w = load(name)
    vp, gp, eps, Ns = w["vp"], w["gp"], w["epsilon"], shape["Ns"]
    _, tb = orc.vpbounds(vp, gp, workloads.VP_OPTIONS)
    prep = cport.Prepared(vp, gp, tb)
    F, dF, G, H, dH, Isk = cport.negelcbo(prep, g["theta"], Ns, eps)
    assert rel(F, g["F"]) < 1e-10 and rel(dF, g["dF"]) < 1e-10 and rel(G, g["G"]) < 1e-10
    assert rel(H, g["H"]) < 1e-12 and rel(dH, g["dH"]) < 1e-10


def test_closed_form_known_answers():
    """K = 1: the mixture entropy estimate and its mu-gradient in closed form (independent of every implementation)."""
    g = np.load(os.path.join(HERE, "k1_closed_form_D4.npz"))
    assert abs(float(g["H"]) - float(g["closed_H"])) < 1e-12 * max(1.0, abs(float(g["closed_H"])))
    D = 4
    assert np.max(np.abs(g["dH"][:D] - g["closed_dH_mu"])) < 1e-13


@pytest.mark.parametrize("name", NAMES)
def test_golden_vectors_against_binary128_truth(name, capsys):
    """every committed vector (an FP64 oracle output, the exact inputs bench/matlab/dump_reference_vectors.m will
    feed to the real reference) against the IEEE binary128 evaluation of the same formulas (oracle/c -DVBMC_ORACLE_QUAD).  The
    entropy side must agree to round-off; the log-joint side to the FP64 noise floor of the case, which is printed: this is the
    distance the MATLAB outputs themselves are expected to have from the truth."""
    g, shape, w = load(name)
    vp, gp, eps, Ns = w["vp"], w["gp"], w["epsilon"], shape["Ns"]
    _, tb = orc.vpbounds(vp, gp, workloads.VP_OPTIONS)
    prep = cport.Prepared(vp, gp, tb)
    F, dF, G, H, dH, Isk = cport.negelcbo(prep, g["theta"], Ns, eps, truth128=True)
    e = dict(F=rel(g["F"], F), dF=rel(g["dF"], dF), G=rel(g["G"], G), H=rel(g["H"], H), dH=rel(g["dH"], dH), I_sk=rel(g["I_sk"], Isk))
    with capsys.disabled():
        print(f"\n[golden {name}] committed FP64 vector vs binary128 truth: " + ", ".join(f"{k} {v:.1e}" for k, v in e.items()))
    assert e["H"] < 1e-13 and e["dH"] < 1e-12
    assert max(e["F"], e["dF"], e["G"], e["I_sk"]) < 1e-8
