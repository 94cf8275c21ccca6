"""CUDA leg of the committed vectors (tests/golden/*.npz; CPU legs and loader in tests/test_golden.py).
(File name: runs after the long-validated GPU tests under `pytest -x`.)"""
import os

import numpy as np
import pytest

from test_golden import HERE, NAMES, load, rel
from vbmc_b200 import workloads


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_reproduces_golden(gpu_ctx, name):
    import vbmc_b200
    g, shape, w = load(name)
    vp, gp, eps, Ns = w["vp"], w["gp"], w["epsilon"], shape["Ns"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    F, dF, G, H, varF, dH = vbmc_b200.negelcbo_vbmc(g["theta"], 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
    # the gate is against the binary128 evaluation of the same inputs; the committed FP64 vector must then agree with the CUDA
    # result to within ITS OWN distance from that truth (tests/test_golden.py prints it)
    from oracle import cport
    Ft, dFt, Gt, Ht, dHt, _ = cport.negelcbo(cport.Prepared(vp, gp, tb), g["theta"], Ns, eps, truth128=True)
    assert max(rel(F, Ft), rel(dF, dFt), rel(G, Gt), rel(H, Ht), rel(dH, dHt)) < 1e-10
    assert rel(F, g["F"]) < 1e-10 + 2 * rel(g["F"], Ft) and rel(dF, g["dF"]) < 1e-10 + 2 * rel(g["dF"], dFt)
    assert rel(G, g["G"]) < 1e-10 + 2 * rel(g["G"], Gt)
    assert rel(H, g["H"]) < 1e-10 and rel(dH, g["dH"]) < 1e-10
    if name == "k1_closed_form_D4":
        assert abs(H - float(g["closed_H"])) < 1e-10 * max(1.0, abs(float(g["closed_H"])))


@pytest.mark.gpu
def test_cuda_component_relabelling_and_antithetic_draws(gpu_ctx):
    """Size-independent properties (tests/test_properties.py) through the CUDA path."""
    import vbmc_b200
    from test_properties import check_antithetic, check_permutation
    check_permutation(vbmc_b200.negelcbo_vbmc, vbmc_b200.vpbounds, tol=1e-11)   # different summation order over k
    check_antithetic(vbmc_b200.negelcbo_vbmc, vbmc_b200.vpbounds, tol=1e-12)
