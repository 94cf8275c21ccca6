"""Three-way pin, CPU leg: the C/OpenMP restatement (oracle/c) agrees with the NumPy oracle."""
import numpy as np
import pytest

from oracle import cport
from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


@pytest.mark.parametrize("shape", [dict(D=2, N=50, K=2, S=8, Ns=100), dict(D=3, N=33, K=5, S=2, Ns=37),
                                   dict(D=6, N=120, K=9, S=3, Ns=256)])
@pytest.mark.parametrize("bounds", [True, False])
def test_c_port_matches_numpy_oracle(shape, bounds):
    cfg = dict(shape, target="rosenbrock", noisy=False)
    w = workloads.build(cfg, orc.gplite_post, seeds=(5, 6, 7, 8))
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"].copy(), w["epsilon"]
    tb = orc.vpbounds(vp, gp, workloads.VP_OPTIONS)[1] if bounds else None
    if bounds:
        theta[0] = tb["ub"][0] + 0.2
        theta[-1] = 0.3
    prep = cport.Prepared(vp, gp, tb)
    F, dF, G, H, dH, Isk = cport.negelcbo(prep, theta, shape["Ns"], eps)
    ref = orc.negelcbo_vbmc(theta, 0.0, vp, gp, shape["Ns"], 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
    assert rel(F, ref[0]) < 1e-10 and rel(dF, ref[1]) < 1e-10
    assert rel(G, ref[2]) < 1e-10 and rel(H, ref[3]) < 1e-12 and rel(dH, ref[5]) < 1e-10
