"""Parity against the REAL reference.  bench/matlab/dump_reference_vectors.m, run once on a box that has MATLAB, writes
tests/golden/reference_<case>.mat — outputs of the unmodified acerbilab/vbmc code on the committed inputs
(tests/golden/matlab_inputs/, entropy draws injected through a randn override).  With those files present this test pins
the NumPy oracle — and with -m gpu the CUDA path — to the reference at the north-star tolerance; without them every case
is SKIPPED, and the oracle stays "parity unpinned" (DESIGN.md 2).  VBMC_B200_REFERENCE_DIR overrides the directory."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import vbmc_oracle as orc
from vbmc_b200 import workloads

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REFDIR = os.environ.get("VBMC_B200_REFERENCE_DIR", GOLD)
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
CASES = ["c1_rosenbrock_D2", "ragged_D3", "k1_closed_form_D4", "c2_reduced_D6"]
TOL = 1e-10      # BASELINE.json north_star: FP64 within 1e-10 relative
TOL_VAR = 1e-7   # variances: J = prior term - z'K^-1 z cancels (the GP is confident near its data)


def rel(a, b):
    a, b = np.asarray(a, dtype=float).ravel(), np.asarray(b, dtype=float).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def load(name):
    path = os.path.join(REFDIR, f"reference_{name}.mat")
    if not os.path.exists(path):
        pytest.skip(f"{path} not present: generate it with bench/matlab/dump_reference_vectors.m on a box that has MATLAB")
    from scipy.io import loadmat
    ref = {k: np.squeeze(v) for k, v in loadmat(path).items() if not k.startswith("__")}
    shape, w = make_golden.build(name)
    theta = np.load(os.path.join(GOLD, name + ".npz"))["theta"]
    return ref, shape, w, theta


def unpacked(vp, theta):
    D, K = vp["D"], vp["K"]
    v = dict(vp)
    v["mu"] = theta[:D * K].reshape(K, D).T.copy()
    v["sigma"] = np.exp(theta[D * K:D * K + K])
    v["lambda"] = np.exp(theta[D * K + K:D * K + K + D])
    v["eta"] = theta[-K:].copy()
    e = np.exp(v["eta"])
    v["w"] = e / e.sum()
    return v


def compare(ref, shape, w, theta, fns):
    """fns: the implementation under test (oracle module functions or the vbmc_b200 mirrors)."""
    vp, gp, eps, Ns = w["vp"], w["gp"], w["epsilon"], shape["Ns"]
    _, tb = fns["vpbounds"](vp, gp, workloads.VP_OPTIONS)
    assert rel(tb["lb"], ref["thetabnd_lb"]) < 1e-14 and rel(tb["ub"], ref["thetabnd_ub"]) < 1e-14
    alpha = np.stack([p["alpha"] for p in gp["post"]], axis=1)
    assert rel(alpha, ref["alpha"]) < 1e-8                      # chol of the reference vs LAPACK here
    F, dF, G, H, _, dH = fns["negelcbo"](theta, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)[:6]
    assert rel(F, ref["F"]) < TOL and rel(dF, ref["dF"]) < TOL and rel(G, ref["G"]) < TOL
    assert rel(H, ref["H"]) < TOL and rel(dH, ref["dH"]) < TOL
    vpt = unpacked(vp, theta)
    He, dHe = fns["entmc"](vpt, Ns, [1, 1, 1, 1], True, epsilon=eps, nargout=2)
    assert rel(He, ref["H_entmc"]) < TOL and rel(dHe, ref["dH_entmc"]) < TOL
    Gg, dGg, vG, dvG, vss = fns["gplogjoint"](vpt, gp, [1, 1, 1, 1], True, True, 2, nargout=5)[:5]
    assert rel(Gg, ref["G_glj"]) < TOL and rel(dGg, ref["dG_glj"]) < TOL
    assert rel(vG, ref["varG_diag"]) < TOL_VAR and rel(vss, ref["varss_diag"]) < TOL_VAR and rel(dvG, ref["dvarG_diag"]) < 1e-6
    full = fns["gplogjoint"](vpt, gp, [0, 0, 0, 0], True, True, 1, True, nargout=7)
    assert rel(full[2], ref["varG_full"]) < TOL_VAR and rel(full[5], ref["I_sk"]) < TOL
    Jref = np.asarray(ref["J_sjk"], dtype=float).reshape(np.asarray(full[6]).shape)
    assert np.max(np.abs(np.asarray(full[6]) - Jref)) < TOL_VAR * max(1e-300, np.max(np.abs(Jref)))
    Hl, dHl = fns["entlb"](vpt, [1, 1, 1, 1], True)
    assert rel(Hl, ref["H_lb"]) < TOL and rel(dHl, ref["dH_lb"]) < TOL
    Fl, dFl = fns["negelcbo"](theta, 0.0, vp, gp, 0, 1, 0, 0, tb, 0, nargout=2)[:2]
    assert rel(Fl, ref["F_lb"]) < TOL and rel(dFl, ref["dF_lb"]) < TOL


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_the_reference(name):
    ref, shape, w, theta = load(name)
    compare(ref, shape, w, theta, dict(vpbounds=orc.vpbounds, negelcbo=orc.negelcbo_vbmc, entmc=orc.entmc_vbmc,
                                       gplogjoint=orc.gplogjoint, entlb=orc.entlb_vbmc))
    gp1 = orc.gplite_post(w["hyp"][:, :1], w["X"], w["y"], 1, 4, [1, 0, 0], None)
    nlZ, dnlZ = orc.gplite_nlZ(w["hyp"][:, 0], gp1, None, nargout=2)[:2]
    assert rel(nlZ, ref["nlZ"]) < 1e-9 and rel(dnlZ, ref["dnlZ"]) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_matches_the_reference(gpu_ctx, name):
    import vbmc_b200
    ref, shape, w, theta = load(name)
    compare(ref, shape, w, theta, dict(vpbounds=vbmc_b200.vpbounds, negelcbo=vbmc_b200.negelcbo_vbmc, entmc=vbmc_b200.entmc_vbmc,
                                       gplogjoint=vbmc_b200.gplogjoint, entlb=vbmc_b200.entlb_vbmc))


def test_harness_self_check(tmp_path, monkeypatch):
    """NOT a parity claim: writes a .mat with the ORACLE's own outputs in the shapes MATLAB would save them (column vectors,
    S x K, S x K x K) and runs the comparison on it, so that the day real reference files arrive the plumbing (names, shapes,
    orientation, savemat/loadmat round trip) is known to work."""
    from scipy.io import savemat
    name = "ragged_D3"
    shape, w = make_golden.build(name)
    theta = np.load(os.path.join(GOLD, name + ".npz"))["theta"]
    vp, gp, eps, Ns = w["vp"], w["gp"], w["epsilon"], shape["Ns"]
    _, tb = orc.vpbounds(vp, gp, workloads.VP_OPTIONS)
    F, dF, G, H, _, dH = orc.negelcbo_vbmc(theta, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)[:6]
    vpt = unpacked(vp, theta)
    He, dHe = orc.entmc_vbmc(vpt, Ns, [1, 1, 1, 1], True, epsilon=eps, nargout=2)
    Gg, dGg, vG, dvG, vss = orc.gplogjoint(vpt, gp, [1, 1, 1, 1], True, True, 2, nargout=5)[:5]
    full = orc.gplogjoint(vpt, gp, [0, 0, 0, 0], True, True, 1, True, nargout=7)
    Hl, dHl = orc.entlb_vbmc(vpt, [1, 1, 1, 1], True)
    Fl, dFl = orc.negelcbo_vbmc(theta, 0.0, vp, gp, 0, 1, 0, 0, tb, 0, nargout=2)[:2]
    gp1 = orc.gplite_post(w["hyp"][:, :1], w["X"], w["y"], 1, 4, [1, 0, 0], None)
    nlZ, dnlZ = orc.gplite_nlZ(w["hyp"][:, 0], gp1, None, nargout=2)[:2]
    col = lambda x: np.asarray(x, dtype=float).reshape(-1, 1)
    savemat(tmp_path / f"reference_{name}.mat", dict(
        alpha=np.stack([p["alpha"] for p in gp["post"]], axis=1), thetabnd_lb=np.atleast_2d(tb["lb"]), thetabnd_ub=np.atleast_2d(tb["ub"]),
        F=F, dF=col(dF), G=G, H=H, dH=col(dH), H_entmc=He, dH_entmc=col(dHe), G_glj=Gg, dG_glj=col(dGg), varG_diag=vG,
        dvarG_diag=col(dvG), varss_diag=vss, varG_full=full[2], varss_full=full[4], I_sk=full[5], J_sjk=full[6], H_lb=Hl, dH_lb=col(dHl),
        F_lb=Fl, dF_lb=col(dFl), nlZ=nlZ, dnlZ=col(dnlZ)))
    import sys
    monkeypatch.setattr(sys.modules[__name__], "REFDIR", str(tmp_path))
    test_oracle_matches_the_reference(name)
