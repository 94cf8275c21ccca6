"""The FP64 gate (1e-10 relative, BASELINE.json north_star) measured against the TRUTH: an IEEE binary128 evaluation of the
reference's formulas (oracle/c/vbmc_oracle.c compiled with -DVBMC_ORACLE_QUAD) on the same double inputs.

Why: for VBMC's default noise floor (sn2 = 1e-5, vbmc.m:307) the GP weights alpha are ~1e4 while the expected log-joint is O(1);
sum_n z_n alpha_n (misc/gplogjoint.m:167-169) cancels 5-8 digits, so ANY plain FP64 evaluation -- the reference's own, the NumPy
oracle, the C port -- is 1e-11..5e-9 away from the exact value of the formula, and two of them differ by as much (VERDICT r1
weak #1).  The CUDA path evaluates the terms and sums in two-word arithmetic (csrc/dd_math.cuh) and must hit the truth to 1e-10
on such posteriors; the FP64 oracles' own distance to the truth is REPORTED (and only sanity-bounded)."""
import numpy as np
import pytest

from oracle import cport
from oracle import vbmc_oracle as orc
from vbmc_b200 import api, workloads

from _truth import errs_vs_truth, rel, truth_negelcbo

TOL = 1e-10
FAMILY = [(N, seed) for N in (40, 60, 200) for seed in range(500, 507)]   # 21 ill-conditioned posteriors (sn2 = 1e-5)


def family_problem(N, seed, D=2, K=3, S=3, Ns=64):
    cfg = dict(workloads.CONFIGS["c1"])
    cfg.update(D=D, K=K, Ns=Ns, S=S, N=N)
    w = workloads.build(cfg, orc.gplite_post, seeds=(seed, seed + 1, seed + 2, seed + 3))
    _, tb = api.vpbounds(w["vp"], w["gp"], workloads.VP_OPTIONS)
    return w, tb, cfg


def test_truth128_matches_double_port_on_a_well_conditioned_problem():
    """log_sn = 0 (sn2 = 1 instead of 1e-5): little cancellation, so the binary128 and the double build of the same source agree to 1e-12."""
    cfg = dict(D=3, N=50, K=4, S=2, Ns=32, target="rosenbrock", noisy=False, log_sn=0.0)
    w = workloads.build(cfg, orc.gplite_post, seeds=(11, 12, 13, 14))
    _, tb = api.vpbounds(w["vp"], w["gp"], workloads.VP_OPTIONS)
    t = truth_negelcbo(w["vp"], w["gp"], w["theta"], cfg["Ns"], w["epsilon"], tb)
    prep = cport.Prepared(w["vp"], w["gp"], tb)
    F, dF, G, H, dH, _ = cport.negelcbo(prep, w["theta"], cfg["Ns"], w["epsilon"])
    e = errs_vs_truth(dict(F=F, dF=dF, G=G, H=H, dH=dH), t)
    assert max(e.values()) < 1e-11, e   # (alpha is still ~1e2 here: Rosenbrock values span thousands)


def test_fp64_oracles_distance_to_truth_is_reported(capsys):
    """The reference-faithful FP64 evaluations against the truth on the ill-conditioned family: this is the noise floor of any
    'match the reference to 1e-10' statement.  Entropy side: exact to round-off; log-joint side: up to ~5e-9."""
    worst = dict(np={}, c={})
    for N, seed in FAMILY[::3]:
        w, tb, cfg = family_problem(N, seed)
        t = truth_negelcbo(w["vp"], w["gp"], w["theta"], cfg["Ns"], w["epsilon"], tb)
        o = orc.negelcbo_vbmc(w["theta"], 0.0, w["vp"], w["gp"], cfg["Ns"], 1, 0, 0, tb, 0, epsilon=w["epsilon"], nargout=6)
        eo = errs_vs_truth(dict(F=o[0], dF=o[1], G=o[2], H=o[3], dH=o[5]), t)
        prep = cport.Prepared(w["vp"], w["gp"], tb)
        F, dF, G, H, dH, _ = cport.negelcbo(prep, w["theta"], cfg["Ns"], w["epsilon"])
        ec = errs_vs_truth(dict(F=F, dF=dF, G=G, H=H, dH=dH), t)
        for k in eo:
            worst["np"][k] = max(worst["np"].get(k, 0.0), eo[k])
            worst["c"][k] = max(worst["c"].get(k, 0.0), ec[k])
    with capsys.disabled():
        print("\n[truth128] worst FP64-oracle distance to the binary128 truth on the sn2=1e-5 family:", worst)
    for impl in worst.values():
        assert impl["H"] < 1e-13 and impl["dH"] < 1e-12          # no cancellation on the entropy side
        assert impl["G"] < 1e-6 and impl["dF"] < 1e-6            # sanity only: 5-8 digits cancel


@pytest.mark.gpu
@pytest.mark.parametrize("N,seed", FAMILY, ids=lambda v: str(v))
def test_cuda_hits_truth_on_ill_conditioned_posteriors(gpu_ctx, N, seed):
    import vbmc_b200
    w, tb, cfg = family_problem(N, seed)
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    t = truth_negelcbo(vp, gp, theta, cfg["Ns"], eps, tb)
    F, dF, G, H, _, dH = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
    e = errs_vs_truth(dict(F=F, dF=dF, G=G, H=H, dH=dH), t)
    assert max(e.values()) < TOL, e
    # the log-joint side is expected to be far inside the gate, not just under it
    assert e["G"] < 1e-13 and e["dF"] < 1e-12, e


@pytest.mark.gpu
def test_cuda_hits_truth_on_the_smoke_shape(gpu_ctx):
    """The exact problem __graft_entry__.smoke() runs (red in round 1: dF 1.8e-10 against the NumPy oracle)."""
    import vbmc_b200
    w, tb, cfg = family_problem(40, 101)
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    t = truth_negelcbo(vp, gp, theta, cfg["Ns"], eps, tb)
    F, dF, G, H = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0, epsilon=eps, nargout=4)
    e = errs_vs_truth(dict(F=F, dF=dF, G=G, H=H), t)
    assert max(e.values()) < TOL, e


@pytest.mark.gpu
@pytest.mark.parametrize("name,over", [
    ("c2", dict(Ns=256)),                                   # D=6, N=400, K=20, S=8
    ("c3", dict(Ns=128)),                                   # D=10, N=2000, K=50, S=20: the benchmark's log-joint in full
    ("c5", dict(Ns=64, S=4, noisy=False)),                  # D=20, N=4000, K=100: two lanes per training point
    ("c3", dict(Ns=64, S=2, N=1500, D=13, K=7)),            # D in (12, 16]
], ids=["c2", "c3_full_glj", "c5_D20", "D13"])
def test_cuda_hits_truth_at_benchmark_shapes(gpu_ctx, name, over):
    import vbmc_b200
    w = workloads.build(name, vbmc_b200.gplite_post, overrides=over)
    vp, gp, theta, eps, cfg = w["vp"], w["gp"], w["theta"], w["epsilon"], w["cfg"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    t = truth_negelcbo(vp, gp, theta, cfg["Ns"], eps, tb)
    F, dF, G, H, _, dH = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
    e = errs_vs_truth(dict(F=F, dF=dF, G=G, H=H, dH=dH), t)
    assert max(e.values()) < TOL, e


@pytest.mark.gpu
@pytest.mark.parametrize("min_points,warps_per_sm", [(32, 12), (100, 12), (777, 3), (5000, 12), (10 ** 9, 12)])
def test_segmentation_of_the_training_set_does_not_change_the_result(min_points, warps_per_sm):
    """The (pair, n) space is cut into contiguous ranges, one warp each (csrc/gplogjoint.cu); VBMC_B200_GLJ_MIN_POINTS /
    VBMC_B200_GLJ_WARPS_PER_SM force other cuts -- from 32-point segments (up to 25 per pair) to ONE warp for everything: the
    two-word segment partials make the result independent of the cut to 1e-14, and every cut hits the binary128 truth."""
    import json
    import os
    import subprocess
    import sys
    code = (
        "import json, sys, numpy as np, vbmc_b200\n"
        "sys.path.insert(0, 'tests')\n"
        "from vbmc_b200 import workloads\n"
        "from _truth import truth_negelcbo, errs_vs_truth\n"
        "w = workloads.build('c2', vbmc_b200.gplite_post, overrides=dict(Ns=64, N=777))\n"
        "_, tb = vbmc_b200.vpbounds(w['vp'], w['gp'], workloads.VP_OPTIONS)\n"
        "F, dF, G, H = vbmc_b200.negelcbo_vbmc(w['theta'], 0.0, w['vp'], w['gp'], 64, 1, 0, 0, tb, 0, epsilon=w['epsilon'], nargout=4)\n"
        "t = truth_negelcbo(w['vp'], w['gp'], w['theta'], 64, w['epsilon'], tb)\n"
        "print(json.dumps(dict(F=F, G=G, dF=list(map(float, dF)), err=errs_vs_truth(dict(F=F, dF=dF, G=G, H=H), t))))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for forced in (False, True):
        env = dict(os.environ, PYTHONPATH=root)
        env.pop("VBMC_B200_GLJ_MIN_POINTS", None)
        env.pop("VBMC_B200_GLJ_WARPS_PER_SM", None)
        if forced:
            env["VBMC_B200_GLJ_MIN_POINTS"] = str(min_points)
            env["VBMC_B200_GLJ_WARPS_PER_SM"] = str(warps_per_sm)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert rel(outs[1]["dF"], outs[0]["dF"]) < 1e-14 and abs(outs[1]["G"] - outs[0]["G"]) < 1e-14 * abs(outs[0]["G"])
    assert max(outs[1]["err"].values()) < TOL, outs[1]["err"]
