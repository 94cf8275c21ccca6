"""Accuracy of the refit (factor backward error, alpha, nlZ) against the oracle on the test-suite problems; a script (python tests/refit_accuracy.py on a GPU box) used to A/B the panel variants (VBMC_B200_REFIT_PANEL_V1 / _TRSM_V1)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import vbmc_b200
from oracle import vbmc_oracle as orc
from test_gpu_gplite import problem

rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1e-300, np.max(np.abs(b))))
for (N, D, S) in ((150, 4, 7), (500, 3, 3)):
    X, y, s2, hyp = problem(N, D, S)
    hyp = hyp.copy()
    if S > 5:
        hyp[:D, 3] += 0.7
        hyp[D + 1, 5] = math.log(0.5)
    ref = orc.gplite_post(hyp, X, y, 1, 4, [1, 0, 0], None)
    gp = vbmc_b200.gplite_post(hyp, X, y, 1, 4, [1, 0, 0], None, want_L=True)
    got = vbmc_b200.gplite_nlZ_batch(hyp, ref, None)
    exp = np.array([orc.gplite_nlZ(hyp[:, s], ref, None, nargout=1)[0] for s in range(S)])
    for s in range(S):
        h = hyp[:, s]
        ell, sf2, sn2 = np.exp(h[:D]), math.exp(2 * h[D]), math.exp(2 * h[D + 1])
        Aex = (sf2 * np.exp(-0.5 * orc.sq_dist((X / ell).T))) / (sn2 * ref["post"][s]["sn2_mult"]) + np.eye(N)
        be = lambda L: float(np.max(np.abs(L.T @ L - Aex)) / np.max(np.abs(Aex)))
        print("   backward error |R'R - A|/|A|: cuda %.2e  oracle %.2e" % (be(gp["post"][s]["L"]), be(ref["post"][s]["L"])))
        A = ref["post"][s]["L"].T @ ref["post"][s]["L"]
        print(N, s, "cond %.2e" % np.linalg.cond(A), "L %.2e" % rel(gp["post"][s]["L"], ref["post"][s]["L"]),
              "alpha %.2e" % rel(gp["post"][s]["alpha"], ref["post"][s]["alpha"]), "nlZ %.2e" % (abs(got[s] - exp[s]) / abs(exp[s])), flush=True)
