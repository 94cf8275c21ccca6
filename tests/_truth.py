"""Helpers shared by the parity tests: the IEEE binary128 evaluation of negelcbo_vbmc (oracle/c/vbmc_oracle.c compiled with
-DVBMC_ORACLE_QUAD, see oracle/cport.py) on a workload dict, and the max-norm relative error the FP64 gate is stated in."""
import numpy as np

from oracle import cport


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def truth_negelcbo(vp, gp, theta, Ns, eps, thetabnd, compute_grad=True):
    """dict(F, dF, G, H, dH, I_sk): every intermediate in binary128, results rounded to double."""
    prep = cport.Prepared(vp, gp, thetabnd)
    F, dF, G, H, dH, Isk = cport.negelcbo(prep, theta, Ns, eps, compute_grad=compute_grad, truth128=True)
    return dict(F=F, dF=dF, G=G, H=H, dH=dH, I_sk=Isk)


def errs_vs_truth(got, truth, keys=("F", "dF", "G", "H", "dH")):
    return {k: rel(got[k], truth[k]) for k in keys if got.get(k) is not None and truth.get(k) is not None}
