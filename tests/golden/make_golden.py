"""Generate tests/golden/*.npz — committed input/output vectors of the hot path.

WHAT THESE ARE (and are not).  The reference (acerbilab/vbmc) is MATLAB and cannot run here, and it holds no golden
vectors for this path, so these files were NOT produced by the reference: they are outputs of the line-cited NumPy
oracle (oracle/vbmc_oracle.py) on seeded inputs, frozen so that (i) an edit to the oracle that changes its results is
caught, (ii) the C/OpenMP port and the CUDA path are compared against the same committed numbers on every box
(nothing at test time reads /root/reference), and (iii) each file also carries `closed_*` entries — closed-form
known answers derived from the reference's formulas, independent of any implementation:
    K = 1:  H = D/2 log(2 pi) + D log(sigma) + sum(log lambda) + 1/2 mean ||eps||^2      (ent/entmc_vbmc.m:60-67)
            dH/dmu = 0 (antithetic pairs), dH/dlog(sigma) = mean ||eps||^2               (:82-88, :112-114)
Re-generate with:  python tests/golden/make_golden.py     (deterministic; CPU only)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import vbmc_oracle as orc  # noqa: E402
from vbmc_b200 import workloads  # noqa: E402

CASES = {
    # name: (shape, seeds)            shapes follow BASELINE.json's configs at sizes the oracle finishes in seconds
    "c1_rosenbrock_D2": (dict(D=2, N=50, K=2, S=8, Ns=100, target="rosenbrock", noisy=False), (101, 102, 103, 104)),
    "c2_reduced_D6": (dict(D=6, N=120, K=20, S=4, Ns=512, target="rosenbrock", noisy=False), (201, 202, 203, 204)),
    "c3_reduced_D10": (dict(D=10, N=200, K=50, S=3, Ns=256, target="lumpy", noisy=False), (301, 302, 303, 304)),
    "ragged_D3": (dict(D=3, N=33, K=5, S=2, Ns=37, target="rosenbrock", noisy=False), (401, 402, 403, 404)),
    "k1_closed_form_D4": (dict(D=4, N=40, K=1, S=2, Ns=400, target="rosenbrock", noisy=False), (501, 502, 503, 504)),
}


def build(name):
    shape, seeds = CASES[name]
    return shape, workloads.build(dict(shape), orc.gplite_post, seeds=seeds)


def main():
    for name in CASES:
        shape, w = build(name)
        vp, gp, theta, eps, Ns = w["vp"], w["gp"], w["theta"], w["epsilon"], shape["Ns"]
        _, tb = orc.vpbounds(vp, gp, workloads.VP_OPTIONS)
        th = theta.copy()
        th[0] = tb["ub"][0] + 0.15           # one active soft bound
        out = orc.negelcbo_vbmc(th, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
        F, dF, G, H, _, dH = out[:6]
        full = orc.negelcbo_vbmc(th, 0.0, vp, gp, Ns, 0, 1, 0, tb, 0, epsilon=eps, nargout=11)
        # inputs: small ones verbatim; the draws (the bulk) as a fingerprint — they are regenerated from the NumPy Philox seed
        d = dict(theta=th, hyp=w["hyp"], X=w["X"], y=w["y"], alpha=np.stack([p["alpha"] for p in gp["post"]], axis=1),
                 eps_fingerprint=np.array([eps.size, eps.sum(), np.sum(eps * eps), *eps.ravel()[:8], *eps.ravel()[-8:]]),
                 F=F, dF=dF, G=G, H=H, dH=dH, I_sk=full[9])
        if eps.size <= 4096:
            d["epsilon"] = eps
        if shape["K"] == 1:
            D = shape["D"]
            sig, lam = float(np.asarray(vp["sigma"]).ravel()[0]), np.asarray(vp["lambda"]).ravel()
            # draws actually used: eps and -eps (antithetic), Ns even here
            ee = np.mean(np.sum(eps[0] ** 2, axis=1))
            # theta may move sigma/lambda: closed forms use the values unpacked from theta (negelcbo_vbmc.m:39-43)
            sig = float(np.exp(th[D * 1]))
            lam = np.exp(th[D * 1 + 1:D * 1 + 1 + D])
            d["closed_H"] = 0.5 * D * np.log(2 * np.pi) + D * np.log(sig) + np.sum(np.log(lam)) + 0.5 * ee
            d["closed_dH_mu"] = np.zeros(D)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "F", F, "H", H, "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    main()
