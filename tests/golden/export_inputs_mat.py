"""Write the inputs of the committed golden cases as MATLAB files (tests/golden/matlab_inputs/<case>.mat) for
bench/matlab/dump_reference_vectors.m, which runs the UNMODIFIED reference on them on a box that has MATLAB and writes
tests/golden/reference_<case>.mat (see bench/matlab/README.md).  Same seeded inputs as make_golden.py.

    python tests/golden/export_inputs_mat.py            # the small committed cases
    python tests/golden/export_inputs_mat.py --full     # also c3_full (D=10, N=2000, K=50, Ns=32768, S=20; ~70 MB, not committed)
"""
import os
import sys

import numpy as np
from scipy.io import savemat

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import make_golden  # noqa: E402
from oracle import vbmc_oracle as orc  # noqa: E402
from vbmc_b200 import workloads  # noqa: E402

SMALL = ["c1_rosenbrock_D2", "ragged_D3", "k1_closed_form_D4", "c2_reduced_D6"]


def export(name, shape, w, theta):
    vp, eps = w["vp"], w["epsilon"]
    D, K = shape["D"], shape["K"]
    rs = np.random.Generator(np.random.Philox(900))
    Xstar = w["X"][rs.integers(0, w["X"].shape[0], 16)] + 0.2 * rs.standard_normal((16, D))
    d = dict(D=float(D), K=float(K), Ns=float(int(np.ceil(shape["Ns"] / 2) * 2)), meanfun=4.0,
             X=w["X"], y=w["y"].reshape(-1, 1), hyp=w["hyp"], Xstar=Xstar,
             mu=np.asarray(vp["mu"]).reshape(D, K), sigma=np.ravel(vp["sigma"]), **{"lambda": np.ravel(vp["lambda"])},
             w=np.ravel(vp["w"]), eta=np.ravel(vp["eta"]), theta=theta.reshape(-1, 1),
             epsilon=np.ascontiguousarray(np.transpose(eps, (2, 1, 0))),        # (K, Ns/2, D) -> D x Ns/2 x K
             **{k: float(v) for k, v in workloads.VP_OPTIONS.items()})
    if w.get("s2") is not None:
        d["s2"] = w["s2"].reshape(-1, 1)
    out = os.path.join(HERE, "matlab_inputs")
    os.makedirs(out, exist_ok=True)
    savemat(os.path.join(out, name + ".mat"), d, do_compression=True)
    print(name, os.path.getsize(os.path.join(out, name + ".mat")), "bytes")


def main():
    for name in SMALL:
        shape, w = make_golden.build(name)
        g = np.load(os.path.join(HERE, name + ".npz"))
        export(name, shape, w, g["theta"])
    if "--full" in sys.argv:
        cfg = dict(workloads.CONFIGS["c3"])
        w = workloads.build(cfg, orc.gplite_post)
        export("c3_full", cfg, w, w["theta"])


if __name__ == "__main__":
    main()
