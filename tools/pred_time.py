"""gplite_pred (Nstar = 1024) and the rank-one update on the c3 / c5 GPs: wall time through the host API."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vbmc_b200
from vbmc_b200 import workloads

ctx = vbmc_b200.default_context()
for name in sys.argv[1:] or ("c3",):
    cfg = dict(workloads.CONFIGS[name])
    X, y, s2 = workloads.make_training_set(cfg)
    hyp = workloads.make_hyp_samples(cfg, X, y)
    nf = [1, 1, 0] if s2 is not None else [1, 0, 0]
    gp = vbmc_b200.gplite_post(hyp, X, y, 1, 4, nf, s2, ctx=ctx, want_L=False)
    rs = np.random.default_rng(5)
    Xs = X[rs.integers(0, cfg["N"], 1024)] + 0.3 * rs.standard_normal((1024, cfg["D"]))
    vbmc_b200.gplite_pred(gp, Xs[:64], nargout=2, ctx=ctx)
    best = 1e9
    for i in range(3):
        ctx.sync(); t0 = time.perf_counter()
        out = vbmc_b200.gplite_pred(gp, Xs, nargout=4, ctx=ctx)
        ctx.sync(); best = min(best, time.perf_counter() - t0)
    fl = cfg["S"] * cfg["N"] ** 2 * 1024.0
    print(name, "gplite_pred Nstar=1024: %.2f ms, %.2f TFLOP/s (S N^2 Nstar)" % (best * 1e3, fl / best / 1e12), "fs2 mean %.6g" % float(np.mean(out[3])), flush=True)
    if s2 is None:
        ctx.sync(); t0 = time.perf_counter()
        gp1 = vbmc_b200.gplite_post_update1(gp, X[0] + 0.1, float(y[0]), ctx=ctx)
        ctx.sync(); dt = time.perf_counter() - t0
        ctx.sync(); t0 = time.perf_counter()
        gp2 = vbmc_b200.gplite_post_update1(gp1, X[1] + 0.1, float(y[1]), ctx=ctx)
        ctx.sync(); dt2 = time.perf_counter() - t0
        print(name, "rank-one update: %.3f ms (second: %.3f ms)" % (dt * 1e3, dt2 * 1e3), flush=True)
