"""Time gplite_pred / rank-one update / batched nlZ on the c3 GP (wall clock around synced host-API calls)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vbmc_b200
from vbmc_b200 import workloads

ctx = vbmc_b200.default_context()
cfg = dict(workloads.CONFIGS["c3"])
X, y, s2 = workloads.make_training_set(cfg)
hyp = workloads.make_hyp_samples(cfg, X, y)
gp = vbmc_b200.gplite_post(hyp, X, y, 1, 4, [1, 0, 0], None, ctx=ctx, want_L=False)
rs = np.random.default_rng(5)
Xs = X[rs.integers(0, cfg["N"], 1024)] + 0.3 * rs.standard_normal((1024, cfg["D"]))
vbmc_b200.gplite_pred(gp, Xs[:64], nargout=2, ctx=ctx)
for n in (1, 64, 1024):
    ctx.sync(); t0 = time.perf_counter()
    vbmc_b200.gplite_pred(gp, Xs[:n], nargout=2, ctx=ctx)
    ctx.sync(); print("gplite_pred Nstar=%d: %.3f ms" % (n, 1e3 * (time.perf_counter() - t0)), flush=True)
g2 = gp
for i in range(3):
    ctx.sync(); t0 = time.perf_counter()
    g2 = vbmc_b200.gplite_post_update1(g2, X[i] + 0.1, float(y[i]), ctx=ctx)
    ctx.sync(); print("rank-one update %d: %.3f ms" % (i, 1e3 * (time.perf_counter() - t0)), flush=True)
ctx.profile_reset(); ctx.profile_enable(True)
g2 = vbmc_b200.gplite_post_update1(g2, X[5] + 0.1, float(y[5]), ctx=ctx)
ctx.profile_enable(False)
print({k: round(ctx.profile_get(k)[0], 4) for k in ("pred_cross", "pred_trsm", "pred_var", "rank1")})
