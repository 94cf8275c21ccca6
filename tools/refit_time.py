"""Wall time of vbmc_b200.gplite_post (full refit: Gram, Cholesky, alpha) at the benchmark sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vbmc_b200
from vbmc_b200 import workloads

ctx = vbmc_b200.default_context()
for name in sys.argv[1:] or ("c3", "c5"):
    cfg = dict(workloads.CONFIGS[name])
    X, y, s2 = workloads.make_training_set(cfg)
    hyp = workloads.make_hyp_samples(cfg, X, y)
    nf = [1, 1, 0] if s2 is not None else [1, 0, 0]
    best = 1e9
    for i in range(5):
        ctx.sync(); t0 = time.perf_counter()
        vbmc_b200.gplite_post(hyp, X, y, 1, 4, nf, s2, ctx=ctx, want_L=False)
        ctx.sync(); best = min(best, time.perf_counter() - t0)
    print(name, "gplite_post wall ms %.3f" % (best * 1e3), "N^3/3*S TFLOP/s %.2f" % (cfg["S"] * cfg["N"] ** 3 / 3 / best / 1e12), flush=True)
    ctx.profile_reset(); ctx.profile_enable(True)   # per-kernel events, serial order (no look-ahead, no graph)
    vbmc_b200.gplite_post(hyp, X, y, 1, 4, nf, s2, ctx=ctx, want_L=False)
    ctx.profile_enable(False)
    print("   serial ms:", {k: round(ctx.profile_get(k)[0], 3) for k in ("gram", "potrf_potf2", "potrf_trsm", "potrf_update", "trsv", "gp_prep")}, flush=True)
