"""Step time of a BASELINE configuration with the ahead-of-time draw generation on and off (VBMC_B200_PREFETCH),
same protocol as bench.py's `value` (per-step CUDA events, L2 flushed between steps), plus the device fminadam loop."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vbmc_b200
from vbmc_b200 import _lib, workloads


def run(cfg_name, prefetch, reps, glj_first="1"):
    os.environ["VBMC_B200_PREFETCH"] = prefetch
    os.environ["VBMC_B200_GLJ_FIRST"] = glj_first
    ctx = vbmc_b200.Context(0)
    cfg = dict(workloads.CONFIGS[cfg_name])
    w = workloads.build(cfg, lambda *a: vbmc_b200.gplite_post(*a, ctx=ctx, want_L=False), with_eps=False)
    _, tb = vbmc_b200.vpbounds(w["vp"], w["gp"], workloads.VP_OPTIONS)
    ctx.vp_set(w["vp"]); ctx.gp_attach(w["gp"]); ctx.thetabnd_set(tb)
    theta = np.ascontiguousarray(w["theta"])
    F, dF, ms = C.c_double(), np.zeros_like(theta), C.c_float()
    a = _lib.NegelcboArgs()
    a.theta, a.ntheta, a.beta, a.Ns = _lib.dptr(theta), theta.size, 0.0, cfg["Ns"]
    a.compute_grad, a.compute_var, a.separate_K, a.use_thetabnd = 1, 0, 0, 1
    a.eps_mode, a.seed, a.stream = _lib.EPS_PHILOX, 1, 0
    a.F, a.dF = C.pointer(F), _lib.dptr(dF)
    times = []
    for i in range(8 + reps):
        ctx.flush_l2(); a.stream = 100 + i
        _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
        if i >= 8:
            times.append(ms.value)
    Fv = F.value
    ctx.profile_reset(); ctx.profile_enable(True)
    for i in range(10):
        ctx.flush_l2(); a.stream = 100 + 8 + reps + i
        _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
    ctx.profile_enable(False)
    kern = {k: round(ctx.profile_get(k)[0] / 10 * 1e3, 1) for k in ("entmc", "philox", "gplogjoint", "reduce", "finalize", "vp_unpack")}
    # e2e through the host API, streaming keys
    x = theta.copy()
    for i in range(8):
        vbmc_b200.negelcbo_vbmc(x, 0.0, w["vp"], w["gp"], cfg["Ns"], 1, 0, 0, tb, 0, rng=(5, i), nargout=2, ctx=ctx)
    ctx.sync(); t0 = time.perf_counter()
    for i in range(reps):
        Fh, dFh = vbmc_b200.negelcbo_vbmc(x, 0.0, w["vp"], w["gp"], cfg["Ns"], 1, 0, 0, tb, 0, rng=(5, 8 + i), nargout=2, ctx=ctx)
        x = x - 1e-4 * dFh
    ctx.sync(); e2e = (time.perf_counter() - t0) / reps
    nit = 200
    vbmc_b200.fminadam_negelcbo(theta, 0.0, w["vp"], w["gp"], cfg["Ns"], 0, tb, None, None, 1e-9, 40, None, rng=(9, 0), ctx=ctx)
    ctx.sync(); t0 = time.perf_counter()
    _, _, _, ftab, it = vbmc_b200.fminadam_negelcbo(theta, 0.0, w["vp"], w["gp"], cfg["Ns"], 0, tb, None, None, 1e-9, nit, None, rng=(9, 1000), ctx=ctx)
    ctx.sync(); dt = time.perf_counter() - t0
    print(f"{cfg_name} prefetch={prefetch} glj_first={glj_first}: step {np.mean(times):.4f} ms (min {np.min(times):.4f}, p90 {np.percentile(times, 90):.4f}) "
          f"e2e {e2e * 1e3:.4f} ms  fminadam {dt / it * 1e3:.4f} ms/it ({it} it)  F={Fv!r} ftab[-1]={ftab[-1]!r} kernels_us={kern}", flush=True)
    del ctx


if __name__ == "__main__":
    cfgs = sys.argv[1:] or ["c3"]
    reps = int(os.environ.get("VBMC_REPS", "40"))
    for cn in cfgs:
        for pf, gf in (("1", "1"), ("1", "0")):
            run(cn, pf, reps, gf)
