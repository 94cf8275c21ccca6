"""Per-iteration time of the device fminadam loop at c3 under different switches (diagnostic)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vbmc_b200
from vbmc_b200 import workloads


def run(tag, env, prof=False, pre_profile=False):
    for k in ("VBMC_B200_PREFETCH", "VBMC_B200_GLJ_FIRST", "VBMC_B200_GRAPHS", "VBMC_B200_ENTMC_BALANCE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = vbmc_b200.Context(0)
    cfg = dict(workloads.CONFIGS["c3"])
    w = workloads.build(cfg, lambda *a: vbmc_b200.gplite_post(*a, ctx=ctx, want_L=False), with_eps=False)
    _, tb = vbmc_b200.vpbounds(w["vp"], w["gp"], workloads.VP_OPTIONS)
    th = np.ascontiguousarray(w["theta"])
    if pre_profile:
        ctx.profile_reset(); ctx.profile_enable(True)
        vbmc_b200.negelcbo_vbmc(th, 0.0, w["vp"], w["gp"], cfg["Ns"], 1, 0, 0, tb, 0, rng=(5, 1), nargout=2, ctx=ctx)
        ctx.profile_enable(False)
    vbmc_b200.fminadam_negelcbo(th, 0.0, w["vp"], w["gp"], cfg["Ns"], 0, tb, None, None, 1e-9, 40, None, rng=(9, 0), ctx=ctx)
    res = []
    for rep in range(2):
        if prof:
            ctx.profile_reset(); ctx.profile_enable(True)
        ctx.sync(); t0 = time.perf_counter()
        _, _, _, ftab, it = vbmc_b200.fminadam_negelcbo(th, 0.0, w["vp"], w["gp"], cfg["Ns"], 0, tb, None, None, 1e-9, 100, None, rng=(9, 1000 + 500 * rep), ctx=ctx)
        ctx.sync(); res.append((time.perf_counter() - t0) / it * 1e3)
        if prof:
            ctx.profile_enable(False)
    kern = {k: round(ctx.profile_get(k)[0] / 100 * 1e3, 1) for k in ("entmc", "philox", "gplogjoint", "reduce", "finalize", "vp_unpack", "adam_step")} if prof else None
    print(f"{tag}: {res[0]:.4f} {res[1]:.4f} ms/it ftab[-1]={ftab[-1]!r} {kern}", flush=True)
    del ctx


if __name__ == "__main__":
    run("default", {})
    run("balance=0", {"VBMC_B200_ENTMC_BALANCE": "0"})
    run("balance=0 profiled (direct launches)", {"VBMC_B200_ENTMC_BALANCE": "0"}, prof=True)
    run("default+pre_profile", {}, pre_profile=True)
    run("glj_first=1", {"VBMC_B200_GLJ_FIRST": "1"})
    run("prefetch=0", {"VBMC_B200_PREFETCH": "0"})
    run("graphs=0", {"VBMC_B200_GRAPHS": "0"})
    run("profiled (direct launches)", {}, prof=True)
