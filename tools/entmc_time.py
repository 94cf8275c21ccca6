"""Time the entropy sweep alone at a BASELINE configuration (CUDA events through the library's profile hooks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vbmc_b200
from vbmc_b200 import workloads, _lib
import ctypes as C

def main(cfg_name="c3", reps=20):
    reps = int(os.environ.get("VBMC_REPS", reps))
    ctx = vbmc_b200.default_context()
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
    ctx.set_precision(int(os.environ.get("VBMC_PREC", "64")))
    for i in range(3):
        _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
    tot = 0.0
    for i in range(reps):
        ctx.flush_l2(); a.stream = 10 + i
        _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
        tot += ms.value
    ctx.entmc_prune_stats(True)
    _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
    kept, total = ctx.entmc_prune_stats(False)
    ctx.profile_reset(); ctx.profile_enable(True)
    for i in range(reps):
        ctx.flush_l2(); a.stream = 100 + i
        _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
    ctx.profile_enable(False)
    out = {k: round(ctx.profile_get(k)[0] / reps, 4) for k in ("entmc", "entmc_f32", "philox", "gplogjoint", "reduce", "finalize", "vp_unpack")}
    if os.environ.get("VBMC_PREC", "64") != "64":
        print(cfg_name, "step_ms", round(tot / reps, 4), out, "kept", round(kept / max(1, total), 4), flush=True)
        return
    # the entropy sweep alone (no gplogjoint branch competing for the SMs), resident draws
    ctx.profile_reset(); ctx.profile_enable(True)
    for i in range(reps):
        ctx.flush_l2()
        vbmc_b200.entmc_vbmc(w["vp"], cfg["Ns"], epsilon="resident", ctx=ctx)
    ctx.profile_enable(False)
    out["entmc_alone"] = round(ctx.profile_get("entmc")[0] / reps, 4)
    ctx.profile_reset(); ctx.profile_enable(True)
    for i in range(reps):
        ctx.flush_l2()
        vbmc_b200.gplogjoint(w["vp"], w["gp"], ctx=ctx)
    ctx.profile_enable(False)
    out["glj_alone"] = round(ctx.profile_get("gplogjoint")[0] / reps, 4)
    out["glj_epi_alone"] = round(ctx.profile_get("gplogjoint_epilogue")[0] / reps, 4)
    print(cfg_name, "step_ms", round(tot / reps, 4), out, "F", F.value, "kept", round(kept / max(1, total), 4), flush=True)

if __name__ == "__main__":
    main(*(sys.argv[1:2]))
