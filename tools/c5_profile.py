"""Per-kernel times of one negelcbo step at a named configuration, FP64 and FP32 sweeps (run on the GPU box)."""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import vbmc_b200
from vbmc_b200 import _lib, workloads


def main(names):
    ctx = vbmc_b200.default_context()
    out = {}
    for name in names:
        cfg = dict(workloads.CONFIGS[name])
        w = workloads.build(cfg, lambda *a: vbmc_b200.gplite_post(*a, ctx=ctx, want_L=False), with_eps=False)
        _, tb = vbmc_b200.vpbounds(w["vp"], w["gp"], workloads.VP_OPTIONS)
        ctx.vp_set(w["vp"]); ctx.gp_attach(w["gp"]); ctx.thetabnd_set(tb)
        theta = np.ascontiguousarray(w["theta"])
        F, dF, ms = C.c_double(), np.zeros_like(theta), C.c_float()
        a = _lib.NegelcboArgs()
        a.theta, a.ntheta, a.beta, a.Ns = _lib.dptr(theta), theta.size, 0.0, cfg["Ns"]
        a.compute_grad, a.compute_var, a.separate_K, a.use_thetabnd = 1, 0, 0, 1
        a.eps_mode, a.seed, a.stream = _lib.EPS_PHILOX, 7, 0
        a.F, a.dF = C.pointer(F), _lib.dptr(dF)
        for bits in (64, 32):
            ctx.set_precision(bits)
            for i in range(2):
                a.stream = i
                _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
            ctx.profile_reset(); ctx.profile_enable(True)
            n = 4
            tot = 0.0
            for i in range(n):
                a.stream = 10 + i
                _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
                tot += ms.value
            ctx.sync()
            prof = {}
            for k in ("vp_unpack", "philox", "entmc", "entmc_direct", "entmc_f32_tables", "entmc_f32", "gplogjoint", "reduce", "finalize"):
                msk, cnt = ctx.profile_get(k)
                if cnt:
                    prof[k] = round(msk / n, 4)
            ctx.profile_enable(False)
            out[f"{name}_{bits}"] = {"ms_per_step": tot / n, "kernels_ms": prof, "F": F.value}
        ctx.set_precision(64)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:] or ["c3", "c5"])
