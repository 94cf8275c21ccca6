"""Per-kernel CUDA-event times of one configuration on N ranks (torchrun), max over ranks of the step time.
    python -m torch.distributed.run --nproc-per-node N tools/mgpu_profile.py c4"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import vbmc_b200
from vbmc_b200 import _lib, workloads


def main(cfg_name="c4", reps=20):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = vbmc_b200.Context(local)
    uid = torch.zeros(_lib.UNIQUE_ID_BYTES, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(vbmc_b200.Context.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
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

    def sync():
        ctx.sync(); dist.barrier(); torch.cuda.synchronize()

    for i in range(5):
        a.stream = i
        _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
    out = {}
    for flush in (True, False):
        tot = 0.0
        sync()
        for i in range(reps):
            if flush:
                ctx.flush_l2()
            a.stream = 100 + i + (0 if flush else 1000)
            _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
            tot += ms.value
        t = torch.tensor([tot / reps], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["step_ms_flush" if flush else "step_ms_noflush"] = round(float(t.item()), 4)
    sync()
    a.stream = 5000
    _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 100, C.byref(ms)))
    t = torch.tensor([ms.value / 100], device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["loop100_ms_per_step"] = round(float(t.item()), 4)
    ctx.profile_reset(); ctx.profile_enable(True)
    sync()
    for i in range(reps):
        ctx.flush_l2(); a.stream = 9000 + i
        _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
    ctx.profile_enable(False)
    prof = {k: round(ctx.profile_get(k)[0] / reps, 4) for k in ("entmc", "philox", "gplogjoint", "reduce", "finalize", "vp_unpack")}
    sync()
    print(f"[r{rank}] {cfg_name} world={world} {out} profile-mode kernels {prof}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(*(sys.argv[1:2]))
