"""Multi-GPU parity check, launched with torchrun (one rank per GPU):
every rank evaluates the same negelcbo_vbmc step on its shard (MC pair axis, hyper-parameter samples);
after the single NCCL all-reduce all ranks must hold the same F, dF, equal to the oracle's (rank 0 checks)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import vbmc_b200
from vbmc_b200 import _lib, workloads


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    def say(m):
        print(f"[r{rank}] {m}", flush=True)
    say("start")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    say("torch pg up")
    ctx = vbmc_b200.Context(local)
    say("ctx up")
    uid = torch.zeros(_lib.UNIQUE_ID_BYTES, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(vbmc_b200.Context.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    say("uid broadcast")
    ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    say("comm_init done")
    from oracle import vbmc_oracle as orc
    worst = 0.0
    for shape in (dict(D=3, N=60, K=5, S=3, Ns=100), dict(D=6, N=200, K=20, S=8, Ns=4096), dict(D=10, N=300, K=50, S=5, Ns=2048)):
        cfg = dict(shape, target="rosenbrock", noisy=False)
        w = workloads.build(cfg, orc.gplite_post)
        vp, gp, theta, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
        _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
        say(f"calling negelcbo {shape}")
        F, dF, G, H = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0, epsilon=eps, nargout=4, ctx=ctx)
        say("negelcbo returned")
        # identical on every rank (the all-reduce result is replicated)
        t = torch.tensor(np.concatenate([[F, G, H], dF]), device=f"cuda:{local}")
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref), f"rank {rank}: results differ from rank 0"
        # device Philox draws do not depend on the sharding: compare against a single-GPU style evaluation of the dump
        F2, dF2 = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0, rng=(42, 3), nargout=2, ctx=ctx)
        if rank == 0:
            Fo, dFo, Go, Ho = orc.negelcbo_vbmc(theta, 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0, epsilon=eps, nargout=4)[:4]
            rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))
            errs = (rel(F, Fo), rel(dF, dFo), rel(G, Go), rel(H, Ho))
            worst = max(worst, *errs)
            print(f"[mgpu_check] world={world} {shape}: rel errors F,dF,G,H = {errs}", flush=True)
    # device-resident fminadam loop: every rank runs the same loop on its shard; iterates must agree bit for bit
    x, f, xtab, ftab, it = vbmc_b200.fminadam_negelcbo(theta, 0.0, vp, gp, cfg["Ns"], 0, tb, None, None, 0.001, 45, None, epsilon=eps, ctx=ctx)
    t = torch.tensor(np.concatenate([[f, it], x, ftab]), device=f"cuda:{local}")
    ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(t, ref), f"rank {rank}: fminadam results differ from rank 0"
    if rank == 0:
        fun = lambda t_: orc.negelcbo_vbmc(t_, 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0, epsilon=eps, nargout=2)[:2]
        xo, fo, xtabo, ftabo, ito = orc.fminadam(fun, theta, None, None, 0.001, 45, None)
        e = float(np.max(np.abs(ftab - ftabo)) / np.max(np.abs(ftabo)))
        print(f"[mgpu_check] fminadam world={world}: iter {it} (oracle {ito}), rel err ftab {e:.2e}", flush=True)
        assert it == ito and e < 1e-8
    if rank == 0:
        assert worst < 1e-10, worst
        print(f"[mgpu_check] OK world={world} worst={worst:.3e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        os._exit(1)   # never leave the other ranks waiting in a collective
