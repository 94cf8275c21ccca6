"""Multi-GPU parity check, launched with torchrun, one rank per GPU (tests/test_mgpu.py runs it under pytest -m gpu on every
box that has at least two GPUs; it lives under tests/ because it uses the oracle as its checker):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_check.py
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
    say(f"comm_init done, peer-memory all-reduce: {ctx.comm_p2p()}")
    from oracle import vbmc_oracle as orc
    worst = 0.0
    quick = os.environ.get("VBMC_MGPU_QUICK", "0") == "1"   # fewer CPU-oracle evaluations (short multi-GPU slots)
    shapes = (dict(D=3, N=60, K=5, S=3, Ns=100), dict(D=6, N=200, K=20, S=8, Ns=4096), dict(D=10, N=300, K=50, S=5, Ns=2048))
    for shape in (shapes[:2] if quick else shapes):
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
            # checker: the binary128 evaluation (tests/test_truth128.py); the FP64 NumPy oracle's own distance is printed beside it
            from oracle import cport
            Ft, dFt, Gt, Ht, _, _ = cport.negelcbo(cport.Prepared(vp, gp, tb), theta, cfg["Ns"], eps, truth128=True)
            Fo, dFo, Go, Ho = orc.negelcbo_vbmc(theta, 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0, epsilon=eps, nargout=4)[:4]
            rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))
            errs = (rel(F, Ft), rel(dF, dFt), rel(G, Gt), rel(H, Ht))
            worst = max(worst, *errs)
            print(f"[mgpu_check] world={world} {shape}: rel errors vs binary128 F,dF,G,H = {errs}; FP64 oracle vs binary128 = "
                  f"{(rel(Fo, Ft), rel(dFo, dFt), rel(Go, Gt), rel(Ho, Ht))}", flush=True)
        dist.barrier()
        # beta != 0 with the diagonal variance and its gradient (negelcbo_vbmc.m:119-130): the per-sample gradients of ALL hyper-
        # parameter samples are needed on every rank although the step shards them (stale rows of glj_out)
        Fb, dFb = vbmc_b200.negelcbo_vbmc(theta, 0.7, vp, gp, cfg["Ns"], 1, 2, 0, tb, 0, epsilon=eps, nargout=2, ctx=ctx)
        t = torch.tensor(np.concatenate([[Fb], dFb]), device=f"cuda:{local}")
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref), f"rank {rank}: beta != 0 results differ from rank 0"
        if rank == 0:
            Fbo, dFbo = orc.negelcbo_vbmc(theta, 0.7, vp, gp, cfg["Ns"], 1, 2, 0, tb, 0, epsilon=eps, nargout=2)[:2]
            eb = (rel(Fb, Fbo), rel(dFb, dFbo))
            print(f"[mgpu_check] world={world} {shape}: beta=0.7, compute_var=2: rel errors F,dF vs oracle = {eb}", flush=True)
            assert max(eb) < 1e-6, eb   # the variance goes through K^-1 (cond ~ sf2/sn2): 1e-7..1e-6 is the FP64 floor (test_gpu_parity.py)
        dist.barrier()
    # streaming generator-mode calls (stream advancing by one): ahead-of-time draws + graph replay of the multi-GPU step
    # (the all-reduce is inside finalize_kernel).  Every call must equal parity mode on that stream's dumped draws, on every rank.
    refs = []
    for i in range(8):
        e = ctx.eps_philox(cfg["D"], cfg["K"], cfg["Ns"], 42, 100 + i, readback=True)   # this rank's shard of the draws
        refs.append(vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0, epsilon=e, nargout=2, ctx=ctx))
    for i in range(8):
        Fi, dFi = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0, rng=(42, 100 + i), nargout=2, ctx=ctx)
        assert Fi == refs[i][0] and np.array_equal(dFi, refs[i][1]), f"rank {rank}: streaming call {i} differs from parity mode"
        t = torch.tensor(np.concatenate([[Fi], dFi]), device=f"cuda:{local}")
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref), f"rank {rank}: streaming call {i} differs from rank 0"
    say("streaming calls OK")
    # device-resident fminadam loop: every rank runs the same loop on its shard; iterates must agree bit for bit
    x, f, xtab, ftab, it = vbmc_b200.fminadam_negelcbo(theta, 0.0, vp, gp, cfg["Ns"], 0, tb, None, None, 0.001, 45, None, epsilon=eps, ctx=ctx)
    t = torch.tensor(np.concatenate([[f, it], x, ftab]), device=f"cuda:{local}")
    ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(t, ref), f"rank {rank}: fminadam results differ from rank 0"
    if rank == 0 and not quick:
        fun = lambda t_: orc.negelcbo_vbmc(t_, 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0, epsilon=eps, nargout=2)[:2]
        xo, fo, xtabo, ftabo, ito = orc.fminadam(fun, theta, None, None, 0.001, 45, None)
        e = float(np.max(np.abs(ftab - ftabo)) / np.max(np.abs(ftabo)))
        print(f"[mgpu_check] fminadam world={world}: iter {it} (oracle {ito}), rel err ftab {e:.2e}", flush=True)
        assert it == ito and e < 1e-8
    dist.barrier()   # rank 0 spent seconds in the CPU oracle: bring the ranks back together before the next GPU call
    # generator mode inside the device loop (one stream per iteration, draws produced ahead of time), graph replay
    x2, f2, xtab2, ftab2, it2 = vbmc_b200.fminadam_negelcbo(theta, 0.0, vp, gp, cfg["Ns"], 0, tb, None, None, 1e-9, 60, None, rng=(7, 500), ctx=ctx)
    t = torch.tensor(np.concatenate([[f2, it2], x2, ftab2]), device=f"cuda:{local}")
    ref = t.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(t, ref), f"rank {rank}: generator-mode fminadam results differ from rank 0"
    Fi, dFi = vbmc_b200.negelcbo_vbmc(xtab2[-2], 0.0, vp, gp, cfg["Ns"], 1, 0, 0, tb, 0,
                                     rng=(7, 500 + it2 - 1), nargout=2, ctx=ctx)
    assert abs(Fi - ftab2[-1]) <= 1e-12 * abs(Fi), (Fi, ftab2[-1])   # iteration it evaluates x_{it-1} with stream 500 + it - 1
    say("generator-mode fminadam OK")
    if os.environ.get("VBMC_MGPU_TIME", "0") == "1":
        import ctypes as C
        import time
        cfg3 = dict(workloads.CONFIGS[os.environ.get("VBMC_MGPU_CFG", "c3")])
        w3 = workloads.build(cfg3, lambda *a: vbmc_b200.gplite_post(*a, ctx=ctx, want_L=False), with_eps=False)
        _, tb3 = vbmc_b200.vpbounds(w3["vp"], w3["gp"], workloads.VP_OPTIONS)
        ctx.vp_set(w3["vp"]); ctx.gp_attach(w3["gp"]); ctx.thetabnd_set(tb3)
        th3 = np.ascontiguousarray(w3["theta"])
        F, dF, ms = C.c_double(), np.zeros_like(th3), C.c_float()
        a = _lib.NegelcboArgs()
        a.theta, a.ntheta, a.beta, a.Ns = _lib.dptr(th3), th3.size, 0.0, cfg3["Ns"]
        a.compute_grad, a.compute_var, a.separate_K, a.use_thetabnd = 1, 0, 0, 1
        a.eps_mode, a.seed, a.stream = _lib.EPS_PHILOX, 1, 0
        a.F, a.dF = C.pointer(F), _lib.dptr(dF)
        tot = 0.0
        for i in range(48):
            ctx.flush_l2(); a.stream = 100 + i
            if i == 8:
                ctx.sync(); dist.barrier(); torch.cuda.synchronize()
            _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
            if i >= 8:
                tot += ms.value
        t = torch.tensor([tot / 40], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        x = th3.copy()
        for i in range(8):
            vbmc_b200.negelcbo_vbmc(x, 0.0, w3["vp"], w3["gp"], cfg3["Ns"], 1, 0, 0, tb3, 0, rng=(5, i), nargout=2, ctx=ctx)
        ctx.sync(); dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(100):
            Fh, dFh = vbmc_b200.negelcbo_vbmc(x, 0.0, w3["vp"], w3["gp"], cfg3["Ns"], 1, 0, 0, tb3, 0, rng=(5, 8 + i), nargout=2, ctx=ctx)
            x = x - 1e-4 * dFh
        ctx.sync(); dist.barrier(); torch.cuda.synchronize(); e2e = (time.perf_counter() - t0) / 100
        vbmc_b200.fminadam_negelcbo(th3, 0.0, w3["vp"], w3["gp"], cfg3["Ns"], 0, tb3, None, None, 1e-9, 40, None, rng=(9, 0), ctx=ctx)
        ctx.sync(); dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        _, _, _, _, it3 = vbmc_b200.fminadam_negelcbo(th3, 0.0, w3["vp"], w3["gp"], cfg3["Ns"], 0, tb3, None, None, 1e-9, 200, None, rng=(9, 1000), ctx=ctx)
        ctx.sync(); dist.barrier(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / it3
        if rank == 0:
            print(f"[mgpu_time] world={world} p2p={ctx.comm_p2p()} {os.environ.get('VBMC_MGPU_CFG', 'c3')}: step {float(t.item()):.4f} ms (max over ranks), "
                  f"e2e {e2e * 1e3:.4f} ms, fminadam {dt * 1e3:.4f} ms/it, F={F.value!r}", flush=True)
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
