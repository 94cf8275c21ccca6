#!/usr/bin/env python
"""bench.py — negelcbo_vbmc grad-steps/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 50 --warmup 5            # this repo's CUDA path
    python bench.py --impl reference --steps 3 --warmup 1     # reference CPU path (oracle port, host cores)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" = one [F,dF] = negelcbo_vbmc(theta, 0, vp, gp, Ns, 1, 0, 0, thetabnd) evaluation with fresh
entropy draws (the reference draws randn inside entmc_vbmc every call), i.e. one fminadam iteration's
objective call (utils/fminadam.m:48).  Workload: synthetic config c3 (D=10, N=2000, K=50, Ns=32768 per
component, S=20) at N=1, c4 (Ns=131072, strong-sharded MC axis) is available with --config c4.

Timed quantities
  value : steps/s with theta, GP posterior and draws resident in HBM; every step is timed with CUDA
          events on the library's launch stream (L2 flushed between steps, outside the events);
          max over ranks.
  e2e   : steps/s through the public host API vbmc_b200.negelcbo_vbmc (host theta in, F/dF out,
          Adam update of fminadam.m:51-60 on the host), wall clock between device syncs.
  roofline : entmc kernel (dominant): algorithmic FLOPs / CUDA-event duration vs the FP64 FMA peak
          measured live on the same device; HBM fraction of the algorithmic bytes reported beside it.
  cpu_baseline : the oracle's C/OpenMP port on the host cores, bounded sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", default=None, help="c2|c3|c4 (default: c3 at 1 GPU, c4 at >1)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-parity", action="store_true", help="skip the parity gate (profiling runs under ncu only; such a line says so)")
    return p.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def scipy_gp_post(hyp, X, y, covfun, meanfun, noisefun, s2):
    """Workload set-up only (NOT the timed path, NOT the oracle): GP posterior via LAPACK."""
    import scipy.linalg as sla
    N, D = X.shape
    S = hyp.shape[1]
    post = []
    for s in range(S):
        h = hyp[:, s]
        ell, sf2, sn2 = np.exp(h[:D]), math.exp(2 * h[D]), math.exp(2 * h[D + 1])
        Z = X / ell
        sq = np.maximum(np.sum(Z * Z, 1)[:, None] + np.sum(Z * Z, 1)[None, :] - 2 * Z @ Z.T, 0)
        Kmat = sf2 * np.exp(-0.5 * sq)
        sn2v = sn2 + (s2 if s2 is not None else 0.0) * np.ones(N)
        m = h[D + 2] - 0.5 * np.sum(((X - h[D + 3:2 * D + 3]) / np.exp(h[2 * D + 3:])) ** 2, axis=1)
        sdiv = float(np.min(sn2v))
        L = sla.cholesky(Kmat / sdiv + np.diag(sn2v / sdiv), lower=False)
        alpha = sla.cho_solve((L, False), y - m) / sdiv
        post.append({"hyp": h.copy(), "alpha": alpha, "sW": np.ones(N) / math.sqrt(sdiv), "L": None, "sn2_mult": 1.0, "Lchol": True})
    return {"X": X, "y": y, "s2": s2, "covfun": 1, "meanfun": 4, "noisefun": noisefun, "Ncov": D + 1, "Nnoise": 1,
            "Nmean": 2 * D + 1, "post": post}


def library_cholesky_bar(w, cfg, local):
    """The on-box library bar for the refit's factorisation (hard part 3): the same S matrices K/sl + diag(sn2/sl)
    factored by the vendor library through torch.linalg.cholesky (cuSOLVER potrfBatched / potrf as torch dispatches; FP64).
    Bench-only: the product never links cuSOLVER.  Factorisation only -- no Gram, no solves -- so it is a LOWER bound of a
    library-built gplite_post."""
    try:
        import torch
        dev = torch.device("cuda", local)
        X = torch.tensor(w["X"], device=dev, dtype=torch.float64)
        hyp = torch.tensor(w["hyp"], device=dev, dtype=torch.float64)
        D, N, S = cfg["D"], cfg["N"], cfg["S"]
        mats = []
        for s_ in range(S):
            h = hyp[:, s_]
            Z = X / torch.exp(h[:D])
            sq = torch.cdist(Z, Z).pow(2)
            sn2 = torch.exp(2 * h[D + 1]) + (1.0 if w["s2"] is not None else 0.0)
            mats.append(torch.exp(2 * h[D]) * torch.exp(-0.5 * sq) / sn2 + torch.eye(N, device=dev, dtype=torch.float64))
        A = torch.stack(mats)
        del mats
        torch.linalg.cholesky(A, upper=True)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for _ in range(3):
            e0.record()
            torch.linalg.cholesky(A, upper=True)
            e1.record()
            torch.cuda.synchronize(dev)
            t = e0.elapsed_time(e1)
            best = t if best is None else min(best, t)
        flops = S * N ** 3 / 3.0
        out = {"torch_linalg_cholesky_ms": best, "tflops": flops / (best * 1e-3) / 1e12, "S": S, "N": N,
               "what": "torch.linalg.cholesky(upper=True) on the S x N x N batch, FP64, best of 3, CUDA events (factorisation only)"}
        del A
        torch.cuda.empty_cache()
        return out
    except Exception as e:   # reporting only
        return {"error": str(e)[:200]}


def algorithmic_counts(cfg):
    """per grad-step."""
    D, K, Ns, N, S = cfg["D"], cfg["K"], cfg["Ns"], cfg["N"], cfg["S"]
    return {"entmc_triples": K * K * Ns, "entmc_flops": K * K * Ns * (5 * D + 12), "entmc_exp": K * K * Ns,
            "entmc_bytes": K * (Ns // 2) * D * 8 + (1 + 2 * D + K) * K * 8,
            "glj_flops": S * K * N * (7 * D + 10), "glj_bytes": (N * D + S * N) * 8}


def adam_update(state, x, grad):
    """utils/fminadam.m:51-60."""
    state["it"] += 1
    it = state["it"]
    state["m"] = 0.9 * state["m"] + 0.1 * grad
    state["v"] = 0.999 * state["v"] + 0.001 * grad * grad
    mhat = state["m"] / (1 - 0.9 ** it)
    vhat = state["v"] / (1 - 0.999 ** it)
    step = 0.001 + (0.1 - 0.001) * math.exp(-it / 200.0)
    return x - step * mhat / (np.sqrt(vhat) + math.sqrt(np.finfo(float).eps))


def numpy_stand_in(w, cfg):
    """The NumPy restatement that materialises the same D x Ns x K temporaries as the .m code (oracle/vbmc_oracle.py) — the
    stand-in for "the reference's MATLAB CPU path" (MATLAB cannot run here) — on a bounded sample: one
    evaluation with Ns = 2048 draws per component, the entropy part (linear in Ns) scaled to the configuration's Ns."""
    from oracle import vbmc_oracle as orc
    from vbmc_b200 import workloads
    vp, gp, theta = w["vp"], w["gp"], w["theta"]
    _, tb = orc.vpbounds(dict(vp), gp, workloads.VP_OPTIONS)
    Ns_s = min(cfg["Ns"], 2048)
    eps = workloads.make_epsilon(cfg, Ns=Ns_s)
    t0 = time.perf_counter()
    orc.negelcbo_vbmc(theta, 0.0, vp, gp, Ns_s, 1, 0, 0, tb, 0, epsilon=eps, nargout=2)
    t_small = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.gplogjoint(vp, gp, [1, 1, 1, 1], True, True, 0, nargout=2)
    t_glj = time.perf_counter() - t0
    t_step = t_glj + max(t_small - t_glj, 0.0) * cfg["Ns"] / Ns_s
    return {"value": 1.0 / t_step, "unit": "steps/s", "cores": 1, "kind": "port",
            "sample": f"NumPy restatement vectorised like the .m code (elementwise NumPy is single-threaded): one evaluation at Ns={Ns_s} "
                      f"({t_small:.2f} s, of which gplogjoint {t_glj:.2f} s), entropy part scaled x{cfg['Ns'] // Ns_s} to Ns={cfg['Ns']}"}



TRAFFIC_C3_BYTES = 65.671e6 + 0.294e6
TRAFFIC_C3_SOURCE = "ncu --set full capture r2b_entmc2 of the final round-2 build (profiles/r2_ncu_summary.md): dram read 65.67 MB + write 0.29 MB per launch; copied, not measured in this run"
PARITY_TOL = 1e-10      # FP64 gate (BASELINE.json north_star); FP32 sweep: 1e-4


def _rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def parity_gate(ctx, vbmc_b200, w, tb, cfg, dist, local, rank, world, *, full=True, truth_Ns=128, truth_S=None, precision=64,
                reduced_Ns=None):
    """Checked BEFORE anything is timed: one GPU step of this configuration against the CPU restatements, same draws.

    full      : the configuration's own Ns, device Philox draws dumped to the host, vs the C port in FP64 (entropy side: no
                cancellation, the port is good to 1e-13; log-joint side: the port itself is ~1e-10 from the truth, reported).
    truth     : theta, GP and the first `truth_Ns` draws per component vs the binary128 evaluation (oracle/c -DVBMC_ORACLE_QUAD):
                the whole of F, dF, G, H, dH at 1e-10 (optionally on the first `truth_S` hyper-parameter samples only).
    N > 1     : F, dF must be bit-identical on all ranks and within 1e-12 of a 1-rank evaluation of the same Philox draws.
    Raises on failure; returns the dict that goes into the bench line as "parity"."""
    from oracle import cport
    vp, gp, theta = w["vp"], w["gp"], w["theta"]
    D, K, Ns = cfg["D"], cfg["K"], (reduced_Ns or cfg["Ns"])
    tol = PARITY_TOL if precision == 64 else 1e-4
    seed, stream = 990001, 7
    out = {"tol": tol, "sweep_bits": precision}
    c1 = ctx if world == 1 else vbmc_b200.Context(local)      # 1-rank evaluator of the same draws
    ctx.set_precision(precision)
    try:
        if full:
            got = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, rng=(seed, stream), nargout=6, ctx=ctx)
            F, dF, G, H, _, dH = got
            if world > 1:
                import torch
                t = torch.tensor(np.concatenate([[F, G, H], dF, dH]), device=f"cuda:{local}")
                allt = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(allt, t)
                same = all(torch.equal(allt[0], x) for x in allt)
                out["ranks_bit_identical"] = bool(same)
                if not same:
                    raise RuntimeError("parity gate: F/dF differ between ranks")
                c1.set_precision(precision)
                F1, dF1, G1, H1, _, dH1 = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, rng=(seed, stream), nargout=6, ctx=c1)
                out["vs_one_rank_same_draws"] = dict(F=_rel(F, F1), dF=_rel(dF, dF1), G=_rel(G, G1), H=_rel(H, H1), dH=_rel(dH, dH1))
                lim = 1e-12 if precision == 64 else 1e-5
                if max(out["vs_one_rank_same_draws"].values()) > lim:
                    raise RuntimeError(f"parity gate: {world}-rank step differs from the 1-rank step on the same draws: {out['vs_one_rank_same_draws']}")
            if rank == 0:
                c1.set_precision(precision)
                eps = c1.eps_philox(D, K, Ns, seed, stream, readback=True)
                c1.set_precision(64)
                prep = cport.Prepared(vp, gp, tb)
                Fp, dFp, Gp, Hp, dHp, _ = cport.negelcbo(prep, theta, Ns, eps)
                e = dict(F=_rel(F, Fp), dF=_rel(dF, dFp), G=_rel(G, Gp), H=_rel(H, Hp), dH=_rel(dH, dHp))
                out["vs_cport_fp64_full"] = dict(e, Ns=Ns, what="C/OpenMP port in FP64 on the dumped device draws; its own log-joint "
                                                 "side is ~1e-10 from the binary128 truth at this conditioning, so only H, dH are gated here")
                if max(e["H"], e["dH"]) > tol:
                    raise RuntimeError(f"parity gate ({cfg.get('name', '')} Ns={Ns}, {precision}-bit sweep): entropy side off: {e}")
                if max(e["F"], e["dF"], e["G"]) > max(1e-8, tol):
                    raise RuntimeError(f"parity gate: log-joint side far from the FP64 port: {e}")
        # ---- truth: binary128 on a reduced number of draws (and optionally of hyper-parameter samples) ----
        gp_t = gp if truth_S is None else dict(gp, post=list(gp["post"][:truth_S]))
        c1.set_precision(64)
        eps_t = c1.eps_philox(D, K, truth_Ns, seed, stream + 1, readback=True)
        c1.set_precision(precision)
        got = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp_t, truth_Ns, 1, 0, 0, tb, 0, epsilon=eps_t, nargout=6, ctx=ctx)
        F, dF, G, H, _, dH = got
        if rank == 0:
            prep = cport.Prepared(vp, gp_t, tb)
            Ft, dFt, Gt, Ht, dHt, _ = cport.negelcbo(prep, theta, truth_Ns, eps_t, truth128=True)
            e = dict(F=_rel(F, Ft), dF=_rel(dF, dFt), G=_rel(G, Gt), H=_rel(H, Ht), dH=_rel(dH, dHt))
            Fp, dFp, Gp, Hp, dHp, _ = cport.negelcbo(prep, theta, truth_Ns, eps_t)
            out["vs_binary128"] = dict(e, Ns=truth_Ns, S=len(gp_t["post"]))
            out["fp64_port_vs_binary128"] = dict(F=_rel(Fp, Ft), dF=_rel(dFp, dFt), G=_rel(Gp, Gt), H=_rel(Hp, Ht), dH=_rel(dHp, dHt))
            if max(e.values()) > tol:
                raise RuntimeError(f"parity gate: CUDA step vs binary128 truth: {e}")
        if truth_S is not None:   # make the full posterior resident again
            ctx.gp_attach(gp)
    finally:
        ctx.set_precision(64)
        if c1 is not ctx:
            c1.close()
    out["ok"] = True
    return out


def guarded_gate(name, dist, local, *a, **kw):
    """Run parity_gate on every rank; a failure on any rank stops all of them (no rank is left waiting in a collective)."""
    err, res = None, None
    try:
        res = parity_gate(*a, **kw)
    except Exception as e:   # noqa: BLE001
        err = f"{name}: {e}"
    bad = 1.0 if err else 0.0
    if dist is not None:
        import torch
        t = torch.tensor([bad], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        bad = float(t.item())
    if bad:
        sys.stderr.write(f"[bench] PARITY GATE FAILED ({err or 'on another rank'}) -- nothing is timed\n")
        sys.stderr.flush()
        if dist is not None:
            dist.destroy_process_group()
        sys.exit(3)
    return res


def run_reference(args, cfg_name, emit):
    """Reference arm: the reference's CPU algorithm (oracle C/OpenMP port) on the host cores."""
    from vbmc_b200 import workloads
    from oracle import cport
    cfg = dict(workloads.CONFIGS[cfg_name])
    w = workloads.build(cfg, scipy_gp_post, with_eps=False)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cport.time_negelcbo(w, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "negelcbo_vbmc grad-steps/sec", "value": res["steps_per_s"], "unit": "steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / res["steps_per_s"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": cfg_name, **{k: cfg[k] for k in ("D", "N", "K", "Ns", "S")}},
            "cpu_baseline": {"value": res["steps_per_s"], "unit": "steps/s", "cores": res["threads"], "kind": "port",
                             "sample": res["sample"]},
            "e2e": {"value": res["steps_per_s"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def quick_value(ctx, vbmc_b200, cfg_name, steps, warmup, dist, local, precision=64, gate=None, rank=0, world=1):
    """Device-resident steps/s of another configuration (same protocol as the headline `value`), behind its own parity gate.
    warmup >= 2: the first call with a signature allocates, the second captures the CUDA graph (host time inside its events)."""
    assert warmup >= 2
    import ctypes as C
    from vbmc_b200 import _lib, workloads
    cfg = dict(workloads.CONFIGS[cfg_name])
    w = workloads.build(cfg, lambda *a: vbmc_b200.gplite_post(*a, ctx=ctx, want_L=False), with_eps=False)
    _, tb = vbmc_b200.vpbounds(w["vp"], w["gp"], workloads.VP_OPTIONS)
    parity = None
    if gate is not None:
        parity = guarded_gate(f"{cfg_name}/{precision}", dist, local, ctx, vbmc_b200, w, tb, cfg, dist, local, rank, world, precision=precision, **gate)
    ctx.vp_set(w["vp"]); ctx.gp_attach(w["gp"]); ctx.thetabnd_set(tb)
    theta = np.ascontiguousarray(w["theta"])
    F, dF, ms = C.c_double(), np.zeros_like(theta), C.c_float()
    a = _lib.NegelcboArgs()
    a.theta, a.ntheta, a.beta, a.Ns = _lib.dptr(theta), theta.size, 0.0, cfg["Ns"]
    a.compute_grad, a.compute_var, a.separate_K, a.use_thetabnd = 1, 0, 0, 1
    a.eps_mode, a.seed, a.stream = _lib.EPS_PHILOX, 20260925, 0
    a.F, a.dF = C.pointer(F), _lib.dptr(dF)
    tot = 0.0
    ctx.set_precision(precision)
    try:
        for i in range(warmup + steps):
            ctx.flush_l2()
            if dist is not None:   # ranks start each timed step together (their L2 flushes are outside everybody's events)
                ctx.sync()
                dist.barrier()
                import torch
                torch.cuda.synchronize()
            a.stream = 50_000 + i
            _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
            if i >= warmup:
                tot += ms.value
    finally:
        ctx.set_precision(64)
    if dist is not None:
        import torch
        t = torch.tensor([tot], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot = float(t.item())
    return {"workload": cfg_name, **{k: cfg[k] for k in ("D", "N", "K", "Ns", "S")}, "steps": steps, "entropy_sweep_bits": precision,
            "ms_per_step": tot / steps, "value": 1e3 * steps / tot, "unit": "steps/s", "parity": parity}


def main():
    # stdout carries exactly ONE JSON line: library banners (e.g. NCCL's version line at NCCL_DEBUG=WARN/VERSION) go to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_name = args.config or "c3"   # the configuration BASELINE.json quotes the metric on, at every N (strong scaling)
    if args.impl == "reference":
        run_reference(args, cfg_name, emit)
        return

    import vbmc_b200
    from vbmc_b200 import _lib, workloads
    import ctypes as C

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = vbmc_b200.Context(local)
    if world > 1:
        import torch
        uid = torch.zeros(_lib.UNIQUE_ID_BYTES, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(vbmc_b200.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))

    cfg = dict(workloads.CONFIGS[cfg_name])
    t_setup = time.time()
    gp_fn = (lambda *a: vbmc_b200.gplite_post(*a, ctx=ctx, want_L=False)) if os.environ.get("VBMC_B200_BENCH_GPU_GP", "1") == "1" \
        else scipy_gp_post
    try:
        w = workloads.build(cfg, gp_fn, with_eps=False)
        gp_source = "vbmc_b200.gplite_post (GPU)" if gp_fn is not scipy_gp_post else "scipy LAPACK (set-up only)"
    except vbmc_b200.VbmcB200Error as e:
        if "NotYet" not in str(e):
            raise
        w = workloads.build(cfg, scipy_gp_post, with_eps=False)
        gp_source = "scipy LAPACK (set-up only)"
    t_setup = time.time() - t_setup
    vp, gp, theta0 = w["vp"], w["gp"], w["theta"]
    # ---- secondary path: GP refit (gplite_post = S x gplite_core), timed per kernel with CUDA events ----
    refit = None
    if gp_source.startswith("vbmc_b200") and rank == 0:
        noisefun = [1, 1, 0] if w["s2"] is not None else [1, 0, 0]
        vbmc_b200.gplite_post(w["hyp"], w["X"], w["y"], 1, 4, noisefun, w["s2"], ctx=ctx, want_L=False)  # warm
        t_refit = float("inf")
        for _ in range(3):   # best of 3 whole calls, like the library bar below (a single call varies by +-0.2 ms with the host)
            ctx.sync()
            t0 = time.perf_counter()
            vbmc_b200.gplite_post(w["hyp"], w["X"], w["y"], 1, 4, noisefun, w["s2"], ctx=ctx, want_L=False)
            ctx.sync()
            t_refit = min(t_refit, time.perf_counter() - t0)
        ctx.profile_reset(); ctx.profile_enable(True)
        gp = vbmc_b200.gplite_post(w["hyp"], w["X"], w["y"], 1, 4, noisefun, w["s2"], ctx=ctx, want_L=False)
        ctx.profile_enable(False)
        kt = {k: ctx.profile_get(k)[0] for k in ("gram", "potrf_potf2", "potrf_trsm", "potrf_update", "trsv", "gp_prep")}
        kt["potrf_panel"] = kt["potrf_potf2"] + kt["potrf_trsm"]   # diagonal-block factorisation + row-panel solve
        N_, S_ = cfg["N"], cfg["S"]
        Np_ = (N_ + 1 + 63) // 64 * 64
        nb_ = Np_ // 64
        upd_flops = S_ * sum((nb_ - kb - 1) * (nb_ - kb) // 2 for kb in range(nb_)) * 2.0 * 64 ** 3
        refit = {"gplite_post_wall_ms": t_refit * 1e3, "S": S_, "N": N_,
                 "end_to_end": {"tflops": S_ * N_ ** 3 / 3.0 / t_refit / 1e12, "frac_of_dmma_peak": S_ * N_ ** 3 / 3.0 / t_refit / 1e12 / 37.1,
                                "what": "S*N^3/3 over the wall time of the whole gplite_post call (Gram, factorisation with look-ahead, "
                                        "solves, host<->device copies of X, y, hyp, alpha)"},
                 "kernels_ms_note": "per-kernel events are taken in the serial (no look-ahead) order",
                 "kernels_ms": {k: round(v, 4) for k, v in kt.items()},
                 "potrf_update_dmma_tflops": upd_flops / (kt["potrf_update"] * 1e-3) / 1e12 if kt["potrf_update"] > 0 else None,
                 "chol_flops_N3_over_3_x_S": S_ * N_ ** 3 / 3.0}
        if refit["potrf_update_dmma_tflops"]:
            # the N^3/3 contraction runs on the FP64 tensor path (mma.sync.m8n8k4.f64 -> DMMA); peak = tools/dmma_probe.cu on B200
            refit["roofline"] = {"kernel": "gp_update_kernel", "bound": "tensor", "achieved": refit["potrf_update_dmma_tflops"], "peak": 37.1,
                                 "unit": "TFLOP/s", "frac": refit["potrf_update_dmma_tflops"] / 37.1,
                                 "peak_source": "FP64 DMMA micro-benchmark tools/dmma_probe.cu, measured on B200 (profiles/r1_ncu_summary.md); "
                                                "ncu: tensor sub-pipe 63.8 % active on the large trailing updates"}
        refit["library_bar"] = library_cholesky_bar(w, cfg, local)
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    Ns = cfg["Ns"]
    counts = algorithmic_counts(cfg)

    # ---- parity gate: nothing is timed unless this configuration's step matches the CPU restatements ----
    parity = None
    if not args.no_parity:
        parity = guarded_gate(cfg_name, dist, local, ctx, vbmc_b200, dict(w, gp=gp), tb, cfg, dist, local, rank, world, full=True, truth_Ns=128)

    # ---- make everything resident, build the args struct for the device-resident loop ----
    ctx.vp_set(vp)
    ctx.gp_attach(gp)
    ctx.thetabnd_set(tb)
    theta = np.ascontiguousarray(theta0)
    F = C.c_double()
    dF = np.zeros_like(theta)
    a = _lib.NegelcboArgs()
    a.theta, a.ntheta, a.beta, a.Ns = _lib.dptr(theta), theta.size, 0.0, Ns
    a.compute_grad, a.compute_var, a.separate_K, a.use_thetabnd = 1, 0, 0, 1
    a.eps_mode, a.seed, a.stream = _lib.EPS_PHILOX, 20260925, 0
    a.F, a.dF = C.pointer(F), _lib.dptr(dF)
    ms = C.c_float()

    def resident_step(i):
        a.stream = i
        _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), 1, C.byref(ms)))
        return ms.value

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # runs through the timed region, the per-kernel pass and the e2e loop
    for i in range(args.warmup):
        resident_step(i)
    barrier()
    l0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    tot_ms = 0.0
    per_step_ms = []
    for i in range(args.steps):
        ctx.flush_l2()               # outside the timed events
        if dist is not None:
            barrier()                # ranks start each timed step together: a host hiccup on one rank is not billed to its peers' events
        v = resident_step(args.warmup + i)
        tot_ms += v
        per_step_ms.append(v)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - l0
    if dist is not None:
        import torch
        t = torch.tensor([tot_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_ms = float(t.item())
    ms_per_step = tot_ms / args.steps
    value = 1e3 / ms_per_step

    # ---- the same steps as ONE resident loop of >= 200 consecutive steps, wall clock between two device syncs (no L2 flush, no gaps):
    # what a caller that keeps theta on the device sees; the per-step events above cannot hide work in un-timed gaps here ----
    nloop = max(200, args.steps)
    barrier()
    tl0 = time.perf_counter()
    a.stream = 300_000
    _lib.check(ctx.lib.vbmc_b200_negelcbo_resident_loop(ctx.handle, C.byref(a), nloop, C.byref(ms)))
    barrier()
    loop_wall = time.perf_counter() - tl0
    loop_dev_ms = ms.value
    if dist is not None:
        import torch
        t = torch.tensor([loop_wall, loop_dev_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        loop_wall, loop_dev_ms = float(t[0].item()), float(t[1].item())
    resident_loop = {"steps": nloop, "wall_s": loop_wall, "steps_per_s_wall": nloop / loop_wall, "steps_per_s_device_events": 1e3 * nloop / loop_dev_ms,
                     "what": "vbmc_b200_negelcbo_resident_loop: consecutive steps, fresh device draws each, host sync every step for F/dF; "
                             "L2 NOT flushed (a step's 65.5 MB of draws are written by the generator right before they are read)"}

    # ---- per-kernel CUDA-event timing of the same steps (roofline of the dominant kernel) ----
    ctx.profile_reset()
    ctx.profile_enable(True)
    nprof = max(3, min(40, args.steps))
    for i in range(nprof):
        ctx.flush_l2()
        resident_step(10_000 + i)
    ctx.profile_enable(False)
    prof = {}
    for name in ("entmc", "gplogjoint", "philox", "reduce", "finalize", "vp_unpack"):
        t_ms, n = ctx.profile_get(name)
        prof[name] = {"ms_per_step": t_ms / nprof, "launches_per_step": n / nprof}

    # ---- how much of the K x K component sweep the entropy kernel actually scores (warp-uniform pruning, csrc/entmc.cu) ----
    ctx.entmc_prune_stats(True)
    resident_step(20_000)
    kept, total = ctx.entmc_prune_stats(False)
    kept_frac = kept / max(1, total)

    # ---- e2e through the public host API: host theta -> F, dF on the host + host Adam ----
    def e2e_loop(nsteps, warm, host_eps=None):
        x = theta0.copy()
        st = {"it": 0, "m": np.zeros_like(x), "v": np.zeros_like(x)}
        for i in range(warm + nsteps):
            if i == warm:
                barrier()
                t0 = time.perf_counter()
            if host_eps is None:
                Fh, dFh = vbmc_b200.negelcbo_vbmc(x, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, rng=(777, i), nargout=2, ctx=ctx)
            else:
                Fh, dFh = vbmc_b200.negelcbo_vbmc(x, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=host_eps, nargout=2, ctx=ctx)
            x = adam_update(st, x, dFh)
        barrier()
        return (time.perf_counter() - t0) / nsteps

    e2e_s = e2e_loop(args.steps, max(3, args.warmup))
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_host_eps = None
    if world == 1:
        heps = workloads.make_epsilon(cfg)
        e2e_host_eps = 1.0 / e2e_loop(max(3, args.steps // 4), 2, host_eps=heps)

    # ---- whole-optimisation call: fminadam with the loop on the device, host x0 in, x/f/xtab/ftab out ----
    fmin = None
    try:
        nit = max(40, args.steps)
        # warm call with the SAME MaxIter: the loop's CUDA graph and the iterate-history buffer are keyed by it, a different
        # value would put the capture, the instantiation and a reallocation inside the timed call (8-80 ms, host dependent)
        vbmc_b200.fminadam_negelcbo(theta0, 0.0, vp, gp, Ns, 0, tb, None, None, 1e-9, nit, None, rng=(778, 0), ctx=ctx)
        barrier()
        t0 = time.perf_counter()
        _, _, xtab_, ftab_, it_ = vbmc_b200.fminadam_negelcbo(theta0, 0.0, vp, gp, Ns, 0, tb, None, None, 1e-9, nit, None,
                                                            rng=(778, 1000), ctx=ctx)
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            import torch
            t = torch.tensor([dt], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        fmin = {"value": it_ / dt, "unit": "steps/s", "iterations": it_, "wall_s": dt, "h2d_bytes_per_call": theta0.size * 8 * 3,
                "d2h_bytes_per_call": int(xtab_.nbytes + ftab_.nbytes + theta0.size * 8 + 64),
                "what": "one vbmc_b200.fminadam_negelcbo call = utils/fminadam.m with negelcbo_vbmc as objective; Adam state, xtab, "
                        "ftab on the device, host reads the termination flag every 20 iterations"}
    except Exception as e:   # reporting only
        fmin = {"error": str(e)[:300]}

    # ---- the callers either side of the path, each timed through the host API (wall clock, synced) ----
    nxt = None
    if rank == 0 and gp_source.startswith("vbmc_b200") and os.environ.get("VBMC_B200_BENCH_NEXT", "1") != "0":
        try:
            nxt = {}
            noisefun = [1, 1, 0] if w["s2"] is not None else [1, 0, 0]
            gpd = vbmc_b200.gplite_post(w["hyp"], w["X"], w["y"], 1, 4, noisefun, w["s2"], ctx=ctx, want_L=False)
            rs = np.random.default_rng(5)
            Xs = w["X"][rs.integers(0, cfg["N"], 1024)] + 0.3 * rs.standard_normal((1024, cfg["D"]))
            vbmc_b200.gplite_pred(gpd, Xs, nargout=2, ctx=ctx)   # warm at the timed shape (right-hand-side buffer, sweep graphs)
            ctx.sync(); t0 = time.perf_counter()
            vbmc_b200.gplite_pred(gpd, Xs, nargout=2, ctx=ctx)
            ctx.sync(); dt = time.perf_counter() - t0
            nxt["gplite_pred"] = {"Nstar": 1024, "S": cfg["S"], "N": cfg["N"], "ms": dt * 1e3,
                                  "trsm_flops": float(cfg["S"]) * cfg["N"] ** 2 * 1024,
                                  "tflops": float(cfg["S"]) * cfg["N"] ** 2 * 1024 / dt / 1e12}
            if w["s2"] is None:
                xnew = w["X"][0] + 0.1
                ctx.sync(); t0 = time.perf_counter()
                gpu1 = vbmc_b200.gplite_post_update1(gpd, xnew, float(w["y"][0]), ctx=ctx)
                ctx.sync(); dt = time.perf_counter() - t0
                nxt["gplite_post_rank1"] = {"ms": dt * 1e3, "full_refit_ms": refit["gplite_post_wall_ms"] if refit else None,
                                            "N_after": int(gpu1["X"].shape[0])}
            hyps = np.repeat(w["hyp"], 2, axis=1)[:, :32] + 0.05 * rs.standard_normal((w["hyp"].shape[0], min(32, 2 * cfg["S"])))
            vbmc_b200.gplite_nlZ_batch(hyps, gpd, None, ctx=ctx)
            ctx.sync(); t0 = time.perf_counter()
            vbmc_b200.gplite_nlZ_batch(hyps, gpd, None, ctx=ctx)
            ctx.sync(); dt = time.perf_counter() - t0
            nxt["gplite_nlZ_batch"] = {"vectors": int(hyps.shape[1]), "ms": dt * 1e3, "ms_per_vector": dt * 1e3 / hyps.shape[1]}
            # make the step's GP resident again for the measurements below
            gp = vbmc_b200.gplite_post(w["hyp"], w["X"], w["y"], 1, 4, noisefun, w["s2"], ctx=ctx, want_L=False)
            ctx.vp_set(vp); ctx.gp_attach(gp); ctx.thetabnd_set(tb)
        except Exception as e:   # reporting only
            nxt = {"error": str(e)[:300]}

    # ---- c4 (Ns=131072): the MC-shard configuration BASELINE.json names for 2/4/8 GPUs, same protocol ----
    c4 = None
    if args.config is None:
        c4 = quick_value(ctx, vbmc_b200, "c4", max(10, args.steps // 5), 3, dist, local, rank=rank, world=world,
                         gate=None if args.no_parity else dict(full=True, truth_Ns=64))
    # ---- c5 (D=20, N=4000, K=100, Ns=262144, S=40): BASELINE.json's FP32 configuration, FP32 and FP64 sweeps ----
    c5 = None
    if args.config is None and os.environ.get("VBMC_B200_BENCH_C5", "1") != "0":
        try:
            g5 = None if args.no_parity else dict(full=True, reduced_Ns=2048, truth_Ns=64, truth_S=2)
            c5 = {"fp32": quick_value(ctx, vbmc_b200, "c5", 5, 2, dist, local, precision=32, gate=g5, rank=rank, world=world),
                  "fp64": quick_value(ctx, vbmc_b200, "c5", 3, 2, dist, local, precision=64, gate=g5, rank=rank, world=world)}
        except Exception as e:   # a sub-measurement must never cost the headline line (every rank fails alike: no collective is left hanging)
            c5 = {"error": str(e)[:300]}
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- roofline ----
    peaks, peak_src = measured_peaks()
    fp64_peak = ctx.measure_fp64_peak()
    ent_ms = prof["entmc"]["ms_per_step"]
    shard = 1.0 / world
    ent_tflops = counts["entmc_flops"] * shard / (ent_ms * 1e-3) / 1e12 if ent_ms > 0 else None
    ent_gbs = counts["entmc_bytes"] * shard / (ent_ms * 1e-3) / 1e9 if ent_ms > 0 else None
    exec_tflops = ent_tflops * kept_frac if ent_tflops else None
    roofline = {"kernel": "entmc_kernel", "bound": "fp64", "achieved": exec_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": exec_tflops / fp64_peak if exec_tflops else None,
                "frac_is": "EXECUTED flops (scored (pair, component) blocks only) / measured DFMA peak; the algorithmic fraction is in `algorithmic`",
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at c3 on one GPU: NOT measured in this run, copied from the
                # ncu capture named in traffic_source (the kernel reads each draw exactly once, so the figure is shape-determined)
                "traffic": TRAFFIC_C3_BYTES if (cfg_name == "c3" and world == 1) else None,
                "traffic_source": TRAFFIC_C3_SOURCE if (cfg_name == "c3" and world == 1) else None,
                "peak_source": "DFMA micro-benchmark run live on this device (MEASURED_PEAKS.json has no FP64 entry)",
                "hbm": {"achieved": ent_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ent_gbs / peaks["hbm_gbs"] if ent_gbs else None,
                        "peak_source": peak_src + " MEASURED_PEAKS.json"},
                "algorithmic": {"flops_per_launch": counts["entmc_flops"] * shard, "bytes_per_launch": counts["entmc_bytes"] * shard,
                                "ms_per_launch": ent_ms, "tflops": ent_tflops, "frac": ent_tflops / fp64_peak if ent_tflops else None},
                "executed": {"kept_fraction": kept_frac, "tflops": ent_tflops * kept_frac if ent_tflops else None,
                             "frac": ent_tflops * kept_frac / fp64_peak if ent_tflops else None,
                             "what": "components whose term is < exp(-50) of q for a whole warp of draws are skipped (below FP64 round-off; "
                                     "parity-tested); `algorithmic` counts the K^2 Ns (5D+12) flops of SURVEY 8d, `executed` (= achieved/frac) only the scored ones"},
                "note": "entmc at K=50 has ~78 flop/B: FP64-pipe bound, not HBM bound (SURVEY.md 8d); both fractions reported"}
    cpu = None
    if not args.no_cpu_baseline:
        try:
            from oracle import cport
            r = cport.time_negelcbo(w, steps=2, warmup=1)
            cpu = {"value": r["steps_per_s"], "unit": "steps/s", "cores": r["threads"], "kind": "port", "sample": r["sample"]}
        except Exception as e:  # the baseline is reporting only; never let it kill the bench line
            cpu = {"value": None, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port", "sample": f"unavailable: {e}"}
    cpu_numpy = None
    if not args.no_cpu_baseline:
        try:
            cpu_numpy = numpy_stand_in(dict(w, gp=gp), cfg)
        except Exception as e:
            cpu_numpy = {"value": None, "unit": "steps/s", "kind": "port", "sample": f"unavailable: {str(e)[:200]}"}
    nth = theta.size
    line = {
        "metric": "negelcbo_vbmc grad-steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step,
        "ms_per_step_percentiles_rank0": {f"p{q}": float(np.percentile(per_step_ms, q)) for q in (10, 50, 90)} if per_step_ms else None,
        "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg_name, **{k: cfg[k] for k in ("D", "N", "K", "Ns", "S")},
                   "parallelism": f"mc-pair-shard x{world} + hyp-sample shard, 1 all-reduce/step" if world > 1 else "single GPU",
                   "allreduce": (("peer memory over NVLink, fused into finalize_kernel (CUDA IPC, no NCCL call per step)" if ctx.comm_p2p()
                                  else "ncclAllReduce") if world > 1 else None),
                   "eps": "device Philox4x32-10 + 1024-strip ziggurat, fresh draws every step (reference: randn per call), generated "
                          "ahead of time in the tail of the previous step",
                   "l2": "flushed (256 MB memset) between timed steps, outside the CUDA events",
                   "gp_posterior_from": gp_source, "setup_s": round(t_setup, 2)},
        "clocks": clocks,
        "e2e": {"value": 1.0 / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": nth * 8, "d2h_bytes_per_step": (nth + 8) * 8,
                "includes": "host theta H2D, F/dF D2H, host Adam update (fminadam.m:51-60); draws from the device generator",
                "host_eps_variant_steps_per_s": e2e_host_eps,
                "host_eps_variant_h2d_bytes_per_step": counts["entmc_bytes"] if e2e_host_eps else None},
        "parity": parity if parity is not None else {"ok": None, "skipped": "--no-parity"},
        "fminadam_device_loop": fmin,
        "next_rows": nxt,
        "gpu_launches": launches,
        "wall_s_timed_region": t_wall,
        "resident_loop": resident_loop,
        "kernels_ms_per_step": {k: round(v["ms_per_step"], 5) for k, v in prof.items()},
        "roofline": roofline,
        "refit": refit,
        "c4": c4,
        "c5": c5,
        "cpu_baseline": cpu,
        "cpu_baseline_numpy": cpu_numpy,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
