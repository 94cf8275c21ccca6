"""ctypes wrapper of oracle/c/libvbmc_oracle.so — the C/OpenMP restatement of the hot path.
TEST INFRASTRUCTURE (see oracle/vbmc_oracle.py): used by tests (three-way agreement) and by
bench.py's cpu_baseline / --impl reference legs only."""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
import time
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent / "c"
_LIB = _DIR / "libvbmc_oracle.so"
_LIB128 = _DIR / "libvbmc_truth128.so"   # the same C source compiled in IEEE binary128 (make -C oracle/c): "truth"
_lib = None
_lib128 = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def host_threads():
    """Threads the CPU baseline should use: the cgroup CPU quota (x2 for SMT) capped by the affinity mask.
    (On the gpurun boxes nproc = 128 but cpu.max = 16 CPUs; 128 OpenMP threads thrash there.)"""
    aff = len(os.sched_getaffinity(0))
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            return max(1, min(aff, 2 * int(math.ceil(int(q) / int(per)))))
    except Exception:
        pass
    return aff


def load_truth128():
    global _lib128
    if _lib128 is None:
        load()
        if not _LIB128.exists():
            subprocess.check_call(["make", "-C", str(_DIR)])
        _lib128 = C.CDLL(str(_LIB128))
        _lib128.vbmc_truth128_negelcbo.restype = C.c_int
        _lib128.vbmc_truth128_set_threads(int(os.environ.get("VBMC_ORACLE_THREADS", host_threads())))
    return _lib128


def load():
    global _lib
    if _lib is None:
        # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must not inherit that,
        # so the thread count is set explicitly unless the caller asks for one with VBMC_ORACLE_THREADS.
        os.environ["OMP_NUM_THREADS"] = os.environ.get("VBMC_ORACLE_THREADS", str(host_threads()))
        if not _LIB.exists():
            subprocess.check_call(["make", "-C", str(_DIR)])
        _lib = C.CDLL(str(_LIB))
        _lib.vbmc_oracle_threads.restype = C.c_int
        _lib.vbmc_oracle_negelcbo.restype = C.c_int
        _lib.vbmc_oracle_set_threads(int(os.environ.get("VBMC_ORACLE_THREADS", host_threads())))
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(dp)


def _c(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class Prepared:
    """Column-major host buffers of one (vp, gp, thetabnd) problem."""

    def __init__(self, vp, gp, thetabnd):
        self.D, self.K = int(vp["D"]), int(vp["K"])
        X = _c(gp["X"])
        self.N = X.shape[0]
        self.X = np.ascontiguousarray(X.T)
        post = gp["post"]
        self.S = len(post)
        self.hyp = _c(np.stack([np.asarray(p["hyp"]).ravel() for p in post]))
        self.Nhyp = self.hyp.shape[1]
        self.alpha = _c(np.stack([np.asarray(p["alpha"]).ravel() for p in post]))
        self.Ncov, self.Nnoise, self.meanfun = int(gp["Ncov"]), int(gp["Nnoise"]), int(gp["meanfun"])
        self.opt = np.ascontiguousarray([int(bool(vp[f])) for f in ("optimize_mu", "optimize_sigma", "optimize_lambda", "optimize_weights")], dtype=np.int32)
        self.mu = np.ascontiguousarray(_c(vp["mu"]).reshape(self.D, self.K).T)
        self.sigma, self.lam, self.w = _c(vp["sigma"]).ravel(), _c(vp["lambda"]).ravel(), _c(vp["w"]).ravel()
        self.eta = _c(vp["eta"]).ravel() if vp.get("eta") is not None else np.log(self.w)
        d = vp.get("delta")
        self.delta = None if d is None or np.size(d) == 0 else _c(d).ravel() * np.ones(self.D)
        if thetabnd is None:
            self.nbnd, self.lb, self.ub, self.tol, self.wt, self.wp = 0, None, None, 0.0, 0.0, 0.0
        else:
            self.lb, self.ub = _c(thetabnd["lb"]).ravel(), _c(thetabnd["ub"]).ravel()
            self.nbnd, self.tol = self.lb.size, float(thetabnd["TolCon"])
            self.wt, self.wp = float(thetabnd.get("WeightThreshold", 0.0)), float(thetabnd.get("WeightPenalty", 0.0))


def negelcbo(prep: Prepared, theta, Ns, eps, compute_grad=True, truth128=False):
    """(F, dF, G, H, dH, I_sk) of negelcbo_vbmc(theta,0,vp,gp,Ns,compute_grad,0,0,thetabnd).
    ``truth128``: evaluate every intermediate in IEEE binary128 (results rounded to double at the end)."""
    lib = load_truth128() if truth128 else load()
    theta = _c(theta).ravel()
    eps = _c(eps)
    F, G, H = C.c_double(), C.c_double(), C.c_double()
    dF, dH = np.zeros(theta.size), np.zeros(theta.size)
    Isk = np.zeros((prep.S, prep.K))
    fn = lib.vbmc_truth128_negelcbo if truth128 else lib.vbmc_oracle_negelcbo
    rc = fn(
        prep.D, prep.K, prep.N, prep.S, prep.Nhyp, prep.Ncov, prep.Nnoise, prep.meanfun, _p(prep.X), _p(prep.hyp), _p(prep.alpha),
        _p(theta), theta.size, prep.opt.ctypes.data_as(ip), _p(prep.mu), _p(prep.sigma), _p(prep.lam), _p(prep.w), _p(prep.eta),
        _p(prep.delta), int(Ns), _p(eps), prep.nbnd, _p(prep.lb), _p(prep.ub), C.c_double(prep.tol), C.c_double(prep.wt),
        C.c_double(prep.wp), int(bool(compute_grad)), C.byref(F), _p(dF), C.byref(G), C.byref(H), _p(dH), _p(Isk))
    if rc != 0:
        raise RuntimeError(f"vbmc_oracle_negelcbo failed rc={rc}")
    return F.value, (dF if compute_grad else None), G.value, H.value, (dH if compute_grad else None), Isk


def time_negelcbo(w, steps=3, warmup=1, max_seconds=60.0):
    """Time the C/OpenMP port on the host cores on workload dict ``w`` (vbmc_b200.workloads.build).
    One step = one full negelcbo evaluation with gradient at the workload's Ns (draws pre-generated:
    the reference's own randn time is NOT charged to the CPU)."""
    from vbmc_b200 import workloads, api
    lib = load()
    cfg = w["cfg"]
    _, tb = api.vpbounds(w["vp"], w["gp"], workloads.VP_OPTIONS)
    prep = Prepared(w["vp"], w["gp"], tb)
    eps = workloads.make_epsilon(cfg)
    theta = w["theta"]
    for _ in range(warmup):
        negelcbo(prep, theta, cfg["Ns"], eps)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        negelcbo(prep, theta, cfg["Ns"], eps)
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = (time.perf_counter() - t0) / done
    return {"steps_per_s": 1.0 / dt, "threads": lib.vbmc_oracle_threads(), "steps": done,
            "sample": f"{done} full steps of {cfg.get('name', '')}D={cfg['D']},N={cfg['N']},K={cfg['K']},Ns={cfg['Ns']},S={cfg['S']} "
                      f"(C/OpenMP port of the reference algorithm, {lib.vbmc_oracle_threads()} threads, draws pre-generated)"}
