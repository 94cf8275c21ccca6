"""Host restatement of the cost-weighted schedule of the FP64 entropy sweep (vbmc_b200/csrc/finalize.cu, last CTA of
vp_unpack2_kernel; consumed by entmc2_kernel / entmc2_reduce_kernel).  Test helper: integer arithmetic identical to the device's."""
import numpy as np


def emax_quantiles(D):
    """10 / 50 / 90 % quantiles of the largest ||eps|| among a warp's 32 draws (Wilson-Hilferty, as in launch_vp_unpack)."""
    out = []
    for z in (1.480, 2.026, 2.718):
        wh = 1.0 - 2.0 / (9.0 * D) + z * np.sqrt(2.0 / (9.0 * D))
        out.append(float(np.sqrt(D * wh ** 3)))
    return out


def survivors3(mu, sigma, lam, w, prune_c):
    """cnt3[j] = sum over the three ||eps||max quantiles of the number of components k that pass the sweep's pruning test for
    source component j (thirds of a component).  mu is (K, D)."""
    K, D = mu.shape
    cnt = np.zeros(K, dtype=np.int64)
    for j in range(K):
        u = (mu[j][None, :] - mu) / (sigma[:, None] * lam[None, :])
        un = np.sqrt((u * u).sum(1))
        r = sigma[j] / sigma
        with np.errstate(divide="ignore", invalid="ignore"):
            lck = np.log(w / w[j]) + D * (np.log(sigma[j]) - np.log(sigma))
        tot = 0
        for emax in emax_quantiles(D):
            tt = un - r * emax
            bb = np.where(tt > 0, -0.5 * tt * tt, 0.0)
            lhs = bb + 0.5 * emax * emax + prune_c + lck + 0.5
            keep = ~(lhs < 0.0) if prune_c > 0 else np.ones(K, dtype=bool)
            tot += int(keep.sum())
        cnt[j] = max(1, tot)
    return cnt


def plan(weights, tpc, G, crun=0):
    """tstart[G+1], jlo[K], jhi[K] for tile weights `weights[j]` (every tile of component j weighs the same) and a fixed cost
    `crun` charged at the start of every component (same units as the weights)."""
    K = len(weights)
    pref = np.concatenate([[0], np.cumsum(np.asarray(weights, dtype=np.int64))])
    C = pref * tpc + np.arange(K + 1, dtype=np.int64) * crun   # cumulative cost in front of component j
    Wtot = int(C[K])
    tstart = np.zeros(G + 1, dtype=np.int64)
    for b in range(G + 1):
        target = (Wtot * b) // G
        lo = int(np.searchsorted(C, target, side="right")) - 1   # largest jj with C[jj] <= target
        lo = min(lo, K)
        if lo >= K:
            t = K * tpc
        else:
            wj = int(weights[lo])
            off = target - int(C[lo]) - crun
            q = 0 if off <= 0 else (2 * off + wj) // (2 * wj)   # nearest tile boundary
            t = lo * tpc + min(q, tpc)
        tstart[b] = t
    tstart[0], tstart[G] = 0, K * tpc
    jlo = np.zeros(K, dtype=np.int64)
    jhi = np.zeros(K, dtype=np.int64)
    for j in range(K):
        tlo, thi = j * tpc, (j + 1) * tpc
        jlo[j] = min([b for b in range(G) if tstart[b + 1] > tlo], default=G - 1)
        jhi[j] = max([b for b in range(G) if tstart[b] < thi], default=0)
    return tstart, jlo, jhi


def rmax_bound(K, tpc, G, c0, crun=0):
    """slots per CTA the host reserves (entmc2.cu make_plan2); c0, crun in whole scored components"""
    T = K * tpc
    per = (T + G - 1) // G
    per = (per * (c0 + K) + c0 - 1) // c0 + 3 + (K * crun) // (G * c0)
    return min(K + 1, (per + tpc - 1) // tpc + 1)
