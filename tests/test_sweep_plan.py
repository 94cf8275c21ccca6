"""Cost-weighted schedule of the FP64 entropy sweep: the host restatement's invariants (CPU) and the device table (GPU)."""
import numpy as np
import pytest

import _sweep_plan as sp
from vbmc_b200 import workloads


def _mixture(name="c3", **over):
    cfg = dict(workloads.CONFIGS[name])
    cfg.update(over)
    X, y, _ = workloads.make_training_set(cfg)
    vp = workloads.make_vp(cfg, X, y)
    mu = np.asarray(vp["mu"], dtype=float)
    if mu.shape != (cfg["K"], cfg["D"]):
        mu = mu.T
    return cfg, vp, np.ascontiguousarray(mu), np.asarray(vp["sigma"], float), np.asarray(vp["lambda"], float), np.asarray(vp["w"], float)


def _check_invariants(tstart, jlo, jhi, K, tpc, G, c0, crun=0):
    assert tstart[0] == 0 and tstart[G] == K * tpc
    assert np.all(np.diff(tstart) >= 0)
    rmax = sp.rmax_bound(K, tpc, G, c0, crun)
    for b in range(G):
        if tstart[b] < tstart[b + 1]:
            runs = (tstart[b + 1] - 1) // tpc - tstart[b] // tpc + 1
            assert runs <= rmax, (b, runs, rmax)
    for j in range(K):
        owners = [b for b in range(G) if tstart[b] < tstart[b + 1] and tstart[b] < (j + 1) * tpc and tstart[b + 1] > j * tpc]
        assert owners, j
        assert jlo[j] <= owners[0] and jhi[j] >= owners[-1]
        # CTAs between jlo and jhi that are not owners must be empty (the reduction skips them)
        for b in range(int(jlo[j]), int(jhi[j]) + 1):
            assert b in owners or tstart[b] >= tstart[b + 1]


@pytest.mark.parametrize("K,tpc,G,c0", [(50, 64, 148, 16), (2, 1, 2, 16), (7, 3, 5, 1), (256, 2, 148, 16), (160, 9, 148, 4),
                                         (50, 3, 148, 16), (3, 50, 148, 16), (20, 16, 37, 64)])
def test_plan_invariants_random_weights(K, tpc, G, c0):
    rs = np.random.default_rng(K * 1000 + tpc)
    G = min(G, K * tpc)
    for trial in range(6):
        cnt = rs.integers(1, K + 1, K) if trial % 2 else np.where(rs.random(K) < 0.5, 1, K)
        crun = (0, 40, 7)[trial % 3]
        tstart, jlo, jhi = sp.plan(c0 + cnt, tpc, G, crun)
        _check_invariants(tstart, jlo, jhi, K, tpc, G, c0, crun)


def test_plan_balances_the_c3_mixture():
    """On the benchmark mixture (two clusters: 17 components that see 17, 33 that see 33) equal-count ranges leave the slowest
    CTA 12-17 % above the mean; cost-weighted ranges bring it within tile granularity."""
    cfg, vp, mu, sigma, lam, w = _mixture("c3")
    K, D = cfg["K"], cfg["D"]
    cnt = sp.survivors3(mu, sigma, lam, w, 50.0)
    assert sorted(set(cnt.tolist())) == [3 * 17, 3 * 33]
    tpc, G, c0, crun = 64, 148, 16, 40
    tstart, jlo, jhi = sp.plan(3 * c0 + cnt, tpc, G, 3 * crun)
    _check_invariants(tstart, jlo, jhi, K, tpc, G, c0, crun)
    cost = np.repeat(3 * c0 + cnt, tpc).astype(float)

    def loads(ts):   # tiles + one table build per source component a range touches
        out = []
        for b in range(G):
            a, e = int(ts[b]), int(ts[b + 1])
            out.append(cost[a:e].sum() + 3 * crun * ((e - 1) // tpc - a // tpc + 1) if a < e else 0.0)
        return np.array(out)

    eq = np.array([(b * K * tpc + G - 1) // G for b in range(G + 1)])
    load_eq, load_w = loads(eq), loads(tstart)
    assert load_eq.max() / load_eq.mean() > 1.10
    assert load_w.max() / load_w.mean() < 1.05


# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,over", [("c3", dict(N=200, S=2, Ns=8192)), ("c3", dict(N=200, S=2, Ns=2048, K=160)),
                                       ("c2", dict(N=100, S=2, Ns=4096))])
def test_device_plan_matches_host_restatement_and_results_do_not_depend_on_it(gpu_ctx, name, over):
    import vbmc_b200
    from oracle import vbmc_oracle as orc
    cfg = dict(workloads.CONFIGS[name]); cfg.update(over)
    w = workloads.build(cfg, orc.gplite_post)
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    _, tb = vbmc_b200.vpbounds(vp, gp, workloads.VP_OPTIONS)
    K, D, Ns = cfg["K"], cfg["D"], cfg["Ns"]
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1e-300, np.max(np.abs(b))))
    try:
        gpu_ctx.entmc_balance(False)
        off = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6, ctx=gpu_ctx)
        assert gpu_ctx.entmc_plan() is None
        for c0 in (16, 3):
            gpu_ctx.entmc_balance(True, c0)
            on = vbmc_b200.negelcbo_vbmc(theta, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6, ctx=gpu_ctx)
            got = gpu_ctx.entmc_plan()
            assert got is not None
            tstart, jlo, jhi, tpc = got
            G = len(tstart) - 1
            _check_invariants(tstart.astype(np.int64), jlo, jhi, K, tpc, G, c0, 0)
            mu = np.asarray(vp["mu"], float)
            mu = mu if mu.shape == (K, D) else mu.T
            cnt = sp.survivors3(np.ascontiguousarray(mu), np.asarray(vp["sigma"], float), np.asarray(vp["lambda"], float),
                                np.asarray(vp["w"], float), 50.0)
            ts_h, jlo_h, jhi_h = sp.plan(3 * c0 + cnt, tpc, G, 0)
            assert np.max(np.abs(ts_h - tstart)) <= 2, (ts_h, tstart)
            for i in (0, 1, 3, 5):   # F, dF, H, dH
                assert rel(on[i], off[i]) < 1e-13, (c0, i)
    finally:
        gpu_ctx.entmc_balance(True, 16)
    ref = orc.negelcbo_vbmc(theta, 0.0, vp, gp, Ns, 1, 0, 0, tb, 0, epsilon=eps, nargout=6)
    for i in (3, 5):
        assert rel(on[i], ref[i]) < 1e-10
