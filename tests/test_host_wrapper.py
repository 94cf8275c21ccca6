"""Host mirror (vbmc_b200/api.py) without a GPU: the library entry points are replaced by Python fakes that read the ctypes
argument blocks and write recognisable outputs.  Checks (i) that negelcbo_vbmc's pre-marshalled call frames hand the C ABI
exactly what a freshly built argument block would (every nargout / flag / draw-mode combination, results returned as
fresh arrays), and (ii) that the cheap change probes of vp_set / thetabnd_set never skip an upload that is needed
(in-place edits, permutations, new objects) and do skip the redundant ones."""
import ctypes as C
import itertools
import types

import numpy as np
import pytest

from oracle import vbmc_oracle as orc
from vbmc_b200 import _lib, api, workloads


class FakeLib:
    def __init__(self):
        self.calls = {"vp_set": 0, "thetabnd_set": 0, "gp_attach": 0, "negelcbo": 0}
        self.last = None

    def vbmc_b200_vp_set(self, h, d):
        self.calls["vp_set"] += 1
        d = d._obj
        self.vp_mu0, self.vp_K = d.mu[0], d.K
        return 0

    def vbmc_b200_thetabnd_set(self, h, n, lb, ub, tol, wt, wp):
        self.calls["thetabnd_set"] += 1
        self.bnd = (n, None if not lb else lb[0], tol, wt, wp)
        return 0

    def vbmc_b200_gp_attach(self, *a):
        self.calls["gp_attach"] += 1
        return 0

    def vbmc_b200_gp_set_sn2_mult(self, *a):
        return 0

    def vbmc_b200_negelcbo(self, h, ref):
        self.calls["negelcbo"] += 1
        a = ref._obj
        n = a.ntheta
        th = np.array([a.theta[i] for i in range(n)])
        self.last = dict(ntheta=n, beta=a.beta, Ns=a.Ns, grad=a.compute_grad, var=a.compute_var, sepK=a.separate_K,
                         bnd=a.use_thetabnd, mode=a.eps_mode, seed=a.seed, stream=a.stream,
                         eps0=(a.eps[0] if a.eps else None), has=dict(dF=bool(a.dF), dH=bool(a.dH), I=bool(a.I_sk), J=bool(a.J_sjk)))
        base = th.sum() + 3 * a.beta + 1e-3 * a.Ns + 1e-6 * a.seed + 1e-9 * a.stream + 7 * a.eps_mode + 11 * a.use_thetabnd
        if a.eps:
            base += a.eps[0]
        for i, ptr in enumerate((a.F, a.G, a.H, a.varF, a.varGss, a.varG, a.varH)):
            if ptr:
                ptr[0] = base + i
        if a.dF:
            for i in range(n):
                a.dF[i] = 2 * th[i] + i
        if a.dH:
            for i in range(n):
                a.dH[i] = 3 * th[i] - i
        S, K = self.S, self.K
        if a.I_sk:
            for i in range(S * K):
                a.I_sk[i] = 0.5 * i + base
        if a.J_sjk:
            for i in range(S * K * K):
                a.J_sjk[i] = 0.25 * i - base
        return 0


class FakeCtx(api.Context):
    def __init__(self, S, K):
        self.lib = FakeLib()
        self.lib.S, self.lib.K = S, K
        self._h = None
        self._gp_key = self._vp_key = self._bnd_key = None
        self._vp_probe = self._bnd_probe = None
        self._frames = {}

    def __del__(self):
        pass


def spec_call(ctx, theta, beta, vp, gp, Ns, compute_grad, compute_var, thetabnd, epsilon, rng, nargout):
    """The argument block built from scratch for one call (the specification the call frames must reproduce)."""
    if compute_grad is None:
        compute_grad = nargout > 1
    if beta is None or not np.isfinite(beta):
        beta = 0.0
    if compute_var is None:
        compute_var = (beta != 0) or nargout > 4
    separate_K = nargout > 9
    theta = _lib.f64(theta).ravel()
    S, K = len(gp["post"]), int(vp["K"])
    a = _lib.NegelcboArgs()
    a.theta, a.ntheta = _lib.dptr(theta), theta.size
    a.beta, a.Ns = float(beta), int(Ns)
    a.compute_grad, a.compute_var, a.separate_K = int(bool(compute_grad)), int(compute_var), int(separate_K)
    a.use_thetabnd = int(thetabnd is not None)
    mode, e, seed, stream = api._eps_args(vp, Ns, epsilon, rng) if Ns > 0 else (_lib.EPS_RESIDENT, None, 0, 0)
    a.eps_mode, a.eps, a.seed, a.stream = mode, _lib.dptr(e), seed, stream
    sc = np.zeros(8)
    dF = np.zeros(theta.size) if compute_grad else None
    dH = np.zeros(theta.size) if compute_grad else None
    Isk = np.zeros((K, S)) if separate_K else None
    Jsjk = np.zeros((K, K, S)) if (separate_K and compute_var) else None
    p = sc.ctypes.data_as(_lib.c_double_p)
    off = lambda i: C.cast(C.addressof(p.contents) + 8 * i, _lib.c_double_p)
    a.F, a.G, a.H, a.varF, a.varGss, a.varG, a.varH = off(0), off(1), off(2), off(3), off(4), off(5), off(6)
    a.dF, a.dH, a.I_sk, a.J_sjk = _lib.dptr(dF), _lib.dptr(dH), _lib.dptr(Isk), _lib.dptr(Jsjk)
    ctx.lib.vbmc_b200_negelcbo(None, C.byref(a))
    seen = dict(ctx.lib.last)
    out = (float(sc[0]), dF, float(sc[1]), float(sc[2]), float(sc[3]), dH, float(sc[4]), float(sc[5]), float(sc[6]),
           None if Isk is None else Isk.T.copy(), None if Jsjk is None else Jsjk.transpose(2, 1, 0).copy())
    return out[:max(1, nargout)], seen


@pytest.fixture(scope="module")
def problem():
    cfg = dict(D=3, N=30, K=4, S=2, Ns=16, target="rosenbrock", noisy=False)
    w = workloads.build(cfg, orc.gplite_post)
    _, tb = orc.vpbounds(w["vp"], w["gp"], workloads.VP_OPTIONS)
    return w, tb


def same(a, b):
    if a is None or b is None:
        return a is None and b is None
    return np.array_equal(np.asarray(a), np.asarray(b))


def test_call_frames_reproduce_the_fresh_argument_block(problem):
    w, tb = problem
    vp, gp, theta, eps = w["vp"], w["gp"], w["theta"], w["epsilon"]
    ctx = FakeCtx(S=2, K=4)
    kept = []
    combos = itertools.product([1, 2, 4, 5, 6, 9, 10, 11], [None, 0, 1], [None, 0, 1, 2], [0.0, 1.5, float("nan")],
                               [None, tb], ["eps", "rng", "resident", "Ns0"])
    for i, (nargout, cg, cv, beta, bnd, draw) in enumerate(combos):
        th = theta + 0.01 * i
        kw = dict(epsilon=eps) if draw == "eps" else dict(rng=(5, i)) if draw == "rng" else dict(epsilon="resident") if draw == "resident" else {}
        Ns = 0 if draw == "Ns0" else 16
        got = api.negelcbo_vbmc(th, beta, vp, gp, Ns, cg, cv, 0, bnd, 0, nargout=nargout, ctx=ctx, **kw)
        seen = dict(ctx.lib.last)
        want, seen_spec = spec_call(ctx, th, beta, vp, gp, Ns, cg, cv, bnd, kw.get("epsilon") if draw in ("eps", "resident") else None,
                                    kw.get("rng"), nargout)
        assert seen == seen_spec, (nargout, cg, cv, beta, draw)
        assert len(got) == len(want) and all(same(g, x) for g, x in zip(got, want)), (nargout, cg, cv, beta, draw)
        kept.append((got, want))
    # results of earlier calls are not overwritten by later ones (fresh arrays, not views of the frame's buffers)
    for got, want in kept[::37]:
        assert all(same(g, x) for g, x in zip(got, want))


def test_change_probes_never_skip_a_needed_upload(problem):
    w, tb = problem
    gp, theta, eps = w["gp"], w["theta"], w["epsilon"]
    vp = dict(w["vp"])
    vp["mu"], vp["sigma"] = np.array(vp["mu"], dtype=float), np.array(vp["sigma"], dtype=float)
    tb = dict(tb, lb=np.array(tb["lb"], dtype=float), ub=np.array(tb["ub"], dtype=float))
    ctx = FakeCtx(S=2, K=4)
    call = lambda v=vp, b=tb: api.negelcbo_vbmc(theta, 0.0, v, gp, 16, 1, 0, 0, b, 0, epsilon=eps, nargout=2, ctx=ctx)
    call(); call(); call()
    assert ctx.lib.calls["vp_set"] == 1 and ctx.lib.calls["thetabnd_set"] == 1 and ctx.lib.calls["gp_attach"] == 1
    vp["mu"][0, 0] += 1e-9                      # in-place edit of one element
    call()
    assert ctx.lib.calls["vp_set"] == 2 and ctx.lib.vp_mu0 == vp["mu"][0, 0]
    vp["sigma"][[0, 1]] = vp["sigma"][[1, 0]]   # permutation: same sum, same multiset
    call()
    assert ctx.lib.calls["vp_set"] == 3
    vp["optimize_weights"] = not vp["optimize_weights"]
    call()
    assert ctx.lib.calls["vp_set"] == 4
    vp["optimize_weights"] = not vp["optimize_weights"]
    v2 = {k: (np.array(x) if isinstance(x, np.ndarray) else x) for k, x in vp.items()}   # equal content, new objects: no upload needed
    call(v2)
    n_equal = ctx.lib.calls["vp_set"]
    call(v2); call(v2)
    assert ctx.lib.calls["vp_set"] == n_equal
    tb["ub"][3] += 0.5
    call()
    assert ctx.lib.calls["thetabnd_set"] == 2
    tb["TolCon"] = tb["TolCon"] * 2
    call()
    assert ctx.lib.calls["thetabnd_set"] == 3
    api.negelcbo_vbmc(theta, 0.0, vp, gp, 16, 1, 0, 0, None, 0, epsilon=eps, nargout=2, ctx=ctx)
    assert ctx.lib.calls["thetabnd_set"] == 4 and ctx.lib.bnd[0] == 0
    call()
    assert ctx.lib.calls["thetabnd_set"] == 5    # bounds are sent again after they were cleared
    vp_list = dict(vp, sigma=list(vp["sigma"]))  # entries that are not float64 arrays always take the full path (content key)
    call(vp_list); call(vp_list)
    assert ctx.lib.calls["negelcbo"] > 0


def test_gp_attach_never_reuses_a_stale_posterior(problem):
    """the device posterior must be re-uploaded after ANY change of the gp dict -- in-place edits of a middle
    sample, scalar fields, replaced arrays -- and must not be re-uploaded when nothing changed (content equal, frozen or not)."""
    import copy
    w = problem
    cfg = dict(D=3, N=30, K=4, S=3, Ns=16, target="rosenbrock", noisy=False)
    gp = workloads.build(cfg, orc.gplite_post)["gp"]
    ctx = FakeCtx(3, 4)
    n = lambda: ctx.lib.calls["gp_attach"]
    ctx.gp_attach(gp)
    assert n() == 1
    ctx.gp_attach(gp)
    assert n() == 1                                     # unchanged (content probe: the oracle's arrays are writable)
    gp["post"][1]["alpha"][7] += 1e-9                   # in-place edit of a MIDDLE sample, middle element
    ctx.gp_attach(gp)
    assert n() == 2
    gp["post"][1]["hyp"][2] += 1e-12
    ctx.gp_attach(gp)
    assert n() == 3
    gp["meanfun"] = 1
    ctx.gp_attach(gp)
    assert n() == 4
    gp["meanfun"] = 4
    gp["post"][2]["sn2_mult"] = 10.0
    ctx.gp_attach(gp)
    assert n() == 6 - 1
    gp["X"][3, 1] += 1e-9
    ctx.gp_attach(gp)
    assert n() == 6
    gp["y"][5] -= 1e-9
    ctx.gp_attach(gp)
    assert n() == 7
    gp2 = copy.deepcopy(gp)                             # a NEW dict (CPython may reuse ids) with the same content: no upload
    ctx.gp_attach(gp2)
    assert n() == 7
    gp2["post"][0]["alpha"] = gp2["post"][0]["alpha"] * (1 + 1e-15)   # replaced array
    ctx.gp_attach(gp2)
    assert n() == 8
    # a posterior attached for S = 3 is not the resident one of its one-sample sub-struct (activeimportancesampling_vbmc.m:170-171)
    gp1 = dict(gp2, post=[gp2["post"][0]])
    ctx.gp_attach(gp1)
    assert n() == 9
    ctx.gp_attach(gp2)
    assert n() == 10
    # frozen arrays (what gplite_post returns): identity is enough, and an attempted in-place edit raises instead of going unnoticed
    api._freeze(gp2["X"], gp2["y"], *[a for p in gp2["post"] for a in (p["alpha"], p["hyp"], p["sW"], p["L"])])
    ctx.gp_attach(gp2)
    assert n() == 10 and ctx._gp_key.frozen is True     # same content: no upload; from now on identity alone decides
    ctx.gp_attach(gp2)
    assert n() == 10
    with pytest.raises(ValueError):
        gp2["post"][1]["alpha"][0] = 0.0
    # the factors: a posterior attached without L is not enough for a caller that needs them
    ctx.gp_attach(gp2, want_L=True)
    assert n() == 11
    ctx.gp_attach(gp2, want_L=False)
    assert n() == 11
