"""Host-side mirror of the reference's MATLAB functions for the hot path, on top of the C ABI.

Same names, argument order/meaning and error identifiers as the reference:

    negelcbo_vbmc(theta,beta,vp,gp,Ns,compute_grad,compute_var,altent_flag,thetabnd,entropy_alpha)
        misc/negelcbo_vbmc.m:1
    gplogjoint(vp,gp,grad_flags,avg_flag,jacobian_flag,compute_var,separate_K)   misc/gplogjoint.m:1
    entmc_vbmc(vp,Ns,grad_flags,jacobian_flag)                                   ent/entmc_vbmc.m:1
    gplite_post(hyp,X,y,covfun,meanfun,noisefun,s2)                              gplite/gplite_post.m:1
    gplite_nlZ(hyp,gp,hprior)                                                    gplite/gplite_nlZ.m:1

MATLAB structs are dicts (``vp``: D,K,mu(D,K),sigma(K),lambda(D),w(K),eta,delta,optimize_*;
``gp``: X(N,D),y,s2,covfun,meanfun,noisefun,Ncov,Nnoise,Nmean,post=[{hyp,alpha,sW,L,sn2_mult,Lchol}]).
MATLAB's ``nargout`` is an explicit keyword.  The reference draws the entropy samples from
MATLAB's global ``randn`` stream (entmc_vbmc.m:53); here they come either from ``epsilon``
(shape (K, Ns/2, D), parity mode) or from the device Philox generator (``rng=(seed, stream)``).

All numerics run in libvbmc_b200.so on the GPU; nothing here computes the path on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import VbmcB200Error, dptr, f64

__all__ = [
    "Context", "default_context", "negelcbo_vbmc", "fminadam_negelcbo", "gplogjoint", "entmc_vbmc", "gplite_post", "gplite_nlZ", "gplite_nlZ_batch", "gplite_pred", "gplite_post_update1",
    "vpbounds", "rescale_params", "get_vptheta", "VbmcB200Error",
]


def _fingerprint(*arrays):
    return hash(tuple(a.tobytes() if isinstance(a, np.ndarray) else repr(a) for a in arrays))


# ---- cheap "did this dict change since the last call" probes -----------------------------------------------------------
# A gradient step calls negelcbo_vbmc with the SAME vp / thetabnd objects a few thousand times; re-marshalling and hashing
# them costs more host time than everything else in the wrapper.  The probe is (identity of the dict and of its arrays,
# a dot product of every array with fixed pseudo-random weights): an in-place edit of any entry changes it (a permutation
# too), and a new object always takes the full path.
_PROBE_W = np.random.Generator(np.random.Philox(20260925)).standard_normal(1 << 14)


def _probe(entries):
    """(ids, value) for a tuple of dict entries, or None when an entry is not a float64 ndarray small enough to probe."""
    ids, val = [], 0.0
    for x in entries:
        if x is None:
            ids.append(0)
            continue
        if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.size <= _PROBE_W.size):
            return None
        ids.append(id(x))
        val += float(x.ravel() @ _PROBE_W[:x.size]) + 0.125 * x.size
    return tuple(ids), val


# ---- "is this gp the posterior that is resident on the device?" -------------------------------------------------------------
# Two levels.  (1) identity: the same array OBJECTS in the same places, all of them read-only (the arrays of a gp dict built by
# gplite_post / gplite_post_update1 are frozen, so they cannot have been edited in place) plus every scalar field -- no data is
# read, ~10 us.  (2) content: shapes, scalars and a weighted sum of EVERY array (X, y, s2 and hyp, alpha, sW(1), sn2_mult, Lchol of
# every sample): any edit, in place or not, of any sample changes it; used for dicts that did not come from this module or that
# hold a writable array.  A stale device posterior is never reused silently.
def _probe_big(x):
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    n, m = x.size, _PROBE_W.size
    val = 0.125 * n
    for o in range(0, n, m):
        c = x[o:o + m]
        val += float(c @ _PROBE_W[:c.size]) * (1.0 + 1e-3 * (o // m))
    return val


def _gp_scalars(gp):
    cov = gp.get("covfun", 1)
    cov = int(cov[0] if isinstance(cov, (list, tuple, np.ndarray)) else cov)
    return (cov, int(gp.get("meanfun", 1)), tuple(int(v) for v in gp.get("noisefun", [1, 0, 0])), len(gp["post"]))


def _gp_identity(gp):
    """(key, frozen): ids of every array + scalars; frozen is True when no array can be edited in place."""
    arrs = [gp["X"], gp.get("y"), gp.get("s2")]
    key = [id(a) for a in arrs]
    for p in gp["post"]:
        pa = (p["alpha"], p["hyp"], p.get("sW"), p.get("L"))
        arrs.extend(pa)
        key.extend(id(a) for a in pa)
        key.append((id(p), float(p.get("sn2_mult", 1.0) or 1.0), bool(p.get("Lchol", True))))
    frozen = all(a is None or (isinstance(a, np.ndarray) and not a.flags.writeable) for a in arrs)
    return (_gp_scalars(gp), tuple(key)), frozen


def _gp_content(gp, with_L):
    X = np.asarray(gp["X"])
    vals = [X.shape, _gp_scalars(gp), _probe_big(X)]
    for name in ("y", "s2"):
        vals.append(None if gp.get(name) is None else _probe_big(gp[name]))
    for p in gp["post"]:
        vals.append((_probe_big(p["hyp"]), _probe_big(p["alpha"]), float(np.ravel(p["sW"])[0]), float(p.get("sn2_mult", 1.0) or 1.0),
                     bool(p.get("Lchol", True)), _probe_big(p["L"]) if (with_L and p.get("L") is not None) else None))
    return tuple(vals)


def _freeze(*arrays):
    for a in arrays:
        if isinstance(a, np.ndarray):
            a.flags.writeable = False


class FrozenStruct(dict):
    """Read-only dict: what gplite_post / gplite_post_update1 return (the gp struct, and every gp['post'][s]).  Nothing inside
    can be replaced or edited in place, so "is this the resident posterior?" is ONE identity comparison per call instead of a
    walk over all S samples (the walk cost ~50 us per negelcbo_vbmc call at S = 20).  To modify, copy first: ``dict(gp)`` /
    ``copy.deepcopy(gp)`` give ordinary writable dicts (MATLAB value semantics), which take the content-probe path."""
    __slots__ = ()

    def _ro(self, *a, **k):
        raise TypeError("this struct is read-only (it mirrors a device-resident posterior): copy it with dict(gp) before editing")

    __setitem__ = __delitem__ = update = pop = popitem = clear = setdefault = __ior__ = _ro

    def __copy__(self):
        return dict(self)

    def __deepcopy__(self, memo):
        import copy
        return {k: copy.deepcopy(v, memo) for k, v in self.items()}

    def __reduce__(self):
        return (dict, (dict(self),))


def _freeze_gp(gp):
    """Freeze every array, turn gp['post'] into a tuple of FrozenStruct and gp itself into a FrozenStruct."""
    posts = []
    for p in gp["post"]:
        _freeze(*[p.get(k) for k in ("hyp", "alpha", "sW", "L")])
        posts.append(p if isinstance(p, FrozenStruct) else FrozenStruct(p))
    _freeze(gp.get("X"), gp.get("y"), gp.get("s2"))
    out = dict(gp)
    out["post"] = tuple(posts)
    if isinstance(out.get("noisefun"), list):
        out["noisefun"] = tuple(out["noisefun"])
    return FrozenStruct(out)


class _GpResident:
    """What the device holds.  For a FrozenStruct produced by this module nothing is read at construction: the struct cannot
    change, so its identity key and content probes are computed from it the first time a DIFFERENT object is compared with it
    (0.5 ms of the 0.8 ms of Python per gplite_post call at S = 20, N = 2000 otherwise)."""
    __slots__ = ("_ident", "_frozen", "_content", "_content_L", "_with_L", "has_L", "ref")

    def __init__(self, gp, has_L):
        self.ref = gp if isinstance(gp, FrozenStruct) else None   # strong reference: the id cannot be recycled while it is resident
        self.has_L = has_L
        self._with_L = has_L and all(p.get("L") is not None for p in gp["post"])
        self._ident = self._frozen = self._content = self._content_L = None
        if self.ref is None:
            self._ident, self._frozen = _gp_identity(gp)
            self._content = _gp_content(gp, False)
            self._content_L = _gp_content(gp, True) if self._with_L else None

    def _fill_ident(self):
        if self._ident is None:
            self._ident, self._frozen = _gp_identity(self.ref)

    @property
    def ident(self):
        self._fill_ident()
        return self._ident

    @ident.setter
    def ident(self, v):
        self._ident = v

    @property
    def frozen(self):
        self._fill_ident()
        return self._frozen

    @frozen.setter
    def frozen(self, v):
        self._frozen = v

    @property
    def content(self):
        if self._content is None:
            self._content = _gp_content(self.ref, False)
        return self._content

    @property
    def content_L(self):
        if self._content_L is None and self._with_L and self.ref is not None:
            self._content_L = _gp_content(self.ref, True)
        return self._content_L


class _CallFrame:
    """Pre-marshalled argument block of one negelcbo_vbmc signature: the ctypes struct, its output buffers and the pointers
    between them are built once; a call copies theta in, sets the scalars and copies the results out."""
    __slots__ = ("a", "theta", "sc", "dF", "dH", "Isk", "Jsjk")

    def __init__(self, ntheta, K, S, grad, var, sepK):
        self.a = a = _lib.NegelcboArgs()
        self.theta = np.zeros(ntheta)
        self.sc = np.zeros(8)
        self.dF = np.zeros(ntheta) if grad else None
        self.dH = np.zeros(ntheta) if grad else None
        self.Isk = np.zeros((K, S)) if sepK else None
        self.Jsjk = np.zeros((K, K, S)) if (sepK and var) else None
        a.theta, a.ntheta = dptr(self.theta), ntheta
        p = self.sc.ctypes.data_as(_lib.c_double_p)
        off = lambda i: C.cast(C.addressof(p.contents) + 8 * i, _lib.c_double_p)
        a.F, a.G, a.H, a.varF, a.varGss, a.varG, a.varH = off(0), off(1), off(2), off(3), off(4), off(5), off(6)
        a.dF, a.dH, a.I_sk, a.J_sjk = dptr(self.dF), dptr(self.dH), dptr(self.Isk), dptr(self.Jsjk)


class Context:
    """One GPU context (vbmc_b200_create).  Caches which gp / vp / thetabnd are resident."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self.lib.vbmc_b200_create(C.byref(self._h), int(device)))
        self.device = device
        self._gp_key = None
        self._vp_key = None
        self._bnd_key = None
        self._keep = []  # host arrays referenced by in-flight descriptors
        self._vp_probe = None    # (id(vp), flags, ids, value) of the vp dict last marshalled (see _probe)
        self._bnd_probe = None
        self._frames = {}        # negelcbo_vbmc call frames by signature

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if self._h:
            self.lib.vbmc_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def sync(self):
        _lib.check(self.lib.vbmc_b200_sync(self._h))

    def set_precision(self, bits: int):
        """64 (default) or 32: arithmetic of the entropy sweep (vbmc_b200_set_precision in include/vbmc_b200.h)."""
        _lib.check(self.lib.vbmc_b200_set_precision(self._h, int(bits)))

    def entmc_prune(self, log_threshold: float):
        """Skip mixture components below exp(-log_threshold) of q for a whole warp of draws (0: score everything)."""
        _lib.check(self.lib.vbmc_b200_entmc_prune(self._h, float(log_threshold)))

    def entmc_prune_stats(self, enable=True):
        """(kept, total) (warp, component) blocks since the last call; ``enable`` keeps the counters running."""
        k, t = C.c_ulonglong(), C.c_ulonglong()
        _lib.check(self.lib.vbmc_b200_entmc_prune_stats(self._h, int(bool(enable)), C.byref(k), C.byref(t)))
        return k.value, t.value

    def entmc_balance(self, on=True, c0=0):
        """Cost-weighted (default) or equal-count tile ranges of the FP64 entropy sweep (vbmc_b200_entmc_balance)."""
        _lib.check(self.lib.vbmc_b200_entmc_balance(self._h, int(bool(on)), int(c0)))

    def entmc_plan(self):
        """(tstart[G+1], jlo[K], jhi[K], tpc) of the last balanced sweep, or None when the last step used equal counts."""
        out = np.zeros(2048, dtype=np.int32)
        G, tpc = C.c_int(), C.c_int()
        _lib.check(self.lib.vbmc_b200_entmc_plan_get(self._h, out.ctypes.data_as(C.POINTER(C.c_int)), out.size, C.byref(G), C.byref(tpc)))
        if G.value == 0:
            return None
        g, t = G.value, tpc.value
        K = int(out[g]) // t   # tstart[G] = K * tpc
        return out[:g + 1].copy(), out[g + 1:g + 1 + K].copy(), out[g + 1 + K:g + 1 + 2 * K].copy(), t

    def launch_count(self) -> int:
        n = C.c_longlong()
        _lib.check(self.lib.vbmc_b200_launch_count(self._h, C.byref(n)))
        return n.value

    # -- multi-GPU --------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(_lib.UNIQUE_ID_BYTES)
        _lib.check(_lib.load().vbmc_b200_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes | None):
        buf = C.create_string_buffer(uid, _lib.UNIQUE_ID_BYTES) if uid is not None else None
        _lib.check(self.lib.vbmc_b200_comm_init(self._h, nranks, rank, buf))

    def comm_p2p(self) -> bool:
        """True when the per-step all-reduce runs inside finalize_kernel over NVLink peer memory (else NCCL)."""
        v = C.c_int()
        _lib.check(self.lib.vbmc_b200_comm_p2p(self._h, C.byref(v)))
        return bool(v.value)

    # -- GP ---------------------------------------------------------------------------------
    @staticmethod
    def _gp_desc(gp, hyp, keep):
        X = f64(gp["X"])
        N, D = X.shape
        Xc = np.ascontiguousarray(X.T)  # column-major N x D
        hyp = f64(hyp)
        if hyp.ndim == 1:
            hyp = hyp[None, :]
        S, Nhyp = hyp.shape
        d = _lib.GpDesc()
        d.N, d.D, d.S, d.Nhyp = N, D, S, Nhyp
        cov = gp.get("covfun", 1)
        d.covfun = int(cov[0] if isinstance(cov, (list, tuple, np.ndarray)) else cov)
        d.meanfun = int(gp.get("meanfun", 1))
        nf = list(gp.get("noisefun", [1, 0, 0]))
        d.noisefun = (C.c_int * 3)(*[int(v) for v in nf])
        y = f64(gp["y"]).ravel() if gp.get("y") is not None else None
        s2 = f64(gp["s2"]).ravel() if gp.get("s2") is not None else None
        d.X, d.y, d.s2, d.hyp = dptr(Xc), dptr(y), dptr(s2), dptr(hyp)
        keep += [Xc, hyp, y, s2]
        return d, (N, D, S, Nhyp)

    def gp_is_resident(self, gp, want_L=False):
        """True when the device holds exactly this posterior (see _gp_identity / _gp_content above)."""
        r = self._gp_key
        if r is None or (want_L and not r.has_L):
            return False
        if r.ref is not None and gp is r.ref:      # a FrozenStruct cannot have changed since it was attached
            return True
        ident, frozen = _gp_identity(gp)
        if frozen and r.frozen and ident == r.ident:
            return True
        if _gp_content(gp, False) != r.content:
            return False
        if want_L and r.content_L is not None and all(p.get("L") is not None for p in gp["post"]) and _gp_content(gp, True) != r.content_L:
            return False
        r.ident, r.frozen = ident, frozen      # same content under new objects: remember them
        return True

    def gp_attach(self, gp, want_L=False):
        """Make ``gp`` (with a posterior computed elsewhere) resident; no-op when the device already holds exactly this posterior."""
        post = gp["post"]
        if self.gp_is_resident(gp, want_L):
            return
        if want_L and any(p.get("L") is None for p in post):
            raise VbmcB200Error(_lib.ESTATE, "vbmc_b200:noL: the variance path needs gp.post(s).L (call gplite_post with want_L=True, "
                                             "or keep the posterior resident)")
        hyp = np.stack([f64(p["hyp"]).ravel() for p in post])
        alpha = np.ascontiguousarray(np.stack([f64(p["alpha"]).ravel() for p in post]))
        keep = []
        d, (N, D, S, _) = self._gp_desc(gp, hyp, keep)
        sW1 = f64([np.asarray(p["sW"]).ravel()[0] for p in post])
        Lchol = np.ascontiguousarray([int(bool(p.get("Lchol", True))) for p in post], dtype=np.int32)
        L = None
        if want_L:
            L = np.ascontiguousarray(np.stack([f64(p["L"]).T for p in post]))  # each column-major
        self._gp_key = None
        _lib.check(self.lib.vbmc_b200_gp_attach(self._h, C.byref(d), dptr(alpha), dptr(sW1),
                                                Lchol.ctypes.data_as(_lib.c_int_p), dptr(L)))
        mult = f64([float(p.get("sn2_mult", 1.0) or 1.0) for p in post])
        _lib.check(self.lib.vbmc_b200_gp_set_sn2_mult(self._h, dptr(mult)))
        self._gp_key = _GpResident(gp, bool(want_L))

    # -- VP ---------------------------------------------------------------------------------
    def vp_set(self, vp):
        D, K = int(vp["D"]), int(vp["K"])
        flags = tuple(int(bool(vp[f])) for f in ("optimize_mu", "optimize_sigma", "optimize_lambda", "optimize_weights"))
        pr = _probe((vp["mu"], vp["sigma"], vp["lambda"], vp["w"], vp.get("eta"), vp.get("delta")))
        probe = None if pr is None else (id(vp), D, K, flags) + pr
        if probe is not None and probe == getattr(self, "_vp_probe", None) and self._vp_key is not None:
            return
        self._vp_probe = None
        mu = np.ascontiguousarray(f64(vp["mu"]).reshape(D, K).T)  # column-major D x K
        sigma = f64(vp["sigma"]).ravel()
        lam = f64(vp["lambda"]).ravel()
        w = f64(vp["w"]).ravel()
        eta = f64(vp["eta"]).ravel() if vp.get("eta") is not None else None
        delta = vp.get("delta")
        if delta is not None and np.size(delta) > 0:
            delta = f64(delta).ravel() * np.ones(D)
        else:
            delta = None
        key = (D, K, flags, _fingerprint(mu, sigma, lam, w, eta, delta))
        if key == self._vp_key:
            self._vp_probe = probe
            return
        d = _lib.VpDesc()
        d.D, d.K = D, K
        d.mu, d.sigma, d.lambda_, d.w, d.eta, d.delta = dptr(mu), dptr(sigma), dptr(lam), dptr(w), dptr(eta), dptr(delta)
        d.optimize_mu, d.optimize_sigma, d.optimize_lambda, d.optimize_weights = flags
        self._vp_key = None
        _lib.check(self.lib.vbmc_b200_vp_set(self._h, C.byref(d)))
        self._vp_key = key
        self._vp_probe = probe

    def thetabnd_set(self, thetabnd):
        if thetabnd is None:
            self._bnd_probe = None
            if self._bnd_key is not None:
                _lib.check(self.lib.vbmc_b200_thetabnd_set(self._h, 0, None, None, 0.0, 0.0, 0.0))
                self._bnd_key = None
            return
        wt, wp = float(thetabnd.get("WeightThreshold", 0.0)), float(thetabnd.get("WeightPenalty", 0.0))
        pr = _probe((thetabnd["lb"], thetabnd["ub"]))
        probe = None if pr is None else (id(thetabnd), float(thetabnd["TolCon"]), wt, wp) + pr
        if probe is not None and probe == getattr(self, "_bnd_probe", None) and self._bnd_key is not None:
            return
        self._bnd_probe = None
        lb, ub = f64(thetabnd["lb"]).ravel(), f64(thetabnd["ub"]).ravel()
        key = (_fingerprint(lb, ub), float(thetabnd["TolCon"]), wt, wp)
        if key == self._bnd_key:
            self._bnd_probe = probe
            return
        self._bnd_key = None
        _lib.check(self.lib.vbmc_b200_thetabnd_set(self._h, lb.size, dptr(lb), dptr(ub), float(thetabnd["TolCon"]), wt, wp))
        self._bnd_key = key
        self._bnd_probe = probe

    # -- draws ------------------------------------------------------------------------------
    def eps_upload(self, epsilon):
        e = f64(epsilon)
        K, half, D = e.shape
        _lib.check(self.lib.vbmc_b200_eps_upload(self._h, D, K, 2 * half, dptr(e)))

    def eps_philox(self, D, K, Ns, seed, stream, readback=False):
        Ns = int(math.ceil(Ns / 2) * 2)
        out = np.empty((K, Ns // 2, D)) if readback else None
        _lib.check(self.lib.vbmc_b200_eps_philox(self._h, D, K, Ns, int(seed), int(stream), dptr(out)))
        return out

    # -- profiling hooks --------------------------------------------------------------------
    def profile_enable(self, on=True):
        _lib.check(self.lib.vbmc_b200_profile_enable(self._h, int(bool(on))))

    def profile_get(self, name):
        ms, n = C.c_double(), C.c_longlong()
        _lib.check(self.lib.vbmc_b200_profile_get(self._h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile_reset(self):
        _lib.check(self.lib.vbmc_b200_profile_reset(self._h))

    def measure_fp64_peak(self):
        v = C.c_double()
        _lib.check(self.lib.vbmc_b200_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def measure_hbm_copy(self):
        v = C.c_double()
        _lib.check(self.lib.vbmc_b200_measure_hbm_copy(self._h, C.byref(v)))
        return v.value

    def flush_l2(self):
        _lib.check(self.lib.vbmc_b200_flush_l2(self._h))


_default = None


def default_context() -> Context:
    """Library-static context, the analogue of the MEX file's mexLock()'d state."""
    global _default
    if _default is None:
        _default = Context(0)
    return _default


# ------------------------------------------------------------------------------------------
# host-side helpers the callers of the hot path use (plain O(DK) NumPy; no GPU work)
# ------------------------------------------------------------------------------------------
def rescale_params(vp, theta=None):
    """misc/rescale_params.m:6-40: assign theta, rescale so that mean(lambda^2) == 1, sum(w) == 1."""
    vp = dict(vp)
    D, K = vp["D"], vp["K"]
    if theta is not None:
        t = f64(theta).ravel()
        i = 0
        if vp["optimize_mu"]:
            vp["mu"] = t[: D * K].reshape(K, D).T.copy()
            i = D * K
        if vp["optimize_sigma"]:
            vp["sigma"] = np.exp(t[i:i + K])
            i += K
        if vp["optimize_lambda"]:
            vp["lambda"] = np.exp(t[i:i + D])
        if vp["optimize_weights"]:
            eta = t[-K:] - np.max(t[-K:])
            vp["w"] = np.exp(eta)
    lam = f64(vp["lambda"]).ravel()
    nl = math.sqrt(float(np.sum(lam * lam)) / D)
    vp["lambda"] = lam / nl
    vp["sigma"] = f64(vp["sigma"]).ravel() * nl
    if vp["optimize_weights"]:
        w = f64(vp["w"]).ravel()
        vp["w"] = w / np.sum(w)
        vp.pop("eta", None)
    vp.pop("mode", None)
    return vp


def get_vptheta(vp):
    """misc/get_vptheta.m:17-21 -> (theta, vp)."""
    vp = rescale_params(vp)
    blocks = []
    if vp["optimize_mu"]:
        blocks.append(f64(vp["mu"]).T.ravel())
    if vp["optimize_sigma"]:
        blocks.append(np.log(vp["sigma"]))
    if vp["optimize_lambda"]:
        blocks.append(np.log(vp["lambda"]))
    if vp["optimize_weights"]:
        blocks.append(np.log(vp["w"]))
    return np.concatenate(blocks), vp


def vpbounds(vp, gp, options, K=None):
    """misc/vpbounds.m:8-52 -> (vp, thetabnd).  options: TolLength, TolWeight, TolConLoss, WeightPenalty."""
    vp = dict(vp)
    K = vp["K"] if K is None else K
    D = vp["D"]
    X = f64(gp["X"])
    lo, hi = X.min(axis=0), X.max(axis=0)
    b = dict(vp.get("bounds") or {"mu_lb": np.full(D, np.inf), "mu_ub": np.full(D, -np.inf),
                                   "lnscale_lb": np.full(D, np.inf), "lnscale_ub": np.full(D, -np.inf)})
    b["mu_lb"] = np.minimum(lo, b["mu_lb"])
    b["mu_ub"] = np.maximum(hi, b["mu_ub"])
    lnrange = np.log(hi - lo)
    b["lnscale_lb"] = np.minimum(b["lnscale_lb"], lnrange + math.log(options["TolLength"]))
    b["lnscale_ub"] = np.maximum(b["lnscale_ub"], lnrange)
    if vp["optimize_weights"]:
        b["eta_lb"], b["eta_ub"] = math.log(0.5 * options["TolWeight"]), 0.0
    vp["bounds"] = b
    lbs, ubs = [], []
    if vp["optimize_mu"]:
        lbs.append(np.tile(b["mu_lb"], K)); ubs.append(np.tile(b["mu_ub"], K))
    if vp["optimize_sigma"] or vp["optimize_lambda"]:
        lbs.append(np.tile(b["lnscale_lb"], K)); ubs.append(np.tile(b["lnscale_ub"], K))
    if vp["optimize_weights"]:
        lbs.append(np.full(K, b["eta_lb"])); ubs.append(np.full(K, b["eta_ub"]))
    tb = {"lb": np.concatenate(lbs), "ub": np.concatenate(ubs), "TolCon": options["TolConLoss"]}
    if vp["optimize_weights"]:
        tb["WeightThreshold"] = max(1.0 / (4 * K), options["TolWeight"])
        tb["WeightPenalty"] = options["WeightPenalty"]
    return vp, tb


# ------------------------------------------------------------------------------------------
# the hot-path functions
# ------------------------------------------------------------------------------------------
def _eps_args(vp, Ns, epsilon, rng):
    Ns2 = int(math.ceil(Ns / 2) * 2)
    if epsilon is not None and not isinstance(epsilon, str):
        e = f64(epsilon)
        want = (vp["K"], Ns2 // 2, vp["D"])
        if e.shape != want:
            raise VbmcB200Error(_lib.EINVAL, f"vbmc_b200:epsilon: epsilon must have shape (K,Ns/2,D)={want}, got {e.shape}")
        return _lib.EPS_HOST, e, 0, 0
    if isinstance(epsilon, str) and epsilon == "resident":
        return _lib.EPS_RESIDENT, None, 0, 0
    if rng is None:
        raise VbmcB200Error(_lib.EINVAL, "vbmc_b200:epsilon: pass epsilon (parity mode), epsilon='resident', or rng=(seed, stream)")
    return _lib.EPS_PHILOX, None, int(rng[0]), int(rng[1])


def negelcbo_vbmc(theta, beta, vp, gp, Ns=0, compute_grad=None, compute_var=None, altent_flag=False,
                  thetabnd=None, entropy_alpha=0, *, epsilon=None, rng=None, nargout=2, ctx=None):
    """[F,dF,G,H,varF,dH,varGss,varG,varH,I_sk,J_sjk] = negelcbo_vbmc(...), misc/negelcbo_vbmc.m:1-164.

    Defaults follow negelcbo_vbmc.m:9-17 (``compute_grad = nargout > 1`` etc.).  ``altent_flag`` and
    ``entropy_alpha`` are accepted and ignored, like the reference (:19).  Returns a tuple of
    ``nargout`` values.
    """
    ctx = ctx or default_context()
    if Ns is None:
        Ns = 0
    if compute_grad is None:
        compute_grad = nargout > 1
    if beta is None or not np.isfinite(beta):
        beta = 0.0
    if compute_var is None:
        compute_var = (beta != 0) or nargout > 4
    separate_K = nargout > 9
    theta = f64(theta).ravel()
    ctx.vp_set(vp)
    ctx.gp_attach(gp, want_L=bool(compute_var))
    ctx.thetabnd_set(thetabnd)
    S, K = len(gp["post"]), int(vp["K"])
    grad, var = bool(compute_grad), int(compute_var)
    frames = ctx.__dict__.setdefault("_frames", {})
    fkey = (theta.size, K, S, grad, bool(var), bool(separate_K))
    fr = frames.get(fkey)
    if fr is None:
        if len(frames) > 64:
            frames.clear()
        fr = frames[fkey] = _CallFrame(theta.size, K, S, grad, bool(var), bool(separate_K))
    a = fr.a
    np.copyto(fr.theta, theta)
    a.beta, a.Ns = float(beta), int(Ns)
    a.compute_grad, a.compute_var, a.separate_K = int(grad), var, int(separate_K)
    a.use_thetabnd = int(thetabnd is not None)
    if Ns > 0:
        mode, e, seed, stream = _eps_args(vp, Ns, epsilon, rng)
    else:
        mode, e, seed, stream = _lib.EPS_RESIDENT, None, 0, 0
    a.eps_mode, a.eps, a.seed, a.stream = mode, dptr(e), seed, stream
    sc = fr.sc
    sc[:] = 0.0
    try:
        _lib.check(ctx.lib.vbmc_b200_negelcbo(ctx.handle, C.byref(a)))
    finally:
        a.eps = None   # do not keep the caller's draws alive through the cached struct
    # results are handed out as fresh arrays: the frame's buffers are overwritten by the next call
    Isk, Jsjk = fr.Isk, fr.Jsjk
    out = (float(sc[0]), fr.dF.copy() if grad else None, float(sc[1]), float(sc[2]), float(sc[3]), fr.dH.copy() if grad else None,
           float(sc[4]), float(sc[5]), float(sc[6]),
           None if Isk is None else Isk.T.copy(), None if Jsjk is None else Jsjk.transpose(2, 1, 0).copy())
    return out[:max(1, nargout)]


def fminadam_negelcbo(x0, beta, vp, gp, Ns, compute_var=0, thetabnd=None, LB=None, UB=None, TolFun=None, MaxIter=None,
                      master_stepsize=None, *, epsilon=None, rng=None, nargout=5, ctx=None):
    """[x,f,xtab,ftab,iter] = fminadam(@(t) negelcbo_vbmc(t,beta,vp,gp,Ns,1,compute_var,0,thetabnd,0), x0, LB, UB,
    TolFun, MaxIter, master_stepsize) — utils/fminadam.m:1-102 driven by the closure of misc/vpoptimize_vbmc.m:71,127,
    with the whole loop on the device (vbmc_b200_fminadam).

    ``master_stepsize``: dict with any of max/min/decay (missing or None -> fminadam.m:11-18 defaults).  ``xtab`` is
    returned as (iter, nvars) — what the reference hands back for a row-vector x0 (fminadam.m:101), the shape
    vpoptimize_vbmc.m:131 indexes (``theta_lst(idx_mid,:)``).  Draws: ``epsilon`` (same draws every iteration, parity
    mode) or ``rng=(seed, stream)`` (iteration i uses Philox stream ``stream+i``: fresh draws per call like ``randn``).
    """
    ctx = ctx or default_context()
    x0 = f64(x0).ravel()
    ctx.vp_set(vp)
    ctx.gp_attach(gp, want_L=bool(compute_var))   # the variance path needs the factors gp.post(s).L on the device
    ctx.thetabnd_set(thetabnd)
    a = _lib.FminadamArgs()
    a.x0, a.nvars = dptr(x0), x0.size
    lb = None if LB is None or np.size(LB) == 0 else f64(LB).ravel() * np.ones(x0.size)
    ub = None if UB is None or np.size(UB) == 0 else f64(UB).ravel() * np.ones(x0.size)
    a.LB, a.UB = dptr(lb), dptr(ub)
    a.TolFun = float("nan") if TolFun is None or np.size(TolFun) == 0 else float(TolFun)
    a.MaxIter = 0 if MaxIter is None or np.size(MaxIter) == 0 else int(MaxIter)
    ms = master_stepsize or {}
    get = lambda k: float("nan") if ms.get(k) is None or np.size(ms.get(k)) == 0 else float(ms[k])
    a.stepsize_max, a.stepsize_min, a.stepsize_decay = get("max"), get("min"), get("decay")
    a.beta = 0.0 if beta is None or not np.isfinite(beta) else float(beta)
    a.Ns, a.compute_var, a.use_thetabnd = int(Ns), int(compute_var or 0), int(thetabnd is not None)
    mode, e, seed, stream = _eps_args(vp, Ns, epsilon, rng)
    a.eps_mode, a.eps, a.seed, a.stream = mode, dptr(e), seed, stream
    maxit = a.MaxIter if a.MaxIter > 0 else 10000
    x = np.zeros(x0.size)
    fval = C.c_double()
    it = C.c_int()
    xtab = np.zeros((maxit, x0.size)) if nargout > 2 else None   # C order (iter, nvars) == column-major nvars x iter
    ftab = np.zeros(maxit) if nargout > 3 else None
    stats = np.zeros(5)
    a.x, a.f, a.xtab, a.ftab, a.iter, a.stats = dptr(x), C.pointer(fval), dptr(xtab), dptr(ftab), C.pointer(it), dptr(stats)
    _lib.check(ctx.lib.vbmc_b200_fminadam(ctx.handle, C.byref(a)))
    n = it.value
    out = (x, fval.value, None if xtab is None else xtab[:n].copy(), None if ftab is None else ftab[:n].copy(), n)
    fminadam_negelcbo.last_stats = dict(zip(("stop", "dx", "slope", "slope_err", "slope_err_max"), stats))
    return out[:max(1, nargout)]


def entmc_vbmc(vp, Ns=10, grad_flags=None, jacobian_flag=True, *, epsilon=None, rng=None, nargout=2, ctx=None):
    """[H,dH] = entmc_vbmc(vp,Ns,grad_flags,jacobian_flag), ent/entmc_vbmc.m:1-125."""
    ctx = ctx or default_context()
    if nargout < 2:
        grad_flags = False
    elif grad_flags is None:
        grad_flags = True
    if np.isscalar(grad_flags):
        grad_flags = [bool(grad_flags)] * 4
    gf = (C.c_int * 4)(*[int(bool(g)) for g in grad_flags])
    ctx.vp_set(vp)
    mode, e, seed, stream = _eps_args(vp, Ns, epsilon, rng)
    D, K = int(vp["D"]), int(vp["K"])
    n = D * K * gf[0] + K * gf[1] + D * gf[2] + K * gf[3]
    H = C.c_double()
    dH = np.zeros(n) if nargout > 1 else None
    _lib.check(ctx.lib.vbmc_b200_entmc(ctx.handle, int(Ns), gf, int(bool(jacobian_flag)), mode, dptr(e), seed, stream,
                                       C.byref(H), dptr(dH)))
    return (H.value, dH)[:max(1, nargout)]


def entlb_vbmc(vp, grad_flags=None, jacobian_flag=True, *, nargout=2, ctx=None):
    """[H,dH] = entlb_vbmc(vp,grad_flags,jacobian_flag), ent/entlb_vbmc.m:1-147: the deterministic entropy lower bound
    negelcbo_vbmc uses when Ns == 0 (negelcbo_vbmc.m:102-109)."""
    ctx = ctx or default_context()
    if nargout < 2:
        grad_flags = False
    elif grad_flags is None:
        grad_flags = True
    if np.isscalar(grad_flags):
        grad_flags = [bool(grad_flags)] * 4
    gf = (C.c_int * 4)(*[int(bool(g)) for g in grad_flags])
    ctx.vp_set(vp)
    D, K = int(vp["D"]), int(vp["K"])
    n = D * K * gf[0] + K * gf[1] + D * gf[2] + K * gf[3]
    H = C.c_double()
    dH = np.zeros(n) if nargout > 1 else None
    _lib.check(ctx.lib.vbmc_b200_entlb(ctx.handle, gf, int(bool(jacobian_flag)), C.byref(H), dptr(dH)))
    return (H.value, dH)[:max(1, nargout)]


def gplogjoint(vp, gp, grad_flags=None, avg_flag=True, jacobian_flag=True, compute_var=None, separate_K=None, *,
               nargout=2, ctx=None):
    """[F,dF,varF,dvarF,varss,I_sk,J_sjk] = gplogjoint(...), misc/gplogjoint.m:1-413."""
    ctx = ctx or default_context()
    if separate_K is None:
        separate_K = nargout > 5
    if compute_var is None:
        compute_var = nargout > 2
    if nargout < 2:
        grad_flags = False
    elif grad_flags is None:
        grad_flags = True
    if np.isscalar(grad_flags):
        grad_flags = [bool(grad_flags)] * 4
    gf = (C.c_int * 4)(*[int(bool(g)) for g in grad_flags])
    ctx.vp_set(vp)
    ctx.gp_attach(gp, want_L=bool(compute_var))
    D, K, S = int(vp["D"]), int(vp["K"]), len(gp["post"])
    n = D * K * gf[0] + K * gf[1] + D * gf[2] + K * gf[3]
    F, varss = C.c_double(), C.c_double()
    dF = np.zeros(n) if nargout > 1 else None
    varF = C.c_double() if compute_var else None
    compute_vargrad = nargout > 3 and compute_var and any(gf)
    dvarF = np.zeros(n) if compute_vargrad else None
    Isk = np.zeros((K, S)) if separate_K else None
    Jsjk = np.zeros((K, K, S)) if (separate_K and compute_var) else None
    _lib.check(ctx.lib.vbmc_b200_gplogjoint(ctx.handle, gf, int(bool(avg_flag)), int(bool(jacobian_flag)), int(compute_var),
                                            C.byref(F), dptr(dF), None if varF is None else C.byref(varF), dptr(dvarF),
                                            C.byref(varss), dptr(Isk), dptr(Jsjk)))
    out = (F.value, dF, None if varF is None else varF.value, dvarF, varss.value,
           None if Isk is None else Isk.T.copy(), None if Jsjk is None else Jsjk.transpose(2, 1, 0).copy())
    return out[:max(1, nargout)]


def _noise_count(noisefun):
    return int(noisefun[0] == 1) + int(noisefun[1] == 2) + 2 * int(noisefun[2] == 1)


def _mean_count(D, meanfun):
    return {0: 0, 1: 1, 4: 1 + 2 * D}.get(meanfun)


def gplite_post(hyp, X, y, covfun=None, meanfun=None, noisefun=None, s2=None, *, ctx=None, want_L=True):
    """gp = gplite_post(hyp,X,y,covfun,meanfun,noisefun,s2), gplite/gplite_post.m:94-172 (full refit).

    The S Cholesky factorisations run on the GPU and the posterior stays resident (it becomes the
    attached GP of ``ctx``); the returned dict mirrors the MATLAB struct.  ``want_L=False`` skips the
    N x N x S device->host copy of the factors (gp.post(s).L is then None).
    """
    ctx = ctx or default_context()
    X = np.array(X, dtype=np.float64, order="C")        # the struct owns frozen copies (MATLAB value semantics): an in-place edit of
    y = np.array(y, dtype=np.float64).ravel()           # the caller's arrays cannot silently diverge from the device posterior
    N, D = X.shape
    hyp = f64(hyp)
    if hyp.ndim == 1:
        hyp = hyp[:, None]
    Nhyp, S = hyp.shape
    covfun = 1 if covfun is None else covfun
    meanfun = 1 if meanfun is None else meanfun
    if noisefun is None:
        noisefun = [1, 0, 0] if s2 is None else [1, 1, 0]
    gp = {"X": X, "y": y, "s2": None if s2 is None else np.array(s2, dtype=np.float64).ravel(), "covfun": covfun, "meanfun": meanfun,
          "noisefun": list(noisefun), "Ncov": D + 1, "Nnoise": _noise_count(noisefun), "Nmean": _mean_count(D, meanfun),
          "meanfun_extras": None, "intmeanfun": 0}
    keep = []
    d, _ = ctx._gp_desc(gp, np.ascontiguousarray(hyp.T), keep)
    alpha = np.zeros((S, N))
    L = np.zeros((S, N, N)) if want_L else None
    sW1, mult = np.zeros(S), np.zeros(S)
    Lchol = np.zeros(S, dtype=np.int32)
    _lib.check(ctx.lib.vbmc_b200_gp_post(ctx.handle, C.byref(d), dptr(alpha), dptr(L), dptr(sW1), dptr(mult),
                                         Lchol.ctypes.data_as(_lib.c_int_p)))
    # per-sample rows are views of the (frozen) result arrays; sW is a constant vector (gplite_core.m:281): a read-only broadcast
    hypT = np.ascontiguousarray(hyp.T)
    gp["post"] = [{"hyp": hypT[s], "alpha": alpha[s], "sW": np.broadcast_to(sW1[s:s + 1], (N,)),
                   "L": None if L is None else L[s].T.copy(), "sn2_mult": float(mult[s]), "Lchol": bool(Lchol[s])}
                  for s in range(S)]
    gp = _freeze_gp(gp)
    ctx._gp_key = _GpResident(gp, True)   # the factors stay on the device whether or not they were copied out
    return gp


def gplite_post_update1(gp, xstar, ystar, s2star=None, *, ctx=None):
    """gp = gplite_post(gp,xstar,ystar,[],[],[],s2star,1) — rank-one update (gplite/gplite_post.m:50-92,173-251) of the
    posterior resident on the GPU.  With ``s2star`` (heteroskedastic noise) the reference itself performs the standard
    update with the enlarged training set (:78-91): so does this mirror (full GPU refit)."""
    ctx = ctx or default_context()
    xs = f64(xstar)
    if xs.ndim > 1 and xs.shape[0] > 1:
        raise VbmcB200Error(_lib.EREFERENCE, "gplite_post:NotRankOne: GPLITE_POST with this input format only supports rank-one updates.")
    if gp is None:
        raise VbmcB200Error(_lib.EREFERENCE, "gplite_post:NoGP: GPLITE_POST can perform rank-one update only with an existing GP struct.")
    xs = xs.ravel()
    ystar = float(np.ravel(ystar)[0])
    X, y = f64(gp["X"]), f64(gp["y"]).ravel()
    N, D = X.shape
    post = gp["post"]
    S = len(post)
    if s2star is not None and np.size(s2star) > 0 or gp.get("intmeanfun", 0):
        hyp = np.stack([f64(p["hyp"]).ravel() for p in post], axis=1)
        s2 = None if s2star is None else np.append(f64(gp["s2"]).ravel(), f64(s2star).ravel())
        return gplite_post(hyp, np.vstack([X, xs[None, :]]), np.append(y, ystar), gp["covfun"], gp["meanfun"], gp["noisefun"], s2,
                           ctx=ctx, want_L=any(p.get("L") is not None for p in post))
    have_L = all(p.get("L") is not None for p in post)
    if not ctx.gp_is_resident(gp, want_L=True):
        ctx.gp_attach(gp, want_L=True)   # raises vbmc_b200:noL when the factors are neither resident nor in the struct
    alpha = np.zeros((S, N + 1))
    Lcol = np.zeros((S, N + 1))
    sWn = np.zeros(S)
    _lib.check(ctx.lib.vbmc_b200_gp_post_update1(ctx.handle, dptr(xs), ystar, dptr(alpha), dptr(Lcol), dptr(sWn)))
    new = dict(gp)
    Xn = np.empty((N + 1, D))
    Xn[:N] = X
    Xn[N] = xs
    yn = np.empty(N + 1)
    yn[:N] = y
    yn[N] = ystar
    new["X"], new["y"] = Xn, yn
    new["post"] = []
    sW_all = np.empty((S, N + 1))       # per-sample rows of the (frozen) result arrays are handed out as views
    for s, p in enumerate(post):
        sW_all[s, :N] = p["sW"]
    sW_all[:, N] = sWn
    for s, p in enumerate(post):
        q = dict(p)
        q["alpha"] = alpha[s]
        q["sW"] = sW_all[s]
        if have_L and not p.get("Lchol", True):
            # low noise (:234-238): every entry of L = -inv(K + diag) changes; the device hands the new matrix over (symmetric)
            Ln = np.zeros((N + 1, N + 1))
            _lib.check(ctx.lib.vbmc_b200_gp_get_factor(ctx.handle, s, dptr(Ln)))
            q["L"] = Ln.T.copy()   # column-major on the wire
        elif have_L:
            Ln = np.zeros((N + 1, N + 1))
            Ln[:N, :N] = p["L"]
            Ln[:, N] = Lcol[s]
            q["L"] = Ln
        else:
            q["L"] = None
        new["post"].append(q)
    new = _freeze_gp(new)
    ctx._gp_key = _GpResident(new, True)
    return new


def gplite_nlZ_batch(hyp, gp, hprior=None, *, ctx=None):
    """nlZ[s] = gplite_nlZ(hyp[:, s], gp, hprior) for all columns of ``hyp`` in one batched factorisation (value only):
    the evaluations gplite_train.m makes one at a time (fminfill design :200-204, slice sampler :318-330)."""
    ctx = ctx or default_context()
    hyp = f64(hyp)
    if hyp.ndim == 1:
        hyp = hyp[:, None]
    Nhyp, Ns = hyp.shape
    D = gp["X"].shape[1]
    Ncov, Nnoise, Nmean = D + 1, _noise_count(gp["noisefun"]), _mean_count(D, gp["meanfun"])
    if Nmean is None or Nhyp != Ncov + Nnoise + Nmean:
        raise VbmcB200Error(_lib.EREFERENCE, "gplite_nlZ:dimmismatch: Number of hyperparameters mismatched with dimension of training inputs.")
    keep = []
    d, _ = ctx._gp_desc(gp, np.ascontiguousarray(hyp.T), keep)
    hp = None
    if hprior is not None:
        hp = _lib.HPrior()
        mu, sg = f64(hprior["mu"]).ravel(), f64(hprior["sigma"]).ravel()
        df = hprior.get("df")
        df = None if df is None or np.size(df) == 0 else f64(df).ravel()
        hp.mu, hp.sigma, hp.df = dptr(mu), dptr(sg), dptr(df)
        keep += [mu, sg, df]
    out = np.zeros(Ns)
    _lib.check(ctx.lib.vbmc_b200_gp_nlz_batch(ctx.handle, C.byref(d), None if hp is None else C.byref(hp), dptr(out)))
    ctx._gp_key = None  # the GP work buffers were reused
    return out


def gplite_pred(gp, Xstar, ystar=None, s2star=None, ssflag=False, nowarpflag=False, *, nargout=2, ctx=None):
    """[ymu,ys2,fmu,fs2,lp] = gplite_pred(gp,Xstar,ystar,s2star,ssflag,nowarpflag), gplite/gplite_pred.m:1-163.

    Arrays come back as (Nstar, S) when ``ssflag`` (or one sample), else (Nstar,); ``lp`` is (Nstar, S) or None.
    """
    ctx = ctx or default_context()
    Xs = f64(Xstar)
    if Xs.ndim == 1:
        Xs = Xs[None, :]
    Nstar = Xs.shape[0]
    ystar = None if ystar is None or np.size(ystar) == 0 else f64(ystar).ravel()
    s2star = None if s2star is None or np.size(s2star) == 0 else f64(s2star).ravel()
    if ystar is not None and ystar.size != Nstar:
        raise VbmcB200Error(_lib.EREFERENCE, "gplite_pred:ydimmismatch: YSTAR should be empty or a column vector of NSTAR observations.")
    if s2star is not None and s2star.size != Nstar:
        raise VbmcB200Error(_lib.EREFERENCE, "gplite_pred:s2dimmismatch: S2STAR should be empty or a column vector of NSTAR estimated variances.")
    if gp.get("intmeanfun") or gp.get("outwarpfun"):
        raise VbmcB200Error(_lib.EUNSUPPORTED, "vbmc_b200:OutOfScope: integrated mean functions / output warping (not VBMC defaults)")
    want_var = nargout > 1
    ctx.gp_attach(gp, want_L=want_var)
    S = len(gp["post"])
    sep = bool(ssflag) or S == 1
    ncol = S if sep else 1
    Xc = np.ascontiguousarray(Xs.T)   # column-major Nstar x D
    mk = lambda n: np.zeros((n, Nstar))
    ymu, fmu = mk(ncol), mk(ncol)
    ys2 = mk(ncol) if want_var else None
    fs2 = mk(ncol) if want_var else None
    lp = mk(S) if (ystar is not None and nargout > 4) else None
    _lib.check(ctx.lib.vbmc_b200_gp_pred(ctx.handle, Nstar, dptr(Xc), dptr(ystar), dptr(s2star), int(bool(ssflag)), int(want_var),
                                         dptr(ymu), dptr(ys2), dptr(fmu), dptr(fs2), dptr(lp)))
    shape = (lambda a: None if a is None else (a.T.copy() if sep else a[0].copy()))
    out = (shape(ymu), shape(ys2), shape(fmu), shape(fs2), None if lp is None else lp.T.copy())
    return out[:max(1, nargout)]


def gplite_nlZ(hyp, gp, hprior=None, *, nargout=1, ctx=None):
    """[nlZ,dnlZ] = gplite_nlZ(hyp,gp,hprior), gplite/gplite_nlZ.m:27-66."""
    ctx = ctx or default_context()
    hyp = f64(hyp)
    if hyp.ndim == 1:
        hyp = hyp[:, None]
    Nhyp, Ns = hyp.shape
    D = gp["X"].shape[1]
    Ncov, Nnoise, Nmean = D + 1, _noise_count(gp["noisefun"]), _mean_count(D, gp["meanfun"])
    if Nmean is None or Nhyp != Ncov + Nnoise + Nmean:
        raise VbmcB200Error(_lib.EREFERENCE, "gplite_nlZ:dimmismatch: Number of hyperparameters mismatched with dimension of training inputs.")
    if nargout > 1 and Ns > 1:
        raise VbmcB200Error(_lib.EREFERENCE, "gplite_nlZ:NoSampling: Computation of the log marginal likelihood is available only for one-sample hyperparameter inputs.")
    if Ns > 1:
        return (gplite_nlZ_batch(hyp, gp, hprior, ctx=ctx),)
    keep = []
    d, _ = ctx._gp_desc(gp, np.ascontiguousarray(hyp[:, :1].T), keep)
    hp = None
    if hprior is not None:
        hp = _lib.HPrior()
        mu, sg = f64(hprior["mu"]).ravel(), f64(hprior["sigma"]).ravel()
        df = hprior.get("df")
        df = None if df is None or np.size(df) == 0 else f64(df).ravel()
        hp.mu, hp.sigma, hp.df = dptr(mu), dptr(sg), dptr(df)
        keep += [mu, sg, df]
    nlZ = C.c_double()
    dnlZ = np.zeros(Nhyp) if nargout > 1 else None
    _lib.check(ctx.lib.vbmc_b200_gp_nlz(ctx.handle, C.byref(d), None if hp is None else C.byref(hp), C.byref(nlZ), dptr(dnlZ)))
    ctx._gp_key = None  # gp_nlz reuses the GP work buffers
    return (nlZ.value, dnlZ)[:max(1, nargout)]
