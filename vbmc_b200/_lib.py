"""ctypes binding of libvbmc_b200.so (the C ABI declared in include/vbmc_b200.h).

This is the Python counterpart of the MEX gateway shown in INTEGRATION.md: it marshals host
NumPy buffers (MATLAB column-major FP64) into the plain-C entry points.  There is no CPU
fallback: if the shared library is missing, or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libvbmc_b200.so"

OK, EINVAL, ECUDA, ENODEV, ESTATE, EUNSUPPORTED, EREFERENCE, ENCCL = range(8)
EPS_HOST, EPS_RESIDENT, EPS_PHILOX = 0, 1, 2
UNIQUE_ID_BYTES = 128

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class VbmcB200Error(RuntimeError):
    """Raised for any non-zero status.  ``identifier`` is the MATLAB error id the reference
    would raise (e.g. 'negelcbo_vbmc:vargrad') or a 'vbmc_b200:*' id for library conditions."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code
        head = message.split(" ", 1)[0]
        self.identifier = head[:-1] if head.endswith(":") and head.count(":") >= 2 else ""


class GpDesc(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("D", C.c_int), ("S", C.c_int), ("Nhyp", C.c_int),
        ("covfun", C.c_int), ("meanfun", C.c_int), ("noisefun", C.c_int * 3),
        ("X", c_double_p), ("y", c_double_p), ("s2", c_double_p), ("hyp", c_double_p),
    ]


class HPrior(C.Structure):
    _fields_ = [("mu", c_double_p), ("sigma", c_double_p), ("df", c_double_p)]


class VpDesc(C.Structure):
    _fields_ = [
        ("D", C.c_int), ("K", C.c_int),
        ("mu", c_double_p), ("sigma", c_double_p), ("lambda_", c_double_p), ("w", c_double_p),
        ("eta", c_double_p), ("delta", c_double_p),
        ("optimize_mu", C.c_int), ("optimize_sigma", C.c_int), ("optimize_lambda", C.c_int),
        ("optimize_weights", C.c_int),
    ]


class NegelcboArgs(C.Structure):
    _fields_ = [
        ("theta", c_double_p), ("ntheta", C.c_int), ("beta", C.c_double), ("Ns", C.c_int),
        ("compute_grad", C.c_int), ("compute_var", C.c_int), ("separate_K", C.c_int),
        ("use_thetabnd", C.c_int), ("eps_mode", C.c_int), ("eps", c_double_p),
        ("seed", C.c_uint64), ("stream", C.c_uint64),
        ("F", c_double_p), ("dF", c_double_p), ("G", c_double_p), ("H", c_double_p),
        ("varF", c_double_p), ("dH", c_double_p), ("varGss", c_double_p), ("varG", c_double_p),
        ("varH", c_double_p), ("I_sk", c_double_p), ("J_sjk", c_double_p),
    ]


class FminadamArgs(C.Structure):
    _fields_ = [
        ("x0", c_double_p), ("nvars", C.c_int), ("LB", c_double_p), ("UB", c_double_p),
        ("TolFun", C.c_double), ("MaxIter", C.c_int),
        ("stepsize_max", C.c_double), ("stepsize_min", C.c_double), ("stepsize_decay", C.c_double),
        ("beta", C.c_double), ("Ns", C.c_int), ("compute_var", C.c_int), ("use_thetabnd", C.c_int),
        ("eps_mode", C.c_int), ("eps", c_double_p), ("seed", C.c_uint64), ("stream", C.c_uint64),
        ("x", c_double_p), ("f", c_double_p), ("xtab", c_double_p), ("ftab", c_double_p), ("iter", c_int_p),
        ("stats", c_double_p),
    ]


# every symbol include/vbmc_b200.h declares (checked by tests/test_abi.py without a GPU)
EXPORTS = [
    "vbmc_b200_version", "vbmc_b200_last_error", "vbmc_b200_create", "vbmc_b200_destroy", "vbmc_b200_sync",
    "vbmc_b200_launch_count", "vbmc_b200_set_precision", "vbmc_b200_comm_unique_id", "vbmc_b200_comm_init", "vbmc_b200_comm_info",
    "vbmc_b200_gp_attach", "vbmc_b200_gp_post", "vbmc_b200_gp_nlz", "vbmc_b200_vp_set", "vbmc_b200_thetabnd_set",
    "vbmc_b200_eps_upload", "vbmc_b200_eps_philox", "vbmc_b200_negelcbo", "vbmc_b200_entmc", "vbmc_b200_gplogjoint",
    "vbmc_b200_negelcbo_resident_loop", "vbmc_b200_profile_enable", "vbmc_b200_profile_get",
    "vbmc_b200_profile_reset", "vbmc_b200_measure_fp64_peak", "vbmc_b200_measure_hbm_copy", "vbmc_b200_flush_l2",
    "vbmc_b200_philox_raw", "vbmc_b200_shard_range", "vbmc_b200_fminadam", "vbmc_b200_entmc_prune", "vbmc_b200_entmc_prune_stats", "vbmc_b200_entmc_balance", "vbmc_b200_entmc_plan_get", "vbmc_b200_gp_pred", "vbmc_b200_gp_post_update1", "vbmc_b200_gp_get_factor", "vbmc_b200_gp_nlz_batch", "vbmc_b200_gp_set_sn2_mult",
    "vbmc_b200_comm_p2p", "vbmc_b200_shared", "vbmc_b200_shared_release", "vbmc_b200_gp_tag_set", "vbmc_b200_gp_tag_get", "vbmc_b200_gp_shape", "vbmc_b200_entlb",
]

_lib = None


def load():
    """dlopen the in-tree library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("VBMC_B200_LIB", LIB_PATH))
    if not path.exists():
        raise VbmcB200Error(ENODEV, f"vbmc_b200:nolib: {path} not found; run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(there is no CPU fallback)")
    lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
    lib.vbmc_b200_last_error.restype = C.c_char_p
    lib.vbmc_b200_version.restype = C.c_int
    vp = C.c_void_p
    lib.vbmc_b200_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.vbmc_b200_destroy.argtypes = [vp]
    lib.vbmc_b200_sync.argtypes = [vp]
    lib.vbmc_b200_launch_count.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.vbmc_b200_set_precision.argtypes = [vp, C.c_int]
    lib.vbmc_b200_comm_unique_id.argtypes = [C.c_void_p]
    lib.vbmc_b200_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_void_p]
    lib.vbmc_b200_comm_info.argtypes = [vp, c_int_p, c_int_p]
    lib.vbmc_b200_comm_p2p.argtypes = [vp, c_int_p]
    lib.vbmc_b200_shard_range.argtypes = [C.c_int, C.c_int, C.c_int, c_int_p, c_int_p]
    lib.vbmc_b200_gp_attach.argtypes = [vp, C.POINTER(GpDesc), c_double_p, c_double_p, c_int_p, c_double_p]
    lib.vbmc_b200_gp_post.argtypes = [vp, C.POINTER(GpDesc), c_double_p, c_double_p, c_double_p, c_double_p, c_int_p]
    lib.vbmc_b200_gp_nlz.argtypes = [vp, C.POINTER(GpDesc), C.POINTER(HPrior), c_double_p, c_double_p]
    lib.vbmc_b200_gp_nlz_batch.argtypes = [vp, C.POINTER(GpDesc), C.POINTER(HPrior), c_double_p]
    lib.vbmc_b200_gp_post_update1.argtypes = [vp, c_double_p, C.c_double, c_double_p, c_double_p, c_double_p]
    lib.vbmc_b200_gp_get_factor.argtypes = [vp, C.c_int, c_double_p]
    lib.vbmc_b200_vp_set.argtypes = [vp, C.POINTER(VpDesc)]
    lib.vbmc_b200_thetabnd_set.argtypes = [vp, C.c_int, c_double_p, c_double_p, C.c_double, C.c_double, C.c_double]
    lib.vbmc_b200_eps_upload.argtypes = [vp, C.c_int, C.c_int, C.c_int, c_double_p]
    lib.vbmc_b200_eps_philox.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, c_double_p]
    lib.vbmc_b200_negelcbo.argtypes = [vp, C.POINTER(NegelcboArgs)]
    lib.vbmc_b200_negelcbo_resident_loop.argtypes = [vp, C.POINTER(NegelcboArgs), C.c_int, C.POINTER(C.c_float)]
    lib.vbmc_b200_entmc_prune.argtypes = [vp, C.c_double]
    lib.vbmc_b200_entmc_prune_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.vbmc_b200_entmc_balance.argtypes = [vp, C.c_int, C.c_int]
    lib.vbmc_b200_entmc_plan_get.argtypes = [vp, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.vbmc_b200_gp_pred.argtypes = [vp, C.c_int, c_double_p, c_double_p, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p,
                                      c_double_p, c_double_p, c_double_p]
    lib.vbmc_b200_gp_set_sn2_mult.argtypes = [vp, c_double_p]
    lib.vbmc_b200_fminadam.argtypes = [vp, C.POINTER(FminadamArgs)]
    lib.vbmc_b200_entmc.argtypes = [vp, C.c_int, c_int_p, C.c_int, C.c_int, c_double_p, C.c_uint64, C.c_uint64,
                                    c_double_p, c_double_p]
    lib.vbmc_b200_entlb.argtypes = [vp, c_int_p, C.c_int, c_double_p, c_double_p]
    lib.vbmc_b200_gplogjoint.argtypes = [vp, c_int_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                         c_double_p, c_double_p, c_double_p, c_double_p]
    lib.vbmc_b200_profile_enable.argtypes = [vp, C.c_int]
    lib.vbmc_b200_profile_get.argtypes = [vp, C.c_char_p, c_double_p, C.POINTER(C.c_longlong)]
    lib.vbmc_b200_profile_reset.argtypes = [vp]
    lib.vbmc_b200_measure_fp64_peak.argtypes = [vp, c_double_p]
    lib.vbmc_b200_measure_hbm_copy.argtypes = [vp, c_double_p]
    lib.vbmc_b200_flush_l2.argtypes = [vp]
    lib.vbmc_b200_philox_raw.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    _lib = lib
    return lib


def check(rc: int):
    if rc != OK:
        msg = load().vbmc_b200_last_error().decode("utf-8", "replace")
        raise VbmcB200Error(rc, msg)


def dptr(a):
    """double* of a C-contiguous float64 array (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))
