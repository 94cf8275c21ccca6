"""vbmc_b200 — B200 (sm_100a) implementation of the VBMC variational-optimisation hot path.

Public surface = the reference's MATLAB function names for this path (see vbmc_b200/api.py);
the numerics live in vbmc_b200/lib/libvbmc_b200.so (C ABI: include/vbmc_b200.h).
"""
from .api import (Context, VbmcB200Error, default_context, entlb_vbmc, entmc_vbmc, fminadam_negelcbo, get_vptheta, gplite_nlZ, gplite_nlZ_batch, gplite_post, gplite_post_update1, gplite_pred,
                  gplogjoint, negelcbo_vbmc, rescale_params, vpbounds)

__all__ = ["Context", "VbmcB200Error", "default_context", "entlb_vbmc", "entmc_vbmc", "fminadam_negelcbo", "get_vptheta", "gplite_nlZ", "gplite_nlZ_batch", "gplite_post", "gplite_post_update1", "gplite_pred",
           "gplogjoint", "negelcbo_vbmc", "rescale_params", "vpbounds"]
__version__ = "0.1.0"
