"""TEST HELPER (not product code): multi-GPU decomposition of one negelcbo_vbmc step — host-side mirror of the
plan the library uses on the device, so that the shard logic can be tested on CPU ranks (gloo).

Every output of a step is a sum over independent units:
  entmc_vbmc  : sum over (source component j, antithetic pair p)  -> the PAIR axis p is sharded (all K
                components on every rank, partners +eps/-eps stay together);
  gplogjoint  : sum over hyper-parameter samples s, then /S       -> the SAMPLE axis s is sharded.
Each rank fills a partial vector R = [Hs(K) | M(K*D) | E(K*D) | Wc(K) | I_sk(S*K) | Gmu(K*D) | Gsig(K) | Glam(D)]
(Wc[l] = sum_j w_j W_jl and Glam are contracted with the weights before the exchange, csrc/common.cuh RLayout);
ONE all-reduce (SUM) makes it identical on all ranks; the O(DK) epilogue (Jacobians, penalties) is replicated.
"""
from __future__ import annotations


def shard_range(total: int, nranks: int, rank: int):
    """Contiguous balanced split: the first (total % nranks) ranks get one extra unit (== vbmc_b200_shard_range)."""
    base, rem = divmod(total, nranks)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def r_layout(D: int, K: int, S: int):
    """Offsets of the all-reduced partial vector (RLayout in csrc/common.cuh)."""
    o = {}
    n = 0
    for name, size in (("Hs", K), ("M", K * D), ("E", K * D), ("Wc", K), ("I", S * K), ("Gmu", K * D), ("Gsig", K), ("Glam", D)):
        o[name] = (n, n + size)
        n += size
    o["total"] = n
    return o
