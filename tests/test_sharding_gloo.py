"""N>1 host logic on CPU ranks (gloo, world_size 2): the shard plan reproduces the unsharded result.
Each rank evaluates the ORACLE on its shard of the MC pair axis / hyper-parameter-sample axis, the
partial sums are all-reduced once, and the combination must equal the unsharded oracle."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_matches_c_abi():
    import _sharding as sharding
    from vbmc_b200 import _lib
    lib = _lib.load()
    for total in (0, 1, 7, 20, 16384, 65536):
        for n in (1, 2, 3, 4, 8):
            cover = []
            for r in range(n):
                b, e = C.c_int(), C.c_int()
                assert lib.vbmc_b200_shard_range(total, n, r, C.byref(b), C.byref(e)) == 0
                assert (b.value, e.value) == sharding.shard_range(total, n, r)
                cover += list(range(b.value, e.value))
            assert cover == list(range(total))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import vbmc_oracle as orc
    import _sharding as sharding
    from vbmc_b200 import workloads
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = dict(D=3, N=40, K=4, S=5, Ns=64, target="rosenbrock", noisy=False, log_sn=np.log(0.1))
    w = workloads.build(cfg, orc.gplite_post, seeds=(11, 12, 13, 14))
    vp, gp, eps, Ns = w["vp"], w["gp"], w["epsilon"], cfg["Ns"]
    D, K, S = cfg["D"], cfg["K"], cfg["S"]
    # ---- entropy: shard the pair axis.  H and dH (pre-normalisation sums) are linear in the pairs ----
    pb, pe = sharding.shard_range(Ns // 2, world, rank)
    n_loc = 2 * (pe - pb)
    if n_loc > 0:
        H_loc, dH_loc = orc.entmc_vbmc(vp, n_loc, True, True, epsilon=eps[:, pb:pe, :])
        part_H = np.concatenate([[H_loc], dH_loc]) * (n_loc / Ns)   # means -> weighted by shard size
    else:
        part_H = np.zeros(1 + D * K + K + D + K)
    # ---- log joint: shard the hyper-parameter samples.  G, dG are means over s ----
    sb, se = sharding.shard_range(S, world, rank)
    part_G = np.zeros(1 + D * K + K + D + K)
    if se > sb:
        gp_loc = dict(gp, post=gp["post"][sb:se])
        G_loc, dG_loc = orc.gplogjoint(vp, gp_loc, True, True, True, 0, nargout=2)[:2]
        part_G = np.concatenate([[G_loc], dG_loc]) * ((se - sb) / S)
    buf = torch.from_numpy(np.concatenate([part_H, part_G]))
    dist.all_reduce(buf)                                           # the ONE exchange step
    tot = buf.numpy()
    nh = part_H.size
    H, dH, G, dG = tot[0], tot[1:nh], tot[nh], tot[nh + 1:]
    Ho, dHo = orc.entmc_vbmc(vp, Ns, True, True, epsilon=eps)
    Go, dGo = orc.gplogjoint(vp, gp, True, True, True, 0, nargout=2)[:2]
    err = max(abs(H - Ho) / abs(Ho), np.max(np.abs(dH - dHo)) / np.max(np.abs(dHo)),
              abs(G - Go) / abs(Go), np.max(np.abs(dG - dGo)) / np.max(np.abs(dGo)))
    q.put((rank, float(err)))
    dist.destroy_process_group()


def test_two_rank_shard_equals_unsharded():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-12, (rank, err)
