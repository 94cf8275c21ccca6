"""The multi-GPU parity script (tests/mgpu_check.py) under pytest: runs on every box with >= 2 GPUs (self-skips otherwise), one
rank per GPU through torch.distributed.run, all GPUs of the box (max 8).  It checks: identical F, dF on all ranks, agreement with
the binary128 evaluation to 1e-10, beta != 0 with the variance gradient, streaming generator-mode calls (graph replay +
ahead-of-time draws) equal to parity mode bit for bit, and the device fminadam loop identical on all ranks."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_multi_gpu_step_matches_truth_on_every_rank():
    n = min(_ngpu(), 8)
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, this box has {n}")
    env = dict(os.environ, PYTHONPATH=ROOT, VBMC_MGPU_QUICK=os.environ.get("VBMC_MGPU_QUICK", "0"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29917", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    assert r.returncode == 0, tail
    assert f"[mgpu_check] OK world={n}" in r.stdout, tail
