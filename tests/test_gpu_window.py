"""Two-rank proof over the IPC exchange window (NVLink peer stores from inside the row-hash kernel,
flag barrier instead of an all-gather): every rank's proof must equal the single-GPU proof byte for
byte.  Needs two GPUs on the box; skipped otherwise (the single-GPU simulation of the sharded path
is tests/test_gpu_sharded.py, the host logic runs under gloo in tests/test_sharded_cpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_window_sharded_proof_two_ranks():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, LOGN="13")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_window_debug.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("equal to the single-GPU proof: True") == 6
