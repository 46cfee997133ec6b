"""GPU (-m gpu), >= 2 GPUs: torchrun-launched check that R ranks + the NCCL gradient all-reduce reproduce the 1-rank gradients on
the concatenated batch, for sample sharding (level 1, DDP mean) and point sharding (level 2, n_norm + SUM).
Skipped on a single-GPU box; the run on 2 / 8 GPUs of this round is kept under profiles/ (r02_multirank_*.txt)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-6), ("f16x3", 2e-5)])
def test_ranks_reproduce_single_rank_gradients(mode, tol):
    """fp32 mode: <= 1e-6 (reduction-order noise only).  f16x3: every rank's tiles get their own power-of-two scales, so the
    sharded and the unsharded run round differently; bound = the mode's own noise floor (2e-5), far below its oracle tolerance."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 8 else 8
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multirank_child.py"), mode, str(tol)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("MULTIRANK")]
    print("\n".join(line))
    assert r.returncode == 0, (line, r.stderr[-3000:])
