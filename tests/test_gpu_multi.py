"""2-rank run of the sharded path on one box (needs two GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_reproduce_the_oracle(built):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "multi_rank_check.py"), "tiny",
           "adversarial", "clip50", "c1_40x", "ref30k", "lowcomplexity"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_RANK_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
