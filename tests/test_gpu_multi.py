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
           "adversarial", "clip50", "c1_40x", "ref30k", "lowcomplexity", "ref30k_cnv"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_RANK_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_two_contexts_on_two_devices_in_one_process(built):
    """One process, a context on device 0 and one on device 1 (no communicator): every kernel that opts in to
    more than 48 KB of shared memory must have done so on BOTH devices (the opt-in is per device), and the two
    contexts must not share state.  Interleaved calls, both against the oracle."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import oracle
    from ped_b200 import api
    from tests.cases import CASES, make_inputs

    fc, ft = make_inputs(CASES["c1_40x"])
    oc, ot, osnp = oracle.run(fc, ft, 20)
    ctxs = [api.Context(0), api.Context(1)]
    try:
        samples = []
        for ctx in ctxs:
            c, t = api.Sample(ctx), api.Sample(ctx)
            c.add_fastq(fc)
            t.add_fastq(ft)
            samples.append((c, t))
        for c, t in samples:      # interleaved: device 0, device 1, device 0, ...
            t.ingest()
        for c, t in samples:
            c.ingest()
        for c, t in samples:
            t.count(20)
            c.count(20)
        for ctx, (c, t) in zip(ctxs, samples):
            r = ctx.join(c, t)
            assert r.text() == osnp
            km, ct = t.table()
            okm, oct_ = ot.table()
            assert (km == okm).all() and (ct == oct_).all()
            r.destroy()
        for c, t in samples:
            c.destroy(); t.destroy()
    finally:
        for ctx in ctxs:
            ctx.close()
