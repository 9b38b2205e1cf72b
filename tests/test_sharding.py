"""Host-side logic of the multi-GPU path on CPU: the bucket plan, and a world_size-2
gloo replay of the sharded data flow against the oracle."""
import os
import subprocess
import sys

import numpy as np

from ped_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bucket_plan_is_contiguous_balanced_and_total(built):
    rng = np.random.default_rng(3)
    for world in (1, 2, 4, 8):
        for hist in (np.full(64, 1000, dtype=np.uint64), rng.integers(0, 10_000, 64).astype(np.uint64),
                     np.zeros(64, dtype=np.uint64), np.eye(1, 64, 5, dtype=np.uint64)[0] * 7):
            owner = api.plan_bucket_owners(hist, world).astype(int)
            assert owner[0] == 0 and owner[-1] == world - 1
            assert (np.diff(owner) >= 0).all() and (np.diff(owner) <= 1).all()  # contiguous, no rank skipped
            if hist.sum() and hist.min() > 0:
                load = np.bincount(owner, weights=hist.astype(float), minlength=world)
                assert load.max() <= hist.sum() / world + hist.max()


def test_owner_functions_are_deterministic_and_spread(built):
    ks = np.arange(1, 20001, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    for world in (2, 8):
        o = np.array([api.owner_of_kmer(int(k), 20, world) for k in ks[:4000]])
        assert o.min() == 0 and o.max() == world - 1
        assert np.bincount(o, minlength=world).min() > 4000 / world * 0.8
        w = np.array([7, 8, 9, 10, 11], dtype=np.uint64)
        assert api.owner_of_read(w, world) == api.owner_of_read(w.copy(), world)


def test_two_rank_gloo_replay_matches_oracle(built):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "sharding_gloo_check.py")]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and "SHARDING_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
