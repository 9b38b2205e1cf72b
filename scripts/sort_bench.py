#!/usr/bin/env python
"""Micro-benchmark of the radix pass (and RLE) alone: PEDK_RANK_MODE is read once per process."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ped_b200 import api

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 28
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ctx = api.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
keys0 = torch.randint(0, 1 << bits, (n,), dtype=torch.int64, device="cuda", generator=g)
for rep in range(3):
    keys = keys0.clone()
    torch.cuda.synchronize()
    ctx.reset_stats()
    t0 = time.perf_counter()
    ctx.test_sort_u64(keys, n, bits)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    st = ctx.stats()
    passes = (bits + 7) // 8
    ms = st["sort_pass_ms"] / passes
    print(f"mode={os.environ.get('PEDK_RANK_MODE', 'default')} n={n} bits={bits} rep={rep}: {ms:.3f} ms/pass "
          f"{16 * n / ms / 1e6:.0f} GB/s  (wall {wall:.1f} ms, sort stage {st['stage_ms']['sort']:.1f} ms)", flush=True)
ok = bool((keys[1:] >= keys[:-1]).all())
print("sorted:", ok)
uk = torch.empty(n, dtype=torch.int64, device="cuda")
uc = torch.empty(n, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for rep in range(2):
    ctx.reset_stats()
    D = ctx.test_rle(keys, n, uk, uc)
    st = ctx.stats()
    print(f"rle: D={D}  stage count {st['stage_ms']['count']:.3f} ms (events around kernels only)", flush=True)
ctx.close()
