"""Launched by torchrun with the gloo backend (CPU): the sharding plan libpedk uses
(owner of a read, owner of a canonical k-mer, bucket ranges for table rows) must turn a
2-rank computation into exactly the single-process result.  The data flow is replayed in
Python on top of the library's own host-side owner functions."""
import os
import sys
from collections import Counter

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from ped_b200 import api  # noqa: E402
from tests.cases import CASES, make_inputs, revcomp  # noqa: E402

CODE = {65: 0, 67: 1, 71: 2, 84: 3}
K = 20


def pack(s: bytes) -> np.ndarray:
    w = np.zeros((len(s) + 31) // 32, dtype=np.uint64)
    for i, c in enumerate(s):
        w[i >> 5] |= np.uint64(CODE[c] << (62 - 2 * (i & 31)))
    return w


def enc(s: bytes) -> int:
    v = 0
    for c in s:
        v = (v << 2) | CODE[c]
    return v


def all_to_all(rank, world, buckets):
    got = [None] * world
    dist.all_to_all_object_list if False else None
    gathered = [None] * world
    dist.all_gather_object(gathered, buckets)  # gloo: gather everything, keep what is mine
    return [x for src in gathered for x in src[rank]]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    case = dict(CASES["adversarial"])
    fc, ft = make_inputs(case)
    tables = []
    for fq in (fc, ft):
        lines = fq.split(b"\n")
        seqs = [l for l in lines[1::4] if l and b"N" not in l]
        mine = seqs[len(seqs) * rank // world:len(seqs) * (rank + 1) // world]
        # reads -> owner by hash of the canonical packed read; dedup locally
        out = [[] for _ in range(world)]
        for s in mine:
            c = min(s, revcomp(s))
            out[api.owner_of_read(pack(c), world)].append(c)
        reads = set(all_to_all(rank, world, out))
        # canonical k-mers -> owner by hash; count locally (palindromic reads weigh one strand)
        out = [[] for _ in range(world)]
        for s in reads:
            w = 1 if s == revcomp(s) else 2
            for i in range(len(s) - K + 1):
                km = s[i:i + K]
                c = min(km, revcomp(km))
                out[api.owner_of_kmer(enc(c), K, world)].append((c, w))
        half = Counter()
        for c, w in all_to_all(rank, world, out):
            half[c] += w
        rows = []
        for c, h in half.items():
            rc = revcomp(c)
            if rc == c:
                rows.append((enc(c), h))
            else:
                assert h % 2 == 0
                rows += [(enc(c), h // 2), (enc(rc), h // 2)]
        tables.append(rows)
    # table rows -> owner of their 3-nt bucket, ranges balanced on the all-reduced histogram
    hist = np.zeros(64, dtype=np.int64)
    for rows in tables:
        for k, _ in rows:
            hist[k >> (2 * K - 6)] += 1
    import torch
    ht = torch.from_numpy(hist)
    dist.all_reduce(ht)
    owner = api.plan_bucket_owners(ht.numpy().astype(np.uint64), world)
    assert (np.diff(owner.astype(int)) >= 0).all() and owner[0] == 0 and owner[-1] == world - 1
    final = []
    for rows in tables:
        out = [[] for _ in range(world)]
        for k, c in rows:
            out[int(owner[k >> (2 * K - 6)])].append((k, c))
        final.append(sorted(all_to_all(rank, world, out)))
    parts = [None] * world
    dist.all_gather_object(parts, final)
    ok = True
    if rank == 0:
        oc, ot, _ = oracle.run(fc, ft, K)
        for i, o in enumerate((oc, ot)):
            km, ct = o.table()
            got = [x for p in parts for x in p[i]]  # rank order must already be key order
            ok &= [k for k, _ in got] == km.tolist() and [c for _, c in got] == ct.tolist()
        print("SHARDING_CHECK", "PASS" if ok else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
