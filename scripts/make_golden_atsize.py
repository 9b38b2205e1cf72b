#!/usr/bin/env python
"""Known answers for bench.py's `parity` key at the sizes the bench runs at.

Runs the oracle (oracle/ped_oracle.c, the CPU restatement of ped.pl pinned to ped.pl's own outputs
by tests/test_oracle.py) ONCE over a whole bench workload and stores
  * sha256 + number of lines of the edge list (<target>.kmer.snp), and
  * rows + order-independent digest of both count tables (ped_b200.api.table_digest)
in tests/golden/atsize_<workload>_s<scale>_k<k>.json -- the file bench.py compares its own values
with at every N.  The inputs come from the same counter-based generator bench.py uses
(libpedk_synth.so, host side); no GPU and nothing of the product library is involved.

Needs the memory of the GPU box's host (configs[1] at k = 20: ~60 GB, a few minutes on 16 cores), so
it is run there through gpurun and the JSON is carried back in gpurun_out/:
    python scripts/make_golden_atsize.py --workload configs1 --out gpurun_out
usage here (small scale):  python scripts/make_golden_atsize.py --workload configs1 --scale 0.01
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (workload table, input recipe)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="configs1", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()

    from oracle import oracle
    import synth  # noqa: F401  (synth/libpedk_synth.so only)
    from ped_b200 import build
    from ped_b200.api import table_digest  # numpy only; loads no library

    build.build_oracle()
    build.build_synth()
    wl = bench.workload(args.workload, args.scale, args.k)
    oracle.set_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    fc, ft = bench.cpu_inputs(wl, wl["reads"])
    t_in = time.perf_counter() - t0
    t0 = time.perf_counter()
    if wl["control"] == "reference":
        oc, ot, snp = oracle.run_ref(fc.tobytes(), ft, wl["k"])
    else:
        oc, ot, snp = oracle.run(fc, ft, wl["k"])
    t_or = time.perf_counter() - t0
    tab = {}
    for name, o in (("control", oc), ("target", ot)):
        km, ct = o.table(copy=False)
        rows, dig = table_digest(km, ct)
        tab[name] = {"rows": rows, "digest": f"{dig:016x}"}
    out = {"workload": wl["name"], "key": wl["key"], "scale": wl["scale"], "k": wl["k"], "genome": wl["genome"],
           "reads_per_sample": wl["reads"], "bases": bench.bases_of(wl, wl["reads"]),
           "edges_sha": hashlib.sha256(snp).hexdigest(), "edges": snp.count(b"\n"), "table_sum": tab,
           "unique_reads": {"control": oc.num_reads, "target": ot.num_reads},
           "oracle_seconds": round(t_or, 1), "input_seconds": round(t_in, 1), "host_cores": os.cpu_count(),
           "made_by": "scripts/make_golden_atsize.py (oracle/ped_oracle.c on the host CPU)"}
    os.makedirs(args.out, exist_ok=True)
    path = os.path.join(args.out, os.path.basename(bench.golden_path(wl)))
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))
    print("wrote", path)


if __name__ == "__main__":
    main()
