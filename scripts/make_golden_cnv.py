#!/usr/bin/env python
"""tests/golden/<case>_cnv.json: the UNMODIFIED reference's `method=cnv` output (ped.pl:362-366, 934-989)

  perl ped.pl ref=REF,file=REF.fa,wd=WD                      reference: chr files, ref20_uniq, control reads
  perl ped.pl target=T,control=C,ref=REF,method=cnv,wd=WD    count tables of T and C, full outer join, positions

-> sha256 / rows / head of WD/T/T.C.cnv (lines `chr pos strand target_count control_count`, sorted by
`sort -k 1 -n -k 2 -n`).  usage: python scripts/make_golden_cnv.py [case ...]
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scripts.make_golden import REF  # noqa: E402
from tests.cases import REF_CASES, make_cnv_inputs  # noqa: E402


def run_case(name: str) -> dict:
    fasta, control_fq, target_fq = make_cnv_inputs(REF_CASES[name])
    wd = tempfile.mkdtemp(prefix=f"ped_golden_cnv_{name}_")
    try:
        os.makedirs(os.path.join(wd, "REF"))
        for s, fq in (("T", target_fq), ("C", control_fq)):
            os.makedirs(os.path.join(wd, s, "read"))
            with open(os.path.join(wd, s, "read", s + ".fastq"), "wb") as f:
                f.write(fq)
        with open(os.path.join(wd, "REF", "REF.fa"), "wb") as f:
            f.write(fasta)
        env = dict(os.environ)
        env.pop("LC_ALL", None)
        t0 = time.time()
        subprocess.run(["perl", os.path.join(REF, "ped.pl"), f"ref=REF,file=REF.fa,thread=8,wd={wd}"], cwd=REF, env=env,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t1 = time.time()
        subprocess.run(["perl", os.path.join(REF, "ped.pl"), f"target=T,control=C,ref=REF,method=cnv,thread=8,wd={wd}"],
                       cwd=REF, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t2 = time.time()
        b = open(os.path.join(wd, "T", "T.C.cnv"), "rb").read()
        lines = b.splitlines(keepends=True)
        return {"case": name, "reference_wall_s": {"mkref": round(t1 - t0, 1), "cnv": round(t2 - t1, 1)},
                "cnv_sha256": hashlib.sha256(b).hexdigest(), "cnv_rows": len(lines),
                "cnv_head": b"".join(lines[:40]).decode(), "cnv_tail": b"".join(lines[-10:]).decode()}
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def main():
    for n in sys.argv[1:] or list(REF_CASES):
        g = run_case(n)
        with open(os.path.join(ROOT, "tests", "golden", n + "_cnv.json"), "w") as f:
            json.dump(g, f, indent=1)
        print(n, g["reference_wall_s"], "rows", g["cnv_rows"])


if __name__ == "__main__":
    main()
