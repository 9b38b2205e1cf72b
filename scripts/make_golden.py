#!/usr/bin/env python
"""Generate tests/golden/*.json by running the UNMODIFIED reference
(/root/reference/ped.pl, method=kmer, no ref) on deterministic inputs.

Runs only in the build container (needs /root/reference, perl, coreutils);
nothing at test time reads /root/reference.  Each golden holds the recipe of
its input (synthetic seeds, or the literal FASTQ for hand-made cases) and, of
the reference's outputs: sha256 of the concatenated `zcat` of sort_uniq/,
count/ and lbc/ (tags AAA..TTT in order), row counts, the `clipping :` log line
and the full text of <target>.kmer.snp.

usage: python scripts/make_golden.py [case ...]
"""
from __future__ import annotations

import gzip
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
from tests.cases import CASES, make_inputs  # noqa: E402

REF = "/root/reference"
TAGS = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT"]


def zcat_all(d: str, pattern: str):
    h = hashlib.sha256()
    rows = 0
    per_tag = {}
    for tag in TAGS:
        p = os.path.join(d, pattern.format(tag=tag))
        with gzip.open(p, "rb") as f:
            b = f.read()
        h.update(b)
        n = b.count(b"\n")
        rows += n
        per_tag[tag] = hashlib.sha256(b).hexdigest()[:16]
    return h.hexdigest(), rows, per_tag


def run_case(name: str) -> dict:
    case = CASES[name]
    control_fq, target_fq = make_inputs(case)
    wd = tempfile.mkdtemp(prefix=f"ped_golden_{name}_")
    try:
        for s, fq in (("C", control_fq), ("T", target_fq)):
            os.makedirs(os.path.join(wd, s, "read"))
            with open(os.path.join(wd, s, "read", s + ".fastq"), "wb") as f:
                f.write(fq)
        arg = f"target=T,control=C,method=kmer,thread=8,wd={wd}"
        if case.get("clipping"):
            arg += f",clipping={case['clipping']}"
        env = dict(os.environ)
        env.pop("LC_ALL", None)
        t0 = time.time()
        subprocess.run(["perl", os.path.join(REF, "ped.pl"), arg], cwd=REF, env=env, check=True,
                       stdout=subprocess.DEVNULL)
        wall = time.time() - t0
        out = {"case": name, "reference_wall_s": round(wall, 1), "k": 20}
        for s in ("C", "T"):
            d = {}
            d["sort_uniq_sha256"], d["sort_uniq_rows"], d["sort_uniq_tags"] = zcat_all(
                os.path.join(wd, s, "sort_uniq"), s + ".sort_uniq.{tag}.gz")
            d["count_sha256"], d["count_rows"], d["count_tags"] = zcat_all(
                os.path.join(wd, s, "count"), s + ".count.{tag}.gz")
            d["lbc_sha256"], d["lbc_rows"], d["lbc_tags"] = zcat_all(os.path.join(wd, s, "lbc"), s + ".lbc.{tag}.gz")
            out["control" if s == "C" else "target"] = d
        with open(os.path.join(wd, "T", "T.kmer.snp"), "rb") as f:
            out["kmer_snp"] = f.read().decode()
        with open(os.path.join(wd, "T", "T.log")) as f:
            out["clipping_lines"] = [l.strip() for l in f if l.startswith("clipping")]
        return out
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def main():
    names = sys.argv[1:] or list(CASES)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for n in names:
        print("running reference on", n, flush=True)
        g = run_case(n)
        with open(os.path.join(ROOT, "tests", "golden", n + ".json"), "w") as f:
            json.dump(g, f, indent=1)
        print(" ", n, g["reference_wall_s"], "s; edges:", g["kmer_snp"].count("\n"), flush=True)


if __name__ == "__main__":
    main()
