#!/usr/bin/env python
"""Generate tests/golden/<case>.json for the reference-as-control cases (tests/cases.py REF_CASES)
by running the UNMODIFIED reference in the build container:

  perl ped.pl ref=REF,file=REF.fa,wd=WD           mkChrFromFile, mkControlRead (ped.pl:1027-1136)
  perl ped.pl target=T,control=REF,method=kmer    count / lbc / join with the tiled reference as control
  perl ped.pl target=T,ref=REF,sub=mapKmerSub,arg=TAG   (groundwork for SURVEY 8f row f1) the edges of each
                                                  bucket joined with REF/ref20_uniq.TAG.gz (ped.pl:641-685)

The first run leaves REF/chrNAME and REF/sort_uniq/; the second REF/count, REF/lbc, T/* and
T/T.kmer.snp (no-ref format: the edge list itself, before ped.pl's mapping stages).

usage: python scripts/make_golden_ref.py [case ...]
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
from scripts.make_golden import REF, zcat_all  # noqa: E402
from tests.cases import REF_CASES, make_ref_inputs  # noqa: E402


def run_case(name: str) -> dict:
    fasta, target_fq = make_ref_inputs(REF_CASES[name])
    wd = tempfile.mkdtemp(prefix=f"ped_golden_{name}_")
    try:
        os.makedirs(os.path.join(wd, "REF"))
        os.makedirs(os.path.join(wd, "T", "read"))
        with open(os.path.join(wd, "REF", "REF.fa"), "wb") as f:
            f.write(fasta)
        with open(os.path.join(wd, "T", "read", "T.fastq"), "wb") as f:
            f.write(target_fq)
        env = dict(os.environ)
        env.pop("LC_ALL", None)
        t0 = time.time()
        subprocess.run(["perl", os.path.join(REF, "ped.pl"), f"ref=REF,file=REF.fa,thread=8,wd={wd}"], cwd=REF, env=env,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t1 = time.time()
        out = {"case": name, "k": 20, "reference_wall_s": {"mkref": round(t1 - t0, 1)}}
        chrs = {}
        for fn in sorted(os.listdir(os.path.join(wd, "REF"))):
            if fn.startswith("chr") and os.path.isfile(os.path.join(wd, "REF", fn)):
                b = open(os.path.join(wd, "REF", fn), "rb").read()
                chrs[fn] = {"length": len(b), "sha256": hashlib.sha256(b).hexdigest()}
        out["chromosomes"] = chrs
        d = {}
        d["sort_uniq_sha256"], d["sort_uniq_rows"], d["sort_uniq_tags"] = zcat_all(
            os.path.join(wd, "REF", "sort_uniq"), "REF.sort_uniq.{tag}.gz")
        subprocess.run(["perl", os.path.join(REF, "ped.pl"), f"target=T,control=REF,method=kmer,thread=8,wd={wd}"],
                       cwd=REF, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        out["reference_wall_s"]["kmer"] = round(time.time() - t1, 1)
        d["count_sha256"], d["count_rows"], d["count_tags"] = zcat_all(os.path.join(wd, "REF", "count"),
                                                                       "REF.count.{tag}.gz")
        d["lbc_sha256"], d["lbc_rows"], d["lbc_tags"] = zcat_all(os.path.join(wd, "REF", "lbc"), "REF.lbc.{tag}.gz")
        out["control"] = d
        t = {}
        t["sort_uniq_sha256"], t["sort_uniq_rows"], t["sort_uniq_tags"] = zcat_all(
            os.path.join(wd, "T", "sort_uniq"), "T.sort_uniq.{tag}.gz")
        t["count_sha256"], t["count_rows"], t["count_tags"] = zcat_all(os.path.join(wd, "T", "count"), "T.count.{tag}.gz")
        t["lbc_sha256"], t["lbc_rows"], t["lbc_tags"] = zcat_all(os.path.join(wd, "T", "lbc"), "T.lbc.{tag}.gz")
        out["target"] = t
        with open(os.path.join(wd, "T", "T.kmer.snp"), "rb") as f:
            out["kmer_snp"] = f.read().decode()
        # ---- groundwork for f1: the unique-20-mer index of the reference and the edge -> position map ----
        from scripts.make_golden import TAGS
        r = {}
        r["sha256"], r["rows"], r["tags"] = zcat_all(os.path.join(wd, "REF"), "ref20_uniq.{tag}.gz")
        out["ref20_uniq"] = r
        tmpdir = os.path.join(wd, "T", "tmp")
        os.makedirs(tmpdir, exist_ok=True)
        os.makedirs(os.path.join(wd, "child"), exist_ok=True)
        by_tag = {t: [] for t in TAGS}
        for line in out["kmer_snp"].splitlines(keepends=True):
            by_tag[line[:3]].append(line)
        maps = {}
        for t in TAGS:
            with open(os.path.join(tmpdir, f"T.snp.{t}"), "w") as f:
                f.write("".join(by_tag[t]))
            subprocess.run(["perl", os.path.join(REF, "ped.pl"), f"target=T,ref=REF,sub=mapKmerSub,arg={t},wd={wd}"],
                           cwd=REF, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            with open(os.path.join(tmpdir, f"T.map.{t}")) as f:
                m = f.read()
            if m:
                maps[t] = m
        out["map"] = maps
        return out
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def main():
    names = sys.argv[1:] or list(REF_CASES)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for n in names:
        g = run_case(n)
        with open(os.path.join(ROOT, "tests", "golden", n + ".json"), "w") as f:
            json.dump(g, f, indent=1)
        print(n, g["reference_wall_s"], "chromosomes", {k: v["length"] for k, v in g["chromosomes"].items()},
              "control reads", g["control"]["sort_uniq_rows"], "edges", g["kmer_snp"].count("\n"))


if __name__ == "__main__":
    main()
