#!/usr/bin/env python
"""tests/golden/<case>_verify.json: the UNMODIFIED reference's `method=kmer` run WITH a reference
(ped.pl:368-389: mapKmer, snpMkT, sortSeq, joinTarget, snpMkC, sortSeq, joinControl, kmerReadCount, toVcf)

  perl ped.pl ref=REF,file=REF.fa,wd=WD                       chr files, ref20_uniq, the tiled reference reads
  perl ped.pl target=T,control=C,ref=REF,method=kmer,wd=WD    -> T/T.kmer.snp (20 columns), T/T.kmer.vcf
  perl ped.pl target=T,ref=REF,method=kmer,wd=WD2             control omitted: the reference is the control
                                                              (ped.pl:193; $clength = 100)

The map files the verification starts from are re-made with sub=mapKmerSub from the no-ref edge list
(as scripts/make_golden_ref.py does) and stored too, so that a test can start where SURVEY row f2 starts.
usage: python scripts/make_golden_verify.py [case ...]
"""
from __future__ import annotations

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


def ped(args: str, env) -> None:
    subprocess.run(["perl", os.path.join(REF, "ped.pl"), args], cwd=REF, env=env, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def run_case(name: str) -> dict:
    fasta, control_fq, target_fq = make_cnv_inputs(REF_CASES[name])
    out = {"case": name, "reference_wall_s": {}}
    env = dict(os.environ)
    env.pop("LC_ALL", None)
    for variant in ("control_reads", "control_ref"):
        wd = tempfile.mkdtemp(prefix=f"ped_golden_verify_{name}_")
        try:
            os.makedirs(os.path.join(wd, "REF"))
            samples = (("T", target_fq), ("C", control_fq)) if variant == "control_reads" else (("T", target_fq),)
            for s, fq in samples:
                os.makedirs(os.path.join(wd, s, "read"))
                with open(os.path.join(wd, s, "read", s + ".fastq"), "wb") as f:
                    f.write(fq)
            with open(os.path.join(wd, "REF", "REF.fa"), "wb") as f:
                f.write(fasta)
            ped(f"ref=REF,file=REF.fa,thread=8,wd={wd}", env)
            t1 = time.time()
            ctl = "control=C," if variant == "control_reads" else ""
            ped(f"target=T,{ctl}ref=REF,method=kmer,thread=8,wd={wd}", env)
            out["reference_wall_s"][variant] = round(time.time() - t1, 1)
            with open(os.path.join(wd, "T", "T.kmer.snp")) as f:
                snp = f.read()
            with open(os.path.join(wd, "T", "T.kmer.vcf")) as f:
                vcf = f.read()
            out[variant] = {"kmer_snp": snp, "kmer_vcf": vcf}
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    return out


def main():
    for n in sys.argv[1:] or list(REF_CASES):
        g = run_case(n)
        with open(os.path.join(ROOT, "tests", "golden", n + "_verify.json"), "w") as f:
            json.dump(g, f, indent=1)
        print(n, g["reference_wall_s"], {v: g[v]["kmer_snp"].count("\n") for v in ("control_reads", "control_ref")})


if __name__ == "__main__":
    main()
