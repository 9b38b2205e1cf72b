#!/usr/bin/env python
"""Row f1 + f2 at the size of the headline configuration (needs a B200): a `method=kmer` run WITH a
reference on configs[1]-like input — a 100 Mb reference, target and control of 20 M reads x 150 bp each —
from the resident FASTQ text to `<t>.kmer.snp` (20 columns) and the VCF, through the C ABI:

  reference:  ingest (chromosomes) -> unique 20-mer index              pedk_sample_ingest_reference, _ref20_uniq
  samples:    ingest -> count -> join -> edge list                     (the hot path bench.py times)
  map:        edges x unique 20-mers -> candidate positions            pedk_sample_map_kmer_text     (f1)
  verify:     candidates x reads of both samples -> cw cm tw tm, M/H/R/N  pedk_verify_text            (f2)
  vcf:        pedk_vcf_text

The oracle cannot run at this size in reasonable time, so the check is a property of the input: the target
was made from the reference by substitutions only (no indels), so the planted SNPs are where the two
genomes differ; a planted SNP that is far enough from the chromosome ends and covered on both samples must
come out at its position with ref/alt as planted and genotype M, and no line may claim M for a position
where the genomes agree.  Prints one JSON line.

usage: python scripts/bench_verify.py [--scale 1.0] [--genome 100000000] [--reads 20000000]
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=100_000_000)
    ap.add_argument("--reads", type=int, default=20_000_000)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--read-len", type=int, default=150)
    args = ap.parse_args()
    import torch

    import bench
    from ped_b200 import api

    G = int(args.genome * args.scale)
    N = int(args.reads * args.scale)
    L = args.read_len
    ctx = api.Context(0)
    torch.cuda.set_device(0)
    g = api.synth_genome(1002, G)
    t = api.synth_mutate(g, 2002, 1000, 0)          # substitutions only: position i of both genomes correspond
    assert len(t) == len(g)
    planted = np.flatnonzero(g != t)
    fa = bench.reference_fasta(g)
    gd, td = torch.from_numpy(g).cuda(), torch.from_numpy(t).cuda()
    nb = api.synth_fastq_bytes(0, N, L)
    dev_t = torch.empty(api.fastq_padded_bytes(nb), dtype=torch.uint8, device="cuda")
    dev_c = torch.empty(api.fastq_padded_bytes(nb), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.synth_fastq_device(api.SynthParams(4002, G, L, 2000, 1000, 0), td, 0, N, dev_t)
    ctx.synth_fastq_device(api.SynthParams(3002, G, L, 2000, 1000, 0), gd, 0, N, dev_c)
    torch.cuda.synchronize()
    del gd, td

    def timed(label, fn, out):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        out[label] = round((time.perf_counter() - t0) * 1000.0, 2)
        return r

    ms = {}
    r, c, tt = api.Sample(ctx), api.Sample(ctx), api.Sample(ctx)
    r.add_fastq(fa.tobytes())
    timed("reference_ingest", r.ingest_reference, ms)
    rows = timed("reference_unique_20mers", lambda: r.ref20_uniq(20), ms)
    tt.borrow_fastq_device(dev_t, nb)
    c.borrow_fastq_device(dev_c, nb)

    def hot_path():
        tt.ingest()
        tt.count(20)
        c.ingest()
        c.count(20)
        return ctx.join(c, tt)

    res = timed("ingest_count_join", hot_path, ms)
    edges = timed("edge_text", res.text, ms)
    map_text = timed("map", lambda: r.map_kmer_text(edges), ms)
    ctx.reset_stats()
    snp = timed("verify", lambda: r.verify_text(c, tt, map_text), ms)
    st = ctx.stats()
    snp2 = r.verify_text(c, tt, map_text)
    vcf = timed("vcf", lambda: api.vcf_text(snp, "T"), ms)

    # ---- properties ----
    lines = [x.split("\t") for x in snp.decode().split("\n")[:-1]]
    full = [x for x in lines if len(x) == 20 and x[19] in "MHRN" and x[19] != ""]
    geno = collections.Counter(x[19] if len(x) == 20 else "" for x in lines)
    planted_set = set((planted + 1).tolist())     # positions are 1-based
    m_pos = {}
    bad_counts = 0
    for x in full:
        pos = int(x[1])
        cw, cm, tw, tm = (int(v) for v in x[15:19])
        if max(cw, cm, tw, tm) > L:
            bad_counts += 1
        if x[19] == "M":
            m_pos[pos] = (x[4], x[5])
    false_m = [p for p in m_pos if p not in planted_set]
    inner = [int(p) for p in (planted + 1) if L <= p <= G - L]
    hit = 0
    wrong_allele = 0
    for p in inner:
        if p in m_pos:
            hit += 1
            ref_b, alt_b = chr(g[p - 1]), chr(t[p - 1])
            if m_pos[p] != (ref_b, alt_b):
                wrong_allele += 1
    windows = 2 * L * 2 * len(set((x[0], x[1], x[4], x[5]) for x in lines))  # wild + mutant, two samples (lower bound)
    out = {"what": "method=kmer with a reference: edge list -> position map (f1) -> read-level verification (f2) -> VCF",
           "genome": G, "reads_per_sample": N, "read_len": L, "unique_20mers_of_reference": int(rows),
           "edges": edges.count(b"\n"), "map_lines": map_text.count(b"\n"), "kmer_snp_lines": len(lines),
           "vcf_records": vcf.count(b"\n") - 8, "genotypes": dict(geno), "ms": ms,
           "verify_device_ms": round(sum(st["stage_ms"].values()), 2),
           "windows_looked_up_lower_bound": windows,
           "properties": {"deterministic": snp == snp2, "planted_snps": int(len(planted)),
                          "planted_snps_away_from_the_ends": len(inner), "of_which_called_M_at_their_position": hit,
                          "recall": round(hit / max(1, len(inner)), 4), "called_M_with_other_alleles": wrong_allele,
                          "M_where_the_genomes_agree": len(false_m), "counts_above_read_length": bad_counts}}
    print(json.dumps(out), flush=True)
    res.destroy(); r.destroy(); c.destroy(); tt.destroy()
    ctx.close()
    ok = snp == snp2 and not false_m and not wrong_allele and not bad_counts and hit >= 0.8 * len(inner)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
