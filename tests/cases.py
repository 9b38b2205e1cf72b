"""Deterministic inputs shared by scripts/make_golden.py (which feeds them to the
unmodified reference) and the tests (which feed them to the oracle and to the
CUDA path).  Synthetic reads follow SURVEY.md section 8d."""
from __future__ import annotations

import numpy as np

from ped_b200 import api

# name -> recipe.  seeds follow SURVEY 8d: genome 1000+c, variants 2000+c, control reads 3000+c, target 4000+c
CASES = {
    # 3 kb genome, 800 reads per sample (survey probe P1 shape)
    "tiny": dict(kind="synth", cfg=1, genome=3000, reads=800, L=150, snp_ppm=1000, indel_ppm=300, err_ppm=2000,
                 n_ppm=1000, dup_ppm=0),
    # config 1 at 40x: edges present
    "c1_40x": dict(kind="synth", cfg=2, genome=30000, reads=8000, L=150, snp_ppm=1000, indel_ppm=100, err_ppm=2000,
                   n_ppm=1000, dup_ppm=0),
    # config 1 (BASELINE.json configs[0]): ~1000x, every count > 100 -> no edges
    "c1_1000x": dict(kind="synth", cfg=3, genome=30000, reads=200000, L=150, snp_ppm=1000, indel_ppm=100,
                     err_ppm=2000, n_ppm=1000, dup_ppm=0),
    # many exact duplicate reads (same placement, error-free copies)
    "dup": dict(kind="synth", cfg=4, genome=5000, reads=3000, L=150, snp_ppm=2000, indel_ppm=200, err_ppm=300,
                n_ppm=5000, dup_ppm=400000),
    # explicit clipping: one 100-base chunk per read / three 50-base chunks per read
    "clip100": dict(kind="synth", cfg=5, genome=5000, reads=2000, L=150, snp_ppm=2000, indel_ppm=200, err_ppm=2000,
                    n_ppm=1000, dup_ppm=0, clipping=100),
    "clip50": dict(kind="synth", cfg=6, genome=4000, reads=1500, L=150, snp_ppm=2000, indel_ppm=200, err_ppm=2000,
                   n_ppm=1000, dup_ppm=0, clipping=50),
    # mixed read lengths, no clipping=: ped.pl's automatic clipping incl. the reads already written unclipped
    "mixed": dict(kind="mixed", cfg=7, genome=5000, reads=2500, L=150, snp_ppm=2000, indel_ppm=200, err_ppm=2000,
                  n_ppm=1000, dup_ppm=0),
    # palindromic reads, palindromic 20-mers, reverse-complement duplicates, no final newline
    "adversarial": dict(kind="adversarial", cfg=8, genome=4000, reads=1200, L=150, snp_ppm=2000, indel_ppm=200,
                        err_ppm=1000, n_ppm=2000, dup_ppm=100000),
}

_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(s: bytes) -> bytes:
    return s.translate(_COMP)[::-1]


def genomes(case):
    c = case["cfg"]
    g = api.synth_genome(1000 + c, case["genome"])
    if case["kind"] == "adversarial":
        # plant two palindromic 20-mers (each equals its own reverse complement)
        for pos, half in ((500, b"ACGTTGCAAT"), (2100, b"GGATCCATGA")):
            pal = half + revcomp(half)
            g[pos:pos + 20] = np.frombuffer(pal, dtype=np.uint8)
    t = api.synth_mutate(g, 2000 + c, case["snp_ppm"], case["indel_ppm"])
    return g, t


def _records(fq: bytes):
    lines = fq.split(b"\n")
    assert lines[-1] == b""
    lines = lines[:-1]
    return [lines[i:i + 4] for i in range(0, len(lines), 4)]


def _join(recs) -> bytes:
    return b"".join(b"\n".join(r) + b"\n" for r in recs)


def _fq(seqs) -> bytes:
    return "".join(f"@r{i}\n{s}\n+\n{'I' * len(s)}\n" for i, s in enumerate(seqs)).encode()


def low_complexity_inputs(n: int = 3000, seed: int = 11):
    """Reads that are mostly poly-A / (AC)n: a handful of k-mers carry almost all instances, so on
    several ranks the optimistic level-A segments overflow and the exchange is redone with exact sizes."""
    import random

    rnd = random.Random(seed)
    reads = []
    for _ in range(n):
        u = "".join(rnd.choice("ACGT") for _ in range(10))
        reads.append(u + "A" * 130 + "".join(rnd.choice("ACGT") for _ in range(10)))
        reads.append(u + "AC" * 65 + "".join(rnd.choice("ACGT") for _ in range(10)))
    return _fq(reads), _fq(reads[: len(reads) // 2] + ["ACGT" * 37 + "AC"] * 3)


def make_inputs(case):
    """-> (control_fastq, target_fastq) as bytes."""
    if case["kind"] == "lowcomplexity":
        return low_complexity_inputs(case["reads"])
    c = case["cfg"]
    g, t = genomes(case)
    kw = dict(read_len=case["L"], err_ppm=case["err_ppm"], n_ppm=case["n_ppm"], dup_ppm=case["dup_ppm"])
    fc = api.synth_fastq(g, 3000 + c, case["reads"], **kw)
    ft = api.synth_fastq(t, 4000 + c, case["reads"], **kw)
    if case["kind"] == "synth":
        return fc, ft
    out = []
    for si, fq in enumerate((fc, ft)):
        recs = _records(fq)
        if case["kind"] == "mixed":
            for i, r in enumerate(recs):
                if i < 300:
                    continue  # a uniform prefix: these are written unclipped before the first length change
                cut = 120 if i % 7 == 3 else 90 if i % 11 == 5 else 0
                if cut:
                    r[1] = r[1][:cut]
                    r[3] = r[3][:cut]
            out.append(_join(recs))
        else:  # adversarial
            gg = g if si == 0 else t
            gb = gg.tobytes()
            extra = []
            # palindromic reads: x + revcomp(x); two copies of the first one
            for j, pos in enumerate((100, 100, 900, 1700)):
                x = gb[pos:pos + 75]
                extra.append([b"@pal%d" % j, x + revcomp(x), b"+", b"I" * 150])
            # the reverse complement of an existing read, and an exact copy of another
            extra.append([b"@rc", revcomp(recs[10][1]) if b"N" not in recs[10][1] else revcomp(recs[11][1]), b"+",
                          b"I" * 150])
            extra.append([b"@copy", recs[20][1], b"+", b"I" * 150])
            # reads covering the planted palindromic 20-mers on both samples
            for j, pos in enumerate((430, 440, 450, 2050, 2060)):
                extra.append([b"@cov%d" % j, gb[pos:pos + 150], b"+", b"I" * 150])
            allr = recs[:600] + extra + recs[600:]
            b = _join(allr)
            out.append(b[:-1])  # no newline after the last quality line
    return out[0], out[1]


# no golden (not run through ped.pl): checked against the oracle only
CASES["lowcomplexity"] = dict(kind="lowcomplexity", cfg=9, reads=3000)

# ---- the reference as the control (BASELINE config 5 shape) ----
REF_CASES = {
    # 30 kb reference in two sequences: 60-column lines, a soft-masked stretch, a run of N, an
    # IUPAC code, a CRLF line, a sequence shorter than one window, a name that comes back
    # (the later sequence replaces the earlier file), text before the first header
    "ref30k": dict(cfg=2, genome=30000, reads=8000, L=150, snp_ppm=1000, indel_ppm=100, err_ppm=2000, n_ppm=1000,
                   dup_ppm=0),
}


def make_ref_inputs(case):
    """-> (reference FASTA bytes, target FASTQ bytes).  The target reads come from the mutated genome."""
    c = case["cfg"]
    g = api.synth_genome(1000 + c, case["genome"])
    t = api.synth_mutate(g, 2000 + c, case["snp_ppm"], case["indel_ppm"])
    ft = api.synth_fastq(t, 4000 + c, case["reads"], read_len=case["L"], err_ppm=case["err_ppm"],
                         n_ppm=case["n_ppm"], dup_ppm=case["dup_ppm"])
    gb = g.tobytes()
    half = len(gb) * 3 // 5
    c1, c2 = bytearray(gb[:half]), bytearray(gb[half:])
    c1[5000:5040] = b"N" * 40
    c1[9000:9100] = bytes(c1[9000:9100]).lower()
    c2[3000] = ord("R")

    def fa(name, seq, width=60, crlf_at=-1):
        s = bytes(seq)
        lines = [s[i:i + width] for i in range(0, len(s), width)]
        if 0 <= crlf_at < len(lines):
            lines[crlf_at] += b"\r"
        return b">" + name + b"\n" + b"\n".join(lines) + b"\n"

    text = (b"ACGTACGT stray line before any header\n" + fa(b"chr01 first sequence", c1, crlf_at=7) +
            fa(b"scaffoldX", gb[100:160]) + fa(b"Chr2\tsecond", gb[200:900], width=70) + fa(b"2 again", c2, width=80))
    return text[:-1], ft  # no newline after the last sequence line


def make_cnv_inputs(case):
    """-> (reference FASTA, control FASTQ, target FASTQ) for `method=cnv` (ped.pl:362-366, 934-989): the
    reference of make_ref_inputs, control reads from the reference genome, target reads from the mutated one."""
    fa, ft = make_ref_inputs(case)
    c = case["cfg"]
    g = api.synth_genome(1000 + c, case["genome"])
    fc = api.synth_fastq(g, 3000 + c, case["reads"], read_len=case["L"], err_ppm=case["err_ppm"], n_ppm=case["n_ppm"],
                         dup_ppm=case["dup_ppm"])
    return fa, fc, ft


def make_verify_inputs(seed: int, L: int = 100, CL: int = 80, chrom_len: int = 5000, coverage: int = 40):
    """-> (reference FASTA, control FASTQ, target FASTQ) for the read-level verification (SURVEY row f2,
    ped.pl:2087-2500): three chromosomes ("1", "2", "scX"), a diploid target with homozygous and
    heterozygous SNPs, SNPs 3 / 25 / 80 bases apart (a window then carries its neighbours, ped.pl:2184-2192),
    one closer to the start of a chromosome than a read is long and one that close to its end, control reads
    of another length (CL) drawn from the reference itself."""
    import random
    rnd = random.Random(seed)
    names = ["1", "2", "scX"]
    ref = {n: [rnd.choice("ACGT") for _ in range(chrom_len)] for n in names}

    def other(b):
        return rnd.choice([x for x in "ACGT" if x != b])

    hap = [{n: list(s) for n, s in ref.items()}, {n: list(s) for n, s in ref.items()}]
    for n in names:
        sites = [(40, "hom"), (chrom_len - 50, "hom"), (700, "hom"), (703, "hom"), (1500, "het"), (1525, "hom"),
                 (2300, "hom"), (2380, "het"), (3000, "het"), (3600, "hom")]
        sites += [(p, rnd.choice(["hom", "het"])) for p in rnd.sample(range(3700, chrom_len - 200), 6)]
        for p, kind in sites:
            a = other(ref[n][p])
            hap[0][n][p] = a
            if kind == "hom":
                hap[1][n][p] = a

    def reads(genomes, length, cov, tag):
        out = []
        i = 0
        for g in genomes:
            for n in names:
                s = "".join(g[n])
                for _ in range(len(s) * cov // (length * len(genomes))):
                    a = rnd.randrange(0, len(s) - length + 1)
                    r = s[a:a + length]
                    if rnd.random() < 0.5:
                        r = r[::-1].translate(str.maketrans("ACGT", "TGCA"))
                    if rnd.random() < 0.05:   # a sequencing error
                        q = rnd.randrange(length)
                        r = r[:q] + other(r[q]) + r[q + 1:]
                    out.append(f"@{tag}{i}\n{r}\n+\n{'I' * length}\n")
                    i += 1
        rnd.shuffle(out)
        return "".join(out).encode()

    fa = "".join(f">{n}\n" + "\n".join("".join(ref[n])[i:i + 60] for i in range(0, chrom_len, 60)) + "\n" for n in names)
    return fa.encode(), reads([ref], CL, coverage, "c"), reads(hap, L, coverage, "t")
