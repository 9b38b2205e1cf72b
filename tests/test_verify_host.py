"""The host side of the read-level verification without a GPU: libpedk's own text logic --
candidates, the reference / mutant strings whose windows are looked up, and everything kmerReadCount makes of
the counters -- through its two host-only entry points, against the line-by-line restatement of the Perl
(oracle/verify_ref.py) and ped.pl's own T.kmer.snp.  The look-up itself (a set membership here) is the kernel's
job and is tested on the GPU (tests/test_gpu_parity.py::test_verification_*)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle, verify_ref
from ped_b200 import api
from tests.cases import REF_CASES, make_cnv_inputs, make_verify_inputs
from tests.test_oracle import verification_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _through_the_host_halves(map_text, t_lines, c_lines, chromosomes):
    length = verify_ref.first_line_length(t_lines)
    clength = verify_ref.first_line_length(c_lines)
    files = {n: s.encode() for n, s in chromosomes.items()}
    snps = verify_ref.snp_list(map_text.decode().splitlines())
    hits = []
    for L, lines, kinds in ((length, t_lines, ("tw", "tm")), (clength, c_lines, ("cw", "cm"))):
        reads = set(r for v in lines.values() for r in v)
        strings = api.verify_strings(map_text, files, L)
        # the candidates are ped.pl's snp.list, in its order
        assert [s[0].decode() for s in strings] == [" ".join([verify_ref._strip_zeros(x.split()[0]),
                                                              str(int(x.split()[1]))] + x.split()[2:]) if x.split() else ""
                                                    for x in snps]
        # the windows of the strings are the lines snpMkT / snpMkC print (those that reach a tag file)
        got = []
        h = np.zeros(2 * len(strings), dtype=np.uint32)
        for c, (cand, ref_seq, mut_seq) in enumerate(strings):
            for j, (s, kind) in enumerate(((ref_seq.decode(), kinds[0]), (mut_seq.decode(), kinds[1]))):
                for i in range(L if s else 0):
                    w = s[i:i + L]
                    if w[:3] in verify_ref.TAGS:
                        got.append((w, f"{cand.decode()} {kind}"))
                    if len(w) == L and w in reads:
                        h[2 * c + j] += 1
        want = list(verify_ref.windows(snps, chromosomes, L, kinds[0], kinds[1]))
        assert sorted(got) == sorted(want)
        hits.append(h)
    return api.verify_counts_text(map_text, hits[0], hits[1], length, clength).decode()


@pytest.mark.parametrize("variant", ["control_reads", "control_ref"])
def test_host_halves_reproduce_the_reference(variant, built):
    g = json.load(open(os.path.join(GOLDEN, "ref30k_verify.json")))[variant]
    fa, fc, ft = make_cnv_inputs(REF_CASES["ref30k"])
    oc, ot, edges = oracle.run(fc, ft, 20) if variant == "control_reads" else oracle.run_ref(fa, ft, 20)
    map_text, t_lines, c_lines, chromosomes = verification_inputs(fa, oc, ot, edges)
    assert _through_the_host_halves(map_text, t_lines, c_lines, chromosomes) == g["kmer_snp"]


@pytest.mark.parametrize("seed,L,CL", [(1, 100, 80), (4, 150, 33), (5, 33, 150), (6, 31, 32)])
def test_host_halves_against_the_restatement(seed, L, CL, built):
    fa, fc, ft = make_verify_inputs(seed, L=L, CL=CL)
    oc, ot, edges = oracle.run(fc, ft, 20)
    map_text, t_lines, c_lines, chromosomes = verification_inputs(fa, oc, ot, edges)
    want = verify_ref.verify(map_text.decode().splitlines(), t_lines, c_lines, chromosomes)
    assert _through_the_host_halves(map_text, t_lines, c_lines, chromosomes) == want


def test_host_halves_on_made_up_map_lines(built):
    """Map lines ped.pl could print but the generators above do not: a multi-base alt, two alts at one position
    (one `chr:pos` key, joined crosswise), a chromosome without a file, a name that is not a number, positions
    next to each other, a line whose fields moved left (an F7 row), a candidate closer to the start than a read."""
    import random
    rnd = random.Random(5)
    L = 40
    chromosomes = {"chr1": "".join(rnd.choice("ACGT") for _ in range(900)),
                   "chrX": "".join(rnd.choice("ACGT") for _ in range(500))}
    def line(chrom, pos, ref, alt, strand="f", counts=(0, 9, 0, 0, 0, 0, 9, 0)):
        return "\t".join([chrom, str(pos), "ACGTACGTACGTACGTACG", ref, ref, alt, strand] + [str(x) for x in counts]) + "\n"
    s1 = chromosomes["chr1"]
    rows = [line("1", 100, s1[99], "G" if s1[99] != "G" else "T"), line("1", 103, s1[102], s1[102] + "T"),
            line("1", 103, s1[102], "CA" if s1[102] not in "CA" else "GT"), line("1", 120, s1[119], "A" if s1[119] != "A" else "C"),
            line("1", 12, s1[11], "G" if s1[11] != "G" else "T"), line("1", 880, s1[879], "T" if s1[879] != "T" else "A"),
            line("X", 200, chromosomes["chrX"][199], "C" if chromosomes["chrX"][199] != "C" else "G", "r"),
            line("007", 300, "A", "C"), line("nofile", 300, "A", "C"),
            "1\t400\tACGTACGTACGTACGTACG\t" + s1[399] + "\t" + s1[399] + "\t\tf\t0\t9\t0\t0\t0\t0\t9\n"]
    map_text = "".join(rows).encode()
    # reads: every window of the reference of both chromosomes, and the mutant windows of the first candidate
    reads = set()
    for s in chromosomes.values():
        for i in range(0, len(s) - L + 1, 3):
            reads.add(s[i:i + L])
    mut = s1[:99] + rows[0].split("\t")[5] + s1[100:]
    for i in range(60, 100):
        reads.add(mut[i:i + L])
    comp = str.maketrans("ACGT", "TGCA")
    lines = {t: sorted(r for r in (reads | {x[::-1].translate(comp) for x in reads}) if r[:3] == t) for t in verify_ref.TAGS}
    want = verify_ref.verify(map_text.decode().splitlines(), lines, lines, chromosomes)
    assert want.count("\n") >= 6
    assert _through_the_host_halves(map_text, lines, lines, chromosomes) == want
