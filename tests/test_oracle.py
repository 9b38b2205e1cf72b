"""The C oracle against (a) a naive Python restatement of ped.pl and (b) the
outputs of the unmodified reference itself (tests/golden/, made by
scripts/make_golden.py).  CPU only."""
import hashlib
import json
import os

import pytest

from oracle import oracle
from tests import pyref
from tests.cases import CASES, REF_CASES, make_inputs, make_ref_inputs, revcomp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TAGS = oracle.TAGS


def _sha_tags(fn):
    h = hashlib.sha256()
    rows = 0
    for t in range(64):
        b = fn(t)
        h.update(b)
        rows += b.count(b"\n")
    return h.hexdigest(), rows


def _check_against_pyref(fc: bytes, ft: bytes, k=20, clipping=0):
    ref = pyref.run(fc.decode(), ft.decode(), k, clipping)
    c, t, snp = oracle.run(fc, ft, k, clipping)
    for name, o in (("control", c), ("target", t)):
        r = ref[name]
        assert o.clipping_used == r["clipping"]
        for i, tag in enumerate(TAGS):
            assert o.sort_uniq_text(i).decode() == "".join(x + "\n" for x in r["sort_uniq"][tag]), (name, tag)
            assert o.count_text(i).decode() == "".join(f"{s}\t{n}\n" for s, n in r["count"][tag]), (name, tag)
            assert o.lbc_text(i).decode() == r["lbc"][tag], (name, tag)
    assert snp.decode() == ref["kmer_snp"]
    return ref


@pytest.mark.parametrize("name", ["tiny", "adversarial", "mixed", "clip50"])
def test_oracle_matches_python_restatement(name, built):
    case = dict(CASES[name])
    if name != "tiny":
        case["reads"] = min(case["reads"], 900)
    fc, ft = make_inputs(case)
    ref = _check_against_pyref(fc, ft, 20, case.get("clipping", 0))
    if name == "tiny":
        assert ref["kmer_snp"].count("\n") > 0


def _fq(seqs):
    return "".join(f"@r{i}\n{s}\n+\n{'I' * len(s)}\n" for i, s in enumerate(seqs)).encode()


def test_oracle_hand_made_corner_cases(built):
    import random

    rnd = random.Random(7)
    g = "".join(rnd.choice("ACGT") for _ in range(400))
    reads = [g[i:i + 60] for i in range(0, 340, 3)]
    half = g[100:130]
    pal = half + revcomp(half.encode()).decode()  # a read equal to its own reverse complement
    reads += [pal, pal, revcomp(reads[5].encode()).decode(), reads[7], "ACGTN" * 12, "AC"]
    tgt = list(reads)
    tgt[30] = tgt[30][:20] + ("A" if tgt[30][20] != "A" else "C") + tgt[30][21:]
    fc, ft = _fq(reads), _fq(tgt)
    _check_against_pyref(fc, ft)
    # file that ends inside a record / without the last newline
    _check_against_pyref(fc[:-1], ft[:-70])
    # other k (the reference hard-codes 20: ped.pl:912-913 with the literal edited)
    for k in (5, 19, 31, 32):
        _check_against_pyref(fc, ft, k)
    # empty and N-only inputs
    _check_against_pyref(b"", ft)
    _check_against_pyref(_fq(["NNNN" * 10]), ft)


def test_oracle_strand_symmetry(built):
    fc, ft = make_inputs(CASES["tiny"])
    c, _, _ = oracle.run(fc, ft)
    km, ct = c.table()
    table = dict(zip(km.tolist(), ct.tolist()))
    comp = {0: 3, 1: 2, 2: 1, 3: 0}
    for kmer, n in table.items():
        rc = 0
        v = kmer
        for _ in range(20):
            rc = (rc << 2) | comp[v & 3]
            v >>= 2
        assert table[rc] == n  # the read set is closed under reverse complement


def _goldens():
    if not os.path.isdir(GOLDEN):
        return []
    return sorted(f[:-5] for f in os.listdir(GOLDEN) if f.endswith(".json") and f[:-5] in CASES)


@pytest.mark.parametrize("name", _goldens())
def test_oracle_matches_reference_golden(name, built):
    """Bit-exact against ped.pl's own sort_uniq/, count/, lbc/ and <t>.kmer.snp."""
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        g = json.load(f)
    case = CASES[name]
    fc, ft = make_inputs(case)
    c, t, snp = oracle.run(fc, ft, g["k"], case.get("clipping", 0))
    for key, o in (("control", c), ("target", t)):
        sha, rows = _sha_tags(o.sort_uniq_text)
        assert (rows, sha) == (g[key]["sort_uniq_rows"], g[key]["sort_uniq_sha256"]), (key, "sort_uniq")
        sha, rows = _sha_tags(o.count_text)
        assert (rows, sha) == (g[key]["count_rows"], g[key]["count_sha256"]), (key, "count")
        sha, rows = _sha_tags(o.lbc_text)
        assert (rows, sha) == (g[key]["lbc_rows"], g[key]["lbc_sha256"]), (key, "lbc")
    assert snp.decode() == g["kmer_snp"]
    # "clipping : N" lines of <t>.log: one per mkSortUniq call, target first (ped.pl:291-297)
    want = [f"clipping : {case['clipping']}"] if case.get("clipping") else []  # the argument echo (ped.pl:105)
    assert g["clipping_lines"] == want + [f"clipping : {t.clipping_used}", f"clipping : {c.clipping_used}"]


@pytest.mark.parametrize("name", sorted(REF_CASES))
def test_oracle_reference_as_control_matches_reference_golden(name, built):
    """a10: mkChrFromFile + mkControlRead (ped.pl:1027-1136) -- the chromosome files, the tiled
    reference's sort_uniq/, count/, lbc/ and the edge list against ped.pl's own outputs
    (scripts/make_golden_ref.py)."""
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        g = json.load(f)
    fa, ft = make_ref_inputs(REF_CASES[name])
    chrs = oracle.ref_chromosomes(fa)
    assert {k: (len(v), hashlib.sha256(v).hexdigest()) for k, v in chrs.items()} == {
        k: (v["length"], v["sha256"]) for k, v in g["chromosomes"].items()}
    c, t, snp = oracle.run_ref(fa, ft, g["k"])
    for key, o in (("control", c), ("target", t)):
        sha, rows = _sha_tags(o.sort_uniq_text)
        assert (rows, sha) == (g[key]["sort_uniq_rows"], g[key]["sort_uniq_sha256"]), (key, "sort_uniq")
        sha, rows = _sha_tags(o.count_text)
        assert (rows, sha) == (g[key]["count_rows"], g[key]["count_sha256"]), (key, "count")
        sha, rows = _sha_tags(o.lbc_text)
        assert (rows, sha) == (g[key]["lbc_rows"], g[key]["lbc_sha256"]), (key, "lbc")
    assert snp.decode() == g["kmer_snp"]


def test_oracle_reference_parsing_corner_cases(built):
    """Header forms (ped.pl:1045-1050) and what the tiling does at the edges."""
    fa = b">chr007 x\nACGT\n>Chr3\nAC\n>7\nGGGG\n>\nTT\n> spaced\nCC\n>scaf_1|a b\nacgtnRYK\n"
    chrs = oracle.ref_chromosomes(fa)
    # chr007 -> 7, then ">7" truncates it; ">" and "> spaced" both name the file "chr"
    assert chrs == {"chr7": b"GGGG", "chr3": b"AC", "chr": b"CC", "chrscaf_1|a": b"ACGTNNNN"}
    s = oracle.OracleSample()
    s.ingest_reference(b">a\n" + b"ACGT" * 30 + b"\n", window=100, step=2)  # 120 bases: windows at 0..20
    assert s.reads_kept == 11
    s = oracle.OracleSample()
    s.ingest_reference(b">a\n" + b"ACGT" * 24 + b"\n")  # 96 bases: no window
    assert s.num_reads == 0


@pytest.mark.parametrize("name", sorted(REF_CASES))
def test_oracle_groundwork_for_position_mapping(name, built):
    """Groundwork for SURVEY 8(f) row f1 (next round; the product does not implement it yet): the
    oracle's unique-20-mer index of the reference (ped.pl:1138-1211) and its edge -> position map
    (ped.pl:641-685) against ped.pl's own ref20_uniq files and `sub=mapKmerSub` outputs."""
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        g = json.load(f)
    fa, ft = make_ref_inputs(REF_CASES[name])
    h = hashlib.sha256()
    rows = 0
    uniq = []
    for t in range(64):
        b = oracle.ref20_uniq_text(fa, t)
        uniq.append(b)
        h.update(b)
        rows += b.count(b"\n")
    assert (rows, h.hexdigest()) == (g["ref20_uniq"]["rows"], g["ref20_uniq"]["sha256"])
    by_tag = {t: b"" for t in TAGS}
    for line in g["kmer_snp"].encode().splitlines(keepends=True):
        by_tag[line[:3].decode()] += line
    got = {}
    for i, t in enumerate(TAGS):
        m = oracle.map_kmer_text(by_tag[t], uniq[i])
        if m:
            got[t] = m.decode()
    assert got == g["map"]


def test_cnv_restatement_against_the_reference(built):
    """f3: the line-by-line restatement of cnvJoin/cnvJoinSub (tests/pyref.py, ped.pl:934-989) over the
    oracle's count tables and unique-20-mer index reproduces ped.pl's own <t>.<c>.cnv
    (tests/golden/ref30k_cnv.json, made by scripts/make_golden_cnv.py from the unmodified reference)."""
    from tests import pyref
    from tests.cases import make_cnv_inputs

    with open(os.path.join(GOLDEN, "ref30k_cnv.json")) as f:
        g = json.load(f)
    fa, fc, ft = make_cnv_inputs(REF_CASES["ref30k"])
    oc, ot, _ = oracle.run(fc, ft, 20)
    ref20 = b"".join(oracle.ref20_uniq_text(fa, t) for t in range(64))
    text = pyref.cnv_text(ref20, oc.table(), ot.table())
    assert (text.count(b"\n"), hashlib.sha256(text).hexdigest()) == (g["cnv_rows"], g["cnv_sha256"])


def verification_inputs(fa, control, target, edges: bytes):
    """What SURVEY row f2 starts from, made with the (pinned) oracle: the position-map lines of all 64
    tags in `cat` order, the two sort_uniq line sets and the chromosome files."""
    by_tag = {t: b"" for t in TAGS}
    for line in edges.splitlines(keepends=True):
        by_tag[line[:3].decode()] += line
    map_text = b"".join(oracle.map_kmer_text(by_tag[t], oracle.ref20_uniq_text(fa, i)) for i, t in enumerate(TAGS))

    def lines(o):
        return {t: o.sort_uniq_text(i).decode().split("\n")[:-1] for i, t in enumerate(TAGS)}

    chromosomes = {n: s.decode() for n, s in oracle.ref_chromosomes(fa).items()}
    return map_text, lines(target), lines(control), chromosomes


@pytest.mark.parametrize("variant", ["control_reads", "control_ref"])
def test_verification_restatement_against_the_reference(variant, built):
    """f2: the line-by-line restatement of snpMkT/sortSeq/joinTarget/snpMkC/joinControl/kmerReadCount and of
    toVcf (oracle/verify_ref.py, ped.pl:2087-2305, 2576-2641, 2388-2500, 1532-1574) over the oracle's edge
    list, position map and sort_uniq sets reproduces ped.pl's own T.kmer.snp and T.kmer.vcf of a
    `method=kmer` run with a reference (tests/golden/ref30k_verify.json, scripts/make_golden_verify.py),
    with sequenced control reads (150 bp) and with the tiled reference as the control (100 bp)."""
    from oracle import verify_ref
    from tests.cases import make_cnv_inputs

    with open(os.path.join(GOLDEN, "ref30k_verify.json")) as f:
        g = json.load(f)[variant]
    fa, fc, ft = make_cnv_inputs(REF_CASES["ref30k"])
    oc, ot, edges = oracle.run(fc, ft, 20) if variant == "control_reads" else oracle.run_ref(fa, ft, 20)
    map_text, t_lines, c_lines, chromosomes = verification_inputs(fa, oc, ot, edges)
    snp = verify_ref.verify(map_text.decode().splitlines(), t_lines, c_lines, chromosomes)
    assert snp == g["kmer_snp"]
    assert verify_ref.to_vcf(snp, "T") == g["kmer_vcf"]
