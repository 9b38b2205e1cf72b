"""Parity of the CUDA path (through the C ABI) with the oracle and with the
reference's own outputs (tests/golden/).  Bit-exact: everything here is byte,
integer or index work.  B200 only."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle
from ped_b200 import api
from tests.cases import CASES, REF_CASES, make_inputs, make_ref_inputs, revcomp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu_run(ctx, fc: bytes, ft: bytes, k=20, clipping=0):
    c, t = api.Sample(ctx), api.Sample(ctx)
    c.add_fastq(fc)
    t.add_fastq(ft)
    ti = t.ingest(clipping)  # ped.pl:291-297: target first; an automatic clipping sticks for the control
    ci = c.ingest(ti.clipping_used if ti.auto_clipped else clipping)
    c.count(k)
    t.count(k)
    r = ctx.join(c, t)
    return c, t, ci, ti, r


def check_sample(gs: api.Sample, info, os_: oracle.OracleSample, texts=True):
    assert info.unique_reads == os_.num_reads
    assert info.clipping_used == os_.clipping_used
    assert bool(info.auto_clipped) == os_.auto_clipped
    km, ct = gs.table()
    okm, oct_ = os_.table()
    assert km.shape == okm.shape and (km == okm).all()
    assert (ct == oct_).all()
    # the order-independent digest bench.py's parity check relies on, taken from the unsorted rows
    assert gs.table_digest() == api.table_digest(okm, oct_)
    if texts:
        for tag in range(64):
            assert gs.sort_uniq_text(tag) == os_.sort_uniq_text(tag), ("sort_uniq", tag)
            assert gs.count_text(tag) == os_.count_text(tag), ("count", tag)
            assert gs.lbc_text(tag) == os_.lbc_text(tag), ("lbc", tag)


def _sha_tags(fn):
    h = hashlib.sha256()
    for t in range(64):
        h.update(fn(t))
    return h.hexdigest()


@pytest.mark.parametrize("name", ["tiny", "adversarial", "mixed", "clip100", "clip50", "dup", "c1_40x", "c1_1000x"])
def test_case_against_oracle_and_reference(ctx, name):
    case = CASES[name]
    fc, ft = make_inputs(case)
    clip = case.get("clipping", 0)
    oc, ot, osnp = oracle.run(fc, ft, 20, clip)
    c, t, ci, ti, r = gpu_run(ctx, fc, ft, 20, clip)
    try:
        big = name == "c1_1000x"
        check_sample(c, ci, oc, texts=not big)
        check_sample(t, ti, ot, texts=not big)
        assert r.text() == osnp
        e = r.edges()
        assert len(e) == osnp.count(b"\n") and (np.diff(e["prefix"].astype(np.int64)) >= 0).all()
        gpath = os.path.join(GOLDEN, name + ".json")
        if os.path.exists(gpath):
            g = json.load(open(gpath))
            assert r.text().decode() == g["kmer_snp"]
            for key, s in (("control", c), ("target", t)):
                assert _sha_tags(s.count_text) == g[key]["count_sha256"]
                assert _sha_tags(s.lbc_text) == g[key]["lbc_sha256"]
                assert _sha_tags(s.sort_uniq_text) == g[key]["sort_uniq_sha256"]
    finally:
        r.destroy(); c.destroy(); t.destroy()


@pytest.mark.parametrize("k", [4, 5, 16, 19, 21, 31, 32])
def test_other_k(ctx, k):
    case = dict(CASES["adversarial"])
    fc, ft = make_inputs(case)
    oc, ot, osnp = oracle.run(fc, ft, k)
    c, t, ci, ti, r = gpu_run(ctx, fc, ft, k)
    try:
        check_sample(c, ci, oc, texts=False)
        check_sample(t, ti, ot, texts=False)
        assert r.text() == osnp
    finally:
        r.destroy(); c.destroy(); t.destroy()


def _fq(seqs):
    return "".join(f"@r{i}\n{s}\n+\n{'I' * len(s)}\n" for i, s in enumerate(seqs)).encode()


def test_corner_inputs(ctx):
    import random

    rnd = random.Random(7)
    g = "".join(rnd.choice("ACGT") for _ in range(600))
    reads = [g[i:i + 64] for i in range(0, 530, 2)]
    half = g[100:132]
    pal = half + revcomp(half.encode()).decode()
    reads += [pal, pal, revcomp(reads[5].encode()).decode(), reads[7], "ACGTN" * 12 + "ACGT"]
    tgt = list(reads)
    tgt[30] = tgt[30][:20] + ("A" if tgt[30][20] != "A" else "C") + tgt[30][21:]
    fc, ft = _fq(reads), _fq(tgt)
    variants = [(fc, ft), (fc[:-1], ft[:-70]), (b"", ft), (fc, b""), (_fq(["NNNN" * 16]), ft), (b"@only\n", ft),
                (_fq(["ACGTACGTACGTACGT"]), _fq(["ACGTACGTACGTACGT"]))]  # reads shorter than k
    for a, b in variants:
        oc, ot, osnp = oracle.run(a, b, 20)
        c, t, ci, ti, r = gpu_run(ctx, a, b, 20)
        try:
            check_sample(c, ci, oc)
            check_sample(t, ti, ot)
            assert r.text() == osnp
        finally:
            r.destroy(); c.destroy(); t.destroy()


def test_alphabet_error_is_loud(ctx):
    s = api.Sample(ctx)
    s.add_fastq(_fq(["ACGTRYACGTACGTACGTACGTACGT"]))
    with pytest.raises(api.PedkError) as e:
        s.ingest()
    assert e.value.code == -5
    s.destroy()


def test_chunked_host_feed_and_fused_call(ctx):
    import torch

    case = CASES["c1_40x"]
    fc, ft = make_inputs(case)
    _, _, osnp = oracle.run(fc, ft, 20)
    # (a) FASTQ fed in arbitrary pieces
    c, t = api.Sample(ctx), api.Sample(ctx)
    for s, fq in ((c, fc), (t, ft)):
        for i in range(0, len(fq), 300_001):
            s.add_fastq(fq[i:i + 300_001])
        s.ingest()
        s.count(20)
    r = ctx.join(c, t)
    assert r.text() == osnp
    r.destroy(); c.destroy(); t.destroy()
    # (b) the one-call path with host buffers (pinned) and with device buffers
    hc = torch.frombuffer(bytearray(fc), dtype=torch.uint8).pin_memory()
    ht = torch.frombuffer(bytearray(ft), dtype=torch.uint8).pin_memory()
    r = ctx.kmer_edges(hc, len(fc), ht, len(ft), device_buffers=False)
    assert r.text() == osnp
    r.destroy()
    dc, dt = hc.cuda(), ht.cuda()
    torch.cuda.synchronize()
    r = ctx.kmer_edges(dc, len(fc), dt, len(ft), device_buffers=True)
    assert r.text() == osnp
    r.destroy()
    # ... and with padded device buffers that are read in place (twice: the text must survive)
    pc = torch.full((api.fastq_padded_bytes(len(fc)),), 0x41, dtype=torch.uint8, device="cuda")
    pt = torch.full((api.fastq_padded_bytes(len(ft)),), 0x0A, dtype=torch.uint8, device="cuda")
    pc[:len(fc)] = dc
    pt[:len(ft)] = dt
    torch.cuda.synchronize()
    for _ in range(2):
        r = ctx.kmer_edges(pc, len(fc), pt, len(ft), device_buffers=True, padded=True)
        assert r.text() == osnp
        r.destroy()
    assert bytes(pc[:len(fc)].cpu().numpy().tobytes()) == fc


@pytest.mark.parametrize("k", [20, 31])
def test_medium_scale_against_oracle(ctx, k):
    """5 Mb genome, 30x, 2 x 1 M reads = 300 Mbases (the survey's config-2 cut-down): table and edges vs
    the oracle, at the reference's k and at k = 31 (BASELINE config 3's other point: 64-bit instances)."""
    case = dict(kind="synth", cfg=12, genome=5_000_000, reads=1_000_000, L=150, snp_ppm=1000, indel_ppm=100,
                err_ppm=2000, n_ppm=1000, dup_ppm=0)
    fc, ft = make_inputs(case)
    oc, ot, osnp = oracle.run(fc, ft, k)
    c, t, ci, ti, r = gpu_run(ctx, fc, ft, k)
    try:
        check_sample(c, ci, oc, texts=False)
        check_sample(t, ti, ot, texts=False)
        assert r.text() == osnp
        assert r.num_edges > 1000
    finally:
        r.destroy(); c.destroy(); t.destroy()


@pytest.mark.parametrize("k,G,N", [(20, 20_000_000, 4_000_000), (31, 8_000_000, 1_500_000)])
def test_full_size_properties(ctx, k, G, N):
    """Size-independent properties at a size the oracle is not asked to follow:
    sortedness, strand symmetry, and conservation of instances.  k = 31 is the
    other point of BASELINE config 3's sweep and takes the 64-bit form of every bucket kernel."""
    import torch

    L = 150
    g = api.synth_genome(501, G)
    gd = torch.from_numpy(g).cuda()
    nb = api.synth_fastq_bytes(0, N, L)
    buf = torch.empty(nb, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.synth_fastq_device(api.SynthParams(502, G, L, 2000, 1000, 0), gd, 0, N, buf)
    s = api.Sample(ctx)
    s.add_fastq_device(buf, nb)
    info = s.ingest()
    ci = s.count(k)
    assert ci.instances * 2 - ci.correction_rows == info.unique_reads * (L - k + 1)
    km, ct = s.table()
    assert (km[1:] > km[:-1]).all()
    assert int(ct.astype(np.uint64).sum()) == info.unique_reads * (L - k + 1)
    # count(K) == count(revcomp K): reverse-complement every key and look it up
    rc = np.zeros_like(km)
    v = km.copy()
    for _ in range(k):
        rc = (rc << np.uint64(2)) | (np.uint64(3) - (v & np.uint64(3)))
        v >>= np.uint64(2)
    idx = np.searchsorted(km, rc)
    assert (km[idx] == rc).all() and (ct[idx] == ct).all()
    s.destroy()


@pytest.mark.parametrize("env", [{"PEDK_BUCKET_SLOTS_LOG2": "10"}, {"PEDK_BUCKET_SLOTS_LOG2": "10", "PEDK_BUCKET_B": "1"},
                                 {"PEDK_BUCKET_B": "2"}, {"PEDK_BUCKET_EXACT": "1"}, {"PEDK_COUNT_MODE": "sort"},
                                 {"PEDK_JOIN_SLOTS_LOG2": "10", "PEDK_JOIN_B": "1"}, {"PEDK_JOIN_B": "3"},
                                 {"PEDK_COUNT_PASSES": "4"}, {"PEDK_COUNT_PASSES": "2", "PEDK_BUCKET_EXACT": "1"},
                                 {"PEDK_DEVICE_BUDGET_MB": "8"}, {"PEDK_BC_THREADS": "1024"},
                                 {"PEDK_SUB_THREADS": "512"}, {"PEDK_BUCKET_B_EXACT": "1"}, {"PEDK_PART_BULK": "1"},
                                 {"PEDK_PART_BIG_TILE": "2"}, {"PEDK_SUB_ITEMS": "8"}, {"PEDK_SUB_ITEMS": "16"},
                                 {"PEDK_SUB_ITEMS": "8", "PEDK_SUB_THREADS": "512"}])
@pytest.mark.parametrize("k", [20, 31])
def test_bucket_table_overflow_and_legacy_path(ctx, env, k):
    """The per-bucket hash table is forced to overflow (1024 slots, buckets 128x larger) so that
    buckets are split by further hash bits and re-read (PEDK_JOIN_*: the same for the tables of the
    hash join, whose groups then meet in several rounds); PEDK_COUNT_PASSES / PEDK_DEVICE_BUDGET_MB
    count the instances in several groups of hash bins, as a sample that does not fit HBM is
    (the reference's `sort -T` spill); PEDK_BUCKET_EXACT=1 sizes the level-A bins
    with a counting pass instead of optimistically; PEDK_COUNT_MODE=sort is the extract +
    LSD radix sort + run-length path the buckets replaced.  All must give the reference's table."""
    case = dict(kind="synth", cfg=13, genome=300_000, reads=60_000, L=150, snp_ppm=1000, indel_ppm=100,
                err_ppm=2000, n_ppm=1000, dup_ppm=0)
    fc, ft = make_inputs(case)
    oc, ot, osnp = oracle.run(fc, ft, k)
    old = {n: os.environ.get(n) for n in env}
    os.environ.update(env)
    try:
        c, t, ci, ti, r = gpu_run(ctx, fc, ft, k)
        if "PEDK_COUNT_PASSES" in env:
            assert c.last_count.count_passes == int(env["PEDK_COUNT_PASSES"])
        if "PEDK_DEVICE_BUDGET_MB" in env:
            assert c.last_count.count_passes > 1
    finally:
        for n, v in old.items():
            if v is None:
                os.environ.pop(n, None)
            else:
                os.environ[n] = v
    try:
        check_sample(c, ci, oc, texts=False)
        check_sample(t, ti, ot, texts=False)
        assert r.text() == osnp
    finally:
        r.destroy(); c.destroy(); t.destroy()


def test_one_kmer_with_many_instances(ctx):
    """Low-complexity reads: a handful of k-mers carry almost all instances (one table slot each)."""
    import random

    rnd = random.Random(11)
    reads = []
    for i in range(3000):
        u = "".join(rnd.choice("ACGT") for _ in range(10))
        reads.append(u + "A" * 130 + "".join(rnd.choice("ACGT") for _ in range(10)))
        reads.append(u + "AC" * 65 + "".join(rnd.choice("ACGT") for _ in range(10)))
    fc = _fq(reads)
    ft = _fq(reads[: len(reads) // 2] + ["ACGT" * 37 + "AC"] * 3)
    oc, ot, osnp = oracle.run(fc, ft, 20)
    c, t, ci, ti, r = gpu_run(ctx, fc, ft, 20)
    try:
        check_sample(c, ci, oc, texts=False)
        check_sample(t, ti, ot, texts=False)
        assert r.text() == osnp
    finally:
        r.destroy(); c.destroy(); t.destroy()


def test_reference_as_control(ctx):
    """a10 (BASELINE config 5 shape): the control is the reference FASTA tiled into 100-base
    windows every 2 bases (ped.pl:1027-1136).  Chromosome files, tables, texts and the edge list
    against the oracle and against ped.pl's own outputs (tests/golden/ref30k.json)."""
    fa, ft = make_ref_inputs(REF_CASES["ref30k"])
    oc, ot, osnp = oracle.run_ref(fa, ft, 20)
    c, t = api.Sample(ctx), api.Sample(ctx)
    for i in range(0, len(fa), 7001):  # arbitrary chunking
        c.add_fastq(fa[i:i + 7001])
    t.add_fastq(ft)
    ti = t.ingest()
    ci = c.ingest_reference()
    c.count(20)
    t.count(20)
    r = ctx.join(c, t)
    try:
        want = oracle.ref_chromosomes(fa)
        got = {name: seq for name, seq, overwritten in c.chromosomes() if not overwritten}
        assert got == want
        assert any(ow for _, _, ow in c.chromosomes())  # the sequence whose name came back
        assert ci.reads_kept == oc.reads_kept and ci.clipping_used == 100
        check_sample(c, ci, oc)
        check_sample(t, ti, ot)
        assert r.text() == osnp
        g = json.load(open(os.path.join(GOLDEN, "ref30k.json")))
        assert r.text().decode() == g["kmer_snp"]
        assert {k: hashlib.sha256(v).hexdigest() for k, v in got.items()} == {
            k: v["sha256"] for k, v in g["chromosomes"].items()}
        for key, s in (("control", c), ("target", t)):
            assert _sha_tags(s.count_text) == g[key]["count_sha256"]
            assert _sha_tags(s.lbc_text) == g[key]["lbc_sha256"]
            assert _sha_tags(s.sort_uniq_text) == g[key]["sort_uniq_sha256"]
    finally:
        r.destroy(); c.destroy(); t.destroy()


def test_reference_corner_inputs(ctx):
    for fa in (b"", b"no header at all\nACGT\n", b">only\n", b">a\n" + b"ACGT" * 24 + b"\n",
               b">a\n" + b"ACGT" * 30 + b"\n>b\n" + b"ACGTN" * 40 + b"\n>a\n" + b"GATTACA" * 20,
               b">x\n" + b"A" * 300 + b"\n"):
        o = oracle.OracleSample()
        o.ingest_reference(fa)
        o.count(20)
        s = api.Sample(ctx)
        s.add_fastq(fa)
        info = s.ingest_reference()
        s.count(20)
        try:
            assert info.unique_reads == o.num_reads and info.reads_kept == o.reads_kept
            km, ct = s.table()
            okm, oct_ = o.table()
            assert km.shape == okm.shape and (km == okm).all() and (ct == oct_).all()
            for tag in range(64):
                assert s.sort_uniq_text(tag) == o.sort_uniq_text(tag)
        finally:
            s.destroy()


def test_reference_tiling_properties_at_scale(ctx):
    """20 Mb reference (config 5 cut down), sizes the oracle is not asked to follow: window
    arithmetic, conservation of k-mer instances and strand symmetry of the table."""
    G, k = 20_000_000, 20
    g = api.synth_genome(777, G).tobytes()
    fa = b">chr1\n" + b"\n".join(g[i:i + 80] for i in range(0, G // 2, 80)) + b"\n>chr2 x\n" + g[G // 2:] + b"\n"
    s = api.Sample(ctx)
    s.add_fastq(fa)
    info = s.ingest_reference()
    n_win = 2 * ((G // 2 - 100) // 2 + 1)
    assert info.records == n_win and info.reads_kept == n_win  # no N in a synthetic genome
    assert info.unique_reads <= 2 * n_win and info.unique_reads > 1.99 * n_win  # a random genome has no repeats
    ci = s.count(k)
    assert ci.instances * 2 - ci.correction_rows == info.unique_reads * (100 - k + 1)
    km, ct = s.table()
    assert (km[1:] > km[:-1]).all()
    assert int(ct.astype(np.uint64).sum()) == info.unique_reads * (100 - k + 1)
    rc = np.zeros_like(km)
    v = km.copy()
    for _ in range(k):
        rc = (rc << np.uint64(2)) | (np.uint64(3) - (v & np.uint64(3)))
        v >>= np.uint64(2)
    idx = np.searchsorted(km, rc)
    assert (km[idx] == rc).all() and (ct[idx] == ct).all()
    s.destroy()


def test_reference_unique_kmer_index_and_position_map(ctx):
    """f1: mk20mer + mkUniq (ped.pl:1138-1211) and mapKmerSub (ped.pl:641-685) against
    ped.pl's own ref20_uniq files and `sub=mapKmerSub` outputs (tests/golden/ref30k.json) and the oracle."""
    g = json.load(open(os.path.join(GOLDEN, "ref30k.json")))
    fa, ft = make_ref_inputs(REF_CASES["ref30k"])
    c, t = api.Sample(ctx), api.Sample(ctx)
    c.add_fastq(fa)
    t.add_fastq(ft)
    t.ingest()
    c.ingest_reference()
    rows = c.ref20_uniq(20)
    c.count(20)
    t.count(20)
    r = ctx.join(c, t)
    try:
        h = hashlib.sha256()
        for tag in range(64):
            b = c.ref20_uniq_text(tag)
            assert b == oracle.ref20_uniq_text(fa, tag), tag
            h.update(b)
        assert (rows, h.hexdigest()) == (g["ref20_uniq"]["rows"], g["ref20_uniq"]["sha256"])
        snp = r.text()
        assert snp.decode() == g["kmer_snp"]
        by_tag = {tg: b"" for tg in oracle.TAGS}
        for line in snp.splitlines(keepends=True):
            by_tag[line[:3].decode()] += line
        got = {}
        for tg in oracle.TAGS:
            m = c.map_kmer_text(by_tag[tg])
            if m:
                got[tg] = m.decode()
        assert got == g["map"]
        # the whole edge list at once maps to the concatenation of the per-tag files
        assert c.map_kmer_text(snp).decode() == "".join(g["map"].get(tg, "") for tg in oracle.TAGS)
    finally:
        r.destroy(); c.destroy(); t.destroy()


@pytest.mark.parametrize("fa", [b"", b">only\n", b">a\n" + b"ACGT" * 4 + b"\n", b">NOP\n" + b"ACGTTGCA" * 9 + b"\n",
                                b">a\n" + b"ACGTTGCAAT" * 7 + b"\n>b x\n" + b"GATTACAGATTACCA" * 5 + b"N" + b"ACGTTGCAAT" * 3,
                                b">p\nACGTTGCAATATTGCAACGT" + b"GGATC" * 9 + b"\n"])
def test_reference_unique_kmer_index_corner_inputs(ctx, fa):
    """Empty reference, a sequence shorter than k, a chromosome called NOP (skipped, ped.pl:1217), repeats
    within and between chromosomes, an N, a palindromic 20-mer (both of its lines carry the same k-mer)."""
    s = api.Sample(ctx)
    s.add_fastq(fa)
    s.ingest_reference()
    s.ref20_uniq(20)
    try:
        for tag in range(64):
            assert s.ref20_uniq_text(tag) == oracle.ref20_uniq_text(fa, tag), tag
    finally:
        s.destroy()


def test_cnv_against_the_reference(ctx):
    """f3: `method=cnv` (ped.pl:362-366, 934-989) -- the full outer join of the two count
    tables mapped through the unique 20-mers of the reference -- against ped.pl's own <t>.<c>.cnv
    (tests/golden/ref30k_cnv.json, scripts/make_golden_cnv.py) and a line-by-line restatement."""
    from tests import pyref
    from tests.cases import make_cnv_inputs

    g = json.load(open(os.path.join(GOLDEN, "ref30k_cnv.json")))
    fa, fc, ft = make_cnv_inputs(REF_CASES["ref30k"])
    r, c, t = api.Sample(ctx), api.Sample(ctx), api.Sample(ctx)
    try:
        r.add_fastq(fa)
        r.ingest_reference()
        r.ref20_uniq(20)
        for s, fq in ((t, ft), (c, fc)):
            s.add_fastq(fq)
            s.ingest()
            s.count(20)
        text = r.cnv_text(c, t)
        assert (text.count(b"\n"), hashlib.sha256(text).hexdigest()) == (g["cnv_rows"], g["cnv_sha256"])
        assert text.decode().startswith(g["cnv_head"]) and text.decode().endswith(g["cnv_tail"])
        ref20 = b"".join(r.ref20_uniq_text(tag) for tag in range(64))
        assert text == pyref.cnv_text(ref20, c.table(), t.table())
        # a k-mer only one sample has counts 0 in the other: the control against an empty sample
        e = api.Sample(ctx)
        e.add_fastq(b"")
        e.ingest()
        e.count(20)
        km, ct = c.table()
        assert r.cnv_text(c, e) == pyref.cnv_text(ref20, (km, ct), (km[:0], ct[:0]))
        e.destroy()
    finally:
        r.destroy(); c.destroy(); t.destroy()


def _verify_through_the_library(ctx, fa, fc, ft):
    """reference + control + target -> edges -> position map -> `<t>.kmer.snp` of a run with a reference,
    all through the C ABI; fc None: the tiled reference is the control (ped.pl:193)."""
    r, t = api.Sample(ctx), api.Sample(ctx)
    c = api.Sample(ctx) if fc is not None else r
    res = None
    try:
        r.add_fastq(fa)
        r.ingest_reference()
        r.ref20_uniq(20)
        t.add_fastq(ft)
        t.ingest()
        if fc is not None:
            c.add_fastq(fc)
            c.ingest()
        c.count(20)
        t.count(20)
        res = ctx.join(c, t)
        edges = res.text()
        map_text = r.map_kmer_text(edges)
        return edges, map_text, r.verify_text(c, t, map_text)
    finally:
        if res is not None:
            res.destroy()
        r.destroy(); t.destroy()
        if c is not r:
            c.destroy()


@pytest.mark.parametrize("variant", ["control_reads", "control_ref"])
def test_verification_against_the_reference(ctx, variant):
    """f2: snpMkT, sortSeq, joinTarget, snpMkC, joinControl, kmerReadCount (ped.pl:2087-2305,
    2576-2641, 2388-2500) and toVcf (1532-1574) against ped.pl's own T.kmer.snp / T.kmer.vcf of a
    `method=kmer` run with a reference (tests/golden/ref30k_verify.json, scripts/make_golden_verify.py):
    with sequenced control reads (both samples 150 bp) and with the tiled reference as the control
    (150 bp target, 100 bp control)."""
    from tests.cases import make_cnv_inputs

    g = json.load(open(os.path.join(GOLDEN, "ref30k_verify.json")))[variant]
    fa, fc, ft = make_cnv_inputs(REF_CASES["ref30k"])
    _, _, snp = _verify_through_the_library(ctx, fa, fc if variant == "control_reads" else None, ft)
    assert snp.decode() == g["kmer_snp"]
    assert api.vcf_text(snp, "T").decode() == g["kmer_vcf"]


@pytest.mark.parametrize("seed,L,CL", [(1, 100, 80), (2, 100, 100), (3, 64, 97), (4, 150, 33), (5, 33, 150), (6, 31, 32),
                                       (7, 129, 64), (8, 97, 161)])
@pytest.mark.parametrize("form", ["strings", "windows"])
def test_verification_against_the_restatement(ctx, seed, L, CL, form, monkeypatch):
    """f2 on a diploid target: heterozygous and homozygous SNPs, SNPs 3 / 25 / 80 bases apart (their windows
    carry each other, ped.pl:2184-2192), candidates closer to a chromosome end than a read is long, control
    reads of another length, F7 rows whose fields shift -- against the line-by-line restatement of the
    Perl (oracle/verify_ref.py, itself pinned to ped.pl's output) over the oracle's map and read sets."""
    from oracle import verify_ref
    from tests.cases import make_verify_inputs
    from tests.test_oracle import verification_inputs

    monkeypatch.setenv("PEDK_PROBE", form)   # both forms of the window look-up (ingest.cu)
    fa, fc, ft = make_verify_inputs(seed, L=L, CL=CL)
    edges, map_text, snp = _verify_through_the_library(ctx, fa, fc, ft)
    oc, ot, oedges = oracle.run(fc, ft, 20)
    assert edges == oedges
    omap, t_lines, c_lines, chromosomes = verification_inputs(fa, oc, ot, oedges)
    assert map_text == omap
    want = verify_ref.verify(omap.decode().splitlines(), t_lines, c_lines, chromosomes)
    assert snp.decode() == want
    kinds = set(line.split("\t")[-1] for line in want.splitlines())
    assert {"M", "R"} <= kinds
    assert api.vcf_text(snp, "T").decode() == verify_ref.to_vcf(want, "T")


def test_verification_corner_inputs(ctx):
    """f2: an empty map gives ped.pl's lone last-group line joined with nothing (an empty file); a map whose
    chromosome has no file, and a sample without reads (ped.pl dies in read(); here an error code)."""
    from tests.cases import make_verify_inputs

    fa, fc, ft = make_verify_inputs(7, L=50, CL=50, chrom_len=4200, coverage=10)
    r, c, t, e = api.Sample(ctx), api.Sample(ctx), api.Sample(ctx), api.Sample(ctx)
    try:
        r.add_fastq(fa); r.ingest_reference()
        c.add_fastq(fc); c.ingest()
        t.add_fastq(ft); t.ingest()
        e.add_fastq(b""); e.ingest()
        assert r.verify_text(c, t, b"") == b""
        line = b"nosuch\t500\tAAAAAAAAAAAAAAAAAAA\tC\tC\tG\tf\t0\t9\t0\t0\t0\t0\t9\t0\n"
        assert r.verify_text(c, t, line) == b""
        with pytest.raises(api.PedkError):
            r.verify_text(e, t, line)
    finally:
        r.destroy(); c.destroy(); t.destroy(); e.destroy()
