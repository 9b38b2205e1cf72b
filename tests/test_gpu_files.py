"""The file-level entry points ped.pl binds (pedk_mk_sort_uniq / count_kmer /
lbc_kmer / join_kmer): same directories, names and `zcat` content as the
reference's file contract.  B200 only."""
import ctypes as C
import gzip
import json
import os

import pytest

from oracle import oracle
from ped_b200 import api
from tests.cases import CASES, REF_CASES, make_inputs, make_ref_inputs

pytestmark = pytest.mark.gpu
TAGS = oracle.TAGS
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _zcat(p):
    with gzip.open(p, "rb") as f:
        return f.read()


def _layout(tmp_path, fc, ft, gz=False):
    wd = str(tmp_path)
    for s, fq in (("C", fc), ("T", ft)):
        os.makedirs(os.path.join(wd, s, "read"))
        half = fq.find(b"\n@r", len(fq) // 2) + 1  # two files, split at a record boundary
        parts = [fq[:half], fq[half:]]
        for i, part in enumerate(parts):
            if gz:
                with gzip.open(os.path.join(wd, s, "read", f"{s}_{i + 1}.fastq.gz"), "wb") as f:
                    f.write(part)
            else:
                with open(os.path.join(wd, s, "read", f"{s}_{i + 1}.fastq"), "wb") as f:
                    f.write(part)
    return wd


def _check_files(wd, oc, ot, osnp):
    for s, o in (("C", oc), ("T", ot)):
        for i, tag in enumerate(TAGS):
            assert _zcat(f"{wd}/{s}/sort_uniq/{s}.sort_uniq.{tag}.gz") == o.sort_uniq_text(i)
            assert _zcat(f"{wd}/{s}/count/{s}.count.{tag}.gz") == o.count_text(i)
            assert _zcat(f"{wd}/{s}/lbc/{s}.lbc.{tag}.gz") == o.lbc_text(i)
    assert open(f"{wd}/T/T.kmer.snp", "rb").read() == osnp
    assert not [f for d, _, fs in os.walk(wd) for f in fs if f.endswith(".tmp")]


@pytest.mark.parametrize("gz", [False, True])
def test_drop_in_file_contract(built, tmp_path, gz):
    case = CASES["c1_40x"]
    fc, ft = make_inputs(case)
    oc, ot, osnp = oracle.run(fc, ft, 20)
    wd = _layout(tmp_path, fc, ft, gz)
    ctx = api.Context(0)
    lib = ctx.lib
    used = C.c_int()
    for s in ("T", "C"):
        ctx.check(lib.pedk_mk_sort_uniq(ctx.h, wd.encode(), s.encode(), 0, C.byref(used), None))
        assert used.value == 150
    for s in ("T", "C"):
        ctx.check(lib.pedk_count_kmer(ctx.h, wd.encode(), s.encode(), 20))
    for s in ("T", "C"):
        ctx.check(lib.pedk_lbc_kmer(ctx.h, wd.encode(), s.encode(), 20))
    ctx.check(lib.pedk_join_kmer(ctx.h, wd.encode(), b"T", b"C", 20, None, 0))
    ctx.close()
    _check_files(wd, oc, ot, osnp)
    g = json.load(open(os.path.join(GOLDEN, "c1_40x.json"))) if os.path.exists(os.path.join(GOLDEN, "c1_40x.json")) else None
    if g:
        assert open(f"{wd}/T/T.kmer.snp").read() == g["kmer_snp"]


def test_resume_from_files_in_fresh_processes(built, tmp_path):
    """Every stage in its own context: nothing but the files carries over
    (ped.pl's gates, ped.pl:291-297, 369-372), incl. the ref-mode per-tag files."""
    case = CASES["tiny"]
    fc, ft = make_inputs(case)
    oc, ot, osnp = oracle.run(fc, ft, 20)
    wd = _layout(tmp_path, fc, ft)

    def call(fn, *a):
        ctx = api.Context(0)
        try:
            ctx.check(getattr(ctx.lib, fn)(ctx.h, *a))
        finally:
            ctx.close()

    for s in (b"T", b"C"):
        call("pedk_mk_sort_uniq", wd.encode(), s, 0, None, None)
    for s in (b"T", b"C"):
        call("pedk_count_kmer", wd.encode(), s, 20)
    for s in (b"T", b"C"):
        call("pedk_lbc_kmer", wd.encode(), s, 20)
    call("pedk_join_kmer", wd.encode(), b"T", b"C", 20, None, 0)
    _check_files(wd, oc, ot, osnp)
    tmpdir = os.path.join(wd, "T", "tmp")
    call("pedk_join_kmer", wd.encode(), b"T", b"C", 20, tmpdir.encode(), 1)
    cat = b"".join(open(os.path.join(tmpdir, f"T.snp.{t}"), "rb").read() for t in TAGS)
    assert cat == osnp


def test_missing_input_is_an_error_not_a_gate_file(built, tmp_path):
    ctx = api.Context(0)
    rc = ctx.lib.pedk_mk_sort_uniq(ctx.h, str(tmp_path).encode(), b"nosuch", 0, None, None)
    assert rc == -4 and b"nosuch" in ctx.lib.pedk_last_error(ctx.h)
    assert not os.path.exists(os.path.join(str(tmp_path), "nosuch", "sort_uniq"))
    ctx.close()


@pytest.mark.parametrize("gz", [False, True])
def test_reference_as_control_file_contract(built, tmp_path, gz):
    """`perl ped.pl ref=REF,file=REF.fa` then `target=T,ref=REF,method=kmer`: pedk_mk_control_read
    leaves REF/chrNAME and REF/sort_uniq/ (ped.pl:1027-1136), the usual calls do the rest with the
    tiled reference as the control; the per-tag edge files go to tmpdir (have_ref, ped.pl:687-701)."""
    fa, ft = make_ref_inputs(REF_CASES["ref30k"])
    oc, ot, osnp = oracle.run_ref(fa, ft, 20)
    wd = str(tmp_path)
    os.makedirs(f"{wd}/REF")
    os.makedirs(f"{wd}/T/read")
    name = "REF.fa.gz" if gz else "REF.fa"
    with (gzip.open if gz else open)(f"{wd}/REF/{name}", "wb") as f:
        f.write(fa)
    open(f"{wd}/T/read/T.fastq", "wb").write(ft)
    ctx = api.Context(0)
    lib = ctx.lib
    ctx.check(lib.pedk_mk_control_read(ctx.h, wd.encode(), b"REF", name.encode()))
    ctx.check(lib.pedk_mk_sort_uniq(ctx.h, wd.encode(), b"T", 0, None, None))
    for s in (b"T", b"REF"):
        ctx.check(lib.pedk_count_kmer(ctx.h, wd.encode(), s, 20))
        ctx.check(lib.pedk_lbc_kmer(ctx.h, wd.encode(), s, 20))
    # mkRef's order (ped.pl:1018-1023): chromosome files, then mk20, then mkControlRead; here the
    # reference is resident after pedk_mk_control_read, so mk20 needs no second parse
    ctx.check(lib.pedk_mk_ref20_uniq(ctx.h, wd.encode(), b"REF", None, 20))
    tmpdir = f"{wd}/T/tmp"
    ctx.check(lib.pedk_join_kmer(ctx.h, wd.encode(), b"T", b"REF", 20, tmpdir.encode(), 1))
    ctx.check(lib.pedk_map_kmer(ctx.h, wd.encode(), b"T", b"REF", tmpdir.encode()))
    # f2 + toVcf (ped.pl:380-387): snpMkT ... kmerReadCount in one call, then the VCF
    ctx.check(lib.pedk_verify_kmer(ctx.h, wd.encode(), b"T", b"REF", b"REF", tmpdir.encode()))
    ctx.check(lib.pedk_to_vcf(ctx.h, wd.encode(), b"T"))
    assert lib.pedk_mk_control_read(ctx.h, wd.encode(), b"REF", b"nosuch.fa") == -4
    ctx.close()
    gv = json.load(open(os.path.join(GOLDEN, "ref30k_verify.json")))["control_ref"]
    assert open(f"{wd}/T/T.kmer.snp").read() == gv["kmer_snp"]
    assert open(f"{wd}/T/T.kmer.vcf").read() == gv["kmer_vcf"]
    os.unlink(f"{wd}/T/T.kmer.snp")
    os.unlink(f"{wd}/T/T.kmer.vcf")
    # f1: <ref>/ref20_uniq.TAG.gz and <tmpdir>/<t>.map.TAG as ped.pl leaves them
    g = json.load(open(os.path.join(GOLDEN, "ref30k.json")))
    for i, tag in enumerate(TAGS):
        assert _zcat(f"{wd}/REF/ref20_uniq.{tag}.gz") == oracle.ref20_uniq_text(fa, i)
        assert open(f"{tmpdir}/T.map.{tag}").read() == g["map"].get(tag, "")
    # ... and again in a fresh process that only has the files (FASTA named: parsed again; map from gz)
    for tag in TAGS:
        os.unlink(f"{wd}/REF/ref20_uniq.{tag}.gz")
        os.unlink(f"{tmpdir}/T.map.{tag}")
    ctx = api.Context(0)
    ctx.check(ctx.lib.pedk_mk_ref20_uniq(ctx.h, wd.encode(), b"REF", name.encode(), 20))
    ctx.close()
    ctx = api.Context(0)
    ctx.check(ctx.lib.pedk_map_kmer(ctx.h, wd.encode(), b"T", b"REF", tmpdir.encode()))
    ctx.close()
    ctx = api.Context(0)   # the verification from the sort_uniq, map and chromosome files alone
    ctx.check(ctx.lib.pedk_verify_kmer(ctx.h, wd.encode(), b"T", b"REF", b"REF", tmpdir.encode()))
    ctx.check(ctx.lib.pedk_to_vcf(ctx.h, wd.encode(), b"T"))
    ctx.close()
    assert open(f"{wd}/T/T.kmer.snp").read() == gv["kmer_snp"]
    assert open(f"{wd}/T/T.kmer.vcf").read() == gv["kmer_vcf"]
    for i, tag in enumerate(TAGS):
        assert _zcat(f"{wd}/REF/ref20_uniq.{tag}.gz") == oracle.ref20_uniq_text(fa, i)
        assert open(f"{tmpdir}/T.map.{tag}").read() == g["map"].get(tag, "")
    for fn, seq in oracle.ref_chromosomes(fa).items():
        assert open(f"{wd}/REF/{fn}", "rb").read() == seq
    for s, o in (("REF", oc), ("T", ot)):
        for i, tag in enumerate(TAGS):
            assert _zcat(f"{wd}/{s}/sort_uniq/{s}.sort_uniq.{tag}.gz") == o.sort_uniq_text(i)
            assert _zcat(f"{wd}/{s}/count/{s}.count.{tag}.gz") == o.count_text(i)
            assert _zcat(f"{wd}/{s}/lbc/{s}.lbc.{tag}.gz") == o.lbc_text(i)
    assert b"".join(open(f"{tmpdir}/T.snp.{t}", "rb").read() for t in TAGS) == osnp
    assert osnp.decode() == g["kmer_snp"]


def test_cnv_file_contract(built, tmp_path):
    """`perl ped.pl ref=REF,file=REF.fa` then `target=T,control=C,ref=REF,method=cnv` (ped.pl:362-366, 934-989):
    <wd>/T/T.C.cnv as the reference leaves it (tests/golden/ref30k_cnv.json), once with everything resident
    and once in a fresh process that only has the count files and the ref20_uniq files."""
    import hashlib

    from tests.cases import make_cnv_inputs

    g = json.load(open(os.path.join(GOLDEN, "ref30k_cnv.json")))
    fa, fc, ft = make_cnv_inputs(REF_CASES["ref30k"])
    wd = str(tmp_path)
    os.makedirs(f"{wd}/REF")
    open(f"{wd}/REF/REF.fa", "wb").write(fa)
    for s, fq in (("T", ft), ("C", fc)):
        os.makedirs(f"{wd}/{s}/read")
        open(f"{wd}/{s}/read/{s}.fastq", "wb").write(fq)
    ctx = api.Context(0)
    lib = ctx.lib
    ctx.check(lib.pedk_mk_control_read(ctx.h, wd.encode(), b"REF", b"REF.fa"))
    ctx.check(lib.pedk_mk_ref20_uniq(ctx.h, wd.encode(), b"REF", None, 20))
    for s in (b"T", b"C"):
        ctx.check(lib.pedk_mk_sort_uniq(ctx.h, wd.encode(), s, 0, None, None))
        ctx.check(lib.pedk_count_kmer(ctx.h, wd.encode(), s, 20))
    ctx.check(lib.pedk_cnv(ctx.h, wd.encode(), b"T", b"C", b"REF", 20))
    ctx.close()
    b = open(f"{wd}/T/T.C.cnv", "rb").read()
    assert (b.count(b"\n"), hashlib.sha256(b).hexdigest()) == (g["cnv_rows"], g["cnv_sha256"])
    os.unlink(f"{wd}/T/T.C.cnv")
    ctx = api.Context(0)
    ctx.check(ctx.lib.pedk_cnv(ctx.h, wd.encode(), b"T", b"C", b"REF", 20))
    ctx.close()
    assert open(f"{wd}/T/T.C.cnv", "rb").read() == b


def test_perl_drop_in_command_line(built, tmp_path):
    """The XS binding (perl/PedK.xs -> perl/lib/auto/PedK/PedK.so) and ped.pl's command line on top of it:
    `perl perl/ped_kmer.pl target=T,control=C,wd=WD` leaves every file ped.pl leaves (zcat-identical), and the
    ref= forms leave ref20_uniq, the per-tag edge files and the position map."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["sh", os.path.join(root, "perl", "build.sh")], check=True)
    case = CASES["c1_40x"]
    fc, ft = make_inputs(case)
    oc, ot, osnp = oracle.run(fc, ft, 20)
    wd = _layout(tmp_path, fc, ft)
    r = subprocess.run(["perl", os.path.join(root, "perl", "ped_kmer.pl"), f"target=T,control=C,wd={wd}"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    _check_files(wd, oc, ot, osnp)
    assert "clipping : 150" in open(f"{wd}/T/T.log").read()
    # reference mode: `ref=REF,file=REF.fa`, then `target=T,ref=REF`
    fa, ft2 = make_ref_inputs(REF_CASES["ref30k"])
    g = json.load(open(os.path.join(GOLDEN, "ref30k.json")))
    wd2 = os.path.join(str(tmp_path), "refmode")
    os.makedirs(f"{wd2}/REF")
    os.makedirs(f"{wd2}/T/read")
    open(f"{wd2}/REF/REF.fa", "wb").write(fa)
    open(f"{wd2}/T/read/T.fastq", "wb").write(ft2)
    for args in (f"ref=REF,file=REF.fa,wd={wd2}", f"target=T,ref=REF,wd={wd2}"):
        r = subprocess.run(["perl", os.path.join(root, "perl", "ped_kmer.pl"), args], capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for i, tag in enumerate(TAGS):
        assert _zcat(f"{wd2}/REF/ref20_uniq.{tag}.gz") == oracle.ref20_uniq_text(fa, i)
        assert open(f"{wd2}/T/tmp/T.map.{tag}").read() == g["map"].get(tag, "")
    assert "".join(open(f"{wd2}/T/tmp/T.snp.{t}").read() for t in TAGS) == g["kmer_snp"]
    # ... and the verified, genotyped candidates and their VCF (ped.pl:380-387)
    gv = json.load(open(os.path.join(GOLDEN, "ref30k_verify.json")))["control_ref"]
    assert open(f"{wd2}/T/T.kmer.snp").read() == gv["kmer_snp"]
    assert open(f"{wd2}/T/T.kmer.vcf").read() == gv["kmer_vcf"]
