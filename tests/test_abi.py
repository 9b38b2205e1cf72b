"""The C-ABI library loads and exports every symbol include/pedk.h declares
(no compute calls: this runs without a GPU)."""
import os
import re

import pytest

from ped_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header="pedk.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(pedk_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_exported(built):
    lib = api.load_library()
    names = declared_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"libpedk.so does not export {n}"


def test_python_mirror_covers_header(built):
    assert sorted(api.SIGNATURES) == declared_functions()


def test_synth_library_is_separate(built):
    """The host generator lives in libpedk_synth.so (include/pedk_synth.h): a CPU arm that makes its
    inputs with it never maps the product library."""
    import synth

    lib = synth.load_library()
    names = declared_functions("pedk_synth.h")
    assert sorted(synth.SIGNATURES) == names
    for n in names:
        assert hasattr(lib, n)
    product = api.load_library()
    for n in names:
        assert not hasattr(product, n), f"libpedk.so still exports {n}"


def test_struct_layouts_match_header(built):
    import ctypes as C

    assert C.sizeof(api.Edge) == 48 and api.EDGE_DTYPE.itemsize == 48
    assert C.sizeof(api.IngestInfo) == 64
    assert C.sizeof(api.CountInfo) == 40
    assert C.sizeof(api.SynthParams) == 32


def test_no_cpu_fallback(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.PedkError) as e:
        api.Context(0)
    assert e.value.code == -9  # PEDK_E_NODEVICE


def test_product_never_imports_oracle():
    bad = []
    for d in ("ped_b200", "perl", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, d)):
            if "build" in dp.split(os.sep):
                continue
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c", ".pm", ".xs", ".pl")):
                    s = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"ped_oracle|from oracle|import oracle|oracle/", s):
                        bad.append(os.path.join(dp, f))
    assert bad == [ROOT + "/ped_b200/build.py"] or bad == []  # build.py only compiles the checker


def test_synth_slices_are_independent(built):
    g = api.synth_genome(11, 2000)
    whole = api.synth_fastq(g, 5, 230, read_len=100, dup_ppm=200000)
    a = api.synth_fastq(g, 5, 100, read_len=100, dup_ppm=200000)
    b = api.synth_fastq(g, 5, 130, read_len=100, dup_ppm=200000, first=100)
    assert a + b == whole
    assert len(whole) == api.synth_fastq_bytes(0, 230, 100)
    t = api.synth_mutate(g, 3, 5000, 1000)
    assert len(t) != 0 and bytes(t) != bytes(g)


def test_perl_binding_builds_and_loads(built):
    """perl/PedK.xs compiles against include/pedk.h and the module loads (no device call: abi_version only)."""
    import shutil
    import subprocess

    if not shutil.which("xsubpp"):
        pytest.skip("no xsubpp")
    subprocess.run(["sh", os.path.join(ROOT, "perl", "build.sh")], check=True)
    r = subprocess.run(["perl", "-I" + os.path.join(ROOT, "perl", "lib"), "-MPedK", "-e",
                        "print PedK::abi_version(), ' ', join(',', map { defined &{\"PedK::$_\"} ? 1 : 0 } "
                        "qw(mk_sort_uniq mk_control_read mk_ref20_uniq count_kmer lbc_kmer join_kmer map_kmer verify_kmer to_vcf cnv))"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.split()[0] == "1" and r.stdout.split()[1] == "1,1,1,1,1,1,1,1,1,1"


def test_vcf_writer_is_host_only_and_matches_the_reference(built):
    """pedk_vcf_text (toVcf, ped.pl:1532-1574) makes no CUDA call, so it runs here: against ped.pl's own
    T.kmer.vcf (tests/golden/ref30k_verify.json) and, on made-up rows that exercise the quirks (rows whose
    fields shifted, an M or H anywhere in the line, a line that repeats, depth 0, names longer than three
    characters), against the line-by-line restatement (oracle/verify_ref.py)."""
    import json
    import random

    from oracle import verify_ref

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref30k_verify.json")))
    for v in ("control_reads", "control_ref"):
        assert api.vcf_text(g[v]["kmer_snp"].encode(), "T").decode() == g[v]["kmer_vcf"]
    rnd = random.Random(11)
    lines = []
    for i in range(400):
        chrom = rnd.choice(["1", "12", "scaffoldM", "chrH", "X", "0007", "MT"])
        f = [chrom, str(rnd.randrange(1, 10 ** 6)), "ACGTACGTACGTACGTACG", "A", "A", rnd.choice(["C", "CT", "G"]),
             rnd.choice("fr")] + [str(rnd.randrange(0, 60)) for _ in range(8)]
        f += [str(rnd.randrange(0, 40)), str(rnd.randrange(0, 3)), str(rnd.choice([0, 0, 1, 17, 333])),
              str(rnd.choice([0, 0, 2, 29, 700])), rnd.choice("MHRN")]
        if rnd.random() < 0.15:            # an F7 row: a field is missing, the rest moved left, empty fields behind
            del f[rnd.randrange(7, 15)]
            f += [""] * 3
        lines.append("\t".join(f) + "\n")
        if rnd.random() < 0.1:
            lines.append(lines[-1])
    text = "".join(lines)
    assert api.vcf_text(text.encode(), "sample_1").decode() == verify_ref.to_vcf(text, "sample_1")
    assert api.vcf_text(b"", "T").decode() == verify_ref.to_vcf("", "T")
