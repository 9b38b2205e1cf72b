"""CPU checks of the bit-level helpers shared by host and device code."""
import ctypes as C
import os
import random
import subprocess
import tempfile

import numpy as np
import pytest

from tests.cases import revcomp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


@pytest.fixture(scope="module")
def hd():
    d = tempfile.mkdtemp(prefix="pedk_hd_")
    so = os.path.join(d, "hd_shim.so")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "hd_shim.cpp"),
                           "-o", so])
    L = C.CDLL(so)
    L.hd_kmer_revcomp.restype = C.c_uint64
    L.hd_kmer_revcomp.argtypes = [C.c_uint64, C.c_int]
    L.hd_kmer_at.restype = C.c_uint64
    L.hd_kmer_at.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.hd_read_revcomp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.hd_words_cmp.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.hd_base_code.restype = C.c_uint32
    L.hd_base_code.argtypes = [C.c_uint32]
    return L


def enc(s: str) -> int:
    v = 0
    for c in s:
        v = (v << 2) | CODE[c]
    return v


def pack(s: str) -> np.ndarray:
    W = (len(s) + 31) // 32
    w = np.zeros(W, dtype=np.uint64)
    for i, c in enumerate(s):
        w[i >> 5] |= np.uint64(CODE[c] << (62 - 2 * (i & 31)))
    return w


def test_base_code(hd):
    for c in "ACGT":
        assert hd.hd_base_code(ord(c)) == CODE[c]


@pytest.mark.parametrize("k", [4, 19, 20, 21, 31, 32])
def test_kmer_revcomp(hd, k):
    rnd = random.Random(k)
    for _ in range(200):
        s = "".join(rnd.choice("ACGT") for _ in range(k))
        assert hd.hd_kmer_revcomp(enc(s), k) == enc(revcomp(s.encode()).decode())
    pal = "ACGTTGCAAT"[: k // 2]
    if k % 2 == 0 and len(pal) == k // 2:
        p = pal + revcomp(pal.encode()).decode()
        assert hd.hd_kmer_revcomp(enc(p), k) == enc(p)


@pytest.mark.parametrize("L", [3, 31, 32, 33, 64, 75, 100, 150, 151, 256, 300, 512])
def test_read_pack_revcomp_and_windows(hd, L):
    rnd = random.Random(L)
    for _ in range(20):
        s = "".join(rnd.choice("ACGT") for _ in range(L))
        w = pack(s)
        W = len(w)
        out = np.zeros(W, dtype=np.uint64)
        hd.hd_read_revcomp(w.ctypes.data, W, L, out.ctypes.data)
        r = revcomp(s.encode()).decode()
        assert (out == pack(r)).all()
        # numeric order of the word tuples == byte order of the strings
        want = (s > r) - (s < r)
        assert hd.hd_words_cmp(w.ctypes.data, out.ctypes.data, W) == want
        for k in (4, 20, 31, 32):
            if k > L:
                continue
            for i in {0, 1, 31, 32, 33, L - k, (L - k) // 2}:
                if 0 <= i <= L - k:
                    assert hd.hd_kmer_at(w.ctypes.data, W, i, k) == enc(s[i:i + k])


@pytest.mark.parametrize("k", [4, 5, 7, 8, 16, 19, 20, 21, 31, 32])
def test_kmer_hash_is_a_bijection_and_spreads(built, k):
    """kmer_hash (common.cuh) bins canonical k-mers by its top bits and keeps only the low bits:
    it must be invertible on 2k-bit integers, and its top bits must spread k-mers that share
    long prefixes (consecutive windows of a read) evenly."""
    from ped_b200 import api

    rnd = random.Random(k)
    n = 2 * k
    top = (1 << n) - 1
    xs = [0, 1, top, top - 1] + [rnd.getrandbits(n) for _ in range(2000)]
    for x in xs:
        h = api.kmer_hash(x, k)
        assert 0 <= h <= top
        assert api.kmer_unhash(h, k) == x
    if k <= 8:  # exhaustive: a permutation of the whole space
        hs = {api.kmer_hash(x, k) for x in range(1 << n)}
        assert len(hs) == 1 << n
    if k >= 16:
        # windows sliding over one random sequence: adjacent keys share k-1 bases
        seq = [rnd.randrange(4) for _ in range(20000 + k)]
        v = 0
        bins = np.zeros(256, dtype=np.int64)
        for i, c in enumerate(seq):
            v = ((v << 2) | c) & top
            if i >= k - 1:
                bins[api.kmer_hash(v, k) >> (n - 8)] += 1
        assert bins.min() > 0.5 * bins.mean() and bins.max() < 1.6 * bins.mean()
        # low-entropy input: keys that differ only in their low bits
        bins[:] = 0
        for x in range(20000):
            bins[api.kmer_hash(x, k) >> (n - 8)] += 1
        assert bins.min() > 0.5 * bins.mean() and bins.max() < 1.6 * bins.mean()


def test_kmer_owner_is_a_contiguous_range_of_bins(built):
    from ped_b200 import api

    rnd = random.Random(5)
    for k in (4, 20, 32):
        a = min(8, k)
        for world in (1, 2, 3, 4, 8):
            for _ in range(300):
                x = rnd.getrandbits(2 * k)
                b = api.kmer_hash(x, k) >> (2 * k - a)
                assert api.owner_of_kmer(x, k, world) == (b * world) >> a


def test_chromosome_file_names_match_the_oracle(built):
    """Host logic of the reference ingest (reference.cu chr_name_of) against the oracle's
    restatement of ped.pl:1045-1050, on header forms the golden case does not hold."""
    from oracle import oracle
    from ped_b200 import api

    headers = [b">chr1", b">1", b">001 leading zeros", b">chrX\tdesc", b">CHR07", b">Chr_x y", b">", b"> spaced",
               b">scaffold_12|len=5", b">chrchr2", b">0", b">000", b">12a", b">>two", b">a>b", b">chr", b">MT  mito"]
    for h in headers:
        want = list(oracle.ref_chromosomes(h + b"\nACGT\n").keys())
        assert [api.chromosome_file_name(h)] == want, h


def test_edge_pass1_truth_table_matches_the_rule_on_counts(built):
    """k_edges decides pass 1 of joinKmerSub (ped.pl:720-744) from 2-bit count classes.  For every one
    of the 65 536 class codes, pass 1 evaluated on actual counts (both ends of every class) must agree
    with the table -- i.e. the rule really only depends on the classes."""
    from ped_b200 import api

    bits = api.edge_pass1_table()
    reps = {0: (0, 1), 1: (2, 2), 2: (3, 100), 3: (101, 2_000_000_000)}

    def pass1(c, t):  # ped.pl:720-744 with cutoff 3
        rep = nohita = nohitb = pola = polb = 0
        a = b = 0
        for x in range(4):
            cv, tv = c[x], t[x]
            if cv > 100:
                rep += 1
            elif cv >= 3:
                a |= 1 << x
                if tv <= 1:
                    pola += 1
            elif cv <= 1:
                nohita += 1
            if tv > 100:
                rep += 1
            elif tv >= 3:
                b |= 1 << x
                if cv <= 1:
                    polb += 1
            elif tv <= 1:
                nohitb += 1
        return a != b and a and b and rep == 0 and nohita >= 2 and nohitb >= 2 and (pola == 1 or polb == 1)

    rnd = random.Random(1)
    n_pass = 0
    for code in range(65536):
        want = bool((int(bits[code >> 5]) >> (code & 31)) & 1)
        n_pass += want
        cls = [(code >> (2 * s)) & 3 for s in range(8)]
        for trial in range(3):
            pick = [reps[c][0] if trial == 0 else reps[c][1] if trial == 1 else rnd.choice(reps[c]) for c in cls]
            assert bool(pass1(pick[:4], pick[4:])) == want, (code, pick)
    assert 0 < n_pass < 65536 // 8
