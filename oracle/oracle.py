"""ctypes wrapper of oracle/ped_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product (ped_b200/, perl/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libped_oracle.so")
SRC_PATH = os.path.join(_HERE, "ped_oracle.c")

_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SRC_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libped_oracle.so"])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        P = C.c_void_p
        L.oracle_sample_new.restype = P
        L.oracle_sample_free.argtypes = [P]
        L.oracle_error.restype = C.c_char_p
        L.oracle_error.argtypes = [P]
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_get_threads.restype = C.c_int
        L.oracle_ingest.argtypes = [P, C.c_void_p, C.c_size_t, C.c_int]
        L.oracle_clipping_used.argtypes = [P]
        L.oracle_auto_clipped.argtypes = [P]
        L.oracle_dedup.argtypes = [P]
        L.oracle_num_reads.restype = C.c_size_t
        L.oracle_num_reads.argtypes = [P]
        L.oracle_reads_kept.restype = C.c_uint64
        L.oracle_reads_kept.argtypes = [P]
        L.oracle_read.restype = C.c_size_t
        L.oracle_read.argtypes = [P, C.c_size_t, C.POINTER(C.c_char_p)]
        L.oracle_count.argtypes = [P, C.c_int]
        L.oracle_num_kmers.restype = C.c_size_t
        L.oracle_num_kmers.argtypes = [P]
        L.oracle_num_instances.restype = C.c_uint64
        L.oracle_num_instances.argtypes = [P]
        L.oracle_kmers.restype = C.POINTER(C.c_uint64)
        L.oracle_kmers.argtypes = [P]
        L.oracle_counts.restype = C.POINTER(C.c_uint32)
        L.oracle_counts.argtypes = [P]
        for fn in (L.oracle_sort_uniq_text, L.oracle_count_text, L.oracle_lbc_text):
            fn.restype = P
            fn.argtypes = [P, C.c_int, C.POINTER(C.c_size_t)]
        L.oracle_join_text.restype = P
        L.oracle_join_text.argtypes = [P, P, C.POINTER(C.c_size_t)]
        L.oracle_free.argtypes = [P]
        L.oracle_ingest_reference.argtypes = [P, C.c_char_p, C.c_size_t, C.c_int, C.c_int]
        L.oracle_ref_parse.restype = P
        L.oracle_ref_parse.argtypes = [C.c_char_p, C.c_size_t]
        L.oracle_ref_num.restype = C.c_size_t
        L.oracle_ref_num.argtypes = [P]
        L.oracle_ref_name.restype = C.c_char_p
        L.oracle_ref_name.argtypes = [P, C.c_size_t]
        L.oracle_ref_seq.restype = C.c_size_t
        L.oracle_ref_seq.argtypes = [P, C.c_size_t, C.POINTER(C.c_char_p)]
        L.oracle_ref_free.argtypes = [P]
        L.oracle_ref20_uniq_text.restype = P
        L.oracle_ref20_uniq_text.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
        L.oracle_map_kmer_text.restype = P
        L.oracle_map_kmer_text.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        _lib = L
    return _lib


TAGS = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT"]


class OracleSample:
    """CPU restatement of ped.pl's per-sample stages (see ped_oracle.c header)."""

    def __init__(self, fastq=b"", clipping: int = 0):
        self.L = lib()
        self.h = self.L.oracle_sample_new()
        self.k = 0
        if len(fastq):
            self.ingest(fastq, clipping)

    def __del__(self):
        try:
            if self.h:
                self.L.oracle_sample_free(self.h)
                self.h = None
        except Exception:
            pass

    def ingest(self, fastq, clipping: int = 0) -> None:
        """fastq: bytes, or a numpy uint8 array (no copy: at-size inputs are many GB)."""
        if isinstance(fastq, np.ndarray):
            ptr, n = C.c_void_p(fastq.ctypes.data), fastq.size
        else:
            ptr, n = C.cast(C.c_char_p(fastq), C.c_void_p), len(fastq)
        rc = self.L.oracle_ingest(self.h, ptr, n, clipping)
        if rc:
            raise RuntimeError(self.L.oracle_error(self.h).decode())
        self.L.oracle_dedup(self.h)

    def ingest_reference(self, fasta: bytes, window: int = 0, step: int = 0) -> None:
        """mkChrFromFile + mkControlReadChr/Sub (ped.pl:1027-1060, 1089-1136)."""
        rc = self.L.oracle_ingest_reference(self.h, fasta, len(fasta), window, step)
        if rc:
            raise RuntimeError(self.L.oracle_error(self.h).decode())
        self.L.oracle_dedup(self.h)

    @property
    def clipping_used(self) -> int:
        return self.L.oracle_clipping_used(self.h)

    @property
    def auto_clipped(self) -> bool:
        return bool(self.L.oracle_auto_clipped(self.h))

    @property
    def num_reads(self) -> int:
        return self.L.oracle_num_reads(self.h)

    @property
    def reads_kept(self) -> int:
        return self.L.oracle_reads_kept(self.h)

    def reads(self) -> List[bytes]:
        out = []
        p = C.c_char_p()
        for i in range(self.num_reads):
            n = self.L.oracle_read(self.h, i, C.byref(p))
            out.append(C.string_at(p, n))
        return out

    def count(self, k: int = 20) -> None:
        rc = self.L.oracle_count(self.h, k)
        if rc:
            raise RuntimeError(self.L.oracle_error(self.h).decode())
        self.k = k

    @property
    def num_instances(self) -> int:
        return self.L.oracle_num_instances(self.h)

    def table(self, copy: bool = True) -> Tuple[np.ndarray, np.ndarray]:
        """copy=False: views into the oracle's own arrays (valid until the next count / free)."""
        n = self.L.oracle_num_kmers(self.h)
        if n == 0:
            return np.empty(0, np.uint64), np.empty(0, np.uint32)
        km = np.ctypeslib.as_array(self.L.oracle_kmers(self.h), shape=(n,))
        ct = np.ctypeslib.as_array(self.L.oracle_counts(self.h), shape=(n,))
        return (km.copy(), ct.copy()) if copy else (km, ct)

    def _text(self, fn, tag: int) -> bytes:
        n = C.c_size_t()
        p = fn(self.h, tag, C.byref(n))
        try:
            return C.string_at(p, n.value)
        finally:
            self.L.oracle_free(p)

    def sort_uniq_text(self, tag: int) -> bytes:
        return self._text(self.L.oracle_sort_uniq_text, tag)

    def count_text(self, tag: int) -> bytes:
        return self._text(self.L.oracle_count_text, tag)

    def lbc_text(self, tag: int) -> bytes:
        return self._text(self.L.oracle_lbc_text, tag)


def join_text(control: OracleSample, target: OracleSample) -> bytes:
    """`<target>.kmer.snp` of the no-ref k-mer mode (ped.pl:703-782, 374-378)."""
    L = lib()
    n = C.c_size_t()
    p = L.oracle_join_text(control.h, target.h, C.byref(n))
    try:
        return C.string_at(p, n.value)
    finally:
        L.oracle_free(p)


def run(control_fastq: bytes, target_fastq: bytes, k: int = 20, clipping: int = 0):
    """ped.pl main (291-297): the target is ingested first; `$clipping` is a global, so a
    value chosen automatically for the target is an explicit clipping= for the control."""
    t = OracleSample(target_fastq, clipping)
    if t.auto_clipped:
        clipping = t.clipping_used
    c = OracleSample(control_fastq, clipping)
    c.count(k)
    t.count(k)
    return c, t, join_text(c, t)


def ref_chromosomes(fasta: bytes):
    """The chromosome files mkChrFromFile leaves: {file name: cleaned sequence}."""
    L = lib()
    r = L.oracle_ref_parse(fasta, len(fasta))
    try:
        out = {}
        p = C.c_char_p()
        for i in range(L.oracle_ref_num(r)):
            n = L.oracle_ref_seq(r, i, C.byref(p))
            out[L.oracle_ref_name(r, i).decode()] = C.string_at(p, n) if n else b""
        return out
    finally:
        L.oracle_ref_free(r)


def run_ref(ref_fasta: bytes, target_fastq: bytes, k: int = 20, clipping: int = 0):
    """`ref=` run: the control is the reference tiled into reads (ped.pl:193, 1062-1136)."""
    t = OracleSample(target_fastq, clipping)
    c = OracleSample()
    c.ingest_reference(ref_fasta)
    c.count(k)
    t.count(k)
    return c, t, join_text(c, t)


# ---- groundwork for SURVEY 8(f) f1 (next round): oracle only, nothing in the product uses it ----

def ref20_uniq_text(ref_fasta: bytes, tag: int) -> bytes:
    """`zcat <ref>/ref20_uniq.TAG.gz` (ped.pl:1162-1211 mk20mer + 1138-1160 mkUniq)."""
    L = lib()
    n = C.c_size_t()
    p = L.oracle_ref20_uniq_text(ref_fasta, len(ref_fasta), tag, C.byref(n))
    try:
        return C.string_at(p, n.value)
    finally:
        L.oracle_free(p)


def map_kmer_text(snp_tag_text: bytes, ref20_uniq_tag_text: bytes) -> bytes:
    """`$tmpdir/<t>.map.TAG` (ped.pl:641-685 mapKmerSub) from `$tmpdir/<t>.snp.TAG` and ref20_uniq.TAG."""
    L = lib()
    n = C.c_size_t()
    p = L.oracle_map_kmer_text(snp_tag_text, len(snp_tag_text), ref20_uniq_tag_text, len(ref20_uniq_tag_text),
                               C.byref(n))
    try:
        return C.string_at(p, n.value)
    finally:
        L.oracle_free(p)


def set_threads(n: int) -> None:
    lib().oracle_set_threads(n)


def get_threads() -> int:
    return lib().oracle_get_threads()
