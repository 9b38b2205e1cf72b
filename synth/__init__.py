"""Synthetic genomes and FASTQ reads on the host: ctypes mirror of
include/pedk_synth.h over synth/libpedk_synth.so (source: synth/synth_host.cc).  Test and bench
infrastructure, kept outside the product package.  Plain C++ without CUDA, so the CPU arms of bench.py and
the golden scripts make their inputs without mapping the product library (libpedk.so)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpedk_synth.so")
_U64P = C.POINTER(C.c_uint64)
_P = C.c_void_p


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("genome_len", C.c_uint64), ("read_len", C.c_uint32), ("err_ppm", C.c_uint32),
                ("n_ppm", C.c_uint32), ("dup_ppm", C.c_uint32)]


SIGNATURES = {
    "pedk_synth_genome": (C.c_int, [C.c_uint64, C.c_uint64, _P]),
    "pedk_synth_mutate": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, _P, C.c_uint64, _U64P]),
    "pedk_synth_fastq_bytes": (C.c_uint64, [C.c_uint64, C.c_uint64, C.c_uint32]),
    "pedk_synth_fastq_host": (C.c_int, [C.POINTER(SynthParams), _P, C.c_uint64, C.c_uint64, _P]),
}

_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m ped_b200.build`")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def synth_genome(seed: int, length: int) -> np.ndarray:
    g = np.empty(length, dtype=np.uint8)
    rc = load_library().pedk_synth_genome(seed, length, _ptr(g))
    if rc:
        raise RuntimeError(f"pedk_synth_genome: {rc}")
    return g


def synth_mutate(genome: np.ndarray, seed: int, snp_ppm: int = 1000, indel_ppm: int = 100) -> np.ndarray:
    cap = len(genome) + len(genome) // 8 + 64
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_uint64()
    rc = load_library().pedk_synth_mutate(_ptr(genome), len(genome), seed, snp_ppm, indel_ppm, _ptr(out), cap,
                                          C.byref(n))
    if rc:
        raise RuntimeError(f"pedk_synth_mutate: {rc}")
    return out[: n.value].copy()


def synth_fastq_bytes(first: int, n_reads: int, read_len: int) -> int:
    return int(load_library().pedk_synth_fastq_bytes(first, n_reads, read_len))


def synth_fastq_array(genome: np.ndarray, seed: int, n_reads: int, read_len: int = 150, err_ppm: int = 2000,
                      n_ppm: int = 1000, dup_ppm: int = 0, first: int = 0) -> np.ndarray:
    p = SynthParams(seed, len(genome), read_len, err_ppm, n_ppm, dup_ppm)
    out = np.empty(synth_fastq_bytes(first, n_reads, read_len), dtype=np.uint8)
    rc = load_library().pedk_synth_fastq_host(C.byref(p), _ptr(genome), first, n_reads, _ptr(out))
    if rc:
        raise RuntimeError(f"pedk_synth_fastq_host: {rc}")
    return out


def synth_fastq(genome: np.ndarray, seed: int, n_reads: int, read_len: int = 150, err_ppm: int = 2000,
                n_ppm: int = 1000, dup_ppm: int = 0, first: int = 0) -> bytes:
    return synth_fastq_array(genome, seed, n_reads, read_len, err_ppm, n_ppm, dup_ppm, first).tobytes()
