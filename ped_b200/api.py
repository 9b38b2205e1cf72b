"""ctypes mirror of include/pedk.h.

The host-side interface of the k-mer hot path follows ped.pl's own subs
(reference: ped.pl mkSortUniq 1365-1410, countKmer 784-815, lbcKmer 817-831,
joinKmer 687-701, kmer 368-378): a `Sample` is ingested (sort_uniq), counted
(count), and two samples are joined into the edge list (`<target>.kmer.snp`).
Everything is computed by libpedk.so on a B200; importing this module without
the built library, or opening a context without a GPU, raises -- there is no
CPU path in the product.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpedk.so")

# synthetic inputs (tests and bench.py): the host generator is test infrastructure outside this package
# (synth/, its own libpedk_synth.so); re-exported here so that test code reads api.synth_*
import sys as _sys  # noqa: E402

if os.path.dirname(_HERE) not in _sys.path:
    _sys.path.insert(0, os.path.dirname(_HERE))
from synth import SynthParams, synth_fastq, synth_fastq_bytes, synth_genome, synth_mutate  # noqa: E402,F401

PEDK_N_STAGES = 10
STAGES = ["ingest", "dedup", "extract", "sort", "count", "expand", "rowsort", "edges", "copy", "exchange"]
KERNEL_CLASSES = ["radix_keys", "radix_pairs", "part_scatter", "sub_hist", "sub_scatter", "bucket_count", "pack",
                  "part_count", "join_count", "join_scatter", "join_sub", "join_table"]
COMM_ID_BYTES = 128


class PedkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pedk error {code}: {msg}")
        self.code = code


class IngestInfo(C.Structure):
    _fields_ = [("fastq_bytes", C.c_uint64), ("records", C.c_uint64), ("reads_kept", C.c_uint64),
                ("strand_reads", C.c_uint64), ("unique_reads", C.c_uint64), ("palindromes", C.c_uint64),
                ("clipping_used", C.c_int32), ("auto_clipped", C.c_int32), ("n_groups", C.c_int32),
                ("pad", C.c_int32)]


class CountInfo(C.Structure):
    _fields_ = [("k", C.c_int32), ("sort_passes", C.c_int32), ("instances", C.c_uint64),
                ("canonical_kmers", C.c_uint64), ("correction_rows", C.c_uint64), ("count_passes", C.c_int32),
                ("pad", C.c_int32)]


class Edge(C.Structure):
    _fields_ = [("prefix", C.c_uint64), ("control", C.c_uint32 * 4), ("target", C.c_uint32 * 4),
                ("cont_mask", C.c_uint8), ("alt_mask", C.c_uint8), ("quirk", C.c_uint8), ("pad", C.c_uint8 * 5)]


EDGE_DTYPE = np.dtype([("prefix", "<u8"), ("control", "<u4", (4,)), ("target", "<u4", (4,)), ("cont_mask", "u1"),
                       ("alt_mask", "u1"), ("quirk", "u1"), ("pad", "u1", (5,))])


class Stats(C.Structure):
    _fields_ = [("stage_ms", C.c_double * PEDK_N_STAGES), ("stage_launches", C.c_uint64 * PEDK_N_STAGES),
                ("kernel_launches", C.c_uint64), ("sort_pass_launches", C.c_uint64), ("sort_pass_keys", C.c_uint64),
                ("sort_pass_ms", C.c_double), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("peak_device_bytes", C.c_uint64), ("kc_ms", C.c_double * 16), ("kc_bytes", C.c_uint64 * 16),
                ("kc_launches", C.c_uint64 * 16)]

    def as_dict(self) -> dict:
        return {
            "stage_ms": {STAGES[i]: self.stage_ms[i] for i in range(PEDK_N_STAGES)},
            "stage_launches": {STAGES[i]: int(self.stage_launches[i]) for i in range(PEDK_N_STAGES)},
            "kernel_launches": int(self.kernel_launches),
            "sort_pass_launches": int(self.sort_pass_launches),
            "sort_pass_keys": int(self.sort_pass_keys),
            "sort_pass_ms": float(self.sort_pass_ms),
            "h2d_bytes": int(self.h2d_bytes),
            "d2h_bytes": int(self.d2h_bytes),
            "peak_device_bytes": int(self.peak_device_bytes),
            "kernels": {KERNEL_CLASSES[i]: {"ms": float(self.kc_ms[i]), "bytes": int(self.kc_bytes[i]),
                                            "launches": int(self.kc_launches[i])} for i in range(len(KERNEL_CLASSES))},
        }




# every symbol include/pedk.h declares: (restype, argtypes)
_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_U64P = C.POINTER(C.c_uint64)
SIGNATURES = {
    "pedk_open": (C.c_int, [_PP, C.POINTER(C.c_int), C.c_int]),
    "pedk_close": (C.c_int, [_P]),
    "pedk_last_error": (C.c_char_p, [_P]),
    "pedk_strerror": (C.c_char_p, [C.c_int]),
    "pedk_abi_version": (C.c_int, []),
    "pedk_stream": (C.c_void_p, [_P]),
    "pedk_sample_create": (C.c_int, [_P, _PP]),
    "pedk_sample_destroy": (C.c_int, [_P]),
    "pedk_sample_reserve": (C.c_int, [_P, C.c_uint64]),
    "pedk_sample_add_fastq": (C.c_int, [_P, _P, C.c_uint64]),
    "pedk_sample_add_fastq_device": (C.c_int, [_P, _P, C.c_uint64]),
    "pedk_sample_ingest": (C.c_int, [_P, C.c_int, C.POINTER(IngestInfo)]),
    "pedk_sample_borrow_fastq_device": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64]),
    "pedk_sample_count": (C.c_int, [_P, C.c_int, C.POINTER(CountInfo)]),
    "pedk_sample_table": (C.c_int, [_P, _P, _P, C.c_uint64, _U64P]),
    "pedk_sample_ref20_uniq": (C.c_int, [_P, C.c_int, _U64P]),
    "pedk_sample_ref20_uniq_text": (C.c_int, [_P, C.c_int, _PP, _U64P]),
    "pedk_sample_map_kmer_text": (C.c_int, [_P, C.c_char_p, C.c_uint64, _PP, _U64P]),
    "pedk_cnv_text": (C.c_int, [_P, _P, _P, _PP, _U64P]),
    "pedk_cnv": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    "pedk_verify_text": (C.c_int, [_P, _P, _P, C.c_char_p, C.c_uint64, _PP, _U64P]),
    "pedk_vcf_text": (C.c_int, [C.c_char_p, C.c_uint64, C.c_char_p, _PP, _U64P]),
    "pedk_verify_strings_text": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _U64P,
                                           C.c_int, C.c_int, _PP, _U64P]),
    "pedk_verify_counts_text": (C.c_int, [C.c_char_p, C.c_uint64, _P, _P, C.c_uint64, C.c_int, C.c_int, _PP, _U64P]),
    "pedk_verify_kmer": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]),
    "pedk_to_vcf": (C.c_int, [_P, C.c_char_p, C.c_char_p]),
    "pedk_mk_ref20_uniq": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    "pedk_map_kmer": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]),
    "pedk_sample_table_digest": (C.c_int, [_P, _U64P, _U64P]),
    "pedk_sample_reads": (C.c_int, [_P, C.c_int, _P, _P, C.c_uint64, _U64P, C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int32)]),
    "pedk_sample_release_reads": (C.c_int, [_P]),
    "pedk_join": (C.c_int, [_P, _P, _P, _PP]),
    "pedk_result_num_edges": (C.c_uint64, [_P]),
    "pedk_result_text": (C.c_int, [_P, C.POINTER(C.c_char_p), _U64P]),
    "pedk_result_edges": (C.c_int, [_P, _P, C.c_uint64, _U64P]),
    "pedk_sample_lbc_text": (C.c_int, [_P, C.c_int, _PP, _U64P]),
    "pedk_sample_count_text": (C.c_int, [_P, C.c_int, _PP, _U64P]),
    "pedk_sample_sort_uniq_text": (C.c_int, [_P, C.c_int, _PP, _U64P]),
    "pedk_free": (None, [_P]),
    "pedk_result_destroy": (C.c_int, [_P]),
    "pedk_kmer_edges": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint64, C.c_int, C.c_int, C.c_int, _PP]),
    "pedk_kmer_edges_ref": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint64, C.c_int, C.c_int, C.c_int, _PP]),
    "pedk_sample_ingest_reference": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(IngestInfo)]),
    "pedk_sample_num_chromosomes": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "pedk_sample_chromosome": (C.c_int, [_P, C.c_int, _P, C.c_uint64, _U64P, C.POINTER(C.c_int), _P, C.c_uint64]),
    "pedk_edge_pass1_table": (C.c_int, [_P]),
    "pedk_chromosome_file_name": (C.c_int, [C.c_char_p, _P, C.c_uint64]),
    "pedk_mk_control_read": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p]),
    "pedk_mk_sort_uniq": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pedk_count_kmer": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_int]),
    "pedk_lbc_kmer": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_int]),
    "pedk_join_kmer": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]),
    "pedk_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "pedk_guard_hits": (C.c_uint64, [_P]),
    "pedk_reset_stats": (C.c_int, [_P]),
    "pedk_comm_unique_id": (C.c_int, [_P, _P]),
    "pedk_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "pedk_owner_of_read": (C.c_int, [_P, C.c_int, C.c_int]),
    "pedk_owner_of_kmer": (C.c_int, [C.c_uint64, C.c_int, C.c_int]),
    "pedk_kmer_hash": (C.c_uint64, [C.c_uint64, C.c_int]),
    "pedk_kmer_unhash": (C.c_uint64, [C.c_uint64, C.c_int]),
    "pedk_plan_bucket_owners": (C.c_int, [_P, C.c_int, _P]),
    "pedk_plan_level_a": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint64, _P, C.c_int, _P, _P, _P, C.POINTER(C.c_int),
                                    _U64P, _U64P]),
    "pedk_synth_fastq_device": (C.c_int, [_P, C.POINTER(SynthParams), _P, C.c_uint64, C.c_uint64, _P]),
    "pedk_plan_level_a_pass": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint64, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P,
                                         C.POINTER(C.c_int), _U64P, _U64P]),
    "pedk_test_sort_u64": (C.c_int, [_P, _P, C.c_uint64, C.c_int]),
    "pedk_test_sort_pairs": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int]),
    "pedk_test_rle": (C.c_int, [_P, _P, C.c_uint64, _P, _P, _U64P]),
}

_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load libpedk.so (built by `python -m ped_b200.build`).  Fails loudly."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(f"{p} is missing: build it with `python -m ped_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _ptr(x) -> C.c_void_p:
    """Raw address of a numpy array, bytes object, int or torch tensor."""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    if isinstance(x, (bytes, bytearray)):
        return C.cast(C.c_char_p(bytes(x)) if isinstance(x, bytearray) else C.c_char_p(x), C.c_void_p)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    raise TypeError(type(x))


class Context:
    """One CUDA device.  Replaces ped.pl's process pool (ped.pl:244-255, 3504-3543)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        dev = (C.c_int * 1)(device)
        rc = self.lib.pedk_open(C.byref(self.h), dev, 1)
        if rc != 0:
            msg = self.lib.pedk_last_error(None).decode()
            raise PedkError(rc, msg)
        self.device = device

    def check(self, rc: int) -> None:
        if rc != 0:
            raise PedkError(rc, self.lib.pedk_last_error(self.h).decode())

    def close(self) -> None:
        if self.h:
            hits = int(self.lib.pedk_guard_hits(self.h))
            self.lib.pedk_close(self.h)
            self.h = C.c_void_p()
            if hits:
                raise PedkError(-8, f"PEDK_GUARD: {hits} device block(s) were written past their end")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def stream(self) -> int:
        """cudaStream_t of the context (wrap with torch.cuda.ExternalStream to time it)."""
        return int(self.lib.pedk_stream(self.h) or 0)

    def stats(self) -> dict:
        st = Stats()
        self.check(self.lib.pedk_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def reset_stats(self) -> None:
        self.check(self.lib.pedk_reset_stats(self.h))

    # ---- multi-GPU: one rank per process ----
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        self.check(self.lib.pedk_comm_unique_id(self.h, buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, nranks: int) -> None:
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        self.check(self.lib.pedk_comm_init(self.h, buf, rank, nranks))

    # ---- whole path ----
    def kmer_edges(self, control_fastq, control_bytes: int, target_fastq, target_bytes: int, *, device_buffers: bool,
                   k: int = 20, clipping: int = 0, padded: bool = False, reference_control: bool = False) -> "Result":
        """padded (device buffers only): the buffers are addressable up to fastq_padded_bytes(n) and are read
        in place instead of being copied.  reference_control: control_fastq is the text of a reference FASTA
        that is tiled into reads like `ped.pl ref=` does (pedk_kmer_edges_ref)."""
        r = C.c_void_p()
        where = (2 if padded else 1) if device_buffers else 0
        fn = self.lib.pedk_kmer_edges_ref if reference_control else self.lib.pedk_kmer_edges
        self.check(fn(self.h, _ptr(control_fastq), control_bytes, _ptr(target_fastq), target_bytes, where, k, clipping,
                      C.byref(r)))
        return Result(self, r)

    def join(self, control: "Sample", target: "Sample") -> "Result":
        r = C.c_void_p()
        self.check(self.lib.pedk_join(self.h, control.h, target.h, C.byref(r)))
        return Result(self, r)

    # ---- raw kernels (unit tests) ----
    def test_sort_u64(self, keys_dev, n: int, key_bits: int) -> None:
        self.check(self.lib.pedk_test_sort_u64(self.h, _ptr(keys_dev), n, key_bits))

    def test_sort_pairs(self, keys_dev, vals_dev, n: int, key_bits: int) -> None:
        self.check(self.lib.pedk_test_sort_pairs(self.h, _ptr(keys_dev), _ptr(vals_dev), n, key_bits))

    def test_rle(self, sorted_dev, n: int, ukey_dev, count_dev) -> int:
        d = C.c_uint64()
        self.check(self.lib.pedk_test_rle(self.h, _ptr(sorted_dev), n, _ptr(ukey_dev), _ptr(count_dev), C.byref(d)))
        return int(d.value)

    def synth_fastq_device(self, params: SynthParams, genome_dev, first: int, n: int, out_dev) -> None:
        self.check(self.lib.pedk_synth_fastq_device(self.h, C.byref(params), _ptr(genome_dev), first, n,
                                                    _ptr(out_dev)))


class Sample:
    """One sample (target or control): sort_uniq -> count, resident in HBM."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.h = C.c_void_p()
        ctx.check(ctx.lib.pedk_sample_create(ctx.h, C.byref(self.h)))

    def destroy(self) -> None:
        if self.h:
            self.ctx.lib.pedk_sample_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.destroy()

    def add_fastq(self, data, nbytes: Optional[int] = None) -> None:
        """Append FASTQ text held in HOST memory (bytes / numpy / pinned tensor)."""
        n = nbytes if nbytes is not None else len(data)
        self.ctx.check(self.ctx.lib.pedk_sample_add_fastq(self.h, _ptr(data), n))

    def add_fastq_device(self, dev_ptr, nbytes: int) -> None:
        self.ctx.check(self.ctx.lib.pedk_sample_add_fastq_device(self.h, _ptr(dev_ptr), nbytes))

    def borrow_fastq_device(self, dev_buf, nbytes: int) -> None:
        """Text already in HBM, read in place: dev_buf must be addressable up to fastq_padded_bytes(nbytes)."""
        self.ctx.check(self.ctx.lib.pedk_sample_borrow_fastq_device(self.h, _ptr(dev_buf), nbytes,
                                                                    fastq_padded_bytes(nbytes)))

    def ingest(self, clipping: int = 0) -> IngestInfo:
        """mkSortUniq (ped.pl:1365-1530)."""
        info = IngestInfo()
        self.ctx.check(self.ctx.lib.pedk_sample_ingest(self.h, clipping, C.byref(info)))
        return info

    def ingest_reference(self, window: int = 0, step: int = 0) -> IngestInfo:
        """mkChrFromFile + mkControlRead (ped.pl:1027-1060, 1089-1136): the text fed with add_fastq is
        a reference FASTA; its 100-base windows every 2 bases become the reads of this sample."""
        info = IngestInfo()
        self.ctx.check(self.ctx.lib.pedk_sample_ingest_reference(self.h, window, step, C.byref(info)))
        return info

    def chromosomes(self, sequences: bool = True):
        """[(file name ped.pl gives it, cleaned sequence or length, overwritten by a later one)]"""
        n = C.c_int()
        self.ctx.check(self.ctx.lib.pedk_sample_num_chromosomes(self.h, C.byref(n)))
        out = []
        for i in range(n.value):
            name = C.create_string_buffer(4096)
            length = C.c_uint64()
            ow = C.c_int()
            self.ctx.check(self.ctx.lib.pedk_sample_chromosome(self.h, i, name, 4096, C.byref(length), C.byref(ow),
                                                               None, 0))
            seq = length.value
            if sequences:
                buf = C.create_string_buffer(max(1, length.value))
                self.ctx.check(self.ctx.lib.pedk_sample_chromosome(self.h, i, None, 0, None, None, buf, length.value))
                seq = buf.raw[:length.value]
            out.append((name.value.decode(), seq, bool(ow.value)))
        return out

    def ref20_uniq(self, k: int = 20) -> int:
        """mk20mer + mkUniq (ped.pl:1138-1211) on an ingested reference: lines of all ref20_uniq files."""
        n = C.c_uint64()
        self.ctx.check(self.ctx.lib.pedk_sample_ref20_uniq(self.h, k, C.byref(n)))
        return int(n.value)

    def ref20_uniq_text(self, tag: int) -> bytes:
        return self._text(self.ctx.lib.pedk_sample_ref20_uniq_text, tag)

    def map_kmer_text(self, snp_text: bytes) -> bytes:
        """mapKmerSub (ped.pl:641-685): edge lines -> `<tmpdir>/<t>.map.TAG` lines."""
        p = C.c_void_p()
        n = C.c_uint64()
        self.ctx.check(self.ctx.lib.pedk_sample_map_kmer_text(self.h, snp_text, len(snp_text), C.byref(p), C.byref(n)))
        try:
            return C.string_at(p, n.value)
        finally:
            self.ctx.lib.pedk_free(p)

    def cnv_text(self, control: "Sample", target: "Sample") -> bytes:
        """cnvJoin (ped.pl:934-989) with this sample as the indexed reference: `<t>.<c>.cnv`."""
        p = C.c_void_p()
        n = C.c_uint64()
        self.ctx.check(self.ctx.lib.pedk_cnv_text(self.h, control.h, target.h, C.byref(p), C.byref(n)))
        try:
            return C.string_at(p, n.value)
        finally:
            self.ctx.lib.pedk_free(p)

    def verify_text(self, control: "Sample", target: "Sample", map_text: bytes) -> bytes:
        """snpMkT ... kmerReadCount (ped.pl:380-386) with this sample as the reference (chromosomes):
        the map lines of all tags -> `<t>.kmer.snp` of a run with a reference."""
        p = C.c_void_p()
        n = C.c_uint64()
        self.ctx.check(self.ctx.lib.pedk_verify_text(target.h, control.h, self.h, map_text, len(map_text),
                                                     C.byref(p), C.byref(n)))
        try:
            return C.string_at(p, n.value)
        finally:
            self.ctx.lib.pedk_free(p)

    def count(self, k: int = 20) -> CountInfo:
        """countKmer (ped.pl:784-932)."""
        info = CountInfo()
        self.ctx.check(self.ctx.lib.pedk_sample_count(self.h, k, C.byref(info)))
        self.last_count = info
        return info

    def table(self) -> Tuple[np.ndarray, np.ndarray]:
        """The count table of count/<s>.count.*.gz: (kmers u64 ascending, counts u32), both strands."""
        n = C.c_uint64()
        self.ctx.check(self.ctx.lib.pedk_sample_table(self.h, None, None, 0, C.byref(n)))
        km = np.empty(n.value, dtype=np.uint64)
        ct = np.empty(n.value, dtype=np.uint32)
        self.ctx.check(self.ctx.lib.pedk_sample_table(self.h, _ptr(km), _ptr(ct), n.value, C.byref(n)))
        return km, ct

    def table_digest(self) -> Tuple[int, int]:
        """(rows, digest) of this rank's part of the count table (pedk_sample_table_digest)."""
        n = C.c_uint64()
        d = C.c_uint64()
        self.ctx.check(self.ctx.lib.pedk_sample_table_digest(self.h, C.byref(n), C.byref(d)))
        return int(n.value), int(d.value)

    def reads(self, group: int = 0) -> Tuple[np.ndarray, np.ndarray, int]:
        """Unique canonical packed reads of a length group: (words [n, W] u64, pal [n] u8, read_len)."""
        n = C.c_uint64()
        L = C.c_int32()
        W = C.c_int32()
        lib = self.ctx.lib
        self.ctx.check(lib.pedk_sample_reads(self.h, group, None, None, 0, C.byref(n), C.byref(L), C.byref(W)))
        words = np.empty((n.value, W.value), dtype=np.uint64)
        pal = np.empty(n.value, dtype=np.uint8)
        self.ctx.check(lib.pedk_sample_reads(self.h, group, _ptr(words), _ptr(pal), n.value, C.byref(n), C.byref(L),
                                             C.byref(W)))
        return words, pal, int(L.value)

    def _text(self, fn, tag: int) -> bytes:
        p = C.c_void_p()
        n = C.c_uint64()
        self.ctx.check(fn(self.h, tag, C.byref(p), C.byref(n)))
        try:
            return C.string_at(p, n.value)
        finally:
            self.ctx.lib.pedk_free(p)

    def count_text(self, tag: int) -> bytes:
        return self._text(self.ctx.lib.pedk_sample_count_text, tag)

    def lbc_text(self, tag: int) -> bytes:
        return self._text(self.ctx.lib.pedk_sample_lbc_text, tag)

    def sort_uniq_text(self, tag: int) -> bytes:
        return self._text(self.ctx.lib.pedk_sample_sort_uniq_text, tag)


class Result:
    """The edge list: `<target>/<target>.kmer.snp` (ped.pl:374-378, 777)."""

    def __init__(self, ctx: Context, h: C.c_void_p):
        self.ctx = ctx
        self.h = h

    def destroy(self) -> None:
        if self.h:
            self.ctx.lib.pedk_result_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.destroy()

    @property
    def num_edges(self) -> int:
        return int(self.ctx.lib.pedk_result_num_edges(self.h))

    def text(self) -> bytes:
        p = C.c_char_p()
        n = C.c_uint64()
        self.ctx.check(self.ctx.lib.pedk_result_text(self.h, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    def edges(self) -> np.ndarray:
        n = C.c_uint64()
        self.ctx.check(self.ctx.lib.pedk_result_edges(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=EDGE_DTYPE)
        self.ctx.check(self.ctx.lib.pedk_result_edges(self.h, _ptr(out), n.value, C.byref(n)))
        return out


def table_digest(kmers: np.ndarray, counts: np.ndarray) -> Tuple[int, int]:
    """The same digest computed with numpy from a (kmers, counts) table, e.g. the oracle's."""
    M = np.uint64
    with np.errstate(over="ignore"):
        x = kmers.astype(M) * M(0x9E3779B97F4A7C15) + counts.astype(M)
        x = x + M(0x9E3779B97F4A7C15)
        x = (x ^ (x >> M(30))) * M(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> M(27))) * M(0x94D049BB133111EB)
        x = x ^ (x >> M(31))
        return int(len(kmers)), int(x.sum(dtype=M))


def edge_pass1_table() -> np.ndarray:
    """2048 words: bit `code` set iff a (k-1)-mer group with that class code passes pass 1 (host-only)."""
    bits = np.zeros(2048, dtype=np.uint32)
    rc = load_library().pedk_edge_pass1_table(_ptr(bits))
    if rc:
        raise PedkError(rc, "pedk_edge_pass1_table")
    return bits


def chromosome_file_name(header_line: bytes) -> str:
    """The chromosome file ped.pl's mkChrFromFile opens for a FASTA header line (host-only)."""
    buf = C.create_string_buffer(len(header_line) + 8)
    rc = load_library().pedk_chromosome_file_name(header_line, buf, len(header_line) + 8)
    if rc:
        raise PedkError(rc, "pedk_chromosome_file_name")
    return buf.value.decode()


def vcf_text(kmer_snp: bytes, target_name: str) -> bytes:
    """toVcf for method kmer (ped.pl:1532-1574): `<t>.kmer.snp` -> `<t>.kmer.vcf` (host-only)."""
    lib = load_library()
    p = C.c_void_p()
    n = C.c_uint64()
    rc = lib.pedk_vcf_text(kmer_snp, len(kmer_snp), target_name.encode(), C.byref(p), C.byref(n))
    if rc:
        raise PedkError(rc, "pedk_vcf_text")
    try:
        return C.string_at(p, n.value)
    finally:
        lib.pedk_free(p)


def verify_strings(map_text: bytes, chromosomes: dict, read_length: int):
    """Host-only: [(candidate "chr pos ref alt", reference string, mutant string)] of a position map for one
    read length (snpMkT / snpMkC up to the windows, ped.pl:2087-2192).  chromosomes: file name -> bytes."""
    lib = load_library()
    names = list(chromosomes)
    n = len(names)
    c_names = (C.c_char_p * max(1, n))(*[x.encode() if isinstance(x, str) else x for x in names])
    c_seqs = (C.c_char_p * max(1, n))(*[chromosomes[x] for x in names])
    c_lens = (C.c_uint64 * max(1, n))(*[len(chromosomes[x]) for x in names])
    p = C.c_void_p()
    ln = C.c_uint64()
    rc = lib.pedk_verify_strings_text(map_text, len(map_text), c_names, c_seqs, c_lens, n, read_length, C.byref(p),
                                      C.byref(ln))
    if rc:
        raise PedkError(rc, "pedk_verify_strings_text")
    try:
        text = C.string_at(p, ln.value)
    finally:
        lib.pedk_free(p)
    return [tuple(x.split(b"\t")) for x in text.split(b"\n")[:-1]]


def verify_counts_text(map_text: bytes, target_hits, control_hits, length: int, clength: int) -> bytes:
    """Host-only: kmerReadCount (ped.pl:2388-2500) from the numbers of windows found per string (uint32 arrays,
    two per candidate in the order of verify_strings) -> `<t>.kmer.snp` of a run with a reference."""
    lib = load_library()
    th = np.ascontiguousarray(target_hits, dtype=np.uint32)
    ch = np.ascontiguousarray(control_hits, dtype=np.uint32)
    p = C.c_void_p()
    ln = C.c_uint64()
    rc = lib.pedk_verify_counts_text(map_text, len(map_text), _ptr(th), _ptr(ch), len(th), length, clength, C.byref(p),
                                     C.byref(ln))
    if rc:
        raise PedkError(rc, "pedk_verify_counts_text")
    try:
        return C.string_at(p, ln.value)
    finally:
        lib.pedk_free(p)


def fastq_padded_bytes(n: int) -> int:
    """Capacity a device buffer of n text bytes needs to be read in place (pedk_sample_borrow_fastq_device)."""
    return (n + 16383) // 16384 * 16384 + 16384


# ---- sharding plan (host side; no GPU needed) ----

def owner_of_read(words: np.ndarray, nranks: int) -> int:
    w = np.ascontiguousarray(words, dtype=np.uint64)
    return int(load_library().pedk_owner_of_read(_ptr(w), len(w), nranks))


def owner_of_kmer(canonical_kmer: int, k: int, nranks: int) -> int:
    return int(load_library().pedk_owner_of_kmer(int(canonical_kmer), k, nranks))


def kmer_hash(kmer: int, k: int) -> int:
    """The bijection on 2k-bit integers whose top bits bin a canonical k-mer."""
    return int(load_library().pedk_kmer_hash(int(kmer), k))


def kmer_unhash(h: int, k: int) -> int:
    return int(load_library().pedk_kmer_unhash(int(h), k))


def plan_level_a(k: int, nranks: int, rank: int, m_max: int = 0, counts=None, exact: bool = False, n_pass: int = 1,
                 pass_: int = 0) -> dict:
    """Host-side plan of the k-mer exchange (pedk_plan_level_a_pass): cursors, limits, segments, buffer size."""
    cur = np.zeros(256, dtype=np.uint64)
    lim = np.zeros(256, dtype=np.uint64)
    seg = np.zeros(3 * 256 * nranks, dtype=np.uint64)
    n_seg = C.c_int()
    elems = C.c_uint64()
    cap = C.c_uint64()
    c = None if counts is None else np.ascontiguousarray(counts, dtype=np.uint64)
    rc = load_library().pedk_plan_level_a_pass(k, nranks, rank, m_max, None if c is None else _ptr(c),
                                               1 if exact else 0, n_pass, pass_, _ptr(cur), _ptr(lim), _ptr(seg),
                                               C.byref(n_seg), C.byref(elems), C.byref(cap))
    if rc:
        raise PedkError(rc, "pedk_plan_level_a")
    n = n_seg.value
    return {"cur": cur, "lim": lim, "seg_start": seg[:n].copy(), "seg_end": seg[n:2 * n].copy(),
            "seg_bin": seg[2 * n:3 * n].copy(), "buffer_elems": int(elems.value), "capacity": int(cap.value)}


def plan_bucket_owners(hist64, nranks: int) -> np.ndarray:
    h = np.ascontiguousarray(hist64, dtype=np.uint64)
    assert h.shape == (64,)
    out = np.zeros(64, dtype=np.uint8)
    rc = load_library().pedk_plan_bucket_owners(_ptr(h), nranks, _ptr(out))
    if rc:
        raise PedkError(rc, "pedk_plan_bucket_owners")
    return out


