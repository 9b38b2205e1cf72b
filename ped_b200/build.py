"""Build recipe for libpedk.so (CUDA, sm_100a only) and the test oracle.

`python -m ped_b200.build` compiles in-tree; __graft_entry__.build() calls the
same functions.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ped_b200", "csrc")
LIB = os.path.join(CSRC, "libpedk.so")
SYNTH_SRC = os.path.join(ROOT, "synth", "synth_host.cc")
SYNTH_LIB = os.path.join(ROOT, "synth", "libpedk_synth.so")
ORACLE_SRC = os.path.join(ROOT, "oracle", "ped_oracle.c")
ORACLE_LIB = os.path.join(ROOT, "oracle", "libped_oracle.so")

CU_SOURCES = ["scan.cu", "ingest.cu", "dedup.cu", "kmer.cu", "bucket.cu", "radix.cu", "edges.cu", "synth.cu", "ctx.cu",
              "sample.cu", "join.cu", "comm.cu", "files.cu", "reference.cu", "plan.cu", "refindex.cu", "verify.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
              "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build_libpedk(force: bool = False, verbose: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "pedk.h"))
    objs = []
    jobs = []
    nvcc = _nvcc()
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, "build", src.replace(".cu", ".o"))
        objs.append(o)
        if force or not _newer(o, [s] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=min(8, len(jobs) or 1)) as ex:
        list(ex.map(run, jobs))
    if jobs or force or not _newer(LIB, objs):
        run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lz", "-ldl",
             "-lpthread"])
    return LIB


def build_synth(force: bool = False) -> str:
    """libpedk_synth.so: the host-side synthetic generator (plain C++, no CUDA; include/pedk_synth.h)."""
    deps = [SYNTH_SRC, os.path.join(CSRC, "synth.h"), os.path.join(CSRC, "common.cuh"),
            os.path.join(ROOT, "include", "pedk_synth.h")]
    if force or not _newer(SYNTH_LIB, deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-shared", "-fPIC", "-Wall",
                               "-Wno-unused-function", "-o", SYNTH_LIB, SYNTH_SRC])
    return SYNTH_LIB


def build_oracle(force: bool = False) -> str:
    if force or not _newer(ORACLE_LIB, [ORACLE_SRC]):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-Wall", "-o", ORACLE_LIB, ORACLE_SRC])
    return ORACLE_LIB


if __name__ == "__main__":
    v = "-v" in sys.argv
    print(build_libpedk(force="-f" in sys.argv, verbose=v))
    print(build_synth(force="-f" in sys.argv))
    print(build_oracle(force="-f" in sys.argv))
