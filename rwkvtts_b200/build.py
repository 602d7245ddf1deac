"""Builds librwkvtts_wkv7.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m rwkvtts_b200.build [--force] [--verbose] [--variants]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.  Every source is
compiled to its own object (in parallel) and linked; a library is rebuilt when the hash of its sources,
headers and flags differs from the stamp written next to it (content, not mtime: a prebuilt .so that
travelled with its sources is recognised as current on the GPU box).

Variants (test infrastructure, built by `--variants` / `build_variants()`):
  librwkvtts_wkv7_delay.so    production protocol + an injected stall of group C2 in the backward
  librwkvtts_wkv7_oldbar.so   round-1 single `out_ready` barrier + the same stall (shows the hazard the
                              per-parity barriers close; scripts/stress_wkv7.py, tests/test_stress_gpu.py)
Select a library with the environment variable RWKVTTS_LIB (rwkvtts_b200/_lib.py).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "librwkvtts_wkv7.so")
SOURCES = ["capi.cu", "wkv7_scan.cu", "wkv7_tc_fwd.cu", "wkv7_tc_bwd.cu", "tmix_fused.cu", "adam.cu", "gather.cu",
           "linear_ce.cu", "wkv7_step_exact.cu", "decode_step.cu"]
# per-file flags: the reference-order step kernel must come out with the reference build's flush-to-zero fast-math
# instructions (model/llm/rwkv_asr_cuda_whisper.py:50) to be bit-identical to it
FILE_FLAGS = {"wkv7_step_exact.cu": ["--use_fast_math"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
VARIANTS = {
    "delay": ["-DRWKVTTS_BWD_DELAY_C2"],
    "oldbar": ["-DRWKVTTS_BWD_SINGLE_OUT_READY", "-DRWKVTTS_BWD_DELAY_C2"],
}


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest(extra) -> str:
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "rwkvtts_wkv7.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    h.update(" ".join(NVCC_FLAGS + list(extra) + _sources() + [repr(sorted(FILE_FLAGS.items()))]).encode())
    return h.hexdigest()


def _lib_path(variant=None) -> str:
    return LIB if variant is None else os.path.join(HERE, f"librwkvtts_wkv7_{variant}.so")


def _stale(lib, digest) -> bool:
    try:
        return not os.path.exists(lib) or open(lib + ".stamp").read().strip() != digest
    except OSError:
        return True


def _run(cmd, verbose):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd[-3:]))


def build(force: bool = False, verbose: bool = False, variant=None) -> str:
    extra = VARIANTS[variant] if variant else []
    lib, digest = _lib_path(variant), _digest(extra)
    if not force and not _stale(lib, digest):
        return lib
    nvcc = os.environ.get("NVCC", "nvcc")
    tag = variant or "main"
    os.makedirs(OBJ, exist_ok=True)
    # a variant only changes the chunked kernels: the other objects are shared with the main build
    special = {"wkv7_tc_fwd.cu", "wkv7_tc_bwd.cu"} if variant else set(_sources())
    jobs, objs = [], []
    for s in _sources():
        t = tag if s in special else "main"
        o = os.path.join(OBJ, f"{os.path.splitext(s)[0]}.{t}.o")
        objs.append(o)
        if t == tag or not os.path.exists(o):
            ex = extra if s in special and variant else []
            jobs.append([nvcc, *NVCC_FLAGS, *FILE_FLAGS.get(s, []), *ex, *(["-Xptxas", "-v"] if verbose else []), "-c",
                         os.path.join(CSRC, s), "-o", o])
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(lambda c: _run(c, verbose), jobs))
    _run([nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", lib], verbose)
    with open(lib + ".stamp", "w") as f:
        f.write(digest + "\n")
    return lib


def build_variants(force: bool = False, verbose: bool = False):
    build(force=False, verbose=verbose)          # shared objects first
    return [build(force=force, verbose=verbose, variant=v) for v in VARIANTS]


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    if "--variants" in sys.argv:
        for p in build_variants(force="--force" in sys.argv, verbose="--verbose" in sys.argv):
            print(p)
