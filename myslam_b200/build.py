"""Build libeslam_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "eslam_b200.cu")
OUT = os.path.join(HERE, "libeslam_b200.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("eslam_b200.cu", "field.cuh", "render.cuh", "sample.cuh", "optim.cuh", "exchange.cuh", "keyframes.cuh", "ingest.cuh", "qplane.cuh", "qform.cuh", "qbwd.cuh", "mcubes.cuh")]
DEPS.append(os.path.join(os.path.dirname(HERE), "include", "eslam_b200.h"))
DEPS.append(os.path.abspath(__file__))  # the flags live here


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the sm_100a library cannot be built")


def up_to_date() -> bool:
    return os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "--split-compile", "1", "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC]
    # A fixed thread count on purpose: the optimiser partitions the module by the number of split-compile threads, and
    # the partitioning changes the code of nearly every kernel (1 thread and 4-8 threads give different builds of the
    # same source; "0" = "as many as there are CPUs right now" made that vary from build to build).  Pinned by
    # measurement (tools/build_variants.py + tools/variant_times.py, profiles/r02_variant_times.txt): the unsplit build
    # runs the dominant kernel in 151.5 us instead of 157.1 and a mapping iteration in 217.5 us instead of 223.7.
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT
