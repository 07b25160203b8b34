"""Build libhfkmc.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hyperfluorescence_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhfkmc.so")
SOURCES = [os.path.join(CSRC, "hfkmc.cu")]
HEADERS = [os.path.join(CSRC, "hf_frm.cuh"), os.path.join(CSRC, "hf_frm_c1.cuh"), os.path.join(CSRC, "hf_frm_small.cuh"), os.path.join(CSRC, "hf_exciton.cuh"), os.path.join(CSRC, "hf_exciton_dense.cuh"), os.path.join(CSRC, "hf_philox.cuh"),
           os.path.join(os.path.dirname(HERE), "include", "hfkmc.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # parity: never contract a*b+c (see hf_frm.cuh)
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return exe


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, "-o", LIB, *SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libhfkmc.so")
    # register / spill report of every kernel; build/ is git-ignored (a copy per round lives in profiles/)
    os.makedirs(os.path.join(os.path.dirname(HERE), "build"), exist_ok=True)
    with open(os.path.join(os.path.dirname(HERE), "build", "libhfkmc.ptxas.log"), "w") as f:
        f.write(res.stdout + res.stderr)
    return LIB


def source_hashes(names=None) -> dict:
    """sha256 of the kernel sources (all of csrc/ by default).  profiles/traffic.json stamps each ncu-derived
    counter with the hashes of the files its kernel is compiled from, and bench.py drops a counter whose stamp
    no longer matches what is built."""
    import hashlib
    out = {}
    for name in sorted(names or os.listdir(CSRC)):
        with open(os.path.join(CSRC, name), "rb") as f:
            out[name] = hashlib.sha256(f.read()).hexdigest()[:16]
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
