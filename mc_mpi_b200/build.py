"""In-tree build of libmcb200.so (CUDA kernels + C ABI) for sm_100a.

    python -m mc_mpi_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so lands next to this file so that it
travels with the repository snapshot to the GPU box (it is git-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libmcb200.so")

CUDA_SOURCES = ["mcb_kernels.cu", "mcb_world_kernel.cu", "mcb_layer.cu", "mcb_world.cu", "mcb_facade.cu"]
HEADERS = ["mcb_math.cuh", "mcb_kernels.cuh", "mcb_event.cuh", "mcb_world.cuh", "mcb_host.hpp", os.path.join(INCLUDE, "mcb200.h"),
           os.path.join(INCLUDE, "mcb200", "layer.hpp"),
           os.path.join(INCLUDE, "mcb200", "culayer.hpp")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",          # belt and braces: the physics uses *_rn intrinsics anyway
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-Xptxas", "-v",
    "-shared", "-cudart", "static",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libmcb200.so)")


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps += [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.isfile(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source of the package into libmcb200.so."""
    if not force and not _stale():
        return LIB
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a gcc without libgomp; use PATH's g++
    env.pop("CC", None)
    env.pop("CXX", None)
    cmd = [nvcc_path(), *NVCC_FLAGS, "-ccbin", shutil.which("g++") or "g++",
           "-I", INCLUDE, "-I", CSRC, "-o", LIB,
           *[os.path.join(CSRC, s) for s in CUDA_SOURCES]]
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed (see {log})")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
