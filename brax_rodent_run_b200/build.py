"""Build the CUDA extension in-tree: csrc/librr_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m brax_rodent_run_b200.build [--force]
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
OUT = os.path.join(CSRC, "librr_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-cudart", "static"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.inl")) +
                  glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(INCLUDE, "*.h")))


def up_to_date(out: str, deps) -> bool:
    return os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(d) for d in deps)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date(OUT, sources()):
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, os.path.join(CSRC, "rr_api.cu")]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
