"""In-tree build of the CUDA library (sm_100a) with nvcc.  No JIT cache: the .so lands in
rmf_crowdsim_b200/_lib/ so that it travels with the repository snapshot to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(ROOT)
CSRC = os.path.join(ROOT, "csrc")
LIBDIR = os.path.join(ROOT, "_lib")
LIB = os.path.join(LIBDIR, "librcs.so")

NVCC_FLAGS = [
    "-std=c++17",
    "-O3",
    # rustc never contracts a*b+c: the parity-critical arithmetic must not be fused either
    "--fmad=false",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC,-ffp-contract=off,-fno-fast-math",
    "-shared",
    "-cudart",
    "static",
]


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps() -> list[str]:
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(REPO, "include", "rcs.h"))
    return d


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile librcs.so for sm_100a.  Returns the path of the library."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build librcs.so")
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB, *sources()]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building librcs.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
