"""Build recipe for libcoati_gpu.so (sm_100a only, in-tree so the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcoati_gpu.so")
CLI = os.path.join(HERE, "bin", "coati-gpu")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# -fmad=false: the Viterbi parity contract is an op-order contract; FADD/FADD must
# never be contracted or re-associated.  No -use_fast_math.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-cudart", "static",
]


def sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cc", ".inc"))]
    out.append(os.path.join(HERE, "..", "include", "coati_gpu.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cus = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    ccs = sorted(os.path.join(CSRC, "host", f) for f in os.listdir(os.path.join(CSRC, "host"))
                 if f.endswith(".cc") and f != "cli_main.cc") if os.path.isdir(os.path.join(CSRC, "host")) else []
    cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + cus + ccs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libcoati_gpu.so")
    # the CLI front end (coati-gpu alignpair|sample), linked against the library next to it
    os.makedirs(os.path.join(HERE, "bin"), exist_ok=True)
    cli = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-Wall", "-o", CLI,
           os.path.join(CSRC, "host", "cli_main.cc"), "-L" + HERE, "-lcoati_gpu", "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cli, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building coati-gpu")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
