"""Build libsstb200.so (sm_100a) in-tree with nvcc.  `python -m sst_core_b200.build [--force]`."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libsstb200.so"
SOURCES = [CSRC / "engine.cu", CSRC / "partition.cc"]
DEPS = SOURCES + [CSRC / "components.cuh", CSRC / "rng.cuh", HERE.parent / "include" / "sstb200.h"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr",
    "-fmad=false",  # IEEE-separate mul/add: RNG nextUniform / distributions must round like the reference on x86-64
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and LIB.exists() and all(LIB.stat().st_mtime >= d.stat().st_mtime for d in DEPS):
        return LIB
    extra = os.environ.get("SSTB200_NVCC_EXTRA", "").split()  # experiments, e.g. -DSSTB200_MIN_BLOCKS=5
    cmd = ["nvcc", *NVCC_FLAGS, *extra, "-o", str(LIB), *map(str, SOURCES)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libsstb200.so")
    (HERE / "build_ptxas.log").write_text(r.stdout)
    if verbose:
        print(r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
