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
DEPS = SOURCES + [CSRC / "components.cuh", CSRC / "rng.cuh", CSRC / "parallel_form.cuh", CSRC / "ordered_form.cuh", CSRC / "clock_form.cuh", CSRC / "elements.def",
                  HERE.parent / "include" / "sstb200.h"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr",
    "-fmad=false",  # IEEE-separate mul/add: RNG nextUniform / distributions must round like the reference on x86-64
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def build(force: bool = False, verbose: bool = False, extra_elements=None, out=None) -> Path:
    """Build the library.  ``extra_elements`` = (header.cuh, list.def): an element library from outside the tree is
    compiled in (csrc/elements.def); ``out``: where to write that library (default: the in-tree libsstb200.so;
    load another one with SSTB200_LIB=/path/to/lib.so)."""
    lib = Path(out) if out else LIB
    deps = DEPS + [Path(p) for p in (extra_elements or ())]
    if not force and lib.exists() and all(lib.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return lib
    extra = os.environ.get("SSTB200_NVCC_EXTRA", "").split()  # experiments, e.g. -DSSTB200_MIN_BLOCKS=5
    if extra_elements:
        hdr, lst = (str(Path(p).resolve()) for p in extra_elements)
        extra += [f'-DSSTB200_EXTRA_ELEMENTS_HEADER="{hdr}"', f'-DSSTB200_EXTRA_ELEMENTS_DEF="{lst}"', f"-I{CSRC}"]
    cmd = ["nvcc", *NVCC_FLAGS, *extra, "-o", str(lib), *map(str, SOURCES)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libsstb200.so")
    if lib == LIB:
        (HERE / "build_ptxas.log").write_text(r.stdout)
    if verbose:
        print(r.stdout)
    return lib


if __name__ == "__main__":
    # python -m sst_core_b200.build [--force] [--with header.cuh:list.def] [--out lib.so]
    argv = sys.argv[1:]
    ext = argv[argv.index("--with") + 1].split(":") if "--with" in argv else None
    dst = argv[argv.index("--out") + 1] if "--out" in argv else None
    print(build(force="--force" in argv, verbose=True, extra_elements=ext, out=dst))
