#!/usr/bin/env python3
"""Hand build of the UNMODIFIED reference sst-core into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference's own build is autotools (absent in this image) and its experimental CMake is stale,
so this recipe compiles the reference sources *where they lie* under /root/reference with plain g++
and writes every output under oracle/_ref/ (git-ignored, travels to the GPU box with gpurun):

    oracle/_ref/gen/sst_config.h, build_info.h   hand-written config headers (macro list: configure.ac:108-163)
    oracle/_ref/obj/*.o                          objects
    oracle/_ref/libexec/sstsim.x                 the simulator binary (what `sst` execs; core/bootsst.cc)
    oracle/_ref/lib/sst-elements-library/libcoreTestElement.so   the reference's test elements
    oracle/_ref/lib/sst-elements-library/libpdes.so       OUR out-of-tree ELI plugins built against the
                                                                 reference headers (oracle/probe/*.cc): the event
                                                                 order trace tool and the PHOLD / clock-population
                                                                 components that the reference core runs as oracle
    oracle/_ref/etc/sst/sstsimulator.conf

No reference source is copied into the repo.  Source lists are read from the reference's own
Makefile.am / Makefile.inc files at build time (core/Makefile.am:174-265 and the included .inc files).
Nothing in the product path uses this; only tests/, __graft_entry__ and bench.py's cpu/reference arm do.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import re
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("SST_REFERENCE", "/root/reference"))
OUT = HERE / "_ref"
CORE = REF / "src" / "sst" / "core"

CONFIG_H = r"""
#pragma once
#define HAVE_LIBZ 1
#define HAVE_BACKTRACE 1
#define HAVE_DLFCN_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_UNISTD_H 1
#define HAVE_EXECINFO_H 1
#define PACKAGE_VERSION "-dev"
#define PACKAGE_STRING "SSTCore -dev"
#define PACKAGE_NAME "SSTCore"
#define PACKAGE_BUGREPORT "sst@sandia.gov"
#define SSTCORE_GIT_BRANCH "N/A"
#define SSTCORE_GIT_HEADSHA "-dev"
#define SSTCORE_GIT_COMMITCOUNT "0"
#define SST_CPPFLAGS ""
#define SST_CFLAGS ""
#define SST_CXXFLAGS "-std=c++17 -O2"
#define SST_LDFLAGS ""
#define SST_CC "gcc"
#define SST_CXX "g++"
#define SST_LD "ld"
#define SST_MPICC ""
#define SST_MPICXX ""
#define SST_CPP "gcc -E"
#define SST_CXXCPP "g++ -E"
#define SST_PYTHON_CPPFLAGS ""
#define SST_PYTHON_LDFLAGS ""
#ifndef __STDC_FORMAT_MACROS
#define __STDC_FORMAT_MACROS 1
#endif
#define SST_INSTALL_PREFIX "@PREFIX@"
#define USE_MEMPOOL 1
#define SST_CONFIG_HAVE_PYTHON 1
#define SST_CONFIG_HAVE_PYTHON3 1
"""


def cc_tokens(path: Path, var_prefixes=None):
    """All tokens ending in .cc from a Makefile fragment (optionally only within given variables)."""
    text = path.read_text()
    if var_prefixes is None:
        return re.findall(r"([\w/\-\.]+\.cc)\b", text)
    out = []
    # join continuation lines
    joined = re.sub(r"\\\n", " ", text)
    for line in joined.splitlines():
        m = re.match(r"\s*([\w]+)\s*\+?=\s*(.*)", line)
        if m and any(m.group(1).startswith(p) for p in var_prefixes):
            out += re.findall(r"([\w/\-\.]+\.cc)\b", m.group(2))
    return out


def source_list():
    srcs = ["main.cc"] + cc_tokens(CORE / "Makefile.am", ["sst_core_sources"])
    for sub in ("model", "shared", "impl", "sync", "util"):
        for inc in sorted((CORE / sub).rglob("Makefile.inc")):
            for tok in cc_tokens(inc):
                srcs.append(tok)
    seen, uniq = set(), []
    for s in srcs:
        if s not in seen and (CORE / s).exists():
            seen.add(s)
            uniq.append(s)
    return uniq


def test_element_sources():
    toks = cc_tokens(CORE / "testElements" / "Makefile.inc")
    out = []
    for t in toks:
        p = CORE / t
        if p.exists() and t not in out:
            out.append(t)
    return out


def py_flags():
    inc = sysconfig.get_paths()["include"]
    # base interpreter include (venv points at the system python headers)
    cfgvars = sysconfig.get_config_vars()
    incs = {inc, cfgvars.get("INCLUDEPY", inc)}
    libdir = cfgvars.get("LIBDIR", "/usr/lib")
    ver = cfgvars.get("LDVERSION") or cfgvars.get("VERSION")
    return [f"-I{i}" for i in incs], [f"-L{libdir}", f"-lpython{ver}"]


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write("FAILED: " + " ".join(map(str, cmd)) + "\n" + r.stdout[-4000:] + "\n")
        raise SystemExit(1)
    return r.stdout


def compile_one(args):
    src, obj, flags = args
    if obj.exists() and obj.stat().st_mtime >= src.stat().st_mtime:
        return
    obj.parent.mkdir(parents=True, exist_ok=True)
    run(["g++", *flags, "-c", str(src), "-o", str(obj)])


def main():
    if not CORE.exists():
        print(f"reference not present at {REF}; nothing to build (prebuilt oracle/_ref is used if shipped)")
        return 0
    jobs = int(os.environ.get("JOBS", os.cpu_count() or 4))
    gen = OUT / "gen"
    (gen / "sst" / "core").mkdir(parents=True, exist_ok=True)
    cfg = CONFIG_H.replace("@PREFIX@", str(OUT))
    for d in (gen, gen / "sst" / "core"):
        for name, body in (("sst_config.h", cfg), ("build_info.h", "#pragma once\n")):
            p = d / name
            if not p.exists() or p.read_text() != body:
                p.write_text(body)
    pyinc, pyld = py_flags()
    flags = ["-std=c++17", "-O2", "-fPIC", "-w", "-DTIXML_USE_STL", "-DSST_BUILDING_CORE=1", "-DHAVE_CONFIG_H",
             f"-I{gen}", f"-I{gen}/sst/core", f"-I{REF}/src", f"-I{REF}/external", *pyinc]

    core_srcs = source_list()
    tasks = []
    core_objs = []
    for s in core_srcs:
        obj = OUT / "obj" / "core" / (s.replace("/", "_")[:-3] + ".o")
        core_objs.append(obj)
        tasks.append((CORE / s, obj, flags))
    te_objs = []
    for s in test_element_sources():
        obj = OUT / "obj" / "te" / (s.replace("/", "_")[:-3] + ".o")
        te_objs.append(obj)
        tasks.append((CORE / s, obj, flags))
    probe_objs = []
    probe_flags = [f for f in flags if f != "-DSST_BUILDING_CORE=1"]
    for p in sorted((HERE / "probe").glob("*.cc")):
        obj = OUT / "obj" / "probe" / (p.stem + ".o")
        probe_objs.append(obj)
        tasks.append((p, obj, probe_flags))

    # reference-side bindings of the C ABI (integration/*.cc): ELI plugins for stock SST that call libsstb200.so
    shim_objs = []
    for p in sorted((HERE.parent / "integration").glob("*.cc")):
        obj = OUT / "obj" / "integration" / (p.stem + ".o")
        shim_objs.append(obj)
        tasks.append((p, obj, probe_flags + [f"-I{HERE.parent / 'include'}"]))

    print(f"[build_ref] {len(tasks)} translation units, {jobs} jobs", flush=True)
    with cf.ThreadPoolExecutor(jobs) as ex:
        list(ex.map(compile_one, tasks))

    libexec = OUT / "libexec"
    libexec.mkdir(exist_ok=True)
    exe = libexec / "sstsim.x"
    run(["g++", "-rdynamic", "-o", str(exe), *map(str, core_objs), *pyld, "-lz", "-ldl", "-lrt", "-lpthread", "-lm"])
    # the reference's generator classes behind a tiny command-line driver (long-stream fixtures)
    tool = HERE / "tools" / "rng_ref_driver.cc"
    if tool.exists():
        run(["g++", *probe_flags, "-rdynamic", "-o", str(libexec / "rng_ref_driver"), str(tool),
             *[str(o) for o in core_objs if o.name != "main.o"], *pyld, "-lz", "-ldl", "-lrt", "-lpthread", "-lm"])
    libdir = OUT / "lib" / "sst-elements-library"
    libdir.mkdir(parents=True, exist_ok=True)
    run(["g++", "-shared", "-o", str(libdir / "libcoreTestElement.so"), *map(str, te_objs)])
    if probe_objs:
        run(["g++", "-shared", "-o", str(libdir / "libpdes.so"), *map(str, probe_objs)])
    product = HERE.parent / "sst_core_b200"
    if shim_objs and (product / "libsstb200.so").exists():
        run(["g++", "-shared", "-o", str(libdir / "libb200.so"), *map(str, shim_objs), f"-L{product}", "-lsstb200",
             "-Wl,-rpath,$ORIGIN/../../../../sst_core_b200"])
    etc = OUT / "etc" / "sst"
    etc.mkdir(parents=True, exist_ok=True)
    (etc / "sstsimulator.conf").write_text(f"[SSTCore]\nprefix={OUT}\nlibdir={libdir}\n")
    print(f"[build_ref] built {exe}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
