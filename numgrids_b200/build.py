"""Build libnumgrids_b200.so in-tree with nvcc for sm_100a (no torch extension machinery).

    python -m numgrids_b200.build [--force]

Flags: ``-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false``.  ``-fmad=false``
keeps every ``a*b+c`` a separately rounded DMUL + DADD, which the bit-exact paths rely on; the
FFT kernels opt in to FMA with explicit ``fma()``.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libnumgrids_b200.so")
OBJDIR = os.path.join(HERE, "build")
SOURCES = ["api.cu", "fd.cu", "fd_row.cu", "fd_tile.cu", "dense.cu", "fft.cu", "fft2.cu", "fft_tile.cu", "fft_mixed.cu", "fdlap.cu", "fdlap3d.cu", "curv.cu", "curvlap3d.cu", "slab.cu", "transpose.cu", "bc.cu", "quad.cu", "interp.cu"]
# translation units compiled more than once with different macros (object name -> (source, defines)):
# the contiguous-axis FD kernel is the slowest file to compile; its half widths 3-4 build as a second object
EXTRA_OBJECTS = {"fd_row_wide.o": ("fd_row.cu", ["-DNG_ROW_PSET=1"])}
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _newer(a: str, b: str) -> bool:
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def _deps() -> list[str]:
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdr.append(os.path.join(HERE, "..", "include", "numgrids_b200.h"))
    return hdr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = _deps()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if force or _newer(s, o) or any(_newer(h, o) for h in hdrs):
            cmd = [NVCC, *FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    for obj, (src, defs) in EXTRA_OBJECTS.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, obj)
        if force or _newer(s, o) or any(_newer(h, o) for h in hdrs):
            cmd = [NVCC, *FLAGS, *defs, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for msg in ex.map(run, jobs):
                if verbose and msg:
                    print(msg)
    objs = [os.path.join(OBJDIR, s.replace(".cu", ".o")) for s in SOURCES] + [os.path.join(OBJDIR, o) for o in EXTRA_OBJECTS]
    if force or jobs or not os.path.exists(LIB):
        run([NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
