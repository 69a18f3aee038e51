"""Build libb200iop.so (CUDA kernels + C ABI) and libb200_modules.so (C module adapters) in-tree.

    python -m ansel_b200.build            # build if sources are newer than the libraries
    python -m ansel_b200.build --force

nvcc cross-compiles for sm_100a without a GPU.  Numerics flags are part of the parity contract
(DESIGN.md): no FMA contraction, IEEE division/sqrt, flush-to-zero like the reference's FTZ|DAZ.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
IOP = os.path.join(HERE, "iop")
LIB = os.path.join(HERE, "libb200iop.so")
MODLIB = os.path.join(HERE, "libb200_modules.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false", "-ftz=true", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off",
    "-shared", "-cudart", "static",
]
HOSTCC = os.environ.get("B200_HOSTCC", "/usr/bin/gcc")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(srcs, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> None:
    cu = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdr = sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh"))
                 + glob.glob(os.path.join(ROOT, "include", "*.h")))
    if force or _newer(cu + hdr, LIB):
        objs = []
        objdir = os.path.join(HERE, "build")
        os.makedirs(objdir, exist_ok=True)
        procs = []
        for src in cu:
            obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
            objs.append(obj)
            if not force and not _newer([src] + hdr, obj):
                continue
            cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f not in ("-shared",)] + ["-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for src, p in procs:
            out, _ = p.communicate()
            if p.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{out}")
            if verbose and out.strip():
                print(out)
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
               "-Xcompiler", "-fPIC", "-o", LIB] + objs
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
    csrcs = sorted(glob.glob(os.path.join(IOP, "*.c")))
    chdr = sorted(glob.glob(os.path.join(IOP, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h")))
    if csrcs and (force or _newer(csrcs + chdr + [LIB], MODLIB)):
        cmd = [HOSTCC, "-std=gnu11", "-O2", "-Wall", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off",
               "-I", os.path.join(ROOT, "include"), "-I", IOP, "-o", MODLIB] + csrcs + [
               "-L", HERE, "-l:libb200iop.so", "-Wl,-rpath,$ORIGIN", "-lm"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed:\n{r.stdout}")


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB)
