#!/usr/bin/env python3
"""Build libphysecs_b200.so (sm_100a CUDA kernels + C ABI) in-tree with nvcc.

Every unit that restates reference arithmetic is compiled with -fmad=false: with the reference's operation order
this makes bounds / pair sets bit-exact and keeps the one-step solve at the fp32 noise floor (contact switching
amplifies FMA-level differences past 1e-4).  The kernels are memory-bound, so this costs nothing.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(ROOT, "physecs_b200", "csrc")
OUT = os.path.join(ROOT, "physecs_b200", "lib")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "--expt-relaxed-constexpr", "-w"]

UNITS = [
    ("capi.cu", []),
    ("primitives.cu", []),
    ("broadphase.cu", ["-fmad=false"]),
    ("narrowphase.cu", ["-fmad=false"]),
    ("contacts.cu", ["-fmad=false"]),
    ("solver.cu", ["-fmad=false"]),
    ("joints.cu", ["-fmad=false"]),
    ("trimesh_build.cpp", []),
]


def newer(src, obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + deps)


def build(verbose=False, force=False):
    os.makedirs(OUT, exist_ok=True)
    objdir = os.path.join(OUT, "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "physecs_b200.h"))
    headers.append(os.path.abspath(__file__))   # flag changes rebuild everything
    jobs = []
    objs = []
    for name, extra in UNITS:
        src = os.path.join(SRC, name)
        obj = os.path.join(objdir, name + ".o")
        objs.append(obj)
        if force or newer(src, obj, headers):
            cmd = ["nvcc", "-c", src, "-o", obj] + ARCH + COMMON + extra
            if verbose:
                cmd += ["-Xptxas", "-v"]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        return cmd, r.returncode, r.stdout.decode()

    with ThreadPoolExecutor(max_workers=8) as ex:
        for cmd, rc, out in ex.map(run, jobs):
            if verbose and out:
                print(out)
            if rc:
                raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + out[-8000:])
    lib = os.path.join(OUT, "libphysecs_b200.so")
    if jobs or not os.path.exists(lib):
        cmd = ["nvcc", "-shared", "-o", lib] + objs + ARCH
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode:
            raise RuntimeError("link failed:\n" + r.stdout.decode())
    return lib


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
