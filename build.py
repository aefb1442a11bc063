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
OUT = os.environ.get("PB_BUILD_OUT") or os.path.join(ROOT, "physecs_b200", "lib")   # PB_BUILD_OUT: compile somewhere else (a check build that leaves the shipped .so alone)
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
    ("islands.cu", []),
    ("queries.cu", ["-fmad=false"]),
    ("trimesh_build.cpp", []),
    ("batch.cpp", []),
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


HOST_SRC = ["Scene.cpp", "Meshes.cpp", "MassUtil.cpp", "scene_harness.cpp"]


def find_ecs_includes():
    """EnTT / GLM are the APPLICATION's dependencies (header-only); they are not vendored in this repo.  Looked up in
    PHYSECS_ENTT_INCLUDE / PHYSECS_GLM_INCLUDE, else in the reference tree's vendor/ directory when it is present."""
    entt = os.environ.get("PHYSECS_ENTT_INCLUDE")
    glm = os.environ.get("PHYSECS_GLM_INCLUDE")
    ref = os.environ.get("PHYSECS_REFERENCE", "/root/reference")
    if not entt:
        cand = os.path.join(ref, "vendor", "entt-3.12.2", "single_include", "entt")
        entt = cand if os.path.exists(os.path.join(cand, "entt.hpp")) else None
    if not glm:
        cand = os.path.join(ref, "vendor", "glm 0.9.9.8")
        glm = cand if os.path.exists(os.path.join(cand, "glm", "glm.hpp")) else None
    return entt, glm


def build_host(verbose=False, force=False):
    """libphysecs_b200_scene.so: the host C++ layer (physecs::Scene over entt::registry) + the flat test harness.
    Returns the path, or None when EnTT / GLM headers are not available (the prebuilt library, if any, is kept)."""
    lib = os.path.join(OUT, "libphysecs_b200_scene.so")
    entt, glm = find_ecs_includes()
    if not entt or not glm:
        return lib if os.path.exists(lib) else None
    host = os.path.join(ROOT, "physecs_b200", "host")
    inc = os.path.join(ROOT, "include")
    srcs = [os.path.join(host, f) for f in HOST_SRC]
    deps = srcs + [os.path.join(dp, f) for dp, _, fns in os.walk(os.path.join(inc, "Physecs")) for f in fns] + [os.path.join(inc, "physecs_b200.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-DGLM_FORCE_INLINE", "-I", os.path.join(inc, "Physecs"),
           "-I", os.path.join(inc, "Physecs", "Joints"), "-I", inc, "-I", glm, "-I", entt] + srcs + \
          ["-o", lib, "-L", OUT, "-lphysecs_b200", "-Wl,-rpath,$ORIGIN", "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if verbose and r.stdout:
        print(r.stdout.decode())
    if r.returncode:
        raise RuntimeError("host layer build failed:\n" + r.stdout.decode()[-8000:])
    return lib


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
    print(build_host(verbose="-v" in sys.argv, force="-f" in sys.argv))
