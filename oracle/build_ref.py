#!/usr/bin/env python3
"""Build the ORACLE: the unmodified Physecs reference, compiled for Linux/g++.

TEST INFRASTRUCTURE ONLY.  Nothing under physecs_b200/ may import, link or
execute anything produced here; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs use it, as the checker.

What it does
  1. copies /root/reference/{src,include} to a scratch dir under /tmp
     (the reference tree is read-only and its sources never enter this repo),
  2. applies the mechanical MSVC->g++ portability patch of SURVEY.md §8(c)
     (no algorithmic change) plus three instrumentation taps:
       - Scene's private section is made public (harness reads
         potentialContacts / contactConstraints),
       - every pushed ContactConstraints records its (pair index, triangle
         index) key in a global (so the harness can name manifolds),
       - a global hook is called between "create joint constraints" and the
         substep loop (Physecs.cpp:355/:363) so a test can impose the GPU's
         colour-batched constraint order (north_star gate 3),
  3. compiles reference + oracle/ref_harness.cpp into
       oracle/_ref/libphysecs_ref.so          (as shipped)
       oracle/_ref/libphysecs_ref_hashfix.so  ("reference + hash fix": the
         2-line ContactHash/CollisionHash repair, results bit-identical, only
         there so CPU baselines above ~4k bodies finish)
     with -O2 -ffp-contract=off (bit-reproducible fp32).

Only the two .so files land in the repo dir (git-ignored, shipped to the GPU box).
"""
import os
import re
import shutil
import subprocess
import sys

REF = os.environ.get("PHYSECS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
BUILD = os.environ.get("PHYSECS_ORACLE_BUILD", "/tmp/physecs_oracle_build")

SHIM = r"""
// force-included portability shim (oracle build only)
#pragma once
#include <immintrin.h>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <utility>
#include <limits>
#include <vector>
#include <array>
#include <unordered_set>
#include <unordered_map>
#include <algorithm>
#include <mutex>
#include <atomic>
#ifndef __forceinline
#define __forceinline inline __attribute__((always_inline))
#endif
static inline unsigned char _BitScanForward(unsigned long* idx, unsigned long mask) {
    if (!mask) { *idx = 0; return 0; }
    *idx = (unsigned long)__builtin_ctzl(mask);
    return 1;
}
"""


def sub(text, pattern, repl, path, count_min=1, flags=0):
    new, n = re.subn(pattern, repl, text, flags=flags)
    if n < count_min:
        raise RuntimeError(f"patch did not apply ({n} < {count_min}): {pattern!r} in {path}")
    return new


def patch_tree(root, hashfix):
    def rd(p):
        with open(os.path.join(root, p), encoding="utf-8", errors="replace") as f:
            return f.read()

    def wr(p, s):
        with open(os.path.join(root, p), "w", encoding="utf-8") as f:
            f.write(s)

    # (3) MSVC __m128 member access -> GCC vector subscripts
    for p in ["include/Physecs/SIMD.h", "include/Physecs/Constraint1DContainer.h", "src/Constraint1DW.cpp"]:
        s = rd(p)
        s = sub(s, r"\.m128_f32\[", "[", p)
        wr(p, s)
    p = "src/Constraint1DW.cpp"
    s = rd(p)
    s = sub(s, r"ngsMask\.m128_i32\[i\]", "((const int*)&ngsMask)[i]", p, 2)
    wr(p, s)

    # (4) raw __m128 operator overloads only for MSVC (GCC vector types have them)
    p = "include/Physecs/SIMD.h"
    s = rd(p)
    a = s.index("    inline __m128 operator+ (__m128 a, __m128 b)")
    b = s.index("    inline Vec3W operator* (__m128 a, const Vec3W &b)")
    s = s[:a] + "#ifdef _MSC_VER\n" + s[a:b] + "#endif\n" + s[b:]
    wr(p, s)

    # (5),(6),(3) container
    p = "include/Physecs/Constraint1DContainer.h"
    s = rd(p)
    s = sub(s, r"__m128i lanes = _mm_setzero_si128\(\);",
            "__m128i lanes; ConstraintList() : lanes(_mm_setzero_si128()) {}", p)
    s = sub(s, r"_mm_and_epi32", "_mm_and_si128", p)
    s = sub(s, r"lanes\.m128i_i32\[lane\]", "((int*)&lanes)[lane]", p)
    wr(p, s)

    # (8) AVX512VL store -> SSE2 store
    p = "src/GJK.h"
    s = rd(p)
    s = sub(s, r"_mm_storeu_epi32\(indicesArr, maxIndices\)", "_mm_storeu_si128((__m128i*)indicesArr, maxIndices)", p)
    wr(p, s)

    # (7) instantiate the vertex container the public ConvexMesh needs
    p = "src/ConvexMesh.cpp"
    s = rd(p) + "\ntemplate class physecs::ConvexMeshVertices<4>;\n"
    wr(p, s)

    # instrumentation: public Scene internals
    p = "include/Physecs/Physecs.h"
    s = rd(p)
    s = sub(s, r"class Scene \{", "class Scene {\n    public:", p)
    if hashfix:
        # "reference + hash fix": h3 uses entity1, CollisionHash mixes instead of OR-ing -1
        s = sub(s, r"std::size_t h3 = std::hash<int>\{\}\(static_cast<int>\(pair\.entity0\)\);",
                "std::size_t h3 = std::hash<int>{}(static_cast<int>(pair.entity1));", p)
        s = sub(s, r"return ContactHash\{\}\(pair\.contactPair\) << 32 \| pair\.triangleIndex;",
                "return (ContactHash{}(pair.contactPair) * 0x9E3779B97F4A7C15ull) ^ (std::size_t)(unsigned)pair.triangleIndex;", p)
        s = sub(s, r"return h1 \^ h2 << 1 \^ \(h3 \^ h4 << 1\);",
                "return (h1 * 0x9E3779B97F4A7C15ull) ^ (h2 << 1) ^ ((h3 * 0xC2B2AE3D27D4EB4Full) ^ (h4 << 7));", p)
    wr(p, s)

    # instrumentation: manifold keys + pre-solve hook
    p = "src/Physecs.cpp"
    s = rd(p)
    s = sub(s, r'const char\* frameName = "Solver";',
            'const char* frameName = "Solver";\n'
            'std::vector<std::array<int, 2>> physecs_oracle_manifold_keys;\n'
            'void (*physecs_oracle_presolve_hook)(physecs::Scene*) = nullptr;\n'
            'double physecs_oracle_phase_ms[4] = {0, 0, 0, 0};\n', p)
    s = sub(s, r"contactPoints\.clear\(\);\n", "contactPoints.clear();\n    physecs_oracle_manifold_keys.clear();\n", p)
    s = sub(s, r"contactConstraints\.push_back\(cc\);",
            "contactConstraints.push_back(cc);\n                physecs_oracle_manifold_keys.push_back({ index, collisionResult.triangleIndex });", p)
    s = sub(s, r"PhysecsZoneEnd\(ctx7\);",
            "PhysecsZoneEnd(ctx7);\n    if (physecs_oracle_presolve_hook) physecs_oracle_presolve_hook(this);", p)
    wr(p, s)


def stage_tree(hashfix=False):
    """Copy of the reference's src/ + include/ under the scratch dir with the portability patch applied (no compile); returns its root.
    Also used by tests/test_cpu_dropin.py to build examples/dropin_app.cpp against the reference's own headers."""
    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, "shim.h"), "w") as f:
        f.write(SHIM)
    root = os.path.join(BUILD, "ref" + ("_hashfix" if hashfix else ""))
    if os.path.isdir(root):
        shutil.rmtree(root)
    os.makedirs(root)
    shutil.copytree(os.path.join(REF, "src"), os.path.join(root, "src"))
    shutil.copytree(os.path.join(REF, "include"), os.path.join(root, "include"))
    patch_tree(root, hashfix)
    return root


def build(verbose=True):
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference tree not found at {REF}")
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, "shim.h"), "w") as f:
        f.write(SHIM)
    glm_inc = os.path.join(REF, "vendor", "glm 0.9.9.8")
    entt_inc = os.path.join(REF, "vendor", "entt-3.12.2", "single_include", "entt")
    harness = os.path.join(HERE, "ref_harness.cpp")
    jobs = []
    for variant, hashfix in (("", False), ("_hashfix", True)):
        root = stage_tree(hashfix)
        srcs = []
        for dp, _, fns in os.walk(os.path.join(root, "src")):
            for fn in fns:
                # Constraint1D.cpp / Constraint1DW.cpp hold only templates and are
                # #included by Constraint1DContainer.cpp.
                if fn.endswith(".cpp") and fn not in ("Constraint1D.cpp", "Constraint1DW.cpp"):
                    srcs.append(os.path.join(dp, fn))
        srcs.append(harness)
        flags = [
            "-std=gnu++23", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
            "-march=x86-64-v3", "-w",
            "-DGLM_FORCE_INLINE", "-DPHYSECS_EXPORTS", "-DENTT_PACKED_PAGE=1048576", "-DNDEBUG",
            "-include", os.path.join(BUILD, "shim.h"),
            "-I", os.path.join(root, "src"), "-I", os.path.join(root, "src", "Joints"),
            "-I", os.path.join(root, "include", "Physecs"),
            "-I", os.path.join(root, "include", "Physecs", "Joints"),
            "-I", os.path.join(root, "include"),
            "-I", glm_inc, "-I", entt_inc,
        ]
        objs = []
        for src in srcs:
            obj = os.path.join(root, os.path.basename(src) + ".o")
            objs.append(obj)
            jobs.append((["g++", "-c", src, "-o", obj] + [f for f in flags if f != "-shared"], None))
        lib = os.path.join(OUT, f"libphysecs_ref{variant}.so")
        jobs.append((None, (["g++", "-shared", "-pthread", "-o", lib] + objs, lib)))
    # compile in parallel batches, link after
    procs = []
    links = []
    maxj = max(1, (os.cpu_count() or 4))
    for cmd, link in jobs:
        if link is not None:
            links.append(link)
            continue
        while len(procs) >= maxj:
            procs = _reap(procs)
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT), cmd))
    while procs:
        procs = _reap(procs, block=True)
    for cmd, lib in links:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode:
            raise RuntimeError("link failed:\n" + r.stdout.decode())
        if verbose:
            print("built", lib)


def _reap(procs, block=False):
    keep = []
    for p, cmd in procs:
        rc = p.wait() if block else p.poll()
        if rc is None:
            keep.append((p, cmd))
        elif rc != 0:
            out = p.stdout.read().decode()
            raise RuntimeError("compile failed: " + " ".join(cmd) + "\n" + out[-6000:])
    if not block and len(keep) == len(procs):
        import time
        time.sleep(0.05)
    return keep


if __name__ == "__main__":
    build()
    sys.exit(0)
