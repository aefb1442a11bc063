"""Builds the recording test double of the C ABI (pb_recorder.cpp) and the host layer (physecs_b200/host/*.cpp, unchanged) linked
against it, into tests/abi_recorder/_build/ (git-ignored).  Test infrastructure only: see the header of pb_recorder.cpp."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build")
sys.path.insert(0, ROOT)
import build as product_build  # noqa: E402  (include paths and the host source list)

RECORDER = os.path.join(OUT, "libpb_recorder.so")
SCENE = os.path.join(OUT, "libphysecs_b200_scene_recorder.so")


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout.decode()[-4000:])


def build():
    """Returns (recorder library, host layer over it), or None when EnTT / GLM headers are not available."""
    entt, glm = product_build.find_ecs_includes()
    if not entt or not glm:
        return (RECORDER, SCENE) if os.path.exists(RECORDER) and os.path.exists(SCENE) else None
    os.makedirs(OUT, exist_ok=True)
    host = os.path.join(ROOT, "physecs_b200", "host")
    inc = os.path.join(ROOT, "include")
    srcs = [os.path.join(host, f) for f in product_build.HOST_SRC]
    rec_src = os.path.join(HERE, "pb_recorder.cpp")
    tri_src = os.path.join(ROOT, "physecs_b200", "csrc", "trimesh_build.cpp")      # setup-time host code of the product: the triangle-mesh BVH build
    deps = srcs + [rec_src, tri_src, os.path.abspath(__file__), os.path.join(inc, "physecs_b200.h")] + \
           [os.path.join(dp, f) for dp, _, fns in os.walk(os.path.join(inc, "Physecs")) for f in fns]
    if os.path.exists(RECORDER) and os.path.exists(SCENE) and all(os.path.getmtime(d) <= min(os.path.getmtime(RECORDER), os.path.getmtime(SCENE)) for d in deps):
        return RECORDER, SCENE
    _run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", rec_src, tri_src, "-o", RECORDER])
    _run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-DGLM_FORCE_INLINE", "-I", os.path.join(inc, "Physecs"),
          "-I", os.path.join(inc, "Physecs", "Joints"), "-I", inc, "-I", glm, "-I", entt] + srcs +
         ["-o", SCENE, "-L", OUT, "-lpb_recorder", "-Wl,-rpath,$ORIGIN", "-lpthread"])
    return RECORDER, SCENE


if __name__ == "__main__":
    print(build())
