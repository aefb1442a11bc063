"""-m "not gpu": source compatibility of the drop-in headers.  examples/dropin_app.cpp is an application written against the
reference's public API only (the idiom of its demo, joints, listeners, filter, edits, queries); it must build, link and run
 (a) against the reference's own headers and library (the oracle build's patched copy of the tree, oracle/_ref/libphysecs_ref.so), and
 (b) against this repo's include/Physecs + host layer -- here over the recording double of the C ABI (tests/abi_recorder: nothing is
     simulated without a device; what is checked is that every name the application uses exists with the same meaning) -- together
     with the reference's OWN CharacterController.{h,cpp}, copied unchanged to a scratch directory.
Needs /root/reference (EnTT / GLM headers, CharacterController, the reference tree); skipped where it is absent."""
import os
import shutil
import subprocess

import pytest

from oracle import build_ref
from oracle import ref as R
from tests.abi_recorder import build as recorder_build

product_build = recorder_build.product_build      # the repo's build.py (include paths of EnTT / GLM)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "examples", "dropin_app.cpp")
SCRATCH = "/tmp/physecs_dropin_test"

pytestmark = pytest.mark.skipif(not os.path.isdir(build_ref.REF) or None in product_build.find_ecs_includes(),
                                reason="reference tree / EnTT / GLM headers not available")


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, **kw)
    assert r.returncode == 0, " ".join(cmd) + "\n" + r.stdout.decode()[-4000:]
    return r.stdout.decode()


def _summary(out):
    line = [l for l in out.splitlines() if l.startswith("bodies ")][-1].split()
    return {"bodies": int(line[1]), "lowest_y": float(line[4]), "overlaps": int(line[line.index("overlaps") + 1]),
            "enter": int(line[line.index("enter") + 1]), "exit": int(line[line.index("exit") + 1])}


@pytest.fixture(scope="module")
def reference_run():
    if not R.available():
        pytest.skip("oracle/_ref not built")
    entt, glm = product_build.find_ecs_includes()
    root = build_ref.stage_tree(False)
    os.makedirs(SCRATCH, exist_ok=True)
    exe = os.path.join(SCRATCH, "dropin_ref")
    lib_dir = os.path.dirname(R.lib_path(False))
    _run(["g++", "-std=gnu++23", "-O1", "-w", "-DGLM_FORCE_INLINE", "-DENTT_PACKED_PAGE=1048576", "-DDROPIN_WITH_CHARACTER_CONTROLLER",
          "-include", os.path.join(build_ref.BUILD, "shim.h"), "-I", os.path.join(root, "include"), "-I", os.path.join(root, "include", "Physecs"),
          "-I", os.path.join(root, "include", "Physecs", "Joints"), "-I", os.path.join(root, "src"), "-I", glm, "-I", entt, APP, "-o", exe,
          "-L", lib_dir, "-lphysecs_ref", "-Wl,-rpath," + lib_dir, "-lpthread"])
    return _run([exe])


@pytest.fixture(scope="module")
def drop_in_run():
    built = recorder_build.build()
    assert built is not None
    entt, glm = product_build.find_ecs_includes()
    os.makedirs(SCRATCH, exist_ok=True)
    for f in ("include/Physecs/CharacterController.h", "src/CharacterController.cpp"):     # the reference's files, unchanged, outside the repo
        shutil.copyfile(os.path.join(build_ref.REF, f), os.path.join(SCRATCH, os.path.basename(f)))
    exe = os.path.join(SCRATCH, "dropin_b200")
    inc = os.path.join(ROOT, "include")
    out_dir = os.path.dirname(built[0])
    _run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-DGLM_FORCE_INLINE", "-DDROPIN_WITH_CHARACTER_CONTROLLER", "-I", inc,
          "-I", os.path.join(inc, "Physecs"), "-I", os.path.join(inc, "Physecs", "Joints"), "-I", SCRATCH, "-I", glm, "-I", entt, APP,
          os.path.join(SCRATCH, "CharacterController.cpp"), "-o", exe, "-L", out_dir, "-lphysecs_b200_scene_recorder", "-lpb_recorder",
          "-Wl,-rpath," + out_dir, "-lpthread"])
    return exe, _run([exe])


def test_app_builds_and_runs_against_the_reference(reference_run):
    s = _summary(reference_run)
    # the reference simulates: the bodies have come down onto the ground, the trigger volume has seen them
    assert s["bodies"] == 25 and 0.0 < s["lowest_y"] < 0.6 and s["enter"] > 0 and s["overlaps"] > 0


def test_app_and_reference_character_controller_build_against_the_drop_in_headers(drop_in_run):
    s = _summary(drop_in_run[1])
    # over the recording double nothing falls (its "step" only shifts x): the run proves the API surface, not the physics
    assert s["bodies"] == 25 and abs(s["lowest_y"] - 0.6) < 1e-6


def test_triangle_mesh_object_is_the_references(reference_run, drop_in_run):
    """physecs::TriangleMesh as the application sees it (public members `triangles`, `bvh`, `overlapBvh`): post-build triangle order, node
    table and query result of this repo's host object (csrc/trimesh_build.cpp) equal the reference's (TriangleMesh.cpp:99-192)."""
    mesh = lambda out: [l for l in out.splitlines() if l.startswith("mesh: ")]
    assert mesh(reference_run) and mesh(reference_run) == mesh(drop_in_run[1])


def test_without_a_device_the_application_gets_an_exception_not_a_cpu_path(drop_in_run):
    """The real failure mode of way A without a GPU: Scene::simulate throws (pb_ctx_create fails); the double can play that too."""
    r = subprocess.run([drop_in_run[0]], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=dict(os.environ, PB_RECORDER_NO_DEVICE="1"))
    assert r.returncode != 0 and b"no CPU fallback" in r.stdout
