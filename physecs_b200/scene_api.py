"""ctypes face of the host C++ layer: physecs::Scene over an entt::registry (include/Physecs/Physecs.h of this repo,
physecs_b200/host/*.cpp), reached through the flat entry points of physecs_b200/host/scene_harness.cpp.

This is the public API an application uses -- registry components in, Scene::simulate, registry components out -- so
the parity tests and bench.py's scene-level end-to-end number go through it.  The library is built by build.py when
EnTT / GLM headers are available (they are the application's own dependency, not vendored here).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHYSECS_SCENE_LIB") or os.path.join(_HERE, "lib", "libphysecs_b200_scene.so")   # the override serves A/B runs of two builds of the host layer

EXPORTS = [
    "psh_create", "psh_destroy", "psh_last_error", "psh_add_convex", "psh_add_trimesh", "psh_add_entities", "psh_destroy_entity",
    "psh_add_collider", "psh_clear_colliders", "psh_add_joint", "psh_set_revolute_drive", "psh_destroy_joint", "psh_set_params",
    "psh_set_can_collide", "psh_set_kinematic", "psh_set_sync_mode", "psh_set_contact_filter", "psh_record_trigger_events",
    "psh_take_trigger_events", "psh_set_state", "psh_simulate", "psh_num_entities", "psh_get_state", "psh_native_context",
    "psh_get_stats", "psh_mass_props", "psh_set_arena_capacity", "psh_raycast", "psh_overlap", "psh_sort_dynamic",
]

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def _declare(lib):
    lib.psh_create.restype = C.c_void_p
    lib.psh_destroy.restype = None
    lib.psh_last_error.restype = C.c_char_p
    lib.psh_simulate.restype = C.c_double
    lib.psh_native_context.restype = C.c_void_p
    for f in ("psh_set_params", "psh_set_can_collide", "psh_set_kinematic", "psh_set_sync_mode", "psh_set_contact_filter",
              "psh_record_trigger_events", "psh_set_state", "psh_get_state", "psh_get_stats", "psh_mass_props", "psh_set_arena_capacity", "psh_sort_dynamic"):
        getattr(lib, f).restype = None
    return lib


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python build.py` where EnTT and GLM headers are available")
    capi.load_library()   # libphysecs_b200.so first (the scene library links against it)
    _lib = _declare(C.CDLL(LIB_PATH))
    return _lib


def load_other_build(path):
    """The same harness entry points from another build of the host layer (tests/abi_recorder: the host layer linked against a recording
    double of the C ABI, for the CPU tests of the host's bookkeeping).  Never used by the product path."""
    return _declare(C.CDLL(path))


def _p(a, ct=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _f(a):
    return np.ascontiguousarray(a, np.float32)


def _i(a):
    return np.ascontiguousarray(a, np.int32)


class SceneError(RuntimeError):
    pass


class HostScene:
    """physecs::Scene (this repo's) over a registry filled from a SceneDesc, entity id == description index."""

    def __init__(self, desc, num_threads=0, device=0, lib=None):
        self.lib = lib if lib is not None else load_library()
        self.desc = desc
        self.h = C.c_void_p(self.lib.psh_create(int(num_threads), int(device)))
        d = desc
        for m in d.convex:
            self.lib.psh_add_convex(self.h, _p(_f(m.verts)), len(m.verts), _p(_i(m.face_offsets), C.c_int), _p(_i(m.face_indices), C.c_int),
                                    len(m.face_offsets) - 1, _p(_f(m.face_normals)), _p(_f(m.face_centroids)))
        for m in d.trimesh:
            idx = np.ascontiguousarray(m.indices, np.uint32)
            self.lib.psh_add_trimesh(self.h, _p(_f(m.verts)), len(m.verts), _p(idx, C.c_uint), len(idx))
        self.lib.psh_add_entities(self.h, d.n, _p(_f(d.pos)), _p(_f(d.quat)), _p(_i(d.flags), C.c_int), _p(_f(d.vel)), _p(_f(d.angvel)),
                                  _p(_f(d.inv_mass)), _p(_f(d.com)), _p(_f(d.inv_inertia)), _p(_i(d.col_offsets), C.c_int), _p(_f(d.col_lpos)),
                                  _p(_f(d.col_lquat)), _p(_i(d.col_type), C.c_int), _p(_f(d.col_params)), _p(_i(d.col_mesh), C.c_int),
                                  _p(_f(d.col_material)), _p(_i(d.col_flags), C.c_int), _p(_i(d.col_data), C.c_int))
        self.joint_colors = []
        for (t, e0, a0p, a0q, e1, a1p, a1q, prm) in d.joints:
            self.joint_colors.append(self.lib.psh_add_joint(self.h, int(t), int(e0), _p(_f(a0p)), _p(_f(a0q)), int(e1), _p(_f(a1p)), _p(_f(a1q)), _p(_f(prm))))
        for (e0, e1) in d.no_collide:
            self.lib.psh_set_can_collide(self.h, int(e0), int(e1), 0)
        self.lib.psh_set_params(self.h, int(d.substeps), int(d.iterations), C.c_float(d.gravity))
        self.n = d.n

    def close(self):
        if self.h:
            self.lib.psh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self):
        return self.lib.psh_last_error(self.h).decode()

    def simulate(self, dt=None):
        ms = float(self.lib.psh_simulate(self.h, C.c_float(self.desc.dt if dt is None else dt)))
        if ms < 0:
            raise SceneError(self._err())
        return ms

    def get_state(self):
        n = self.lib.psh_num_entities(self.h)
        pos = np.zeros((n, 3), np.float32); quat = np.zeros((n, 4), np.float32)
        vel = np.zeros((n, 3), np.float32); ang = np.zeros((n, 3), np.float32)
        self.lib.psh_get_state(self.h, _p(pos), _p(quat), _p(vel), _p(ang))
        return pos, quat, vel, ang

    def set_state(self, ents, pos, quat, vel=None, angvel=None, patch=False):
        e = _i(ents)
        self.lib.psh_set_state(self.h, len(e), _p(e, C.c_int), _p(_f(pos)), _p(_f(quat)), _p(_f(vel)) if vel is not None else None,
                               _p(_f(angvel)) if angvel is not None else None, int(patch))

    def set_kinematic(self, e, kin):
        self.lib.psh_set_kinematic(self.h, int(e), int(kin))

    def set_can_collide(self, e0, e1, can):
        self.lib.psh_set_can_collide(self.h, int(e0), int(e1), int(can))

    def set_sync_mode(self, device_authoritative):
        self.lib.psh_set_sync_mode(self.h, int(device_authoritative))

    def set_arena_capacity(self, max_pairs, max_manifolds):
        self.lib.psh_set_arena_capacity(self.h, int(max_pairs), int(max_manifolds))

    def set_contact_filter(self, mode):
        self.lib.psh_set_contact_filter(self.h, int(mode))

    def record_trigger_events(self):
        self.lib.psh_record_trigger_events(self.h)

    def take_trigger_events(self):
        cap = 1 << 16
        out = np.zeros((cap, 5), np.int32)
        n = self.lib.psh_take_trigger_events(self.h, _p(out, C.c_int), cap)
        return out[:min(n, cap)]

    def add_entities(self, d):
        """Append the entities of another SceneDesc to the live registry; returns the first new entity id."""
        first = self.lib.psh_add_entities(self.h, d.n, _p(_f(d.pos)), _p(_f(d.quat)), _p(_i(d.flags), C.c_int), _p(_f(d.vel)), _p(_f(d.angvel)),
                                          _p(_f(d.inv_mass)), _p(_f(d.com)), _p(_f(d.inv_inertia)), _p(_i(d.col_offsets), C.c_int), _p(_f(d.col_lpos)),
                                          _p(_f(d.col_lquat)), _p(_i(d.col_type), C.c_int), _p(_f(d.col_params)), _p(_i(d.col_mesh), C.c_int),
                                          _p(_f(d.col_material)), _p(_i(d.col_flags), C.c_int), _p(_i(d.col_data), C.c_int))
        self.n += d.n
        return first

    def destroy_entity(self, e):
        self.lib.psh_destroy_entity(self.h, int(e))

    def sort_dynamic(self, greater_first=True):
        """registry.sort<RigidBodyDynamicComponent> by entity id with the comparator a > b (True) or a < b (False).  EnTT iterates a pool back
        to front: a > b leaves a pool filled in creation order as it is, a < b REVERSES its packed order -- the order the body rows follow --
        behind the Scene's back (no signal fires)."""
        self.lib.psh_sort_dynamic(self.h, int(greater_first))

    def add_collider(self, e, lpos, lquat, ctype, params, mesh=-1, material=(0.4, 0.2, 0.0), flags=2, data=0):
        """Scene::addCollider on a live entity (flags: bit 0 trigger, bit 1 enableSimulation)."""
        prm = _f(list(params) + [0.0] * (4 - len(params)))
        self.lib.psh_add_collider(self.h, int(e), _p(_f(lpos)), _p(_f(lquat)), int(ctype), _p(prm), int(mesh), _p(_f(material)), int(flags), int(data))

    def clear_colliders(self, e):
        self.lib.psh_clear_colliders(self.h, int(e))

    def add_joint(self, t, e0, a0p, a0q, e1, a1p, a1q, prm):
        return self.lib.psh_add_joint(self.h, int(t), int(e0), _p(_f(a0p)), _p(_f(a0q)), int(e1), _p(_f(a1p)), _p(_f(a1q)), _p(_f(prm)))

    def destroy_joint(self, j):
        self.lib.psh_destroy_joint(self.h, int(j))

    def set_revolute_drive(self, j, enabled, velocity, max_torque):
        if self.lib.psh_set_revolute_drive(self.h, int(j), int(enabled), C.c_float(velocity), C.c_float(max_torque)) != 0:
            raise SceneError("joint is not a RevoluteJoint")

    def raycast(self, orig, direction, max_dist, mod=0, skip=0):
        """Scene::raycastClosest with the filter `entity % mod != skip` (mod 0 = accept all): (entity or -1, hit position)."""
        hit = np.zeros(3, np.float32)
        e = self.lib.psh_raycast(self.h, _p(_f(orig)), _p(_f(direction)), C.c_float(max_dist), int(mod), int(skip), _p(hit))
        if e == -2:
            raise SceneError(self._err())
        return e, hit

    def overlap(self, pos, quat, gtype, params, mesh=-1, flt=0):
        cap = 4096
        out = np.zeros((cap, 2), np.int32)
        prm = _f(list(params) + [0.0] * (4 - len(params)))
        n = self.lib.psh_overlap(self.h, _p(_f(pos)), _p(_f(quat)), int(gtype), _p(prm), int(mesh), int(flt), cap, _p(out, C.c_int))
        if n < 0:
            raise SceneError(self._err())
        return out[:n]

    def overlap_mtd(self, pos, quat, gtype, params, mesh=-1):
        """Scene::overlapWithMinTranslationalDistance: (rows (entity, colIndex), rows (normal xyz, mtd))."""
        cap = 4096
        ids = np.zeros((cap, 2), np.int32); val = np.zeros((cap, 4), np.float32)
        prm = _f(list(params) + [0.0] * (4 - len(params)))
        n = self.lib.psh_overlap_mtd(self.h, _p(_f(pos)), _p(_f(quat)), int(gtype), _p(prm), int(mesh), cap, _p(ids, C.c_int), _p(val))
        if n < 0:
            raise SceneError(self._err())
        return ids[:n], val[:n]

    def bvh_leaves(self):
        """Scene::getBVH walked from getBHVRootId (structure checked in the harness): rows (entity, colIndex), bounds rows."""
        cap = max(len(self.desc.col_type) + 64, 64)
        ids = np.zeros((cap, 2), np.int32); b = np.zeros((cap, 6), np.float32)
        n = self.lib.psh_bvh_leaves(self.h, cap, _p(ids, C.c_int), _p(b))
        if n < 0:
            raise SceneError(f"malformed BVH snapshot ({n}): {self._err()}")
        return ids[:n], b[:n]

    def stats(self):
        out = (C.c_double * 10)()
        self.lib.psh_get_stats(self.h, out)
        names = ["pairs", "manifolds", "points", "colors", "triggers", "device_ms", "gather_ms", "scatter_ms", "total_ms", "prepare_ms"]
        return {k: out[i] for i, k in enumerate(names)}

    def mass_props(self, e, mass):
        com = np.zeros(3, np.float32); inv = np.zeros(9, np.float32)
        self.lib.psh_mass_props(self.h, int(e), C.c_float(mass), _p(com), _p(inv))
        return com, inv

    def taps(self):
        """Parity taps on the Scene's device context (pairs / manifolds / triggers of the last step)."""
        ptr = self.lib.psh_native_context(self.h)
        if not ptr:
            raise SceneError("the Scene has no device context yet (call simulate first)")
        return capi.Context.from_handle(ptr, self.desc)
