"""ctypes wrapper of the oracle harness (oracle/ref_harness.cpp + the reference compiled by oracle/build_ref.py).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (physecs_b200/) never imports this module.
Parity status: the reference ships no tests / golden vectors, so the oracle is "pinned" by being
the reference implementation itself, compiled from /root/reference (oracle/_ref/*.so).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path(hashfix=False):
    return os.path.join(_HERE, "_ref", "libphysecs_ref_hashfix.so" if hashfix else "libphysecs_ref.so")


_libs = {}


def load(hashfix=False):
    if hashfix in _libs:
        return _libs[hashfix]
    p = lib_path(hashfix)
    if not os.path.exists(p):
        raise RuntimeError(f"{p} missing: run `python oracle/build_ref.py` where /root/reference is available")
    lib = C.CDLL(p)
    lib.ph_create.restype = C.c_void_p
    lib.ph_simulate.restype = C.c_double
    lib.ph_destroy.restype = None
    _libs[hashfix] = lib
    return lib


def available():
    return os.path.exists(lib_path(False)) and os.path.exists(lib_path(True))


def _p(a, ct=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _f(a):
    return np.ascontiguousarray(a, np.float32)


def _i(a):
    return np.ascontiguousarray(a, np.int32)


class RefScene:
    """The reference physecs::Scene over an entt::registry filled from a SceneDesc (physecs_b200.scenes)."""

    def __init__(self, desc, num_threads=0, hashfix=True):
        self.lib = load(hashfix)
        self.desc = desc
        self.h = C.c_void_p(self.lib.ph_create(int(num_threads)))
        d = desc
        for m in d.convex:
            self.lib.ph_add_convex(self.h, _p(_f(m.verts)), len(m.verts), _p(_i(m.face_offsets), C.c_int), _p(_i(m.face_indices), C.c_int),
                                   len(m.face_offsets) - 1, _p(_f(m.face_normals)), _p(_f(m.face_centroids)))
        for m in d.trimesh:
            idx = np.ascontiguousarray(m.indices, np.uint32)
            self.lib.ph_add_trimesh(self.h, _p(_f(m.verts)), len(m.verts), _p(idx, C.c_uint), len(idx))
        self.lib.ph_add_entities(self.h, d.n, _p(_f(d.pos)), _p(_f(d.quat)), _p(_i(d.flags), C.c_int), _p(_f(d.vel)), _p(_f(d.angvel)),
                                 _p(_f(d.inv_mass)), _p(_f(d.com)), _p(_f(d.inv_inertia)), _p(_i(d.col_offsets), C.c_int), _p(_f(d.col_lpos)),
                                 _p(_f(d.col_lquat)), _p(_i(d.col_type), C.c_int), _p(_f(d.col_params)), _p(_i(d.col_mesh), C.c_int),
                                 _p(_f(d.col_material)), _p(_i(d.col_flags), C.c_int), _p(_i(d.col_data), C.c_int))
        self.joint_colors = []
        for (t, e0, a0p, a0q, e1, a1p, a1q, prm) in d.joints:
            c = self.lib.ph_add_joint(self.h, int(t), int(e0), _p(_f(a0p)), _p(_f(a0q)), int(e1), _p(_f(a1p)), _p(_f(a1q)), _p(_f(prm)))
            self.joint_colors.append(c)
        for (e0, e1) in d.no_collide:
            self.lib.ph_set_can_collide(self.h, int(e0), int(e1), 0)
        self.lib.ph_set_params(self.h, int(d.substeps), int(d.iterations), C.c_float(d.gravity))

    def close(self):
        if self.h:
            self.lib.ph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def presort(self):
        """Stable-sort the Scene's broadphase entries by bounds.min.x, as its own first-step insertion sort would (O(n log n)
        instead of O(n^2)); makes the oracle usable on freshly created 100k..1M-collider scenes."""
        self.lib.ph_presort(self.h)

    def simulate(self, dt=None):
        """One Scene::simulate; returns its wall time in ms."""
        return float(self.lib.ph_simulate(self.h, C.c_float(self.desc.dt if dt is None else dt)))

    def get_state(self):
        n = self.lib.ph_num_entities(self.h)
        pos = np.zeros((n, 3), np.float32); quat = np.zeros((n, 4), np.float32)
        vel = np.zeros((n, 3), np.float32); ang = np.zeros((n, 3), np.float32)
        self.lib.ph_get_state(self.h, _p(pos), _p(quat), _p(vel), _p(ang))
        return pos, quat, vel, ang

    def set_state(self, ents, pos, quat, vel=None, angvel=None, patch=False):
        e = _i(ents)
        self.lib.ph_set_state(self.h, len(e), _p(e, C.c_int), _p(_f(pos)), _p(_f(quat)), _p(_f(vel)) if vel is not None else None,
                              _p(_f(angvel)) if angvel is not None else None, int(patch))

    def pairs(self):
        n = self.lib.ph_num_pairs(self.h)
        out = np.zeros((max(n, 1), 4), np.int32)
        self.lib.ph_get_pairs(self.h, _p(out, C.c_int))
        return out[:n]

    def bounds(self):
        n = self.lib.ph_num_bounds(self.h)
        ids = np.zeros((max(n, 1), 2), np.int32); b = np.zeros((max(n, 1), 6), np.float32)
        self.lib.ph_get_bounds(self.h, _p(ids, C.c_int), _p(b))
        return ids[:n], b[:n]

    def manifold_keys(self):
        n = self.lib.ph_num_manifolds(self.h)
        out = np.zeros((max(n, 1), 5), np.int32)
        self.lib.ph_get_manifold_keys(self.h, _p(out, C.c_int))
        return out[:n]

    def set_manifold_order(self, keys5):
        k = _i(keys5)
        self._order = k
        self.lib.ph_set_manifold_order(self.h, _p(k, C.c_int), len(k))

    def order_stats(self):
        out = np.zeros(3, np.int32)
        self.lib.ph_get_order_stats(self.h, _p(out, C.c_int))
        return tuple(int(x) for x in out)   # matched, missing, extra

    def narrowphase(self, pairs4):
        """physecs::collision on the current state for pairs (e0,c0,e1,c1). Returns dict like Context.manifolds()."""
        pr = _i(pairs4).reshape(-1, 4)
        cap = max(8 * len(pr), 64) if len(pr) < 100_000 else 2 * len(pr)   # retried below if a mesh pair list outgrows it
        while True:
            keys = np.zeros((cap, 5), np.int32); nrm = np.zeros((cap, 3), np.float32); pts = np.zeros((cap, 4, 2, 3), np.float32)
            m = self.lib.ph_narrowphase(self.h, _p(pr, C.c_int), len(pr), cap, _p(keys, C.c_int), _p(nrm), _p(pts))
            if m <= cap:
                break
            cap = m
        keys = keys[:m]
        full = np.concatenate([pr[keys[:, 0]], keys[:, 1:2]], 1) if m else np.zeros((0, 5), np.int32)
        return dict(keys=full, num_points=keys[:, 2].copy(), normal=nrm[:m], points=pts[:m])

    def add_entities(self, d):
        """Append the entities of another SceneDesc to the live registry (same meshes); returns the first new entity id."""
        first = self.lib.ph_add_entities(self.h, d.n, _p(_f(d.pos)), _p(_f(d.quat)), _p(_i(d.flags), C.c_int), _p(_f(d.vel)), _p(_f(d.angvel)),
                                         _p(_f(d.inv_mass)), _p(_f(d.com)), _p(_f(d.inv_inertia)), _p(_i(d.col_offsets), C.c_int), _p(_f(d.col_lpos)),
                                         _p(_f(d.col_lquat)), _p(_i(d.col_type), C.c_int), _p(_f(d.col_params)), _p(_i(d.col_mesh), C.c_int),
                                         _p(_f(d.col_material)), _p(_i(d.col_flags), C.c_int), _p(_i(d.col_data), C.c_int))
        self._extra = getattr(self, "_extra", 0) + d.n
        return first

    def destroy_entity(self, e):
        self.lib.ph_destroy_entity(self.h, int(e))

    def sort_dynamic(self, greater_first=True):
        """registry.sort<RigidBodyDynamicComponent> by entity id, comparator a > b (True) / a < b (False); see scene_api.HostScene.sort_dynamic."""
        self.lib.ph_sort_dynamic(self.h, int(greater_first))

    def add_collider(self, e, lpos, lquat, ctype, params, mesh=-1, material=(0.4, 0.2, 0.0), flags=2, data=0):
        """Scene::addCollider on a live entity (Physecs.cpp:738-751; flags: bit 0 trigger, bit 1 enableSimulation)."""
        prm = _f(list(params) + [0.0] * (4 - len(params)))
        self.lib.ph_add_collider(self.h, int(e), _p(_f(lpos)), _p(_f(lquat)), int(ctype), _p(prm), int(mesh), _p(_f(material)), int(flags), int(data))

    def clear_colliders(self, e):
        """Scene::clearColliders (Physecs.cpp:725-736)."""
        self.lib.ph_clear_colliders(self.h, int(e))

    def add_joint(self, t, e0, a0p, a0q, e1, a1p, a1q, prm):
        return self.lib.ph_add_joint(self.h, int(t), int(e0), _p(_f(a0p)), _p(_f(a0q)), int(e1), _p(_f(a1p)), _p(_f(a1q)), _p(_f(prm)))

    def destroy_joint(self, j):
        self.lib.ph_destroy_joint(self.h, int(j))

    def set_revolute_drive(self, j, enabled, velocity, max_torque):
        self.lib.ph_set_revolute_drive(self.h, int(j), int(enabled), C.c_float(velocity), C.c_float(max_torque))

    def set_kinematic(self, e, kin):
        self.lib.ph_set_kinematic(self.h, int(e), int(kin))

    def set_can_collide(self, e0, e1, can):
        self.lib.ph_set_can_collide(self.h, int(e0), int(e1), int(can))

    def raycast(self, orig, direction, max_dist, mod=0, skip=0):
        hit = np.zeros(3, np.float32)
        e = self.lib.ph_raycast(self.h, _p(_f(orig)), _p(_f(direction)), C.c_float(max_dist), int(mod), int(skip), _p(hit))
        return e, hit

    def overlap(self, pos, quat, gtype, params, mesh=-1, flt=0):
        cap = 4096
        out = np.zeros((cap, 2), np.int32)
        prm = _f(list(params) + [0.0] * (4 - len(params)))
        n = self.lib.ph_overlap(self.h, _p(_f(pos)), _p(_f(quat)), int(gtype), _p(prm), int(mesh), int(flt), cap, _p(out, C.c_int))
        return out[:n]

    def overlap_mtd(self, pos, quat, gtype, params, mesh=-1):
        """Scene::overlapWithMinTranslationalDistance: (rows (entity, colIndex), rows (normal xyz, mtd))."""
        cap = 4096
        ids = np.zeros((cap, 2), np.int32); val = np.zeros((cap, 4), np.float32)
        prm = _f(list(params) + [0.0] * (4 - len(params)))
        n = self.lib.ph_overlap_mtd(self.h, _p(_f(pos)), _p(_f(quat)), int(gtype), _p(prm), int(mesh), cap, _p(ids, C.c_int), _p(val))
        return ids[:n], val[:n]

    def triggers(self):
        """Overlapping trigger pairs of the last simulate, rows (e0, c0, e1, c1)."""
        n = self.lib.ph_num_triggers(self.h)
        out = np.zeros((max(n, 1), 4), np.int32)
        self.lib.ph_get_triggers(self.h, _p(out, C.c_int))
        return out[:n]

    def record_trigger_events(self):
        self.lib.ph_record_trigger_events(self.h)

    def take_trigger_events(self):
        """Listener calls since the last take, rows (0 enter | 1 exit, e0, c0, e1, c1)."""
        cap = 1 << 16
        out = np.zeros((cap, 5), np.int32)
        n = self.lib.ph_take_trigger_events(self.h, _p(out, C.c_int), cap)
        return out[:min(n, cap)]

    def set_contact_filter(self, mode):
        self.lib.ph_set_contact_filter(self.h, int(mode))

    def trimesh(self, mesh_id=0):
        nt, nn = C.c_int(), C.c_int()
        self.lib.ph_trimesh_sizes(self.h, mesh_id, C.byref(nt), C.byref(nn))
        tri = np.zeros((nt.value, 3), np.uint32); nrm = np.zeros((nt.value, 3), np.float32)
        nb = np.zeros((nn.value, 6), np.float32); ci = np.zeros((nn.value, 2), np.int32)
        self.lib.ph_trimesh_get(self.h, mesh_id, _p(tri, C.c_uint), _p(nrm), _p(nb), _p(ci, C.c_int))
        return tri, nrm, nb, ci
