"""ctypes binding of libphysecs_b200.so (include/physecs_b200.h) + a small Context helper.

This is the Python face of the C ABI used by tests/ and bench.py.  There is no CPU fallback: if the CUDA
library is missing or no device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import scenes as S

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHYSECS_B200_LIB") or os.path.join(_HERE, "lib", "libphysecs_b200.so")

PB_OK, PB_ECUDA, PB_ECAPACITY, PB_EINVAL, PB_EUNSUPPORTED = 0, 1, 2, 3, 4
# pb_counts.cause bits (include/physecs_b200.h)
PB_CAUSE_PAIRS, PB_CAUSE_MANIFOLDS, PB_CAUSE_TRIGGERS, PB_CAUSE_WALK_STACK = 0x1, 0x2, 0x4, 0x8
PB_CAUSE_SPILLED_EPA_FACES, PB_CAUSE_SPILLED_EPA_LOOSE, PB_CAUSE_SPILLED_EPA_VERTS, PB_CAUSE_SPILLED_CLIP = 0x10, 0x20, 0x40, 0x80
PB_CAUSE_SPILLED_TRI_CAND, PB_CAUSE_SPILLED_TRI_CONTACTS, PB_CAUSE_SPILLED_MESH_STACK = 0x100, 0x200, 0x400
PB_CAUSE_SPILL_LIST, PB_CAUSE_SPILL_SCRATCH = 0x1000, 0x2000

EXPORTS = [
    "pb_ctx_create", "pb_ctx_destroy", "pb_last_error", "pb_host_alloc", "pb_host_free", "pb_stream",
    "pb_upload_bodies", "pb_upload_colliders", "pb_register_convex", "pb_register_trimesh", "pb_upload_joints",
    "pb_set_noncolliding_pairs", "pb_set_state", "pb_move_rows", "pb_refresh_bounds", "pb_step", "pb_get_state", "pb_sync",
    "pb_get_counts", "pb_get_timings", "pb_get_pairs", "pb_get_bounds", "pb_get_manifolds", "pb_build_trimesh",
    "pb_set_profile", "pb_get_profile", "pb_get_launches", "pb_profiler_range",
    "pb_update_joint_params", "pb_set_contact_filter", "pb_set_kinematic", "pb_set_mass", "pb_set_bounds", "pb_keep_bounds_begin", "pb_keep_bounds",
    "pb_set_static_poses", "pb_get_triggers", "pb_keep_contact_cache", "pb_keep_joint_state", "pb_grow_arenas", "pb_query_raycast", "pb_query_overlap",
    "pb_debug_sort_pairs",
]


class Caps(C.Structure):
    _fields_ = [("max_bodies", C.c_int), ("max_colliders", C.c_int), ("max_pairs", C.c_int), ("max_manifolds", C.c_int),
                ("max_joints", C.c_int), ("reserved", C.c_int * 3)]


class Counts(C.Structure):
    _fields_ = [("n_pairs", C.c_int), ("n_manifolds", C.c_int), ("n_points", C.c_int), ("n_colors", C.c_int), ("n_overflow", C.c_int),
                ("status", C.c_int), ("n_mesh_pairs", C.c_int), ("n_triggers", C.c_int), ("cause", C.c_int), ("n_spilled", C.c_int)]


class Timings(C.Structure):
    _fields_ = [("broadphase", C.c_float), ("narrowphase", C.c_float), ("contact_build", C.c_float), ("solve", C.c_float),
                ("total", C.c_float), ("solve_kernel", C.c_float), ("reserved", C.c_float * 2)]


_lib = None


def load_library():
    """Load the CUDA library; fail loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python build.py` (CUDA extension is mandatory)")
    lib = C.CDLL(LIB_PATH)
    lib.pb_last_error.restype = C.c_char_p
    lib.pb_stream.restype = C.c_void_p
    lib.pb_ctx_destroy.restype = None
    lib.pb_host_free.restype = None
    lib.pb_get_launches.restype = C.c_ulonglong
    _lib = lib
    return lib


def _p(a, ct=C.c_float):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ct))


def _f(a):
    return np.ascontiguousarray(a, np.float32)


def _i(a):
    return np.ascontiguousarray(a, np.int32)


def build_trimesh(verts, indices):
    """Host-side registration-time BVH build (same construction as the reference TriangleMesh ctor)."""
    lib = load_library()
    v = _f(verts); idx = np.ascontiguousarray(indices, np.uint32)
    nt = len(idx) // 3
    tri = np.zeros((nt, 3), np.uint32); orig = np.zeros(nt, np.int32)
    nb = np.zeros((2 * nt, 6), np.float32); ci = np.zeros((2 * nt, 2), np.int32)
    nn = C.c_int()
    rc = lib.pb_build_trimesh(_p(v), len(v), _p(idx, C.c_uint), len(idx), _p(tri, C.c_uint), _p(orig, C.c_int), _p(nb), _p(ci, C.c_int), C.byref(nn))
    if rc != PB_OK:
        raise RuntimeError("pb_build_trimesh failed")
    return tri, orig, nb[:nn.value], ci[:nn.value]


class PbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"physecs_b200 error {code}: {msg}")
        self.code = code


class Context:
    """Device-resident scene behind the C ABI."""

    def __init__(self, desc: Optional[S.SceneDesc] = None, device=0, max_pairs=None, max_manifolds=None, max_bodies=None, max_colliders=None):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        self.desc = None
        n = desc.n if desc is not None else (max_bodies or 1024)
        nc = len(desc.col_type) if desc is not None else (max_colliders or n)
        caps = Caps()
        caps.max_bodies = max_bodies or max(n, 16)
        caps.max_colliders = max_colliders or max(nc, 16)
        caps.max_pairs = max_pairs or max(16 * nc, 4096)
        caps.max_manifolds = max_manifolds or max(8 * nc, 4096)
        caps.max_joints = max(len(desc.joints) if desc is not None else 0, 16)
        self.caps = caps
        rc = self.lib.pb_ctx_create(int(device), C.byref(caps), C.byref(self.ctx))
        if rc != PB_OK:
            raise PbError(rc, "pb_ctx_create failed (is a CUDA device visible? there is no CPU fallback)")
        if desc is not None:
            self.upload(desc)

    @classmethod
    def from_handle(cls, ptr, desc):
        """Non-owning view of an existing pb_ctx* (the host Scene's context): parity taps only."""
        self = cls.__new__(cls)
        self.lib = load_library()
        self.ctx = C.c_void_p(ptr)
        self.desc = desc
        self._borrowed = True
        return self

    def close(self):
        if self.ctx and not getattr(self, "_borrowed", False):
            self.lib.pb_ctx_destroy(self.ctx)
        self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != PB_OK:
            raise PbError(rc, self.lib.pb_last_error(self.ctx).decode())

    # ---- scene upload ------------------------------------------------------------------------------------------
    def upload(self, d: S.SceneDesc):
        self.desc = d
        dyn = d.dynamic_entities()
        sta = d.static_entities()
        order = np.concatenate([dyn, sta]).astype(np.int32)
        self.row_entity = order
        self.entity_row = np.empty(d.n, np.int32)
        self.entity_row[order] = np.arange(d.n, dtype=np.int32)
        self.n_dyn = len(dyn)
        self.dyn_entities = dyn
        self.tri_order = []
        for m in d.convex:
            h = C.c_int()
            self._check(self.lib.pb_register_convex(self.ctx, _p(_f(m.verts)), len(m.verts), _p(_i(m.face_offsets), C.c_int),
                                                    _p(_i(m.face_indices), C.c_int), len(m.face_offsets) - 1, _p(_f(m.face_normals)),
                                                    _p(_f(m.face_centroids)), C.byref(h)))
        for m in d.trimesh:
            h = C.c_int()
            order_out = np.zeros(len(m.indices) // 3, np.int32)
            idx = np.ascontiguousarray(m.indices, np.uint32)
            self._check(self.lib.pb_register_trimesh(self.ctx, _p(_f(m.verts)), len(m.verts), _p(idx, C.c_uint), len(idx), C.byref(h),
                                                     _p(order_out, C.c_int)))
            self.tri_order.append(order_out)
        pos = _f(d.pos[order]); quat = _f(d.quat[order])
        kin = _i((d.flags[dyn] & S.F_KINEMATIC) != 0)
        self._check(self.lib.pb_upload_bodies(self.ctx, len(dyn), len(sta), _p(order, C.c_int), _p(pos), _p(quat), _p(kin, C.c_int),
                                              _p(_f(d.vel[dyn])), _p(_f(d.angvel[dyn])), _p(_f(d.inv_mass[dyn])), _p(_f(d.com[dyn])),
                                              _p(_f(d.inv_inertia[dyn]))))
        nc = len(d.col_type)
        counts = np.diff(d.col_offsets)
        col_entity = np.repeat(np.arange(d.n, dtype=np.int32), counts)
        col_index = (np.arange(nc, dtype=np.int32) - np.repeat(d.col_offsets[:-1], counts)).astype(np.int32)
        self.col_entity = col_entity
        self.col_index = col_index
        body_row = self.entity_row[col_entity]
        self._check(self.lib.pb_upload_colliders(self.ctx, nc, _p(_i(body_row), C.c_int), _p(col_index, C.c_int), _p(_f(d.col_lpos)),
                                                 _p(_f(d.col_lquat)), _p(_i(d.col_type), C.c_int), _p(_f(d.col_params)), _p(_i(d.col_mesh), C.c_int),
                                                 _p(_f(d.col_material)), _p(_i(d.col_flags), C.c_int), _p(_i(d.col_data), C.c_int)))
        if d.joints:
            self.upload_joints(d)
        nocoll = list(d.no_collide) + [(j[1], j[4]) for j in d.joints]
        if nocoll:
            arr = _i(np.array(nocoll, np.int32).reshape(-1, 2))
            self._check(self.lib.pb_set_noncolliding_pairs(self.ctx, len(arr), _p(arr, C.c_int)))

    def upload_joints(self, d: S.SceneDesc):
        from .joint_colors import color_joints
        nj = len(d.joints)
        colors = color_joints([(j[1], j[4]) for j in d.joints])
        t = _i([j[0] for j in d.joints])
        r0 = _i(self.entity_row[[j[1] for j in d.joints]]); r1 = _i(self.entity_row[[j[4] for j in d.joints]])
        a0p = _f(np.stack([j[2] for j in d.joints])); a0q = _f(np.stack([j[3] for j in d.joints]))
        a1p = _f(np.stack([j[5] for j in d.joints])); a1q = _f(np.stack([j[6] for j in d.joints]))
        prm = _f(np.stack([j[7] for j in d.joints]))
        self.joint_colors = colors
        self._check(self.lib.pb_upload_joints(self.ctx, nj, _p(t, C.c_int), _p(r0, C.c_int), _p(r1, C.c_int), _p(a0p), _p(a0q), _p(a1p), _p(a1q),
                                              _p(prm), _p(_i(colors), C.c_int)))

    # ---- stepping --------------------------------------------------------------------------------------------------
    def step(self, dt=None, substeps=None, iterations=None, gravity=None):
        d = self.desc
        self._check(self.lib.pb_step(self.ctx, C.c_float(d.dt if dt is None else dt), int(d.substeps if substeps is None else substeps),
                                     int(d.iterations if iterations is None else iterations), C.c_float(d.gravity if gravity is None else gravity)))

    def sync(self):
        self._check(self.lib.pb_sync(self.ctx))

    def set_state(self, pos=None, quat=None, vel=None, angvel=None):
        """Arrays over the dynamic rows (storage order)."""
        a = [None if x is None else _f(x) for x in (pos, quat, vel, angvel)]
        self._keep = a
        self._check(self.lib.pb_set_state(self.ctx, self.n_dyn, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3])))

    def set_state_entities(self, pos, quat, vel, angvel):
        """Arrays over all entities (oracle layout); only dynamic entities are pushed."""
        e = self.dyn_entities
        self.set_state(pos[e], quat[e], vel[e], angvel[e])

    def refresh_bounds(self):
        self._check(self.lib.pb_refresh_bounds(self.ctx))

    def move_rows(self, entities, pos, quat):
        rows = _i(self.entity_row[np.asarray(entities)])
        p, q = _f(pos), _f(quat)
        self._check(self.lib.pb_move_rows(self.ctx, len(rows), _p(rows, C.c_int), _p(p), _p(q)))

    def get_state(self):
        n = self.n_dyn
        pos = np.zeros((n, 3), np.float32); quat = np.zeros((n, 4), np.float32)
        vel = np.zeros((n, 3), np.float32); ang = np.zeros((n, 3), np.float32)
        self._check(self.lib.pb_get_state(self.ctx, _p(pos), _p(quat), _p(vel), _p(ang)))
        return pos, quat, vel, ang

    def get_state_entities(self):
        """State scattered back to entity order (statics keep their description values)."""
        d = self.desc
        pos, quat, vel, ang = self.get_state()
        P = d.pos.copy(); Q = d.quat.copy(); V = np.zeros_like(d.vel); W = np.zeros_like(d.angvel)
        e = self.dyn_entities
        P[e], Q[e], V[e], W[e] = pos, quat, vel, ang
        return P, Q, V, W

    # ---- taps ----------------------------------------------------------------------------------------------------------
    def counts(self) -> Counts:
        c = Counts()
        self._check(self.lib.pb_get_counts(self.ctx, C.byref(c)))
        return c

    BIN_NAMES = ["sphere_sphere", "sphere_capsule", "capsule_capsule", "sphere_box", "capsule_box", "box_box", "gjk_epa", "mesh_sphere", "mesh_capsule",
                 "mesh_box", "mesh_convex", "trigger"]

    def bin_counts(self):
        """candidate pairs of the last step per narrowphase bin"""
        out = (C.c_int * 12)()
        self._check(self.lib.pb_get_bin_counts(self.ctx, out))
        return {n: int(out[i]) for i, n in enumerate(self.BIN_NAMES)}

    def timings(self) -> Timings:
        t = Timings()
        self._check(self.lib.pb_get_timings(self.ctx, C.byref(t)))
        return t

    def set_profile(self, on=True):
        self._check(self.lib.pb_set_profile(self.ctx, int(on)))

    def profile(self):
        ms = (C.c_double * 8)(); cnt = (C.c_longlong * 8)()
        self._check(self.lib.pb_get_profile(self.ctx, ms, cnt))
        # (milliseconds, number of phases) per phase kind of the persistent substep kernel
        names = ["integrate_v", "prep", "contact_pass", "joint_solve", "integrate_x", "local_sweeps"]
        out = {n: (ms[i], cnt[i]) for i, n in enumerate(names)}
        out["k_substep_solve_last_step"] = (ms[6], cnt[6])      # CUDA events around the launches of the last step: (ms summed, launches)
        return out

    def profile_colors(self):
        """(ms, phases) accumulated per contact colour since profiling was switched on."""
        ms = (C.c_double * 64)(); cnt = (C.c_longlong * 64)()
        self._check(self.lib.pb_get_profile_colors(self.ctx, ms, cnt))
        return [(ms[i], cnt[i]) for i in range(64)]

    def set_islands(self, mode):
        """0 off, 1 on, 2 auto: small simulation islands solved inside one CTA each (results are identical either way)."""
        self._check(self.lib.pb_set_islands(self.ctx, int(mode)))

    def set_deterministic(self, on=True):
        """colours from fixed priorities: two runs of the same scene give bit-identical states (reference: numThreads = 0)"""
        self._check(self.lib.pb_set_deterministic(self.ctx, int(on)))

    def island_stats(self):
        out = (C.c_int * 3)()
        self._check(self.lib.pb_get_island_stats(self.ctx, out))
        return dict(on=bool(out[0]), local=out[1], total=out[2])

    def broadphase_info(self):
        out = (C.c_int * 3)()
        self._check(self.lib.pb_get_broadphase_info(self.ctx, out))
        return dict(all_pairs=bool(out[0]), tile_hits=out[1], tiles=out[2])

    def launches(self):
        return int(self.lib.pb_get_launches(self.ctx))

    def stream_ptr(self):
        return int(self.lib.pb_stream(self.ctx))

    def pairs(self):
        n = C.c_int()
        cap = self.counts().n_pairs
        out = np.zeros((max(cap, 1), 4), np.int32)
        self._check(self.lib.pb_get_pairs(self.ctx, _p(out, C.c_int), cap, C.byref(n)))
        return out[:n.value]

    def bounds(self):
        out = np.zeros((len(self.desc.col_type), 6), np.float32)
        self._check(self.lib.pb_get_bounds(self.ctx, _p(out)))
        return out

    def triggers(self):
        """Overlapping trigger pairs of the last step, rows (entity0, colIdx0, entity1, colIdx1)."""
        cap = max(self.counts().n_triggers, 1)
        out = np.zeros((cap, 4), np.int32)
        n = C.c_int()
        self._check(self.lib.pb_get_triggers(self.ctx, _p(out, C.c_int), cap, C.byref(n)))
        return out[:n.value]

    def set_contact_filter(self, fn=None):
        """Tabulate a contact filter fn(isTrigger0, data0, isTrigger1, data1) -> bool (True = TRIGGER) over the (isTrigger, data)
        classes present in the scene and hand the table to the device (reference Scene::setContactFilter).  None = default."""
        d = self.desc
        if fn is None:
            self._check(self.lib.pb_set_contact_filter(self.ctx, 0, None, 0, None))
            return
        keys = list(zip((d.col_flags & S.COL_TRIGGER).astype(bool).tolist(), d.col_data.tolist()))
        classes = sorted(set(keys))
        index = {k: i for i, k in enumerate(classes)}
        cls = _i([index[k] for k in keys])
        K = len(classes)
        lut = np.zeros((K, K), np.uint8)
        for i, (t0, d0) in enumerate(classes):
            for j, (t1, d1) in enumerate(classes):
                lut[i, j] = 1 if fn(t0, d0, t1, d1) else 0
        self._check(self.lib.pb_set_contact_filter(self.ctx, len(cls), _p(cls, C.c_int), K, lut.ctypes.data_as(C.POINTER(C.c_ubyte))))

    def manifolds(self):
        cap = max(self.counts().n_manifolds, 1)
        keys = np.zeros((cap, 5), np.int32); npts = np.zeros(cap, np.int32); nrm = np.zeros((cap, 3), np.float32)
        pts = np.zeros((cap, 4, 2, 3), np.float32); col = np.zeros(cap, np.int32)
        n = C.c_int()
        self._check(self.lib.pb_get_manifolds(self.ctx, cap, _p(keys, C.c_int), _p(npts, C.c_int), _p(nrm), _p(pts), _p(col, C.c_int), C.byref(n)))
        k = n.value
        return dict(keys=keys[:k], num_points=npts[:k], normal=nrm[:k], points=pts[:k], color=col[:k])
