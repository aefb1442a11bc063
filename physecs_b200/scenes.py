"""Flat scene descriptions + the synthetic scene generators of BASELINE.json's configs.

A SceneDesc is the array form of what an application would put in its entt::registry
(TransformComponent / RigidBodyCollisionComponent / RigidBodyDynamicComponent, reference
include/Physecs/Components.h:6-17, src/Transform.h:6-10).  The same description feeds the oracle harness
(oracle/ref.py) and the device context (physecs_b200/capi.py), so both sides see identical inputs.

Mass properties follow the reference's setup-time helper (src/MassUtil.cpp:6-28, :73-192) including its
box quirk (half extents used in the full-extent formula); they are inputs to both sides.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional

import numpy as np

SPHERE, CAPSULE, BOX, CONVEX_MESH, TRIANGLE_MESH = 0, 1, 2, 3, 4
J_FIXED, J_REVOLUTE, J_SPHERICAL, J_UNIVERSAL, J_PRISMATIC, J_GEAR, J_SERVO = range(7)
F_COLLISION, F_DYNAMIC, F_KINEMATIC = 1, 2, 4
COL_TRIGGER, COL_ENABLE_SIM = 1, 2

f32 = np.float32


@dataclasses.dataclass
class ConvexMeshDesc:
    verts: np.ndarray          # [nv,3]
    face_offsets: np.ndarray   # [nf+1]
    face_indices: np.ndarray
    face_normals: np.ndarray   # [nf,3]
    face_centroids: np.ndarray # [nf,3]


@dataclasses.dataclass
class TriMeshDesc:
    verts: np.ndarray    # [nv,3] float32
    indices: np.ndarray  # [3*nt] uint32


@dataclasses.dataclass
class SceneDesc:
    """Entities in creation order (entity id == index)."""
    pos: np.ndarray        # [n,3]
    quat: np.ndarray       # [n,4] xyzw
    flags: np.ndarray      # [n] F_* bits
    vel: np.ndarray        # [n,3]
    angvel: np.ndarray     # [n,3]
    inv_mass: np.ndarray   # [n]
    com: np.ndarray        # [n,3]
    inv_inertia: np.ndarray  # [n,9] column-major
    col_offsets: np.ndarray  # [n+1]
    col_lpos: np.ndarray     # [nc,3]
    col_lquat: np.ndarray    # [nc,4]
    col_type: np.ndarray     # [nc]
    col_params: np.ndarray   # [nc,4]
    col_mesh: np.ndarray     # [nc]
    col_material: np.ndarray # [nc,3] friction, restitution, damping
    col_flags: np.ndarray    # [nc]
    col_data: np.ndarray     # [nc]
    convex: List[ConvexMeshDesc] = dataclasses.field(default_factory=list)
    trimesh: List[TriMeshDesc] = dataclasses.field(default_factory=list)
    joints: list = dataclasses.field(default_factory=list)       # (type, e0, a0p, a0q, e1, a1p, a1q, params)
    no_collide: list = dataclasses.field(default_factory=list)   # (e0, e1)
    substeps: int = 8
    iterations: int = 2
    gravity: float = 9.81
    dt: float = 1.0 / 60.0
    name: str = ""

    @property
    def n(self):
        return len(self.pos)

    @property
    def n_dynamic(self):
        return int(np.count_nonzero(self.flags & F_DYNAMIC))

    def dynamic_entities(self):
        return np.nonzero(self.flags & F_DYNAMIC)[0].astype(np.int32)

    def static_entities(self):
        return np.nonzero((self.flags & F_DYNAMIC) == 0)[0].astype(np.int32)


def dynamic_only(desc, lift=(0.0, 0.0, 0.0)) -> SceneDesc:
    """The dynamic bodies of a description whose bodies hold one collider each (mixed_bin, terrain, ...), moved by `lift`: what a
    running application spawns into a live registry (HostScene.add_entities)."""
    assert np.array_equal(desc.col_offsets, np.arange(desc.n + 1)), "one collider per body expected"
    sel = desc.dynamic_entities()
    cut = lambda a: np.ascontiguousarray(a[sel])
    return dataclasses.replace(desc, pos=cut(desc.pos) + np.asarray(lift, np.float32), quat=cut(desc.quat), flags=cut(desc.flags), vel=cut(desc.vel),
                               angvel=cut(desc.angvel), inv_mass=cut(desc.inv_mass), com=cut(desc.com), inv_inertia=cut(desc.inv_inertia),
                               col_offsets=np.arange(len(sel) + 1, dtype=np.int32), col_lpos=cut(desc.col_lpos), col_lquat=cut(desc.col_lquat),
                               col_type=cut(desc.col_type), col_params=cut(desc.col_params), col_mesh=cut(desc.col_mesh),
                               col_material=cut(desc.col_material), col_flags=cut(desc.col_flags), col_data=cut(desc.col_data))


class SceneBuilder:
    def __init__(self, name=""):
        self.name = name
        self.ent = []
        self.cols = []
        self.convex = []
        self.trimesh = []
        self.joints = []
        self.no_collide = []

    def add_convex(self, m: ConvexMeshDesc):
        self.convex.append(m)
        return len(self.convex) - 1

    def add_trimesh(self, m: TriMeshDesc):
        self.trimesh.append(m)
        return len(self.trimesh) - 1

    def add_body(self, pos, quat=(0, 0, 0, 1), colliders=(), dynamic=True, kinematic=False, mass=1.0, vel=(0, 0, 0), angvel=(0, 0, 0)):
        """colliders: list of dicts(type, params, lpos, lquat, mesh, material, flags, data)"""
        e = len(self.ent)
        cols = []
        for c in colliders:
            d = dict(lpos=(0, 0, 0), lquat=(0, 0, 0, 1), mesh=-1, material=(0.4, 0.2, 0.0), flags=COL_ENABLE_SIM, data=0)
            d.update(c)
            p = list(d["params"]) + [0.0] * (4 - len(d["params"]))
            d["params"] = p
            cols.append(d)
        flags = (F_COLLISION if cols else 0) | (F_DYNAMIC if dynamic else 0) | (F_KINEMATIC if (dynamic and kinematic) else 0)
        com = np.zeros(3, f32)
        inv_i = np.zeros(9, f32)
        inv_m = f32(0)
        if dynamic:
            com, inv_inertia = compute_mass_props(cols, mass, self.convex)
            inv_i = inv_inertia.T.reshape(9)  # column-major
            inv_m = f32(1.0) / f32(mass)
        self.ent.append(dict(pos=pos, quat=quat, flags=flags, vel=vel, angvel=angvel, inv_mass=inv_m, com=com, inv_i=inv_i, cols=cols))
        return e

    def add_joint(self, jtype, e0, a0p, a0q, e1, a1p, a1q, params=()):
        p = list(params) + [0.0] * (8 - len(params))
        self.joints.append((jtype, e0, np.asarray(a0p, f32), np.asarray(a0q, f32), e1, np.asarray(a1p, f32), np.asarray(a1q, f32), np.asarray(p, f32)))

    def build(self, **kw) -> SceneDesc:
        n = len(self.ent)
        A = lambda key, w: np.ascontiguousarray(np.array([e[key] for e in self.ent], dtype=f32).reshape(n, w) if w > 1 else np.array([e[key] for e in self.ent], dtype=f32))
        cols = [c for e in self.ent for c in e["cols"]]
        nc = len(cols)
        offs = np.zeros(n + 1, np.int32)
        offs[1:] = np.cumsum([len(e["cols"]) for e in self.ent])
        C = lambda key, w, dt=f32: np.ascontiguousarray(np.array([c[key] for c in cols], dtype=dt).reshape(nc, w) if w > 1 else np.array([c[key] for c in cols], dtype=dt))
        d = SceneDesc(
            pos=A("pos", 3), quat=A("quat", 4), flags=np.array([e["flags"] for e in self.ent], np.int32), vel=A("vel", 3), angvel=A("angvel", 3),
            inv_mass=A("inv_mass", 1), com=A("com", 3), inv_inertia=A("inv_i", 9), col_offsets=offs,
            col_lpos=C("lpos", 3), col_lquat=C("lquat", 4), col_type=C("type", 1, np.int32), col_params=C("params", 4), col_mesh=C("mesh", 1, np.int32),
            col_material=C("material", 3), col_flags=C("flags", 1, np.int32), col_data=C("data", 1, np.int32),
            convex=self.convex, trimesh=self.trimesh, joints=self.joints, no_collide=self.no_collide, name=self.name)
        for k, v in kw.items():
            setattr(d, k, v)
        return d


def bulk_scene(name, pos, quat, flags, col_type, col_params, mass, material=(0.4, 0.2, 0.0), vel=None, angvel=None,
               col_mesh=None, trimesh=(), convex=(), **kw) -> SceneDesc:
    """Vectorised builder for the big configs: exactly one collider per entity at the body origin."""
    n = len(pos)
    pos = np.ascontiguousarray(pos, f32); quat = np.ascontiguousarray(quat, f32)
    flags = np.ascontiguousarray(flags, np.int32)
    col_type = np.ascontiguousarray(col_type, np.int32)
    col_params = np.ascontiguousarray(col_params, f32)
    dyn = (flags & F_DYNAMIC) != 0
    inv_mass = np.where(dyn, f32(1.0) / np.asarray(mass, f32), f32(0)).astype(f32)
    inv_i = np.zeros((n, 9), f32)
    m = np.broadcast_to(np.asarray(mass, f32), (n,))
    diag = single_shape_inertia_diag(col_type, col_params, m)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv_diag = np.where(dyn[:, None] & (diag != 0), f32(1.0) / diag, f32(0)).astype(f32)
    inv_i[:, 0] = inv_diag[:, 0]; inv_i[:, 4] = inv_diag[:, 1]; inv_i[:, 8] = inv_diag[:, 2]
    mat = np.broadcast_to(np.asarray(material, f32), (n, 3)).copy()
    d = SceneDesc(
        pos=pos, quat=quat, flags=flags | F_COLLISION,
        vel=np.zeros((n, 3), f32) if vel is None else np.ascontiguousarray(vel, f32),
        angvel=np.zeros((n, 3), f32) if angvel is None else np.ascontiguousarray(angvel, f32),
        inv_mass=inv_mass, com=np.zeros((n, 3), f32), inv_inertia=inv_i,
        col_offsets=np.arange(n + 1, dtype=np.int32), col_lpos=np.zeros((n, 3), f32),
        col_lquat=np.tile(np.array([0, 0, 0, 1], f32), (n, 1)), col_type=col_type, col_params=col_params,
        col_mesh=np.full(n, -1, np.int32) if col_mesh is None else np.ascontiguousarray(col_mesh, np.int32),
        col_material=mat, col_flags=np.full(n, COL_ENABLE_SIM, np.int32), col_data=np.zeros(n, np.int32),
        convex=list(convex), trimesh=list(trimesh), name=name)
    for k, v in kw.items():
        setattr(d, k, v)
    return d


# ---- mass properties (reference src/MassUtil.cpp) ---------------------------------------------------------------
def single_shape_inertia_diag(col_type, prm, mass):
    """Diagonal inertia of a single collider at the origin with identity local orientation (fp32 like the reference)."""
    n = len(col_type)
    out = np.ones((n, 3), f32)
    m = mass.astype(f32)
    s = col_type == SPHERE
    r = prm[:, 0]
    v = f32(2.0) * m * r * r / f32(5.0)
    out[s] = np.stack([v, v, v], 1)[s]
    c = col_type == CAPSULE
    hh, rr = prm[:, 0], prm[:, 1]
    pi = f32(math.pi)
    vc = f32(2) * rr * rr * pi * hh
    vs = f32(4) * rr ** 3 * pi / f32(3)
    vt = vc + vs
    with np.errstate(divide="ignore", invalid="ignore"):
        mc = vc * m / vt
        ms = vs * m / vt
    ixx = mc * (hh * hh / f32(3) + rr * rr / f32(4)) + ms * (hh * hh + f32(3) * hh * rr / f32(4) + f32(2) * rr * rr / f32(5))
    iyy = mc * rr * rr / f32(2) + ms * f32(2) * rr * rr / f32(5)
    out[c] = np.stack([ixx, iyy, ixx], 1)[c].astype(f32)
    b = col_type == BOX
    mm = m / f32(12)
    x, y, z = prm[:, 0] ** 2, prm[:, 1] ** 2, prm[:, 2] ** 2
    out[b] = np.stack([mm * (y + z), mm * (x + z), mm * (x + y)], 1)[b].astype(f32)
    return out


def quat_to_mat(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], dtype=np.float64)


def compute_mass_props(cols, mass, convex_meshes):
    """computeCOMAndInvInertiaTensor (reference src/MassUtil.cpp:73-185) for sphere/capsule/box/convex colliders."""
    pi = math.pi
    com = np.zeros(3)
    total = 0.0
    vols = []
    for c in cols:
        if c["flags"] & COL_TRIGGER:
            vols.append(0.0)
            continue
        t, p = c["type"], c["params"]
        if t == SPHERE:
            v = 4 * pi * p[0] ** 3 / 3
        elif t == CAPSULE:
            v = 4 * pi * p[1] ** 3 / 3 + p[1] * p[1] * pi * p[0] * 2
        elif t == BOX:
            v = p[0] * p[1] * p[2]
        elif t == CONVEX_MESH:
            v = 0.0
            for (vol, cen) in _convex_tets(convex_meshes[c["mesh"]], p):
                v += vol
            vols.append(v)
            cen_sum = np.zeros(3)
            for (vol, cen) in _convex_tets(convex_meshes[c["mesh"]], p):
                cen_sum += (np.asarray(c["lpos"], float) + cen) * vol
            com += cen_sum
            total += v
            continue
        else:
            vols.append(0.0)
            continue
        vols.append(v)
        com += np.asarray(c["lpos"], float) * v
        total += v
    if total == 0:
        return np.zeros(3, f32), np.eye(3, dtype=f32)
    com /= total
    I = np.zeros((3, 3))
    for c, v in zip(cols, vols):
        if c["flags"] & COL_TRIGGER or v == 0.0:
            continue
        t, p = c["type"], c["params"]
        m = mass * v / total
        r = com - np.asarray(c["lpos"], float)
        pa = m * (np.eye(3) * r.dot(r) - np.outer(r, r))
        if t == SPHERE:
            I += np.eye(3) * (2 * m * p[0] ** 2 / 5) + pa
        elif t == CAPSULE:
            hh, rr = p[0], p[1]
            vc = 2 * rr * rr * pi * hh
            vs = 4 * rr ** 3 * pi / 3
            mc, ms = vc * m / (vc + vs), vs * m / (vc + vs)
            ixx = mc * (hh * hh / 3 + rr * rr / 4) + ms * (hh * hh + 3 * hh * rr / 4 + 2 * rr * rr / 5)
            iyy = mc * rr * rr / 2 + ms * 2 * rr * rr / 5
            R = quat_to_mat(c["lquat"])
            I += R @ np.diag([ixx, iyy, ixx]) @ R.T + pa
        elif t == BOX:
            mm = m / 12
            x, y, z = p[0] ** 2, p[1] ** 2, p[2] ** 2
            R = quat_to_mat(c["lquat"])
            I += R @ np.diag([mm * (y + z), mm * (x + z), mm * (x + y)]) @ R.T + pa
        elif t == CONVEX_MESH:
            mesh = convex_meshes[c["mesh"]]
            d = np.asarray(c["lpos"], float) - com
            center = (np.asarray(p[:3]) * mesh.verts.astype(float)).sum(0)
            nv_padded = (len(mesh.verts) + 3) // 4 * 4
            center = center + (nv_padded - len(mesh.verts)) * np.asarray(p[:3]) * mesh.verts[-1].astype(float)
            center /= len(mesh.verts)
            for f in range(len(mesh.face_offsets) - 1):
                idx = mesh.face_indices[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]
                v0 = np.asarray(p[:3]) * mesh.verts[idx[0]]
                for i in range(1, len(idx) - 1):
                    v1 = np.asarray(p[:3]) * mesh.verts[idx[i]]
                    v2 = np.asarray(p[:3]) * mesh.verts[idx[i + 1]]
                    vol = abs(np.dot(np.cross(v0 - center, v1 - center), v2 - center)) / 6
                    mt = mass * vol / total
                    I += _inertia_tet(mt, [center + d, v0 + d, v1 + d, v2 + d])
    return com.astype(f32), np.linalg.inv(I).astype(f32)


def _convex_tets(mesh, p):
    s = np.asarray(p[:3], float)
    nv_padded = (len(mesh.verts) + 3) // 4 * 4
    center = (s * mesh.verts.astype(float)).sum(0) + (nv_padded - len(mesh.verts)) * s * mesh.verts[-1].astype(float)
    center /= len(mesh.verts)   # reference iterates the padded buffer but divides by size() (MassUtil.cpp:99-103)
    for f in range(len(mesh.face_offsets) - 1):
        idx = mesh.face_indices[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]
        v0 = s * mesh.verts[idx[0]]
        for i in range(1, len(idx) - 1):
            v1 = s * mesh.verts[idx[i]]
            v2 = s * mesh.verts[idx[i + 1]]
            vol = abs(np.dot(np.cross(v0 - center, v1 - center), v2 - center)) / 6
            yield vol, (center + v0 + v1 + v2) / 4


def _inertia_tet(mass, v):
    v = [np.asarray(x, float) for x in v]
    d = np.zeros(3)
    for i in range(4):
        for j in range(i + 1):
            d += v[j] * v[i]
    a, b, c = (d[1] + d[2]) / 10, (d[0] + d[2]) / 10, (d[0] + d[1]) / 10
    ap = bp = cp = 0.0
    for i in range(4):
        for j in range(4):
            mult = 2 if i == j else 1
            ap += mult * v[i][1] * v[j][2]
            bp += mult * v[i][0] * v[j][2]
            cp += mult * v[i][0] * v[j][1]
    ap /= 20; bp /= 20; cp /= 20
    I = np.zeros((3, 3))
    I[0, 0], I[1, 1], I[2, 2] = a, b, c
    I[0, 1] = I[1, 0] = -bp
    I[2, 0] = I[0, 2] = -cp
    I[1, 2] = I[2, 1] = -ap
    return mass * I


# ---- deterministic RNG (splitmix64, top 24 bits -> uniform float) --------------------------------
class SplitMix:
    def __init__(self, seed):
        self.state = np.uint64(seed)

    def u64(self, n):
        with np.errstate(over="ignore"):
            idx = np.arange(1, n + 1, dtype=np.uint64)
            z = self.state + idx * np.uint64(0x9E3779B97F4A7C15)
            self.state = self.state + np.uint64(n) * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def uniform(self, n, lo=0.0, hi=1.0):
        u = (self.u64(n) >> np.uint64(40)).astype(np.float64) / float(1 << 24)
        return (lo + (hi - lo) * u).astype(f32)

    def unit_quat(self, n):
        q = np.stack([self.uniform(n, -1, 1) for _ in range(4)], 1).astype(np.float64)
        # a few rejection rounds toward uniform-in-ball, then normalise
        nrm = np.linalg.norm(q, axis=1)
        bad = (nrm > 1) | (nrm < 1e-3)
        for _ in range(8):
            k = int(bad.sum())
            if not k:
                break
            q[bad] = np.stack([self.uniform(k, -1, 1) for _ in range(4)], 1)
            nrm = np.linalg.norm(q, axis=1)
            bad = (nrm > 1) | (nrm < 1e-3)
        q /= np.linalg.norm(q, axis=1)[:, None]
        return q.astype(f32)


IDQ = np.array([0, 0, 0, 1], f32)


# ---- config 1: box pyramid -----------------------------------------------------------------------------------------
def pyramid(n_boxes=1000, substeps=8, iterations=2) -> SceneDesc:
    """C1: rows r=0.., row r has (R-r) boxes, half extent 0.5, on a static ground box."""
    rows = 1
    while rows * (rows + 1) // 2 < n_boxes:
        rows += 1
    pos = [(0.0, -1.0, 0.0)]
    for r in range(rows):
        cnt = rows - r
        for c in range(cnt):
            if len(pos) - 1 >= n_boxes:
                break
            pos.append(((c - cnt / 2.0) * 1.05, 0.5 + r, 0.0))
    n = len(pos)
    flags = np.full(n, F_DYNAMIC, np.int32); flags[0] = 0
    prm = np.zeros((n, 4), f32); prm[:, :3] = 0.5; prm[0, :3] = (100, 1, 100)
    return bulk_scene("pyramid_%d" % n_boxes, np.array(pos, f32), np.tile(IDQ, (n, 1)), flags, np.full(n, BOX), prm, 1.0,
                      material=(0.4, 0.2, 0.0), substeps=substeps, iterations=iterations)


# ---- config 2: mixed primitives falling into a bin -------------------------------------------------------------------
def mixed_bin(n_bodies=100_000, seed=0xC2, substeps=4, iterations=2, spacing=1.0) -> SceneDesc:
    rng = SplitMix(seed)
    nx, nz = 50, 50
    ny = (n_bodies + nx * nz - 1) // (nx * nz)
    i = np.arange(n_bodies)
    gx, gz, gy = i % nx, (i // nx) % nz, i // (nx * nz)
    jit = np.stack([rng.uniform(n_bodies, -0.1, 0.1) for _ in range(3)], 1)
    p = np.stack([(gx - nx / 2 + 0.5) * spacing, 1.0 + gy * spacing, (gz - nz / 2 + 0.5) * spacing], 1).astype(f32) + jit
    t = (i % 3).astype(np.int32)
    prm = np.zeros((n_bodies, 4), f32)
    a, b, c = rng.uniform(n_bodies, 0.2, 0.4), rng.uniform(n_bodies, 0.15, 0.3), rng.uniform(n_bodies, 0.2, 0.4)
    prm[t == SPHERE, 0] = a[t == SPHERE]
    prm[t == CAPSULE, 0] = a[t == CAPSULE]; prm[t == CAPSULE, 1] = b[t == CAPSULE]
    prm[t == BOX, 0] = a[t == BOX]; prm[t == BOX, 1] = c[t == BOX]; prm[t == BOX, 2] = rng.uniform(n_bodies, 0.2, 0.4)[t == BOX]
    q = rng.unit_quat(n_bodies)
    half = nx * spacing / 2 + 2.0
    height = ny * spacing + 4.0
    spos = np.array([(0, -1, 0), (half + 1, height / 2, 0), (-half - 1, height / 2, 0), (0, height / 2, half + 1), (0, height / 2, -half - 1)], f32)
    sprm = np.array([(half + 2, 1, half + 2, 0), (1, height / 2 + 1, half + 2, 0), (1, height / 2 + 1, half + 2, 0),
                     (half + 2, height / 2 + 1, 1, 0), (half + 2, height / 2 + 1, 1, 0)], f32)
    pos = np.concatenate([spos, p]); quat = np.concatenate([np.tile(IDQ, (5, 1)), q])
    flags = np.concatenate([np.zeros(5, np.int32), np.full(n_bodies, F_DYNAMIC, np.int32)])
    types = np.concatenate([np.full(5, BOX, np.int32), t]); params = np.concatenate([sprm, prm])
    return bulk_scene("mixed_bin_%d" % n_bodies, pos, quat, flags, types, params, 1.0, material=(0.4, 0.0, 0.0),
                      substeps=substeps, iterations=iterations)


# ---- config 4: spheres + capsules over a triangle-mesh terrain ----------------------------------------------------------
def terrain_mesh(cells=1024, spacing=1.0, seed=0xC4) -> TriMeshDesc:
    rng = SplitMix(seed ^ 0x7E44A1)
    nv = cells + 1
    ix, iz = np.meshgrid(np.arange(nv), np.arange(nv), indexing="ij")
    x = (ix - cells / 2) * spacing
    z = (iz - cells / 2) * spacing
    u = rng.uniform(nv * nv, -1, 1).reshape(nv, nv)
    y = 2.0 * np.sin(0.05 * x) * np.cos(0.05 * z) + 0.25 * u
    verts = np.stack([x, y, z], -1).reshape(-1, 3).astype(f32)
    cx, cz = np.meshgrid(np.arange(cells), np.arange(cells), indexing="ij")
    v00 = (cx * nv + cz).ravel(); v10 = ((cx + 1) * nv + cz).ravel(); v01 = (cx * nv + cz + 1).ravel(); v11 = ((cx + 1) * nv + cz + 1).ravel()
    # CCW seen from +Y so normals point up
    tris = np.stack([np.stack([v00, v01, v11], 1), np.stack([v00, v11, v10], 1)], 1).reshape(-1, 3)
    return TriMeshDesc(verts=np.ascontiguousarray(verts), indices=np.ascontiguousarray(tris.reshape(-1).astype(np.uint32)))


def terrain_height(x, z):
    return 2.0 * np.sin(0.05 * x) * np.cos(0.05 * z)


def terrain(n_bodies=1_000_000, cells=1024, seed=0xC4, substeps=4, iterations=2, spacing=1.0, drop=1.0, mesh_spacing=1.0) -> SceneDesc:
    """C4: spheres (even i) and capsules (odd i) on a sqrt(n) x sqrt(n) grid over a static triangle-mesh terrain.
    mesh_spacing < body size gives every body tens of candidate triangles (tests only)."""
    rng = SplitMix(seed)
    mesh = terrain_mesh(cells, mesh_spacing, seed)
    side = int(math.ceil(math.sqrt(n_bodies)))
    i = np.arange(n_bodies)
    gx, gz = i % side, i // side
    x = (gx - side / 2 + 0.5) * spacing + rng.uniform(n_bodies, -0.2, 0.2)
    z = (gz - side / 2 + 0.5) * spacing + rng.uniform(n_bodies, -0.2, 0.2)
    y = terrain_height(x, z) + 0.25 + drop
    p = np.stack([x, y, z], 1).astype(f32)
    t = np.where(i % 2 == 0, SPHERE, CAPSULE).astype(np.int32)
    prm = np.zeros((n_bodies, 4), f32)
    rs, hh, rc = rng.uniform(n_bodies, 0.2, 0.35), rng.uniform(n_bodies, 0.15, 0.3), rng.uniform(n_bodies, 0.15, 0.25)
    prm[t == SPHERE, 0] = rs[t == SPHERE]
    prm[t == CAPSULE, 0] = hh[t == CAPSULE]; prm[t == CAPSULE, 1] = rc[t == CAPSULE]
    q = rng.unit_quat(n_bodies)
    pos = np.concatenate([np.zeros((1, 3), f32), p]); quat = np.concatenate([IDQ[None], q])
    flags = np.concatenate([np.zeros(1, np.int32), np.full(n_bodies, F_DYNAMIC, np.int32)])
    types = np.concatenate([np.array([TRIANGLE_MESH], np.int32), t]); params = np.concatenate([np.zeros((1, 4), f32), prm])
    col_mesh = np.full(n_bodies + 1, -1, np.int32); col_mesh[0] = 0
    return bulk_scene("terrain_%d" % n_bodies, pos, quat, flags, types, params, 1.0, material=(0.4, 0.0, 0.0), col_mesh=col_mesh,
                      trimesh=[mesh], substeps=substeps, iterations=iterations)


# ---- config 5: batched ragdoll scenes -------------------------------------------------------------------------------------
def qmul(a, b):
    """Hamilton product of xyzw quaternions (broadcasting)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], -1)


def qconj(q):
    q = np.asarray(q, np.float64).copy()
    q[..., :3] *= -1
    return q


def qrot(q, v):
    q = np.asarray(q, np.float64); v = np.asarray(v, np.float64)
    qv = q[..., :3]
    uv = np.cross(qv, v)
    uuv = np.cross(qv, uv)
    return v + 2.0 * (q[..., 3:4] * uv + uuv)


def axis_angle(axis, ang):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    return np.concatenate([axis * math.sin(ang / 2), [math.cos(ang / 2)]])


def ragdoll_template():
    """One ragdoll (11 bodies, 10 joints) in T-pose + its ground box; entity 0 = ground."""
    b = SceneBuilder("ragdoll")
    mat = (0.6, 0.0, 0.0)
    b.add_body((0, -0.5, 0), colliders=[dict(type=BOX, params=(3.95, 0.5, 3.95), material=mat)], dynamic=False)
    qz = axis_angle((0, 0, 1), math.pi / 2)   # capsule Y axis -> world -X .. arms along X
    parts = [
        ("pelvis", BOX, (0.15, 0.10, 0.10), (0, 1.00, 0), IDQ, 8.0),
        ("torso", BOX, (0.18, 0.22, 0.11), (0, 1.35, 0), IDQ, 20.0),
        ("head", SPHERE, (0.11,), (0, 1.72, 0), IDQ, 5.0),
        ("uarm_l", CAPSULE, (0.13, 0.045), (0.36, 1.5, 0), qz, 2.0),
        ("larm_l", CAPSULE, (0.12, 0.04), (0.70, 1.5, 0), qz, 1.5),
        ("uarm_r", CAPSULE, (0.13, 0.045), (-0.36, 1.5, 0), qz, 2.0),
        ("larm_r", CAPSULE, (0.12, 0.04), (-0.70, 1.5, 0), qz, 1.5),
        ("uleg_l", CAPSULE, (0.18, 0.06), (0.09, 0.68, 0), IDQ, 7.0),
        ("lleg_l", CAPSULE, (0.18, 0.05), (0.09, 0.24, 0), IDQ, 4.0),
        ("uleg_r", CAPSULE, (0.18, 0.06), (-0.09, 0.68, 0), IDQ, 7.0),
        ("lleg_r", CAPSULE, (0.18, 0.05), (-0.09, 0.24, 0), IDQ, 4.0),
    ]
    ids = {}
    for name, t, prm, pos, q, mass in parts:
        ids[name] = b.add_body(pos, quat=tuple(q), colliders=[dict(type=t, params=prm, material=mat)], mass=mass)
    qhingeZ = axis_angle((0, 1, 0), -math.pi / 2)   # local X -> world Z

    def joint(jt, a, c, world_point, world_frame=IDQ, params=(), world_frame1=None):
        ea, ec = ids[a], ids[c]
        pa, qa = np.asarray(b.ent[ea]["pos"], float), np.asarray(b.ent[ea]["quat"], float)
        pc, qc = np.asarray(b.ent[ec]["pos"], float), np.asarray(b.ent[ec]["quat"], float)
        wp = np.asarray(world_point, float)
        wf1 = world_frame if world_frame1 is None else world_frame1
        b.add_joint(jt, ea, qrot(qconj(qa), wp - pa), qmul(qconj(qa), world_frame), ec, qrot(qconj(qc), wp - pc), qmul(qconj(qc), wf1), params)

    # universal joint: the two anchor frames' z axes must be perpendicular at rest (UniversalJoint.cpp:8-20 drives u0[2].u1[2] to 0)
    joint(J_UNIVERSAL, "pelvis", "torso", (0, 1.115, 0), IDQ, (), axis_angle((1, 0, 0), math.pi / 2))
    joint(J_SPHERICAL, "torso", "head", (0, 1.59, 0))
    joint(J_SPHERICAL, "torso", "uarm_l", (0.20, 1.5, 0))
    joint(J_REVOLUTE, "uarm_l", "larm_l", (0.53, 1.5, 0), qhingeZ)
    joint(J_SPHERICAL, "torso", "uarm_r", (-0.20, 1.5, 0))
    joint(J_REVOLUTE, "uarm_r", "larm_r", (-0.53, 1.5, 0), qhingeZ)
    joint(J_SPHERICAL, "pelvis", "uleg_l", (0.09, 0.89, 0))
    joint(J_REVOLUTE, "uleg_l", "lleg_l", (0.09, 0.46, 0))
    joint(J_SPHERICAL, "pelvis", "uleg_r", (-0.09, 0.89, 0))
    joint(J_REVOLUTE, "uleg_r", "lleg_r", (-0.09, 0.46, 0))
    for i in range(1, 12):
        for j in range(i + 1, 12):
            b.no_collide.append((i, j))
    return b.build()


def ragdolls(n_scenes=4096, seed=0xC5, substeps=4, iterations=2, spacing=8.0, drop=0.6, first_scene=0, total_scenes=None) -> SceneDesc:
    """C5: n independent ragdoll scenes (ground box + 11 bodies + 10 joints each) laid out on a grid in one registry.
    first_scene / total_scenes select scenes [first_scene, first_scene + n_scenes) of a larger batch (multi-GPU sharding):
    layout and random root orientations depend on the GLOBAL scene index, so a shard equals that slice of the whole batch."""
    t = ragdoll_template()
    m = t.n  # 12 entities per scene
    total = n_scenes if total_scenes is None else total_scenes
    rng = SplitMix(seed)
    side = int(math.ceil(math.sqrt(total)))
    s = np.arange(first_scene, first_scene + n_scenes)
    off = np.stack([(s % side - side / 2.0) * spacing, np.zeros(n_scenes), (s // side - side / 2.0) * spacing], 1)
    rootq = rng.unit_quat(total).astype(np.float64)[first_scene:first_scene + n_scenes]
    pos = np.tile(t.pos.astype(np.float64), (n_scenes, 1)).reshape(n_scenes, m, 3)
    quat = np.tile(t.quat.astype(np.float64), (n_scenes, 1)).reshape(n_scenes, m, 4)
    centre = np.array([0, 1.0, 0])
    body = np.arange(1, m)
    rel = pos[:, body] - centre
    pos[:, body] = centre + qrot(rootq[:, None, :], rel) + np.array([0, drop + 0.9, 0])
    quat[:, body] = qmul(rootq[:, None, :], quat[:, body])
    pos += off[:, None, :]
    rep = lambda a: np.tile(a, (n_scenes,) + (1,) * (a.ndim - 1))
    d = SceneDesc(
        pos=np.ascontiguousarray(pos.reshape(-1, 3), f32), quat=np.ascontiguousarray(quat.reshape(-1, 4), f32), flags=rep(t.flags),
        vel=rep(t.vel), angvel=rep(t.angvel), inv_mass=rep(t.inv_mass), com=rep(t.com), inv_inertia=rep(t.inv_inertia),
        col_offsets=np.arange(n_scenes * m + 1, dtype=np.int32), col_lpos=rep(t.col_lpos), col_lquat=rep(t.col_lquat), col_type=rep(t.col_type),
        col_params=rep(t.col_params), col_mesh=rep(t.col_mesh), col_material=rep(t.col_material), col_flags=rep(t.col_flags), col_data=rep(t.col_data),
        name="ragdolls_%d" % n_scenes, substeps=substeps, iterations=iterations)
    joints = []
    nocol = []
    for k in range(n_scenes):
        base = k * m
        for (jt, e0, a0p, a0q, e1, a1p, a1q, prm) in t.joints:
            joints.append((jt, e0 + base, a0p, a0q, e1 + base, a1p, a1q, prm))
        nocol.extend((a + base, c + base) for a, c in t.no_collide)
    d.joints = joints
    d.no_collide = nocol
    return d


def joint_zoo(seed=5) -> SceneDesc:
    """Small scene exercising every joint type (chains hanging from static anchors, one resting on the ground)."""
    b = SceneBuilder("joint_zoo")
    b.add_body((0, -0.5, 0), colliders=[dict(type=BOX, params=(20.0, 0.5, 20.0))], dynamic=False)
    x = -6.0
    specs = [(J_SPHERICAL, ()), (J_REVOLUTE, ()), (J_UNIVERSAL, ()), (J_FIXED, ()), (J_REVOLUTE, (1.0, 2.0, 5.0)),
             (J_PRISMATIC, (0.5, -0.5, 0.0, 0.0, 5.0, 1.0)), (J_PRISMATIC, (0.2, -0.2, 1.0, 0.1, 5.0, 1.0)), (J_SERVO, (0.5, 30.0, 1.0))]
    for jt, prm in specs:
        anchor = b.add_body((x, 3.0, 0), colliders=[dict(type=BOX, params=(0.1, 0.1, 0.1))], dynamic=False)
        prev = anchor
        prev_pos = np.array([x, 3.0, 0.0])
        for link in range(3):
            p = prev_pos + np.array([0.35, -0.05, 0.02 * link])
            cur = b.add_body(tuple(p), colliders=[dict(type=CAPSULE if link % 2 == 0 else BOX, params=(0.12, 0.05) if link % 2 == 0 else (0.1, 0.06, 0.08))], mass=1.0 + link)
            mid = (prev_pos + p) / 2
            q1 = axis_angle((1, 0, 0), math.pi / 2) if jt == J_UNIVERSAL else IDQ   # universal: z axes perpendicular at rest
            b.add_joint(jt, prev, mid - prev_pos, IDQ, cur, mid - p, q1, prm)
            prev, prev_pos = cur, p
        x += 1.6
    return b.build(substeps=4, iterations=2)


# ---- config 3: convex meshes ---------------------------------------------------------------------------------------------------
def terrain_mixed(n_bodies=2000, cells=48, seed=0xC6, substeps=4, iterations=2, spacing=1.0, drop=0.3) -> SceneDesc:
    """All four dynamic shape kinds (sphere, capsule, box, convex mesh) over a triangle-mesh terrain: exercises every
    X-vs-triangle routine of CollisionTriangleMesh.cpp."""
    rng = SplitMix(seed)
    mesh = terrain_mesh(cells, 1.0, seed)
    meshes = convex_templates()
    side = int(math.ceil(math.sqrt(n_bodies)))
    i = np.arange(n_bodies)
    gx, gz = i % side, i // side
    x = (gx - side / 2 + 0.5) * spacing + rng.uniform(n_bodies, -0.2, 0.2)
    z = (gz - side / 2 + 0.5) * spacing + rng.uniform(n_bodies, -0.2, 0.2)
    y = terrain_height(x, z) + 0.25 + drop
    p = np.stack([x, y, z], 1).astype(f32)
    t = (i % 4).astype(np.int32)
    prm = np.zeros((n_bodies, 4), f32)
    a, b, c = rng.uniform(n_bodies, 0.2, 0.35), rng.uniform(n_bodies, 0.15, 0.3), rng.uniform(n_bodies, 0.2, 0.35)
    prm[:, 0] = a
    prm[t == CAPSULE, 1] = b[t == CAPSULE]
    prm[t >= BOX, 1] = b[t >= BOX] + 0.05
    prm[t >= BOX, 2] = c[t >= BOX]
    cmesh = np.where(t == CONVEX_MESH, (i // 4) % len(meshes), -1).astype(np.int32)
    q = rng.unit_quat(n_bodies)
    pos = np.concatenate([np.zeros((1, 3), f32), p]); quat = np.concatenate([IDQ[None], q])
    flags = np.concatenate([np.zeros(1, np.int32), np.full(n_bodies, F_DYNAMIC, np.int32)])
    types = np.concatenate([np.array([TRIANGLE_MESH], np.int32), t]); params = np.concatenate([np.zeros((1, 4), f32), prm])
    col_mesh = np.concatenate([np.array([0], np.int32), cmesh])
    return bulk_scene("terrain_mixed_%d" % n_bodies, pos, quat, flags, types, params, 1.0, material=(0.4, 0.1, 0.0), col_mesh=col_mesh,
                      trimesh=[mesh], convex=meshes, substeps=substeps, iterations=iterations)


def convex_from_points(points) -> ConvexMeshDesc:
    """ConvexMesh (reference include/Physecs/ConvexMesh.h:7-34) from a point set: faces as CCW index loops (seen from outside),
    unit outward normals and face centroids."""
    from scipy.spatial import ConvexHull
    pts = np.asarray(points, np.float64)
    hull = ConvexHull(pts)
    used = np.unique(hull.simplices)
    remap = -np.ones(len(pts), int); remap[used] = np.arange(len(used))
    verts = pts[used]
    centre = verts.mean(0)
    groups = {}
    for simplex, eq in zip(hull.simplices, hull.equations):
        key = tuple(np.round(eq, 6))
        groups.setdefault(key, set()).update(remap[simplex].tolist())
    offsets, indices, normals, cents = [0], [], [], []
    for key, vs in sorted(groups.items()):
        n = np.array(key[:3]); n /= np.linalg.norm(n)
        vs = sorted(vs)
        c = verts[vs].mean(0)
        # order CCW around the outward normal
        ref = verts[vs[0]] - c
        ref -= n * ref.dot(n)
        ref /= np.linalg.norm(ref)
        tang = np.cross(n, ref)
        ang = [math.atan2((verts[v] - c).dot(tang), (verts[v] - c).dot(ref)) for v in vs]
        order = [v for _, v in sorted(zip(ang, vs))]
        indices.extend(order); offsets.append(len(indices)); normals.append(n); cents.append(c)
    return ConvexMeshDesc(verts=verts.astype(f32), face_offsets=np.array(offsets, np.int32), face_indices=np.array(indices, np.int32),
                          face_normals=np.array(normals, f32), face_centroids=np.array(cents, f32))


def convex_templates():
    """Six hull templates: tetrahedron, cube, octahedron, hexagonal prism, icosahedron, dodecahedron."""
    phi = (1 + math.sqrt(5)) / 2
    tet = [(1, 1, 1), (1, -1, -1), (-1, 1, -1), (-1, -1, 1)]
    cube = [(x, y, z) for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)]
    octa = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    hexp = [(math.cos(k * math.pi / 3), y, math.sin(k * math.pi / 3)) for k in range(6) for y in (-0.8, 0.8)]
    ico = [(0, s1, s2 * phi) for s1 in (-1, 1) for s2 in (-1, 1)] + [(s1, s2 * phi, 0) for s1 in (-1, 1) for s2 in (-1, 1)] + \
          [(s2 * phi, 0, s1) for s1 in (-1, 1) for s2 in (-1, 1)]
    dod = [(x, y, z) for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)] + [(0, s1 / phi, s2 * phi) for s1 in (-1, 1) for s2 in (-1, 1)] + \
          [(s1 / phi, s2 * phi, 0) for s1 in (-1, 1) for s2 in (-1, 1)] + [(s2 * phi, 0, s1 / phi) for s1 in (-1, 1) for s2 in (-1, 1)]
    out = []
    for pts in (tet, cube, octa, hexp, ico, dod):
        p = np.array(pts, np.float64)
        p /= np.abs(p).max()
        out.append(convex_from_points(p))
    return out


def convex_pile(n_bodies=250_000, seed=0xC3, substeps=4, iterations=2, spacing=1.2, mix_prims=False) -> SceneDesc:
    """C3: convex-mesh bodies (6 templates, per-body scale in [0.3,0.5]^3) on a lattice over a static box floor.
    mix_prims=True replaces every 4th body by a sphere / capsule / box so all X-convex pair routines are exercised."""
    rng = SplitMix(seed)
    meshes = convex_templates()
    nx = nz = max(2, int(math.ceil((n_bodies / 2.0) ** (1 / 3.0))))
    i = np.arange(n_bodies)
    gx, gz, gy = i % nx, (i // nx) % nz, i // (nx * nz)
    jit = np.stack([rng.uniform(n_bodies, -0.1, 0.1) for _ in range(3)], 1)
    p = np.stack([(gx - nx / 2 + 0.5) * spacing, 0.8 + gy * spacing, (gz - nz / 2 + 0.5) * spacing], 1).astype(f32) + jit
    prm = np.zeros((n_bodies, 4), f32)
    prm[:, 0], prm[:, 1], prm[:, 2] = rng.uniform(n_bodies, 0.3, 0.5), rng.uniform(n_bodies, 0.3, 0.5), rng.uniform(n_bodies, 0.3, 0.5)
    t = np.full(n_bodies, CONVEX_MESH, np.int32)
    mesh = (i % len(meshes)).astype(np.int32)
    if mix_prims:
        sel = i % 4 == 3
        t[sel] = (i[sel] // 4) % 3
        mesh[sel] = -1
        prm[t == CAPSULE, 1] = prm[t == CAPSULE, 1] * 0.6
    q = rng.unit_quat(n_bodies)
    half = nx * spacing / 2 + 3.0
    pos = np.concatenate([np.array([[0, -1, 0]], f32), p]); quat = np.concatenate([IDQ[None], q])
    flags = np.concatenate([np.zeros(1, np.int32), np.full(n_bodies, F_DYNAMIC, np.int32)])
    types = np.concatenate([np.array([BOX], np.int32), t]); params = np.concatenate([np.array([[half, 1, half, 0]], f32), prm])
    col_mesh = np.concatenate([np.array([-1], np.int32), mesh])
    d = bulk_scene("convex_pile_%d" % n_bodies, pos, quat, flags, types, params, 1.0, material=(0.4, 0.0, 0.0), col_mesh=col_mesh,
                   convex=meshes, substeps=substeps, iterations=iterations)
    # convex inertia: use the solid-box approximation of the scaled bounding box (inputs to both sides; MassUtil is setup-time)
    cm = types == CONVEX_MESH
    hx, hy, hz = params[cm, 0], params[cm, 1], params[cm, 2]
    mm = f32(1.0 / 12.0)
    d.inv_inertia[cm, 0] = f32(1) / (mm * (hy * hy + hz * hz) * f32(4)); d.inv_inertia[cm, 4] = f32(1) / (mm * (hx * hx + hz * hz) * f32(4))
    d.inv_inertia[cm, 8] = f32(1) / (mm * (hx * hx + hy * hy) * f32(4))
    return d


# ---- joint overflow bucket, gears, triggers -------------------------------------------------------------------------------------
def joint_star(arms=12, seed=7) -> SceneDesc:
    """A hub with `arms` two-link arms jointed to it: the hub's 8 colour bits run out, so arms 9.. land in the reference's
    sequential overflow bucket (Physecs.cpp:701-704, quirk Q21) and are solved with the scalar Constraint1D semantics.
    Arm types cycle through spherical / revolute / prismatic(limit) / servo / fixed so every flag list appears in the bucket."""
    b = SceneBuilder("joint_star")
    b.add_body((0, -0.5, 0), colliders=[dict(type=BOX, params=(20.0, 0.5, 20.0))], dynamic=False)
    hub_pos = np.array([0.0, 2.0, 0.0])
    post = b.add_body((0, 3.0, 0), colliders=[dict(type=BOX, params=(0.1, 0.1, 0.1))], dynamic=False)
    hub = b.add_body(tuple(hub_pos), colliders=[dict(type=SPHERE, params=(0.3,))], mass=5.0)
    b.add_joint(J_SPHERICAL, post, (0, -0.5, 0), IDQ, hub, (0, 0.5, 0), IDQ)
    kinds = [(J_SPHERICAL, ()), (J_REVOLUTE, ()), (J_PRISMATIC, (0.1, -0.1, 0.0, 0.0, 5.0, 1.0)), (J_SERVO, (0.3, 30.0, 1.0)), (J_FIXED, ()),
             (J_REVOLUTE, (1.0, 1.5, 4.0)), (J_PRISMATIC, (0.3, -0.3, 1.0, 0.05, 5.0, 1.0)), (J_UNIVERSAL, ())]
    for a in range(arms):
        ang = 2 * math.pi * a / arms
        dirv = np.array([math.cos(ang), -0.1 * (a % 3), math.sin(ang)])
        p1 = hub_pos + 0.7 * dirv
        p2 = hub_pos + 1.3 * dirv
        e1 = b.add_body(tuple(p1), colliders=[dict(type=CAPSULE, params=(0.12, 0.05))], mass=1.0)
        e2 = b.add_body(tuple(p2), colliders=[dict(type=BOX, params=(0.08, 0.06, 0.1))], mass=0.5 + 0.1 * a)
        jt, prm = kinds[a % len(kinds)]
        q1 = axis_angle((1, 0, 0), math.pi / 2) if jt == J_UNIVERSAL else IDQ
        mid = hub_pos + 0.35 * dirv
        b.add_joint(jt, hub, mid - hub_pos, IDQ, e1, mid - p1, q1, prm)
        mid2 = (p1 + p2) / 2
        b.add_joint(J_SPHERICAL if a % 2 else J_REVOLUTE, e1, mid2 - p1, IDQ, e2, mid2 - p2, IDQ)
    return b.build(substeps=4, iterations=2)


def gear_train(pairs=3) -> SceneDesc:
    """Pairs of wheels on revolute axles (axis = anchor-frame x) coupled by a GearJoint (GearJoint.cpp:10-50); one wheel of each
    pair is driven by its revolute motor, the other has to follow with the gear ratio."""
    b = SceneBuilder("gear_train")
    b.add_body((0, -0.5, 0), colliders=[dict(type=BOX, params=(20.0, 0.5, 20.0))], dynamic=False)
    for k in range(pairs):
        z = 2.0 * k
        frame = b.add_body((0, 2.0, z), colliders=[dict(type=BOX, params=(0.05, 0.05, 0.05))], dynamic=False)
        r0, r1 = 0.4, 0.2 + 0.1 * k
        w0 = b.add_body((0.0, 2.0, z), colliders=[dict(type=BOX, params=(0.05, r0, r0))], mass=2.0, angvel=(0.5, 0, 0))
        w1 = b.add_body((0.0, 2.0 + r0 + r1, z), colliders=[dict(type=BOX, params=(0.05, r1, r1))], mass=1.0)
        b.add_joint(J_REVOLUTE, frame, (0, 0, 0), IDQ, w0, (0, 0, 0), IDQ, (1.0, 1.0 + k, 10.0))
        b.add_joint(J_REVOLUTE, frame, (0, r0 + r1, 0), IDQ, w1, (0, 0, 0), IDQ)
        b.add_joint(J_GEAR, w0, (0, 0, 0), IDQ, w1, (0, 0, 0), IDQ, (r0 / r1,))
        b.no_collide.append((w0, w1))
    return b.build(substeps=4, iterations=2)


def trigger_zoo(n=160, seed=0x7216) -> SceneDesc:
    """Trigger volumes of every geometry kind (static and carried by dynamic bodies) swept by falling bodies of every kind:
    exercises physecs::overlap (Overlap.cpp) for all shape pairs, incl. the GJK boolean path for convex meshes."""
    rng = SplitMix(seed)
    b = SceneBuilder("trigger_zoo")
    meshes = convex_templates()
    for m in meshes:
        b.add_convex(m)
    b.add_body((0, -0.5, 0), colliders=[dict(type=BOX, params=(30.0, 0.5, 30.0))], dynamic=False)
    u = rng.uniform(16 * n, 0.0, 1.0)
    qs = rng.unit_quat(2 * n)
    k = 0

    def shape(t, s):
        if t == SPHERE: return dict(type=SPHERE, params=(0.25 * s,))
        if t == CAPSULE: return dict(type=CAPSULE, params=(0.25 * s, 0.15 * s))
        if t == BOX: return dict(type=BOX, params=(0.2 * s, 0.25 * s, 0.15 * s))
        return dict(type=CONVEX_MESH, params=(0.35 * s, 0.3 * s, 0.4 * s), mesh=int(s * 97) % len(meshes))

    side = int(math.ceil(math.sqrt(n)))
    for i in range(n):
        x, z = (i % side - side / 2) * 1.1, (i // side - side / 2) * 1.1
        # a static trigger volume near the ground, type cycles through all four kinds
        zone = shape(i % 4, 1.6 + 0.8 * u[k]); k += 1
        zone["flags"] = COL_ENABLE_SIM | COL_TRIGGER
        zone["data"] = i % 3
        b.add_body((x + 0.2 * (u[k] - 0.5), 0.6, z + 0.2 * (u[k + 1] - 0.5)), quat=tuple(qs[2 * i]), colliders=[zone], dynamic=False); k += 2
        # a dynamic body falling through it; every 5th carries its own trigger collider next to the solid one
        solid = shape((i // 4) % 4, 0.8 + 0.6 * u[k]); k += 1
        solid["data"] = i % 2
        cols = [solid]
        if i % 5 == 0:
            halo = shape((i // 5) % 4, 2.0)
            halo["flags"] = COL_ENABLE_SIM | COL_TRIGGER
            halo["lpos"] = (0.1, 0.0, 0.0)
            cols.append(halo)
        b.add_body((x, 1.6 + 1.2 * u[k], z), quat=tuple(qs[2 * i + 1]), colliders=cols, mass=1.0, vel=(0.5 * (u[k + 1] - 0.5), 0, 0.5 * (u[k + 2] - 0.5))); k += 3
    return b.build(substeps=4, iterations=2)


# ---- adversarial per-pair workloads: the spill paths of the narrowphase (narrowphase.cu k_np_gjk_spill / k_np_mesh_spill) --------
def prism(sides=32, half_height=0.5, radius=1.0) -> ConvexMeshDesc:
    """A `sides`-gon prism ("cylinder"): two faces with `sides` vertices -- more than the 24-point polygons the per-thread clipping holds."""
    pts = [(radius * math.cos(2 * math.pi * k / sides), y, radius * math.sin(2 * math.pi * k / sides)) for k in range(sides) for y in (-half_height, half_height)]
    return convex_from_points(np.array(pts, np.float64))


def big_on_fine_mesh(cells=80, mesh_spacing=0.1, seed=0xAD, substeps=4, iterations=2) -> SceneDesc:
    """Large shapes resting on / sunk into a finely tessellated mesh: every (shape, mesh) pair meets hundreds to thousands of
    triangles (the per-thread path holds 128 candidates / 24 contacts), including a 3 m capsule on a 0.1 m mesh, a flat box,
    a big sphere, big convex hulls (one with 32-gon faces), next to ordinary small bodies that stay on the fast path."""
    rng = SplitMix(seed)
    nv = cells + 1
    ix, iz = np.meshgrid(np.arange(nv), np.arange(nv), indexing="ij")
    x = (ix - cells / 2) * mesh_spacing; z = (iz - cells / 2) * mesh_spacing
    u = rng.uniform(nv * nv, -1, 1).reshape(nv, nv)
    y = 0.05 * np.sin(1.3 * x) * np.cos(1.1 * z) + 0.01 * u
    verts = np.stack([x, y, z], -1).reshape(-1, 3).astype(f32)
    cx, cz = np.meshgrid(np.arange(cells), np.arange(cells), indexing="ij")
    v00 = (cx * nv + cz).ravel(); v10 = ((cx + 1) * nv + cz).ravel(); v01 = (cx * nv + cz + 1).ravel(); v11 = ((cx + 1) * nv + cz + 1).ravel()
    tris = np.stack([np.stack([v00, v01, v11], 1), np.stack([v00, v11, v10], 1)], 1).reshape(-1, 3)
    mesh = TriMeshDesc(verts=np.ascontiguousarray(verts), indices=np.ascontiguousarray(tris.reshape(-1).astype(np.uint32)))
    b = SceneBuilder("big_on_fine_mesh")
    tm = b.add_trimesh(mesh)
    templates = convex_templates()
    hexp = b.add_convex(templates[3]); dod = b.add_convex(templates[5]); cyl = b.add_convex(prism(32, 0.3, 1.0))
    b.add_body((0, 0, 0), colliders=[dict(type=TRIANGLE_MESH, params=(0, 0, 0), mesh=tm)], dynamic=False)
    lying = tuple(axis_angle((0, 0, 1), math.pi / 2))
    tilted = tuple(axis_angle((1, 0, 0.3), 0.4))
    b.add_body((-2.0, 0.25, -2.5), quat=lying, colliders=[dict(type=CAPSULE, params=(1.5, 0.3))], mass=8.0)            # 3 m capsule lying on the mesh
    b.add_body((1.5, 0.9, -2.0), colliders=[dict(type=SPHERE, params=(1.0,))], mass=10.0)                                   # sunk 0.1 m: a wide contact cap
    b.add_body((-1.0, 0.15, 0.0), colliders=[dict(type=BOX, params=(1.2, 0.2, 0.9))], mass=6.0)                             # flat box: face contacts on ~400 triangles
    b.add_body((2.0, 0.5, 1.0), quat=tilted, colliders=[dict(type=BOX, params=(0.8, 0.5, 0.6))], mass=6.0)
    b.add_body((-2.2, 0.45, 2.4), colliders=[dict(type=CONVEX_MESH, params=(1.0, 0.6, 1.0), mesh=hexp)], mass=5.0)         # hexagonal prism, face down
    b.add_body((0.6, 0.75, 2.6), quat=tilted, colliders=[dict(type=CONVEX_MESH, params=(0.9, 0.9, 0.9), mesh=dod)], mass=5.0)
    b.add_body((2.4, 0.28, -0.6), colliders=[dict(type=CONVEX_MESH, params=(0.8, 1.0, 0.8), mesh=cyl)], mass=5.0)          # 32-gon cylinder, cap down
    for k in range(24):      # ordinary bodies: the fast path next to the spilled pairs
        t = k % 4
        px, pz = -3.2 + 0.28 * k, 3.4 - 0.05 * k
        q = tuple(rng.unit_quat(1)[0])
        if t == SPHERE: col = dict(type=SPHERE, params=(0.07,))
        elif t == CAPSULE: col = dict(type=CAPSULE, params=(0.06, 0.05))
        elif t == BOX: col = dict(type=BOX, params=(0.06, 0.05, 0.07))
        else: col = dict(type=CONVEX_MESH, params=(0.07, 0.07, 0.07), mesh=dod)
        b.add_body((px, 0.16, pz), quat=q, colliders=[col], mass=0.2)
    return b.build(substeps=substeps, iterations=iterations)


def degenerate_convex(seed=0xDE6, substeps=4, iterations=2) -> SceneDesc:
    """Convex pairs that stress GJK / EPA and the face clipping: coincident and nested hulls, deep overlaps, needle / plate scales,
    face-to-face stacks of 32-gon prisms (clip polygons of up to 64 points), prisms against every primitive."""
    rng = SplitMix(seed)
    b = SceneBuilder("degenerate_convex")
    templates = convex_templates()
    ids = [b.add_convex(m) for m in templates]
    cyl = b.add_convex(prism(32, 0.4, 0.5)); cyl48 = b.add_convex(prism(48, 0.2, 0.6))
    b.add_body((0, -1, 0), colliders=[dict(type=BOX, params=(30, 1, 30))], dynamic=False)
    x = -12.0
    def place(cols_and_offsets):
        nonlocal x
        for col, off, q in cols_and_offsets:
            b.add_body((x + off[0], off[1], off[2]), quat=q, colliders=[col], mass=1.0)
        x += 2.0
    I = (0, 0, 0, 1)
    for k, cid in enumerate(ids):     # coincident centres, same hull twice (EPA starts from a degenerate simplex)
        place([(dict(type=CONVEX_MESH, params=(0.4, 0.4, 0.4), mesh=cid), (0, 0.5, 0), I), (dict(type=CONVEX_MESH, params=(0.4, 0.4, 0.4), mesh=cid), (0, 0.5, 0), I)])
    for k, cid in enumerate(ids):     # nested: a small hull deep inside a big one, slightly off centre
        q = tuple(rng.unit_quat(1)[0])
        place([(dict(type=CONVEX_MESH, params=(0.8, 0.8, 0.8), mesh=cid), (0, 0.9, 6), I), (dict(type=CONVEX_MESH, params=(0.15, 0.15, 0.15), mesh=ids[(k + 1) % 6]), (0.01 * k, 0.9, 6.02), q)])
    x = -12.0
    for k in range(6):                # needles and plates
        q = tuple(rng.unit_quat(1)[0])
        place([(dict(type=CONVEX_MESH, params=(0.02, 0.9, 0.02), mesh=ids[k]), (0, 1.0, -6), q), (dict(type=CONVEX_MESH, params=(0.9, 0.01, 0.9), mesh=ids[(k + 2) % 6]), (0.05, 1.0, -6), I)])
    x = -12.0
    tilt = tuple(axis_angle((0, 1, 0), 0.1))
    # 32 / 48-gon prisms: stacked cap to cap (convex-convex clip of two big faces), on the ground box, under a box, beside a capsule / sphere
    place([(dict(type=CONVEX_MESH, params=(1, 1, 1), mesh=cyl), (0, 0.39, -12), I), (dict(type=CONVEX_MESH, params=(1, 1, 1), mesh=cyl), (0.1, 1.17, -12), tilt)])
    place([(dict(type=CONVEX_MESH, params=(1, 1, 1), mesh=cyl48), (0, 0.19, -12), I), (dict(type=CONVEX_MESH, params=(1.2, 1, 0.8), mesh=cyl), (0.05, 0.78, -12), tilt)])
    place([(dict(type=CONVEX_MESH, params=(1, 1, 1), mesh=cyl48), (0, 0.19, -12), I), (dict(type=BOX, params=(0.3, 0.2, 0.3)), (0.02, 0.58, -12), tilt)])
    place([(dict(type=CONVEX_MESH, params=(1, 1, 1), mesh=cyl), (0, 0.39, -12), I), (dict(type=CAPSULE, params=(0.3, 0.1)), (0, 0.88, -12), tuple(axis_angle((0, 0, 1), math.pi / 2)))])
    place([(dict(type=CONVEX_MESH, params=(1, 1, 1), mesh=cyl), (0, 0.39, -12), I), (dict(type=SPHERE, params=(0.2,)), (0.1, 0.97, -12), I)])
    return b.build(substeps=substeps, iterations=iterations)
