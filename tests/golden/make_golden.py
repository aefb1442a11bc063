"""Generate the committed golden vectors from the REFERENCE implementation (oracle/_ref, built from /root/reference).

Run in the build container:  python tests/golden/make_golden.py
Outputs (small .npz files, committed):
  narrowphase_prims.npz   manifolds of physecs::collision over all pairs of a dense random soup of spheres/capsules/boxes
  narrowphase_mesh.npz    manifolds of spheres/capsules against a small triangle-mesh terrain
  pyramid_steps.npz       state of a 60-box pyramid after 1..5 reference steps (numThreads=0), + pair counts
The reference ships no tests or golden files of its own; these pin the oracle build and give the
GPU box fixtures that do not need /root/reference.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from physecs_b200 import scenes as S  # noqa: E402
from oracle.ref import RefScene  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def soup(n=90, seed=11, extent=2.2):
    rng = S.SplitMix(seed)
    p = np.stack([rng.uniform(n, -extent, extent) for _ in range(3)], 1)
    q = rng.unit_quat(n)
    t = (np.arange(n) % 3).astype(np.int32)
    prm = np.zeros((n, 4), np.float32)
    a, b, c = rng.uniform(n, 0.25, 0.6), rng.uniform(n, 0.2, 0.45), rng.uniform(n, 0.25, 0.6)
    prm[:, 0] = a
    prm[t == S.CAPSULE, 1] = b[t == S.CAPSULE]
    prm[t == S.BOX, 1] = b[t == S.BOX] + 0.1
    prm[t == S.BOX, 2] = c[t == S.BOX]
    flags = np.full(n, S.F_DYNAMIC, np.int32)
    return S.bulk_scene("soup", p, q, flags, t, prm, 1.0, substeps=4)


def convex_soup(n=80, seed=13, extent=1.8):
    """Dense soup of convex meshes mixed with spheres / capsules / boxes (GJK/EPA routines)."""
    rng = S.SplitMix(seed)
    meshes = S.convex_templates()
    p = np.stack([rng.uniform(n, -extent, extent) for _ in range(3)], 1)
    q = rng.unit_quat(n)
    i = np.arange(n)
    t = np.where(i % 3 == 2, (i // 3) % 3, S.CONVEX_MESH).astype(np.int32)
    prm = np.zeros((n, 4), np.float32)
    prm[:, 0], prm[:, 1], prm[:, 2] = rng.uniform(n, 0.3, 0.55), rng.uniform(n, 0.25, 0.5), rng.uniform(n, 0.3, 0.55)
    mesh = np.where(t == S.CONVEX_MESH, i % len(meshes), -1).astype(np.int32)
    flags = np.full(n, S.F_DYNAMIC, np.int32)
    return S.bulk_scene("convex_soup", p, q, flags, t, prm, 1.0, col_mesh=mesh, convex=meshes, substeps=4)


def all_pairs(n):
    i, j = np.triu_indices(n, 1)
    z = np.zeros_like(i)
    return np.stack([i, z, j, z], 1).astype(np.int32)


def main():
    d = soup()
    r = RefScene(d, 0, hashfix=True)
    m = r.narrowphase(all_pairs(d.n))
    np.savez_compressed(os.path.join(HERE, "narrowphase_prims.npz"), **m)
    print("prims manifolds", len(m["keys"]), "points hist", np.bincount(m["num_points"]))
    r.close()

    d = S.terrain(400, cells=24, drop=-0.15)
    r = RefScene(d, 0, hashfix=True)
    pr = np.stack([np.zeros(400, np.int32), np.zeros(400, np.int32), np.arange(1, 401, dtype=np.int32), np.zeros(400, np.int32)], 1)
    m = r.narrowphase(pr)
    np.savez_compressed(os.path.join(HERE, "narrowphase_mesh.npz"), **m)
    print("mesh manifolds", len(m["keys"]), "points hist", np.bincount(m["num_points"]))
    r.close()

    d = convex_soup()
    r = RefScene(d, 0, hashfix=True)
    m = r.narrowphase(all_pairs(d.n))
    np.savez_compressed(os.path.join(HERE, "narrowphase_convex.npz"), **m)
    print("convex manifolds", len(m["keys"]), "points hist", np.bincount(m["num_points"]))
    r.close()

    d = S.pyramid(60)
    r = RefScene(d, 0, hashfix=False)   # as shipped
    states = []
    npairs = []
    for _ in range(5):
        r.simulate()
        states.append(np.concatenate(r.get_state(), 1))
        npairs.append(len(r.pairs()))
    np.savez_compressed(os.path.join(HERE, "pyramid_steps.npz"), states=np.stack(states), npairs=np.array(npairs))
    print("pyramid pairs", npairs)
    r.close()


if __name__ == "__main__":
    main()
