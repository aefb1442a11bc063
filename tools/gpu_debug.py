"""Development helper: isolate where device and oracle diverge (tiny scenes, absolute diffs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from physecs_b200 import scenes as S  # noqa: E402
from physecs_b200.capi import Context  # noqa: E402
from oracle.ref import RefScene  # noqa: E402


def run(desc, steps, label):
    ref = RefScene(desc, 0, hashfix=True)
    ctx = Context(desc)
    worst = np.zeros(4)
    for k in range(steps):
        p, q, v, w = ref.get_state()
        ctx.set_state_entities(p, q, v, w)
        if k > 0:
            ctx.refresh_bounds()
        ctx.step()
        gm = ctx.manifolds()
        ref.set_manifold_order(gm["keys"])
        ref.simulate()
        st = ref.order_stats()
        if st[1] or st[2]:
            print(f"  [{label}] step {k}: MANIFOLD SET MISMATCH matched/missing/extra={st}")
        P, Q, V, W = ctx.get_state_entities()
        p, q, v, w = ref.get_state()
        d = [np.abs(P - p).max(), np.abs(Q - q).max(), np.abs(V - v).max(), np.abs(W - w).max()]
        worst = np.maximum(worst, d)
        if max(d) > 1e-5 or k < 3:
            i = int(np.argmax(np.abs(W - w).max(1)))
            print(f"  [{label}] step {k}: manifolds={len(gm['keys'])} dpos={d[0]:.2e} dquat={d[1]:.2e} dvel={d[2]:.2e} dangvel={d[3]:.2e} worst body {i}: W={W[i]} w={w[i]} V={V[i]} v={v[i]}")
            ks = gm["keys"]
            sel = (ks[:, 0] == i) | (ks[:, 2] == i)
            print("     manifolds of that body:", ks[sel].tolist(), "np", gm["num_points"][sel].tolist(), "colors", gm["color"][sel].tolist())
    print(f"[{label}] worst abs diffs pos/quat/vel/angvel: {worst}")
    ctx.close(); ref.close()


if __name__ == "__main__":
    rng = S.SplitMix(7)
    # 1. free spinning bodies, no contacts
    b = S.SceneBuilder("free")
    for i in range(20):
        t = i % 3
        prm = [(0.3,), (0.3, 0.2), (0.3, 0.2, 0.4)][t]
        b.add_body((i * 3.0, 5.0, 0.0), quat=rng.unit_quat(1)[0], colliders=[dict(type=t, params=prm)], angvel=tuple(rng.uniform(3, -5, 5)), vel=(1, 2, 3))
    run(b.build(substeps=4), 5, "free")
    # 2. one sphere on ground
    b = S.SceneBuilder("sphere")
    b.add_body((0, -1, 0), colliders=[dict(type=S.BOX, params=(50, 1, 50))], dynamic=False)
    b.add_body((0, 0.31, 0), colliders=[dict(type=S.SPHERE, params=(0.3,))], vel=(1, 0, 0.5), angvel=(0, 3, 1))
    run(b.build(substeps=4), 20, "sphere")
    # 3. one box on ground
    b = S.SceneBuilder("box")
    b.add_body((0, -1, 0), colliders=[dict(type=S.BOX, params=(50, 1, 50))], dynamic=False)
    b.add_body((0, 0.52, 0), quat=rng.unit_quat(1)[0], colliders=[dict(type=S.BOX, params=(0.5, 0.5, 0.5))], vel=(1, 0, 0.5))
    run(b.build(substeps=4), 40, "box")
    # 4. two spheres + ground (body-body contact)
    b = S.SceneBuilder("two")
    b.add_body((0, -1, 0), colliders=[dict(type=S.BOX, params=(50, 1, 50))], dynamic=False)
    b.add_body((0, 0.3, 0), colliders=[dict(type=S.SPHERE, params=(0.3,))])
    b.add_body((0.1, 0.85, 0), colliders=[dict(type=S.SPHERE, params=(0.3,))])
    run(b.build(substeps=4), 40, "two")
    run(S.terrain(1500, cells=48, drop=0.3), 16, "terrain")
