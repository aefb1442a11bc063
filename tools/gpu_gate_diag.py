"""Teacher-forced gates through the landing phase of a bin scene; on the first mismatching step, which bodies differ and where their manifolds sit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import parity
from physecs_b200 import scenes as S
from oracle.ref import RefScene
from physecs_b200.capi import Context
n = int(sys.argv[1]); steps = int(sys.argv[2]); mult = int(sys.argv[3]) if len(sys.argv) > 3 else 16
d = S.mixed_bin(n)
ref = RefScene(d, 0, hashfix=True); ref.presort()
ctx = Context(d, max_pairs=64 * d.n + 4096, max_manifolds=mult * d.n + 4096)
for k in range(steps):
    parity.sync_device_to_oracle(ctx, ref)
    if k > 0:
        ctx.refresh_bounds()
    ctx.step()
    gm = ctx.manifolds(); c = ctx.counts()
    ref.set_manifold_order(gm["keys"]); ref.simulate()
    P, Q, V, W = ctx.get_state_entities(); p, q, v, w = ref.get_state()
    err = np.abs(V - v).max(axis=1)
    if err.max() > 1e-4:
        bad = np.where(err > 1e-4)[0]
        print(f"step {k}: {len(bad)} bodies differ; manifolds {c.n_manifolds} points {c.n_points if hasattr(c, 'n_points') else '?'} colours {c.n_colors} islands {ctx.island_stats()}")
        keys = gm["keys"]
        slots = np.where(np.isin(keys[:, 0], bad) | np.isin(keys[:, 1], bad))[0]
        print("   slots of manifolds touching differing bodies: count", len(slots), "min", slots.min(), "max", slots.max(), "first 20", slots[:20].tolist())
        print("   their colours:", np.unique(gm["color"][slots], return_counts=True))
        print("   their point counts:", np.unique(gm["num_points"][slots], return_counts=True))
        cum = np.concatenate([[0], np.cumsum(gm["num_points"])])
        print("   point offsets of those slots: min", cum[slots].min(), "max", cum[slots].max(), "total points", cum[-1])
        allbad = np.zeros(len(keys), bool); allbad[slots] = True
        good = np.where(~allbad)[0]
        print("   slots NOT touching a differing body: count", len(good), "max", good.max() if len(good) else -1)
        nbad = globals().get("nbad", 0) + 1
        if nbad >= int(os.environ.get("DIAG_MAX_BAD", "1")):
            break
    elif os.environ.get("DIAG_VERBOSE"):
        print(f"step {k}: exact; manifolds {c.n_manifolds} colours {c.n_colors}")
ctx.close(); ref.close()
