"""Colour-size histogram of a settled scene (how many manifolds each solver colour holds)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from physecs_b200 import scenes as S
from physecs_b200.capi import Context
name = sys.argv[1]; n = int(sys.argv[2]); settle = int(sys.argv[3])
mk = {"ragdolls": lambda: S.ragdolls(n), "mixed": lambda: S.mixed_bin(n), "pyramid": lambda: S.pyramid(n), "convex": lambda: S.convex_pile(n),
      "terrain": lambda: S.terrain(n, cells=int(max(16, (n ** 0.5) * 1.024)), drop=0.3)}[name]
d = mk()
ctx = Context(d, max_pairs=64 * d.n, max_manifolds=16 * d.n)
for _ in range(settle):
    ctx.step()
m = ctx.manifolds()
col = m["color"]
print(d.name, "manifolds", len(col), "points", int(m["num_points"].sum()), "hist", np.bincount(col).tolist())
k = m["keys"]
dyn = set(d.dynamic_entities().tolist())
both = np.array([(a in dyn) and (b in dyn) for a, b in zip(k[:, 0], k[:, 2])])
print("body-body manifolds:", int(both.sum()), "body-static:", int((~both).sum()))
