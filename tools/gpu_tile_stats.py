"""Tile statistics behind the broadphase choice (pb_get_broadphase_info) for a few scenes: tools/gpu_tile_stats.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physecs_b200 import scenes as S
from physecs_b200.capi import Context
for name, mk, steps in (("mixed_bin_20000", lambda: S.mixed_bin(20000), 150), ("mixed_bin_100000", lambda: S.mixed_bin(100000), 150), ("convex_pile_20000", lambda: S.convex_pile(20000), 120),
                        ("terrain_30000", lambda: S.terrain(30000, cells=180, drop=0.3), 120), ("ragdolls_2048", lambda: S.ragdolls(2048), 60)):
    d = mk(); ctx = Context(d, max_pairs=64 * d.n + 4096, max_manifolds=16 * d.n + 4096)
    seen = []
    for k in range(steps):
        ctx.step()
        if k % 30 == 29:
            ctx.sync(); i = ctx.broadphase_info(); seen.append((k, i["all_pairs"], round(i["tile_hits"] / max(i["tiles"], 1), 1)))
    print(name, "colliders", len(d.col_type), "(step, all-pairs?, group-tile pairs per tile; threshold 24):", seen, flush=True)
    ctx.close()
