"""Offline prototype input: world AABBs (+0.01 margin) of the 1 M-body terrain scene's dynamic colliders (spheres / capsules resting
on the terrain height), written as float32 [n][6] for tools/proto/tree_visits.cpp.  No GPU involved."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from physecs_b200 import scenes as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = S.terrain(n, drop=0.0) if n == 1_000_000 else S.terrain(n, cells=int(max(16, (n ** 0.5) * 1.05)), drop=0.0)
pos = d.pos[1:].astype(np.float64); q = d.quat[1:].astype(np.float64)
t = d.col_type[1:]; prm = d.col_params[1:].astype(np.float64)
# capsule axis = rotate(q, (0, 1, 0)), xyzw quaternion
x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
axis = np.stack([2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)], 1)
ext = np.where((t == S.SPHERE)[:, None], prm[:, :1].repeat(3, 1), np.abs(axis) * prm[:, :1] + prm[:, 1:2])
ext += 0.01
out = np.concatenate([pos - ext, pos + ext], 1).astype(np.float32)
out.tofile(sys.argv[2] if len(sys.argv) > 2 else "/tmp/aabbs.bin")
print(out.shape, out[:2])
