"""Development helper (run under gpurun): verbose gate runs on several scenes, one failure does not stop the rest."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from physecs_b200 import scenes as S  # noqa: E402
from tests import parity  # noqa: E402

def zoo_part(i):
    import copy
    z = S.joint_zoo()
    d = copy.copy(z)
    d.joints = z.joints[3 * i:3 * i + 3]
    d.name = "zoo%d_type%d" % (i, d.joints[0][0])
    return d


CASES = {
    "pyramid": lambda: (S.pyramid(120), 30),
    "bin": lambda: (S.mixed_bin(1200, spacing=0.8), 50),
    "terrain": lambda: (S.terrain(1500, cells=48, drop=0.3), 50),
    "zoo": lambda: (S.joint_zoo(), 60),
    "ragdolls": lambda: (S.ragdolls(8), 90),
    "convex": lambda: (S.convex_pile(300, mix_prims=True), 90),
    "tmix": lambda: (S.terrain_mixed(800, cells=40), 90),
    **{"zoo%d" % i: (lambda i=i: (zoo_part(i), 40)) for i in range(8)},
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for name in names:
        desc, steps = CASES[name]()
        t0 = time.time()
        try:
            s = parity.run_gates(desc, steps=steps, verbose=True)
            print(f"[{name}] OK {s} in {time.time() - t0:.1f}s", flush=True)
        except Exception:
            print(f"[{name}] FAILED after {time.time() - t0:.1f}s", flush=True)
            traceback.print_exc()
