"""CTA 0's own colour phases in the whole-step kernel of a ragdoll batch: time per NGS / contact / joint phase (in-kernel stamps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physecs_b200 import scenes as S
from physecs_b200.capi import Context
n = int(sys.argv[1])
d = S.ragdolls(n)
ctx = Context(d)
for _ in range(60):
    ctx.step()
ctx.sync(); ctx.set_profile(True)
K = 20
for _ in range(K):
    ctx.step()
ctx.sync()
cols = ctx.profile_colors()
t = ctx.timings()
print(f"ragdolls {n}: solve stage {t.solve:.3f} ms; CTA 0 local phases per step:")
for name, (ms, cnt) in zip(("joint NGS", "contact", "joint solve"), cols[61:64]):
    print(f"   {name:12s} {ms / K * 1e3:8.1f} us/step in {cnt / K:6.1f} phases = {1e3 * ms / max(cnt, 1):6.2f} us/phase")
print("   all:", ctx.profile())
ctx.close()
