"""Batched independent scenes across GPUs (BASELINE.json configs[4]).

Scenes never interact, so a batch shards by scene: rank r of G simulates a contiguous block of scenes in its own device
context (one process per GPU).  There is no collective on the data path; torch.distributed is used only for the barrier
around the timed region and the max-over-ranks reduction of the measured time (bench.py).
"""
from __future__ import annotations


def shard_range(n_scenes: int, world: int, rank: int):
    """Contiguous block [begin, end) of scene indices owned by `rank`; blocks differ in size by at most one scene."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world / rank")
    base, extra = divmod(n_scenes, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (a device time) over all ranks; identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def scene_seed(base_seed: int, scene_index: int) -> int:
    """Per-scene RNG seed: depends on the GLOBAL scene index, so a scene is the same whichever rank simulates it."""
    return (base_seed + 0x9E3779B1 * (scene_index + 1)) & 0xFFFFFFFF


class Batch:
    """A batch of independent scenes sharded over devices through the C ABI's pb_batch_* entry points (include/physecs_b200.h):
    shard k = one context on devices[k] with its own host thread and stream.  `descs[k]` is the scene description of shard k
    (its block of scenes concatenated, e.g. scenes.ragdolls(count, first_scene=b, total_scenes=n))."""

    def __init__(self, descs, devices, pairs_per_body=64, manifolds_per_body=16):
        import ctypes as C
        from . import capi
        self.C, self.capi = C, capi
        self.lib = capi.load_library()
        self.lib.pb_batch_ctx.restype = C.c_void_p
        self.lib.pb_batch_last_error.restype = C.c_char_p
        self.lib.pb_batch_destroy.restype = None
        n = len(descs)
        assert n == len(devices) and n >= 1
        caps = (capi.Caps * n)()
        for k, d in enumerate(descs):
            nc = len(d.col_type)
            caps[k].max_bodies = max(d.n, 16); caps[k].max_colliders = max(nc, 16)
            caps[k].max_pairs = max(pairs_per_body * d.n, 4096); caps[k].max_manifolds = max(manifolds_per_body * d.n, 4096)
            caps[k].max_joints = max(len(d.joints), 16)
        self.handle = C.c_void_p()
        dev = (C.c_int * n)(*[int(x) for x in devices])
        rc = self.lib.pb_batch_create(n, dev, caps, C.byref(self.handle))
        if rc != capi.PB_OK:
            raise capi.PbError(rc, "pb_batch_create failed (is a CUDA device visible? there is no CPU fallback)")
        self.descs = list(descs)
        self.shards = []
        for k, d in enumerate(descs):
            ctx = capi.Context.from_handle(self.lib.pb_batch_ctx(self.handle, k), d)
            ctx.upload(d)
            self.shards.append(ctx)
        self.n_dyn = [c.n_dyn for c in self.shards]

    def _check(self, rc):
        if rc != self.capi.PB_OK:
            raise self.capi.PbError(rc, self.lib.pb_batch_last_error(self.handle).decode())

    def step(self, n_steps=1):
        d = self.descs[0]
        C = self.C
        self._check(self.lib.pb_batch_step(self.handle, int(n_steps), C.c_float(d.dt), int(d.substeps), int(d.iterations), C.c_float(d.gravity)))

    def sync(self):
        self._check(self.lib.pb_batch_sync(self.handle))

    def _ptrs(self, arrays):
        C = self.C
        P = C.POINTER(C.c_float)
        return (P * len(arrays))(*[a.ctypes.data_as(P) for a in arrays])

    def set_state(self, pos, quat, vel, ang):
        """lists of per-shard host arrays (pinned recommended); asynchronous"""
        C = self.C
        nd = (C.c_int * len(self.n_dyn))(*self.n_dyn)
        self._keep = (pos, quat, vel, ang)
        self._check(self.lib.pb_batch_set_state(self.handle, self._ptrs(pos), self._ptrs(quat), self._ptrs(vel), self._ptrs(ang), nd))

    def get_state(self, pos, quat, vel, ang):
        self._check(self.lib.pb_batch_get_state(self.handle, self._ptrs(pos), self._ptrs(quat), self._ptrs(vel), self._ptrs(ang)))

    def alloc_state(self):
        import numpy as np
        mk = lambda w: [np.zeros((n, w), np.float32) for n in self.n_dyn]
        return mk(3), mk(4), mk(3), mk(3)

    def close(self):
        if self.handle:
            self.lib.pb_batch_destroy(self.handle)
            self.handle = self.C.c_void_p()
            for c in self.shards:
                c.ctx = self.C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
