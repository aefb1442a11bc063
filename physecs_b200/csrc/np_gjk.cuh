// GJK / EPA bin (convex meshes).  Placeholder until the convex path lands: pairs that involve a convex mesh
// raise PB_EUNSUPPORTED in the step status instead of being silently skipped.
//   reference: src/GJK.h:226-292, src/EPA.h:172-186, src/Collision.cpp:352-426, :488-499, :694-887
#pragma once
#include "np_clip.cuh"
#include "pb_ctx.h"

#define PB_STATUS_UNSUPPORTED_SHAPE 0x100

__device__ inline bool collideGjkPair(int t0, float4 q0, V3 pos0, Q4 or0, int mesh0, int t1, float4 q1, V3 pos1, Q4 or1, int mesh1,
                                      const PbConvexDev* convexes, Manifold& m, bool& flip, int* counters) {
    atomicOr(&counters[CNT_STATUS], PB_STATUS_UNSUPPORTED_SHAPE);
    return false;
}
