// Drop-in header name of the reference (include/Physecs/BVH.h).  The reference keeps an incremental SAH tree for its scene
// queries (src/BVH.cpp); here queries and broadphase share the device LBVH, and Scene::getBVH() hands out a snapshot of it in the
// reference's node format (BVH.h:31-49) for debug drawing.
#pragma once
#include "Bounds.h"
#include <entt.hpp>

namespace physecs {

struct BVH { static const int null = -1; };

struct InternalNodeData {
    int left = BVH::null;
    int right = BVH::null;
};

struct LeafNodeData {
    entt::entity entity = entt::null;
    int colliderIndex = 0;
};

struct BVHNode {
    Bounds bounds;
    int parent = BVH::null;
    bool isLeaf = false;
    union {
        InternalNodeData internal;
        LeafNodeData leaf;
    };
    BVHNode() : internal() {}
};

} // namespace physecs
