// Host-side, registration-time build of the static triangle-mesh BVH (setup, not on the per-step path).
//
// The reference builds this tree in the TriangleMesh constructor (src/TriangleMesh.cpp:144-164) with a
// 6-bucket binned-SAH split (chooseSplit :66-97, subdivide :99-142) that REORDERS the triangle array; the
// post-build triangle index is what manifolds carry as `triangleIndex` and what keys the contact cache
// (src/Physecs.cpp:237).  To hand out the same indices the device mesh must be built the same way, so this
// file restates that construction (fp32, same comparisons) and returns flat arrays for upload.
#include "trimesh_build.h"
#include <algorithm>
#include <cfloat>
#include <cmath>

namespace {

struct B3 { float mn[3], mx[3]; };

inline B3 emptyB() { B3 b; for (int k = 0; k < 3; ++k) { b.mn[k] = FLT_MAX; b.mx[k] = -FLT_MAX; } return b; }
inline B3 uni(const B3& a, const B3& b) {
    B3 r;
    for (int k = 0; k < 3; ++k) { r.mn[k] = std::min(a.mn[k], b.mn[k]); r.mx[k] = std::max(a.mx[k], b.mx[k]); }
    return r;
}
inline float area(const B3& b) {
    float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    return 2 * (dx * dy + dy * dz + dz * dx);
}

struct Tri { unsigned idx[3]; B3 bounds; float normal[3]; float centroid[3]; int orig; };
struct Node { B3 bounds; int triCount; int index; };

const int kBuckets = 6;

struct Builder {
    std::vector<Tri> tris;
    std::vector<Node> nodes;

    bool fillBuckets(int nodeId, int axis, B3* bb, int* bc) {
        const Node& node = nodes[nodeId];
        for (int i = 0; i < kBuckets; ++i) { bb[i] = emptyB(); bc[i] = 0; }
        float start = node.bounds.mn[axis], end = node.bounds.mx[axis];
        float len = end - start;
        if (!len) return false;
        for (int i = 0; i < node.triCount; ++i) {
            const Tri& t = tris[node.index + i];
            int b = std::min(kBuckets - 1, static_cast<int>((t.centroid[axis] - start) / len * kBuckets));
            bb[b] = uni(bb[b], t.bounds);
            bc[b]++;
        }
        return true;
    }

    static bool evalSplit(int split, const B3* bb, const int* bc, float& cost, B3& bl, B3& br, int& cl, int& cr) {
        bl = emptyB(); br = emptyB(); cl = 0; cr = 0;
        for (int i = 0; i < kBuckets; ++i) {
            if (i < split) { bl = uni(bl, bb[i]); cl += bc[i]; }
            else { br = uni(br, bb[i]); cr += bc[i]; }
        }
        if (!cl || !cr) return false;
        cost = cl * area(bl) + cr * area(br);
        return true;
    }

    void subdivide(int rootId) {
        // explicit stack with the reference's depth-first order (left subtree fully first) so node indices match
        std::vector<int> stack;
        stack.push_back(rootId);
        while (!stack.empty()) {
            int nodeId = stack.back();
            stack.pop_back();
            int bestAxis = 0, splitIndex = 0, bestCl = 0, bestCr = 0;
            B3 bestL = emptyB(), bestR = emptyB();
            float bestCost = FLT_MAX;
            for (int axis = 0; axis < 3; ++axis) {
                B3 bb[kBuckets]; int bc[kBuckets];
                if (!fillBuckets(nodeId, axis, bb, bc)) continue;
                for (int i = 1; i < kBuckets; ++i) {
                    float cost; B3 bl, br; int cl = 0, cr = 0;
                    if (evalSplit(i, bb, bc, cost, bl, br, cl, cr) && cost < bestCost) {
                        bestCost = cost; bestAxis = axis; splitIndex = i; bestL = bl; bestR = br; bestCl = cl; bestCr = cr;
                    }
                }
            }
            if (!(bestCl && bestCr)) continue;
            Node node = nodes[nodeId];
            float start = node.bounds.mn[bestAxis], end = node.bounds.mx[bestAxis];
            float len = end - start;
            int i = node.index, j = i + node.triCount - 1;
            while (i <= j) {
                int b = static_cast<int>((tris[i].centroid[bestAxis] - start) / len * kBuckets);
                if (b < splitIndex) ++i;
                else std::swap(tris[i], tris[j--]);
            }
            nodes.push_back({ bestL, bestCl, node.index });
            nodes.push_back({ bestR, bestCr, i });
            int first = static_cast<int>(nodes.size()) - 2;
            nodes[nodeId].triCount = 0;
            nodes[nodeId].index = first;
            stack.push_back(first + 1);   // right later
            stack.push_back(first);       // left first
        }
    }
};

} // namespace

void pb_build_trimesh_host(const float* verts, int nVerts, const unsigned* indices, int nIndices, PbHostTriMesh& out) {
    Builder b;
    int nTris = nIndices / 3;
    b.tris.reserve(nTris);
    auto V = [&](unsigned i, int k) { return verts[3 * i + k]; };
    for (int t = 0; t < nTris; ++t) {
        Tri tri;
        tri.orig = t;
        for (int k = 0; k < 3; ++k) tri.idx[k] = indices[3 * t + k];
        unsigned i0 = tri.idx[0], i1 = tri.idx[1], i2 = tri.idx[2];
        for (int k = 0; k < 3; ++k) {
            tri.bounds.mn[k] = std::min(std::min(V(i0, k), V(i1, k)), V(i2, k));
            tri.bounds.mx[k] = std::max(std::max(V(i0, k), V(i1, k)), V(i2, k));
        }
        float e0[3], e1[3];
        for (int k = 0; k < 3; ++k) { e0[k] = V(i1, k) - V(i0, k); e1[k] = V(i2, k) - V(i0, k); }
        float n[3] = { e0[1] * e1[2] - e1[1] * e0[2], e0[2] * e1[0] - e1[2] * e0[0], e0[0] * e1[1] - e1[0] * e0[1] };
        float inv = 1.0f / std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int k = 0; k < 3; ++k) {
            tri.normal[k] = n[k] * inv;
            tri.centroid[k] = (V(i0, k) + V(i1, k) + V(i2, k)) / 3.f;
        }
        b.tris.push_back(tri);
    }
    b.nodes.reserve(2 * (size_t)nTris);
    Node root; root.bounds = emptyB(); root.triCount = nTris; root.index = 0;
    for (int t = 0; t < nTris; ++t) root.bounds = uni(root.bounds, b.tris[t].bounds);
    b.nodes.push_back(root);
    b.subdivide(0);

    out.nTris = nTris;
    out.triIdx.resize(3 * (size_t)nTris); out.triNormal.resize(3 * (size_t)nTris); out.triCentroid.resize(3 * (size_t)nTris);
    out.triOrig.resize(nTris);
    for (int t = 0; t < nTris; ++t) {
        for (int k = 0; k < 3; ++k) {
            out.triIdx[3 * t + k] = b.tris[t].idx[k];
            out.triNormal[3 * t + k] = b.tris[t].normal[k];
            out.triCentroid[3 * t + k] = b.tris[t].centroid[k];
        }
        out.triOrig[t] = b.tris[t].orig;
    }
    out.nNodes = (int)b.nodes.size();
    out.nodeBounds.resize(6 * (size_t)out.nNodes); out.nodeCountIndex.resize(2 * (size_t)out.nNodes);
    for (int i = 0; i < out.nNodes; ++i) {
        for (int k = 0; k < 3; ++k) { out.nodeBounds[6 * i + k] = b.nodes[i].bounds.mn[k]; out.nodeBounds[6 * i + 3 + k] = b.nodes[i].bounds.mx[k]; }
        out.nodeCountIndex[2 * i] = b.nodes[i].triCount;
        out.nodeCountIndex[2 * i + 1] = b.nodes[i].index;
    }
}
