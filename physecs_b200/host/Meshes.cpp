// physecs::TriangleMesh host object (include/Physecs/detail/b200_types.hpp).  The BVH comes from the same registration-time
// builder the device context uses (pb_build_trimesh -> csrc/trimesh_build.cpp, which reproduces the reference's binned-SAH
// construction, src/TriangleMesh.cpp:99-164), so `triangles` is in post-build order like the reference's member.
#include <TriangleMesh.h>
#include "../../include/physecs_b200.h"
#include <stdexcept>

namespace physecs {

TriangleMesh::TriangleMesh(const std::vector<glm::vec3>& verts, const std::vector<unsigned int>& indices) : sourceIndices(indices), vertices(verts) {
    const int nTris = (int)indices.size() / 3;
    if (nTris == 0) return;
    std::vector<unsigned> triIdx((size_t)3 * nTris);
    std::vector<int> triOrig(nTris), nodeCI((size_t)4 * nTris);
    std::vector<float> nodeBounds((size_t)12 * nTris);
    int nNodes = 0;
    if (pb_build_trimesh((const float*)verts.data(), (int)verts.size(), indices.data(), (int)indices.size(), triIdx.data(), triOrig.data(),
                         nodeBounds.data(), nodeCI.data(), &nNodes) != PB_OK)
        throw std::runtime_error("physecs_b200: triangle mesh BVH build failed");
    triangles.resize(nTris);
    for (int t = 0; t < nTris; ++t) {
        Triangle& T = triangles[t];
        for (int k = 0; k < 3; ++k) T.indices[k] = triIdx[3 * (size_t)t + k];
        const glm::vec3 &a = verts[T.indices[0]], &b = verts[T.indices[1]], &c = verts[T.indices[2]];
        T.bounds = { glm::min(a, glm::min(b, c)), glm::max(a, glm::max(b, c)) };
        T.normal = glm::normalize(glm::cross(b - a, c - a));
        T.centroid = (a + b + c) / 3.f;
    }
    bvh.resize(nNodes);
    for (int n = 0; n < nNodes; ++n) {
        const float* nb = &nodeBounds[6 * (size_t)n];
        bvh[n].bounds = { glm::vec3(nb[0], nb[1], nb[2]), glm::vec3(nb[3], nb[4], nb[5]) };
        bvh[n].triCount = nodeCI[2 * (size_t)n];
        bvh[n].index = nodeCI[2 * (size_t)n + 1];
    }
}

const std::vector<int>& TriangleMesh::overlapBvh(const Bounds& q) {
    overlapScratch.clear();
    if (bvh.empty()) return overlapScratch;
    std::vector<int> stack{ rootId };
    while (!stack.empty()) {
        const TriangleMeshBVHNode& n = bvh[stack.back()];
        stack.pop_back();
        const Bounds& b = n.bounds;
        bool hit = !(q.max.x < b.min.x || q.min.x > b.max.x || q.max.y < b.min.y || q.min.y > b.max.y || q.max.z < b.min.z || q.min.z > b.max.z);
        if (!hit) continue;
        if (n.triCount) for (int i = 0; i < n.triCount; ++i) overlapScratch.push_back(n.index + i);
        else { stack.push_back(n.index + 1); stack.push_back(n.index); }   // left child is visited first
    }
    return overlapScratch;
}

} // namespace physecs
