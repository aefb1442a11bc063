// Flat result of the registration-time triangle-mesh BVH build (see trimesh_build.cpp).
#pragma once
#include <vector>

struct PbHostTriMesh {
    int nTris = 0, nNodes = 0;
    std::vector<unsigned> triIdx;      // 3 per triangle, post-build order
    std::vector<float> triNormal;      // 3 per triangle
    std::vector<float> triCentroid;    // 3 per triangle
    std::vector<int> triOrig;          // original triangle index of each built triangle
    std::vector<float> nodeBounds;     // 6 per node (min xyz, max xyz)
    std::vector<int> nodeCountIndex;   // 2 per node (triCount, index)
};

void pb_build_trimesh_host(const float* verts, int nVerts, const unsigned* indices, int nIndices, PbHostTriMesh& out);
