// bvh_build.h - host builder of the two-level BVH (PTAP_ACCEL_BVH).
// New design: the reference has only the fixed 25^3 grid of Scene::addMeshesToGrid (Scene.cpp:318-396).
#pragma once
#include <vector>

#include "../../include/ptap.h"
#include "device_types.h"

namespace ptap {

struct BvhBuildResult {
    std::vector<BvhNode> nodes;      // all BLASes back to back
    std::vector<TriRec> tris;        // triangles in leaf order (all meshes)
    std::vector<int> tri_id;         // leaf-order position -> global triangle id
    std::vector<int> mesh_root;      // per mesh: index of its BLAS root node, -1 if the mesh has no triangles
    std::vector<InstanceCull> cull;  // per model: conservative world bounds
    int max_depth = 0;
};

// tris: global triangle table as uploaded (v0, e1, e2 in .xyz).  One BLAS per mesh over [t_start, t_end).
void buildSceneBvh(const std::vector<TriRec>& tris, const std::vector<PtapMesh>& meshes, const std::vector<PtapModel>& models,
                   BvhBuildResult& out);

}  // namespace ptap
