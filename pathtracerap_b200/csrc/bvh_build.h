// bvh_build.h - host builder of the two-level BVH (PTAP_ACCEL_BVH).
// New design: the reference has only the fixed 25^3 grid of Scene::addMeshesToGrid (Scene.cpp:318-396).
#pragma once
#include <functional>
#include <vector>

#include "../../include/ptap.h"
#include "device_types.h"

namespace ptap {

struct BvhBuildResult {
    std::vector<BvhNode> nodes;      // all BLASes back to back (8-wide compressed nodes)
    std::vector<int> tri_id;         // leaf-order position -> global triangle id
    std::vector<int> mesh_root;      // per mesh: index of its BLAS root node, -1 if the mesh has no triangles
    int max_depth = 0;               // levels of 8-wide nodes of the deepest BLAS
};

// IEEE binary16 with directed rounding (host): the largest half <= x / the smallest half >= x.  |x| must be < 65504.
unsigned short halfRoundDown(float x);
unsigned short halfRoundUp(float x);
float halfToFloat(unsigned short h);

struct ChildBox { float lo[3], hi[3]; };
// Fills node `nd` from up to eight (slot, box) pairs: local origin, power-of-two scale, outward-rounded half planes; unused slots get an
// inverted box.  `boxes[c]` is read only when bit c of `used` is set.
void quantiseNode(BvhNode& nd, const ChildBox boxes[8], unsigned used);
// Decoded box of slot c (what the traversal kernel tests), in double.
void decodeChild(const BvhNode& nd, int c, double lo[3], double hi[3]);

// Collapses the binary subtree rooted at node `root2` of `n2` into 8-wide nodes appended to `out` (absolute index = base + position in
// `out`): a child is replaced by its own two children, largest box first, until the node has eight; children are then assigned to
// slots by octant (device_types.h: BvhNode) and their boxes compressed.  Every binary leaf link is handed to `emit_leaf(link, order)`,
// which appends the leaf's primitives to `order` (the new leaf order) and returns how many it appended (1..4).  Returns the root's index.
int collapseBvhWide(const Bvh2Node* n2, int root2, std::vector<BvhNode>& out, int base, std::vector<int>& order,
                 const std::function<int(int, std::vector<int>&)>& emit_leaf, int& max_depth);

// tris: global triangle table (v0, e1, e2 in .xyz).  One BLAS per mesh over [t_start, t_end).
void buildSceneBvh(const TriRec* tris, int ntris, const PtapMesh* meshes, int nmeshes, BvhBuildResult& out);
// (v0, e1 = v1 - v0, e2 = v2 - v0, flat normal in .w lanes) with the reference's arithmetic (Renderer.cpp:183-184, 203)
void makeTriRecs(const PtapVertex* vertices, const PtapTriangle* triangles, int ntris, TriRec* out);
// Walks the nodes from `root` and returns the number of violations of the structure's contract: every child box must contain all of its
// own children's boxes / its triangles' fattened boxes (`prim_box(leaf position)`), links in range, leaf counts 1..4.  Used by the tests.
long long validateBvh(const BvhNode* nodes, int nnodes, int root, int nleafprims, const std::function<ChildBox(int)>& prim_box, int* depth);

}  // namespace ptap
