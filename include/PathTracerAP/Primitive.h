// Primitive.h - the scene PODs of the drop-in facade.
// Same namespaces, type names and byte layouts as the reference's Primitive.h:10-179 (sizes: Vertex 32, Triangle 12, Mesh 40,
// Model 160, Voxel 12, Grid 28, Pixel 12), expressed over the C ABI's records (include/ptap.h) instead of glm: a vec3 is float[3],
// a column-major mat4 is float[16].  Code that only moves these records around (the reference's main.cpp, Scene's public vectors)
// compiles unchanged; code that does glm arithmetic on their members keeps the reference's own header and uses
// integration/Renderer_ptap.cpp instead.
#pragma once
#include "../ptap.h"
#include "Config.h"

namespace Common { typedef int EntityIndex; }
namespace Geometry { typedef PtapVertex Vertex; typedef PtapTriangle Triangle; }
namespace SceneElements {
using namespace Common;
using namespace Geometry;
typedef PtapMaterial Material;
typedef PtapMesh Mesh;
typedef PtapModel Model;
}
namespace SpatialAcceleration {
using namespace Common;
enum EntityType { MODEL, SCENE, TRIANGLE, SPHERE };
typedef PtapVoxel Voxel;
typedef PtapGrid Grid;
}
namespace Camera {
struct Pixel { float color[3]; };          // running sum over iterations, as render_data.dev_image_data->pool holds it (Renderer.cpp:49)
}

static_assert(sizeof(Geometry::Vertex) == 32 && sizeof(Geometry::Triangle) == 12 && sizeof(SceneElements::Mesh) == 40 &&
              sizeof(SceneElements::Model) == 160 && sizeof(SpatialAcceleration::Voxel) == 12 && sizeof(SpatialAcceleration::Grid) == 28 &&
              sizeof(Camera::Pixel) == 12, "facade records must keep the reference's layouts");
