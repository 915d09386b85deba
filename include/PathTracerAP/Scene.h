// Scene.h - drop-in facade of the reference's Scene (Scene.h:21-39) over libptap's host scene (ptap_scene_*).
//
//   Scene(std::string config)   the reference ignores `config` and builds its hard-coded scene from "Input data\\*.obj" in the
//                               current directory (Scene.cpp:3-224); so does this class, unless `config` names a readable
//                               Config.txt-style file (schema of the reference's Config.txt:1-31), which is then parsed.
//   seven public vectors        same names, element layouts and meaning (Scene.h:26-32); they are COPIES owned by the object,
//                               the caller may edit them before Renderer::allocateOnGPU, which reads them back.
#pragma once
#include <cstdio>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "Primitive.h"

using namespace std;                    // the reference's header does this and its main.cpp relies on it
using namespace Common;
using namespace Geometry;
using namespace SceneElements;
using namespace SpatialAcceleration;

class Scene {
public:
    explicit Scene(string config)
    {
        ptap_scene* s = nullptr;
        ifstream probe(config.c_str());
        const bool is_config = probe.good() && !(config.size() > 4 && config.compare(config.size() - 4, 4, ".obj") == 0);
        const int rc = is_config ? ptap_scene_create_from_config(config.c_str(), &s) : ptap_scene_create_builtin(".", &s);
        if (rc != PTAP_OK || !s) throw runtime_error("Scene: libptap error " + to_string(rc));
        int32_t p[4] = {0, 0, 0, 0};
        ptap_scene_config_params(s, p);
        config_width = p[0]; config_height = p[1]; config_iter = p[2]; config_depth = p[3];
        config_has_camera = ptap_scene_config_camera(s, &config_camera) != 0;
        PtapSceneView v;
        ptap_scene_view(s, &v);
        models.assign(v.models, v.models + v.nmodels);
        meshes.assign(v.meshes, v.meshes + v.nmeshes);
        vertices.assign(v.vertices, v.vertices + v.nvertices);
        triangles.assign(v.triangles, v.triangles + v.ntriangles);
        if (v.ngrids > 0) {
            grids.assign(v.grids, v.grids + v.ngrids);
            voxels.assign(v.voxels, v.voxels + v.nvoxels);
            per_voxel_data_pool.assign(v.refs, v.refs + v.nrefs);
        }
        ptap_scene_destroy(s);
    }

    vector<Model> models;
    vector<Mesh> meshes;
    vector<Vertex> vertices;
    vector<Triangle> triangles;
    vector<Grid> grids;
    vector<Voxel> voxels;
    vector<EntityIndex> per_voxel_data_pool;

    // RESOLUTION / ITER / DEPTH keys of a parsed Config.txt, 0 when absent (an extension: the reference has compile-time macros only)
    int config_width = 0, config_height = 0, config_iter = 0, config_depth = 0;
    // CAMERA_ORIGIN / CAMERA_PLANE / CAMERA_SPAN / JITTER keys (the reference hard-codes its camera, Renderer.cpp:538-545)
    bool config_has_camera = false;
    PtapCamera config_camera{{0.0f, 0.0f, 920.0f}, {-10.0f, -4.0f, 900.0f}, {20.0f, 16.0f}, 0, 0u};
};
