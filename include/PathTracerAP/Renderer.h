// Renderer.h - drop-in facade of the reference's Renderer (Renderer.h:19-55) over the C ABI of libptap (include/ptap.h).
//
//   allocateOnGPU(Scene&)  Renderer.cpp:65-130   copies the scene's seven vectors to the device (the Scene may die afterwards)
//   renderLoop()           Renderer.cpp:567-648  ITER iterations of the wavefront; prints the reference's two timing lines
//   renderImage()          Renderer.cpp:15-63    writes Render.bmp into the current directory
//   free()                 Renderer.cpp:132-148
//   render_data            Renderer.h:19-35      dev_image_data->pool[x + y * W].color is host-readable after renderLoop()
//
// Resolution, iteration count and depth come from Config.h's macros (the reference's only configuration), overridden by the
// RESOLUTION / ITER / DEPTH keys of a parsed Config.txt, overridden by the environment (PTAP_WIDTH, PTAP_HEIGHT, PTAP_ITER,
// PTAP_DEPTH, PTAP_ACCEL=grid|bvh|lbvh, PTAP_DEVICE).  The acceleration structure defaults to the reference's own 25^3 grid walk
// (bit-compatible hits); PTAP_ACCEL=bvh selects the BVH (built on the host), lbvh the same built on the GPU.  All errors throw std::runtime_error: there is no CPU fallback.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>

#include "Config.h"
#include "GPUMemoryPool.h"
#include "Primitive.h"
#include "Scene.h"

using namespace Camera;

struct RenderData {
    GPUMemoryPool<Pixel>* dev_image_data = nullptr;
};

class Renderer {
public:
    void allocateOnGPU(Scene& scene)
    {
        width = pick("PTAP_WIDTH", scene.config_width, RESOLUTION_X);
        height = pick("PTAP_HEIGHT", scene.config_height, RESOLUTION_Y);
        iters = pick("PTAP_ITER", scene.config_iter, ITER);
        depth = pick("PTAP_DEPTH", scene.config_depth, MAX_DEPTH);
        samples_x = std::max(1, pick("PTAP_SAMPLESX", 0, SAMPLESX));   // Config.h:14-15: camera rays per pixel, on one W*SX x H*SY lattice
        samples_y = std::max(1, pick("PTAP_SAMPLESY", 0, SAMPLESY));
        const char* a = std::getenv("PTAP_ACCEL");
        const bool lbvh = a && std::string(a) == "lbvh";          // tree built on the GPU at this call
        const bool bvh = a && std::string(a) == "bvh";
        check(ptap_create(pick("PTAP_DEVICE", 0, 0), 0, &ctx), "ptap_create");
        PtapSceneView v{};
        v.models = scene.models.data(); v.nmodels = (int32_t)scene.models.size();
        v.meshes = scene.meshes.data(); v.nmeshes = (int32_t)scene.meshes.size();
        v.vertices = scene.vertices.data(); v.nvertices = (int32_t)scene.vertices.size();
        v.triangles = scene.triangles.data(); v.ntriangles = (int32_t)scene.triangles.size();
        v.grids = scene.grids.data(); v.ngrids = (int32_t)scene.grids.size();
        v.voxels = scene.voxels.data(); v.nvoxels = (int32_t)scene.voxels.size();
        v.refs = scene.per_voxel_data_pool.data(); v.nrefs = (int32_t)scene.per_voxel_data_pool.size();
        v.grid_dim[0] = GRID_X; v.grid_dim[1] = GRID_Y; v.grid_dim[2] = GRID_Z;
        check(ptap_upload_scene(ctx, &v), "ptap_upload_scene");
        check(ptap_build_accel(ctx, lbvh ? PTAP_ACCEL_BVH_DEVICE : bvh || scene.grids.empty() ? PTAP_ACCEL_BVH : PTAP_ACCEL_GRID_COMPAT), "ptap_build_accel");
        check(ptap_set_render_params(ctx, width * samples_x, height * samples_y, depth, PTAP_FLAG_FIRST_HIT_CACHE), "ptap_set_render_params");
    }

    void renderLoop()
    {
        need();
        const auto t0 = std::chrono::high_resolution_clock::now();
        check(ptap_frame_begin(ctx), "ptap_frame_begin");
        check(ptap_render(ctx, 0, iters), "ptap_render");
        image.size = width * height;
        host_film.resize((size_t)image.size);
        image.pool = host_film.data();
        check(ptap_read_film_resolved(ctx, samples_x, samples_y, &host_film[0].color[0]), "ptap_read_film_resolved");
        render_data.dev_image_data = &image;
        const auto t1 = std::chrono::high_resolution_clock::now();
        PtapStats st{};
        ptap_get_stats(ctx, &st);
        std::cout << "Full run: " << std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count() << " microseconds ("
                  << iters << " iterations, " << st.rays_traced << " rays traced, " << st.ms_render << " ms on the device)" << std::endl;
    }

    void renderImage()
    {
        need();
        check(ptap_write_bmp_resolved(ctx, "Render.bmp", iters, samples_x, samples_y), "ptap_write_bmp");
    }

    void free()
    {
        if (ctx) ptap_destroy(ctx);
        ctx = nullptr;
        render_data.dev_image_data = nullptr;
    }

    RenderData render_data;
    int width = RESOLUTION_X, height = RESOLUTION_Y, iters = ITER, depth = MAX_DEPTH;
    int samples_x = SAMPLESX, samples_y = SAMPLESY;

private:
    static int pick(const char* env, int from_config, int dflt)
    {
        const char* e = std::getenv(env);
        if (e && *e) return std::atoi(e);
        return from_config > 0 ? from_config : dflt;
    }
    void need() const { if (!ctx) throw std::runtime_error("Renderer: allocateOnGPU has not been called"); }
    void check(int rc, const char* what) const
    {
        if (rc != PTAP_OK) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + (ctx ? ptap_last_error(ctx) : "no usable CUDA device"));
    }

    ptap_ctx* ctx = nullptr;
    GPUMemoryPool<Pixel> image;
    std::vector<Pixel> host_film;
};
