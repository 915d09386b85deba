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
// PTAP_DEPTH, PTAP_ACCEL=emu|grid|bvh|lbvh, PTAP_DEVICE, PTAP_RANKS = GPUs of this process that share the iterations).  The default reproduces the reference's own 25^3 grid walk
// bit for bit - through the BVH (PTAP_ACCEL_GRID_EMULATED) when the grids have the shape Scene.cpp builds, else by walking them (also PTAP_ACCEL=grid); PTAP_ACCEL=bvh selects the BVH (built on the host), lbvh the same built on the GPU.  All errors throw std::runtime_error: there is no CPU fallback.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

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
        const std::string a = std::getenv("PTAP_ACCEL") ? std::getenv("PTAP_ACCEL") : "";
        const bool lbvh = a == "lbvh";                            // tree built on the GPU at this call
        const bool bvh = a == "bvh";
        const bool walk = a == "grid";                            // walk the grids instead of emulating the walk through the BVH
        PtapSceneView v{};
        v.models = scene.models.data(); v.nmodels = (int32_t)scene.models.size();
        v.meshes = scene.meshes.data(); v.nmeshes = (int32_t)scene.meshes.size();
        v.vertices = scene.vertices.data(); v.nvertices = (int32_t)scene.vertices.size();
        v.triangles = scene.triangles.data(); v.ntriangles = (int32_t)scene.triangles.size();
        v.grids = scene.grids.data(); v.ngrids = (int32_t)scene.grids.size();
        v.voxels = scene.voxels.data(); v.nvoxels = (int32_t)scene.voxels.size();
        v.refs = scene.per_voxel_data_pool.data(); v.nrefs = (int32_t)scene.per_voxel_data_pool.size();
        v.grid_dim[0] = GRID_X; v.grid_dim[1] = GRID_Y; v.grid_dim[2] = GRID_Z;
        // PTAP_RANKS = N: N GPUs of this process (devices PTAP_DEVICE ..), scene replicated, iterations split, films summed onto rank 0
        const int ranks = std::max(1, pick("PTAP_RANKS", 0, 1)), dev0 = pick("PTAP_DEVICE", 0, 0);
        std::vector<int> devices;                             // PTAP_RANK_DEVICES=0,0,1: explicit device of every rank (default dev0 + rank)
        if (const char* list = std::getenv("PTAP_RANK_DEVICES")) for (const char* q = list; *q;) { devices.push_back(std::atoi(q)); while (*q && *q != ',') ++q; if (*q) ++q; }
        all.assign((size_t)ranks, nullptr);
        for (int r = 0; r < ranks; ++r) {
            ctx = nullptr;
            check(ptap_create(r < (int)devices.size() ? devices[(size_t)r] : dev0 + r, 0, &ctx), "ptap_create");
            all[(size_t)r] = ctx;
            check(ptap_upload_scene(ctx, &v), "ptap_upload_scene");
            const int kind = lbvh ? PTAP_ACCEL_BVH_DEVICE : bvh || scene.grids.empty() ? PTAP_ACCEL_BVH : walk ? PTAP_ACCEL_GRID_COMPAT : PTAP_ACCEL_GRID_EMULATED;
            int rc = ptap_build_accel(ctx, kind);
            const int rc_first = rc;
            if (rc == PTAP_E_UNSUPPORTED && kind == PTAP_ACCEL_GRID_EMULATED) rc = ptap_build_accel(ctx, PTAP_ACCEL_GRID_COMPAT);   // same results, walked
            check(rc, "ptap_build_accel");
            if (r == 0) std::fprintf(stderr, "ptap: acceleration structure: %s\n", kind == PTAP_ACCEL_BVH ? "BVH (host build)" : kind == PTAP_ACCEL_BVH_DEVICE ? "BVH (device build)" :
                                     kind == PTAP_ACCEL_GRID_COMPAT ? "the reference's grids, walked" : rc_first == 0 ? "the reference's grids, emulated through the BVH" : "the reference's grids, walked (lists not box-shaped)");
            check(ptap_set_render_params(ctx, width * samples_x, height * samples_y, depth, PTAP_FLAG_FIRST_HIT_CACHE | PTAP_FLAG_ITER_TIMES), "ptap_set_render_params");
            if (scene.config_has_camera) check(ptap_set_camera(ctx, &scene.config_camera), "ptap_set_camera");
        }
        ctx = all[0];
    }

    void renderLoop()
    {
        need();
        const auto t0 = std::chrono::high_resolution_clock::now();
        const int ranks = (int)all.size();
        std::vector<int> first((size_t)ranks + 1, 0);
        for (int r = 0; r < ranks; ++r) first[(size_t)r + 1] = first[(size_t)r] + iters / ranks + (r < iters % ranks ? 1 : 0);
        for (int r = 0; r < ranks; ++r) {                     // asynchronous: every GPU starts before any is waited for
            check(ptap_frame_begin(all[(size_t)r]), "ptap_frame_begin");
            if (first[(size_t)r + 1] > first[(size_t)r]) check(ptap_render(all[(size_t)r], first[(size_t)r], first[(size_t)r + 1]), "ptap_render");
        }
        for (int r = 1; r < ranks; ++r) check(ptap_reduce_peer(all[0], all[(size_t)r]), "ptap_reduce_peer");
        image.size = width * height;
        host_film.resize((size_t)image.size);
        image.pool = host_film.data();
        check(ptap_read_film_resolved(ctx, samples_x, samples_y, &host_film[0].color[0]), "ptap_read_film_resolved");
        render_data.dev_image_data = &image;
        const auto t1 = std::chrono::high_resolution_clock::now();
        PtapStats st{};
        ptap_get_stats(ctx, &st);
        // Renderer.cpp:641-643: one line per iteration (device time between the completions of consecutive iterations of rank 0's share)
        std::vector<float> ms((size_t)std::max(first[1], 1));
        int32_t n = 0;
        check(ptap_get_iteration_times(ctx, ms.data(), (int32_t)ms.size(), &n), "ptap_get_iteration_times");
        for (int k = 0; k < n && k < (int)ms.size(); ++k)
            std::cout << "Iteration " << k + 1 << ": " << (long long)((ms[(size_t)k] - (k ? ms[(size_t)k - 1] : 0.0f)) * 1000.0f) << " microseconds" << std::endl;
        std::cout << "Full run: " << std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count() << " microseconds ("
                  << iters << " iterations on " << ranks << " GPU(s), " << st.rays_traced << " rays traced on rank 0, " << st.ms_render << " ms on its device)" << std::endl;
    }

    void renderImage()
    {
        need();
        check(ptap_write_bmp_resolved(ctx, "Render.bmp", iters, samples_x, samples_y), "ptap_write_bmp");
    }

    void free()
    {
        for (ptap_ctx* c : all) if (c) ptap_destroy(c);
        all.clear();
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

    ptap_ctx* ctx = nullptr;             // rank 0: owns the final film
    std::vector<ptap_ctx*> all;          // PTAP_RANKS contexts, one per GPU (all[0] == ctx)
    GPUMemoryPool<Pixel> image;
    std::vector<Pixel> host_film;
};
