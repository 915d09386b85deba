// Renderer_ptap.cpp - what a maintainer of purvakulkarni15/PathTracerAP adds to switch the render path to libptap.
//
// Compile this file INSTEAD OF the reference's Renderer.cpp, inside the reference tree, with the reference's own unmodified
// headers (Renderer.h, Scene.h, Primitive.h, Config.h, GPUMemoryPool.h, glm).  It defines the four Renderer methods
// (Renderer.h:46-55) on top of the C ABI of include/ptap.h.  The reference's PODs have the layouts the ABI expects
// (Primitive.h: Model 160 B, Mesh 40 B, Vertex 32 B, Triangle 12 B, Grid 28 B, Voxel 12 B - checked below), so the
// seven vectors of Scene are passed through as they are, without conversion or copy on the host.
//
//   nvcc -x cu -std=c++17 -gencode arch=compute_100a,code=sm_100a -I<reference>/PathTracerAP -I<reference>/PathTracerAP/external/include -I<repo>/include \
//        -c integration/Renderer_ptap.cpp          (g++ works too; nothing in this file is device code)
//   link:  main.o Scene.o Renderer_ptap.o -L<repo>/pathtracerap_b200 -lptap
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "Renderer.h"
#include "ptap.h"

#ifndef MAX_DEPTH
#define MAX_DEPTH 5     // the literal of Renderer.cpp:550
#endif

static_assert(sizeof(Model) == sizeof(PtapModel) && sizeof(Mesh) == sizeof(PtapMesh) && sizeof(Vertex) == sizeof(PtapVertex) &&
              sizeof(Triangle) == sizeof(PtapTriangle) && sizeof(Grid) == sizeof(PtapGrid) && sizeof(Voxel) == sizeof(PtapVoxel) &&
              sizeof(Pixel) == 3 * sizeof(float), "libptap's ABI records must match Primitive.h");

namespace {

struct Binding {                      // per-Renderer state the reference keeps in managed memory
    ptap_ctx* ctx = nullptr;          // rank 0: owns the final film
    std::vector<ptap_ctx*> all;       // PTAP_RANKS contexts, one per GPU (all[0] == ctx)
    GPUMemoryPool<Pixel> image;       // host view handed out through render_data.dev_image_data
    std::vector<Pixel> film;
};
std::map<const Renderer*, Binding> g_bindings;

void check(const Binding& b, int rc, const char* what)
{
    if (rc == PTAP_OK) return;
    std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, b.ctx ? ptap_last_error(b.ctx) : "no usable CUDA device");
    std::exit(1);                     // the reference drops CUDA errors silently; failing loudly is deliberate
}

int envInt(const char* name, int dflt) { const char* v = std::getenv(name); return v && *v ? std::atoi(v) : dflt; }

}  // namespace

// PTAP_RANKS=N renders on N GPUs of this process (devices PTAP_DEVICE .. PTAP_DEVICE + N - 1): the scene is replicated, every GPU takes
// a contiguous range of the ITER iterations, and the films are summed onto rank 0 in rank order over peer copies (ptap_reduce_peer).
void Renderer::allocateOnGPU(Scene& scene)
{
    Binding& b = g_bindings[this];
    PtapSceneView v{};
    v.models = reinterpret_cast<const PtapModel*>(scene.models.data()); v.nmodels = (int32_t)scene.models.size();
    v.meshes = reinterpret_cast<const PtapMesh*>(scene.meshes.data()); v.nmeshes = (int32_t)scene.meshes.size();
    v.vertices = reinterpret_cast<const PtapVertex*>(scene.vertices.data()); v.nvertices = (int32_t)scene.vertices.size();
    v.triangles = reinterpret_cast<const PtapTriangle*>(scene.triangles.data()); v.ntriangles = (int32_t)scene.triangles.size();
    v.grids = reinterpret_cast<const PtapGrid*>(scene.grids.data()); v.ngrids = (int32_t)scene.grids.size();
    v.voxels = reinterpret_cast<const PtapVoxel*>(scene.voxels.data()); v.nvoxels = (int32_t)scene.voxels.size();
    v.refs = scene.per_voxel_data_pool.data(); v.nrefs = (int32_t)scene.per_voxel_data_pool.size();
    v.grid_dim[0] = GRID_X; v.grid_dim[1] = GRID_Y; v.grid_dim[2] = GRID_Z;
    // default: the reference's own hits, bit for bit (tier R0) - through the BVH when the grids have the shape Scene.cpp builds
    // (PTAP_ACCEL_GRID_EMULATED), else by walking them (PTAP_ACCEL_GRID_COMPAT; also PTAP_ACCEL=grid).  bvh / lbvh: exact closest hit.
    const std::string accel = std::getenv("PTAP_ACCEL") ? std::getenv("PTAP_ACCEL") : "";
    const int kind = accel == "bvh" ? PTAP_ACCEL_BVH : accel == "lbvh" ? PTAP_ACCEL_BVH_DEVICE : accel == "grid" ? PTAP_ACCEL_GRID_COMPAT : PTAP_ACCEL_GRID_EMULATED;
    const int ranks = std::max(1, envInt("PTAP_RANKS", 1)), dev0 = envInt("PTAP_DEVICE", 0);
    std::vector<int> devices;                                 // PTAP_RANK_DEVICES=0,0,1: explicit device of every rank (default dev0 + rank)
    if (const char* list = std::getenv("PTAP_RANK_DEVICES")) for (const char* q = list; *q;) { devices.push_back(std::atoi(q)); while (*q && *q != ',') ++q; if (*q) ++q; }
    b.all.assign(ranks, nullptr);
    for (int r = 0; r < ranks; ++r) {
        b.ctx = nullptr;
        check(b, ptap_create(r < (int)devices.size() ? devices[r] : dev0 + r, 0, &b.ctx), "ptap_create");
        b.all[r] = b.ctx;
        check(b, ptap_upload_scene(b.ctx, &v), "ptap_upload_scene");
        int rc = ptap_build_accel(b.ctx, kind);
        const int rc_first = rc;
        if (rc == PTAP_E_UNSUPPORTED && kind == PTAP_ACCEL_GRID_EMULATED) rc = ptap_build_accel(b.ctx, PTAP_ACCEL_GRID_COMPAT);   // same results, walked
        check(b, rc, "ptap_build_accel");
        if (r == 0) std::fprintf(stderr, "ptap: acceleration structure: %s\n", kind == PTAP_ACCEL_BVH ? "BVH (host build)" : kind == PTAP_ACCEL_BVH_DEVICE ? "BVH (device build)" :
                                 kind == PTAP_ACCEL_GRID_COMPAT ? "the reference's grids, walked" : rc_first == 0 ? "the reference's grids, emulated through the BVH" : "the reference's grids, walked (lists not box-shaped)");
        // nrays = RESOLUTION * SAMPLES camera rays on one lattice (Renderer.cpp:96, 527-542); SAMPLESX = SAMPLESY = 1 in Config.h:14-15
        check(b, ptap_set_render_params(b.ctx, RESOLUTION_X * SAMPLESX, RESOLUTION_Y * SAMPLESY, MAX_DEPTH, PTAP_FLAG_FIRST_HIT_CACHE | PTAP_FLAG_ITER_TIMES), "ptap_set_render_params");
    }
    b.ctx = b.all[0];
    render_data = RenderData{};
}

void Renderer::renderLoop()
{
    Binding& b = g_bindings[this];
    const int iters = envInt("PTAP_ITER", ITER), ranks = (int)b.all.size();
    const auto t0 = std::chrono::high_resolution_clock::now();
    std::vector<int> first(ranks + 1, 0);
    for (int r = 0; r < ranks; ++r) first[r + 1] = first[r] + iters / ranks + (r < iters % ranks ? 1 : 0);
    for (int r = 0; r < ranks; ++r) {                         // asynchronous: every GPU starts before any is waited for
        check(b, ptap_frame_begin(b.all[r]), "ptap_frame_begin");
        if (first[r + 1] > first[r]) check(b, ptap_render(b.all[r], first[r], first[r + 1]), "ptap_render");
    }
    for (int r = 1; r < ranks; ++r) check(b, ptap_reduce_peer(b.all[0], b.all[r]), "ptap_reduce_peer");
    b.film.resize((size_t)RESOLUTION_X * SAMPLESX * RESOLUTION_Y * SAMPLESY);
    check(b, ptap_read_film(b.ctx, reinterpret_cast<float*>(b.film.data())), "ptap_read_film");
    b.image.size = (int)b.film.size();
    b.image.pool = b.film.data();
    render_data.dev_image_data = &b.image;                    // Renderer.cpp:49 reads the image through this pointer
    const auto t1 = std::chrono::high_resolution_clock::now();
    // Renderer.cpp:641-643 prints one line per iteration.  Iterations are enqueued without host round trips and overlap on the device, so
    // the figure is the device time between the completions of consecutive iterations (rank 0's share when several GPUs render).
    std::vector<float> ms((size_t)std::max(first[1], 1));
    int32_t n = 0;
    check(b, ptap_get_iteration_times(b.ctx, ms.data(), (int32_t)ms.size(), &n), "ptap_get_iteration_times");
    for (int k = 0; k < n && k < (int)ms.size(); ++k)
        std::cout << "Iteration " << k + 1 << ": " << (long long)((ms[k] - (k ? ms[k - 1] : 0.0f)) * 1000.0f) << " microseconds" << std::endl;
    std::cout << "Full run: " << std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count() << " microseconds" << std::endl;   // Renderer.cpp:645-647
}

void Renderer::renderImage()
{
    Binding& b = g_bindings[this];
    // SAMPLES > 1: each pixel is the mean of its lattice samples (the reference's own gather leaves the image black, Renderer.cpp:493)
    check(b, ptap_write_bmp_resolved(b.ctx, "Render.bmp", envInt("PTAP_ITER", ITER), SAMPLESX, SAMPLESY), "ptap_write_bmp");
}

void Renderer::free()
{
    auto it = g_bindings.find(this);
    if (it == g_bindings.end()) return;
    for (ptap_ctx* c : it->second.all) if (c) ptap_destroy(c);
    g_bindings.erase(it);
    render_data = RenderData{};
}
