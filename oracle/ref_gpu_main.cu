// oracle/_ref/pt_ref_gpu: the reference's OWN CUDA kernels (Renderer.cpp:363-648) compiled for sm_100a and run on the B200, as the
// GPU-side "before" number next to the CPU baseline.  TEST / MEASUREMENT INFRASTRUCTURE: only bench.py's
// baseline legs execute this binary; the product never does.
//
// The reference sources are compiled from a patched throw-away copy (oracle/build_ref_gpu.sh): P1 (duplicate `inline`, utility.h:44),
// P3 (`return false` on the path that falls off the end, Renderer.cpp:359), the Config.h macros turned into __managed__ ints so that one
// binary serves every resolution, a ray counter next to the two closest-hit launches, and an optional switch that skips the
// cudaDeviceSynchronize after each launch.  No kernel arithmetic is touched.
//
//   pt_ref_gpu <scene.bin> <W> <H> <iters> <warmup iters> <sync 0|1> [depth [bmp-dir]]
// scene.bin: int32 counts[4] (models, meshes, vertices, triangles) followed by the four arrays in the reference's own layouts
// (Primitive.h), as bench.py dumps them (the OBJ files do not travel to the GPU box); the grids are then built by the reference's own
// Scene::addMeshesToGrid (Scene.cpp:318-396), exactly as oracle/ref_harness.cpp does for the CPU build.
// prints one JSON line: rays traced, seconds (cudaDeviceSynchronize-bracketed wall clock around Renderer::renderLoop), Mrays/s.
#include <cuda.h>
#include <cuda_runtime_api.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>

// ptap_cfg_{res_x,res_y,iter,depth,grid_x,grid_y,grid_z}: __managed__ ints defined by the patched Config.h (build_ref_gpu.sh, patch K)
long long ptap_rays_traced = 0;
int ptap_cfg_sync = 1;

#include "Scene.cpp"
#undef CLAMP
#include "Renderer.cpp"

int main(int argc, char** argv)
{
    if (argc < 7) { fprintf(stderr, "usage: %s <scene.bin> <W> <H> <iters> <warmup> <sync 0|1> [depth [bmp-dir]]\n", argv[0]); return 2; }
    const int W = atoi(argv[2]), H = atoi(argv[3]), iters = atoi(argv[4]), warm = atoi(argv[5]);
    ptap_cfg_sync = atoi(argv[6]);
    if (W <= 0 || H <= 0 || (W * H) % 32 || iters <= 0) { fprintf(stderr, "W*H must be a positive multiple of 32 (Renderer.cpp:573)\n"); return 2; }
    FILE* sf = fopen(argv[1], "rb");
    int cnt[4];
    if (!sf || fread(cnt, 4, 4, sf) != 4) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    ptap_cfg_res_x = W; ptap_cfg_res_y = H;
    if (argc > 7) ptap_cfg_depth = atoi(argv[7]);
    int dev = 0; cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 3; }
    fflush(stdout);
    FILE* real_out = fdopen(dup(1), "w");
    if (!freopen("/dev/null", "w", stdout)) return 2;                   // the reference prints one line per iteration
    // Scene's only constructor hard-codes the bundled scene and its OBJ paths (Scene.cpp:3-224); the class is exactly seven std::vectors
    // (Scene.h:26-32), so the members are built in place from the dumped arrays and the reference's own grid build runs on them.
    Scene* sp = static_cast<Scene*>(::operator new(sizeof(Scene)));
    new (&sp->models) std::vector<Model>(cnt[0]); new (&sp->meshes) std::vector<Mesh>(cnt[1]); new (&sp->vertices) std::vector<Vertex>(cnt[2]);
    new (&sp->triangles) std::vector<Triangle>(cnt[3]); new (&sp->grids) std::vector<Grid>(); new (&sp->voxels) std::vector<Voxel>();
    new (&sp->per_voxel_data_pool) std::vector<EntityIndex>();
    bool ok = fread(sp->models.data(), sizeof(Model), cnt[0], sf) == (size_t)cnt[0] && fread(sp->meshes.data(), sizeof(Mesh), cnt[1], sf) == (size_t)cnt[1] &&
              fread(sp->vertices.data(), sizeof(Vertex), cnt[2], sf) == (size_t)cnt[2] && fread(sp->triangles.data(), sizeof(Triangle), cnt[3], sf) == (size_t)cnt[3];
    fclose(sf);
    if (!ok) { fprintf(stderr, "short read of %s\n", argv[1]); return 2; }
    sp->addMeshesToGrid();
    Scene& scene = *sp;
    Renderer renderer;
    renderer.allocateOnGPU(scene);                                      // main.cpp:17
    if (warm > 0) { ptap_cfg_iter = warm; renderer.renderLoop(); }
    cudaDeviceSynchronize();
    ptap_rays_traced = 0;
    ptap_cfg_iter = iters;
    const auto t0 = std::chrono::high_resolution_clock::now();
    renderer.renderLoop();                                              // main.cpp:20
    cudaDeviceSynchronize();
    const auto t1 = std::chrono::high_resolution_clock::now();
    const cudaError_t err = cudaGetLastError();
    const double s = std::chrono::duration<double>(t1 - t0).count();
    // channel means of the film (un-normalised sum / iters), so that the caller can check this run against its own frame
    double mean[3] = {0, 0, 0};
    for (int i = 0; i < W * H; ++i) { const glm::vec3 c = renderer.render_data.dev_image_data->pool[i].color; mean[0] += c.x; mean[1] += c.y; mean[2] += c.z; }
    if (argc > 8) { if (chdir(argv[8]) == 0) renderer.renderImage(); }
    fprintf(real_out, "{\"device\": \"%s\", \"W\": %d, \"H\": %d, \"iters\": %d, \"depth\": %d, \"sync_per_launch\": %s, \"rays\": %lld, \"seconds\": %.6f, "
            "\"ms_per_iter\": %.4f, \"Mrays_s\": %.3f, \"film_mean\": [%.6f, %.6f, %.6f], \"cuda_error\": \"%s\"}\n",
            prop.name, W, H, iters, (int)ptap_cfg_depth, ptap_cfg_sync ? "true" : "false", ptap_rays_traced, s, s / iters * 1e3, ptap_rays_traced / s / 1e6,
            mean[0] / iters / (W * H), mean[1] / iters / (W * H), mean[2] / iters / (W * H), cudaGetErrorString(err));
    fflush(real_out);
    renderer.free();
    return err == cudaSuccess ? 0 : 4;
}
