// oracle/_ref harness: a C entry-point layer around the UNMODIFIED reference sources
// (/root/reference/PathTracerAP/{Renderer,Scene}.cpp, compiled for the host by
// oracle/build_ref.sh).  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the resulting library.
//
// The reference has no FFI of its own, its kernels are file-static, and it
// records neither primitive ids nor barycentrics, so this translation unit #includes the
// patched copies of the two .cpp files and drives the reference's own functions:
//   * wavefront steps  = the launches of Renderer::renderLoop (Renderer.cpp:567-648), one
//     call per launch so that a test can dump the state between them;
//   * ref_trace mode 0 = computeRaySceneIntersectionKernel (Renderer.cpp:363-409) on a
//     caller-supplied ray set (oracle tier R0);
//   * ref_trace mode 1 = the same per-model transform + computeRayTriangleIntersection
//     (Renderer.cpp:174-215) on EVERY triangle of the model's mesh, no bbox/grid culling
//     (oracle tier R1, structure independent).
// Probe variables (ptap_probe_*) are written by two statements that build_ref.sh appends
// next to Renderer.cpp:210 and :395; they do not touch any arithmetic.

#include <omp.h>
#include <unistd.h>
#include <fcntl.h>
#include <cstdint>
#include <cstdio>
#include <chrono>

// runtime knobs that replace the compile-time macros of Config.h (see build_ref.sh)
int ptap_cfg_res_x = 1000, ptap_cfg_res_y = 800, ptap_cfg_iter = 500, ptap_cfg_depth = 5;
int ptap_cfg_grid_x = 25, ptap_cfg_grid_y = 25, ptap_cfg_grid_z = 25;

struct PtapProbe { int model, tri; float t_model, u, v; };
thread_local int ptap_probe_tri = -1;
thread_local float ptap_probe_t = 0.f, ptap_probe_u = 0.f, ptap_probe_v = 0.f;
PtapProbe* ptap_probe_out = nullptr;   // indexed by iray when non-null

#include "cuda_runtime_api.h"
thread_local ptap_uint3 threadIdx, blockIdx;
thread_local dim3 blockDim;
int ptap_ref_quiet = 1;

#include "Scene.cpp"
#undef CLAMP
#include "Renderer.cpp"

namespace {

struct RefRenderer {
    Renderer r;
    int capacity = 0;
    int nrays = 0;       // active rays of the wavefront in flight
    std::vector<PtapProbe> probe;
};

dim3 blocksFor(int capacity) { return dim3((unsigned)(capacity / 32)); }   // Renderer.cpp:572-573

struct Silence {
    int saved;
    Silence() { fflush(stdout); saved = dup(1); int nul = ::open("/dev/null", 1); dup2(nul, 1); ::close(nul); }
    ~Silence() { fflush(stdout); dup2(saved, 1); ::close(saved); }
};

}  // namespace

#include <fcntl.h>

extern "C" {

struct RefHit {
    int model, tri;          // winning model index and global triangle index (-1: miss)
    float t_model, dist;     // model-space t of the winner, world distance (FLOAT_MAX: miss)
    float u, v;
    float nx, ny, nz;        // world normal as stored by the reference
    int mat_type;
};

int ref_sizeof(int which)
{
    switch (which) {
        case 0: return sizeof(Model); case 1: return sizeof(Mesh); case 2: return sizeof(Vertex);
        case 3: return sizeof(Triangle); case 4: return sizeof(Grid); case 5: return sizeof(Voxel);
        case 6: return sizeof(EntityIndex); case 7: return sizeof(Ray); case 8: return sizeof(IntersectionData);
        case 9: return sizeof(Pixel); case 10: return sizeof(Material); case 11: return sizeof(RefHit);
    }
    return -1;
}

void ref_set_grid(int gx, int gy, int gz) { ptap_cfg_grid_x = gx; ptap_cfg_grid_y = gy; ptap_cfg_grid_z = gz; }
void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_max_threads() { return omp_get_max_threads(); }

// The scene exactly as Scene::Scene codes it (Scene.cpp:3-224); `root` is the directory that
// contains "Input data/".
void* ref_scene_builtin(const char* root)
{
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return nullptr;
    if (chdir(root) != 0) return nullptr;
    Scene* s = new Scene(std::string("Input data\\lucy.obj"));
    if (chdir(cwd) != 0) { delete s; return nullptr; }
    // Mesh/Model structs carry uninitialised padding-free but unset fields (refractive_index,
    // reflectivity); zero them so exported bytes are deterministic.
    for (auto& m : s->models) { m.mat.refractive_index = 0.f; m.mat.reflectivity = 0.f; }
    return s;
}

// A scene from caller arrays in the reference layouts; grids are built by the reference's own
// Scene::addMeshesToGrid (Scene.cpp:318-396).
void* ref_scene_from_arrays(const void* models, int nmodels, const void* meshes, int nmeshes,
                            const void* vertices, int nvertices, const void* triangles, int ntriangles)
{
    // Scene has no default constructor and its only constructor hard-codes a scene (Scene.cpp:3-224) and, when the
    // OBJ files are absent, walks uninitialised Mesh ranges.  The class is exactly seven std::vectors (Scene.h:26-32),
    // so build the members in place instead of running that constructor.
    Scene* s = static_cast<Scene*>(::operator new(sizeof(Scene)));
    new (&s->models) std::vector<Model>(); new (&s->meshes) std::vector<Mesh>(); new (&s->vertices) std::vector<Vertex>();
    new (&s->triangles) std::vector<Triangle>(); new (&s->grids) std::vector<Grid>(); new (&s->voxels) std::vector<Voxel>();
    new (&s->per_voxel_data_pool) std::vector<EntityIndex>();
    s->models.assign((const Model*)models, (const Model*)models + nmodels);
    s->meshes.assign((const Mesh*)meshes, (const Mesh*)meshes + nmeshes);
    s->vertices.assign((const Vertex*)vertices, (const Vertex*)vertices + nvertices);
    s->triangles.assign((const Triangle*)triangles, (const Triangle*)triangles + ntriangles);
    s->grids.clear(); s->voxels.clear(); s->per_voxel_data_pool.clear();
    s->addMeshesToGrid();
    return s;
}

void ref_scene_free(void* h) { delete (Scene*)h; }

void ref_scene_counts(void* h, int* out7)
{
    Scene* s = (Scene*)h;
    out7[0] = (int)s->models.size(); out7[1] = (int)s->meshes.size(); out7[2] = (int)s->vertices.size();
    out7[3] = (int)s->triangles.size(); out7[4] = (int)s->grids.size(); out7[5] = (int)s->voxels.size();
    out7[6] = (int)s->per_voxel_data_pool.size();
}

void ref_scene_get(void* h, int which, void* dst)
{
    Scene* s = (Scene*)h;
    switch (which) {
        case 0: memcpy(dst, s->models.data(), s->models.size() * sizeof(Model)); break;
        case 1: memcpy(dst, s->meshes.data(), s->meshes.size() * sizeof(Mesh)); break;
        case 2: memcpy(dst, s->vertices.data(), s->vertices.size() * sizeof(Vertex)); break;
        case 3: memcpy(dst, s->triangles.data(), s->triangles.size() * sizeof(Triangle)); break;
        case 4: memcpy(dst, s->grids.data(), s->grids.size() * sizeof(Grid)); break;
        case 5: memcpy(dst, s->voxels.data(), s->voxels.size() * sizeof(Voxel)); break;
        case 6: memcpy(dst, s->per_voxel_data_pool.data(), s->per_voxel_data_pool.size() * sizeof(EntityIndex)); break;
    }
}

// Overwrite the models array (material / transform edits); grid topology must not change.
void ref_scene_set_models(void* h, const void* models, int nmodels)
{
    Scene* s = (Scene*)h;
    s->models.assign((const Model*)models, (const Model*)models + nmodels);
}

void* ref_renderer_create(void* scene, int W, int H, int depth)
{
    ptap_cfg_res_x = W; ptap_cfg_res_y = H; ptap_cfg_depth = depth;
    RefRenderer* rr = new RefRenderer();
    {
        Silence q;
        rr->r.allocateOnGPU(*(Scene*)scene);   // Renderer.cpp:65-130
    }
    rr->capacity = W * H;
    rr->probe.assign(rr->capacity, PtapProbe{-1, -1, 0.f, 0.f, 0.f});
    return rr;
}

void ref_renderer_free(void* h)
{
    RefRenderer* rr = (RefRenderer*)h;
    rr->r.free();
    delete rr;
}

// ---- wavefront steps: the launches of Renderer::renderLoop, one call each ----------------

void ref_init_image(void* h)                               // Renderer.cpp:577
{
    RefRenderer* rr = (RefRenderer*)h;
    LAUNCH(initImageKernel, blocksFor(rr->capacity), dim3(32), rr->capacity, rr->r.render_data);
}

void ref_generate(void* h)                                 // Renderer.cpp:588-589
{
    RefRenderer* rr = (RefRenderer*)h;
    rr->nrays = rr->capacity;
    LAUNCH(generateRaysKernel, blocksFor(rr->capacity), dim3(32), rr->nrays, rr->r.render_data);
}

void ref_trace_step(void* h, int with_probe)               // Renderer.cpp:604 / 617
{
    RefRenderer* rr = (RefRenderer*)h;
    if (with_probe) {
        for (int i = 0; i < rr->nrays; ++i) rr->probe[i] = PtapProbe{-1, -1, 0.f, 0.f, 0.f};
        ptap_probe_out = rr->probe.data();
    }
    LAUNCH(computeRaySceneIntersectionKernel, blocksFor(rr->capacity), dim3(32), rr->nrays, rr->r.render_data);
    ptap_probe_out = nullptr;
}

void ref_shade_step(void* h, int iter)                     // Renderer.cpp:622
{
    RefRenderer* rr = (RefRenderer*)h;
    LAUNCH(shadeRayKernel, blocksFor(rr->capacity), dim3(32), rr->nrays, iter, rr->r.render_data);
}

int ref_compact_step(void* h)                              // Renderer.cpp:625-630
{
    RefRenderer* rr = (RefRenderer*)h;
    RenderData& rd = rr->r.render_data;
    LAUNCH(compactStencilKernel, blocksFor(rr->capacity), dim3(32), rr->nrays, rd.dev_ray_data->pool, rd.dev_stencil->pool);
    Ray* itr = thrust::stable_partition(thrust::device, rd.dev_ray_data->pool, rd.dev_ray_data->pool + rr->nrays,
                                        rd.dev_stencil->pool, hasTerminated());
    rr->nrays = (int)(itr - rd.dev_ray_data->pool);
    return rr->nrays;
}

void ref_gather(void* h)                                   // Renderer.cpp:638
{
    RefRenderer* rr = (RefRenderer*)h;
    LAUNCH(gatherImageDataKernel, blocksFor(rr->capacity), dim3(32), rr->r.render_data);
}

int ref_nrays(void* h) { return ((RefRenderer*)h)->nrays; }

// state read-back: whole arrays in the reference's own layouts
void ref_get_rays(void* h, void* dst, int n) { memcpy(dst, ((RefRenderer*)h)->r.render_data.dev_ray_data->pool, (size_t)n * sizeof(Ray)); }
void ref_set_rays(void* h, const void* src, int n) { RefRenderer* rr = (RefRenderer*)h; memcpy(rr->r.render_data.dev_ray_data->pool, src, (size_t)n * sizeof(Ray)); rr->nrays = n; }
void ref_get_hits(void* h, void* dst, int n) { memcpy(dst, ((RefRenderer*)h)->r.render_data.dev_intersection_data->pool, (size_t)n * sizeof(IntersectionData)); }
void ref_set_hits(void* h, const void* src, int n) { memcpy(((RefRenderer*)h)->r.render_data.dev_intersection_data->pool, src, (size_t)n * sizeof(IntersectionData)); }
void ref_get_image(void* h, void* dst) { RefRenderer* rr = (RefRenderer*)h; memcpy(dst, rr->r.render_data.dev_image_data->pool, (size_t)rr->capacity * sizeof(Pixel)); }
void ref_get_probe(void* h, void* dst, int n) { memcpy(dst, ((RefRenderer*)h)->probe.data(), (size_t)n * sizeof(PtapProbe)); }

// The reference's own loop, untouched (Renderer.cpp:567-648), for `iters` iterations.
// Used to check that the step-wise calls above reproduce it bit for bit, and as the
// `--impl reference` timing arm.
double ref_render_loop(void* h, int iters)
{
    RefRenderer* rr = (RefRenderer*)h;
    ptap_cfg_iter = iters;
    auto t0 = std::chrono::high_resolution_clock::now();
    {
        Silence q;
        rr->r.renderLoop();
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// The reference's BMP writer (Renderer.cpp:15-63) writes "Render.bmp" into the CWD.
int ref_write_bmp(void* h, const char* dir, int iters)
{
    RefRenderer* rr = (RefRenderer*)h;
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return -1;
    if (chdir(dir) != 0) return -2;
    ptap_cfg_iter = iters;
    rr->r.renderImage();
    return chdir(cwd);
}

// Oracle tier R1 for one ray slot: the statements of computeRaySceneIntersectionKernel (Renderer.cpp:363-409) with the grid query
// (Renderer.cpp:386) replaced by computeRayTriangleIntersection (Renderer.cpp:174-215) on EVERY triangle of the model's mesh.
// Everything that decides the result - per-model ray set-up, predicate, model t -> world distance, nearest-model rule - is the
// reference's own code; only the culling structure is absent.
static PtapProbe r1One(RenderData& rd, int iray)
{
    const int nmodels = rd.dev_model_data->size;
    Ray* ray = &rd.dev_ray_data->pool[iray];
    IntersectionData* hit_info = &rd.dev_intersection_data->pool[iray];
    float best = hit_info->impact_distance;
    glm::vec3 best_n(0.f); Material best_mat{}; PtapProbe best_p{-1, -1, 0.f, 0.f, 0.f};
    for (int imodel = 0; imodel < nmodels; ++imodel) {
        Model* model = &rd.dev_model_data->pool[imodel];
        // per-model ray set-up: the statements of Renderer.cpp:381-384
        ray->transformed.orig = transformPosition(ray->base.orig, model->world_to_model);
        ray->transformed.dir = glm::normalize(transformDirection(ray->base.dir, model->world_to_model));
        ray->cache.inv_dir = glm::vec3(1 / ray->transformed.dir.x, 1 / ray->transformed.dir.y, 1 / ray->transformed.dir.z);
        hit_info->impact_distance = FLOAT_MAX;
        const Mesh& mesh = rd.dev_mesh_data->pool[model->mesh_index];
        bool any = false;
        ptap_probe_tri = -1;
        for (int t = mesh.triangle_indices.start_index; t < mesh.triangle_indices.end_index; ++t)
            if (computeRayTriangleIntersection(rd, ray, hit_info, t)) any = true;
        if (any) {
            // model t -> world distance: the statements of Renderer.cpp:388-398
            glm::vec3 nd = glm::normalize(ray->transformed.dir);
            glm::vec3 pm = ray->transformed.orig + nd * hit_info->impact_distance;
            glm::vec3 pw = transformPosition(pm, model->model_to_world);
            hit_info->impact_distance = glm::length(pw - ray->base.orig);
            if (best > hit_info->impact_distance) {
                best = hit_info->impact_distance;
                best_mat = model->mat;
                best_n = glm::normalize(transformNormal(hit_info->impact_normal, model->model_to_world));
                best_p = PtapProbe{imodel, ptap_probe_tri, ptap_probe_t, ptap_probe_u, ptap_probe_v};
            }
        }
    }
    if (best < FLOAT_MAX) {
        hit_info->impact_distance = best; hit_info->impact_normal = best_n; hit_info->impact_mat = best_mat;
    }
    return best_p;
}

// The wavefront's trace launch at tier R1 (used for BVH film parity: the rest of the loop stays the reference's own kernels).
void ref_trace_step_r1(void* h)
{
    RefRenderer* rr = (RefRenderer*)h;
    RenderData& rd = rr->r.render_data;
    const int n = rr->nrays;
#pragma omp parallel for schedule(dynamic, 64)
    for (int iray = 0; iray < n; ++iray) rr->probe[iray] = r1One(rd, iray);
}

// ---- closest hit on a caller-supplied ray set ----------------------------------------------
// rays_od: n x 6 floats (origin, direction; the direction need not be normalised, exactly as
// Ray::base).  mode 0 = R0 (reference grid walk), mode 1 = R1 (brute force, same predicate).
int ref_trace(void* h, const float* rays_od, int n, int mode, RefHit* out)
{
    RefRenderer* rr = (RefRenderer*)h;
    RenderData& rd = rr->r.render_data;
    const int cap = rr->capacity - rr->capacity % 32;
    if (cap <= 0) return -1;
    for (int base = 0; base < n; base += cap) {
        const int m = std::min(cap, n - base);
        for (int i = 0; i < m; ++i) {
            Ray& ray = rd.dev_ray_data->pool[i];
            const float* p = rays_od + (size_t)(base + i) * 6;
            ray.base.orig = glm::vec3(p[0], p[1], p[2]);
            ray.base.dir = glm::vec3(p[3], p[4], p[5]);
            rd.dev_intersection_data->pool[i].impact_distance = FLOAT_MAX;        // Renderer.cpp:553
            rr->probe[i] = PtapProbe{-1, -1, 0.f, 0.f, 0.f};
        }
        if (mode == 0) {
            ptap_probe_out = rr->probe.data();
            LAUNCH(computeRaySceneIntersectionKernel, dim3((unsigned)((m + 31) / 32)), dim3(32), m, rd);
            ptap_probe_out = nullptr;
        } else {
#pragma omp parallel for schedule(dynamic, 64)
            for (int iray = 0; iray < m; ++iray) rr->probe[iray] = r1One(rd, iray);
        }
        for (int i = 0; i < m; ++i) {
            const IntersectionData& hd = rd.dev_intersection_data->pool[i];
            const PtapProbe& p = rr->probe[i];
            RefHit& o = out[base + i];
            if (hd.impact_distance < FLOAT_MAX && p.model >= 0) {
                o.model = p.model; o.tri = p.tri; o.t_model = p.t_model; o.dist = hd.impact_distance;
                o.u = p.u; o.v = p.v; o.nx = hd.impact_normal.x; o.ny = hd.impact_normal.y; o.nz = hd.impact_normal.z;
                o.mat_type = (int)hd.impact_mat.material_type;
            } else {
                o.model = -1; o.tri = -1; o.t_model = 0.f; o.dist = hd.impact_distance; o.u = o.v = 0.f;
                o.nx = o.ny = o.nz = 0.f; o.mat_type = -1;
            }
        }
    }
    return 0;
}

// known-answer probes for the unit tests of the restatement (utility.h:43-170)
unsigned ref_util_hash(unsigned a) { return utilHash(a); }
void ref_rng_u01(int iter, int index, int depth, int n, float* out)
{
    thrust::default_random_engine rng = makeSeededRandomEngine(iter, index, depth);
    thrust::uniform_real_distribution<float> u01(0, 1);
    for (int i = 0; i < n; ++i) out[i] = u01(rng);
}
void ref_scatter(int kind, const float* normal, const float* dir, int iter, int index, int depth, float* out3)
{
    thrust::default_random_engine rng = makeSeededRandomEngine(iter, index, depth);
    glm::vec3 n(normal[0], normal[1], normal[2]), d(dir[0], dir[1], dir[2]), r;
    if (kind == 0) r = calculateRandomDirectionInHemisphere(n, rng);
    else if (kind == 1) r = calculateMetalScattering(n, d, rng);
    else if (kind == 2) r = calculateCoatScattering(n, d, rng);
    else r = reflectRay(d, n);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

}  // extern "C"
