/* ptap_oracle.h - CPU restatement of PathTracerAP's render hot path (plain C).
 *
 * TEST INFRASTRUCTURE.  This is the parity oracle: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build, load or call it.  The product
 * (pathtracerap_b200/csrc, libptap.so) never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so the oracle
 * is pinned against the reference's own sources compiled for the host (oracle/_ref, built by
 * oracle/build_ref.sh from /root/reference): tests/test_oracle_vs_ref.py requires bit-equal
 * hits, wavefront state and film, and tests/golden/ holds vectors dumped from that build
 * (tools/make_golden.py) for machines where /root/reference is absent.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/PathTracerAP/).  Arithmetic is IEEE binary32, never contracted
 * (compile with -ffp-contract=off), in the evaluation order of the vendored glm 0.9.6.3.
 */
#ifndef PTAP_ORACLE_H
#define PTAP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Config.h:4-6 */
#define O_EPSILON 0.005f
#define O_FLOAT_MAX 9999999.0f
#define O_FLOAT_MIN (-9999990.0f)

/* PODs with the reference's layouts (Primitive.h:23-178) */
typedef struct { int type; float refractive_index, reflectivity; float color[3]; } OMaterial;          /* 24 B */
typedef struct { int grid_index, mesh_index; float m2w[16], w2m[16]; OMaterial mat; } OModel;          /* 160 B, column-major */
typedef struct { int v_start, v_end, t_start, t_end; float bb_min[3], bb_max[3]; } OMesh;              /* 40 B */
typedef struct { float position[3], normal[3], uv[2]; } OVertex;                                       /* 32 B */
typedef struct { int v[3]; } OTriangle;                                                                /* 12 B */
typedef struct { int v_start, v_end; float width[3]; int entity_type, entity_index; } OGrid;           /* 28 B */
typedef struct { int start, end, entity_type; } OVoxel;                                                /* 12 B */
typedef struct { float orig[3], dir[3], t_orig[3], t_dir[3], inv_dir[3]; int ipixel, remaining_bounces; float color[3]; } ORay; /* 80 B */
typedef struct { float dist; float normal[3]; OMaterial mat; int ipixel; } OHitRecord;                 /* 44 B */

enum { O_DIFFUSE = 0, O_SPECULAR, O_REFLECTIVE, O_REFRACTIVE, O_EMISSIVE, O_COAT, O_METAL };           /* Primitive.h:70-79 */
enum { O_ENTITY_MODEL = 0, O_ENTITY_SCENE, O_ENTITY_TRIANGLE, O_ENTITY_SPHERE };                       /* Primitive.h:110-115 */

typedef struct {
    const OModel* models; int nmodels;
    const OMesh* meshes; int nmeshes;
    const OVertex* vertices; int nvertices;
    const OTriangle* triangles; int ntriangles;
    const OGrid* grids; int ngrids;
    const OVoxel* voxels; int nvoxels;
    const int* refs; int nrefs;
    int grid_dim[3];                      /* GRID_X, GRID_Y, GRID_Z (Config.h:8-10) */
} OScene;

/* Closest-hit result with the primitive id and barycentrics the reference does not record. */
typedef struct {
    int model, tri;          /* winning model and GLOBAL triangle index; -1 = miss */
    float t_model, dist;     /* model-space t of the winner; world distance (O_FLOAT_MAX = miss) */
    float u, v;
    float normal[3];         /* world normal as the reference stores it */
    int mat_type;
} OHit;                      /* 40 B */

/* utility.h:43-62 */
unsigned oracle_hash(unsigned a);
unsigned oracle_rng_seed(int iter, int index, int depth);       /* state after construction */
float oracle_rng_next(unsigned* state);                         /* one uniform_real_distribution<float>(0,1) draw */
/* utility.h:64-170; kind 0 hemisphere, 1 metal, 2 coat, 3 mirror ("reflectRay") */
void oracle_scatter(int kind, const float normal[3], const float dir[3], int iter, int index, int depth, float out[3]);

/* Scene.cpp:293-396.  Two-call protocol: pass voxels/refs NULL to get the counts. */
int oracle_build_grids(OModel* models, int nmodels, const OMesh* meshes, int nmeshes, const OVertex* vertices,
                       const OTriangle* triangles, const int grid_dim[3],
                       OGrid* grids, int* ngrids, OVoxel* voxels, int* nvoxels, int* refs, int* nrefs);

/* Renderer.cpp:363-409 on a caller ray set (n x 6 floats: origin, un-normalised direction).
 * mode 0 = R0 (grid walk, Renderer.cpp:238-360); mode 1 = R1 (every triangle of the model's mesh). */
void oracle_trace(const OScene* s, const float* rays_od, int n, int mode, OHit* out);

/* Probes of single steps of the path (tests/test_emulation_model.py): all hits of one model under the reference's predicate, the voxel
 * sequence of one model's grid walk when nothing is hit, and the reference's model-t -> world-distance conversion. */
int oracle_model_hits(const OScene* s, const float* ray_od, int imodel, int cap, int* tri, float* t);
int oracle_grid_path(const OScene* s, const float* ray_od, int imodel, int cap, int* ixyz);
float oracle_hit_distance(const OScene* s, const float* ray_od, int imodel, float t);

/* Wavefront state for one frame buffer; mirrors RenderData (Renderer.h:19-35). */
typedef struct OWavefront OWavefront;
OWavefront* oracle_wavefront_create(const OScene* s, int W, int H, int depth);
void oracle_wavefront_free(OWavefront* w);
/* Closest-hit tier of the wavefront's trace step: 0 (default) = R0, the reference's grid walk; 1 = R1, the same per-model ray set-up,
 * predicate and nearest-model rule applied to EVERY triangle (what an exact acceleration structure must reproduce).  Everything else
 * of the loop (shade, compaction, gather, seeds) is the reference's. */
void oracle_wavefront_set_mode(OWavefront* w, int mode);
void oracle_init_image(OWavefront* w);                          /* Renderer.cpp:557-565 */
void oracle_generate(OWavefront* w);                            /* Renderer.cpp:521-555 */
void oracle_trace_step(OWavefront* w);                          /* Renderer.cpp:363-409 */
void oracle_shade_step(OWavefront* w, int iter);                /* Renderer.cpp:411-479 */
int oracle_compact_step(OWavefront* w);                         /* Renderer.cpp:506-519, 628-630 */
void oracle_gather(OWavefront* w);                              /* Renderer.cpp:481-496 */
int oracle_nrays(const OWavefront* w);
ORay* oracle_rays(OWavefront* w);
OHitRecord* oracle_hits(OWavefront* w);
OHit* oracle_probe(OWavefront* w);                              /* ids/barycentrics of the last trace step */
float* oracle_image(OWavefront* w);                             /* W*H*3 running sum */
/* Renderer.cpp:567-648 for iterations [iter_begin, iter_end); rays_traced (may be NULL) += rays actually traced. */
void oracle_render(OWavefront* w, int iter_begin, int iter_end, int first_hit_cache, long long* rays_traced);
/* Renderer.cpp:15-63 */
int oracle_write_bmp(const float* image_sum, int W, int H, int iters, const char* path);

/* synth.c: scene inputs without the product library.  glm::translate * rotate(Y) * scale and its inverse (Scene.cpp:34-39); the displaced
 * icosphere of the synthetic workloads (3 vertices per triangle, triangle t = vertices 3t .. 3t+2; returns the triangle count). */
void oracle_compose_trs(const float translate[3], float rotate_y_degrees, const float scale[3], float model_to_world[16], float world_to_model[16]);
int oracle_icosphere_triangles(int level);
int oracle_icosphere(int level, float radius, float displacement, unsigned seed, OVertex* vertices, float bb_min[3], float bb_max[3]);

void oracle_set_threads(int n);
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
