/* ptap_oracle.c - CPU restatement of PathTracerAP's render hot path.  TEST INFRASTRUCTURE
 * (see ptap_oracle.h for the rules on who may call it and for the parity-pinning status).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -fopenmp -fPIC -shared ptap_oracle.c -lm
 * -ffp-contract=off is mandatory: every product below must round before the following sum,
 * exactly as the host-compiled reference does.
 *
 * File:line citations are relative to /root/reference/PathTracerAP/ ; "glm/" means
 * external/include/glm/ (vendored glm 0.9.6.3).
 */
#include "ptap_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 ld3(const float* p) { return V(p[0], p[1], p[2]); }
static inline void st3(float* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
/* glm/detail/func_geometric.inl:65-72: tmp = a*b; tmp.x + tmp.y + tmp.z */
static inline float dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* glm/detail/func_geometric.inl:134-142 */
static inline v3 cross(v3 x, v3 y) { return V(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
/* glm/detail/func_geometric.inl:154-159 with func_exponential.inl:150-152: v * (1 / sqrt(dot(v,v))) */
static inline v3 normalize(v3 v) { return scale(v, 1.0f / sqrtf(dot(v, v))); }
/* glm/detail/func_geometric.inl:95-100 */
static inline float length3(v3 v) { return sqrtf(dot(v, v)); }
/* the shim compiles the reference with std::min/std::max (a CUDA build uses fminf/fmaxf; they agree for non-NaN) */
static inline float fmin_std(float a, float b) { return (b < a) ? b : a; }
static inline float fmax_std(float a, float b) { return (a < b) ? b : a; }
#define O_ABS(x) ((x) < 0 ? -(x) : (x))                                  /* utility.h:14 */
#define O_CLAMP(v, lo, hi) ((v) < (lo) ? (lo) : ((v) > (hi) ? (hi) : (v)))   /* utility.h:18 */

/* glm/detail/type_mat4x4.inl:617-627: (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w), column-major, xyz only.
 * utility.h:71-80: transformPosition uses w = 1, transformDirection w = 0. */
static inline v3 mat4_mul(const float* m, v3 v, float w)
{
    v3 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * w);
    return r;
}

/* utility.h:82-88: transpose(inverse(mat3(M))) * n.
 * inverse: glm/detail/type_mat3x3.inl:37-56; transpose: glm/detail/func_matrix.inl; mat3*vec3: type_mat3x3.inl:487-493.
 * m[c][r] of the upper-left 3x3 is M[4*c + r].  (T*n).r = Inv[r][0]*n.x + Inv[r][1]*n.y + Inv[r][2]*n.z. */
void oracle_normal_matrix(const float* M, float inv[9])
{
#define m(c, r) M[4 * (c) + (r)]
    const float ood = 1.0f / (+m(0, 0) * (m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2))
                              - m(1, 0) * (m(0, 1) * m(2, 2) - m(2, 1) * m(0, 2))
                              + m(2, 0) * (m(0, 1) * m(1, 2) - m(1, 1) * m(0, 2)));
    /* inv[3*c + r] = Inverse[c][r] */
    inv[0] = +(m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2)) * ood;
    inv[3] = -(m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2)) * ood;
    inv[6] = +(m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1)) * ood;
    inv[1] = -(m(0, 1) * m(2, 2) - m(2, 1) * m(0, 2)) * ood;
    inv[4] = +(m(0, 0) * m(2, 2) - m(2, 0) * m(0, 2)) * ood;
    inv[7] = -(m(0, 0) * m(2, 1) - m(2, 0) * m(0, 1)) * ood;
    inv[2] = +(m(0, 1) * m(1, 2) - m(1, 1) * m(0, 2)) * ood;
    inv[5] = -(m(0, 0) * m(1, 2) - m(1, 0) * m(0, 2)) * ood;
    inv[8] = +(m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1)) * ood;
#undef m
}
static inline v3 transform_normal(v3 n, const float* M)
{
    float inv[9];
    oracle_normal_matrix(M, inv);
    /* T = transpose(Inverse): T[c][r] = Inverse[r][c]; (T*n).r = T[0][r]*n.x + T[1][r]*n.y + T[2][r]*n.z */
    v3 r;
    r.x = inv[0] * n.x + inv[1] * n.y + inv[2] * n.z;
    r.y = inv[3] * n.x + inv[4] * n.y + inv[5] * n.z;
    r.z = inv[6] * n.x + inv[7] * n.y + inv[8] * n.z;
    return r;
}

/* ------------------------------------------------------------------ RNG: utility.h:43-62 */

unsigned oracle_hash(unsigned a)                                         /* utility.h:43-53 */
{
    a = (a + 0x7ed55d16u) + (a << 12);
    a = (a ^ 0xc761c23cu) ^ (a >> 19);
    a = (a + 0x165667b1u) + (a << 5);
    a = (a + 0xd3a2646cu) ^ (a << 9);
    a = (a + 0xfd7046c5u) + (a << 3);
    a = (a ^ 0xb55a4f09u) ^ (a >> 16);
    return a;
}

/* utility.h:57-62 then thrust::minstd_rand's seed rule (thrust/random/detail/linear_congruential_engine.inl:45-55):
 * state = seed mod (2^31-1), and 0 becomes 1. */
unsigned oracle_rng_seed(int iter, int index, int depth)
{
    unsigned h = oracle_hash(0x80000000u | ((unsigned)depth << 22) | (unsigned)iter) ^ oracle_hash((unsigned)index);
    unsigned x = h % 2147483647u;
    return x == 0u ? 1u : x;
}

/* minstd_rand step x <- 48271 x mod (2^31-1), then thrust's uniform_real_distribution<float>(0,1)
 * (thrust/random/detail/uniform_real_distribution.inl:61-75): float(x - min) / (1.0f + float(max - min)) with
 * min = 1, max = 2^31-2; the divisor is 2^31 after rounding, so 1.0f can be returned. */
float oracle_rng_next(unsigned* state)
{
    unsigned x = (unsigned)(((unsigned long long)*state * 48271ull) % 2147483647ull);
    *state = x;
    float result = (float)(x - 1u);
    result /= (1.0f + (float)2147483645u);
    return (result * (1.0f - 0.0f)) + 0.0f;
}

/* ------------------------------------------------------------------ scattering: utility.h:64-170 */

#define O_TWO_PI 6.2831853071795864769252867665590057683943f
#define O_SQRT_OF_ONE_THIRD 0.5773502691896257645091487805019574556476f

static v3 reflect_ray(v3 incident, v3 n)                                 /* utility.h:64-69: n - 2(i.n)n */
{
    return sub(n, scale(n, 2.0f * dot(incident, n)));
}

static v3 hemisphere(v3 normal, unsigned* rng)                           /* utility.h:91-123 */
{
    float up = sqrtf(oracle_rng_next(rng));
    float over = sqrtf(1 - up * up);
    float around = oracle_rng_next(rng) * O_TWO_PI;
    v3 not_normal;
    if (O_ABS(normal.x) < O_SQRT_OF_ONE_THIRD) not_normal = V(1, 0, 0);
    else if (O_ABS(normal.y) < O_SQRT_OF_ONE_THIRD) not_normal = V(0, 1, 0);
    else not_normal = V(0, 0, 1);
    v3 p1 = normalize(cross(normal, not_normal));
    v3 p2 = normalize(cross(normal, p1));
    /* up*normal + cos(around)*over*p1 + sin(around)*over*p2, left to right */
    return add(add(scale(normal, up), scale(p1, cosf(around) * over)), scale(p2, sinf(around) * over));
}

static v3 coat(v3 normal, v3 dir, unsigned* rng)                         /* utility.h:125-143 */
{
    float roulette = oracle_rng_next(rng);
    if (roulette < 0.5f) return reflect_ray(dir, normal);
    return hemisphere(normal, rng);
}

static v3 metal(v3 normal, v3 dir, unsigned* rng)                        /* utility.h:145-170 */
{
    float up = sqrtf(oracle_rng_next(rng));
    float over = sqrtf(1 - up * up);
    float around = oracle_rng_next(rng) * O_TWO_PI;
    (void)over; (void)around;                                            /* drawn but unused (utility.h:150-152) */
    float phi = O_TWO_PI * oracle_rng_next(rng);
    float r2 = oracle_rng_next(rng);
    float phongexponent = 30;
    float cosTheta = powf(1 - r2, 1.0f / (phongexponent + 1));
    float sinTheta = sqrtf(1 - cosTheta * cosTheta);
    /* w = normalize(dir - normal * 2.0f * dot(normal, dir)): (normal*2.0f)*dot */
    v3 w = normalize(sub(dir, scale(scale(normal, 2.0f), dot(normal, dir))));
    v3 a = ((double)O_ABS(w.x) > .1) ? V(0, 1, 0) : V(1, 0, 0);          /* `.1` is a double literal */
    v3 u = normalize(cross(a, w));
    v3 v = cross(w, u);
    /* u*cosf(phi)*sinTheta + v*sinf(phi)*sinTheta + w*cosTheta */
    return add(add(scale(scale(u, cosf(phi)), sinTheta), scale(scale(v, sinf(phi)), sinTheta)), scale(w, cosTheta));
}

void oracle_scatter(int kind, const float normal[3], const float dir[3], int iter, int index, int depth, float out[3])
{
    unsigned rng = oracle_rng_seed(iter, index, depth);
    v3 n = ld3(normal), d = ld3(dir), r;
    if (kind == 0) r = hemisphere(n, &rng);
    else if (kind == 1) r = metal(n, d, &rng);
    else if (kind == 2) r = coat(n, d, &rng);
    else r = reflect_ray(d, n);
    st3(out, r);
}

/* float -> int as x86 cvttss2si does it for the host-compiled reference: out of range -> INT_MIN */
static inline int f2i(float x) { return (x >= -2147483648.0f && x < 2147483648.0f) ? (int)x : (-2147483647 - 1); }

/* ------------------------------------------------------------------ grid build: Scene.cpp:293-396 */

static void bb_update(float* mn, float* mx, const float* p)              /* Primitive.h:49-58 */
{
    for (int k = 0; k < 3; ++k) {
        mn[k] = mn[k] > p[k] ? p[k] : mn[k];
        mx[k] = mx[k] < p[k] ? p[k] : mx[k];
    }
}

int oracle_build_grids(OModel* models, int nmodels, const OMesh* meshes, int nmeshes, const OVertex* vertices,
                       const OTriangle* triangles, const int gd[3],
                       OGrid* grids, int* ngrids, OVoxel* voxels, int* nvoxels, int* refs, int* nrefs)
{
    const int ncell = gd[0] * gd[1] * gd[2];
    char* processed = (char*)calloc((size_t)nmeshes, 1);
    int* cache = (int*)calloc((size_t)nmeshes, sizeof(int));
    int ng = 0, nv = 0, nr = 0;
    for (int i = 0; i < nmodels; ++i) {                                  /* Scene.cpp:323-395 */
        const int mi = models[i].mesh_index;
        if (processed[mi]) { models[i].grid_index = cache[mi]; continue; }
        processed[mi] = 1; cache[mi] = ng; models[i].grid_index = ng;
        const OMesh* mesh = &meshes[mi];
        float w[3];
        for (int k = 0; k < 3; ++k) w[k] = (mesh->bb_max[k] - mesh->bb_min[k]) / gd[k];   /* Scene.cpp:341-347 */
        /* The reference appends triangle t to a vector per covered cell, then flattens cell by cell (x fastest).
         * Counting pass + prefix sum + fill pass gives the same order: triangles ascending inside each cell. */
        int* offs = (int*)calloc((size_t)ncell + 1, sizeof(int));
        int* cursor = (int*)calloc((size_t)ncell, sizeof(int));
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                int run = 0;
                for (int c = 0; c < ncell; ++c) { int n = offs[c]; offs[c] = run; cursor[c] = run; run += n; }
                offs[ncell] = run;
                if (!refs) break;
            }
            for (int t = mesh->t_start; t < mesh->t_end; ++t) {
                float tmn[3] = {O_FLOAT_MAX, O_FLOAT_MAX, O_FLOAT_MAX}, tmx[3] = {O_FLOAT_MIN, O_FLOAT_MIN, O_FLOAT_MIN};
                for (int k = 0; k < 3; ++k) bb_update(tmn, tmx, vertices[triangles[t].v[k]].position);
                int lo[3], hi[3];
                for (int k = 0; k < 3; ++k) {                            /* Scene.cpp:300-315: floor(abs(.)/w) then clamp */
                    lo[k] = f2i(floorf(fabsf(mesh->bb_min[k] - tmn[k]) / w[k]));
                    hi[k] = f2i(floorf(fabsf(mesh->bb_min[k] - tmx[k]) / w[k]));
                    lo[k] = O_CLAMP(lo[k], 0, gd[k] - 1);
                    hi[k] = O_CLAMP(hi[k], 0, gd[k] - 1);
                }
                for (int z = lo[2]; z <= hi[2]; ++z)
                    for (int y = lo[1]; y <= hi[1]; ++y)
                        for (int x = lo[0]; x <= hi[0]; ++x) {
                            int c = x + y * gd[0] + gd[0] * gd[1] * z;
                            if (pass == 0) offs[c]++;
                            else refs[nr + cursor[c]++] = t;
                        }
            }
        }
        if (voxels)
            for (int c = 0; c < ncell; ++c) {
                voxels[nv + c].start = nr + offs[c]; voxels[nv + c].end = nr + offs[c + 1];
                voxels[nv + c].entity_type = O_ENTITY_TRIANGLE;
            }
        if (grids) {
            grids[ng].v_start = nv; grids[ng].v_end = nv + ncell;
            grids[ng].width[0] = w[0]; grids[ng].width[1] = w[1]; grids[ng].width[2] = w[2];
            grids[ng].entity_type = O_ENTITY_MODEL; grids[ng].entity_index = i;
        }
        nr += offs[ncell]; nv += ncell; ng += 1;
        free(offs); free(cursor);
    }
    free(processed); free(cache);
    *ngrids = ng; *nvoxels = nv; *nrefs = nr;
    return 0;
}

/* ------------------------------------------------------------------ intersection: Renderer.cpp:150-409 */

typedef struct {
    v3 o, d, inv;            /* ray->transformed.orig/dir, ray->cache.inv_dir */
    float dist;              /* hit_info->impact_distance */
    v3 normal;               /* hit_info->impact_normal */
    int tri; float t, u, v;  /* probe of the winner (not in the reference) */
} RayCtx;

/* Renderer.cpp:174-215 */
static int ray_triangle(const OScene* s, RayCtx* c, int itriangle)
{
    const OTriangle* tr = &s->triangles[itriangle];
    const OVertex* a = &s->vertices[tr->v[0]];
    const OVertex* b = &s->vertices[tr->v[1]];
    const OVertex* cc = &s->vertices[tr->v[2]];
    v3 v0 = ld3(a->position);
    v3 v0v1 = sub(ld3(b->position), v0);
    v3 v0v2 = sub(ld3(cc->position), v0);
    v3 pvec = cross(c->d, v0v2);
    float det = dot(v0v1, pvec);
    if (O_ABS(det - 0.0f) < O_EPSILON) return 0;                         /* IS_EQUAL(det, 0) */
    float invDet = 1 / det;
    v3 tvec = sub(c->o, v0);
    float u = dot(tvec, pvec) * invDet;
    if (u < 0.0f - O_EPSILON || u > 1.0f + O_EPSILON) return 0;
    v3 qvec = cross(tvec, v0v1);
    float v = dot(c->d, qvec) * invDet;
    if (v < 0.0f - O_EPSILON || (u + v) > 1.0f + O_EPSILON) return 0;
    float t = dot(v0v2, qvec) * invDet;
    if (t < 0.0f - O_EPSILON) return 0;
    /* flat normal: normalize((n0 + n1 + n2) * (1/3.0f)) */
    v3 normal = normalize(scale(add(add(ld3(a->normal), ld3(b->normal)), ld3(cc->normal)), 1 / 3.0f));
    if (c->dist > t) {
        c->dist = t; c->normal = normal;
        c->tri = itriangle; c->t = t; c->u = u; c->v = v;
    }
    return 1;
}

/* Renderer.cpp:150-170 */
static int ray_bbox(const RayCtx* c, const float* mn, const float* mx, float* t)
{
    float t1 = c->d.x == 0.0f ? O_FLOAT_MIN : (mn[0] - c->o.x) * c->inv.x;
    float t2 = c->d.x == 0.0f ? O_FLOAT_MAX : (mx[0] - c->o.x) * c->inv.x;
    float t3 = c->d.y == 0.0f ? O_FLOAT_MIN : (mn[1] - c->o.y) * c->inv.y;
    float t4 = c->d.y == 0.0f ? O_FLOAT_MAX : (mx[1] - c->o.y) * c->inv.y;
    float t5 = c->d.z == 0.0f ? O_FLOAT_MIN : (mn[2] - c->o.z) * c->inv.z;
    float t6 = c->d.z == 0.0f ? O_FLOAT_MAX : (mx[2] - c->o.z) * c->inv.z;
    float tmin = fmax_std(fmax_std(fmin_std(t1, t2), fmin_std(t3, t4)), fmin_std(t5, t6));
    float tmax = fmin_std(fmin_std(fmax_std(t1, t2), fmax_std(t3, t4)), fmax_std(t5, t6));
    if (tmax < 0 || tmin > tmax) return 0;
    *t = tmin;
    return 1;
}

/* Renderer.cpp:238-360 (with the missing `return false` of the bbox-miss path) */
static int ray_grid(const OScene* s, RayCtx* c, int igrid)
{
    const OGrid* g = &s->grids[igrid];
    /* bbox is reached through the grid's creating model (Renderer.cpp:245-249) */
    const OMesh* mesh = &s->meshes[s->models[g->entity_index].mesh_index];
    const float* mn = mesh->bb_min;
    const int GX = s->grid_dim[0], GY = s->grid_dim[1], GZ = s->grid_dim[2];
    float t_box;
    if (!ray_bbox(c, mn, mesh->bb_max, &t_box)) return 0;
    v3 p = add(c->o, scale(c->d, t_box));
    if ((p.x - mn[0]) < -O_EPSILON || (p.y - mn[1]) < -O_EPSILON || (p.z - mn[2]) < -O_EPSILON) return 0;
    int ix = f2i(O_ABS(p.x - mn[0] + O_EPSILON) / g->width[0]);
    int iy = f2i(O_ABS(p.y - mn[1] + O_EPSILON) / g->width[1]);
    int iz = f2i(O_ABS(p.z - mn[2] + O_EPSILON) / g->width[2]);
    ix = O_CLAMP(ix, 0, GX - 1); iy = O_CLAMP(iy, 0, GY - 1); iz = O_CLAMP(iz, 0, GZ - 1);
    float tmx = O_FLOAT_MAX, tmy = O_FLOAT_MAX, tmz = O_FLOAT_MAX;
    float dx = O_FLOAT_MAX, dy = O_FLOAT_MAX, dz = O_FLOAT_MAX;
    int step_x = c->d.x > 0.0f ? 1 : -1, step_y = c->d.y > 0.0f ? 1 : -1, step_z = c->d.z > 0.0f ? 1 : -1;
    int out_x = c->d.x > 0.0f ? GX : -1, out_y = c->d.y > 0.0f ? GY : -1, out_z = c->d.z > 0.0f ? GZ : -1;
    int nx = c->d.x > 0.0f ? ix + 1 : ix; float px = mn[0] + nx * g->width[0];
    int ny = c->d.y > 0.0f ? iy + 1 : iy; float py = mn[1] + ny * g->width[1];
    int nz = c->d.z > 0.0f ? iz + 1 : iz; float pz = mn[2] + nz * g->width[2];
    if (c->d.x != 0) { dx = O_ABS(g->width[0] * c->inv.x); tmx = (px - p.x) * c->inv.x; }
    if (c->d.y != 0) { dy = O_ABS(g->width[1] * c->inv.y); tmy = (py - p.y) * c->inv.y; }
    if (c->d.z != 0) { dz = O_ABS(g->width[2] * c->inv.z); tmz = (pz - p.z) * c->inv.z; }
    int cx = 0, cy = 0, cz = 0, is_intersect = 0;
    for (;;) {
        const OVoxel* vox = &s->voxels[g->v_start + ix + iy * GX + iz * GX * GY];
        int any = 0;                                                     /* Renderer.cpp:217-236 */
        if (vox->entity_type == O_ENTITY_TRIANGLE)
            for (int i = vox->start; i < vox->end; ++i)
                if (ray_triangle(s, c, s->refs[i])) any = 1;
        if (any) { cx = ix; cy = iy; cz = iz; is_intersect = 1; }
        if (is_intersect && (O_ABS(cx - ix) > 2 || O_ABS(cy - iy) > 2 || O_ABS(cz - iz) > 2)) return 1;
        if (tmx < tmy && tmx < tmz) {
            ix += step_x;
            if (ix == out_x || tmx >= O_FLOAT_MAX) return is_intersect;
            tmx += dx;
        } else if (tmy < tmz) {
            iy += step_y;
            if (iy == out_y || tmy >= O_FLOAT_MAX) return is_intersect;
            tmy += dy;
        } else {
            iz += step_z;
            if (iz == out_z || tmz >= O_FLOAT_MAX) return is_intersect;
            tmz += dz;
        }
    }
}

/* Per-ray accumulator of computeRaySceneIntersectionKernel's model loop (Renderer.cpp:372-409). */
typedef struct { float g_dist; v3 g_normal; OMaterial g_mat; OHit g_probe; float last_dist; } ModelLoop;

static void loop_begin(ModelLoop* L, const OHitRecord* h)
{
    L->g_dist = h->dist; L->g_normal = ld3(h->normal); L->g_mat = h->mat; L->last_dist = h->dist;
    memset(&L->g_probe, 0, sizeof L->g_probe); L->g_probe.model = -1; L->g_probe.tri = -1; L->g_probe.mat_type = -1;
}

static void model_setup(const OModel* model, v3 bo, v3 bd, RayCtx* c)
{
    c->o = mat4_mul(model->w2m, bo, 1.0f);                               /* Renderer.cpp:381 */
    c->d = normalize(mat4_mul(model->w2m, bd, 0.0f));                    /* Renderer.cpp:382 */
    c->inv = V(1 / c->d.x, 1 / c->d.y, 1 / c->d.z);                      /* Renderer.cpp:383 */
    c->dist = O_FLOAT_MAX;                                               /* Renderer.cpp:384 */
    c->tri = -1;
}

static void model_finish(ModelLoop* L, const OModel* model, int imodel, v3 bo, RayCtx* c, int hit)
{
    if (hit) {
        v3 nd = normalize(c->d);                                         /* Renderer.cpp:388 */
        v3 pm = add(c->o, scale(nd, c->dist));                           /* Renderer.cpp:389 */
        v3 pw = mat4_mul(model->m2w, pm, 1.0f);                          /* Renderer.cpp:390 */
        c->dist = length3(sub(pw, bo));                                  /* Renderer.cpp:391 */
        if (L->g_dist > c->dist) {                                       /* Renderer.cpp:393-398 */
            L->g_dist = c->dist; L->g_mat = model->mat;
            L->g_normal = normalize(transform_normal(c->normal, model->m2w));
            L->g_probe.model = imodel; L->g_probe.tri = c->tri; L->g_probe.t_model = c->t; L->g_probe.u = c->u; L->g_probe.v = c->v;
        }
    }
    L->last_dist = c->dist;                                              /* what this model left in the slot */
}

static void loop_end(ModelLoop* L, OHitRecord* h, OHit* probe)
{
    h->dist = L->last_dist;
    if (L->g_dist < O_FLOAT_MAX) {                                       /* Renderer.cpp:402-408 */
        h->dist = L->g_dist; st3(h->normal, L->g_normal); h->mat = L->g_mat;
        L->g_probe.dist = L->g_dist; st3(L->g_probe.normal, L->g_normal); L->g_probe.mat_type = L->g_mat.type;
    } else {
        L->g_probe.model = -1; L->g_probe.tri = -1; L->g_probe.dist = h->dist; L->g_probe.t_model = 0; L->g_probe.u = L->g_probe.v = 0;
    }
    if (probe) *probe = L->g_probe;
}

/* ---- probes of single steps of the path, for tests of algorithms that must reproduce ray_grid without walking the lists
 * (tests/test_emulation_model.py: the CPU model of PTAP_ACCEL_GRID_EMULATED).  Nothing here is used by oracle_trace. */

/* Every triangle of model `imodel`'s mesh that the reference's predicate accepts for the ray (Renderer.cpp:174-201 with the model's ray
 * set-up, :381-384), in triangle order: global triangle ids and model-space t.  Returns the number of hits (at most `cap` are stored). */
int oracle_model_hits(const OScene* s, const float* ray_od, int imodel, int cap, int* tri, float* t)
{
    const OModel* model = &s->models[imodel];
    const OMesh* mesh = &s->meshes[model->mesh_index];
    RayCtx c; memset(&c, 0, sizeof c);
    model_setup(model, ld3(ray_od), ld3(ray_od + 3), &c);
    int n = 0;
    for (int k = mesh->t_start; k < mesh->t_end; ++k) {
        c.dist = O_FLOAT_MAX; c.tri = -1;                                /* every triangle on its own: no nearest-hit state */
        if (!ray_triangle(s, &c, k)) continue;
        if (n < cap) { tri[n] = k; t[n] = c.tri == k ? c.t : O_FLOAT_MAX; }
        ++n;
    }
    return n;
}

/* The voxels computeRayGridIntersection (Renderer.cpp:238-360) visits for the ray in model `imodel` when no listed triangle is ever hit:
 * its slab test, entry point, entry voxel and DDA stepping, to the end of the grid.  ixyz receives up to `cap` (ix, iy, iz) triplets;
 * returns the number of voxels, 0 when the walk is not entered.  The statements are those of ray_grid above. */
int oracle_grid_path(const OScene* s, const float* ray_od, int imodel, int cap, int* ixyz)
{
    const OModel* model = &s->models[imodel];
    RayCtx cc; memset(&cc, 0, sizeof cc);
    model_setup(model, ld3(ray_od), ld3(ray_od + 3), &cc);
    const RayCtx* c = &cc;
    const OGrid* g = &s->grids[model->grid_index];
    const OMesh* mesh = &s->meshes[s->models[g->entity_index].mesh_index];
    const float* mn = mesh->bb_min;
    const int GX = s->grid_dim[0], GY = s->grid_dim[1], GZ = s->grid_dim[2];
    float t_box;
    if (!ray_bbox(c, mn, mesh->bb_max, &t_box)) return 0;
    v3 p = add(c->o, scale(c->d, t_box));
    if ((p.x - mn[0]) < -O_EPSILON || (p.y - mn[1]) < -O_EPSILON || (p.z - mn[2]) < -O_EPSILON) return 0;
    int ix = f2i(O_ABS(p.x - mn[0] + O_EPSILON) / g->width[0]);
    int iy = f2i(O_ABS(p.y - mn[1] + O_EPSILON) / g->width[1]);
    int iz = f2i(O_ABS(p.z - mn[2] + O_EPSILON) / g->width[2]);
    ix = O_CLAMP(ix, 0, GX - 1); iy = O_CLAMP(iy, 0, GY - 1); iz = O_CLAMP(iz, 0, GZ - 1);
    float tmx = O_FLOAT_MAX, tmy = O_FLOAT_MAX, tmz = O_FLOAT_MAX;
    float dx = O_FLOAT_MAX, dy = O_FLOAT_MAX, dz = O_FLOAT_MAX;
    int step_x = c->d.x > 0.0f ? 1 : -1, step_y = c->d.y > 0.0f ? 1 : -1, step_z = c->d.z > 0.0f ? 1 : -1;
    int out_x = c->d.x > 0.0f ? GX : -1, out_y = c->d.y > 0.0f ? GY : -1, out_z = c->d.z > 0.0f ? GZ : -1;
    int nx = c->d.x > 0.0f ? ix + 1 : ix; float px = mn[0] + nx * g->width[0];
    int ny = c->d.y > 0.0f ? iy + 1 : iy; float py = mn[1] + ny * g->width[1];
    int nz = c->d.z > 0.0f ? iz + 1 : iz; float pz = mn[2] + nz * g->width[2];
    if (c->d.x != 0) { dx = O_ABS(g->width[0] * c->inv.x); tmx = (px - p.x) * c->inv.x; }
    if (c->d.y != 0) { dy = O_ABS(g->width[1] * c->inv.y); tmy = (py - p.y) * c->inv.y; }
    if (c->d.z != 0) { dz = O_ABS(g->width[2] * c->inv.z); tmz = (pz - p.z) * c->inv.z; }
    int n = 0;
    for (;;) {
        if (n < cap) { ixyz[3 * n] = ix; ixyz[3 * n + 1] = iy; ixyz[3 * n + 2] = iz; }
        ++n;
        if (tmx < tmy && tmx < tmz) {
            ix += step_x;
            if (ix == out_x || tmx >= O_FLOAT_MAX) return n;
            tmx += dx;
        } else if (tmy < tmz) {
            iy += step_y;
            if (iy == out_y || tmy >= O_FLOAT_MAX) return n;
            tmy += dy;
        } else {
            iz += step_z;
            if (iz == out_z || tmz >= O_FLOAT_MAX) return n;
            tmz += dz;
        }
    }
}

/* World distance of the hit at model-space parameter t of model `imodel` along the ray, as Renderer.cpp:388-391 computes it. */
float oracle_hit_distance(const OScene* s, const float* ray_od, int imodel, float t)
{
    const OModel* model = &s->models[imodel];
    RayCtx c; memset(&c, 0, sizeof c);
    v3 bo = ld3(ray_od);
    model_setup(model, bo, ld3(ray_od + 3), &c);
    v3 nd = normalize(c.d);
    v3 pm = add(c.o, scale(nd, t));
    v3 pw = mat4_mul(model->m2w, pm, 1.0f);
    return length3(sub(pw, bo));
}

/* One thread of computeRaySceneIntersectionKernel (Renderer.cpp:363-409); h->dist on entry is the slot's incoming
 * hit_info->impact_distance.  Returns the final hit_info fields through *h, ids through *probe. */
static void trace_one(const OScene* s, v3 bo, v3 bd, int mode, OHitRecord* h, OHit* probe)
{
    ModelLoop L; loop_begin(&L, h);
    RayCtx c; memset(&c, 0, sizeof c);
    for (int imodel = 0; imodel < s->nmodels; ++imodel) {
        const OModel* model = &s->models[imodel];
        model_setup(model, bo, bd, &c);
        int hit;
        if (mode == 0) hit = ray_grid(s, &c, model->grid_index);
        else {
            const OMesh* mesh = &s->meshes[model->mesh_index];
            hit = 0;
            for (int t = mesh->t_start; t < mesh->t_end; ++t) if (ray_triangle(s, &c, t)) hit = 1;
        }
        model_finish(&L, model, imodel, bo, &c, hit);
    }
    loop_end(&L, h, probe);
}

/* Tier R1 for a block of rays: the same statements per (ray, model, triangle) as trace_one(mode 1), with the triangle loop outermost
 * inside a model so that a large mesh streams through the cache once per block instead of once per ray.  Each ray still sees its
 * model's triangles in ascending order, so every comparison (and therefore every result bit) is the one trace_one makes. */
#define O_R1_BLOCK 128
static void trace_block_r1(const OScene* s, int n, const v3* bo, const v3* bd, OHitRecord** h, OHit** probe)
{
    ModelLoop L[O_R1_BLOCK]; RayCtx c[O_R1_BLOCK]; int hit[O_R1_BLOCK];
    for (int r = 0; r < n; ++r) { loop_begin(&L[r], h[r]); memset(&c[r], 0, sizeof c[r]); }
    for (int imodel = 0; imodel < s->nmodels; ++imodel) {
        const OModel* model = &s->models[imodel];
        const OMesh* mesh = &s->meshes[model->mesh_index];
        for (int r = 0; r < n; ++r) { model_setup(model, bo[r], bd[r], &c[r]); hit[r] = 0; }
        for (int t = mesh->t_start; t < mesh->t_end; ++t)
            for (int r = 0; r < n; ++r) if (ray_triangle(s, &c[r], t)) hit[r] = 1;
        for (int r = 0; r < n; ++r) model_finish(&L[r], model, imodel, bo[r], &c[r], hit[r]);
    }
    for (int r = 0; r < n; ++r) loop_end(&L[r], h[r], probe ? probe[r] : NULL);
}

void oracle_trace(const OScene* s, const float* rays_od, int n, int mode, OHit* out)
{
    if (mode == 1) {
        int nblocks = (n + O_R1_BLOCK - 1) / O_R1_BLOCK;
#pragma omp parallel for schedule(dynamic, 1)
        for (int b = 0; b < nblocks; ++b) {
            int i0 = b * O_R1_BLOCK, m = n - i0 < O_R1_BLOCK ? n - i0 : O_R1_BLOCK;
            v3 bo[O_R1_BLOCK], bd[O_R1_BLOCK]; OHitRecord hr[O_R1_BLOCK]; OHitRecord* hp[O_R1_BLOCK]; OHit* pp[O_R1_BLOCK];
            for (int r = 0; r < m; ++r) {
                bo[r] = ld3(rays_od + 6 * (size_t)(i0 + r)); bd[r] = ld3(rays_od + 6 * (size_t)(i0 + r) + 3);
                memset(&hr[r], 0, sizeof hr[r]); hr[r].dist = O_FLOAT_MAX;   /* Renderer.cpp:553 */
                hp[r] = &hr[r]; pp[r] = &out[i0 + r];
            }
            trace_block_r1(s, m, bo, bd, hp, pp);
        }
        return;
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        OHitRecord h; memset(&h, 0, sizeof h); h.dist = O_FLOAT_MAX;     /* Renderer.cpp:553 */
        trace_one(s, ld3(rays_od + 6 * (size_t)i), ld3(rays_od + 6 * (size_t)i + 3), mode, &h, &out[i]);
    }
}

/* ------------------------------------------------------------------ wavefront: Renderer.cpp:411-648 */

struct OWavefront {
    const OScene* s; int W, H, depth, N, nrays;
    ORay* rays; ORay* tmp; OHitRecord* hits; OHitRecord* cache; OHit* probe; float* image; int* stencil;
    int cached;
    int mode;            /* closest-hit tier of oracle_trace_step: 0 = R0 (grid walk, the reference as written), 1 = R1 (every triangle) */
};

OWavefront* oracle_wavefront_create(const OScene* s, int W, int H, int depth)
{
    OWavefront* w = (OWavefront*)calloc(1, sizeof *w);
    w->s = s; w->W = W; w->H = H; w->depth = depth; w->N = W * H;
    size_t n = (size_t)w->N;
    w->rays = (ORay*)calloc(n, sizeof(ORay)); w->tmp = (ORay*)calloc(n, sizeof(ORay));
    w->hits = (OHitRecord*)calloc(n, sizeof(OHitRecord)); w->cache = (OHitRecord*)calloc(n, sizeof(OHitRecord));
    w->probe = (OHit*)calloc(n, sizeof(OHit)); w->image = (float*)calloc(n * 3, sizeof(float));
    w->stencil = (int*)calloc(n, sizeof(int));
    return w;
}

void oracle_wavefront_free(OWavefront* w)
{
    if (!w) return;
    free(w->rays); free(w->tmp); free(w->hits); free(w->cache); free(w->probe); free(w->image); free(w->stencil); free(w);
}

void oracle_wavefront_set_mode(OWavefront* w, int mode) { w->mode = mode; }
int oracle_nrays(const OWavefront* w) { return w->nrays; }
ORay* oracle_rays(OWavefront* w) { return w->rays; }
OHitRecord* oracle_hits(OWavefront* w) { return w->hits; }
OHit* oracle_probe(OWavefront* w) { return w->probe; }
float* oracle_image(OWavefront* w) { return w->image; }
void oracle_set_threads(int n) { omp_set_num_threads(n); }
int oracle_max_threads(void) { return omp_get_max_threads(); }

/* the reference launches N/32 (integer division) blocks of 32 once and reuses that grid (Renderer.cpp:572-573) */
static int launch_span(const OWavefront* w) { return (w->N / 32) * 32; }

void oracle_init_image(OWavefront* w)                                    /* Renderer.cpp:557-565 */
{
    int span = launch_span(w);
    for (int i = 0; i < span; ++i) { w->image[3 * i] = w->image[3 * i + 1] = w->image[3 * i + 2] = 0.0f; }
    w->cached = 0;
}

void oracle_generate(OWavefront* w)                                      /* Renderer.cpp:521-555 */
{
    w->nrays = w->N;
    int span = launch_span(w);
#pragma omp parallel for
    for (int i = 0; i < span; ++i) {
        int y = i / (w->W * 1);
        int x = i % (w->W * 1);
        float step_x = (float)(20.0 / (w->W * 1));
        float step_y = (float)(16.0 / (w->H * 1));
        float world_x = (float)(-10.0 + x * step_x);                      /* double + float*float (int converted to float) */
        float world_y = (float)(-4.0 + y * step_y);
        float world_z = 900.0f;
        ORay* r = &w->rays[i];
        r->orig[0] = 0; r->orig[1] = 0; r->orig[2] = 920.0f;
        r->dir[0] = world_x - 0.0f; r->dir[1] = world_y - 0.0f; r->dir[2] = world_z - 920.0f;
        r->color[0] = r->color[1] = r->color[2] = 1.0f;
        r->remaining_bounces = w->depth;
        r->ipixel = i;
        w->hits[i].dist = O_FLOAT_MAX;
        w->hits[i].ipixel = i;
    }
}

void oracle_trace_step(OWavefront* w)                                    /* Renderer.cpp:363-409 */
{
    int span = launch_span(w);
    int n = w->nrays < span ? w->nrays : span;
    if (w->mode == 1) {
        int nblocks = (n + O_R1_BLOCK - 1) / O_R1_BLOCK;
#pragma omp parallel for schedule(dynamic, 1)
        for (int b = 0; b < nblocks; ++b) {
            int i0 = b * O_R1_BLOCK, m = n - i0 < O_R1_BLOCK ? n - i0 : O_R1_BLOCK;
            v3 bo[O_R1_BLOCK], bd[O_R1_BLOCK]; OHitRecord* hp[O_R1_BLOCK]; OHit* pp[O_R1_BLOCK];
            for (int r = 0; r < m; ++r) {
                bo[r] = ld3(w->rays[i0 + r].orig); bd[r] = ld3(w->rays[i0 + r].dir);
                hp[r] = &w->hits[i0 + r]; pp[r] = &w->probe[i0 + r];
            }
            trace_block_r1(w->s, m, bo, bd, hp, pp);
        }
        return;
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i)
        trace_one(w->s, ld3(w->rays[i].orig), ld3(w->rays[i].dir), w->mode, &w->hits[i], &w->probe[i]);
}

void oracle_shade_step(OWavefront* w, int iter)                          /* Renderer.cpp:411-479 */
{
    int span = launch_span(w);
    int n = w->nrays < span ? w->nrays : span;
#pragma omp parallel for schedule(static, 1024)
    for (int i = 0; i < n; ++i) {
        ORay* ray = &w->rays[i];
        OHitRecord* hit = &w->hits[i];
        v3 col = ld3(ray->color);
        if (ray->remaining_bounces <= 0) col = mul(col, V(0.01f, 0.01f, 0.01f));      /* Renderer.cpp:421-424 */
        if (hit->dist < O_FLOAT_MAX) {
            v3 dir = normalize(ld3(ray->dir));                           /* Renderer.cpp:428 */
            v3 n_ = ld3(hit->normal);
            v3 pt = add(ld3(ray->orig), scale(dir, hit->dist));          /* Renderer.cpp:429 */
            if (ray->remaining_bounces > 0) {
                int type = hit->mat.type;
                v3 albedo = ld3(hit->mat.color);
                if (type == O_DIFFUSE || type == O_METAL || type == O_COAT) {
                    unsigned rng = oracle_rng_seed(iter, i, ray->remaining_bounces);
                    v3 nd = type == O_DIFFUSE ? hemisphere(n_, &rng) : type == O_METAL ? metal(n_, dir, &rng) : coat(n_, dir, &rng);
                    st3(ray->dir, nd);
                    st3(ray->orig, add(pt, scale(n_, 0.1f)));            /* intersection_pt + 0.1f * normal */
                    col = mul(col, albedo);
                } else if (type == O_EMISSIVE) {                         /* Renderer.cpp:454-460 */
                    ray->remaining_bounces = 0;
                    col = mul(col, albedo);
                    st3(ray->color, col);
                    hit->dist = O_FLOAT_MAX;
                    continue;
                } else if (type == O_REFLECTIVE) {                       /* Renderer.cpp:461-467 */
                    col = mul(col, albedo);
                    v3 refl = reflect_ray(dir, n_);
                    st3(ray->orig, add(pt, scale(n_, 0.1f)));
                    st3(ray->dir, refl);
                }
            }
            hit->dist = O_FLOAT_MAX;
        } else {                                                         /* Renderer.cpp:471-477 */
            ray->remaining_bounces = 0;
            col = mul(col, V(0.01f, 0.01f, 0.01f));
            st3(ray->color, col);
            hit->dist = O_FLOAT_MAX;
            continue;
        }
        st3(ray->color, col);
        ray->remaining_bounces--;
    }
}

int oracle_compact_step(OWavefront* w)                                   /* Renderer.cpp:506-519, 628-630 */
{
    int span = launch_span(w);
    int n = w->nrays;
    for (int i = 0; i < n && i < span; ++i) w->stencil[i] = w->rays[i].remaining_bounces <= 0 ? 0 : 1;
    /* thrust::stable_partition by stencil == 1: alive first, terminated after, both in order */
    int k = 0;
    for (int i = 0; i < n; ++i) if (w->stencil[i] == 1) w->tmp[k++] = w->rays[i];
    int alive = k;
    for (int i = 0; i < n; ++i) if (w->stencil[i] != 1) w->tmp[k++] = w->rays[i];
    memcpy(w->rays, w->tmp, (size_t)n * sizeof(ORay));
    w->nrays = alive;
    return alive;
}

void oracle_gather(OWavefront* w)                                        /* Renderer.cpp:481-496 */
{
    int span = launch_span(w);
    for (int i = 0; i < span; ++i) {
        ORay* r = &w->rays[i];
        r->color[0] = sqrtf(r->color[0]); r->color[1] = sqrtf(r->color[1]); r->color[2] = sqrtf(r->color[2]);
        float avg = (float)(1 / (1 * 1));
        float* px = &w->image[3 * (size_t)r->ipixel];
        px[0] += avg * r->color[0]; px[1] += avg * r->color[1]; px[2] += avg * r->color[2];
    }
}

void oracle_render(OWavefront* w, int iter_begin, int iter_end, int first_hit_cache, long long* rays_traced)
{
    long long traced = 0;
    for (int iter = iter_begin; iter < iter_end; ++iter) {               /* Renderer.cpp:582-644 */
        oracle_generate(w);
        int ibounce = 0;
        while (1) {
            if (ibounce == 0 && first_hit_cache && w->cached) {
                memcpy(w->hits, w->cache, (size_t)w->N * sizeof(OHitRecord));          /* Renderer.cpp:596-600 */
            } else {
                oracle_trace_step(w); traced += w->nrays;
                if (ibounce == 0 && first_hit_cache) { memcpy(w->cache, w->hits, (size_t)w->N * sizeof(OHitRecord)); w->cached = 1; }
            }
            oracle_shade_step(w, iter);
            if (oracle_compact_step(w) == 0) break;
            ibounce++;
        }
        oracle_gather(w);
    }
    if (rays_traced) *rays_traced += traced;
}

/* Renderer.cpp:15-63: 54-byte header, rows bottom-up as stored, bytes (char)(sum/ITER*255) in x,y,z order, no padding */
int oracle_write_bmp(const float* image_sum, int W, int H, int iters, const char* path)
{
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    unsigned char hdr[54]; memset(hdr, 0, sizeof hdr);
    hdr[0] = 'B'; hdr[1] = 'M'; hdr[10] = 54; hdr[14] = 40;
    for (int k = 0; k < 4; ++k) { hdr[18 + k] = (unsigned char)(W >> (8 * k)); hdr[22 + k] = (unsigned char)(H >> (8 * k)); }
    hdr[26] = 1; hdr[28] = 24;
    int fileSize = 54 + 3 * W * H, imageSize = 3 * W * H;
    memcpy(hdr + 2, &fileSize, 4); memcpy(hdr + 34, &imageSize, 4);
    fwrite(hdr, 1, 54, f);
    unsigned char* row = (unsigned char*)malloc((size_t)3 * W);
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            float div = 1 / (float)iters;
            const float* c = &image_sum[3 * ((size_t)x + (size_t)y * W)];
            for (int k = 0; k < 3; ++k) row[3 * x + k] = (unsigned char)(int)((c[k] * div) * 255.0f);
        }
        fwrite(row, 1, (size_t)3 * W, f);
    }
    free(row);
    fclose(f);
    return 0;
}
