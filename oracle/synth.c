/* synth.c - inputs for the CPU arms of bench.py and for the tests: the synthetic displaced icosphere of BASELINE.json's configs[1]/[3]
 *  and the translate * rotateY * scale model matrices every model of the reference's scene is composed from
 * (Scene.cpp:34-39), as plain C.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE (part of libptap_oracle.so).  It exists so that `bench.py --impl reference` and the oracle-side tests
 * can build their scene arrays WITHOUT loading the product library; tests/test_host.py checks that these arrays are byte-identical to
 * the ones libptap's own host-side generator (ptap_scene_add_icosphere, ptap_compose_trs) produces. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ptap_oracle.h"

/* ---- glm 0.9.6.3 in binary32, evaluation order of external/include/glm/gtc/matrix_transform.inl:40-134 and
 * glm/detail/type_mat4x4.inl:37-92, 686-704; column-major ---- */
typedef struct { float v[16]; } M4;

static M4 m4_identity(void) { M4 m; memset(&m, 0, sizeof m); m.v[0] = m.v[5] = m.v[10] = m.v[15] = 1.0f; return m; }

static M4 m4_mul(const M4* a, const M4* b)                       /* type_mat4x4.inl:686-704 */
{
    M4 r;
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 4; ++k)
            r.v[4 * c + k] = ((a->v[k] * b->v[4 * c] + a->v[4 + k] * b->v[4 * c + 1]) + a->v[8 + k] * b->v[4 * c + 2]) + a->v[12 + k] * b->v[4 * c + 3];
    return r;
}

static M4 m4_scale(const float s[3])                             /* matrix_transform.inl:122-134 */
{
    M4 m = m4_identity(), r;
    for (int k = 0; k < 4; ++k) { r.v[k] = m.v[k] * s[0]; r.v[4 + k] = m.v[4 + k] * s[1]; r.v[8 + k] = m.v[8 + k] * s[2]; r.v[12 + k] = m.v[12 + k]; }
    return r;
}

static M4 m4_translate(const float t[3])                         /* matrix_transform.inl:40-49 */
{
    M4 m = m4_identity(), r = m;
    for (int k = 0; k < 4; ++k) r.v[12 + k] = ((m.v[k] * t[0] + m.v[4 + k] * t[1]) + m.v[8 + k] * t[2]) + m.v[12 + k];
    return r;
}

static M4 m4_rotate(float angle, const float axis_in[3])         /* matrix_transform.inl:52-85 */
{
    const float c = cosf(angle), s = sinf(angle);
    const float len2 = (axis_in[0] * axis_in[0] + axis_in[1] * axis_in[1]) + axis_in[2] * axis_in[2];
    const float il = 1.0f / sqrtf(len2);
    const float ax[3] = {axis_in[0] * il, axis_in[1] * il, axis_in[2] * il};
    const float t[3] = {(1.0f - c) * ax[0], (1.0f - c) * ax[1], (1.0f - c) * ax[2]};
    float R[3][3];
    R[0][0] = c + t[0] * ax[0];
    R[0][1] = 0 + t[0] * ax[1] + s * ax[2];
    R[0][2] = 0 + t[0] * ax[2] - s * ax[1];
    R[1][0] = 0 + t[1] * ax[0] - s * ax[2];
    R[1][1] = c + t[1] * ax[1];
    R[1][2] = 0 + t[1] * ax[2] + s * ax[0];
    R[2][0] = 0 + t[2] * ax[0] + s * ax[1];
    R[2][1] = 0 + t[2] * ax[1] - s * ax[0];
    R[2][2] = c + t[2] * ax[2];
    M4 m = m4_identity(), r;
    for (int col = 0; col < 3; ++col)
        for (int k = 0; k < 4; ++k)
            r.v[4 * col + k] = (m.v[k] * R[col][0] + m.v[4 + k] * R[col][1]) + m.v[8 + k] * R[col][2];
    for (int k = 0; k < 4; ++k) r.v[12 + k] = m.v[12 + k];
    return r;
}

#define E(c, r) (M->v[4 * (c) + (r)])
static M4 m4_inverse(const M4* M)                                /* type_mat4x4.inl:37-92 */
{
    const float C00 = E(2, 2) * E(3, 3) - E(3, 2) * E(2, 3), C02 = E(1, 2) * E(3, 3) - E(3, 2) * E(1, 3), C03 = E(1, 2) * E(2, 3) - E(2, 2) * E(1, 3);
    const float C04 = E(2, 1) * E(3, 3) - E(3, 1) * E(2, 3), C06 = E(1, 1) * E(3, 3) - E(3, 1) * E(1, 3), C07 = E(1, 1) * E(2, 3) - E(2, 1) * E(1, 3);
    const float C08 = E(2, 1) * E(3, 2) - E(3, 1) * E(2, 2), C10 = E(1, 1) * E(3, 2) - E(3, 1) * E(1, 2), C11 = E(1, 1) * E(2, 2) - E(2, 1) * E(1, 2);
    const float C12 = E(2, 0) * E(3, 3) - E(3, 0) * E(2, 3), C14 = E(1, 0) * E(3, 3) - E(3, 0) * E(1, 3), C15 = E(1, 0) * E(2, 3) - E(2, 0) * E(1, 3);
    const float C16 = E(2, 0) * E(3, 2) - E(3, 0) * E(2, 2), C18 = E(1, 0) * E(3, 2) - E(3, 0) * E(1, 2), C19 = E(1, 0) * E(2, 2) - E(2, 0) * E(1, 2);
    const float C20 = E(2, 0) * E(3, 1) - E(3, 0) * E(2, 1), C22 = E(1, 0) * E(3, 1) - E(3, 0) * E(1, 1), C23 = E(1, 0) * E(2, 1) - E(2, 0) * E(1, 1);
    const float F0[4] = {C00, C00, C02, C03}, F1[4] = {C04, C04, C06, C07}, F2[4] = {C08, C08, C10, C11};
    const float F3[4] = {C12, C12, C14, C15}, F4[4] = {C16, C16, C18, C19}, F5[4] = {C20, C20, C22, C23};
    const float V0[4] = {E(1, 0), E(0, 0), E(0, 0), E(0, 0)}, V1[4] = {E(1, 1), E(0, 1), E(0, 1), E(0, 1)};
    const float V2[4] = {E(1, 2), E(0, 2), E(0, 2), E(0, 2)}, V3[4] = {E(1, 3), E(0, 3), E(0, 3), E(0, 3)};
    const float SA[4] = {+1, -1, +1, -1}, SB[4] = {-1, +1, -1, +1};
    M4 inv;
    for (int k = 0; k < 4; ++k) {
        inv.v[0 + k] = ((V1[k] * F0[k] - V2[k] * F1[k]) + V3[k] * F2[k]) * SA[k];
        inv.v[4 + k] = ((V0[k] * F0[k] - V2[k] * F3[k]) + V3[k] * F4[k]) * SB[k];
        inv.v[8 + k] = ((V0[k] * F1[k] - V1[k] * F3[k]) + V3[k] * F5[k]) * SA[k];
        inv.v[12 + k] = ((V0[k] * F2[k] - V1[k] * F4[k]) + V2[k] * F5[k]) * SB[k];
    }
    const float d0 = E(0, 0) * inv.v[0], d1 = E(0, 1) * inv.v[4], d2 = E(0, 2) * inv.v[8], d3 = E(0, 3) * inv.v[12];
    const float ood = 1.0f / ((d0 + d1) + (d2 + d3));
    for (int k = 0; k < 16; ++k) inv.v[k] = inv.v[k] * ood;
    return inv;
}
#undef E

/* translation_matrix * rotate_matrix(Y) * scale_matrix and its glm::inverse, as every model of Scene.cpp:32-221 is composed */
void oracle_compose_trs(const float translate[3], float rotate_y_degrees, const float scale[3], float model_to_world[16], float world_to_model[16])
{
    const float Y[3] = {0.0f, 1.0f, 0.0f};
    const float radians = rotate_y_degrees * 0.01745329251994329576923690768489f;      /* glm::radians, func_trigonometric.inl:41-46 */
    const M4 T = m4_translate(translate), R = m4_rotate(radians, Y), S = m4_scale(scale);
    const M4 TR = m4_mul(&T, &R), M = m4_mul(&TR, &S);
    const M4 I = m4_inverse(&M);
    memcpy(model_to_world, M.v, sizeof M.v);
    memcpy(world_to_model, I.v, sizeof I.v);
}

/* ---- displaced icosphere: 20 * 4^level triangles, one vertex per face corner like the OBJ import ---- */

static uint32_t hash32(uint32_t a) { a ^= a >> 16; a *= 0x7feb352du; a ^= a >> 15; a *= 0x846ca68bu; a ^= a >> 16; return a; }

typedef struct { uint64_t* key; int* val; size_t cap; } EdgeMap;

static int edge_find(EdgeMap* m, uint64_t key, int** slot)
{
    size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & (m->cap - 1);
    while (m->key[h] != ~0ull && m->key[h] != key) h = (h + 1) & (m->cap - 1);
    *slot = &m->val[h];
    if (m->key[h] == key) return 1;
    m->key[h] = key;
    return 0;
}

int oracle_icosphere_triangles(int level) { int n = 20; for (int l = 0; l < level; ++l) n *= 4; return n; }

/* vertices: 3 * ntriangles OVertex records (positions and normals scaled by BASE_MODEL_SCALE = 1000 like the OBJ import, Config.h:17);
 * bb_min / bb_max: the mesh bounds as Scene::loadAndProcessMeshFile accumulates them.  Triangle t uses vertices 3t, 3t+1, 3t+2. */
int oracle_icosphere(int level, float radius, float displacement, unsigned seed, OVertex* vertices, float bb_min[3], float bb_max[3])
{
    if (level < 0 || level > 10) return -1;
    const float kBaseModelScale = 1000.0f;
    const double t = (1.0 + sqrt(5.0)) / 2.0;
    const double base[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t}, {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    const int faces[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4}, {11, 10, 2}, {10, 7, 6}, {7, 1, 8},
                              {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8}, {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
    const int ntri_final = oracle_icosphere_triangles(level);
    const size_t nv_max = (size_t)ntri_final / 2 + 2 + 12;           /* V = F / 2 + 2 on a sphere */
    double* pos = (double*)malloc(3 * nv_max * sizeof(double));
    int* tri = (int*)malloc(3 * (size_t)ntri_final * sizeof(int));
    int* next = (int*)malloc(3 * (size_t)ntri_final * sizeof(int));
    int nv = 0, ntri = 20;
    for (int i = 0; i < 12; ++i) {
        const double l = sqrt(base[i][0] * base[i][0] + base[i][1] * base[i][1] + base[i][2] * base[i][2]);
        pos[3 * nv] = base[i][0] / l; pos[3 * nv + 1] = base[i][1] / l; pos[3 * nv + 2] = base[i][2] / l; ++nv;
    }
    for (int f = 0; f < 20; ++f) { tri[3 * f] = faces[f][0]; tri[3 * f + 1] = faces[f][1]; tri[3 * f + 2] = faces[f][2]; }
    for (int l = 0; l < level; ++l) {
        EdgeMap map; map.cap = 1;
        while (map.cap < (size_t)ntri * 4) map.cap <<= 1;
        map.key = (uint64_t*)malloc(map.cap * sizeof(uint64_t)); map.val = (int*)malloc(map.cap * sizeof(int));
        memset(map.key, 0xff, map.cap * sizeof(uint64_t));
        int nn = 0;
        for (int f = 0; f < ntri; ++f) {
            const int a = tri[3 * f], b = tri[3 * f + 1], c = tri[3 * f + 2];
            const int e[3][2] = {{a, b}, {b, c}, {c, a}};
            int mid[3];
            for (int k = 0; k < 3; ++k) {
                const int x = e[k][0], y = e[k][1];
                const uint64_t key = x < y ? ((uint64_t)x << 32) | (uint32_t)y : ((uint64_t)y << 32) | (uint32_t)x;
                int* slot;
                if (!edge_find(&map, key, &slot)) {
                    const double m[3] = {pos[3 * x] + pos[3 * y], pos[3 * x + 1] + pos[3 * y + 1], pos[3 * x + 2] + pos[3 * y + 2]};
                    const double len = sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
                    pos[3 * nv] = m[0] / len; pos[3 * nv + 1] = m[1] / len; pos[3 * nv + 2] = m[2] / len;
                    *slot = nv++;
                }
                mid[k] = *slot;
            }
            const int ab = mid[0], bc = mid[1], ca = mid[2];
            const int sub[12] = {a, ab, ca, b, bc, ab, c, ca, bc, ab, bc, ca};
            memcpy(next + 3 * nn, sub, sizeof sub); nn += 4;
        }
        free(map.key); free(map.val);
        int* sw = tri; tri = next; next = sw; ntri = nn;
    }
    double* disp = (double*)malloc((size_t)nv * sizeof(double));
    const double ph = (seed % 1000) * 0.01;
    for (int i = 0; i < nv; ++i) {
        const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
        const double smooth = 0.5 * sin(5 * x + ph) * cos(4 * y - ph) + 0.3 * sin(9 * z + 2 * ph) * sin(7 * x) + 0.2 * cos(13 * y + 3 * z);
        const double rough = (hash32((uint32_t)i * 2654435761u + seed) & 0xffff) / 65535.0 - 0.5;
        disp[i] = 1.0 + displacement * (smooth + 0.1 * rough);
    }
    double* P = (double*)malloc(3 * (size_t)nv * sizeof(double));
    double* Nrm = (double*)calloc(3 * (size_t)nv, sizeof(double));
    for (int i = 0; i < nv; ++i) for (int k = 0; k < 3; ++k) P[3 * i + k] = pos[3 * i + k] * disp[i];
    for (int f = 0; f < ntri; ++f) {
        const double* a = &P[3 * tri[3 * f]]; const double* b = &P[3 * tri[3 * f + 1]]; const double* c = &P[3 * tri[3 * f + 2]];
        const double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        const double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) Nrm[3 * tri[3 * f + j] + k] += n[k];
    }
    const double unit = (double)radius / kBaseModelScale;            /* OBJ-space radius; the import scales by BASE_MODEL_SCALE */
    for (int k = 0; k < 3; ++k) { bb_min[k] = 9999999.0f; bb_max[k] = -9999990.0f; }
    for (int i = 0; i < 3 * ntri; ++i) {
        const int vi = tri[i];
        const double l = sqrt(Nrm[3 * vi] * Nrm[3 * vi] + Nrm[3 * vi + 1] * Nrm[3 * vi + 1] + Nrm[3 * vi + 2] * Nrm[3 * vi + 2]);
        OVertex pv; memset(&pv, 0, sizeof pv);
        for (int k = 0; k < 3; ++k) {
            pv.position[k] = (float)(P[3 * vi + k] * unit) * kBaseModelScale;
            pv.normal[k] = (float)(l > 0 ? Nrm[3 * vi + k] / l : pos[3 * vi + k]) * kBaseModelScale;
            if (pv.position[k] < bb_min[k]) bb_min[k] = pv.position[k];
            if (pv.position[k] > bb_max[k]) bb_max[k] = pv.position[k];
        }
        vertices[i] = pv;
    }
    free(pos); free(tri); free(next); free(disp); free(P); free(Nrm);
    return ntri;
}
