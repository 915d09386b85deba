// trace_bvh.cu - closest hit through a two-level BVH (PTAP_ACCEL_BVH).
//
// New design; the reference has no BVH.  Semantics = oracle tier R1: for every model, the reference's own per-model
// ray set-up (Renderer.cpp:381-384) and its tolerant Moller-Trumbore predicate (Renderer.cpp:174-215) applied to
// every triangle of the model's mesh, nearest model-space t wins (ties: lowest triangle index), converted to a world
// distance as Renderer.cpp:388-391, nearest world distance wins (ties: lowest model index).  The BVH only decides
// WHICH triangles get tested: node bounds are conservative for the predicate's tolerance band (bvh_build.cpp), the
// slab test is evaluated with outward rounding slack, and the triangle arithmetic is the un-contracted exact one,
// so the winner is bit-identical to brute force.
//
// Node = 64 B holding both children's boxes (4 x 16-byte loads); triangles in leaf order, 48 B each.
#include "kernels.cuh"

namespace ptap {

namespace {

constexpr int kStack = 64;

struct BvhRay {
    V3 o, d, inv;
    float best_t; int best_tri; float best_u, best_v;
};

// Renderer.cpp:174-215; tie rule of brute force in index order: strictly nearer, or equal t and lower global id
template <bool COUNT>
__device__ __forceinline__ void leafTriangle(const SceneDev& sc, BvhRay& r, int k, int4& cnt)
{
    const TriRec* __restrict__ tp = &sc.bvh_tris[k];
    const float4 a = ldg4(&tp->v0), b = ldg4(&tp->e1), c = ldg4(&tp->e2);
    if (COUNT) cnt.z++;
    const V3 v0 = v3(a), v0v1 = v3(b), v0v2 = v3(c);
    const V3 pvec = xcross(r.d, v0v2);
    const float det = xdot(v0v1, pvec);
    if (xabs(xsub(det, 0.0f)) < kEpsilon) return;
    const float invDet = xdiv(1.0f, det);
    const V3 tvec = xsub(r.o, v0);
    const float u = xmul(xdot(tvec, pvec), invDet);
    if (u < (0.0f - kEpsilon) || u > (1.0f + kEpsilon)) return;
    const V3 qvec = xcross(tvec, v0v1);
    const float v = xmul(xdot(r.d, qvec), invDet);
    if (v < (0.0f - kEpsilon) || xadd(u, v) > (1.0f + kEpsilon)) return;
    const float t = xmul(xdot(v0v2, qvec), invDet);
    if (t < (0.0f - kEpsilon)) return;
    if (t > r.best_t) return;
    const int id = __ldg(&sc.bvh_tri_id[k]);
    if (t < r.best_t || (r.best_tri >= 0 && id < r.best_tri)) { r.best_t = t; r.best_tri = id; r.best_u = u; r.best_v = v; }
}

// conservative slab test: (plane - o) * inv per plane (no cancellation-prone FMA form), interval test with relative slack
__device__ __forceinline__ bool slab(const V3& o, const V3& inv, float lox, float hix, float loy, float hiy, float loz, float hiz,
                                     float tmin_ray, float tmax_ray, float& tnear)
{
    const float x0 = (lox - o.x) * inv.x, x1 = (hix - o.x) * inv.x;
    const float y0 = (loy - o.y) * inv.y, y1 = (hiy - o.y) * inv.y;
    const float z0 = (loz - o.z) * inv.z, z1 = (hiz - o.z) * inv.z;
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin_ray));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax_ray));
    tnear = tn;
    return tn <= tf + (fabsf(tf) * 2e-6f + 1e-6f);
}

}  // namespace

template <bool UV, bool COUNT>
__global__ void __launch_bounds__(kTraceBlock)
k_trace_bvh(SceneDev sc, const float4* __restrict__ O, const float4* __restrict__ D, float4* __restrict__ hit,
            float2* __restrict__ uv, int4* __restrict__ counts, FrameState* st, int round, int n_fixed)
{
    const int n = n_fixed >= 0 ? n_fixed : st->n_active[round];
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_fixed < 0) st->rays_traced += (unsigned long long)n;
    unsigned long long tot_x = 0, tot_y = 0, tot_z = 0;
    int stack[kStack];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = O[i], d4 = D[i];
        const V3 bo = v3(o4), bd = v3(d4);
        float g_dist = kFloatMax, g_t = 0.0f, g_u = 0.0f, g_v = 0.0f, last_dist = kFloatMax;
        int g_model = -1, g_tri = -1;
        int4 cnt = make_int4(0, 0, 0, 0);
        BvhRay r;
        for (int im = 0; im < sc.nmodels; ++im) {
            const InstanceTrace* __restrict__ inst = &sc.inst[im];
            const int root = __float_as_int(__ldg(&inst->grid.z));
            if (root < 0) continue;
            const float4 w0 = ldg4(&inst->w2m[0]), w1 = ldg4(&inst->w2m[1]), w2 = ldg4(&inst->w2m[2]);
            r.o = xmat4(w0, w1, w2, bo, 1.0f);                                   // Renderer.cpp:381
            r.d = xnormalize(xmat4(w0, w1, w2, bd, 0.0f));                       // Renderer.cpp:382
            // traversal-only reciprocal: guarded against 0 (the exact predicate never uses it)
            const float ooeps = 1e-30f;
            r.inv = v3(1.0f / (fabsf(r.d.x) > ooeps ? r.d.x : copysignf(ooeps, r.d.x)),
                       1.0f / (fabsf(r.d.y) > ooeps ? r.d.y : copysignf(ooeps, r.d.y)),
                       1.0f / (fabsf(r.d.z) > ooeps ? r.d.z : copysignf(ooeps, r.d.z)));
            r.best_t = kFloatMax; r.best_tri = -1; r.best_u = 0.0f; r.best_v = 0.0f;
            last_dist = kFloatMax;
            const float tmin_ray = -(kEpsilon + 1e-4f);                          // the predicate accepts t >= -EPSILON

            int sp = 0;
            int node = root;
            for (;;) {
                // inner node: test both children, descend into the nearer, push the farther
                const BvhNode* __restrict__ np = &sc.nodes[node];
                const float4 xy0 = ldg4(&np->xy0), xy1 = ldg4(&np->xy1), z01 = ldg4(&np->z01);
                const int4 link = __ldg(&np->link);
                if (COUNT) cnt.x++;
                float tn0, tn1;
                const bool h0 = slab(r.o, r.inv, xy0.x, xy0.y, xy0.z, xy0.w, z01.x, z01.y, tmin_ray, r.best_t, tn0);
                const bool h1 = slab(r.o, r.inv, xy1.x, xy1.y, xy1.z, xy1.w, z01.z, z01.w, tmin_ray, r.best_t, tn1);
                int next = 0; bool have_next = false;
                if (h0 || h1) {
                    int c0 = link.x, c1 = link.y;
                    const bool two = h0 && h1;
                    if (two ? (tn1 < tn0) : h1) { const int t = c0; c0 = c1; c1 = t; }   // c0 = nearer (or the only) child
                    if (c0 < 0) {                                                     // leaf: intersect now
                        const int code = ~c0; const int k0 = code >> 3, kc = (code & 7) + 1;
                        for (int k = 0; k < kc; ++k) leafTriangle<COUNT>(sc, r, k0 + k, cnt);
                    } else { next = c0; have_next = true; }
                    if (two) {
                        if (c1 < 0) {
                            const int code = ~c1; const int k0 = code >> 3, kc = (code & 7) + 1;
                            for (int k = 0; k < kc; ++k) leafTriangle<COUNT>(sc, r, k0 + k, cnt);
                        } else if (have_next) {
                            if (sp < kStack) stack[sp++] = c1;
                        } else { next = c1; have_next = true; }
                    }
                }
                if (have_next) { node = next; continue; }
                if (sp == 0) break;
                node = stack[--sp];
            }

            if (r.best_tri >= 0) {
                const V3 nd = xnormalize(r.d);                                   // Renderer.cpp:388
                const V3 pm = xadd(r.o, xscale(nd, r.best_t));                   // Renderer.cpp:389
                const V3 pw = xmat4(ldg4(&inst->m2w[0]), ldg4(&inst->m2w[1]), ldg4(&inst->m2w[2]), pm, 1.0f);   // :390
                const float dist = xlength(xsub(pw, bo));                        // Renderer.cpp:391
                last_dist = dist;
                if (g_dist > dist) {                                             // Renderer.cpp:393-398
                    g_dist = dist; g_model = im; g_tri = r.best_tri; g_t = r.best_t; g_u = r.best_u; g_v = r.best_v;
                }
            }
        }
        const bool found = g_dist < kFloatMax;
        hit[i] = make_float4(found ? g_dist : last_dist, __int_as_float(found ? g_tri : -1), __int_as_float(found ? g_model : -1), g_t);
        if (UV && uv) uv[i] = make_float2(g_u, g_v);
        if (COUNT) { if (counts) counts[i] = cnt; tot_x += cnt.x; tot_y += cnt.y; tot_z += cnt.z; }
    }
    if (COUNT) {        // counting build: per-warp totals into the frame state (never used for timing)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            tot_x += __shfl_xor_sync(0xffffffffu, tot_x, d); tot_y += __shfl_xor_sync(0xffffffffu, tot_y, d); tot_z += __shfl_xor_sync(0xffffffffu, tot_z, d);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&st->count_nodes, tot_x); atomicAdd(&st->count_refs, tot_y); atomicAdd(&st->count_tris, tot_z);
        }
    }
}

void launchTraceBvh(const SceneDev& sc, const float4* O, const float4* D, float4* hit, float2* uv, int4* counts, bool count_totals,
                     FrameState* st, int round, int n_fixed, int grid, cudaStream_t stream)
{
    if (counts || count_totals) k_trace_bvh<true, true><<<grid, kTraceBlock, 0, stream>>>(sc, O, D, hit, uv, counts, st, round, n_fixed);
    else if (uv) k_trace_bvh<true, false><<<grid, kTraceBlock, 0, stream>>>(sc, O, D, hit, uv, counts, st, round, n_fixed);
    else k_trace_bvh<false, false><<<grid, kTraceBlock, 0, stream>>>(sc, O, D, hit, uv, counts, st, round, n_fixed);
}

int traceBvhOccupancy()
{
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_bvh<false, false>, kTraceBlock, 0);
    return nb;
}

}  // namespace ptap
