// trace_grid.cu - closest hit through the reference's per-mesh uniform grid (PTAP_ACCEL_GRID_COMPAT).
//
// Replaces computeRaySceneIntersectionKernel and its helpers (Renderer.cpp:150-409) and is bit-compatible
// with them (oracle tier R0): same model loop, same slab test, same 3D-DDA with its early exit, same
// tolerant Moller-Trumbore predicate, every hit-deciding operation un-contracted (exact_math.cuh).
// What changes is everything around the arithmetic: the ray lives in registers instead of being written back
// to global memory per model (Renderer.cpp:381-384), the scene is SoA/AoSoA fetched as 16-byte vectors
// (TriRec, int2 cells) instead of 112 B of AoS gathers per test through two pointer hops, the flat normal and
// the 3x3 inverse-transpose are not evaluated in the loop at all (the kernel returns ids; the shade kernel
// looks the precomputed normal up), and the active count is read from device memory so that no host
// round trip separates the bounces.
#include "kernels.cuh"

namespace ptap {

namespace {

struct GridRay {
    V3 o, d, inv;
    float best_t;     // hit_info->impact_distance inside the current model
    int best_tri;
    float best_u, best_v;
};

// Renderer.cpp:174-215 without the normal (looked up later from the id)
template <bool COUNT>
__device__ __forceinline__ bool rayTriangle(const TriRec* __restrict__ tris, GridRay& r, int itri, int4& cnt)
{
    const float4 a = ldg4(&tris[itri].v0), b = ldg4(&tris[itri].e1), c = ldg4(&tris[itri].e2);
    if (COUNT) cnt.z++;
    const V3 v0 = v3(a), v0v1 = v3(b), v0v2 = v3(c);
    const V3 pvec = xcross(r.d, v0v2);
    const float det = xdot(v0v1, pvec);
    if (xabs(xsub(det, 0.0f)) < kEpsilon) return false;
    const float invDet = xdiv(1.0f, det);
    const V3 tvec = xsub(r.o, v0);
    const float u = xmul(xdot(tvec, pvec), invDet);
    if (u < (0.0f - kEpsilon) || u > (1.0f + kEpsilon)) return false;
    const V3 qvec = xcross(tvec, v0v1);
    const float v = xmul(xdot(r.d, qvec), invDet);
    if (v < (0.0f - kEpsilon) || xadd(u, v) > (1.0f + kEpsilon)) return false;
    const float t = xmul(xdot(v0v2, qvec), invDet);
    if (t < (0.0f - kEpsilon)) return false;
    if (r.best_t > t) { r.best_t = t; r.best_tri = itri; r.best_u = u; r.best_v = v; }
    return true;
}

// Renderer.cpp:238-360 (+ the `return false` of the bbox-miss path)
template <bool COUNT>
__device__ __forceinline__ bool rayGrid(const SceneDev& sc, const float4 bbmin_wx, const float4 bbmax_wy, const float4 gridrec,
                                        GridRay& r, int4& cnt)
{
    const V3 mn = v3(bbmin_wx), mx = v3(bbmax_wy);
    const float wx = bbmin_wx.w, wy = bbmax_wy.w, wz = gridrec.x;
    const int vox0 = __float_as_int(gridrec.y);
    const int GX = sc.gx, GY = sc.gy, GZ = sc.gz;

    // Renderer.cpp:150-170
    const float t1 = r.d.x == 0.0f ? kFloatMin : xmul(xsub(mn.x, r.o.x), r.inv.x);
    const float t2 = r.d.x == 0.0f ? kFloatMax : xmul(xsub(mx.x, r.o.x), r.inv.x);
    const float t3 = r.d.y == 0.0f ? kFloatMin : xmul(xsub(mn.y, r.o.y), r.inv.y);
    const float t4 = r.d.y == 0.0f ? kFloatMax : xmul(xsub(mx.y, r.o.y), r.inv.y);
    const float t5 = r.d.z == 0.0f ? kFloatMin : xmul(xsub(mn.z, r.o.z), r.inv.z);
    const float t6 = r.d.z == 0.0f ? kFloatMax : xmul(xsub(mx.z, r.o.z), r.inv.z);
    const float tmin = max_std(max_std(min_std(t1, t2), min_std(t3, t4)), min_std(t5, t6));
    const float tmax = min_std(min_std(max_std(t1, t2), max_std(t3, t4)), max_std(t5, t6));
    if (tmax < 0 || tmin > tmax) return false;

    const V3 p = xadd(r.o, xscale(r.d, tmin));
    if (xsub(p.x, mn.x) < -kEpsilon || xsub(p.y, mn.y) < -kEpsilon || xsub(p.z, mn.z) < -kEpsilon) return false;

    int ix = f2i_x86(xdiv(xabs(xadd(xsub(p.x, mn.x), kEpsilon)), wx));
    int iy = f2i_x86(xdiv(xabs(xadd(xsub(p.y, mn.y), kEpsilon)), wy));
    int iz = f2i_x86(xdiv(xabs(xadd(xsub(p.z, mn.z), kEpsilon)), wz));
    ix = min(max(ix, 0), GX - 1); iy = min(max(iy, 0), GY - 1); iz = min(max(iz, 0), GZ - 1);

    float tmx = kFloatMax, tmy = kFloatMax, tmz = kFloatMax, dx = kFloatMax, dy = kFloatMax, dz = kFloatMax;
    const int sx = r.d.x > 0.0f ? 1 : -1, sy = r.d.y > 0.0f ? 1 : -1, sz = r.d.z > 0.0f ? 1 : -1;
    const int ox = r.d.x > 0.0f ? GX : -1, oy = r.d.y > 0.0f ? GY : -1, oz = r.d.z > 0.0f ? GZ : -1;
    if (r.d.x != 0) {
        const int nx = r.d.x > 0.0f ? ix + 1 : ix;
        dx = xabs(xmul(wx, r.inv.x));
        tmx = xmul(xsub(xadd(mn.x, xmul((float)nx, wx)), p.x), r.inv.x);
    }
    if (r.d.y != 0) {
        const int ny = r.d.y > 0.0f ? iy + 1 : iy;
        dy = xabs(xmul(wy, r.inv.y));
        tmy = xmul(xsub(xadd(mn.y, xmul((float)ny, wy)), p.y), r.inv.y);
    }
    if (r.d.z != 0) {
        const int nz = r.d.z > 0.0f ? iz + 1 : iz;
        dz = xabs(xmul(wz, r.inv.z));
        tmz = xmul(xsub(xadd(mn.z, xmul((float)nz, wz)), p.z), r.inv.z);
    }

    int cx = 0, cy = 0, cz = 0;
    bool is_intersect = false;
    const int strideY = GX, strideZ = GX * GY;
    for (;;) {
        const int2 cell = __ldg(&sc.cells[vox0 + ix + iy * strideY + iz * strideZ]);
        if (COUNT) cnt.x++;
        bool any = false;                                   // Renderer.cpp:217-236
        for (int i = cell.x; i < cell.y; ++i) {
            const int itri = __ldg(&sc.refs[i]);
            if (COUNT) cnt.y++;
            if (rayTriangle<COUNT>(sc.tris, r, itri, cnt)) any = true;
        }
        if (any) { cx = ix; cy = iy; cz = iz; is_intersect = true; }
        if (is_intersect && (abs(cx - ix) > 2 || abs(cy - iy) > 2 || abs(cz - iz) > 2)) return true;
        if (tmx < tmy && tmx < tmz) {
            ix += sx;
            if (ix == ox || tmx >= kFloatMax) return is_intersect;
            tmx = xadd(tmx, dx);
        } else if (tmy < tmz) {
            iy += sy;
            if (iy == oy || tmy >= kFloatMax) return is_intersect;
            tmy = xadd(tmy, dy);
        } else {
            iz += sz;
            if (iz == oz || tmz >= kFloatMax) return is_intersect;
            tmz = xadd(tmz, dz);
        }
    }
}

}  // namespace

// One thread per active ray slot; persistent grid-stride loop (grid = SMs x resident CTAs).
template <bool UV, bool COUNT>
__global__ void __launch_bounds__(kTraceBlock)
k_trace_grid(SceneDev sc, const float4* __restrict__ O, const float4* __restrict__ D, float4* __restrict__ hit,
             float2* __restrict__ uv, int4* __restrict__ counts, FrameState* st, int round, int n_fixed)
{
    const int n = n_fixed >= 0 ? n_fixed : st->n_active[round];
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_fixed < 0) st->rays_traced += (unsigned long long)n;
    unsigned long long tot_x = 0, tot_y = 0, tot_z = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = O[i], d4 = D[i];
        const V3 bo = v3(o4), bd = v3(d4);
        float g_dist = kFloatMax, g_t = 0.0f, g_u = 0.0f, g_v = 0.0f;
        int g_model = -1, g_tri = -1;
        float last_dist = kFloatMax;                        // what the slot's impact_distance holds after the loop
        int4 cnt = make_int4(0, 0, 0, 0);
        GridRay r;
        for (int im = 0; im < sc.nmodels; ++im) {
            const InstanceTrace* __restrict__ inst = &sc.inst[im];
            const float4 w0 = ldg4(&inst->w2m[0]), w1 = ldg4(&inst->w2m[1]), w2 = ldg4(&inst->w2m[2]);
            r.o = xmat4(w0, w1, w2, bo, 1.0f);                                   // Renderer.cpp:381
            r.d = xnormalize(xmat4(w0, w1, w2, bd, 0.0f));                       // Renderer.cpp:382
            r.inv = v3(xdiv(1.0f, r.d.x), xdiv(1.0f, r.d.y), xdiv(1.0f, r.d.z)); // Renderer.cpp:383
            r.best_t = kFloatMax; r.best_tri = -1; r.best_u = 0.0f; r.best_v = 0.0f;   // Renderer.cpp:384
            last_dist = kFloatMax;
            if (rayGrid<COUNT>(sc, ldg4(&inst->bb_min), ldg4(&inst->bb_max), ldg4(&inst->grid), r, cnt)) {
                const V3 nd = xnormalize(r.d);                                   // Renderer.cpp:388
                const V3 pm = xadd(r.o, xscale(nd, r.best_t));                   // Renderer.cpp:389
                const V3 pw = xmat4(ldg4(&inst->m2w[0]), ldg4(&inst->m2w[1]), ldg4(&inst->m2w[2]), pm, 1.0f);   // :390
                const float dist = xlength(xsub(pw, bo));                        // Renderer.cpp:391
                last_dist = dist;
                if (g_dist > dist) {                                             // Renderer.cpp:393-398
                    g_dist = dist; g_model = im; g_tri = r.best_tri; g_t = r.best_t; g_u = r.best_u; g_v = r.best_v;
                }
            } else {
                last_dist = r.best_t;
            }
        }
        const bool found = g_dist < kFloatMax;                                   // Renderer.cpp:402-408
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
            atomicAdd(&st->count_cells, tot_x); atomicAdd(&st->count_refs, tot_y); atomicAdd(&st->count_tris, tot_z);
        }
    }
}

void launchTraceGrid(const SceneDev& sc, const float4* O, const float4* D, float4* hit, float2* uv, int4* counts, bool count_totals,
                     FrameState* st, int round, int n_fixed, int grid, cudaStream_t stream)
{
    if (counts || count_totals) k_trace_grid<true, true><<<grid, kTraceBlock, 0, stream>>>(sc, O, D, hit, uv, counts, st, round, n_fixed);
    else if (uv) k_trace_grid<true, false><<<grid, kTraceBlock, 0, stream>>>(sc, O, D, hit, uv, counts, st, round, n_fixed);
    else k_trace_grid<false, false><<<grid, kTraceBlock, 0, stream>>>(sc, O, D, hit, uv, counts, st, round, n_fixed);
}

int traceGridOccupancy()
{
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_grid<false, false>, kTraceBlock, 0);
    return nb;
}

}  // namespace ptap
