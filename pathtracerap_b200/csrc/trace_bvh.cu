// trace_bvh.cu - closest hit through a two-level wide BVH with node-local child boxes (PTAP_ACCEL_BVH / PTAP_ACCEL_BVH_DEVICE).
//
// New design; the reference has no BVH.  Semantics = oracle tier R1: for every model, the reference's own per-model
// ray set-up (Renderer.cpp:381-384) and its tolerant Moller-Trumbore predicate (Renderer.cpp:174-215) applied to
// every triangle of the model's mesh, nearest model-space t wins (ties: lowest triangle index), converted to a world
// distance as Renderer.cpp:388-391, nearest world distance wins (ties: lowest model index).  The BVH only decides
// WHICH triangles get tested: node bounds are conservative for the predicate's tolerance band and rounded outward when
// compressed (bvh_build.cpp), the slab test carries relative slack, pruning bounds carry slack, and the triangle
// arithmetic is the un-contracted exact one, so the winner is bit-identical to brute force.
//
// Node = 128 B = one L1 line (device_types.h: BvhNode): child boxes as offsets in a node-local frame, children in ordered slots.
// Default: four children, binary32 offsets (PTAP_BVH_WIDTH=8 builds the eight-children / IEEE-half variant, measured slower on B200).
// What the format buys, per node visit:
//  * one FMA per plane.  t = offset * inv + (p - o) * inv: the (p - o) term is evaluated once per node RELATIVE TO THE NODE, so the
//    cancellation error of the FMA form is bounded by the node's own extent (the global o * inv form of the textbook kernels is
//    unbounded for axis-parallel rays, which is why round 1's kernel spent a subtract and a multiply per plane);
//  * no sorting: the hit children are visited in ascending (slot ^ key), front to back by construction of the slots, so the hit mask
//    is permuted by conditional bit swaps instead of a comparator network on entry distances;
//  * near / far planes are chosen by ADDRESS (the lower / upper planes of an axis are 16 bytes apart), not by min / max per child;
//  * one 8-byte stack entry per node (child base, pending-children mask) instead of one entry per hit child.
//
// Execution model (DESIGN.md "Closest hit"): a persistent grid of warps, each lane owning one ray at a time, scheduled as a warp-wide
// state machine.  A lane holds a NODE GROUP (child base + mask of hit inner children still to visit) and a TRIANGLE GROUP (leaf base +
// mask of triangles of hit leaf children still to test; at the TLAS level: instances still to enter).  Every round
//  (1) lanes with nothing pending pop their stack; lanes whose triangle group is empty advance by one node (pop the nearest pending
//      child of the node group, push the rest of the group, fetch the child, test its eight boxes);
//  (2) one warp reduction counts the lanes waiting in each other state;
//  (3) a state's step runs only when enough lanes wait for it (or when it is at least as popular as descending): one Moller-Trumbore
//      test per lane, the instance entry (the reference's world->model ray set-up), the instance exit (model t -> world distance,
//      nearest-model bookkeeping), retire + refill from the work-stealing cursor.
// Cross-instance pruning, deferred exact distances and work stealing are those of the previous kernel (see the steps below).
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace ptap {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr unsigned kExitMark = 0x80000000u, kDoneMark = 0xC0000000u;     // stack entry .y of the instance marker / the bottom of the stack
// lane flags
constexpr unsigned kOctMask = 7u, kInBlas = 8u, kExit = 16u, kDone = 32u;

// Renderer.cpp:174-215; tie rule of brute force in index order: strictly nearer, or equal t and lower global id.
// `tmax` is the model-space bound of the traversal AND the best t so far (they are the same number once best_tri >= 0).
template <bool UV, bool COUNT>
__device__ __forceinline__ void leafTriangle(const SceneDev& sc, const V3& o, const V3& d, float& tmax, int& best_tri, float& best_u, float& best_v,
                                             int k, int4& cnt)
{
    const F8 ta = ldg8(&sc.bvh_tris[k]), tb = ldg8(reinterpret_cast<const char*>(&sc.bvh_tris[k]) + 32);
    if (COUNT) cnt.z++;
    // Straight-line evaluation: the reference's early returns (Renderer.cpp:188-201) have no side effects, so testing all of its
    // rejection conditions at the end gives the same verdict, and the handful of lanes in a triangle step do not diverge further.
    const V3 v0 = v3(ta.v[0], ta.v[1], ta.v[2]), v0v1 = v3(ta.v[3], ta.v[4], ta.v[5]), v0v2 = v3(ta.v[6], ta.v[7], tb.v[0]);
    const V3 pvec = xcross(d, v0v2);
    const float det = xdot(v0v1, pvec);
    const float invDet = xdiv(1.0f, det);
    const V3 tvec = xsub(o, v0);
    const float u = xmul(xdot(tvec, pvec), invDet);
    const V3 qvec = xcross(tvec, v0v1);
    const float v = xmul(xdot(d, qvec), invDet);
    const float t = xmul(xdot(v0v2, qvec), invDet);
    const bool reject = (xabs(xsub(det, 0.0f)) < kEpsilon) | (u < (0.0f - kEpsilon)) | (u > (1.0f + kEpsilon)) |
                        (v < (0.0f - kEpsilon)) | (xadd(u, v) > (1.0f + kEpsilon)) | (t < (0.0f - kEpsilon)) | (t > tmax);
    if (reject || !(t <= tmax)) return;                        // the second test only catches NaN (a degenerate det passes none of the above as true)
    const int id = __float_as_int(tb.v[1]);
    if (t < tmax || (best_tri >= 0 && id < best_tri)) { tmax = t; best_tri = id; if (UV) { best_u = u; best_v = v; } }
}

// traversal-only reciprocal: guarded against 0 (the exact predicate never uses it)
__device__ __forceinline__ float safeInv(float d)
{
    const float ooeps = 1e-30f;
    return __fdividef(1.0f, fabsf(d) > ooeps ? d : copysignf(ooeps, d));     // 1-ulp reciprocal: inside the slab test's 2e-6 slack
}

// direction-sign bits, taken from the SIGN BIT (not d < 0) so that they agree with safeInv for -0.0f: the octant decides which half of a
// plane pair is the near one, and a near / far swap against the sign of the reciprocal would turn every box into a miss
__device__ __forceinline__ unsigned octOf(const V3& d)
{
    return (__float_as_uint(d.x) >> 31) | ((__float_as_uint(d.y) >> 31) << 1) | ((__float_as_uint(d.z) >> 31) << 2);
}

// byte offsets (x | y << 8 | z << 16) of the near planes of the three axes inside a node: lower planes at 32 / 64 / 96, upper planes 16 further
__device__ __forceinline__ unsigned nearOffsets(unsigned oct)
{
    return (32u + ((oct & 1u) << 4)) | ((64u + ((oct & 2u) << 3)) << 8) | ((96u + ((oct & 4u) << 2)) << 16);
}

// TLAS pruning bound once some instance reported world distance g_dist: g_dist may be the approximation t * |d_w| / |W3 d_w| of the exact
// distance (off by at most g_dist * tie + cb, tie < prune - 1), so the bound carries the same absolute slack cb as the instance-entry bound
__device__ __forceinline__ float worldBound(float g_dist, float prune, float cb)
{
    return g_dist < kFloatMax ? g_dist * prune + cb + 1e-3f : 3.0e38f;
}

__device__ __forceinline__ uint4 ldg4u(const void* p)
{
    uint4 r;
    asm("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float2 halves(unsigned w)
{
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
}

// One child: entry / exit parameters from the six planes (near / far already selected by the ray's octant), interval test with slack.
__device__ __forceinline__ bool childHit(float nx, float ny, float nz, float fx, float fy, float fz, float sx, float sy, float sz,
                                         float cx, float cy, float cz, float tmin, float tmax)
{
    const float tn = fmaxf(fmaxf(fmaxf(__fmaf_rn(nx, sx, cx), __fmaf_rn(ny, sy, cy)), __fmaf_rn(nz, sz, cz)), tmin);
    const float tf = fminf(fminf(fminf(__fmaf_rn(fx, sx, cx), __fmaf_rn(fy, sy, cy)), __fmaf_rn(fz, sz, cz)), tmax);
    return tn <= tf + __fmaf_rn(fabsf(tf), 2e-6f, 1e-6f);
}

// new position k holds old position k ^ key (mask of kBvhWidth bits)
__device__ __forceinline__ unsigned permuteByKey(unsigned h, unsigned key)
{
    const unsigned a = ((h & 0x55u) << 1) | ((h >> 1) & 0x55u);
    h = (key & 1u) ? a : h;
    const unsigned b = ((h & 0x33u) << 2) | ((h >> 2) & 0x33u);
    h = (key & 2u) ? b : h;
    if (kBvhWidth == 8) {
        const unsigned c = ((h & 0x0fu) << 4) | ((h >> 4) & 0x0fu);
        h = (key & 4u) ? c : h;
    }
    return h;
}

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)          // FFMA2 (sm_100): two binary32 FMAs per issue slot
{
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ bool intervalHit(float nx, float ny, float nz, float fx, float fy, float fz, float tmin, float tmax)
{
    const float tn = fmaxf(fmaxf(fmaxf(nx, ny), nz), tmin);
    const float tf = fminf(fminf(fminf(fx, fy), fz), tmax);
    return tn <= tf + __fmaf_rn(fabsf(tf), 2e-6f, 1e-6f);
}

}  // namespace

template <bool UV, bool COUNT>
__global__ void __launch_bounds__(kTraceBlock)
k_trace_bvh(SceneDev sc, const float4* __restrict__ O, const float4* __restrict__ D, float4* __restrict__ hit,
            float2* __restrict__ uv, int4* __restrict__ counts, FrameState* st, int round, int n_fixed, unsigned long long* __restrict__ stamp)
{
    const int n = n_fixed >= 0 ? n_fixed : st->n_active[round];
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_fixed < 0) st->rays_traced += (unsigned long long)n;
    if (stamp && threadIdx.x == 0) atomicMin(stamp, globalTimerNs());
    unsigned int* cursor = &st->fetch[round];
    const int lane = threadIdx.x & 31;
    unsigned long long tot_x = 0, tot_z = 0;
    int4 cnt = make_int4(0, 0, 0, 0);

    if (sc.tlas_root < 0) {                            // no instance has triangles: every ray misses
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
            hit[k] = make_float4(kFloatMax, __int_as_float(-1), __int_as_float(-1), 0.0f);
            if (UV && uv) uv[k] = make_float2(0.0f, 0.0f);
            if (COUNT && counts) counts[k] = cnt;
        }
        return;
    }

    uint2 stack[kBvhStack];
    int sp = 0, i = -1;
    unsigned flags = kDone;                         // octant of the current level | kInBlas | kExit | kDone
    unsigned near_off = 0u;                         // byte offsets of the near planes within a node, one byte per axis (from the octant)
    int g_base = 0; unsigned g_bits = 0u;           // node group: child base, pending hit inner children (bits 0-7, traversal order) | inner mask << 8
    int t_base = 0; unsigned t_mask = 0u, t_lm = 0u;   // triangle group: leaf base, pending triangles, the node's leaf mask
    int exit_im = 0;
    V3 bo = v3(0, 0, 0), bd = v3(0, 0, 0);          // the ray as stored (Ray::base, Primitive.h:160-164)
    V3 ro = v3(0, 0, 0), rd = v3(0, 0, 1), rinv = v3(0, 0, 0);   // the ray of the current level: world, or model space of the entered instance
    float tmin = 0.0f, tmax = 0.0f;
    int best_tri = -1; float best_u = 0.0f, best_v = 0.0f;
    float g_dist = kFloatMax, g_t = 0.0f, g_u = 0.0f, g_v = 0.0f;
    int g_model = -1, g_tri = -1;
    int w_next = 0, w_end = 0;
    bool exhausted = false;
    const int vote_tri = sc.vote_tri, vote_inst = sc.vote_inst, vote_refill = sc.vote_refill;

    for (;;) {
        // ---- (1a) nothing pending: pop.  A node group comes back as it was pushed; a marker ends the instance or the ray.
        if (!(flags & (kExit | kDone)) && t_mask == 0u && (g_bits & 0xffu) == 0u) {
            const uint2 e = stack[--sp];
            if (e.y & kExitMark) { flags |= (e.y == kDoneMark) ? kDone : kExit; exit_im = (int)e.x; }
            else { g_base = (int)e.x; g_bits = e.y; }
        }
        // ---- (1b) inner step: the nearest pending child of the node group
        if (!(flags & (kExit | kDone)) && t_mask == 0u && (g_bits & 0xffu) != 0u) {
            unsigned h = g_bits & 0xffu;
            const unsigned k = (unsigned)__ffs((int)h) - 1u, c = k ^ ((g_bits >> 16) & 7u);
            const int node = g_base + __popc((g_bits >> 8) & ((1u << c) - 1u));
            h &= h - 1u;
            if (h) stack[sp++] = make_uint2((unsigned)g_base, (g_bits & 0xffff00u) | h);
            const char* __restrict__ np = reinterpret_cast<const char*>(&sc.nodes[node]);
            const F8 hd = ldg8(np);                                              // origin, scale / order, bases, masks
            // near / far planes by address: the lower and upper planes of an axis are 16 bytes apart, `near_off` holds the three byte
            // offsets of the near ones (set when the level's ray is set up), the far ones are at offset ^ 16
            const unsigned nox = near_off & 0xffu, noy = (near_off >> 8) & 0xffu, noz = near_off >> 16;
            const uint4 NX = ldg4u(np + nox), FX = ldg4u(np + (nox ^ 16u));
            const uint4 NY = ldg4u(np + noy), FY = ldg4u(np + (noy ^ 16u));
            const uint4 NZ = ldg4u(np + noz), FZ = ldg4u(np + (noz ^ 16u));
            if (COUNT) cnt.x++;
            const float cx = (hd.v[0] - ro.x) * rinv.x, cy = (hd.v[1] - ro.y) * rinv.y, cz = (hd.v[2] - ro.z) * rinv.z;
            const unsigned lm = __float_as_uint(hd.v[6]), im = __float_as_uint(hd.v[7]);
            const unsigned nxw[4] = {NX.x, NX.y, NX.z, NX.w}, fxw[4] = {FX.x, FX.y, FX.z, FX.w};
            const unsigned nyw[4] = {NY.x, NY.y, NY.z, NY.w}, fyw[4] = {FY.x, FY.y, FY.z, FY.w};
            const unsigned nzw[4] = {NZ.x, NZ.y, NZ.z, NZ.w}, fzw[4] = {FZ.x, FZ.y, FZ.z, FZ.w};
            unsigned hits = 0u, tm = 0u, key;
            if (kBvhWidth == 8) {
                const float sx = hd.v[3] * rinv.x, sy = hd.v[3] * rinv.y, sz = hd.v[3] * rinv.z;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float2 ax = halves(nxw[w]), bx = halves(fxw[w]), ay = halves(nyw[w]), by = halves(fyw[w]), az = halves(nzw[w]), bz = halves(fzw[w]);
                    if (childHit(ax.x, ay.x, az.x, bx.x, by.x, bz.x, sx, sy, sz, cx, cy, cz, tmin, tmax)) { hits |= 1u << (2 * w); tm |= lm & (0xfu << (8 * w)); }
                    if (childHit(ax.y, ay.y, az.y, bx.y, by.y, bz.y, sx, sy, sz, cx, cy, cz, tmin, tmax)) { hits |= 2u << (2 * w); tm |= lm & (0xf0u << (8 * w)); }
                }
                key = flags & kOctMask;
            } else {
                // four children, two per FFMA2: t = offset * inv + (p - o) * inv
                const float2 ix = make_float2(rinv.x, rinv.x), iy = make_float2(rinv.y, rinv.y), iz = make_float2(rinv.z, rinv.z);
                const float2 ccx = make_float2(cx, cx), ccy = make_float2(cy, cy), ccz = make_float2(cz, cz);
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const float2 tnx = fma2(make_float2(__uint_as_float(nxw[2 * w]), __uint_as_float(nxw[2 * w + 1])), ix, ccx);
                    const float2 tny = fma2(make_float2(__uint_as_float(nyw[2 * w]), __uint_as_float(nyw[2 * w + 1])), iy, ccy);
                    const float2 tnz = fma2(make_float2(__uint_as_float(nzw[2 * w]), __uint_as_float(nzw[2 * w + 1])), iz, ccz);
                    const float2 tfx = fma2(make_float2(__uint_as_float(fxw[2 * w]), __uint_as_float(fxw[2 * w + 1])), ix, ccx);
                    const float2 tfy = fma2(make_float2(__uint_as_float(fyw[2 * w]), __uint_as_float(fyw[2 * w + 1])), iy, ccy);
                    const float2 tfz = fma2(make_float2(__uint_as_float(fzw[2 * w]), __uint_as_float(fzw[2 * w + 1])), iz, ccz);
                    if (intervalHit(tnx.x, tny.x, tnz.x, tfx.x, tfy.x, tfz.x, tmin, tmax)) hits |= 1u << (2 * w);
                    if (intervalHit(tnx.y, tny.y, tnz.y, tfx.y, tfy.y, tfz.y, tmin, tmax)) hits |= 2u << (2 * w);
                }
                tm = lm & (((hits * 0x249u) & 0x1111u) * 15u);                   // hit bit c -> nibble c: the triangles of the hit leaf children
                key = (__float_as_uint(hd.v[3]) >> (2u * (flags & kOctMask))) & 3u;   // the node's slot key for this sign octant
            }
            g_base = __float_as_int(hd.v[4]);
            g_bits = permuteByKey(hits & im, key) | (im << 8) | (key << 16);
            t_base = __float_as_int(hd.v[5]); t_mask = tm; t_lm = lm;
        }
        // ---- (2) who waits for what: one warp reduction over 6-bit counters, one per state
        const bool s_done = (flags & kDone) != 0u, s_exit = !s_done && (flags & kExit) != 0u;
        const bool s_leaf = !s_done && !s_exit && t_mask != 0u;
        const bool s_tri = s_leaf && (flags & kInBlas) != 0u, s_enter = s_leaf && !(flags & kInBlas);
        const unsigned state = s_done ? 3u : s_exit ? 2u : s_tri ? 0u : s_enter ? 1u : 4u;      // 0 tri, 1 enter, 2 exit, 3 done, 4 inner
        const bool live = !s_done || i >= 0 || !exhausted;      // a retired lane with nothing left to fetch takes no part
        const unsigned sum = __reduce_add_sync(kFull, live ? 1u << (6u * state) : 0u);
        if (sum == 0u) break;                                   // every ray of the launch is retired
        const int n_tri = sum & 63u, n_enter = (sum >> 6) & 63u, n_exit = (sum >> 12) & 63u, n_done = (sum >> 18) & 63u, n_inner = (sum >> 24) & 63u;

        // ---- (3a) one triangle of the triangle group (Renderer.cpp:174-215)
        if (s_tri && n_tri >= min(vote_tri, n_inner)) {
            const unsigned k = (unsigned)__ffs((int)t_mask) - 1u;
            const int pos = t_base + __popc(t_lm & ((1u << k) - 1u));
            t_mask &= t_mask - 1u;
            leafTriangle<UV, COUNT>(sc, ro, rd, tmax, best_tri, best_u, best_v, pos, cnt);
        }
        // ---- (3b) TLAS leaf: enter the next pending instance (Renderer.cpp:381-384)
        if (s_enter && n_enter >= min(vote_inst, n_inner)) {
            const unsigned k = (unsigned)__ffs((int)t_mask) - 1u;
            const int im = __ldg(&sc.tlas_order[t_base + __popc(t_lm & ((1u << k) - 1u))]);
            t_mask &= t_mask - 1u;
            const InstanceTrace* __restrict__ inst = &sc.inst[im];
            const float4 w0 = ldg4(&inst->w2m[0]), w1 = ldg4(&inst->w2m[1]), w2 = ldg4(&inst->w2m[2]);
            ro = xmat4(w0, w1, w2, bo, 1.0f);                                   // Renderer.cpp:381
            const V3 dm = xmat4(w0, w1, w2, bd, 0.0f);
            const float mlen = xsqrt(xdot(dm, dm));
            rd = xscale(dm, xdiv(1.0f, mlen));                                  // glm::normalize, Renderer.cpp:382
            rinv = v3(safeInv(rd.x), safeInv(rd.y), safeInv(rd.z));
            const float wil = rsqrtf(bd.x * bd.x + bd.y * bd.y + bd.z * bd.z);
            const float ascale = __fdividef(1.0f, mlen * wil);                  // world distance per unit of model-space t, for the exit step
            // the TLAS state waits under the instance marker: node group, instance group, world-distance scale
            stack[sp++] = make_uint2((unsigned)g_base, g_bits);
            stack[sp++] = make_uint2((unsigned)t_base, t_mask);
            stack[sp++] = make_uint2(t_lm, __float_as_uint(ascale));
            stack[sp++] = make_uint2((unsigned)im, kExitMark);
            tmin = -(kEpsilon + 1e-4f);                                         // the predicate accepts t >= -EPSILON
            tmax = kFloatMax; best_tri = -1;                                    // Renderer.cpp:384
            if (g_dist < kFloatMax) {
                // World distance of a model-space t along this ray is t * |d_w| / |W3 d_w| when model_to_world inverts world_to_model
                // (checked at upload, else sc.prune = inf): a winner needs t <= g * |W3 d_w| / |d_w|, plus slack for the matrices' residual.
                const float tb = (g_dist * sc.prune + sc.c_pad + 1e-4f * (fabsf(bo.x) + fabsf(bo.y) + fabsf(bo.z))) * (mlen * wil) * 1.0001f;
                if (tb < kFloatMax) tmax = tb;                                  // false for NaN / inf: no bound
            }
            const unsigned oct = octOf(rd);
            flags = oct | kInBlas; near_off = nearOffsets(oct);
            g_base = __float_as_int(__ldg(&inst->grid.z));                      // BLAS root of the instance's mesh, as a one-child group
            g_bits = 1u | (1u << 8);                                             // hit bit 0, inner bit 0, key 0
            t_mask = 0u;
        }
        // ---- (3c) marker popped: leave instance `exit_im` (Renderer.cpp:388-398).  The nearest-model decision of the reference compares exact
        // world distances; here the instance's winner is ranked by |t| * |d_w| / |W3 d_w| (its world distance up to `tie`), and the
        // exact distances are evaluated only when two candidates are closer than that slack, or when pruning is off.
        if (s_exit && n_exit >= min(vote_inst, n_inner)) {
            const int im = exit_im;
            const uint2 e2 = stack[--sp], e1 = stack[--sp], e0 = stack[--sp];
            const float ascale = __uint_as_float(e2.y);
            t_lm = e2.x; t_base = (int)e1.x; t_mask = e1.y; g_base = (int)e0.x; g_bits = e0.y;
            const float cb = sc.c_pad + 1e-4f * (fabsf(bo.x) + fabsf(bo.y) + fabsf(bo.z));
            if (best_tri >= 0) {
                // |t|: the reference ranks instances by length(hit - origin) >= 0 (Renderer.cpp:391-393), and the predicate accepts
                // t down to -EPSILON, which is up to EPSILON * (world units per model unit) behind the origin
                const float a = fabsf(tmax * ascale);
                const float a_lo = a - (a * sc.tie + cb), a_hi = a + (a * sc.tie + cb);
                const float g_lo = g_dist - (g_dist * sc.tie + cb), g_hi = g_dist + (g_dist * sc.tie + cb);
                bool take;
                float nd = a;
                if (a_hi < g_lo && a_hi < 0.99f * kFloatMax) take = true;        // clearly nearer than the current winner (or the first one)
                else if (g_model >= 0 && a_lo > g_hi) take = false;              // clearly farther
                else {                                                           // near-tie, or no usable bound: the reference's comparison
                    const float dn = exactHitDistance(sc, bo, bd, im, tmax);
                    const float dg = g_model >= 0 ? exactHitDistance(sc, bo, bd, g_model, g_t) : kFloatMax;
                    take = dg > dn || (dg == dn && g_model >= 0 && im < g_model);   // Renderer.cpp:393 in model order
                    nd = dn;
                    if (!take && g_model >= 0) g_dist = dg;
                }
                if (take) {
                    g_dist = nd; g_model = im; g_tri = best_tri; g_t = tmax;
                    if (UV) { g_u = best_u; g_v = best_v; }
                }
            }
            // back to the world ray (its reciprocal is recomputed: three registers fewer across the whole BLAS traversal)
            const float il = rsqrtf(bd.x * bd.x + bd.y * bd.y + bd.z * bd.z);
            const V3 wd = v3(bd.x * il, bd.y * il, bd.z * il);
            ro = bo; rinv = v3(safeInv(wd.x), safeInv(wd.y), safeInv(wd.z));
            flags = octOf(wd); near_off = nearOffsets(flags);
            tmin = sc.tmin_world; tmax = worldBound(g_dist, sc.prune, cb);
        }
        // ---- (3d) retire finished rays, refill the lanes from the warp's batch
        if (n_done > 0 && n_done >= min(vote_refill, n_inner)) {
            const unsigned m_done = __ballot_sync(kFull, s_done && live);
            if (s_done && i >= 0) {
                const bool found = g_dist < kFloatMax;
                // hit.x < 0 (a constant: the consumers only test the sign) tells the consumer to evaluate the exact world distance
                // from (model, t) (kernels.cuh: exactHitDistance)
                hit[i] = make_float4(found ? -1.0f : kFloatMax, __int_as_float(found ? g_tri : -1), __int_as_float(found ? g_model : -1), g_t);
                if (UV && uv) uv[i] = make_float2(g_u, g_v);
                if (COUNT) { if (counts) counts[i] = cnt; tot_x += cnt.x; tot_z += cnt.z; }
                i = -1;
            }
            if (w_next >= w_end && !exhausted) {      // the warp's batch is used up: take the next one
                unsigned b = 0;
                if (lane == 0) b = atomicAdd(cursor, (unsigned)sc.batch);
                b = __shfl_sync(kFull, b, 0);
                if (b >= (unsigned)n) { exhausted = true; w_next = w_end = n; }
                else { w_next = (int)b; w_end = min((int)b + sc.batch, n); }
            }
            const int avail = w_end - w_next;
            const int rank = __popc(m_done & ((1u << lane) - 1u));
            if (s_done && live && rank < avail) {
                i = w_next + rank;
                const float4 o4 = O[i], d4 = D[i];
                bo = v3(o4); bd = v3(d4);
                const float il = rsqrtf(bd.x * bd.x + bd.y * bd.y + bd.z * bd.z);
                const V3 wd = v3(bd.x * il, bd.y * il, bd.z * il);
                ro = bo; rinv = v3(safeInv(wd.x), safeInv(wd.y), safeInv(wd.z));
                tmin = sc.tmin_world; tmax = 3.0e38f;
                g_dist = kFloatMax; g_model = -1; g_tri = -1; g_t = 0.0f; g_u = 0.0f; g_v = 0.0f;
                best_tri = -1;
                if (COUNT) cnt = make_int4(0, 0, 0, 0);
                stack[0] = make_uint2(0u, kDoneMark); sp = 1;
                flags = octOf(wd); near_off = nearOffsets(flags);
                g_base = sc.tlas_root; g_bits = 1u | (1u << 8);                  // the TLAS root as a one-child group (key 0)
                t_mask = 0u;
            }
            w_next += min(__popc(m_done), avail);
        }
    }
    if (stamp && (threadIdx.x & 31) == 0) atomicMax(stamp + 1, globalTimerNs());
    if (COUNT) {        // counting build: per-warp totals into the frame state (never used for timing)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { tot_x += __shfl_xor_sync(kFull, tot_x, d); tot_z += __shfl_xor_sync(kFull, tot_z, d); }
        if (lane == 0) { atomicAdd(&st->count_nodes, tot_x); atomicAdd(&st->count_tris, tot_z); }
    }
}

void launchTraceBvh(const SceneDev& sc, const float4* O, const float4* D, float4* hit, float2* uv, int4* counts, bool count_totals,
                     FrameState* st, int round, int n_fixed, int grid, cudaStream_t stream, unsigned long long* stamp)
{
    if (counts || count_totals) k_trace_bvh<true, true><<<grid, kTraceBlock, 0, stream>>>(sc, O, D, hit, uv, counts, st, round, n_fixed, stamp);
    else if (uv) k_trace_bvh<true, false><<<grid, kTraceBlock, 0, stream>>>(sc, O, D, hit, uv, counts, st, round, n_fixed, stamp);
    else k_trace_bvh<false, false><<<grid, kTraceBlock, 0, stream>>>(sc, O, D, hit, uv, counts, st, round, n_fixed, stamp);
}

int traceBvhOccupancy()
{
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_bvh<false, false>, kTraceBlock, 0);
    return nb;
}

}  // namespace ptap
