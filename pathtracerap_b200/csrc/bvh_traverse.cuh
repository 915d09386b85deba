// bvh_traverse.cuh - what the kernels that walk the two-level BVH share (k_trace_bvh in trace_bvh.cu, k_trace_emu in trace_emu.cu): the
// encoding of a lane's state in its `node` register, the one-FMA-per-plane child test on node-local offsets, and the pruning bounds.
// See trace_bvh.cu for the design.
#pragma once
#include "kernels.cuh"

namespace ptap {
namespace bvh {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kDone = (int)0x80000000u;        // ~0x7fffffff: bottom-of-stack sentinel
// A negative `node` is ~code with the lane's state in code >> 29:
//   0: triangle leaf (first << 3 | count - 1), 1: TLAS leaf = enter instance (code & kIndexMask), 2: marker = leave instance, 3: done
constexpr unsigned kEnterBit = 0x20000000u, kExitBit = 0x40000000u, kIndexMask = 0x1fffffffu;
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)          // FFMA2 (sm_100): two binary32 FMAs per issue slot
{
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ uint4 ldg4u(const void* p)
{
    uint4 r;
    asm("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Entry distance of a child as an order-preserving integer key (a missed child sorts last): interval test with relative slack on the
// entry / exit parameters of the six planes (near / far already selected by the ray's direction signs).
__device__ __forceinline__ int childKey(float nx, float ny, float nz, float fx, float fy, float fz, float tmin, float tmax)
{
    const float tn = fmaxf(fmaxf(fmaxf(nx, ny), nz), tmin);
    const float tf = fminf(fminf(fminf(fx, fy), fz), tmax);
    const bool h = tn <= tf + __fmaf_rn(fabsf(tf), 2e-6f, 1e-6f);
    return h ? __float_as_int(fmaxf(tn, 0.0f)) : 0x7f800000;
}

// direction-sign bits, taken from the SIGN BIT (not d < 0) so that they agree with safeInv for -0.0f: they decide which plane of a pair
// is the near one, and a near / far swap against the sign of the reciprocal would turn every box into a miss.
// Result: byte offsets (x | y << 8 | z << 16) of the near planes inside a node: lower planes at 32 / 64 / 96, upper planes 16 further.
__device__ __forceinline__ unsigned nearOffsets(const V3& d)
{
    const unsigned sx = __float_as_uint(d.x) >> 31, sy = __float_as_uint(d.y) >> 31, sz = __float_as_uint(d.z) >> 31;
    return (32u + (sx << 4)) | ((64u + (sy << 4)) << 8) | ((96u + (sz << 4)) << 16);
}

// traversal-only reciprocal: guarded against 0 (the exact predicate never uses it)
__device__ __forceinline__ float safeInv(float d)
{
    const float ooeps = 1e-30f;
    return __fdividef(1.0f, fabsf(d) > ooeps ? d : copysignf(ooeps, d));     // 1-ulp reciprocal: inside the slab test's 2e-6 slack
}

// TLAS pruning bound once some instance reported world distance g_dist: g_dist may be the approximation t * |d_w| / |W3 d_w| of the exact
// distance (off by at most g_dist * tie + cb, tie < prune - 1), so the bound carries the same absolute slack cb as the instance-entry bound
__device__ __forceinline__ float worldBound(float g_dist, float prune, float cb)
{
    return g_dist < kFloatMax ? g_dist * prune + cb + 1e-3f : 3.0e38f;
}

// The four children of an inner node, tested against the ray interval [tmin, tmax] and sorted by entry distance: key[c] = order-preserving
// integer key of child c's entry distance (0x7f800000: missed), lnk[c] its link.  One FMA per plane on node-local offsets (two children per
// FFMA2), near / far planes picked by address from the direction's sign bits (near_off: nearOffsets()).
__device__ __forceinline__ void nodeChildren(const char* __restrict__ np, const V3& ro, const V3& rinv, unsigned near_off, float tmin, float tmax, int key[4], int lnk[4])
{
    const F8 hd = ldg8(np);                                           // origin, four links
    const unsigned nox = near_off & 0xffu, noy = (near_off >> 8) & 0xffu, noz = near_off >> 16;
    const uint4 NX = ldg4u(np + nox), FX = ldg4u(np + (nox ^ 16u));   // near / far planes by address
    const uint4 NY = ldg4u(np + noy), FY = ldg4u(np + (noy ^ 16u));
    const uint4 NZ = ldg4u(np + noz), FZ = ldg4u(np + (noz ^ 16u));
    const float cx = (hd.v[0] - ro.x) * rinv.x, cy = (hd.v[1] - ro.y) * rinv.y, cz = (hd.v[2] - ro.z) * rinv.z;
    const float2 ix = make_float2(rinv.x, rinv.x), iy = make_float2(rinv.y, rinv.y), iz = make_float2(rinv.z, rinv.z);
    const float2 ccx = make_float2(cx, cx), ccy = make_float2(cy, cy), ccz = make_float2(cz, cz);
    const float2 nxa = fma2(make_float2(__uint_as_float(NX.x), __uint_as_float(NX.y)), ix, ccx), nxb = fma2(make_float2(__uint_as_float(NX.z), __uint_as_float(NX.w)), ix, ccx);
    const float2 nya = fma2(make_float2(__uint_as_float(NY.x), __uint_as_float(NY.y)), iy, ccy), nyb = fma2(make_float2(__uint_as_float(NY.z), __uint_as_float(NY.w)), iy, ccy);
    const float2 nza = fma2(make_float2(__uint_as_float(NZ.x), __uint_as_float(NZ.y)), iz, ccz), nzb = fma2(make_float2(__uint_as_float(NZ.z), __uint_as_float(NZ.w)), iz, ccz);
    const float2 fxa = fma2(make_float2(__uint_as_float(FX.x), __uint_as_float(FX.y)), ix, ccx), fxb = fma2(make_float2(__uint_as_float(FX.z), __uint_as_float(FX.w)), ix, ccx);
    const float2 fya = fma2(make_float2(__uint_as_float(FY.x), __uint_as_float(FY.y)), iy, ccy), fyb = fma2(make_float2(__uint_as_float(FY.z), __uint_as_float(FY.w)), iy, ccy);
    const float2 fza = fma2(make_float2(__uint_as_float(FZ.x), __uint_as_float(FZ.y)), iz, ccz), fzb = fma2(make_float2(__uint_as_float(FZ.z), __uint_as_float(FZ.w)), iz, ccz);
    key[0] = childKey(nxa.x, nya.x, nza.x, fxa.x, fya.x, fza.x, tmin, tmax);
    key[1] = childKey(nxa.y, nya.y, nza.y, fxa.y, fya.y, fza.y, tmin, tmax);
    key[2] = childKey(nxb.x, nyb.x, nzb.x, fxb.x, fyb.x, fzb.x, tmin, tmax);
    key[3] = childKey(nxb.y, nyb.y, nzb.y, fxb.y, fyb.y, fzb.y, tmin, tmax);
#pragma unroll
    for (int c = 0; c < 4; ++c) lnk[c] = __float_as_int(hd.v[4 + c]);
#define PTAP_CSWAP(a, b) { const bool sw = key[b] < key[a]; const int ka = sw ? key[b] : key[a], kb = sw ? key[a] : key[b]; \
                           const int la = sw ? lnk[b] : lnk[a], lb = sw ? lnk[a] : lnk[b]; key[a] = ka; key[b] = kb; lnk[a] = la; lnk[b] = lb; }
    PTAP_CSWAP(0, 1) PTAP_CSWAP(2, 3) PTAP_CSWAP(0, 2) PTAP_CSWAP(1, 3) PTAP_CSWAP(1, 2)
#undef PTAP_CSWAP
}

}  // namespace bvh
}  // namespace ptap
