// exact_math.cuh - binary32 arithmetic in the evaluation order of the reference's vendored glm 0.9.6.3,
// never contracted to FMA: every operation is an explicit round-to-nearest intrinsic, so results are
// bit-equal to the host-compiled reference (g++ -ffp-contract=off) whatever -fmad says.
// Citations: /root/reference/PathTracerAP/external/include/glm/detail/*.inl.
#pragma once
#include <cuda_runtime.h>

namespace ptap {

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 v3(const float4& a) { return v3(a.x, a.y, a.z); }
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ V3 xadd(V3 a, V3 b) { return v3(xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)); }
__device__ __forceinline__ V3 xsub(V3 a, V3 b) { return v3(xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)); }
__device__ __forceinline__ V3 xmul(V3 a, V3 b) { return v3(xmul(a.x, b.x), xmul(a.y, b.y), xmul(a.z, b.z)); }
__device__ __forceinline__ V3 xscale(V3 a, float s) { return v3(xmul(a.x, s), xmul(a.y, s), xmul(a.z, s)); }
// func_geometric.inl:65-72: (x + y) + z
__device__ __forceinline__ float xdot(V3 a, V3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
// func_geometric.inl:134-142
__device__ __forceinline__ V3 xcross(V3 x, V3 y)
{
    return v3(xsub(xmul(x.y, y.z), xmul(y.y, x.z)), xsub(xmul(x.z, y.x), xmul(y.z, x.x)), xsub(xmul(x.x, y.y), xmul(y.x, x.y)));
}
// func_geometric.inl:154-159 + func_exponential.inl:150-152: v * (1 / sqrt(dot(v, v)))
__device__ __forceinline__ V3 xnormalize(V3 v) { return xscale(v, xdiv(1.0f, xsqrt(xdot(v, v)))); }
__device__ __forceinline__ float xlength(V3 v) { return xsqrt(xdot(v, v)); }
// type_mat4x4.inl:617-627 restricted to xyz: (m0*x + m1*y) + (m2*z + m3*w); rows r[k] = (m[0][k], m[1][k], m[2][k], m[3][k])
__device__ __forceinline__ V3 xmat4(const float4& r0, const float4& r1, const float4& r2, V3 v, float w)
{
    return v3(xadd(xadd(xmul(r0.x, v.x), xmul(r0.y, v.y)), xadd(xmul(r0.z, v.z), xmul(r0.w, w))),
              xadd(xadd(xmul(r1.x, v.x), xmul(r1.y, v.y)), xadd(xmul(r1.z, v.z), xmul(r1.w, w))),
              xadd(xadd(xmul(r2.x, v.x), xmul(r2.y, v.y)), xadd(xmul(r2.z, v.z), xmul(r2.w, w))));
}
// the host-compiled reference resolves min/max to std::min/std::max (oracle/shim); equal to fminf/fmaxf for non-NaN
__device__ __forceinline__ float min_std(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float max_std(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float xabs(float x) { return x < 0 ? -x : x; }   // utility.h:14
// float -> int as the host reference converts (x86 cvttss2si): out of range -> INT_MIN
__device__ __forceinline__ int f2i_x86(float x) { return (x >= -2147483648.0f && x < 2147483648.0f) ? __float2int_rz(x) : (-2147483647 - 1); }

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// 32-byte read-only vector load (LDG.E.256, sm_100): halves the load instructions and L1 tag look-ups of a 64-byte record
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ldg8(const void* p)
{
    F8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}

}  // namespace ptap
