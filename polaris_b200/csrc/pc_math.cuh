// pc_math.cuh -- vector helpers with OpenCL C semantics for the CUDA tracer.
//
// The reference's device code is OpenCL C (tracer/opencl/CL/**).  These helpers give CUDA's
// float2/3/4 the operator and built-in semantics that code relies on:
// component-wise operators with scalar widening, mix(a,b,t) = a + (b-a)*t, normalize(v) =
// v / sqrt(dot(v,v)), min/max/clamp/sign as the OpenCL 1.2 spec defines them, native_recip
// == IEEE 1/x.  The library is compiled with -fmad=false -prec-div=true -prec-sqrt=true so
// every +,-,*,/ and sqrt below is a single correctly rounded float32 operation, which is what
// makes traversal results bit-comparable with the CPU oracle.
//
// Everything is PC_HD (host+device) so tests/emul can compile the same functions with g++.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define PC_HD __host__ __device__ __forceinline__
// Out-of-line helpers (all arguments by value, so nothing is forced into local memory): they keep
// k_shade's code footprint inside the instruction caches; inlining them at every call site made
// the kernel 212 KB of SASS and the warps stalled on instruction fetch.
#ifdef PC_INLINE_ALL
#define PC_HD_NOINLINE __host__ __device__ __forceinline__
#else
#define PC_HD_NOINLINE static __host__ __device__ __noinline__  // static: one copy per translation unit (pc_host.cu, pc_shade.cu)
#endif
#define PC_D __device__ __forceinline__
#else
#define PC_HD inline
#define PC_HD_NOINLINE inline
#endif

#if defined(__CUDA_ARCH__)
#define PC_LDG(p) __ldg(p)
#else
#define PC_LDG(p) (*(p))
#endif

namespace pc {

PC_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PC_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

PC_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
PC_HD float3 f3s(float s) { return make_float3(s, s, s); }
PC_HD float3 xyz(float4 v) { return make_float3(v.x, v.y, v.z); }
PC_HD float4 f4(float3 v, float w) { return make_float4(v.x, v.y, v.z, w); }

PC_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
PC_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
PC_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
PC_HD float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
PC_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
PC_HD float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
PC_HD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
PC_HD float3 operator+(float3 a, float s) { return f3(a.x + s, a.y + s, a.z + s); }
PC_HD float3 operator-(float3 a, float s) { return f3(a.x - s, a.y - s, a.z - s); }
PC_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }

PC_HD float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
PC_HD float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
PC_HD float4 operator*(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
PC_HD float4 operator*(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
PC_HD float4 operator/(float4 a, float s) { return make_float4(a.x / s, a.y / s, a.z / s, a.w / s); }

PC_HD float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
PC_HD float2 operator*(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
PC_HD float2 operator*(float s, float2 a) { return make_float2(s * a.x, s * a.y); }

PC_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PC_HD float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
PC_HD float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PC_HD float length(float3 a) { return sqrtf(dot(a, a)); }
PC_HD float3 normalize(float3 a) { return a / sqrtf(dot(a, a)); }
PC_HD float4 normalize(float4 a) { return a / sqrtf(dot(a, a)); }
PC_HD float cl_min(float a, float b) { return b < a ? b : a; }
PC_HD float cl_max(float a, float b) { return a < b ? b : a; }
PC_HD float cl_clamp(float x, float lo, float hi) { return cl_min(cl_max(x, lo), hi); }
PC_HD uint32_t cl_clampu(uint32_t x, uint32_t lo, uint32_t hi) { uint32_t m = x < lo ? lo : x; return m > hi ? hi : m; }
PC_HD int cl_clampi(int x, int lo, int hi) { int m = x < lo ? lo : x; return m > hi ? hi : m; }
PC_HD float cl_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : (x == 0.0f ? x : 0.0f)); }
PC_HD float mix(float a, float b, float t) { return a + (b - a) * t; }
PC_HD float4 mix(float4 a, float4 b, float t) { return a + (b - a) * t; }
PC_HD float max3(float3 v) { return cl_max(v.x, cl_max(v.y, v.z)); }

// constants.cl
#define PC_PI 3.14159265358979323846f
#define PC_TWO_PI 6.28318530718f
#define PC_1_PI 0.31830988618379067154f
#define PC_EPS 0.00001f
#define PC_LIGHT_EPS (0.00001f * 1e3f)
#define PC_MIN_ROUGHNESS 0.1f

}  // namespace pc
