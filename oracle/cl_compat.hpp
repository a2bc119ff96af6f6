// cl_compat.hpp -- just enough of OpenCL C 1.2 for g++ to compile the reference's own kernel sources.
//
// TEST INFRASTRUCTURE (oracle/_ref): oracle/build_ref.py concatenates /root/reference/tracer/opencl/CL/
// main.cl with its includes where they lie, rewrites the one piece of OpenCL C syntax C++ cannot
// parse -- vector literals `(float3)(a, b, c)` become `make_float3(a, b, c)` -- and compiles the
// result against this header into oracle/_ref/libpolaris_clref.so.  Nothing of the reference is
// copied into the repository; the generated translation unit lives in a temporary directory.
//
// Semantics follow the OpenCL 1.2 specification for every built-in the kernels use (list taken from
// a scan of the .cl files): vectors are component-wise with scalar widening, float3 has the size
// and alignment of float4 (6.1.5), `mix(a,b,t) = a + (b-a)*t` (6.12.4), `fmin/fmax` return the
// non-NaN operand (6.12.2), conversions truncate (float->int) or round to nearest even
// (uint->float, 6.2.3.3), `sign(+-0) = +-0`.  The runtime-defined `native_*` functions and
// `normalize`/`length` are given their correctly rounded IEEE meaning (1.0f/x, sqrtf, libm cosf/
// sinf, v / sqrtf(dot(v,v))), exactly the choices DESIGN.md states for the CUDA kernels and the
// oracle port -- a real OpenCL CPU runtime may differ from this in the last
// ulp of those built-ins, nowhere else.
//
// Execution model: one work-item at a time, work-group size 1 (so `barrier` is a no-op and
// `__local` variables are plain automatics); ref_driver.cpp owns the NDRange loops.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <math.h>

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong;

#define __kernel
#define __global
#define global
#define __local
#define __constant const
#define __private
#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2
#define printf(...) ((void)0) /* printSurface / printIntersection pass vectors through %v3hlf */

// ---- work-item functions ------------------------------------------------------------------------
struct ClWorkItem {
    size_t global_id[3];
    size_t local_id[3];
    size_t group_id[3];
    size_t global_size[3];
};
static thread_local ClWorkItem cl_wi;
static inline size_t get_global_id(uint d) { return cl_wi.global_id[d]; }
static inline size_t get_local_id(uint d) { return cl_wi.local_id[d]; }
static inline size_t get_group_id(uint d) { return cl_wi.group_id[d]; }
static inline size_t get_global_size(uint d) { return cl_wi.global_size[d]; }
static inline size_t get_local_size(uint) { return 1; }
static inline void barrier(int) {}

// one work-item runs at a time per counter (ref_driver.cpp gives every OpenMP chunk private counters)
static inline int atomic_inc(volatile int *p) { int o = *p; *p = o + 1; return o; }
static inline int atomic_add(volatile int *p, int v) { int o = *p; *p = o + v; return o; }

// ---- vector types -------------------------------------------------------------------------------
template <class T, int N> struct ClVecStorage;

// swizzle proxy for `.xyz` (the only multi-component swizzle the kernels use) on 3- and 4-vectors.
// It must stay trivially copyable to live in the unions below, so swizzle = swizzle is NOT defined
// (it would copy .w too); build_ref.py refuses sources that contain such an assignment.
template <class V3, class T> struct ClSwzXYZ {
    T d[4];
    operator V3() const { return V3(d[0], d[1], d[2]); }
    ClSwzXYZ &operator=(const V3 &v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; return *this; }
};

#define CL_VEC2(NAME, T)                                                                           \
    struct alignas(2 * sizeof(T)) NAME {                                                           \
        union { struct { T x, y; }; struct { T s0, s1; }; T d[2]; };                               \
        NAME() = default;                                                                          \
        NAME(T a, T b) { x = a; y = b; }                                                             \
        NAME(T a) { x = a; y = a; }                                                       \
    };
#define CL_VEC3(NAME, T)                                                                           \
    struct alignas(4 * sizeof(T)) NAME {                                                           \
        union { struct { T x, y, z; }; struct { T r, g, b; }; struct { T s0, s1, s2; }; T d[4];    \
                ClSwzXYZ<NAME, T> xyz; };                                                          \
        NAME() = default;                                                                          \
        NAME(T a, T b_, T c) { x = a; y = b_; z = c; }                                                \
        NAME(T a) { x = a; y = a; z = a; }                                                   \
    };
#define CL_VEC4(NAME, NAME3, T)                                                                    \
    struct alignas(4 * sizeof(T)) NAME {                                                           \
        union { struct { T x, y, z, w; }; struct { T r, g, b, a; }; struct { T s0, s1, s2, s3; };  \
                T d[4]; ClSwzXYZ<NAME3, T> xyz; };                                                 \
        NAME() = default;                                                                          \
        NAME(T a_, T b_, T c, T e) { x = a_; y = b_; z = c; w = e; }                                   \
        NAME(const NAME3 &v, T e) { x = v.x; y = v.y; z = v.z; w = e; }                                \
        NAME(T a_) { x = a_; y = a_; z = a_; w = a_; }                                        \
    };

CL_VEC2(float2, float)
CL_VEC3(float3, float)
CL_VEC4(float4, float3, float)
CL_VEC2(uint2, uint)
CL_VEC2(int2, int)
CL_VEC3(int3, int)
CL_VEC4(int4, int3, int)
CL_VEC3(uint3, uint)
CL_VEC4(uint4, uint3, uint)
CL_VEC3(uchar3, uchar)
CL_VEC4(uchar4, uchar3, uchar)

static_assert(sizeof(float2) == 8 && sizeof(float3) == 16 && sizeof(float4) == 16, "OpenCL vector sizes");
static_assert(alignof(float3) == 16 && sizeof(uchar4) == 4 && sizeof(int4) == 16 && sizeof(uint2) == 8, "OpenCL vector sizes");

// (the broadcast constructors are implicit: OpenCL C widens the scalar in `cond ? vec : 0.0f`)
// vector literals: `(floatN)(...)` is rewritten to make_floatN(...) by build_ref.py
template <class A, class B> static inline float2 make_float2(A a, B b) { return float2((float)a, (float)b); }
template <class A> static inline float2 make_float2(A a) { return float2((float)a); }
template <class A, class B, class C> static inline float3 make_float3(A a, B b, C c) { return float3((float)a, (float)b, (float)c); }
template <class A> static inline float3 make_float3(A a) { return float3((float)a); }
static inline float3 make_float3(const float3 &v) { return v; }
template <class A, class B, class C, class D> static inline float4 make_float4(A a, B b, C c, D d) { return float4((float)a, (float)b, (float)c, (float)d); }
template <class D> static inline float4 make_float4(const float3 &v, D d) { return float4(v, (float)d); }
template <class D> static inline float4 make_float4(const ClSwzXYZ<float3, float> &v, D d) { return float4((float3)v, (float)d); }
template <class A> static inline float4 make_float4(A a) { return float4((float)a); }
template <class A, class B> static inline uint2 make_uint2(A a, B b) { return uint2((uint)a, (uint)b); }
template <class A, class B, class C, class D> static inline uchar4 make_uchar4(A a, B b, C c, D d) { return uchar4((uchar)a, (uchar)b, (uchar)c, (uchar)d); }
template <class A, class B, class C, class D> static inline int4 make_int4(A a, B b, C c, D d) { return int4((int)a, (int)b, (int)c, (int)d); }

// ---- operators (component-wise, scalar widening) ---------------------------------------------------
#define CL_BINOP2(V, T, OP)                                                                        \
    static inline V operator OP(const V &a, const V &b) { return V(a.x OP b.x, a.y OP b.y); }      \
    static inline V operator OP(const V &a, T s) { return V(a.x OP s, a.y OP s); }                 \
    static inline V operator OP(T s, const V &a) { return V(s OP a.x, s OP a.y); }                 \
    static inline V &operator OP##=(V &a, const V &b) { a = a OP b; return a; }                    \
    static inline V &operator OP##=(V &a, T s) { a = a OP s; return a; }
#define CL_BINOP3(V, T, OP)                                                                        \
    static inline V operator OP(const V &a, const V &b) { return V(a.x OP b.x, a.y OP b.y, a.z OP b.z); } \
    static inline V operator OP(const V &a, T s) { return V(a.x OP s, a.y OP s, a.z OP s); }       \
    static inline V operator OP(T s, const V &a) { return V(s OP a.x, s OP a.y, s OP a.z); }       \
    static inline V &operator OP##=(V &a, const V &b) { a = a OP b; return a; }                    \
    static inline V &operator OP##=(V &a, T s) { a = a OP s; return a; }
#define CL_BINOP4(V, T, OP)                                                                        \
    static inline V operator OP(const V &a, const V &b) { return V(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
    static inline V operator OP(const V &a, T s) { return V(a.x OP s, a.y OP s, a.z OP s, a.w OP s); } \
    static inline V operator OP(T s, const V &a) { return V(s OP a.x, s OP a.y, s OP a.z, s OP a.w); } \
    static inline V &operator OP##=(V &a, const V &b) { a = a OP b; return a; }                    \
    static inline V &operator OP##=(V &a, T s) { a = a OP s; return a; }
#define CL_ARITH(M, V, T) M(V, T, +) M(V, T, -) M(V, T, *) M(V, T, /)
CL_ARITH(CL_BINOP2, float2, float)
CL_ARITH(CL_BINOP3, float3, float)
CL_ARITH(CL_BINOP4, float4, float)
CL_ARITH(CL_BINOP2, uint2, uint)
static inline float2 operator-(const float2 &a) { return float2(-a.x, -a.y); }
static inline float3 operator-(const float3 &a) { return float3(-a.x, -a.y, -a.z); }
static inline float4 operator-(const float4 &a) { return float4(-a.x, -a.y, -a.z, -a.w); }
// the swizzle proxy only converts implicitly for non-template calls; unary minus on it needs this
static inline float3 operator-(const ClSwzXYZ<float3, float> &a) { return -(float3)a; }

// ---- built-ins ---------------------------------------------------------------------------------
static inline float dot(const float2 &a, const float2 &b) { return a.x * b.x + a.y * b.y; }
static inline float dot(const float3 &a, const float3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float dot(const float4 &a, const float4 &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
static inline float3 cross(const float3 &a, const float3 &b) {
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float length(const float2 &a) { return sqrtf(dot(a, a)); }
static inline float length(const float3 &a) { return sqrtf(dot(a, a)); }
static inline float length(const float4 &a) { return sqrtf(dot(a, a)); }
static inline float2 normalize(const float2 &a) { return a / sqrtf(dot(a, a)); }
static inline float3 normalize(const float3 &a) { return a / sqrtf(dot(a, a)); }
static inline float4 normalize(const float4 &a) { return a / sqrtf(dot(a, a)); }

static inline float native_recip(float x) { return 1.0f / x; }
static inline float3 native_recip(const float3 &v) { return float3(1.0f / v.x, 1.0f / v.y, 1.0f / v.z); }
static inline float native_sqrt(float x) { return sqrtf(x); }
static inline float native_cos(float x) { return cosf(x); }
static inline float native_sin(float x) { return sinf(x); }

static inline float mix(float a, float b, float t) { return a + (b - a) * t; }
static inline float2 mix(const float2 &a, const float2 &b, float t) { return a + (b - a) * t; }
static inline float3 mix(const float3 &a, const float3 &b, float t) { return a + (b - a) * t; }
static inline float4 mix(const float4 &a, const float4 &b, float t) { return a + (b - a) * t; }

static inline float3 fmin(const float3 &a, const float3 &b) { return float3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
static inline float3 fmax(const float3 &a, const float3 &b) { return float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }

// min / max / clamp on scalars (6.12.4: undefined for NaN, none occur); explicit overloads keep
// calls with mixed int literals unambiguous
static inline float min(float a, float b) { return b < a ? b : a; }
static inline float max(float a, float b) { return a < b ? b : a; }
static inline int min(int a, int b) { return b < a ? b : a; }
static inline int max(int a, int b) { return a < b ? b : a; }
static inline uint min(uint a, uint b) { return b < a ? b : a; }
static inline uint max(uint a, uint b) { return a < b ? b : a; }
static inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
static inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
static inline uint clamp(uint x, uint lo, uint hi) { return min(max(x, lo), hi); }
static inline int clamp(int x, int lo, uint hi) { return min(max(x, lo), (int)hi); }
static inline float3 clamp(const float3 &v, float lo, float hi) { return float3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
static inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : (x == 0.0f ? x : 0.0f)); }
static inline float3 pow(const float3 &v, float e) { return float3(powf(v.x, e), powf(v.y, e), powf(v.z, e)); }
static inline float2 floor(const float2 &v) { return float2(floorf(v.x), floorf(v.y)); }

static inline float2 convert_float2(const uint2 &v) { return float2((float)v.x, (float)v.y); }  // rte
static inline float4 convert_float4(const uchar4 &v) { return float4((float)v.x, (float)v.y, (float)v.z, (float)v.w); }
