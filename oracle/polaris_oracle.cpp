// polaris_oracle.cpp -- CPU restatement of the reference's tracer hot path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this library; the product path
// (libpolaris_cuda.so) never does and has no CPU fallback.
//
// What it restates (file:line into /root/reference, the achilleasa/polaris checkout):
//   device code  tracer/opencl/CL/**            (every kernel except kernels/debug.cl)
//   host loops   tracer/opencl/tracer.go:194-286, pipeline.go:94-213, resources.go:81-360
// Each function below cites the lines it follows; statement order, operator association and
// variable names are kept so the two can be read side by side.  Arithmetic is IEEE float32:
// build with -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile does), `native_x` is the
// correctly rounded `x` (1.0f/x, sqrtf) or libm (cosf, sinf), see SURVEY appendix A.
//
// Parity status: PINNED against the reference's own kernels compiled for the CPU
// (oracle/_ref, built from /root/reference/tracer/opencl/CL by oracle/build_ref.py) by
// tests/test_oracle_vs_ref.py whenever that library is present.  The reference's own unit
// tests hold no golden vectors for this path.
//
// Deliberate differences from the reference, all listed in DESIGN.md:
//   * compaction order inside shadeHits is ray-index order (the reference's is atomic arrival
//     order inside a work-group, pt_integrator.cl:162,176 -- SURVEY Q13); with a work-group
//     size of 1 and in-order groups the reference produces exactly this order;
//   * uninitialised reads are defined: a missed ray's hit record is zero except wuvt.w ==
//     tmax (intersect.cl:221-222), light-sampling temporaries start at 0;
//   * emissive hits accumulate at paths[..].pixelIndex unless fix_q4 == 0, in which case the
//     ray's path index is used like pt_integrator.cl:106 (identical whenever BlockY == 0);
//   * the primary rays use rayIntersectionQuery, what the reference does on CPU devices
//     (pipeline.go:107-111); the packet kernel is GPU-only.
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/polaris_cuda.h"  // pc_scene_view, pc_block_request: the ABI structs only

namespace {

// ------------------------------------------------------------------------------------------
// OpenCL C value types with OpenCL C operator semantics (component-wise, scalar widening).
// ------------------------------------------------------------------------------------------
struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };
struct u2 { uint32_t x, y; };

inline f3 F3(float x, float y, float z) { return f3{x, y, z}; }
inline f3 F3(float s) { return f3{s, s, s}; }
inline f3 xyz(const f4 &v) { return f3{v.x, v.y, v.z}; }

inline f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator*(f3 a, f3 b) { return f3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline f3 operator/(f3 a, f3 b) { return f3{a.x / b.x, a.y / b.y, a.z / b.z}; }
inline f3 operator*(f3 a, float s) { return f3{a.x * s, a.y * s, a.z * s}; }
inline f3 operator*(float s, f3 a) { return f3{s * a.x, s * a.y, s * a.z}; }
inline f3 operator/(f3 a, float s) { return f3{a.x / s, a.y / s, a.z / s}; }
inline f3 operator+(f3 a, float s) { return f3{a.x + s, a.y + s, a.z + s}; }
inline f3 operator-(f3 a, float s) { return f3{a.x - s, a.y - s, a.z - s}; }
inline f3 operator-(f3 a) { return f3{-a.x, -a.y, -a.z}; }
inline f3 &operator*=(f3 &a, f3 b) { a = a * b; return a; }
inline f3 &operator/=(f3 &a, float s) { a = a / s; return a; }
inline f3 &operator+=(f3 &a, f3 b) { a = a + b; return a; }

inline f4 operator+(f4 a, f4 b) { return f4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline f4 operator-(f4 a, f4 b) { return f4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline f4 operator*(f4 a, float s) { return f4{a.x * s, a.y * s, a.z * s, a.w * s}; }
inline f4 operator*(float s, f4 a) { return f4{s * a.x, s * a.y, s * a.z, s * a.w}; }
inline f4 operator/(f4 a, float s) { return f4{a.x / s, a.y / s, a.z / s, a.w / s}; }

inline f2 operator+(f2 a, f2 b) { return f2{a.x + b.x, a.y + b.y}; }
inline f2 operator*(f2 a, f2 b) { return f2{a.x * b.x, a.y * b.y}; }
inline f2 operator*(float s, f2 a) { return f2{s * a.x, s * a.y}; }

// OpenCL built-ins (OpenCL 1.2 spec 6.12.2/6.12.4/6.12.5), see SURVEY appendix A.
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(f4 a, f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline f3 cross(f3 a, f3 b) { return f3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(f3 a) { return sqrtf(dot(a, a)); }
inline f3 normalize(f3 a) { return a / sqrtf(dot(a, a)); }
inline f4 normalize(f4 a) { return a / sqrtf(dot(a, a)); }
inline float cl_min(float a, float b) { return b < a ? b : a; }
inline float cl_max(float a, float b) { return a < b ? b : a; }
inline float cl_clamp(float x, float lo, float hi) { return cl_min(cl_max(x, lo), hi); }
inline uint32_t cl_clamp(uint32_t x, uint32_t lo, uint32_t hi) { uint32_t m = x < lo ? lo : x; return m > hi ? hi : m; }
inline int cl_clamp(int x, int lo, int hi) { int m = x < lo ? lo : x; return m > hi ? hi : m; }
inline float cl_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : (x == 0.0f ? x : 0.0f)); }
inline float mix(float a, float b, float t) { return a + (b - a) * t; }
inline f4 mix(f4 a, f4 b, float t) { return a + (b - a) * t; }
inline f3 fmin3(f3 a, f3 b) { return f3{fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
inline f3 fmax3(f3 a, f3 b) { return f3{fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
inline float native_recip(float x) { return 1.0f / x; }
inline f3 native_recip(f3 v) { return f3{1.0f / v.x, 1.0f / v.y, 1.0f / v.z}; }
inline float native_sqrt(float x) { return sqrtf(x); }
inline float native_cos(float x) { return cosf(x); }
inline float native_sin(float x) { return sinf(x); }

// constants.cl
constexpr float C_PI = 3.14159265358979323846f;
constexpr float C_TWO_TIMES_PI = 6.28318530718f;
constexpr float C_1_PI = 0.31830988618379067154f;
constexpr float INTERSECTION_EPSILON = 0.00001f;
constexpr float INTERSECTION_WITH_LIGHT_EPSILON = INTERSECTION_EPSILON * 1e3f;
constexpr float MIN_ROUGHNESS = 0.1f;

// ------------------------------------------------------------------------------------------
// types.cl
// ------------------------------------------------------------------------------------------
struct Ray { f4 origin; f4 dir; };                               // types.cl:4-10
struct Path { f3 throughput; float _pad; uint32_t pixelIndex, flags, _r1, _r2; };  // :12-25
struct BvhNode { f3 minExtent; int32_t left; f3 maxExtent; int32_t right; };       // :27-50
struct MeshInstance { uint32_t meshIndex, bvhRoot, _r1, _r2; f4 m0, m1, m2, m3; };  // :52-67
struct Intersection { f4 wuvt; uint32_t meshInstance, triIndex, _r1, _r2; };        // :69-83
struct Surface { f3 point, normal; f2 uv; uint32_t matNodeIndex; };                 // :85-97
struct TextureMetadata { uint32_t format, width, height, dataOffset; };             // :99-108
struct MaterialNode {                                                               // :110-165
    uint32_t type, leftChild;
    int32_t rightChild_or_transmittanceTex;
    int32_t tex;  // bumpTex | mixWeightsTex | reflectanceTex | specularityTex | radianceTex
    f3 u2; float _p2;  // reflectance | specularity | radiance | intDispersionIORs | mixWeight(.x)
    f3 u3; float _p3;  // transmittance | extDispersionIORs
    float intIOR, extIOR, scale_or_roughness;
    int32_t roughnessTex;
};
struct Emissive { f4 m0, m1, m2, m3; float area; uint32_t triIndex, matNodeIndex, type; };  // :167-187
static_assert(sizeof(Ray) == 32 && sizeof(Path) == 32 && sizeof(BvhNode) == 32, "layout");
static_assert(sizeof(MeshInstance) == 80 && sizeof(Intersection) == 32, "layout");
static_assert(sizeof(MaterialNode) == 64 && sizeof(Emissive) == 80 && sizeof(TextureMetadata) == 16, "layout");

// bxdf.cl:10-21, material_sampler.cl:4-9, path.cl:4-6, emissive_sampler.cl:4-5, texture_sampler.cl:4-7
constexpr uint32_t BXDF_INVALID = 0, BXDF_TYPE_EMISSIVE = 1u << 1, BXDF_TYPE_DIFFUSE = 1u << 2,
                   BXDF_TYPE_CONDUCTOR = 1u << 3, BXDF_TYPE_ROUGHT_CONDUCTOR = 1u << 4,
                   BXDF_TYPE_DIELECTRIC = 1u << 5, BXDF_TYPE_ROUGH_DIELECTRIC = 1u << 6;
constexpr uint32_t MAT_OP_MIX = 10001, MAT_OP_MIX_MAP = 10002, MAT_OP_BUMP_MAP = 10003,
                   MAT_OP_NORMAL_MAP = 10004, MAT_OP_DISPERSE = 10005;
constexpr uint32_t PATH_FLAG_DISPERSE_R = 1, PATH_FLAG_DISPERSE_G = 2, PATH_FLAG_DISPERSE_B = 4;
constexpr uint32_t EMISSIVE_TYPE_AREA_LIGHT = 0, EMISSIVE_TYPE_ENVIRONMENT_LIGHT = 1;
constexpr uint32_t TEX_FMT_LUMINANCE8 = 0, TEX_FMT_LUMINANCE32F = 1, TEX_FMT_RGBA8 = 2, TEX_FMT_RGBA32F = 3;
inline bool BXDF_IS_EMISSIVE(uint32_t t) { return t == BXDF_TYPE_EMISSIVE; }
inline bool BXDF_IS_SINGULAR(uint32_t t) { return (t & (BXDF_TYPE_CONDUCTOR | BXDF_TYPE_DIELECTRIC)) != 0; }

struct SceneRefs {
    const BvhNode *bvhNodes = nullptr;
    const MeshInstance *meshInstances = nullptr;
    const MaterialNode *materialNodes = nullptr;
    const uint8_t *texData = nullptr;
    const TextureMetadata *texMeta = nullptr;
    const f4 *vertices = nullptr;
    const f4 *normals = nullptr;
    const f2 *uv = nullptr;
    const uint32_t *materialIndices = nullptr;
    const Emissive *emissives = nullptr;
    uint32_t numEmissives = 0;
    int32_t sceneDiffuseMatIndex = -1;
    bool loaded = false;
};

// ------------------------------------------------------------------------------------------
// samplers/random_sampler.cl:7-16
// ------------------------------------------------------------------------------------------
inline f2 randomGetSample2f(u2 *state) {
    const float invMaxInt = 1.0f / 4294967296.0f;
    uint32_t x = state->x * 17u + state->y * 13123u;
    state->x = (x << 13) ^ x;
    state->y ^= (x << 7);
    uint32_t t0 = x * (x * x * 15731u + 74323u) + 871483u;
    uint32_t t1 = x * (x * x * 13734u + 37828u) + 234234u;
    return f2{(float)t0 * invMaxInt, (float)t1 * invMaxInt};  // convert_float2: round to nearest even
}

// ------------------------------------------------------------------------------------------
// util/transform.cl:9-38
// ------------------------------------------------------------------------------------------
inline f3 mul4x1(f3 vec, f4 mat0, f4 mat1, f4 mat2, f4 mat3) {
    f3 out;
    out.x = mat0.x * vec.x + mat1.x * vec.y + mat2.x * vec.z + mat3.x;
    out.y = mat0.y * vec.x + mat1.y * vec.y + mat2.y * vec.z + mat3.y;
    out.z = mat0.z * vec.x + mat1.z * vec.y + mat2.z * vec.z + mat3.z;
    return out;
}
inline f3 mul3x1(f3 vec, f3 mat0, f3 mat1, f3 mat2) {
    f3 out;
    out.x = mat0.x * vec.x + mat1.x * vec.y + mat2.x * vec.z;
    out.y = mat0.y * vec.x + mat1.y * vec.y + mat2.y * vec.z;
    out.z = mat0.z * vec.x + mat1.z * vec.y + mat2.z * vec.z;
    return out;
}
inline f2 rayToLatLongUV(f3 vec) {
    float at2 = atan2f(vec.x, vec.z);
    float r = length(vec);
    return f2{(at2 >= 0.0f ? at2 : (at2 + C_TWO_TIMES_PI)) / C_TWO_TIMES_PI, acosf(vec.y / r) / C_PI};
}

// util/fresnel.cl:8-16 (Schlick)
inline float fresnelForDielectric(float etaI, float etaT, float iDotN) {
    float eta = etaI / etaT;
    float r0 = ((1.0f - eta) * (1.0f - eta)) / ((1.0f + eta) * (1.0f + eta));
    float c = 1.0f - fabsf(iDotN);
    float c1 = c * c;
    return r0 + (1.0f - r0) * c1 * c1 * c;
}

// util/surface.cl:4-6
inline void TANGENT_VECTORS(f3 normal, f3 &u, f3 &v) {
    u = normalize(cross((fabsf(normal.z) < .999f ? F3(0.0f, 0.0f, 1.0f) : F3(1.0f, 0.0f, 0.0f)), normal));
    v = cross(normal, u);
}

// util/surface.cl:12-33
inline void surfaceInit(Surface *surface, const Intersection *intersection, const SceneRefs &sc) {
    f3 wuv = xyz(intersection->wuvt);
    int offset = (int)(intersection->triIndex * 3);
    surface->point = xyz(wuv.x * sc.vertices[offset] + wuv.y * sc.vertices[offset + 1] + wuv.z * sc.vertices[offset + 2]);
    surface->normal = normalize(xyz(wuv.x * sc.normals[offset] + wuv.y * sc.normals[offset + 1] + wuv.z * sc.normals[offset + 2]));
    surface->uv = wuv.x * sc.uv[offset] + wuv.y * sc.uv[offset + 1] + wuv.z * sc.uv[offset + 2];
    surface->matNodeIndex = sc.materialIndices[intersection->triIndex];
}

// ------------------------------------------------------------------------------------------
// samplers/texture_sampler.cl
// ------------------------------------------------------------------------------------------
struct TexTaps { uint32_t tx, ty, bx, by, w; float coeffX, coeffY; };
inline TexTaps texTaps(f2 uv, const TextureMetadata &m) {  // :15-34, :106-125, :188-204
    TexTaps t;
    f2 scaledUV = f2{uv.x - floorf(uv.x), uv.y - floorf(uv.y)};
    scaledUV.x *= (float)m.width;
    scaledUV.y *= (float)m.height;
    t.tx = cl_clamp((uint32_t)scaledUV.x, 0u, m.width - 1);
    t.ty = cl_clamp((uint32_t)scaledUV.y, 0u, m.height - 1);
    t.bx = cl_clamp(t.tx + 1, 0u, m.width - 1);
    t.by = cl_clamp(t.ty + 1, 0u, m.height - 1);
    t.coeffX = scaledUV.x - (float)t.tx;
    t.coeffY = scaledUV.y - (float)t.ty;
    t.w = m.width;
    return t;
}
inline float ldf(const uint8_t *p) { float f; memcpy(&f, p, 4); return f; }
inline f4 ldf4(const uint8_t *p) { f4 f; memcpy(&f, p, 16); return f; }

f3 texGetSample3f(f2 uv, int texIndex, const SceneRefs &sc) {  // :14-101
    const TextureMetadata &m = sc.texMeta[texIndex];
    TexTaps t = texTaps(uv, m);
    const uint8_t *basePtr = sc.texData + m.dataOffset;
    switch (m.format) {
        case TEX_FMT_RGBA8: {
            auto cv = [&](uint32_t i) { const uint8_t *p = basePtr + 4 * (size_t)i; return f4{(float)p[0], (float)p[1], (float)p[2], (float)p[3]}; };
            f4 rgbTL = cv(t.ty * t.w + t.tx), rgbTR = cv(t.ty * t.w + t.bx), rgbBL = cv(t.by * t.w + t.tx), rgbBR = cv(t.by * t.w + t.bx);
            return xyz(mix(mix(rgbTL, rgbBL, t.coeffY), mix(rgbTR, rgbBR, t.coeffY), t.coeffX)) / 255.0f;
        }
        case TEX_FMT_RGBA32F: {
            auto ld = [&](uint32_t i) { return ldf4(basePtr + 16 * (size_t)i); };
            f4 rgbTL = ld(t.ty * t.w + t.tx), rgbTR = ld(t.ty * t.w + t.bx), rgbBL = ld(t.by * t.w + t.tx), rgbBR = ld(t.by * t.w + t.bx);
            return xyz(mix(mix(rgbTL, rgbBL, t.coeffY), mix(rgbTR, rgbBR, t.coeffY), t.coeffX));
        }
        case TEX_FMT_LUMINANCE8: {
            float rTL = (float)basePtr[t.ty * t.w + t.tx], rTR = (float)basePtr[t.ty * t.w + t.bx];
            float rBL = (float)basePtr[t.by * t.w + t.tx], rBR = (float)basePtr[t.by * t.w + t.bx];
            float r = mix(mix(rTL, rBL, t.coeffY), mix(rTR, rBR, t.coeffY), t.coeffX) / 255.0f;
            return F3(r, r, r);
        }
        case TEX_FMT_LUMINANCE32F: {
            auto ld = [&](uint32_t i) { return ldf(basePtr + 4 * (size_t)i); };
            float rTL = ld(t.ty * t.w + t.tx), rTR = ld(t.ty * t.w + t.bx), rBL = ld(t.by * t.w + t.tx), rBR = ld(t.by * t.w + t.bx);
            float r = mix(mix(rTL, rBL, t.coeffY), mix(rTR, rBR, t.coeffY), t.coeffX);
            return F3(r, r, r);
        }
    }
    return F3(0.0f, 0.0f, 0.0f);
}

float texGetSample1f(f2 uv, int texIndex, const SceneRefs &sc) {  // :105-184 (red channel only)
    const TextureMetadata &m = sc.texMeta[texIndex];
    TexTaps t = texTaps(uv, m);
    const uint8_t *basePtr = sc.texData + m.dataOffset;
    switch (m.format) {
        case TEX_FMT_RGBA8: {
            auto ld = [&](uint32_t y, uint32_t x) { return (float)basePtr[((size_t)(y * t.w) << 2) + ((size_t)x << 2)]; };
            float rTL = ld(t.ty, t.tx), rTR = ld(t.ty, t.bx), rBL = ld(t.by, t.tx), rBR = ld(t.by, t.bx);
            return mix(mix(rTL, rBL, t.coeffY), mix(rTR, rBR, t.coeffY), t.coeffX) / 255.0f;
        }
        case TEX_FMT_RGBA32F: {
            auto ld = [&](uint32_t y, uint32_t x) { return ldf(basePtr + 4 * (((size_t)(y * t.w) << 2) + ((size_t)x << 2))); };
            float rTL = ld(t.ty, t.tx), rTR = ld(t.ty, t.bx), rBL = ld(t.by, t.tx), rBR = ld(t.by, t.bx);
            return mix(mix(rTL, rBL, t.coeffY), mix(rTR, rBR, t.coeffY), t.coeffX);
        }
        case TEX_FMT_LUMINANCE8: {
            float rTL = (float)basePtr[t.ty * t.w + t.tx], rTR = (float)basePtr[t.ty * t.w + t.bx];
            float rBL = (float)basePtr[t.by * t.w + t.tx], rBR = (float)basePtr[t.by * t.w + t.bx];
            return mix(mix(rTL, rBL, t.coeffY), mix(rTR, rBR, t.coeffY), t.coeffX) / 255.0f;
        }
        case TEX_FMT_LUMINANCE32F: {
            auto ld = [&](uint32_t i) { return ldf(basePtr + 4 * (size_t)i); };
            float rTL = ld(t.ty * t.w + t.tx), rTR = ld(t.ty * t.w + t.bx), rBL = ld(t.by * t.w + t.tx), rBR = ld(t.by * t.w + t.bx);
            return mix(mix(rTL, rBL, t.coeffY), mix(rTR, rBR, t.coeffY), t.coeffX);
        }
    }
    return 0.0f;
}

f3 texGetBumpSample3f(f2 uv, int texIndex, const SceneRefs &sc) {  // :187-251
    const TextureMetadata &m = sc.texMeta[texIndex];
    TexTaps t = texTaps(uv, m);
    const uint8_t *basePtr = sc.texData + m.dataOffset;
    f3 halfVec = F3(0.5f, 0.5f, 0.5f);
    float s0, s1, s2;
    switch (m.format) {
        case TEX_FMT_RGBA8:
            s0 = (float)basePtr[4 * (size_t)(t.ty * t.w + t.tx)] / 255.0f;
            s1 = (float)basePtr[4 * (size_t)(t.ty * t.w + t.bx)] / 255.0f;
            s2 = (float)basePtr[4 * (size_t)(t.by * t.w + t.tx)] / 255.0f;
            break;
        case TEX_FMT_RGBA32F:
            s0 = ldf(basePtr + 16 * (size_t)(t.ty * t.w + t.tx));
            s1 = ldf(basePtr + 16 * (size_t)(t.ty * t.w + t.bx));
            s2 = ldf(basePtr + 16 * (size_t)(t.by * t.w + t.tx));
            break;
        case TEX_FMT_LUMINANCE8:
            s0 = (float)basePtr[t.ty * t.w + t.tx] / 255.0f;
            s1 = (float)basePtr[t.ty * t.w + t.bx] / 255.0f;
            s2 = (float)basePtr[t.by * t.w + t.tx] / 255.0f;
            break;
        case TEX_FMT_LUMINANCE32F:
            s0 = ldf(basePtr + 4 * (size_t)(t.ty * t.w + t.tx));
            s1 = ldf(basePtr + 4 * (size_t)(t.ty * t.w + t.bx));
            s2 = ldf(basePtr + 4 * (size_t)(t.by * t.w + t.tx));
            break;
        default:
            return F3(0.0f, 0.0f, 0.0f);
    }
    return halfVec + 0.5f * normalize(F3(s1 - s0, s2 - s0, 1.0f));
}

// ------------------------------------------------------------------------------------------
// samplers/material_sampler.cl
// ------------------------------------------------------------------------------------------
inline f3 matGetSample3f(f2 uv, f3 defaultValue, int texIndex, const SceneRefs &sc) {  // :92-98
    if (texIndex == -1) return defaultValue;
    return texGetSample3f(uv, texIndex, sc);
}
inline float matGetSample1f(f2 uv, float defaultValue, int texIndex, const SceneRefs &sc) {  // :102-108
    if (texIndex == -1) return defaultValue;
    return texGetSample1f(uv, texIndex, sc);
}
f3 matGetNormalSample3f(f3 normal, f2 uv, int texIndex, const SceneRefs &sc) {  // :111-121
    f3 u, v;
    TANGENT_VECTORS(normal, u, v);
    f3 sample = (texGetSample3f(uv, texIndex, sc) * 2.0f) - 1.0f;
    return normalize(u * sample.x + v * sample.y + 0.5f * normal * sample.z);
}
f3 matGetBumpSample3f(f3 normal, f2 uv, int texIndex, const SceneRefs &sc) {  // :124-131
    f3 u, v;
    TANGENT_VECTORS(normal, u, v);
    f3 sample = (texGetBumpSample3f(uv, texIndex, sc) * 2.0f) - 1.0f;
    return normalize(u * sample.x + v * sample.y + normal * sample.z);
}

struct MatStats { uint32_t nodesVisited = 0; };

// :21-88
void matSelectNode(Path *path, Surface *surface, f3 inRayDir, MaterialNode *selectedMaterial, f3 *tint,
                   const SceneRefs &sc, u2 *rndState, MatStats *st) {
    (void)inRayDir;
    const MaterialNode *node = sc.materialNodes + surface->matNodeIndex;
    f2 sample;
    f2 forceIOR = f2{0.0f, 0.0f};
    uint32_t flags;
    st->nodesVisited = 1;
    while (node->type >= MAT_OP_MIX) {
        switch (node->type) {
            case MAT_OP_MIX:
                sample = randomGetSample2f(rndState);
                node = sc.materialNodes + (sample.x < node->u2.x ? node->leftChild : (uint32_t)node->rightChild_or_transmittanceTex);
                break;
            case MAT_OP_MIX_MAP:
                sample = randomGetSample2f(rndState);
                sample.y = texGetSample1f(surface->uv, node->tex, sc);
                node = sc.materialNodes + (sample.x < sample.y ? node->leftChild : (uint32_t)node->rightChild_or_transmittanceTex);
                break;
            case MAT_OP_BUMP_MAP:
                surface->normal = matGetBumpSample3f(surface->normal, surface->uv, node->tex, sc);
                node = sc.materialNodes + node->leftChild;
                break;
            case MAT_OP_NORMAL_MAP:
                surface->normal = matGetNormalSample3f(surface->normal, surface->uv, node->tex, sc);
                node = sc.materialNodes + node->leftChild;
                break;
            case MAT_OP_DISPERSE:
                flags = path->flags;
                if ((flags & PATH_FLAG_DISPERSE_R) != 0) {
                    *tint = F3(1.0f, 0.0f, 0.0f);
                    forceIOR = f2{node->u2.x, node->u3.x};
                } else if ((flags & PATH_FLAG_DISPERSE_G) != 0) {
                    *tint = F3(0.0f, 1.0f, 0.0f);
                    forceIOR = f2{node->u2.y, node->u3.y};
                } else if ((flags & PATH_FLAG_DISPERSE_B) != 0) {
                    *tint = F3(0.0f, 0.0f, 1.0f);
                    forceIOR = f2{node->u2.z, node->u3.z};
                } else {
                    sample = randomGetSample2f(rndState);
                    if (sample.x < 0.333f) {
                        *tint = F3(1.0f, 0.0f, 0.0f);
                        forceIOR = f2{node->u2.x, node->u3.x};
                        path->flags |= PATH_FLAG_DISPERSE_R;
                    } else if (sample.x < 0.666f) {
                        *tint = F3(0.0f, 1.0f, 0.0f);
                        forceIOR = f2{node->u2.y, node->u3.y};
                        path->flags |= PATH_FLAG_DISPERSE_G;
                    } else {
                        *tint = F3(0.0f, 0.0f, 1.0f);
                        forceIOR = f2{node->u2.z, node->u3.z};
                        path->flags |= PATH_FLAG_DISPERSE_B;
                    }
                }
                node = sc.materialNodes + node->leftChild;
                break;
            default:  // unknown op: the reference would spin; stop on an invalid leaf instead
                *selectedMaterial = *node;
                selectedMaterial->type = BXDF_INVALID;
                return;
        }
        st->nodesVisited++;
    }
    *selectedMaterial = *node;
    selectedMaterial->intIOR = cl_max(selectedMaterial->intIOR, forceIOR.x);
    selectedMaterial->extIOR = cl_max(selectedMaterial->extIOR, forceIOR.y);
}

// ------------------------------------------------------------------------------------------
// samplers/distribution_sampler.cl
// ------------------------------------------------------------------------------------------
inline float _ggxGetG1(float roughness, f3 v, f3 n, f3 m) {  // :16-29
    float nDotV = dot(n, v);
    float mDotV = dot(m, v);
    if (nDotV * mDotV <= 0.0f) return 0.0f;
    float nDotVSq = nDotV * nDotV;
    float tanSq = nDotVSq > 0.0f ? (1.0f - nDotVSq) / nDotVSq : 0.0f;
    float aSq = roughness * roughness;
    return 2.0f / (1.0f + sqrtf(1.0f + aSq * tanSq));
}
inline float ggxGetG(float roughness, f3 inRayDir, f3 outRayDir, f3 n, f3 m) {  // :33-35
    return _ggxGetG1(roughness, inRayDir, n, m) * _ggxGetG1(roughness, outRayDir, n, m);
}
inline float ggxGetD(float roughness, f3 n, f3 m) {  // :38-52
    float nDotM = dot(n, m);
    if (nDotM <= 0.0f) return 0.0f;
    float nDotMSq = nDotM * nDotM;
    float tanSq = nDotM != 0.0f ? ((1.0f - nDotMSq) / nDotMSq) : 0.0f;
    float aSq = roughness * roughness;
    float denom = C_PI * nDotMSq * nDotMSq * (aSq + tanSq) * (aSq + tanSq);
    return denom > 0.0f ? (aSq / denom) : 0.0f;
}
inline f3 ggxGetSample(float roughness, f3 inRayDir, f3 n, f2 randSample) {  // :55-74
    (void)inRayDir;
    f3 u, v;
    TANGENT_VECTORS(n, u, v);
    float theta = atanf(roughness * sqrtf(randSample.x / (1.0f - randSample.x)));
    theta = theta >= 0.0f ? theta : (theta + C_TWO_TIMES_PI);
    float cosTheta = native_cos(theta);
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    float cosPhi = native_cos(C_TWO_TIMES_PI * randSample.y);
    float sinPhi = sqrtf(1.0f - cosPhi * cosPhi);
    return normalize(u * sinTheta * cosPhi + v * sinTheta * sinPhi + n * cosTheta);
}
inline float ggxGetReflectionPdf(float roughness, f3 inRayDir, f3 outRayDir, f3 n, f3 h) {  // :76-85
    (void)inRayDir;
    float nDotH = fabsf(dot(n, h));
    float oDotH = fabsf(dot(outRayDir, h));
    float denom = 4.0f * oDotH;
    return denom == 0.0f ? 0.0f : ggxGetD(roughness, n, h) * nDotH / denom;
}
inline float ggxGetRefractionPdf(float roughness, float etaI, float etaT, f3 inRayDir, f3 outRayDir, f3 n, f3 h) {  // :87-96
    float iDotH = fabsf(dot(inRayDir, h));
    float oDotH = fabsf(dot(outRayDir, h));
    float hDotN = fabsf(dot(h, n));
    float denom = (etaI * iDotH + etaT * oDotH) * (etaI * iDotH + etaT * oDotH);
    return denom > 0.0f ? ggxGetD(roughness, n, h) * hDotN * oDotH * etaT * etaT / denom : 0.0f;
}
inline f3 cosWeightedHemisphereGetSample(f3 normal, f2 randSample) {  // :101-112
    float rd = sqrtf(randSample.x);
    float phi = C_TWO_TIMES_PI * randSample.y;
    f3 u, v;
    TANGENT_VECTORS(normal, u, v);
    return normalize(u * rd * native_cos(phi) + v * rd * native_sin(phi) + normal * native_sqrt(1 - randSample.x));
}

// ------------------------------------------------------------------------------------------
// bxdf/*.cl
// ------------------------------------------------------------------------------------------
inline int reflTex(const MaterialNode *m) { return m->tex; }
inline int transTex(const MaterialNode *m) { return m->rightChild_or_transmittanceTex; }

// diffuse.cl:12-32
f3 diffuseSample(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f2 randSample, f3 *rayOutDir, float *pdf) {
    *rayOutDir = cosWeightedHemisphereGetSample(surface->normal, randSample);
    *pdf = dot(surface->normal, *rayOutDir) * C_1_PI;
    f3 kd = matGetSample3f(surface->uv, matNode->u2, reflTex(matNode), sc);
    return kd * C_1_PI;
}
float diffusePdf(Surface *surface, f3 rayOutDir) { return dot(surface->normal, rayOutDir) * C_1_PI; }
f3 diffuseEval(Surface *surface, MaterialNode *matNode, const SceneRefs &sc) {
    f3 kd = matGetSample3f(surface->uv, matNode->u2, reflTex(matNode), sc);
    return kd * C_1_PI;
}

// conductor.cl:12-62
f3 conductorSample(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f3 inRayDir, f3 *outRayDir, float *pdf) {
    float iDotN = dot(inRayDir, surface->normal);
    *outRayDir = 2.0f * iDotN * surface->normal - inRayDir;
    *pdf = 1.0f;
    float f = matNode->intIOR != 0.0f ? fresnelForDielectric(matNode->extIOR, matNode->intIOR, iDotN) : 1.0f;
    f3 ks = matGetSample3f(surface->uv, matNode->u2, reflTex(matNode), sc);
    return iDotN != 0.0f ? f * ks / iDotN : F3(0.0f);
}
float conductorPdf(Surface *surface, f3 inRayDir, f3 outRayDir) {
    float iDotN = dot(inRayDir, surface->normal);
    f3 expOutDir = 2.0f * iDotN * surface->normal - inRayDir;
    float expDot = dot(expOutDir, outRayDir);
    return expDot >= 0.0f && expDot <= 0.001f ? 1.0f : 0.0f;  // SURVEY Q9, as written
}
f3 conductorEval(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f3 inRayDir, f3 outRayDir) {
    float iDotN = dot(inRayDir, surface->normal);
    f3 expOutDir = 2.0f * iDotN * surface->normal - inRayDir;
    float expDot = dot(expOutDir, outRayDir);
    if (expDot < 0.0f || expDot > 0.001f) return F3(0.0f, 0.0f, 0.0f);
    float f = matNode->intIOR != 0.0f ? fresnelForDielectric(matNode->extIOR, matNode->intIOR, iDotN) : 1.0f;
    f3 ks = matGetSample3f(surface->uv, matNode->u2, reflTex(matNode), sc);
    return iDotN != 0.0f ? f * ks / iDotN : F3(0.0f);
}

// dielectric.cl:12-61
f3 dielecticSample(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f2 randSample, f3 inRayDir, f3 *outRayDir, float *pdf) {
    float iDotN = dot(inRayDir, surface->normal);
    float etaI = matNode->extIOR;
    float etaT = matNode->intIOR;
    if (iDotN < 0.0f) { float tmp = etaI; etaI = etaT; etaT = tmp; }
    float eta = etaI / etaT;
    float f = fresnelForDielectric(etaI, etaT, iDotN);
    f3 kVal;
    float cosTSq = 1.0f + eta * (iDotN * iDotN - 1.0f);  // eta, not eta^2
    if (cosTSq <= 0.0f || randSample.x <= f) {
        *outRayDir = -cl_sign(iDotN) * 2.0f * iDotN * surface->normal - inRayDir;
        kVal = matGetSample3f(surface->uv, matNode->u2, reflTex(matNode), sc);
        *pdf = cosTSq <= 0.0f ? 1.0f : f;
    } else {
        *outRayDir = (eta * iDotN - cl_sign(iDotN) * sqrtf(cosTSq)) * surface->normal - eta * inRayDir;
        kVal = eta * eta * matGetSample3f(surface->uv, matNode->u3, transTex(matNode), sc);
        *pdf = 1.0f - f;
    }
    return iDotN != 0.0f ? *pdf * kVal / fabsf(iDotN) : F3(0.0f);
}

inline float roughnessOf(Surface *surface, MaterialNode *matNode, const SceneRefs &sc) {
    float roughness = cl_clamp(matGetSample1f(surface->uv, matNode->scale_or_roughness, matNode->roughnessTex, sc), MIN_ROUGHNESS, 1.0f);
    roughness *= roughness;
    return roughness;
}

// rough_conductor.cl:9-78
f3 roughConductorSample(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f2 randSample, f3 inRayDir, f3 *outRayDir, float *pdf) {
    float roughness = roughnessOf(surface, matNode, sc);
    f3 ks = matGetSample3f(surface->uv, matNode->u2, reflTex(matNode), sc);
    f3 h = ggxGetSample(roughness, inRayDir, surface->normal, randSample);
    *outRayDir = 2.0f * dot(inRayDir, h) * h - inRayDir;
    *pdf = ggxGetReflectionPdf(roughness, inRayDir, *outRayDir, surface->normal, h);
    float iDotN = dot(inRayDir, surface->normal);
    float oDotN = dot(*outRayDir, surface->normal);
    h = normalize(inRayDir + *outRayDir);
    float d = ggxGetD(roughness, surface->normal, h);
    float g = ggxGetG(roughness, inRayDir, *outRayDir, surface->normal, h);
    float f = matNode->intIOR != 0.0f ? fresnelForDielectric(matNode->extIOR, matNode->intIOR, iDotN) : 1.0f;
    float denom = 4.0f * iDotN * oDotN;
    return denom > 0.0f ? ks * f * d * g / denom : F3(0.0f);
}
float roughConductorPdf(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f3 inRayDir, f3 outRayDir) {
    float roughness = roughnessOf(surface, matNode, sc);
    f3 h = normalize(inRayDir + outRayDir);
    return ggxGetReflectionPdf(roughness, inRayDir, outRayDir, surface->normal, h);
}
f3 roughConductorEval(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f3 inRayDir, f3 outRayDir) {
    float roughness = roughnessOf(surface, matNode, sc);
    f3 ks = matGetSample3f(surface->uv, matNode->u2, reflTex(matNode), sc);
    float iDotN = dot(inRayDir, surface->normal);
    float oDotN = dot(outRayDir, surface->normal);
    float f = matNode->intIOR != 0.0f ? fresnelForDielectric(matNode->extIOR, matNode->intIOR, iDotN) : 1.0f;
    f3 h = normalize(inRayDir + outRayDir);
    float d = ggxGetD(roughness, surface->normal, h);
    float g = ggxGetG(roughness, inRayDir, outRayDir, surface->normal, h);
    float denom = 4.0f * iDotN * oDotN;
    return denom > 0.0f ? ks * f * d * g / denom : F3(0.0f);
}

// rough_dielectric.cl:9-166
f3 roughDielectricSample(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f2 randSample, f3 inRayDir, f3 *outRayDir, float *pdf) {
    float iDotN = dot(inRayDir, surface->normal);
    float roughness = roughnessOf(surface, matNode, sc);
    float etaI = matNode->extIOR;
    float etaT = matNode->intIOR;
    if (iDotN < 0.0f) { float tmp = etaI; etaI = etaT; etaT = tmp; }
    float eta = etaI / etaT;
    f3 h = ggxGetSample(roughness, inRayDir, surface->normal, randSample);
    float f = fresnelForDielectric(etaI, etaT, iDotN);
    float cosTSq = 1.0f + eta * (iDotN * iDotN - 1.0f);
    if (cosTSq <= 0.0f || randSample.x <= f) {
        *outRayDir = 2.0f * dot(inRayDir, h) * h - inRayDir;
        f3 ks = matGetSample3f(surface->uv, matNode->u2, reflTex(matNode), sc);
        float iDotN2 = dot(inRayDir, surface->normal);
        float oDotN = dot(*outRayDir, surface->normal);
        h = normalize(inRayDir + *outRayDir);
        *pdf = cosTSq <= 0.0f ? 1.0f : ggxGetReflectionPdf(roughness, inRayDir, *outRayDir, surface->normal, h);
        float d = ggxGetD(roughness, surface->normal, h);
        float g = ggxGetG(roughness, inRayDir, *outRayDir, surface->normal, h);
        float denom = 4.0f * iDotN2 * oDotN;
        return denom > 0.0f ? ks * f * d * g / denom : F3(0.0f);
    }
    *outRayDir = (eta * iDotN - cl_sign(iDotN) * sqrtf(cosTSq)) * h - eta * inRayDir;
    h = normalize(-(etaI * inRayDir + etaT * *outRayDir));
    *pdf = ggxGetRefractionPdf(roughness, etaI, etaT, inRayDir, *outRayDir, surface->normal, h);
    float iDotH = fabsf(dot(inRayDir, h));
    float oDotH = fabsf(dot(*outRayDir, h));
    float oDotN = dot(*outRayDir, surface->normal);
    float focusTermDenom = iDotN * oDotN * (etaI * iDotH + etaT * oDotH) * (etaI * iDotH + etaT * oDotH);
    if (focusTermDenom == 0.0f) return F3(0.0f, 0.0f, 0.0f);
    float focusTerm = fabsf(etaT * etaT * iDotH * oDotH / focusTermDenom);
    float d = ggxGetD(roughness, surface->normal, h);
    float g = ggxGetG(roughness, inRayDir, *outRayDir, surface->normal, h);
    f3 tf = matGetSample3f(surface->uv, matNode->u3, transTex(matNode), sc);
    return tf * (1.0f - f) * d * g * focusTerm;
}
float roughDielectricPdf(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f3 inRayDir, f3 outRayDir) {
    float iDotN = dot(inRayDir, surface->normal);
    float roughness = roughnessOf(surface, matNode, sc);
    if (iDotN > 0.0f) {
        f3 h = normalize(inRayDir + outRayDir);
        return ggxGetReflectionPdf(roughness, inRayDir, outRayDir, surface->normal, h);
    }
    float etaI = matNode->extIOR;
    float etaT = matNode->intIOR;
    if (iDotN < 0.0f) { float tmp = etaI; etaI = etaT; etaT = tmp; }
    f3 h = normalize(-(etaI * inRayDir + etaT * outRayDir));
    return ggxGetRefractionPdf(roughness, etaI, etaT, inRayDir, outRayDir, surface->normal, h);
}
f3 roughDielectricEval(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f3 inRayDir, f3 outRayDir) {
    float iDotN = dot(inRayDir, surface->normal);
    float oDotN = dot(outRayDir, surface->normal);
    float roughness = roughnessOf(surface, matNode, sc);
    float etaI = matNode->extIOR;
    float etaT = matNode->intIOR;
    if (iDotN < 0.0f) { float tmp = etaI; etaI = etaT; etaT = tmp; }
    float f = fresnelForDielectric(etaI, etaT, iDotN);
    if (iDotN > 0.0f) {
        f3 ks = matGetSample3f(surface->uv, matNode->u2, reflTex(matNode), sc);
        f3 h = normalize(inRayDir + outRayDir);
        float d = ggxGetD(roughness, surface->normal, h);
        float g = ggxGetG(roughness, inRayDir, outRayDir, surface->normal, h);
        float denom = 4.0f * iDotN * oDotN;
        return denom > 0.0f ? ks * f * d * g / denom : F3(0.0f);
    }
    f3 h = normalize(-(etaI * inRayDir + etaT * outRayDir));
    float iDotH = fabsf(dot(inRayDir, h));
    float oDotH = fabsf(dot(outRayDir, h));
    float focusTermDenom = iDotN * oDotN * (etaI * iDotH + etaT * oDotH) * (etaI * iDotH + etaT * oDotH);
    if (focusTermDenom == 0.0f) return F3(0.0f, 0.0f, 0.0f);
    float focusTerm = fabsf(etaT * etaT * iDotH * oDotH / focusTermDenom);
    float d = ggxGetD(roughness, surface->normal, h);
    float g = ggxGetG(roughness, inRayDir, outRayDir, surface->normal, h);
    f3 tf = matGetSample3f(surface->uv, matNode->u3, transTex(matNode), sc);
    return tf * (1.0f - f) * d * g * focusTerm;
}

// bxdf.cl:29-105
f3 bxdfGetSample(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f2 randSample, f3 inRayDir, f3 *outRayDir, float *pdf) {
    switch (matNode->type) {
        case BXDF_TYPE_DIFFUSE: return diffuseSample(surface, matNode, sc, randSample, outRayDir, pdf);
        case BXDF_TYPE_CONDUCTOR: return conductorSample(surface, matNode, sc, inRayDir, outRayDir, pdf);
        case BXDF_TYPE_DIELECTRIC: return dielecticSample(surface, matNode, sc, randSample, inRayDir, outRayDir, pdf);
        case BXDF_TYPE_ROUGHT_CONDUCTOR: return roughConductorSample(surface, matNode, sc, randSample, inRayDir, outRayDir, pdf);
        case BXDF_TYPE_ROUGH_DIELECTRIC: return roughDielectricSample(surface, matNode, sc, randSample, inRayDir, outRayDir, pdf);
    }
    return F3(0.0f, 0.0f, 0.0f);
}
float bxdfGetPdf(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f3 inRayDir, f3 outRayDir) {
    switch (matNode->type) {
        case BXDF_TYPE_DIFFUSE: return diffusePdf(surface, outRayDir);
        case BXDF_TYPE_CONDUCTOR: return conductorPdf(surface, inRayDir, outRayDir);
        case BXDF_TYPE_DIELECTRIC: return 0.0f;  // dielectric.cl:51-54 "cheat and always return 0"
        case BXDF_TYPE_ROUGHT_CONDUCTOR: return roughConductorPdf(surface, matNode, sc, inRayDir, outRayDir);
        case BXDF_TYPE_ROUGH_DIELECTRIC: return roughDielectricPdf(surface, matNode, sc, inRayDir, outRayDir);
    }
    return 0.0f;
}
f3 bxdfEval(Surface *surface, MaterialNode *matNode, const SceneRefs &sc, f3 inRayDir, f3 outRayDir) {
    switch (matNode->type) {
        case BXDF_TYPE_DIFFUSE: return diffuseEval(surface, matNode, sc);
        case BXDF_TYPE_CONDUCTOR: return conductorEval(surface, matNode, sc, inRayDir, outRayDir);
        case BXDF_TYPE_DIELECTRIC: return F3(0.0f, 0.0f, 0.0f);  // dielectric.cl:59-61
        case BXDF_TYPE_ROUGHT_CONDUCTOR: return roughConductorEval(surface, matNode, sc, inRayDir, outRayDir);
        case BXDF_TYPE_ROUGH_DIELECTRIC: return roughDielectricEval(surface, matNode, sc, inRayDir, outRayDir);
    }
    return F3(0.0f, 0.0f, 0.0f);
}

// ------------------------------------------------------------------------------------------
// samplers/emissive_sampler.cl
// ------------------------------------------------------------------------------------------
f3 environmentLightGetSample(Surface *surface, const Emissive *emissive, const SceneRefs &sc, f2 randSample,
                             f3 *outRayDir, float *pdf, float *distToEmissive) {  // :16-37
    *outRayDir = cosWeightedHemisphereGetSample(surface->normal, randSample);
    *pdf = cl_max(0.0f, dot(surface->normal, *outRayDir)) * C_1_PI;
    *distToEmissive = FLT_MAX;
    f2 uv = rayToLatLongUV(*outRayDir);
    MaterialNode matNode = sc.materialNodes[emissive->matNodeIndex];
    return matNode.scale_or_roughness * matGetSample3f(uv, matNode.u2, matNode.tex, sc) * C_1_PI;
}
float environmentLightGetPdf(Surface *surface, f3 outRayDir) {  // :39-47
    return cl_max(0.0f, dot(surface->normal, outRayDir) * C_1_PI);
}
f3 areaLightGetSample(Surface *surface, const Emissive *emissive, const SceneRefs &sc, f2 randSample,
                      f3 *outRayDir, float *pdf, float *distToEmissive) {  // :51-113
    float r1sqrt = native_sqrt(randSample.x);
    float ru = (1.0f - randSample.y) * r1sqrt;
    float rv = randSample.y * r1sqrt;
    f3 wuv = F3(1.0f - ru - rv, ru, rv);
    int offset = (int)(emissive->triIndex * 3);
    f3 emissivePoint = mul4x1(xyz(wuv.x * sc.vertices[offset] + wuv.y * sc.vertices[offset + 1] + wuv.z * sc.vertices[offset + 2]),
                              emissive->m0, emissive->m1, emissive->m2, emissive->m3);
    f3 emissiveNormal = mul4x1(xyz(wuv.x * sc.normals[offset] + wuv.y * sc.normals[offset + 1] + wuv.z * sc.normals[offset + 2]),
                               emissive->m0, emissive->m1, emissive->m2, emissive->m3);
    f2 emissiveUV = wuv.x * sc.uv[offset] + wuv.y * sc.uv[offset + 1] + wuv.z * sc.uv[offset + 2];
    MaterialNode matNode = sc.materialNodes[emissive->matNodeIndex];
    f3 emissiveRay = emissivePoint - surface->point;
    float squaredDistToLight = dot(emissiveRay, emissiveRay);
    *outRayDir = normalize(emissiveRay);
    *distToEmissive = native_sqrt(squaredDistToLight);
    float nDotOutRay = dot(emissiveNormal, -*outRayDir);
    if (nDotOutRay > 0.0f) {
        *pdf = 1.0f / emissive->area;
        f3 ke = matGetSample3f(emissiveUV, matNode.u2, matNode.tex, sc);
        return matNode.scale_or_roughness * ke * nDotOutRay / squaredDistToLight;
    }
    *pdf = 0.0f;
    return F3(0.0f, 0.0f, 0.0f);
}
float areaLightGetPdf(Surface *surface, const Emissive *emissive, const SceneRefs &sc, f3 outRayDir) {  // :117-173
    int offset = (int)(emissive->triIndex * 3);
    f3 v0 = xyz(sc.vertices[offset]);
    f3 edge01 = xyz(sc.vertices[offset + 1]) - v0;
    f3 edge02 = xyz(sc.vertices[offset + 2]) - v0;
    v0 = mul4x1(v0, emissive->m0, emissive->m1, emissive->m2, emissive->m3);
    edge01 = mul4x1(edge01, emissive->m0, emissive->m1, emissive->m2, emissive->m3);
    edge02 = mul4x1(edge02, emissive->m0, emissive->m1, emissive->m2, emissive->m3);
    f3 pVec = cross(outRayDir, edge02);
    float det = dot(edge01, pVec);
    if (fabsf(det) < INTERSECTION_EPSILON) return 0.0f;
    float invDet = native_recip(det);
    f3 tVec = surface->point - v0;
    float u = dot(tVec, pVec) * invDet;
    if (u < 0.0f || u > 1.0f) return 0.0f;
    f3 qVec = cross(tVec, edge01);
    float v = dot(outRayDir, qVec) * invDet;
    if (v < 0.0f || u + v > 1.0f) return 0.0f;
    float t = dot(edge02, qVec) * invDet;
    if (t < INTERSECTION_EPSILON) return 0.0f;
    f3 emissiveNormal = normalize(cross(edge01, edge02));
    float denominator = emissive->area * fabsf(dot(emissiveNormal, outRayDir));
    return denominator > 0.0f ? (t * t) / denominator : 0.0f;
}
f3 emissiveGetSample(Surface *surface, const Emissive *emissive, const SceneRefs &sc, f2 randSample,
                     f3 *outRayDir, float *pdf, float *distToEmissive) {  // :178-201
    switch (emissive->type) {
        case EMISSIVE_TYPE_AREA_LIGHT: return areaLightGetSample(surface, emissive, sc, randSample, outRayDir, pdf, distToEmissive);
        case EMISSIVE_TYPE_ENVIRONMENT_LIGHT: return environmentLightGetSample(surface, emissive, sc, randSample, outRayDir, pdf, distToEmissive);
    }
    return F3(0.0f, 0.0f, 0.0f);
}
float emissiveGetPdf(Surface *surface, const Emissive *emissive, const SceneRefs &sc, f3 outRayDir) {  // :204-224
    switch (emissive->type) {
        case EMISSIVE_TYPE_AREA_LIGHT: return areaLightGetPdf(surface, emissive, sc, outRayDir);
        case EMISSIVE_TYPE_ENVIRONMENT_LIGHT: return environmentLightGetPdf(surface, outRayDir);
    }
    return 0.0f;
}
inline uint32_t emissiveSelect(const int numLights, float randSample, float *pdf) {  // :227-237
    *pdf = native_recip((float)numLights);
    return (uint32_t)cl_clamp((int)(randSample * numLights), 0, numLights - 1);
}

// ------------------------------------------------------------------------------------------
// kernels/intersect.cl -- the traversal state machine of SURVEY appendix B
// ------------------------------------------------------------------------------------------
constexpr int BVH_MAX_STACK_SIZE = 256;  // reference: 32, unchecked (intersect.cl:4)

struct TraverseCounters { uint64_t nodes = 0, tris = 0, instances = 0; };

inline bool slabWant(const BvhNode &c, f3 o, f3 invDir, float tmaxRay) {  // intersect.cl:302-309
    f3 tmin = (c.minExtent - o) * invDir;
    f3 tmax = (c.maxExtent - o) * invDir;
    f3 rmin = fmin3(tmin, tmax);
    f3 rmax = fmax3(tmin, tmax);
    float minmax = fminf(fminf(rmax.x, rmax.y), rmax.z);
    float maxmin = fmaxf(fmaxf(rmin.x, rmin.y), rmin.z);
    float hitDist = minmax < 0 || maxmin > minmax ? FLT_MAX : (maxmin >= tmaxRay ? FLT_MAX : maxmin);
    return hitDist < FLT_MAX;
}

// anyHit == false: rayIntersectionQuery (intersect.cl:184-347)
// anyHit == true : rayIntersectionTest  (intersect.cl:26-180)
template <bool anyHit>
int traverse(const Ray &rayIn, const SceneRefs &sc, Intersection *out, TraverseCounters *cnt) {
    uint32_t nodeStack[BVH_MAX_STACK_SIZE];
    BvhNode curNode, childNodes[2];
    int meshInstanceId = 0;
    Ray ray = rayIn;
    f3 origRayOrigin = xyz(ray.origin);
    f3 origRayDir = xyz(ray.dir);
    f3 o = origRayOrigin, d = origRayDir;

    Intersection intersection;
    memset(&intersection, 0, sizeof(intersection));
    intersection.wuvt.w = ray.origin.w;

    int stackIndex = 0;
    int meshBvhStackStartIndex = -1;
    curNode = sc.bvhNodes[0];
    int wantLeft, wantRight;
    int gotHit = 0;
    while (stackIndex > -1) {
        if (curNode.left <= 0) {  // BVH_IS_LEAF
            int numTriangles = curNode.right;
            if (numTriangles == 0) {
                meshInstanceId = -curNode.left;
                const MeshInstance &meshInstance = sc.meshInstances[meshInstanceId];
                meshBvhStackStartIndex = stackIndex;
                if (stackIndex >= BVH_MAX_STACK_SIZE) { fprintf(stderr, "oracle: BVH stack overflow\n"); abort(); }
                nodeStack[stackIndex++] = meshInstance.bvhRoot;
                o = mul4x1(o, meshInstance.m0, meshInstance.m1, meshInstance.m2, meshInstance.m3);
                d = mul3x1(d, xyz(meshInstance.m0), xyz(meshInstance.m1), xyz(meshInstance.m2));
                cnt->instances++;
            } else {
                int triStartIndex = -curNode.left;
                for (int vIndex = triStartIndex * 3; vIndex < (triStartIndex + numTriangles) * 3; vIndex += 3) {
                    cnt->tris++;
                    f3 v0 = xyz(sc.vertices[vIndex]);
                    f3 edge01 = xyz(sc.vertices[vIndex + 1]) - v0;
                    f3 edge02 = xyz(sc.vertices[vIndex + 2]) - v0;
                    f3 pVec = cross(d, edge02);
                    float det = dot(edge01, pVec);
                    if (fabsf(det) < INTERSECTION_EPSILON) continue;
                    float invDet = native_recip(det);
                    f3 tVec = o - v0;
                    float u = dot(tVec, pVec) * invDet;
                    if (u < 0.0f || u > 1.0f) continue;
                    f3 qVec = cross(tVec, edge01);
                    float v = dot(d, qVec) * invDet;
                    if (v < 0.0f || u + v > 1.0f) continue;
                    float t = dot(edge02, qVec) * invDet;
                    if (anyHit) {
                        if (t > INTERSECTION_EPSILON && t < ray.origin.w) {  // :120-124
                            gotHit = 1;
                            stackIndex = -1;
                            break;
                        }
                    } else if (t > INTERSECTION_EPSILON && t < intersection.wuvt.w) {  // :281-290
                        intersection.wuvt = f4{1.0f - (u + v), u, v, t};
                        intersection.triIndex = (uint32_t)(vIndex / 3);
                        intersection.meshInstance = (uint32_t)meshInstanceId;
                    }
                }
            }
            wantLeft = 0;
            wantRight = 0;
        } else {
            cnt->nodes++;
            childNodes[0] = sc.bvhNodes[curNode.left];
            childNodes[1] = sc.bvhNodes[curNode.right];
            f3 invDir = native_recip(d);
            wantLeft = slabWant(childNodes[0], o, invDir, ray.origin.w) ? 1 : 0;
            wantRight = slabWant(childNodes[1], o, invDir, ray.origin.w) ? 1 : 0;
        }
        if (wantLeft && wantRight) {  // :324-326: always left first
            if (stackIndex >= BVH_MAX_STACK_SIZE) { fprintf(stderr, "oracle: BVH stack overflow\n"); abort(); }
            nodeStack[stackIndex++] = (uint32_t)curNode.right;
            curNode = childNodes[0];
        } else if (wantLeft || wantRight) {
            curNode = wantLeft ? childNodes[0] : childNodes[1];
        } else {
            if (stackIndex == meshBvhStackStartIndex) {  // :330-335
                o = origRayOrigin;
                d = origRayDir;
                meshBvhStackStartIndex = -1;
            }
            if (--stackIndex >= 0) curNode = sc.bvhNodes[nodeStack[stackIndex]];
        }
    }
    if (anyHit) return gotHit;
    *out = intersection;
    return intersection.wuvt.w < ray.origin.w ? 1 : 0;  // :345
}

// ------------------------------------------------------------------------------------------
// The tracer: buffers of tracer/opencl/buffers.go:21-70 + the host loops.
// ------------------------------------------------------------------------------------------
struct Oracle {
    SceneRefs sc;
    uint32_t frameW = 0, frameH = 0;
    f3 eye{0, 0, 0};
    f4 frustrum[4]{};
    std::vector<Ray> rays[3];
    std::vector<Path> paths;
    std::vector<uint32_t> hitFlags;
    std::vector<Intersection> intersections;
    std::vector<f4> emissiveSamples, traceAcc, frameAcc;
    std::vector<uint8_t> frameBuffer, debugOutput;
    int32_t numRays[3] = {0, 0, 0};
    int fixQ4 = 1;
    pc_stats stats{};
    // shadeHits scratch (phase 1 results, see shadeHits below)
    std::vector<uint8_t> wantOcc, wantInd;
    std::vector<Ray> occRay, indRay;
    std::vector<f4> occSample;
};

// kernels/camera.cl:5-58
void generatePrimaryRays(Oracle &o, const pc_block_request &req, uint32_t randSeed) {
    const f4 frustrumTL = o.frustrum[0], frustrumTR = o.frustrum[1], frustrumBL = o.frustrum[2], frustrumBR = o.frustrum[3];
    const f2 texelDims = f2{1.0f / (float)req.frame_w, 1.0f / (float)req.frame_h};  // resources.go:130-133
    const uint32_t frameW = req.frame_w, blockH = req.block_h, blockY = req.block_y;
    o.numRays[0] = (int32_t)(frameW * blockH);
#pragma omp parallel for schedule(static)
    for (int64_t gy = 0; gy < (int64_t)blockH; gy++) {
        for (uint32_t gx = 0; gx < frameW; gx++) {
            uint32_t index = ((uint32_t)gy * frameW) + gx;
            uint32_t pixelIndex = (((uint32_t)gy + blockY) * frameW) + gx;
            u2 rndState = u2{gx + randSeed, (uint32_t)gy + randSeed};
            f2 sample0 = randomGetSample2f(&rndState);
            f2 offset = f2{
                sample0.x < 0.5f ? native_sqrt(2.0f * sample0.x) - 0.5f : 1.5f - native_sqrt(2.0f - 2.0f * sample0.x),
                sample0.y < 0.5f ? native_sqrt(2.0f * sample0.y) - 0.5f : 1.5f - native_sqrt(2.0f - 2.0f * sample0.y)};
            f2 texel = (f2{(float)gx, (float)((uint32_t)gy + blockY)} + offset) * texelDims;
            f4 dir = normalize(mix(mix(frustrumTL, frustrumBL, texel.y), mix(frustrumTR, frustrumBR, texel.y), texel.x));
            // rayNew (util/ray.cl:13-16), pathNew (util/path.cl:13-17)
            o.rays[0][index].origin = f4{o.eye.x, o.eye.y, o.eye.z, FLT_MAX};
            o.rays[0][index].dir = f4{dir.x, dir.y, dir.z, (float)index};
            Path &p = o.paths[index];
            p.throughput = F3(1.0f, 1.0f, 1.0f);
            p.pixelIndex = pixelIndex;
            p.flags = 0;
        }
    }
}

void rayIntersectionQuery(Oracle &o, int buf) {  // resources.go:182-199
    const int n = o.numRays[buf];
    uint64_t nodes = 0, tris = 0, inst = 0, missed = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nodes, tris, inst, missed)
    for (int i = 0; i < n; i++) {
        TraverseCounters c;
        o.hitFlags[i] = (uint32_t)traverse<false>(o.rays[buf][i], o.sc, &o.intersections[i], &c);
        nodes += c.nodes; tris += c.tris; inst += c.instances;
        missed += o.hitFlags[i] ? 0 : 1;
    }
    o.stats.query_rays += (uint64_t)n;
    o.stats.nodes_tested += nodes; o.stats.tris_tested += tris; o.stats.instances_entered += inst;
    o.stats.missed_query_rays += missed;
}

void rayIntersectionTest(Oracle &o, int buf) {  // resources.go:162-178
    const int n = o.numRays[buf];
    uint64_t nodes = 0, tris = 0, inst = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nodes, tris, inst)
    for (int i = 0; i < n; i++) {
        TraverseCounters c;
        o.hitFlags[i] = (uint32_t)traverse<true>(o.rays[buf][i], o.sc, nullptr, &c);
        nodes += c.nodes; tris += c.tris; inst += c.instances;
    }
    o.stats.occlusion_rays += (uint64_t)n;
    o.stats.nodes_tested += nodes; o.stats.tris_tested += tris; o.stats.instances_entered += inst;
}

// kernels/pt_integrator.cl:214-275; primary == true is shadePrimaryRayMisses
void shadeMisses(Oracle &o, int buf, bool primary) {
    const int n = o.numRays[buf];
#pragma omp parallel for schedule(static)
    for (int globalId = 0; globalId < n; globalId++) {
        if (o.hitFlags[globalId]) continue;
        MaterialNode matNode = o.sc.materialNodes[o.sc.sceneDiffuseMatIndex];
        const Ray &r = o.rays[buf][globalId];
        uint32_t rayPathIndex = (uint32_t)r.dir.w;
        f2 uv = rayToLatLongUV(xyz(r.dir));
        f3 kd = matGetSample3f(uv, matNode.u2, matNode.tex, o.sc);
        f4 &acc = o.traceAcc[o.paths[rayPathIndex].pixelIndex];
        f3 add = primary ? kd : o.paths[rayPathIndex].throughput * kd;
        acc.x += add.x; acc.y += add.y; acc.z += add.z;
    }
}

// kernels/pt_integrator.cl:17-211.  Phase 1 is the body of the kernel per work item; phase 2
// replaces the local/global atomics (:162,176,188-197) by a prefix in ray-index order.
void shadeHits(Oracle &o, int a, uint32_t bounce, uint32_t minBouncesForRR, uint32_t randSeed) {
    const int n = o.numRays[a];
    const SceneRefs &sc = o.sc;
    const uint32_t numEmissives = sc.numEmissives;
    o.numRays[2] = 0;       // resources.go:230-238
    o.numRays[1 - a] = 0;
    uint64_t shaded = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : shaded)
    for (int globalId = 0; globalId < n; globalId++) {
        o.wantOcc[globalId] = 0;
        o.wantInd[globalId] = 0;
        if (!o.hitFlags[globalId]) continue;
        shaded++;
        Surface surface;
        uint32_t rayPathIndex;
        f3 outBxdfRayOrigin, outEmissiveRayOrigin;
        f3 curPathThroughput;
        f3 bxdfTint = F3(1.0f, 1.0f, 1.0f);
        f3 bxdfOutRayDir = F3(0.0f), bxdfSample, bxdfEmissiveSample, emissiveOutRayDir = F3(0.0f), emissiveSample = F3(0.0f);
        float bxdfPdf, bxdfEmissivePdf, emissivePdf = 0.0f, emissiveBxdfPdf, emissiveSelectionPdf = 0.0f;
        float emissiveWeight = 0.0f, bxdfWeight, distToEmissive = 0.0f;

        bxdfPdf = 1.0f;
        bxdfWeight = 1.0f;
        u2 rndState = u2{randSeed, (uint32_t)globalId};
        f2 sample0 = randomGetSample2f(&rndState);
        f2 sample1 = randomGetSample2f(&rndState);
        f2 sample2 = randomGetSample2f(&rndState);

        const Ray &inRay = o.rays[a][globalId];
        rayPathIndex = (uint32_t)inRay.dir.w;  // rayGetDirAndPathIndex (util/ray.cl:19-23)
        f3 inRayDir = -xyz(inRay.dir);
        curPathThroughput = o.paths[rayPathIndex].throughput;

        surfaceInit(&surface, &o.intersections[globalId], sc);

        MaterialNode materialNode;
        MatStats ms;
        matSelectNode(&o.paths[rayPathIndex], &surface, inRayDir, &materialNode, &bxdfTint, sc, &rndState, &ms);

        float inRayDotNormal = dot(inRayDir, surface.normal);
        if (BXDF_IS_EMISSIVE(materialNode.type)) {
            if (inRayDotNormal > 0.0f) {
                f3 add = curPathThroughput * materialNode.scale_or_roughness * matGetSample3f(surface.uv, materialNode.u2, materialNode.tex, sc);
                uint32_t dst = o.fixQ4 ? o.paths[rayPathIndex].pixelIndex : rayPathIndex;  // :106, SURVEY Q4
                f4 &acc = o.traceAcc[dst];
                acc.x += add.x; acc.y += add.y; acc.z += add.z;
            }
            continue;
        }
        bool rejectSample = materialNode.type == BXDF_INVALID;
        if (bounce >= minBouncesForRR) {
            float rrProbability = cl_max(cl_min(0.5f, 0.2126f * curPathThroughput.x + 0.7152f * curPathThroughput.y + 0.0722f * curPathThroughput.z), 0.01f);
            if (rrProbability < sample2.x) {
                rejectSample = true;
            } else {
                curPathThroughput /= rrProbability;
            }
        }
        if (rejectSample) continue;

        bxdfSample = bxdfGetSample(&surface, &materialNode, sc, sample0, inRayDir, &bxdfOutRayDir, &bxdfPdf);
        float displaceDir = cl_sign(dot(surface.normal, bxdfOutRayDir));
        outBxdfRayOrigin = surface.point + surface.normal * displaceDir * INTERSECTION_EPSILON;  // DISPLACE_BY_EPSILON
        outEmissiveRayOrigin = surface.point + surface.normal * INTERSECTION_EPSILON;

        int emissiveIndex = numEmissives > 0 ? (int)emissiveSelect((int)numEmissives, sample1.x, &emissiveSelectionPdf) : -1;
        if (emissiveIndex > -1) {
            emissiveSample = emissiveGetSample(&surface, sc.emissives + emissiveIndex, sc, sample1, &emissiveOutRayDir, &emissivePdf, &distToEmissive);
            bxdfEmissivePdf = bxdfGetPdf(&surface, &materialNode, sc, inRayDir, emissiveOutRayDir);
            emissiveWeight = (emissivePdf * emissivePdf) / (emissivePdf * emissivePdf + bxdfEmissivePdf * bxdfEmissivePdf);  // POWER_HEURISTIC
            emissiveBxdfPdf = emissiveGetPdf(&surface, sc.emissives + emissiveIndex, sc, bxdfOutRayDir);
            bxdfWeight = (bxdfPdf * bxdfPdf) / (bxdfPdf * bxdfPdf + emissiveBxdfPdf * emissiveBxdfPdf);
        }
        float nDotEmissiveOutRay = cl_max(0.0f, dot(surface.normal, emissiveOutRayDir));
        if (cl_max(emissiveSample.x, cl_max(emissiveSample.y, emissiveSample.z)) > 0.0f && emissivePdf > 0.0f && nDotEmissiveOutRay > 0.0f) {
            bxdfEmissiveSample = bxdfEval(&surface, &materialNode, sc, inRayDir, emissiveOutRayDir);
            emissiveSample *= emissiveWeight * bxdfEmissiveSample * curPathThroughput * nDotEmissiveOutRay / (emissivePdf * emissiveSelectionPdf);
            if (cl_max(emissiveSample.x, cl_max(emissiveSample.y, emissiveSample.z)) > 0.0f) {
                o.wantOcc[globalId] = 1;
                o.occSample[globalId] = f4{emissiveSample.x, emissiveSample.y, emissiveSample.z, 0.0f};
                // rayNew(.., outEmissiveRayOrigin, emissiveOutRayDir, dist - eps, rayPathIndex) (:203)
                o.occRay[globalId].origin = f4{outEmissiveRayOrigin.x, outEmissiveRayOrigin.y, outEmissiveRayOrigin.z, distToEmissive - INTERSECTION_WITH_LIGHT_EPSILON};
                o.occRay[globalId].dir = f4{emissiveOutRayDir.x, emissiveOutRayDir.y, emissiveOutRayDir.z, (float)rayPathIndex};
            }
        }
        if (BXDF_IS_SINGULAR(materialNode.type)) bxdfWeight = 1.0f;
        f3 throughput = bxdfWeight * bxdfSample * bxdfTint * fabsf(dot(surface.normal, bxdfOutRayDir));
        if (cl_max(throughput.x, cl_max(throughput.y, throughput.z)) > 0.0f && bxdfPdf > 0.0f) {
            o.paths[rayPathIndex].throughput = curPathThroughput * throughput / bxdfPdf;  // pathSetThroughput
            o.wantInd[globalId] = 1;
            o.indRay[globalId].origin = f4{outBxdfRayOrigin.x, outBxdfRayOrigin.y, outBxdfRayOrigin.z, FLT_MAX};
            o.indRay[globalId].dir = f4{bxdfOutRayDir.x, bxdfOutRayDir.y, bxdfOutRayDir.z, (float)rayPathIndex};
        }
    }
    int nOcc = 0, nInd = 0;
    for (int i = 0; i < n; i++) {
        if (o.wantOcc[i]) {
            o.emissiveSamples[nOcc] = o.occSample[i];
            o.rays[2][nOcc] = o.occRay[i];
            nOcc++;
        }
        if (o.wantInd[i]) o.rays[1 - a][nInd++] = o.indRay[i];
    }
    o.numRays[2] = nOcc;
    o.numRays[1 - a] = nInd;
    o.stats.shaded_hits += shaded;
    o.stats.occlusion_emitted += (uint64_t)nOcc;
    o.stats.indirect_emitted += (uint64_t)nInd;
}

// kernels/pt_integrator.cl:278-296
void accumulateEmissiveSamples(Oracle &o, int buf) {
    const int n = o.numRays[buf];
    uint64_t un = 0;
    for (int globalId = 0; globalId < n; globalId++) {
        if (o.hitFlags[globalId]) continue;
        un++;
        uint32_t pathIndex = (uint32_t)o.rays[buf][globalId].dir.w;
        f4 &acc = o.traceAcc[o.paths[pathIndex].pixelIndex];
        const f4 &s = o.emissiveSamples[globalId];
        acc.x += s.x; acc.y += s.y; acc.z += s.z;
    }
    o.stats.unoccluded += un;
}

// kernels/hdr.cl:5-28
inline void tonemapOne(const f4 &accum, float sampleWeight, float exposure, uint8_t *out) {
    f3 hdrColor = F3(accum.x, accum.y, accum.z) * sampleWeight * exposure;
    f3 mapped = hdrColor / (hdrColor + 1.0f);
    const float e = 1.0f / 2.2f;
    f3 p = F3(powf(mapped.x, e), powf(mapped.y, e), powf(mapped.z, e));
    f3 normalizedOutput = F3(cl_clamp(p.x, 0.0f, 1.0f), cl_clamp(p.y, 0.0f, 1.0f), cl_clamp(p.z, 0.0f, 1.0f)) * 255.0f;
    out[0] = (uint8_t)normalizedOutput.x;
    out[1] = (uint8_t)normalizedOutput.y;
    out[2] = (uint8_t)normalizedOutput.z;
    out[3] = 255;
}

// splitmix64, used when the caller passes no seeds (the reference uses Go's global math/rand)
inline uint32_t nextSeed(uint64_t &s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)z;
}

// ------------------------------------------------------------------------------------------
// kernels/debug.cl:16-156 + their launches (resources.go:362-520)
// ------------------------------------------------------------------------------------------
inline f3 debugToneMapAndGammaCorrect(f3 sample) {  // debug.cl:8-13 (DEBUG_TONEMAP_EXPOSURE 1.0f)
    sample = sample * 1.0f;
    f3 mapped = sample / (sample + 1.0f);
    const float e = 1.0f / 2.2f;
    return F3(cl_clamp(powf(mapped.x, e), 0.0f, 1.0f), cl_clamp(powf(mapped.y, e), 0.0f, 1.0f), cl_clamp(powf(mapped.z, e), 0.0f, 1.0f)) * 255.0f;
}
inline void debugPut(Oracle &o, uint32_t pixel, uint8_t r, uint8_t g, uint8_t b) {
    uint8_t *p = &o.debugOutput[4 * (size_t)pixel];
    p[0] = r; p[1] = g; p[2] = b; p[3] = 255;
}
void debugClearBuffer(Oracle &o) {  // debug.cl:16-20 over FrameW*FrameH (resources.go:362-375)
    for (size_t i = 0; i < (size_t)o.frameW * o.frameH; i++) debugPut(o, (uint32_t)i, 0, 0, 0);
}
void debugRayIntersectionDepth(Oracle &o, const pc_block_request &req, int buf) {  // debug.cl:23-48, resources.go:378-421
    debugClearBuffer(o);
    float maxDepth = 1.0f;
    for (const Intersection &i : o.intersections)
        if (i.wuvt.w != FLT_MAX && i.wuvt.w > maxDepth) maxDepth = i.wuvt.w;
    const int n = (int)((size_t)req.frame_w * req.block_h);
    for (int globalId = 0; globalId < n && globalId < o.numRays[buf]; globalId++) {
        uint32_t pixelIndex = o.paths[globalId].pixelIndex;
        float hitDist = o.intersections[globalId].wuvt.w;
        if (!o.hitFlags[globalId] || hitDist == FLT_MAX) { debugPut(o, pixelIndex, 0, 0, 0); continue; }
        uint8_t sd = (uint8_t)(255.0f * (1.0f - hitDist / (maxDepth + 1.0f)));
        debugPut(o, pixelIndex, sd, sd, sd);
    }
}
void debugRayIntersectionNormals(Oracle &o, const pc_block_request &req, int buf) {  // debug.cl:51-97, resources.go:424-455
    debugClearBuffer(o);
    const int n = (int)((size_t)req.frame_w * req.block_h);
    for (int globalId = 0; globalId < n && globalId < o.numRays[buf]; globalId++) {
        uint32_t pixelIndex = o.paths[globalId].pixelIndex;
        float hitDist = o.intersections[globalId].wuvt.w;
        if (!o.hitFlags[globalId] || hitDist == FLT_MAX) { debugPut(o, pixelIndex, 0, 0, 0); continue; }
        Surface surface;
        surfaceInit(&surface, &o.intersections[globalId], o.sc);
        f3 inRayDir = -xyz(o.rays[buf][globalId].dir);
        MaterialNode materialNode;
        u2 rndState = u2{(uint32_t)globalId, (uint32_t)globalId};
        f3 bxdfTint;
        MatStats ms;
        matSelectNode(&o.paths[globalId], &surface, inRayDir, &materialNode, &bxdfTint, o.sc, &rndState, &ms);  // may set dispersion bits
        f3 val = (surface.normal + 1.0f) * 255.0f * 0.5f;
        debugPut(o, pixelIndex, (uint8_t)val.x, (uint8_t)val.y, (uint8_t)val.z);
    }
}
void debugEmissiveSamples(Oracle &o, const pc_block_request &req, uint32_t maskOccluded, uint32_t maskNotOccluded) {  // debug.cl:100-126
    debugClearBuffer(o);
    const int n = (int)((size_t)req.frame_w * req.block_h);
    for (int globalId = 0; globalId < n && globalId < o.numRays[2]; globalId++) {
        uint32_t pathIndex = (uint32_t)o.rays[2][globalId].dir.w;
        uint32_t pixelIndex = o.paths[pathIndex].pixelIndex;
        if ((maskOccluded && o.hitFlags[globalId]) || (maskNotOccluded && !o.hitFlags[globalId])) { debugPut(o, pixelIndex, 0, 0, 0); continue; }
        f3 val = debugToneMapAndGammaCorrect(xyz(o.emissiveSamples[globalId]));
        debugPut(o, pixelIndex, (uint8_t)val.x, (uint8_t)val.y, (uint8_t)val.z);
    }
}
void debugThroughput(Oracle &o, const pc_block_request &req) {  // debug.cl:129-140: every path of the block, no ray count
    debugClearBuffer(o);
    const size_t n = (size_t)req.frame_w * req.block_h;
    for (size_t globalId = 0; globalId < n; globalId++) {
        f3 val = debugToneMapAndGammaCorrect(o.paths[globalId].throughput);
        debugPut(o, o.paths[globalId].pixelIndex, (uint8_t)val.x, (uint8_t)val.y, (uint8_t)val.z);
    }
}
void debugAccumulator(Oracle &o, const pc_block_request &req) {  // debug.cl:143-154: indexed by work-item, not by pixelIndex
    debugClearBuffer(o);
    const size_t n = (size_t)req.frame_w * req.block_h;
    float sampleWeight = 1.0f / (float)(req.accumulated_samples + req.samples_per_pixel);  // resources.go:509
    for (size_t globalId = 0; globalId < n; globalId++) {
        f3 val = debugToneMapAndGammaCorrect(xyz(o.traceAcc[globalId]) * sampleWeight);
        debugPut(o, (uint32_t)globalId, (uint8_t)val.x, (uint8_t)val.y, (uint8_t)val.z);
    }
}
struct DebugSink {
    uint32_t flags = 0;
    uint8_t *frames = nullptr;
    uint64_t cap = 0;
    pc_debug_frame *infos = nullptr;
    uint32_t infosCap = 0, n = 0;
    bool capture = false;
    int overflow = 0;
};
void debugDump(Oracle &o, DebugSink &d, uint32_t flag, uint32_t bounce) {  // dumpDebugBuffer (pipeline.go:259-277)
    if (!d.capture) return;
    const uint64_t bytes = (uint64_t)o.frameW * o.frameH * 4;
    if ((uint64_t)(d.n + 1) * bytes > d.cap || d.n >= d.infosCap) { d.overflow = 1; return; }
    memcpy(d.frames + (uint64_t)d.n * bytes, o.debugOutput.data(), bytes);
    d.infos[d.n] = pc_debug_frame{flag, bounce};
    d.n++;
}

// Tracer.Trace (tracer.go:194-247) + MonteCarloIntegrator (pipeline.go:94-213); dbg adds the debug stages
int traceImpl(Oracle &o, pc_block_request *req, const uint32_t *seeds, size_t n_seeds, pc_stats *stats, DebugSink *dbg) {
    if (!o.sc.loaded) return PC_ERR_NO_SCENE_DATA;
    if (o.frameW != req->frame_w || o.frameH != req->frame_h) return PC_ERR_NO_FRAME;
    const size_t per_sample = 1 + (size_t)req->num_bounces;
    if (seeds && n_seeds < per_sample * req->samples_per_pixel) return PC_ERR_INVALID_ARGUMENT;
    auto t0 = std::chrono::steady_clock::now();
    memset(&o.stats, 0, sizeof(o.stats));
    const size_t px = (size_t)req->frame_w * req->frame_h;
    if (req->accumulated_samples == 0)  // pipeline.Reset -> ClearFrameAccumulator (tracer.go:208-213)
        std::fill(o.frameAcc.begin(), o.frameAcc.begin() + px, f4{0, 0, 0, 0});
    std::fill(o.traceAcc.begin(), o.traceAcc.begin() + px, f4{0, 0, 0, 0});  // tracer.go:215
    uint64_t own = 0x501A2150ull + req->seed;
    const uint32_t df = dbg ? dbg->flags : 0u;
    for (uint32_t sample = 0; sample < req->samples_per_pixel; sample++) {
        if (dbg) dbg->capture = sample + 1 == req->samples_per_pixel;  // the reference overwrites its PNGs every sample
        const uint32_t *ss = seeds ? seeds + per_sample * sample : nullptr;
        req->seed = ss ? ss[0] : nextSeed(own);  // tracer.go:222
        generatePrimaryRays(o, *req, req->seed);
        int activeRayBuf = 0;
        rayIntersectionQuery(o, activeRayBuf);  // CPU device: pipeline.go:110
        if (df & PC_DEBUG_PRIMARY_DEPTH) { debugRayIntersectionDepth(o, *req, activeRayBuf); debugDump(o, *dbg, PC_DEBUG_PRIMARY_DEPTH, 0); }        // :113-119
        if (df & PC_DEBUG_PRIMARY_NORMALS) { debugRayIntersectionNormals(o, *req, activeRayBuf); debugDump(o, *dbg, PC_DEBUG_PRIMARY_NORMALS, 0); }  // :120-126
        for (uint32_t bounce = 0; bounce < req->num_bounces; bounce++) {
            if (o.sc.sceneDiffuseMatIndex != -1) shadeMisses(o, activeRayBuf, bounce == 0);  // :134-143
            uint32_t shadeSeed = ss ? ss[1 + bounce] : nextSeed(own);                        // :146
            shadeHits(o, activeRayBuf, bounce, req->min_bounces_for_rr, shadeSeed);
            if (df & PC_DEBUG_THROUGHPUT) { debugThroughput(o, *req); debugDump(o, *dbg, PC_DEBUG_THROUGHPUT, bounce); }  // :151-157
            rayIntersectionTest(o, 2);       // :160
            accumulateEmissiveSamples(o, 2); // :165
            if (df & PC_DEBUG_ALL_EMISSIVE) { debugEmissiveSamples(o, *req, 0, 0); debugDump(o, *dbg, PC_DEBUG_ALL_EMISSIVE, bounce); }            // :170-176
            if (df & PC_DEBUG_VISIBLE_EMISSIVE) { debugEmissiveSamples(o, *req, 1, 0); debugDump(o, *dbg, PC_DEBUG_VISIBLE_EMISSIVE, bounce); }    // :178-184
            if (df & PC_DEBUG_OCCLUDED_EMISSIVE) { debugEmissiveSamples(o, *req, 0, 1); debugDump(o, *dbg, PC_DEBUG_OCCLUDED_EMISSIVE, bounce); }  // :186-192
            if (df & PC_DEBUG_ACCUMULATOR) { debugAccumulator(o, *req); debugDump(o, *dbg, PC_DEBUG_ACCUMULATOR, bounce); }                        // :194-200
            if (bounce + 1 < req->num_bounces) {  // :203-209
                activeRayBuf = 1 - activeRayBuf;
                rayIntersectionQuery(o, activeRayBuf);
            }
        }
        req->accumulated_samples++;  // tracer.go:240
    }
    auto t1 = std::chrono::steady_clock::now();
    o.stats.block_w = req->block_w;
    o.stats.block_h = req->block_h;
    o.stats.render_time_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
    if (stats) *stats = o.stats;
    return dbg && dbg->overflow ? PC_ERR_INVALID_ARGUMENT : 0;
}

}  // namespace

// ==========================================================================================
// C interface (ctypes): same shape as the pc_* ABI so tests drive both sides identically.
// ==========================================================================================
extern "C" {

void *po_create(void) { return new Oracle(); }
void po_destroy(void *h) { delete (Oracle *)h; }

int po_set_option(void *h, int option, int value) {
    Oracle &o = *(Oracle *)h;
    if (option == PC_OPT_FIX_Q4) { o.fixQ4 = value; return 0; }
    return 0;
}

// bufferSet.Resize (buffers.go:127-175): everything is frame sized.
int po_resize(void *h, uint32_t w, uint32_t hgt) {
    Oracle &o = *(Oracle *)h;
    o.frameW = w; o.frameH = hgt;
    size_t px = (size_t)w * hgt;
    for (auto &r : o.rays) r.assign(px, Ray{});
    o.paths.assign(px, Path{});
    o.hitFlags.assign(px, 0);
    o.intersections.assign(px, Intersection{});
    o.emissiveSamples.assign(px, f4{0, 0, 0, 0});
    o.traceAcc.assign(px, f4{0, 0, 0, 0});
    o.frameAcc.assign(px, f4{0, 0, 0, 0});
    o.frameBuffer.assign(px * 4, 0);
    o.debugOutput.assign(px * 4, 0);
    o.wantOcc.assign(px, 0); o.wantInd.assign(px, 0);
    o.occRay.assign(px, Ray{}); o.indRay.assign(px, Ray{});
    o.occSample.assign(px, f4{0, 0, 0, 0});
    return 0;
}

// The oracle borrows the caller's arrays (they must outlive the handle's use).
int po_upload_scene(void *h, const pc_scene_view *v) {
    Oracle &o = *(Oracle *)h;
    SceneRefs &s = o.sc;
    s.bvhNodes = (const BvhNode *)v->bvh_nodes;
    s.meshInstances = (const MeshInstance *)v->mesh_instances;
    s.materialNodes = (const MaterialNode *)v->material_nodes;
    s.texData = (const uint8_t *)v->texture_data;
    s.texMeta = (const TextureMetadata *)v->texture_metadata;
    s.vertices = (const f4 *)v->vertices;
    s.normals = (const f4 *)v->normals;
    s.uv = (const f2 *)v->uvs;
    s.materialIndices = (const uint32_t *)v->material_indices;
    s.emissives = (const Emissive *)v->emissives;
    s.numEmissives = (uint32_t)(v->emissives_bytes / sizeof(Emissive));
    s.sceneDiffuseMatIndex = v->scene_diffuse_mat_index;
    s.loaded = v->bvh_nodes_bytes >= sizeof(BvhNode);
    return 0;
}

int po_set_camera(void *h, const float eye[3], const float frustum[16]) {
    Oracle &o = *(Oracle *)h;
    o.eye = f3{eye[0], eye[1], eye[2]};
    memcpy(o.frustrum, frustum, 64);
    return 0;
}

int po_trace(void *h, pc_block_request *req, const uint32_t *seeds, size_t n_seeds, pc_stats *stats) {
    return traceImpl(*(Oracle *)h, req, seeds, n_seeds, stats, nullptr);
}

// MonteCarloIntegrator(debugFlags): same call shape as pc_trace_debug (include/polaris_cuda.h)
int po_trace_debug(void *h, pc_block_request *req, const uint32_t *seeds, size_t n_seeds, uint32_t debug_flags, uint8_t *frames_out,
                   uint64_t frames_cap_bytes, pc_debug_frame *infos, uint32_t infos_cap, uint32_t *n_frames, pc_stats *stats) {
    DebugSink d;
    d.flags = debug_flags; d.frames = frames_out; d.cap = frames_cap_bytes; d.infos = infos; d.infosCap = infos_cap;
    int rc = traceImpl(*(Oracle *)h, req, seeds, n_seeds, stats, &d);
    if (n_frames) *n_frames = d.n;
    return rc;
}

// Tracer.MergeOutput -> aggregateAccumulator (accumulator.cl:13-19, resources.go:108-124)
int po_merge_output(void *dst_h, void *src_h, const pc_block_request *req) {
    Oracle &dst = *(Oracle *)dst_h;
    Oracle &src = *(Oracle *)src_h;
    size_t off = (size_t)req->frame_w * req->block_y, n = (size_t)req->block_w * req->block_h;
    for (size_t g = off; g < off + n; g++) {
        dst.frameAcc[g].x += src.traceAcc[g].x;
        dst.frameAcc[g].y += src.traceAcc[g].y;
        dst.frameAcc[g].z += src.traceAcc[g].z;
    }
    return 0;
}

// Tracer.SyncFramebuffer -> tonemapSimpleReinhard (resources.go:344-360, hdr.cl:5-28)
int po_sync_framebuffer(void *h, const pc_block_request *req, uint8_t *rgba_out) {
    Oracle &o = *(Oracle *)h;
    if (!o.sc.loaded) return PC_ERR_NO_SCENE_DATA;
    size_t n = (size_t)req->frame_w * req->block_h;
    float sampleWeight = 1.0f / (float)(req->accumulated_samples + req->samples_per_pixel);  // resources.go:347, float32 division
    for (size_t g = 0; g < n; g++) tonemapOne(o.frameAcc[g], sampleWeight, req->exposure, &o.frameBuffer[4 * g]);
    if (rgba_out) memcpy(rgba_out, o.frameBuffer.data(), (size_t)o.frameW * o.frameH * 4);
    return 0;
}

int po_read_buffer(void *h, int which, void *dst, uint64_t bytes) {
    Oracle &o = *(Oracle *)h;
    const void *src = nullptr;
    size_t have = 0;
    switch (which) {
        case PC_BUF_RAYS0: case PC_BUF_RAYS1: case PC_BUF_RAYS2:
            src = o.rays[which].data(); have = o.rays[which].size() * sizeof(Ray); break;
        case PC_BUF_PATHS: src = o.paths.data(); have = o.paths.size() * sizeof(Path); break;
        case PC_BUF_HIT_FLAGS: src = o.hitFlags.data(); have = o.hitFlags.size() * 4; break;
        case PC_BUF_INTERSECTIONS: src = o.intersections.data(); have = o.intersections.size() * sizeof(Intersection); break;
        case PC_BUF_EMISSIVE_SAMPLES: src = o.emissiveSamples.data(); have = o.emissiveSamples.size() * 16; break;
        case PC_BUF_TRACE_ACCUMULATOR: src = o.traceAcc.data(); have = o.traceAcc.size() * 16; break;
        case PC_BUF_FRAME_ACCUMULATOR: src = o.frameAcc.data(); have = o.frameAcc.size() * 16; break;
        case PC_BUF_FRAME_BUFFER: src = o.frameBuffer.data(); have = o.frameBuffer.size(); break;
        case PC_BUF_RAY_COUNTERS: src = o.numRays; have = 12; break;
        default: return PC_ERR_INVALID_ARGUMENT;
    }
    if (bytes > have) return PC_ERR_INVALID_ARGUMENT;
    memcpy(dst, src, bytes);
    return 0;
}

int po_debug_intersect(void *h, const void *rays, uint32_t n, int mode, uint32_t *out_flags, void *out_hits,
                       uint64_t *counters /* nodes, tris, instances; may be NULL */) {
    Oracle &o = *(Oracle *)h;
    if (!o.sc.loaded) return PC_ERR_NO_SCENE_DATA;
    const Ray *r = (const Ray *)rays;
    Intersection *hits = (Intersection *)out_hits;
    uint64_t nodes = 0, tris = 0, inst = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : nodes, tris, inst)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        TraverseCounters c;
        if (mode == 1) {
            out_flags[i] = (uint32_t)traverse<true>(r[i], o.sc, nullptr, &c);
        } else {
            Intersection it;
            out_flags[i] = (uint32_t)traverse<false>(r[i], o.sc, &it, &c);
            if (hits) hits[i] = it;
        }
        nodes += c.nodes; tris += c.tris; inst += c.instances;
    }
    if (counters) { counters[0] = nodes; counters[1] = tris; counters[2] = inst; }
    return 0;
}

// BxDF table hook.  Input record (64 B): normal.xyz, matNodeIndex(u32 bits) | inRayDir.xyz, pad |
// outRayDir.xyz, pad | rand.xy, uv.xy.  Output record (48 B): sample.xyz, samplePdf |
// sampledDir.xyz, pdf(outRayDir) | eval(outRayDir).xyz, pad.  The node must be a leaf.
struct BxdfIn { float n[3]; uint32_t matNode; float in[3], p0; float out[3], p1; float rnd[2], uv[2]; };
struct BxdfOut { float sample[3], samplePdf; float dir[3], pdf; float eval[3], p; };
int po_debug_bxdf(void *h, const void *in_records, uint32_t n, void *out_records) {
    Oracle &o = *(Oracle *)h;
    if (!o.sc.loaded) return PC_ERR_NO_SCENE_DATA;
    const BxdfIn *in = (const BxdfIn *)in_records;
    BxdfOut *out = (BxdfOut *)out_records;
    for (uint32_t i = 0; i < n; i++) {
        Surface s;
        s.point = F3(0.0f);
        s.normal = F3(in[i].n[0], in[i].n[1], in[i].n[2]);
        s.uv = f2{in[i].uv[0], in[i].uv[1]};
        s.matNodeIndex = in[i].matNode;
        MaterialNode m = o.sc.materialNodes[in[i].matNode];
        f3 inDir = F3(in[i].in[0], in[i].in[1], in[i].in[2]);
        f3 outDir = F3(in[i].out[0], in[i].out[1], in[i].out[2]);
        f3 dir = F3(0.0f);
        float pdf = 1.0f;
        f3 smp = bxdfGetSample(&s, &m, o.sc, f2{in[i].rnd[0], in[i].rnd[1]}, inDir, &dir, &pdf);
        float p = bxdfGetPdf(&s, &m, o.sc, inDir, outDir);
        f3 e = bxdfEval(&s, &m, o.sc, inDir, outDir);
        out[i] = BxdfOut{{smp.x, smp.y, smp.z}, pdf, {dir.x, dir.y, dir.z}, p, {e.x, e.y, e.z}, 0.0f};
    }
    return 0;
}

int po_debug_rng(uint32_t *states_inout, uint32_t n, uint32_t draws, float *out) {
    for (uint32_t i = 0; i < n; i++) {
        u2 s = u2{states_inout[2 * i], states_inout[2 * i + 1]};
        for (uint32_t d = 0; d < draws; d++) {
            f2 v = randomGetSample2f(&s);
            out[2 * ((size_t)i * draws + d)] = v.x;
            out[2 * ((size_t)i * draws + d) + 1] = v.y;
        }
        states_inout[2 * i] = s.x;
        states_inout[2 * i + 1] = s.y;
    }
    return 0;
}

int po_debug_tonemap(const float *acc, uint32_t n, float sample_weight, float exposure, uint8_t *rgba_out) {
    for (uint32_t i = 0; i < n; i++) tonemapOne(f4{acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], 0.0f}, sample_weight, exposure, rgba_out + 4 * i);
    return 0;
}

// ------------------------------------------------------------------------------------------
// bvh.Build restated literally (asset/compiler/bvh/bvh_builder.go:124-308): every candidate
// plane is scored with a full pass over the node's items, O(planes x items).  Used by the
// tests to validate the binned builder in polaris_b200/csrc/scene_compiler.cpp.
// leaf encoding: ldata = -(first slot in out_order), rdata = count.
// ------------------------------------------------------------------------------------------
struct LitBuilder {
    const float *bmin, *bmax, *center;
    int minLeaf;
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> order;

    float scorePartition(const std::vector<uint32_t> &w) {
        if (w.empty()) return FLT_MAX;
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (uint32_t it : w)
            for (int k = 0; k < 3; k++) {
                if (bmin[3 * it + k] < mn[k]) mn[k] = bmin[3 * it + k];
                if (bmax[3 * it + k] > mx[k]) mx[k] = bmax[3 * it + k];
            }
        float s0 = mx[0] - mn[0], s1 = mx[1] - mn[1], s2 = mx[2] - mn[2];
        return (float)w.size() * (s0 * s1 + s1 * s2 + s0 * s2);
    }
    float scoreSplit(const std::vector<uint32_t> &w, int axis, float splitPoint, int *lc, int *rc) {
        float lmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, rmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        float lmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, rmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        int leftCount = 0, rightCount = 0;
        for (uint32_t it : w) {
            if (center[3 * it + axis] < splitPoint) {
                leftCount++;
                for (int k = 0; k < 3; k++) { if (bmin[3 * it + k] < lmin[k]) lmin[k] = bmin[3 * it + k]; if (bmax[3 * it + k] > lmax[k]) lmax[k] = bmax[3 * it + k]; }
            } else {
                rightCount++;
                for (int k = 0; k < 3; k++) { if (bmin[3 * it + k] < rmin[k]) rmin[k] = bmin[3 * it + k]; if (bmax[3 * it + k] > rmax[k]) rmax[k] = bmax[3 * it + k]; }
            }
        }
        *lc = leftCount; *rc = rightCount;
        if (leftCount == 0 || rightCount == 0) return FLT_MAX;
        float l0 = lmax[0] - lmin[0], l1 = lmax[1] - lmin[1], l2 = lmax[2] - lmin[2];
        float r0 = rmax[0] - rmin[0], r1 = rmax[1] - rmin[1], r2 = rmax[2] - rmin[2];
        return ((float)leftCount * (l0 * l1 + l1 * l2 + l0 * l2)) + ((float)rightCount * (r0 * r1 + r1 * r2 + r0 * r2));
    }
    uint32_t createLeaf(BvhNode node, const std::vector<uint32_t> &w) {
        node.left = -(int32_t)order.size();
        node.right = (int32_t)w.size();
        for (uint32_t it : w) order.push_back(it);
        nodes.push_back(node);
        return (uint32_t)nodes.size() - 1;
    }
    uint32_t partition(const std::vector<uint32_t> &w, int depth) {
        BvhNode node;
        node.minExtent = F3(FLT_MAX); node.maxExtent = F3(-FLT_MAX); node.left = node.right = 0;
        float *mn = &node.minExtent.x, *mx = &node.maxExtent.x;
        for (uint32_t it : w)
            for (int k = 0; k < 3; k++) {
                if (bmin[3 * it + k] < mn[k]) mn[k] = bmin[3 * it + k];
                if (bmax[3 * it + k] > mx[k]) mx[k] = bmax[3 * it + k];
            }
        if ((int)w.size() <= minLeaf) return createLeaf(node, w);
        float bestScore = scorePartition(w);
        bool haveBest = false;
        int bestAxis = 0, bestL = 0, bestR = 0;
        float bestSplit = 0;
        for (int axis = 0; axis < 3; axis++) {
            float side = mx[axis] - mn[axis];
            if (side < 1e-3f) continue;
            float splitStep = side / (1024.0f / (float)(depth + 1));
            if (splitStep < 1e-5f) continue;
            for (float splitPoint = mn[axis]; splitPoint < mx[axis]; splitPoint += splitStep) {
                int lc, rc;
                float score = scoreSplit(w, axis, splitPoint, &lc, &rc);
                if (score < bestScore) { bestScore = score; haveBest = true; bestAxis = axis; bestSplit = splitPoint; bestL = lc; bestR = rc; }
                if (!(splitPoint + splitStep > splitPoint)) { fprintf(stderr, "oracle bvh: stalled split loop\n"); abort(); }
            }
        }
        if (!haveBest) return createLeaf(node, w);
        std::vector<uint32_t> l, r;
        l.reserve(bestL); r.reserve(bestR);
        for (uint32_t it : w) (center[3 * it + bestAxis] < bestSplit ? l : r).push_back(it);
        uint32_t nodeIndex = (uint32_t)nodes.size();
        nodes.push_back(node);
        uint32_t li = partition(l, depth + 1);
        uint32_t ri = partition(r, depth + 1);
        nodes[nodeIndex].left = (int32_t)li;
        nodes[nodeIndex].right = (int32_t)ri;
        return nodeIndex;
    }
};

// returns node count; out_nodes must hold 2*n entries (32 B each), out_order n entries.
uint32_t po_build_bvh(const float *bmin, const float *bmax, const float *center, uint32_t n, int min_leaf_items,
                      void *out_nodes, uint32_t *out_order) {
    LitBuilder b{bmin, bmax, center, min_leaf_items, {}, {}};
    std::vector<uint32_t> w(n);
    for (uint32_t i = 0; i < n; i++) w[i] = i;
    b.partition(w, 0);
    memcpy(out_nodes, b.nodes.data(), b.nodes.size() * sizeof(BvhNode));
    memcpy(out_order, b.order.data(), b.order.size() * 4);
    return (uint32_t)b.nodes.size();
}

}  // extern "C"
