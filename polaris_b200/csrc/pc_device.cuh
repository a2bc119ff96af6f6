// pc_device.cuh -- per-ray device functions of the CUDA tracer (host+device portable).
//
// One function per reference device function (tracer/opencl/CL/**), written against CUDA vector
// types; each cites the .cl lines it implements.  The __global__ kernels that call these live in
// pc_kernels.cu; tests/emul compiles this header with g++ to compare it with the CPU oracle
// without a GPU (it is never part of the product path).
#pragma once
#include "pc_math.cuh"

namespace pc {

// ------------------------------------------------------------------------------------------------
// Device view of an uploaded scene (see pc_layout.hpp for the derived records)
// ------------------------------------------------------------------------------------------------
struct DScene {
    // derived traversal layout
    const float4 *node64;   // 4 x float4 per inner node
    const float4 *tri48;    // 3 x float4 per triangle
    const float4 *inst80;   // 5 x float4 per instance
#ifdef PC_WIDE_BVH
    const float4 *node128;  // 8 x float4 per inner node: 4-ary collapse (experiment)
#endif
    uint32_t rootRef;
    // reference buffers (shading + reference-order traversal)
    const float4 *bvhNodes;       // 2 x float4 per node
    const float4 *meshInstances;  // 5 x float4 per instance
    const float4 *vertices;
    const float4 *normals;
    const float2 *uvs;
    const uint32_t *matIndex;
    const float4 *matNodes;   // 4 x float4 per node
    const float4 *emissives;  // 5 x float4 per emissive
    const uint4 *texMeta;
    const uint8_t *texData;
    uint32_t numEmissives;
    int32_t sceneDiffuseMat;
    // world bounds (top-level BVH root box) as origin-cell transform of the traversal-order sort key: cell = (o - min) * scale
    float3 worldMin, worldCellScale;
};

constexpr uint32_t REF_LEAF = 0x80000000u;
constexpr uint32_t REF_TOP = 0x40000000u;
constexpr uint32_t REF_POP_INSTANCE = 0xFFFFFFFFu;  // stack marker: restore the world-space ray
constexpr uint32_t REF_DONE = 0xFFFFFFFEu;          // traversal finished (never stored)
constexpr uint32_t REF_POP_TRANSLATED = 0xFFFFFFFDu;  // stack marker: restore the origin only (translation-only instance)
constexpr uint32_t INST_FLAG_IDENTITY = 1u;
constexpr uint32_t INST_FLAG_TRANSLATION = 2u;

// bxdf.cl:10-21, material_sampler.cl:4-9, path.cl:4-6, emissive_sampler.cl:4-5, texture_sampler.cl:4-7
constexpr uint32_t BXDF_INVALID = 0, BXDF_EMISSIVE = 2, BXDF_DIFFUSE = 4, BXDF_CONDUCTOR = 8,
                   BXDF_ROUGH_CONDUCTOR = 16, BXDF_DIELECTRIC = 32, BXDF_ROUGH_DIELECTRIC = 64;
constexpr uint32_t OP_MIX = 10001, OP_MIX_MAP = 10002, OP_BUMP_MAP = 10003, OP_NORMAL_MAP = 10004,
                   OP_DISPERSE = 10005;
constexpr uint32_t PATH_DISPERSE_R = 1, PATH_DISPERSE_G = 2, PATH_DISPERSE_B = 4;
constexpr uint32_t TEX_L8 = 0, TEX_L32F = 1, TEX_RGBA8 = 2, TEX_RGBA32F = 3;

struct MatNode {  // types.cl:110-165, loaded as 4 x float4
    uint32_t type, left;
    int32_t rightOrTransTex, tex;
    float3 u2;  // reflectance | specularity | radiance | intDispersionIORs | mixWeight(.x)
    float3 u3;  // transmittance | extDispersionIORs
    float intIOR, extIOR, scaleOrRoughness;
    int32_t roughnessTex;
};

PC_HD MatNode loadMatNode(const DScene &sc, uint32_t i) {
    float4 a = PC_LDG(sc.matNodes + 4 * (size_t)i), b = PC_LDG(sc.matNodes + 4 * (size_t)i + 1);
    float4 c = PC_LDG(sc.matNodes + 4 * (size_t)i + 2), d = PC_LDG(sc.matNodes + 4 * (size_t)i + 3);
    MatNode m;
    m.type = f2u(a.x); m.left = f2u(a.y); m.rightOrTransTex = (int32_t)f2u(a.z); m.tex = (int32_t)f2u(a.w);
    m.u2 = xyz(b); m.u3 = xyz(c);
    m.intIOR = d.x; m.extIOR = d.y; m.scaleOrRoughness = d.z; m.roughnessTex = (int32_t)f2u(d.w);
    return m;
}

struct Surface {  // types.cl:85-97
    float3 point, normal;
    float2 uv;
    uint32_t matNodeIndex;
};

// ------------------------------------------------------------------------------------------------
// samplers/random_sampler.cl:7-16
// ------------------------------------------------------------------------------------------------
PC_HD float2 randomGetSample2f(uint2 &state) {
    const float invMaxInt = 1.0f / 4294967296.0f;
    uint32_t x = state.x * 17u + state.y * 13123u;
    state.x = (x << 13) ^ x;
    state.y ^= (x << 7);
    uint32_t t0 = x * (x * x * 15731u + 74323u) + 871483u;
    uint32_t t1 = x * (x * x * 13734u + 37828u) + 234234u;
#if defined(__CUDA_ARCH__)
    return make_float2(__uint2float_rn(t0) * invMaxInt, __uint2float_rn(t1) * invMaxInt);
#else
    return make_float2((float)t0 * invMaxInt, (float)t1 * invMaxInt);
#endif
}

// ------------------------------------------------------------------------------------------------
// util/transform.cl:9-38, util/fresnel.cl:8-16, util/surface.cl:4-33
// ------------------------------------------------------------------------------------------------
PC_HD float3 mul4x1(float3 v, float4 m0, float4 m1, float4 m2, float4 m3) {
    float3 o;
    o.x = m0.x * v.x + m1.x * v.y + m2.x * v.z + m3.x;
    o.y = m0.y * v.x + m1.y * v.y + m2.y * v.z + m3.y;
    o.z = m0.z * v.x + m1.z * v.y + m2.z * v.z + m3.z;
    return o;
}
PC_HD float3 mul3x1(float3 v, float4 m0, float4 m1, float4 m2) {
    float3 o;
    o.x = m0.x * v.x + m1.x * v.y + m2.x * v.z;
    o.y = m0.y * v.x + m1.y * v.y + m2.y * v.z;
    o.z = m0.z * v.x + m1.z * v.y + m2.z * v.z;
    return o;
}
PC_HD_NOINLINE float2 rayToLatLongUV(float3 v) {
    float at2 = atan2f(v.x, v.z);
    float r = length(v);
    return make_float2((at2 >= 0.0f ? at2 : (at2 + PC_TWO_PI)) / PC_TWO_PI, acosf(v.y / r) / PC_PI);
}
PC_HD float fresnelForDielectric(float etaI, float etaT, float iDotN) {
    float eta = etaI / etaT;
    float r0 = ((1.0f - eta) * (1.0f - eta)) / ((1.0f + eta) * (1.0f + eta));
    float c = 1.0f - fabsf(iDotN);
    float c1 = c * c;
    return r0 + (1.0f - r0) * c1 * c1 * c;
}
PC_HD void tangentVectors(float3 n, float3 &u, float3 &v) {
    u = normalize(cross((fabsf(n.z) < .999f ? f3(0.0f, 0.0f, 1.0f) : f3(1.0f, 0.0f, 0.0f)), n));
    v = cross(n, u);
}
PC_HD void surfaceInit(Surface &s, float4 wuvt, uint32_t triIndex, const DScene &sc) {
    float3 wuv = xyz(wuvt);
    size_t off = (size_t)triIndex * 3;
    float4 v0 = PC_LDG(sc.vertices + off), v1 = PC_LDG(sc.vertices + off + 1), v2 = PC_LDG(sc.vertices + off + 2);
    float4 n0 = PC_LDG(sc.normals + off), n1 = PC_LDG(sc.normals + off + 1), n2 = PC_LDG(sc.normals + off + 2);
    float2 t0 = PC_LDG(sc.uvs + off), t1 = PC_LDG(sc.uvs + off + 1), t2 = PC_LDG(sc.uvs + off + 2);
    s.point = xyz(wuv.x * v0 + wuv.y * v1 + wuv.z * v2);
    s.normal = normalize(xyz(wuv.x * n0 + wuv.y * n1 + wuv.z * n2));
    s.uv = wuv.x * t0 + wuv.y * t1 + wuv.z * t2;
    s.matNodeIndex = PC_LDG(sc.matIndex + triIndex);
}

// ------------------------------------------------------------------------------------------------
// samplers/texture_sampler.cl -- bilinear fetch, clamp-to-edge on the +1 tap, 4 formats
// ------------------------------------------------------------------------------------------------
struct TexTaps {
    uint32_t tx, ty, bx, by, w, format;
    float cx, cy;
    const uint8_t *base;
};
PC_HD TexTaps texTaps(float2 uv, int texIndex, const uint4 *texMeta, const uint8_t *texData) {  // :15-36
    uint4 m = PC_LDG(texMeta + texIndex);  // format, width, height, dataOffset
    TexTaps t;
    float sx = uv.x - floorf(uv.x), sy = uv.y - floorf(uv.y);
    sx *= (float)m.y;
    sy *= (float)m.z;
    t.tx = cl_clampu((uint32_t)sx, 0u, m.y - 1);
    t.ty = cl_clampu((uint32_t)sy, 0u, m.z - 1);
    t.bx = cl_clampu(t.tx + 1, 0u, m.y - 1);
    t.by = cl_clampu(t.ty + 1, 0u, m.z - 1);
    t.cx = sx - (float)t.tx;
    t.cy = sy - (float)t.ty;
    t.w = m.y;
    t.format = m.x;
    t.base = texData + m.w;
    return t;
}
PC_HD float ldF(const uint8_t *p) { return PC_LDG((const float *)p); }
PC_HD float ldB(const uint8_t *p) { return (float)PC_LDG(p); }
PC_HD float4 ldRGBA32F(const uint8_t *base, uint32_t i) {
    const uint8_t *p = base + 16 * (size_t)i;
    if ((((uintptr_t)p) & 15) == 0) return PC_LDG((const float4 *)p);  // texture blobs are only 4 B aligned
    return make_float4(ldF(p), ldF(p + 4), ldF(p + 8), ldF(p + 12));
}
PC_HD float4 ldRGBA8(const uint8_t *base, uint32_t i) {
    uint32_t v = PC_LDG((const uint32_t *)(base + 4 * (size_t)i));  // little endian r,g,b,a
    return make_float4((float)(v & 255u), (float)((v >> 8) & 255u), (float)((v >> 16) & 255u), (float)(v >> 24));
}

PC_HD_NOINLINE float3 texSample3(float2 uv, int texIndex, const uint4 *texMeta, const uint8_t *texData) {  // :14-101
    TexTaps t = texTaps(uv, texIndex, texMeta, texData);
    uint32_t iTL = t.ty * t.w + t.tx, iTR = t.ty * t.w + t.bx, iBL = t.by * t.w + t.tx, iBR = t.by * t.w + t.bx;
    switch (t.format) {
        case TEX_RGBA8: {
            float4 a = ldRGBA8(t.base, iTL), b = ldRGBA8(t.base, iTR), c = ldRGBA8(t.base, iBL), d = ldRGBA8(t.base, iBR);
            return xyz(mix(mix(a, c, t.cy), mix(b, d, t.cy), t.cx)) / 255.0f;
        }
        case TEX_RGBA32F: {
            float4 a = ldRGBA32F(t.base, iTL), b = ldRGBA32F(t.base, iTR), c = ldRGBA32F(t.base, iBL), d = ldRGBA32F(t.base, iBR);
            return xyz(mix(mix(a, c, t.cy), mix(b, d, t.cy), t.cx));
        }
        case TEX_L8: {
            float a = ldB(t.base + iTL), b = ldB(t.base + iTR), c = ldB(t.base + iBL), d = ldB(t.base + iBR);
            float r = mix(mix(a, c, t.cy), mix(b, d, t.cy), t.cx) / 255.0f;
            return f3(r, r, r);
        }
        case TEX_L32F: {
            float a = ldF(t.base + 4 * (size_t)iTL), b = ldF(t.base + 4 * (size_t)iTR);
            float c = ldF(t.base + 4 * (size_t)iBL), d = ldF(t.base + 4 * (size_t)iBR);
            float r = mix(mix(a, c, t.cy), mix(b, d, t.cy), t.cx);
            return f3(r, r, r);
        }
    }
    return f3(0.0f, 0.0f, 0.0f);
}
// red channel of the four taps, by format (:105-184 and the three taps of :187-251)
PC_HD float texRed(const TexTaps &t, uint32_t i) {
    switch (t.format) {
        case TEX_RGBA8: return ldB(t.base + 4 * (size_t)i);
        case TEX_RGBA32F: return ldF(t.base + 16 * (size_t)i);
        case TEX_L8: return ldB(t.base + i);
        case TEX_L32F: return ldF(t.base + 4 * (size_t)i);
    }
    return 0.0f;
}
PC_HD_NOINLINE float texSample1(float2 uv, int texIndex, const uint4 *texMeta, const uint8_t *texData) {  // :105-184
    TexTaps t = texTaps(uv, texIndex, texMeta, texData);
    if (t.format > TEX_RGBA32F) return 0.0f;
    float a = texRed(t, t.ty * t.w + t.tx), b = texRed(t, t.ty * t.w + t.bx);
    float c = texRed(t, t.by * t.w + t.tx), d = texRed(t, t.by * t.w + t.bx);
    float r = mix(mix(a, c, t.cy), mix(b, d, t.cy), t.cx);
    return (t.format == TEX_RGBA8 || t.format == TEX_L8) ? r / 255.0f : r;
}
PC_HD_NOINLINE float3 texBump3(float2 uv, int texIndex, const uint4 *texMeta, const uint8_t *texData) {  // :187-251
    TexTaps t = texTaps(uv, texIndex, texMeta, texData);
    if (t.format > TEX_RGBA32F) return f3(0.0f, 0.0f, 0.0f);
    float s0 = texRed(t, t.ty * t.w + t.tx), s1 = texRed(t, t.ty * t.w + t.bx), s2 = texRed(t, t.by * t.w + t.tx);
    if (t.format == TEX_RGBA8 || t.format == TEX_L8) {
        s0 = s0 / 255.0f; s1 = s1 / 255.0f; s2 = s2 / 255.0f;
    }
    return f3(0.5f, 0.5f, 0.5f) + 0.5f * normalize(f3(s1 - s0, s2 - s0, 1.0f));
}

PC_HD float3 texGetSample3f(float2 uv, int texIndex, const DScene &sc) { return texSample3(uv, texIndex, sc.texMeta, sc.texData); }
PC_HD float texGetSample1f(float2 uv, int texIndex, const DScene &sc) { return texSample1(uv, texIndex, sc.texMeta, sc.texData); }
PC_HD float3 texGetBumpSample3f(float2 uv, int texIndex, const DScene &sc) { return texBump3(uv, texIndex, sc.texMeta, sc.texData); }

// ------------------------------------------------------------------------------------------------
// samplers/material_sampler.cl
// ------------------------------------------------------------------------------------------------
PC_HD float3 matGetSample3f(float2 uv, float3 def, int tex, const DScene &sc) {  // :92-98
    if (tex == -1) return def;
    return texGetSample3f(uv, tex, sc);
}
PC_HD float matGetSample1f(float2 uv, float def, int tex, const DScene &sc) {  // :102-108
    if (tex == -1) return def;
    return texGetSample1f(uv, tex, sc);
}
PC_HD float3 matGetNormalSample3f(float3 normal, float2 uv, int tex, const DScene &sc) {  // :111-121
    float3 u, v;
    tangentVectors(normal, u, v);
    float3 s = (texGetSample3f(uv, tex, sc) * 2.0f) - 1.0f;
    return normalize(u * s.x + v * s.y + 0.5f * normal * s.z);
}
PC_HD float3 matGetBumpSample3f(float3 normal, float2 uv, int tex, const DScene &sc) {  // :124-131
    float3 u, v;
    tangentVectors(normal, u, v);
    float3 s = (texGetBumpSample3f(uv, tex, sc) * 2.0f) - 1.0f;
    return normalize(u * s.x + v * s.y + normal * s.z);
}

// matSelectNode (:21-88). pathFlags is read-modify-written by the caller's copy; the walk is
// bounded (a malformed tree cannot hang the GPU).
PC_HD MatNode matSelectNode(uint32_t &pathFlags, Surface &surface, float3 &tint, const DScene &sc, uint2 &rnd) {
    MatNode node = loadMatNode(sc, surface.matNodeIndex);
    float2 forceIOR = make_float2(0.0f, 0.0f);
    float2 smp;
    for (int guard = 0; node.type >= OP_MIX; guard++) {
        if (guard >= 64) {
            node.type = BXDF_INVALID;
            return node;
        }
        uint32_t next = node.left;
        switch (node.type) {
            case OP_MIX:
                smp = randomGetSample2f(rnd);
                next = smp.x < node.u2.x ? node.left : (uint32_t)node.rightOrTransTex;
                break;
            case OP_MIX_MAP:
                smp = randomGetSample2f(rnd);
                smp.y = texGetSample1f(surface.uv, node.tex, sc);
                next = smp.x < smp.y ? node.left : (uint32_t)node.rightOrTransTex;
                break;
            case OP_BUMP_MAP:
                surface.normal = matGetBumpSample3f(surface.normal, surface.uv, node.tex, sc);
                break;
            case OP_NORMAL_MAP:
                surface.normal = matGetNormalSample3f(surface.normal, surface.uv, node.tex, sc);
                break;
            case OP_DISPERSE: {
                uint32_t ch;
                if ((pathFlags & PATH_DISPERSE_R) != 0) ch = 0;
                else if ((pathFlags & PATH_DISPERSE_G) != 0) ch = 1;
                else if ((pathFlags & PATH_DISPERSE_B) != 0) ch = 2;
                else {
                    smp = randomGetSample2f(rnd);
                    ch = smp.x < 0.333f ? 0 : (smp.x < 0.666f ? 1 : 2);
                    pathFlags |= (1u << ch);
                }
                tint = f3(ch == 0 ? 1.0f : 0.0f, ch == 1 ? 1.0f : 0.0f, ch == 2 ? 1.0f : 0.0f);
                forceIOR = ch == 0 ? make_float2(node.u2.x, node.u3.x)
                                   : (ch == 1 ? make_float2(node.u2.y, node.u3.y) : make_float2(node.u2.z, node.u3.z));
                break;
            }
            default:
                node.type = BXDF_INVALID;
                return node;
        }
        node = loadMatNode(sc, next);
    }
    node.intIOR = cl_max(node.intIOR, forceIOR.x);
    node.extIOR = cl_max(node.extIOR, forceIOR.y);
    return node;
}

// ------------------------------------------------------------------------------------------------
// samplers/distribution_sampler.cl
// ------------------------------------------------------------------------------------------------
PC_HD float ggxG1(float roughness, float3 v, float3 n, float3 m) {  // :16-29
    float nDotV = dot(n, v);
    float mDotV = dot(m, v);
    if (nDotV * mDotV <= 0.0f) return 0.0f;
    float nDotVSq = nDotV * nDotV;
    float tanSq = nDotVSq > 0.0f ? (1.0f - nDotVSq) / nDotVSq : 0.0f;
    float aSq = roughness * roughness;
    return 2.0f / (1.0f + sqrtf(1.0f + aSq * tanSq));
}
PC_HD float ggxGetG(float roughness, float3 i, float3 o, float3 n, float3 m) {  // :33-35
    return ggxG1(roughness, i, n, m) * ggxG1(roughness, o, n, m);
}
PC_HD float ggxGetD(float roughness, float3 n, float3 m) {  // :38-52
    float nDotM = dot(n, m);
    if (nDotM <= 0.0f) return 0.0f;
    float nDotMSq = nDotM * nDotM;
    float tanSq = nDotM != 0.0f ? ((1.0f - nDotMSq) / nDotMSq) : 0.0f;
    float aSq = roughness * roughness;
    float denom = PC_PI * nDotMSq * nDotMSq * (aSq + tanSq) * (aSq + tanSq);
    return denom > 0.0f ? (aSq / denom) : 0.0f;
}
PC_HD_NOINLINE float3 ggxGetSample(float roughness, float3 n, float2 r) {  // :55-74 (sinPhi >= 0: SURVEY Q8)
    float3 u, v;
    tangentVectors(n, u, v);
    float theta = atanf(roughness * sqrtf(r.x / (1.0f - r.x)));
    theta = theta >= 0.0f ? theta : (theta + PC_TWO_PI);
    float cosTheta = cosf(theta);
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    float cosPhi = cosf(PC_TWO_PI * r.y);
    float sinPhi = sqrtf(1.0f - cosPhi * cosPhi);
    return normalize(u * sinTheta * cosPhi + v * sinTheta * sinPhi + n * cosTheta);
}
PC_HD_NOINLINE float ggxGetReflectionPdf(float roughness, float3 o, float3 n, float3 h) {  // :76-85
    float nDotH = fabsf(dot(n, h));
    float oDotH = fabsf(dot(o, h));
    float denom = 4.0f * oDotH;
    return denom == 0.0f ? 0.0f : ggxGetD(roughness, n, h) * nDotH / denom;
}
PC_HD_NOINLINE float ggxGetRefractionPdf(float roughness, float etaI, float etaT, float3 i, float3 o, float3 n, float3 h) {  // :87-96
    float iDotH = fabsf(dot(i, h));
    float oDotH = fabsf(dot(o, h));
    float hDotN = fabsf(dot(h, n));
    float denom = (etaI * iDotH + etaT * oDotH) * (etaI * iDotH + etaT * oDotH);
    return denom > 0.0f ? ggxGetD(roughness, n, h) * hDotN * oDotH * etaT * etaT / denom : 0.0f;
}
PC_HD_NOINLINE float3 cosWeightedHemisphereGetSample(float3 normal, float2 r) {  // :101-112
    float rd = sqrtf(r.x);
    float phi = PC_TWO_PI * r.y;
    float3 u, v;
    tangentVectors(normal, u, v);
    return normalize(u * rd * cosf(phi) + v * rd * sinf(phi) + normal * sqrtf(1 - r.x));
}

// ------------------------------------------------------------------------------------------------
// bxdf/*.cl
// ------------------------------------------------------------------------------------------------
PC_HD float roughnessOf(const Surface &s, const MatNode &m, const DScene &sc) {
    float r = cl_clamp(matGetSample1f(s.uv, m.scaleOrRoughness, m.roughnessTex, sc), PC_MIN_ROUGHNESS, 1.0f);
    r *= r;  // Disney remapping a = roughness^2
    return r;
}
PC_HD float conductorFresnel(const MatNode &m, float iDotN) {
    return m.intIOR != 0.0f ? fresnelForDielectric(m.extIOR, m.intIOR, iDotN) : 1.0f;
}
// GGX reflection lobe shared by roughConductor and the reflected branch of roughDielectric
// (rough_conductor.cl:36-39,61-77; rough_dielectric.cl:41-55,135-146)
PC_HD_NOINLINE float3 ggxReflectionEval(float roughness, float3 ks, float f, float3 i, float3 o, float3 n) {
    float iDotN = dot(i, n);
    float oDotN = dot(o, n);
    float3 h = normalize(i + o);
    float d = ggxGetD(roughness, n, h);
    float g = ggxGetG(roughness, i, o, n, h);
    float denom = 4.0f * iDotN * oDotN;
    return denom > 0.0f ? ks * f * d * g / denom : f3s(0.0f);
}
// equation 21 of Walter et al. as written in rough_dielectric.cl:59-82,148-165 (tf is the
// transmittance sample; the reference fetches it after the zero-denominator test, it is pure)
PC_HD_NOINLINE float3 ggxRefractionEval(float roughness, float3 tf, float f, float etaI, float etaT, float3 i, float3 o,
                                        float3 h, float3 n) {
    float iDotN = dot(i, n);
    float oDotN = dot(o, n);
    float iDotH = fabsf(dot(i, h));
    float oDotH = fabsf(dot(o, h));
    float focusTermDenom = iDotN * oDotN * (etaI * iDotH + etaT * oDotH) * (etaI * iDotH + etaT * oDotH);
    if (focusTermDenom == 0.0f) return f3(0.0f, 0.0f, 0.0f);
    float focusTerm = fabsf(etaT * etaT * iDotH * oDotH / focusTermDenom);
    float d = ggxGetD(roughness, n, h);
    float g = ggxGetG(roughness, i, o, n, h);
    return tf * (1.0f - f) * d * g * focusTerm;
}

// bxdfGetSample (bxdf.cl:29-54) and the five *Sample functions
PC_HD float3 bxdfGetSample(const Surface &s, const MatNode &m, const DScene &sc, float2 r, float3 in, float3 &out, float &pdf) {
    const float3 n = s.normal;
    switch (m.type) {
        case BXDF_DIFFUSE: {  // diffuse.cl:12-20
            out = cosWeightedHemisphereGetSample(n, r);
            pdf = dot(n, out) * PC_1_PI;
            return matGetSample3f(s.uv, m.u2, m.tex, sc) * PC_1_PI;
        }
        case BXDF_CONDUCTOR: {  // conductor.cl:12-29
            float iDotN = dot(in, n);
            out = 2.0f * iDotN * n - in;
            pdf = 1.0f;
            float f = conductorFresnel(m, iDotN);
            float3 ks = matGetSample3f(s.uv, m.u2, m.tex, sc);
            return iDotN != 0.0f ? f * ks / iDotN : f3s(0.0f);
        }
        case BXDF_DIELECTRIC: {  // dielectric.cl:12-45
            float iDotN = dot(in, n);
            float etaI = m.extIOR, etaT = m.intIOR;
            if (iDotN < 0.0f) { float t = etaI; etaI = etaT; etaT = t; }
            float eta = etaI / etaT;
            float f = fresnelForDielectric(etaI, etaT, iDotN);
            float3 kVal;
            float cosTSq = 1.0f + eta * (iDotN * iDotN - 1.0f);  // eta, not eta^2
            if (cosTSq <= 0.0f || r.x <= f) {
                out = -cl_sign(iDotN) * 2.0f * iDotN * n - in;
                kVal = matGetSample3f(s.uv, m.u2, m.tex, sc);
                pdf = cosTSq <= 0.0f ? 1.0f : f;
            } else {
                out = (eta * iDotN - cl_sign(iDotN) * sqrtf(cosTSq)) * n - eta * in;
                kVal = eta * eta * matGetSample3f(s.uv, m.u3, m.rightOrTransTex, sc);
                pdf = 1.0f - f;
            }
            return iDotN != 0.0f ? pdf * kVal / fabsf(iDotN) : f3s(0.0f);
        }
        case BXDF_ROUGH_CONDUCTOR: {  // rough_conductor.cl:9-40
            float roughness = roughnessOf(s, m, sc);
            float3 ks = matGetSample3f(s.uv, m.u2, m.tex, sc);
            float3 h = ggxGetSample(roughness, n, r);
            out = 2.0f * dot(in, h) * h - in;
            pdf = ggxGetReflectionPdf(roughness, out, n, h);
            return ggxReflectionEval(roughness, ks, conductorFresnel(m, dot(in, n)), in, out, n);
        }
        case BXDF_ROUGH_DIELECTRIC: {  // rough_dielectric.cl:9-83
            float iDotN = dot(in, n);
            float roughness = roughnessOf(s, m, sc);
            float etaI = m.extIOR, etaT = m.intIOR;
            if (iDotN < 0.0f) { float t = etaI; etaI = etaT; etaT = t; }
            float eta = etaI / etaT;
            float3 h = ggxGetSample(roughness, n, r);
            float f = fresnelForDielectric(etaI, etaT, iDotN);
            float cosTSq = 1.0f + eta * (iDotN * iDotN - 1.0f);
            if (cosTSq <= 0.0f || r.x <= f) {
                out = 2.0f * dot(in, h) * h - in;
                float3 ks = matGetSample3f(s.uv, m.u2, m.tex, sc);
                h = normalize(in + out);
                pdf = cosTSq <= 0.0f ? 1.0f : ggxGetReflectionPdf(roughness, out, n, h);
                return ggxReflectionEval(roughness, ks, f, in, out, n);
            }
            out = (eta * iDotN - cl_sign(iDotN) * sqrtf(cosTSq)) * h - eta * in;
            h = normalize(-(etaI * in + etaT * out));
            pdf = ggxGetRefractionPdf(roughness, etaI, etaT, in, out, n, h);
            return ggxRefractionEval(roughness, matGetSample3f(s.uv, m.u3, m.rightOrTransTex, sc), f, etaI, etaT, in, out, h, n);
        }
    }
    return f3(0.0f, 0.0f, 0.0f);
}

// bxdfGetPdf (bxdf.cl:57-79)
PC_HD float bxdfGetPdf(const Surface &s, const MatNode &m, const DScene &sc, float3 in, float3 out) {
    const float3 n = s.normal;
    switch (m.type) {
        case BXDF_DIFFUSE: return dot(n, out) * PC_1_PI;  // diffuse.cl:24-26
        case BXDF_CONDUCTOR: {                            // conductor.cl:33-40 (as written)
            float iDotN = dot(in, n);
            float3 expOut = 2.0f * iDotN * n - in;
            float expDot = dot(expOut, out);
            return expDot >= 0.0f && expDot <= 0.001f ? 1.0f : 0.0f;
        }
        case BXDF_DIELECTRIC: return 0.0f;  // dielectric.cl:49-52
        case BXDF_ROUGH_CONDUCTOR: {        // rough_conductor.cl:43-51
            float roughness = roughnessOf(s, m, sc);
            float3 h = normalize(in + out);
            return ggxGetReflectionPdf(roughness, out, n, h);
        }
        case BXDF_ROUGH_DIELECTRIC: {  // rough_dielectric.cl:86-111
            float iDotN = dot(in, n);
            float roughness = roughnessOf(s, m, sc);
            if (iDotN > 0.0f) {
                float3 h = normalize(in + out);
                return ggxGetReflectionPdf(roughness, out, n, h);
            }
            float etaI = m.extIOR, etaT = m.intIOR;
            if (iDotN < 0.0f) { float t = etaI; etaI = etaT; etaT = t; }
            float3 h = normalize(-(etaI * in + etaT * out));
            return ggxGetRefractionPdf(roughness, etaI, etaT, in, out, n, h);
        }
    }
    return 0.0f;
}

// bxdfEval (bxdf.cl:83-105)
PC_HD float3 bxdfEval(const Surface &s, const MatNode &m, const DScene &sc, float3 in, float3 out) {
    const float3 n = s.normal;
    switch (m.type) {
        case BXDF_DIFFUSE: return matGetSample3f(s.uv, m.u2, m.tex, sc) * PC_1_PI;  // diffuse.cl:29-32
        case BXDF_CONDUCTOR: {                                                      // conductor.cl:45-62
            float iDotN = dot(in, n);
            float3 expOut = 2.0f * iDotN * n - in;
            float expDot = dot(expOut, out);
            if (expDot < 0.0f || expDot > 0.001f) return f3(0.0f, 0.0f, 0.0f);
            float f = conductorFresnel(m, iDotN);
            float3 ks = matGetSample3f(s.uv, m.u2, m.tex, sc);
            return iDotN != 0.0f ? f * ks / iDotN : f3s(0.0f);
        }
        case BXDF_DIELECTRIC: return f3(0.0f, 0.0f, 0.0f);  // dielectric.cl:58-61
        case BXDF_ROUGH_CONDUCTOR: {                        // rough_conductor.cl:54-78
            float roughness = roughnessOf(s, m, sc);
            float3 ks = matGetSample3f(s.uv, m.u2, m.tex, sc);
            return ggxReflectionEval(roughness, ks, conductorFresnel(m, dot(in, n)), in, out, n);
        }
        case BXDF_ROUGH_DIELECTRIC: {  // rough_dielectric.cl:114-166
            float iDotN = dot(in, n);
            float roughness = roughnessOf(s, m, sc);
            float etaI = m.extIOR, etaT = m.intIOR;
            if (iDotN < 0.0f) { float t = etaI; etaI = etaT; etaT = t; }
            float f = fresnelForDielectric(etaI, etaT, iDotN);
            if (iDotN > 0.0f) {
                float3 ks = matGetSample3f(s.uv, m.u2, m.tex, sc);
                return ggxReflectionEval(roughness, ks, f, in, out, n);
            }
            float3 h = normalize(-(etaI * in + etaT * out));
            return ggxRefractionEval(roughness, matGetSample3f(s.uv, m.u3, m.rightOrTransTex, sc), f, etaI, etaT, in, out, h, n);
        }
    }
    return f3(0.0f, 0.0f, 0.0f);
}

// ------------------------------------------------------------------------------------------------
// samplers/emissive_sampler.cl
// ------------------------------------------------------------------------------------------------
struct Emissive {  // types.cl:167-187
    float4 m0, m1, m2, m3;
    float area;
    uint32_t triIndex, matNodeIndex, type;
};
PC_HD Emissive loadEmissive(const DScene &sc, uint32_t i) {
    Emissive e;
    const float4 *p = sc.emissives + 5 * (size_t)i;
    e.m0 = PC_LDG(p); e.m1 = PC_LDG(p + 1); e.m2 = PC_LDG(p + 2); e.m3 = PC_LDG(p + 3);
    float4 q = PC_LDG(p + 4);
    e.area = q.x; e.triIndex = f2u(q.y); e.matNodeIndex = f2u(q.z); e.type = f2u(q.w);
    return e;
}
PC_HD float3 emissiveGetSample(const Surface &s, const Emissive &e, const DScene &sc, float2 r, float3 &outDir, float &pdf,
                               float &dist) {  // :178-201
    if (e.type == 1) {  // environmentLightGetSample (:16-37)
        outDir = cosWeightedHemisphereGetSample(s.normal, r);
        pdf = cl_max(0.0f, dot(s.normal, outDir)) * PC_1_PI;
        dist = FLT_MAX;
        float2 uv = rayToLatLongUV(outDir);
        MatNode m = loadMatNode(sc, e.matNodeIndex);
        return m.scaleOrRoughness * matGetSample3f(uv, m.u2, m.tex, sc) * PC_1_PI;
    }
    if (e.type != 0) return f3(0.0f, 0.0f, 0.0f);
    // areaLightGetSample (:51-113)
    float r1sqrt = sqrtf(r.x);
    float ru = (1.0f - r.y) * r1sqrt;
    float rv = r.y * r1sqrt;
    float3 wuv = f3(1.0f - ru - rv, ru, rv);
    size_t off = (size_t)e.triIndex * 3;
    float4 v0 = PC_LDG(sc.vertices + off), v1 = PC_LDG(sc.vertices + off + 1), v2 = PC_LDG(sc.vertices + off + 2);
    float4 n0 = PC_LDG(sc.normals + off), n1 = PC_LDG(sc.normals + off + 1), n2 = PC_LDG(sc.normals + off + 2);
    float2 t0 = PC_LDG(sc.uvs + off), t1 = PC_LDG(sc.uvs + off + 1), t2 = PC_LDG(sc.uvs + off + 2);
    float3 ePoint = mul4x1(xyz(wuv.x * v0 + wuv.y * v1 + wuv.z * v2), e.m0, e.m1, e.m2, e.m3);
    float3 eNormal = mul4x1(xyz(wuv.x * n0 + wuv.y * n1 + wuv.z * n2), e.m0, e.m1, e.m2, e.m3);  // with translation, SURVEY Q7
    float2 eUV = wuv.x * t0 + wuv.y * t1 + wuv.z * t2;
    MatNode m = loadMatNode(sc, e.matNodeIndex);
    float3 eRay = ePoint - s.point;
    float sqDist = dot(eRay, eRay);
    outDir = normalize(eRay);
    dist = sqrtf(sqDist);
    float nDotOut = dot(eNormal, -outDir);
    if (nDotOut > 0.0f) {
        pdf = 1.0f / e.area;  // area measure; the geometry term is folded into the radiance
        float3 ke = matGetSample3f(eUV, m.u2, m.tex, sc);
        return m.scaleOrRoughness * ke * nDotOut / sqDist;
    }
    pdf = 0.0f;
    return f3(0.0f, 0.0f, 0.0f);
}
PC_HD float emissiveGetPdf(const Surface &s, const Emissive &e, const DScene &sc, float3 outDir) {  // :204-224
    if (e.type == 1) return cl_max(0.0f, dot(s.normal, outDir) * PC_1_PI);  // :39-47
    if (e.type != 0) return 0.0f;
    // areaLightGetPdf (:117-173)
    size_t off = (size_t)e.triIndex * 3;
    float3 v0 = xyz(PC_LDG(sc.vertices + off));
    float3 edge01 = xyz(PC_LDG(sc.vertices + off + 1)) - v0;
    float3 edge02 = xyz(PC_LDG(sc.vertices + off + 2)) - v0;
    v0 = mul4x1(v0, e.m0, e.m1, e.m2, e.m3);
    edge01 = mul4x1(edge01, e.m0, e.m1, e.m2, e.m3);
    edge02 = mul4x1(edge02, e.m0, e.m1, e.m2, e.m3);
    float3 pVec = cross(outDir, edge02);
    float det = dot(edge01, pVec);
    if (fabsf(det) < PC_EPS) return 0.0f;
    float invDet = 1.0f / det;
    float3 tVec = s.point - v0;
    float u = dot(tVec, pVec) * invDet;
    if (u < 0.0f || u > 1.0f) return 0.0f;
    float3 qVec = cross(tVec, edge01);
    float v = dot(outDir, qVec) * invDet;
    if (v < 0.0f || u + v > 1.0f) return 0.0f;
    float t = dot(edge02, qVec) * invDet;
    if (t < PC_EPS) return 0.0f;
    float3 eNormal = normalize(cross(edge01, edge02));
    float denominator = e.area * fabsf(dot(eNormal, outDir));
    return denominator > 0.0f ? (t * t) / denominator : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// kernels/camera.cl:5-58 for one pixel of the block
// ------------------------------------------------------------------------------------------------
struct CameraParams {
    float4 frustrumTL, frustrumTR, frustrumBL, frustrumBR;
    float3 eye;
    float2 texelDims;  // (1/FrameW, 1/FrameH), resources.go:130-133
};
PC_HD float3 primaryRayDir(const CameraParams &cam, uint32_t gx, uint32_t gy, uint32_t blockY, uint32_t randSeed) {
    uint2 rnd = make_uint2(gx + randSeed, gy + randSeed);
    float2 s0 = randomGetSample2f(rnd);
    float2 offset = make_float2(s0.x < 0.5f ? sqrtf(2.0f * s0.x) - 0.5f : 1.5f - sqrtf(2.0f - 2.0f * s0.x),
                                s0.y < 0.5f ? sqrtf(2.0f * s0.y) - 0.5f : 1.5f - sqrtf(2.0f - 2.0f * s0.y));
    float2 texel = (make_float2((float)gx, (float)(gy + blockY)) + offset) * cam.texelDims;
    float4 dir = normalize(mix(mix(cam.frustrumTL, cam.frustrumBL, texel.y), mix(cam.frustrumTR, cam.frustrumBR, texel.y), texel.x));
    return xyz(dir);
}

// ------------------------------------------------------------------------------------------------
// kernels/hdr.cl:5-28
// ------------------------------------------------------------------------------------------------
PC_HD uchar4 tonemapReinhard(float4 acc, float sampleWeight, float exposure) {
    float3 hdr = xyz(acc) * sampleWeight * exposure;
    float3 mapped = hdr / (hdr + 1.0f);
    const float e = 1.0f / 2.2f;
    float3 p = f3(powf(mapped.x, e), powf(mapped.y, e), powf(mapped.z, e));
    float3 o = f3(cl_clamp(p.x, 0.0f, 1.0f), cl_clamp(p.y, 0.0f, 1.0f), cl_clamp(p.z, 0.0f, 1.0f)) * 255.0f;
    return make_uchar4((unsigned char)o.x, (unsigned char)o.y, (unsigned char)o.z, 255);
}

// ------------------------------------------------------------------------------------------------
// Ray / box and ray / triangle tests (kernels/intersect.cl)
// ------------------------------------------------------------------------------------------------
// Slab test of intersect.cl:302-309: returns the entry distance, or FLT_MAX when the reference
// would not descend (exit < 0, entry > exit, or entry >= the RAY's max distance).  fminf/fmaxf
// return the non-NaN operand like OpenCL's fmin/fmax (0*inf appears for axis-parallel rays).
PC_HD float slabEntry(float3 bmin, float3 bmax, float3 o, float3 invDir, float tmaxRay) {
    float3 t1 = (bmin - o) * invDir;
    float3 t2 = (bmax - o) * invDir;
    float minmax = fminf(fminf(fmaxf(t1.x, t2.x), fmaxf(t1.y, t2.y)), fmaxf(t1.z, t2.z));
    float maxmin = fmaxf(fmaxf(fminf(t1.x, t2.x), fminf(t1.y, t2.y)), fminf(t1.z, t2.z));
    return (minmax < 0 || maxmin > minmax) ? FLT_MAX : (maxmin >= tmaxRay ? FLT_MAX : maxmin);
}

// Moeller-Trumbore exactly as intersect.cl:253-280 on precomputed edges; returns false where the
// reference `continue`s.  t is not range-checked here.
PC_HD bool triTest(float3 v0, float3 e1, float3 e2, float3 o, float3 d, float &u, float &v, float &t) {
    float3 pVec = cross(d, e2);
    float det = dot(e1, pVec);
    if (fabsf(det) < PC_EPS) return false;
    float3 tVec = o - v0;
    float a = dot(tVec, pVec);
#ifdef PC_TRI_PRETEST  // MEASURED AND REJECTED: -4 % on config 2 (3 818 -> 3 672 Mrays/s, profiles/ab_r01h.txt), kept for the record
    // Division-free early outs, decided only when the reference's own test `u < 0 || u > 1` on
    // u = RN(a * RN(1/det)) is certain to reject (the IEEE reciprocal is ~10 instructions and most tested
    // triangles die here -- but the two extra compares and the extra divergent branch cost more than it saves):
    //   signs differ, |a| >= 2^-20: |RN(1/det)| >= 2^-129 (no flush to zero), so the product is a nonzero negative
    //     number after rounding -> u < 0;
    //   signs equal, |a| > |det| * (1 + 1e-5): a/det > 1 + 9.9e-6 and the two roundings take at most 1.2e-7
    //     of it -> u > 1.
    // Everything else (including NaN / inf operands, for which both comparisons are false) takes the exact path.
    {
        const bool neg = (a < 0.0f) != (det < 0.0f);
        const float aa = fabsf(a);
        if (neg ? aa >= 9.5367431640625e-07f : aa > fabsf(det) * 1.00001f) return false;
    }
#endif
    float invDet = 1.0f / det;
    u = a * invDet;
    if (u < 0.0f || u > 1.0f) return false;
    float3 qVec = cross(tVec, e1);
    v = dot(d, qVec) * invDet;
    if (v < 0.0f || u + v > 1.0f) return false;
    t = dot(e2, qVec) * invDet;
    return true;
}

struct Hit {  // Intersection (types.cl:69-83) + the tie-break key
    float4 wuvt;
    uint32_t inst, tri, rank;
};
struct TravStats {
    uint32_t nodes, tris, instances;
};

#ifndef PC_STACK_SIZE
#define PC_STACK_SIZE 64
#endif
// Closest-hit culling slack: a child box is skipped only when its entry distance exceeds the
// best hit by more than this relative margin, so triangles whose computed t ties (or nearly
// ties) with the current best are still tested and the winner is decided by exact comparison
// of t plus the reference-order key -- never by the rounding of the box test.
#define PC_CULL_SLACK 1.0001f

// Where the traversal stack lives.  The reference keeps 32 uints of private memory per work-item (intersect.cl:4,42,201);
// a CUDA local array indexed by a run-time stack pointer is the same thing: every push / pop is an LSU round trip through
// L1 that competes with the node fetches.  PC_SMEM_STACK > 0 keeps the first PC_SMEM_STACK entries of each thread's stack
// in SHARED memory, one 4-byte column per thread (entry i of thread t at word i * blockDim + t: a warp's accesses to one
// level hit 32 different banks), and only the rare deeper entries spill to a local array.  The trav* steps below take the
// stack as a template parameter: a plain uint32_t* (host build, tests, PC_SMEM_STACK == 0) or a SmemStack.
#ifndef PC_SMEM_STACK
#define PC_SMEM_STACK 12
#endif
PC_HD void stackPush(uint32_t *s, int &sp, uint32_t v) { s[sp++] = v; }
PC_HD uint32_t stackPop(uint32_t *s, int &sp) { return s[--sp]; }
#if defined(__CUDACC__)
template <int DEPTH, int STRIDE>
struct SmemStack {
    uint32_t *sm;     // this thread's column of the CTA's shared stack
    uint32_t *spill;  // PC_STACK_SIZE - DEPTH local entries
};
template <int DEPTH, int STRIDE>
__device__ __forceinline__ void stackPush(SmemStack<DEPTH, STRIDE> s, int &sp, uint32_t v) {
    if (sp < DEPTH) s.sm[sp * STRIDE] = v;
    else s.spill[sp - DEPTH] = v;
    sp++;
}
template <int DEPTH, int STRIDE>
__device__ __forceinline__ uint32_t stackPop(SmemStack<DEPTH, STRIDE> s, int &sp) {
    --sp;
    return sp < DEPTH ? s.sm[sp * STRIDE] : s.spill[sp - DEPTH];
}
#endif

// PC_POP_CULL: a closest-hit walk pushes the far child's ENTRY DISTANCE next to its reference (two stack words per entry) and
// tests it again when the entry is popped: by then the near subtree has usually produced a closer hit, and without the test
// a popped leaf reference goes straight to its (up to 10) triangle tests and a popped inner node costs a fetch and two slab
// tests before its children are culled.  Same predicate as at push time (entry > best * PC_CULL_SLACK), so the hit record is
// unchanged: a culled subtree cannot hold a hit with t <= best.  Instance markers carry distance 0 and are never culled.
// Any-hit walks have no "best so far" and keep one-word entries.
// MEASURED AND REJECTED (round 2, profiles/ab_r02l.txt): with nearest-first order and culling at push time there is little
// left to cull at pop time -- on real bounce rays (host build of this header) 1 % fewer node / triangle steps on config 2,
// 8 % on config 3, 2 % on config 4, hit records identical -- and the second stack word plus the pop loop cost more:
// k_trace 3 681 -> 3 885 us on config 3, 1 017 -> 1 064 us on config 4, unchanged on config 2.  Compiled out.
#ifndef PC_POP_CULL
#define PC_POP_CULL 0
#endif
// Stack-based traversal of the derived layout, written as a per-ray state machine so that the same
// three steps serve the plain loop below (host build / tests) and the persistent kernels, which
// interleave them with warp-level refilling of finished lanes (pc_kernels.cuh).
//   ANY_HIT == false : rayIntersectionQuery semantics (intersect.cl:184-347) -- closest hit; the
//                      result equals the reference's left-first walk because ties on t go to the
//                      smaller (instance dfsRank, triangle index), the order that walk visits.
//   ANY_HIT == true  : rayIntersectionTest semantics (:26-180) -- same cull predicate as the
//                      reference (entry >= ray tmax), so the boolean is identical in any order.
// Children are visited nearest first.
struct Trav {
    float3 o0, d0;      // the ray as given (world space)
    float3 o, d, invDir;  // the ray in the current (world or mesh) space
    float tmaxRay;
    uint32_t curInst, curRank;
    uint32_t cur;       // current reference; REF_DONE when the walk is over
    int sp;
    Hit best;
};

PC_HD void travInit(Trav &t, const DScene &sc, float3 o0, float3 d0, float tmaxRay) {
    t.o0 = o0; t.d0 = d0; t.o = o0; t.d = d0;
    t.invDir = f3(1.0f / d0.x, 1.0f / d0.y, 1.0f / d0.z);  // native_recip(ray.dir) (:302)
    t.tmaxRay = tmaxRay;
    t.curInst = 0; t.curRank = 0;
    t.best.wuvt = make_float4(0.0f, 0.0f, 0.0f, tmaxRay);
    t.best.inst = 0; t.best.tri = 0; t.best.rank = 0;
    t.cur = sc.rootRef;
    t.sp = 0;
}

template <bool ANY_HIT, class Stack>
PC_HD void travPush(Trav &t, Stack stack, uint32_t ref, float entry) {
    stackPush(stack, t.sp, ref);
#if PC_POP_CULL
    if (!ANY_HIT) stackPush(stack, t.sp, f2u(entry));
#endif
}
// next reference to visit, REF_DONE when the stack is exhausted
template <bool ANY_HIT, class Stack>
PC_HD uint32_t travPop(Trav &t, Stack stack) {
#if PC_POP_CULL
    if (!ANY_HIT) {
        const float lim = t.best.wuvt.w * PC_CULL_SLACK;
        while (t.sp) {
            const float entry = u2f(stackPop(stack, t.sp));
            const uint32_t ref = stackPop(stack, t.sp);
            if (!(entry > lim)) return ref;
        }
        return REF_DONE;
    }
#endif
    return t.sp ? stackPop(stack, t.sp) : REF_DONE;
}

// Reference classes: inner node (bit 31 clear), triangle leaf (bits 31:30 == 10), and the "other"
// leaf-type references with bits 31:30 == 11 (instance entry, the exit marker, REF_DONE).
PC_HD bool refIsTriLeaf(uint32_t c) { return (c >> 30) == 2u; }

// One inner-node step: slab-test both children (one aligned 64 B record), descend into the nearer
// accepted child, push the other.  Requires !(t.cur & REF_LEAF).
template <bool ANY_HIT, bool COUNT, class Stack>
PC_HD void travInner(Trav &t, const DScene &sc, Stack stack, TravStats &st) {
    if (COUNT) st.nodes++;
    const float4 *np = sc.node64 + 4 * (size_t)t.cur;
    float4 q0 = PC_LDG(np), q1 = PC_LDG(np + 1), q2 = PC_LDG(np + 2), q3 = PC_LDG(np + 3);
    float tl = slabEntry(xyz(q0), xyz(q1), t.o, t.invDir, t.tmaxRay);
    float tr = slabEntry(xyz(q2), xyz(q3), t.o, t.invDir, t.tmaxRay);
    bool wl = tl < FLT_MAX, wr = tr < FLT_MAX;
    if (!ANY_HIT) {
        float lim = t.best.wuvt.w * PC_CULL_SLACK;
        wl = wl && !(tl > lim);
        wr = wr && !(tr > lim);
    }
    uint32_t lref = f2u(q0.w), rref = f2u(q1.w);
    if (wl && wr) {
        bool leftFirst = tl <= tr;
        const uint32_t farRef = leftFirst ? rref : lref;
        travPush<ANY_HIT>(t, stack, farRef, leftFirst ? tr : tl);
        t.cur = leftFirst ? lref : rref;
#if defined(__CUDA_ARCH__) && defined(PC_PREFETCH_FAR)
        // the far child is fetched when it is popped, many steps later: ask for its record now (experiment)
        if (!(farRef & REF_LEAF)) asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.node64 + 4 * (size_t)farRef));
        else if (refIsTriLeaf(farRef)) asm volatile("prefetch.global.L1 [%0];" ::"l"(sc.tri48 + 3 * (size_t)(farRef & 0x3FFFFFFFu)));
#endif
    } else if (wl || wr) {
        t.cur = wl ? lref : rref;
    } else {
        t.cur = travPop<ANY_HIT>(t, stack);
    }
}


#ifdef PC_WIDE_BVH
// One step over a 4-ary node (pc_layout.hpp build_wide): four slab tests on one 128 B record, the accepted children sorted
// by entry distance with a 5-comparator network, nearest next, the others pushed far to near.  EXPERIMENT, default off:
// modelled on real bounce rays (tools/wide_bvh_model.py) it halves the node iterations of a warp at +4-10 % instructions;
// hit records are bit-identical to the binary walk's.  Requires !(t.cur & REF_LEAF).
PC_HD void wideSwap(float &ea, uint32_t &ra, float &eb, uint32_t &rb) {
    const bool s = eb < ea;
    const float e = s ? eb : ea, f = s ? ea : eb;
    const uint32_t r = s ? rb : ra, q = s ? ra : rb;
    ea = e; eb = f; ra = r; rb = q;
}
template <bool ANY_HIT, bool COUNT, class Stack>
PC_HD void travInnerWide(Trav &t, const DScene &sc, Stack stack, TravStats &st) {
    if (COUNT) st.nodes++;
    const float4 *np = sc.node128 + 8 * (size_t)t.cur;
    const float4 q0 = PC_LDG(np), q1 = PC_LDG(np + 1), q2 = PC_LDG(np + 2), q3 = PC_LDG(np + 3);
    const float4 q4 = PC_LDG(np + 4), q5 = PC_LDG(np + 5), q6 = PC_LDG(np + 6), q7 = PC_LDG(np + 7);
    const uint32_t n = f2u(q1.w);
    float e0 = slabEntry(xyz(q0), xyz(q1), t.o, t.invDir, t.tmaxRay);
    float e1 = slabEntry(xyz(q2), xyz(q3), t.o, t.invDir, t.tmaxRay);
    float e2 = n > 2u ? slabEntry(xyz(q4), xyz(q5), t.o, t.invDir, t.tmaxRay) : FLT_MAX;
    float e3 = n > 3u ? slabEntry(xyz(q6), xyz(q7), t.o, t.invDir, t.tmaxRay) : FLT_MAX;
    if (!ANY_HIT) {
        const float lim = t.best.wuvt.w * PC_CULL_SLACK;
        if (e0 > lim) e0 = FLT_MAX;
        if (e1 > lim) e1 = FLT_MAX;
        if (e2 > lim) e2 = FLT_MAX;
        if (e3 > lim) e3 = FLT_MAX;
    }
    uint32_t r0 = f2u(q0.w), r1 = f2u(q2.w), r2 = f2u(q4.w), r3 = f2u(q6.w);
    wideSwap(e0, r0, e1, r1);
    wideSwap(e2, r2, e3, r3);
    wideSwap(e0, r0, e2, r2);
    wideSwap(e1, r1, e3, r3);
    wideSwap(e1, r1, e2, r2);  // e0 <= e1 <= e2 <= e3, rejected children (FLT_MAX) last
    if (e0 == FLT_MAX) {
        t.cur = travPop<ANY_HIT>(t, stack);
        return;
    }
    if (e3 < FLT_MAX) travPush<ANY_HIT>(t, stack, r3, e3);
    if (e2 < FLT_MAX) travPush<ANY_HIT>(t, stack, r2, e2);
    if (e1 < FLT_MAX) travPush<ANY_HIT>(t, stack, r1, e1);
    t.cur = r0;
}
#endif

// Instance entry (:237-249) or the exit marker that restores the world-space ray (:330-335).
// Returns 0 to continue, 1 when the walk is over.  Requires bits 31:30 == 11 and cur != REF_DONE.
template <bool ANY_HIT, bool COUNT, class Stack>
PC_HD int travOther(Trav &t, const DScene &sc, Stack stack, TravStats &st) {
    if (t.cur == REF_POP_INSTANCE || t.cur == REF_POP_TRANSLATED) {
        t.o = t.o0;
        if (t.cur == REF_POP_INSTANCE) {
            t.d = t.d0;
            t.invDir = f3(1.0f / t.d.x, 1.0f / t.d.y, 1.0f / t.d.z);
        }
        t.cur = travPop<ANY_HIT>(t, stack);
        return t.cur == REF_DONE ? 1 : 0;
    }
    if (COUNT) st.instances++;
    t.curInst = t.cur & 0x3FFFFFFFu;
    const float4 *ip = sc.inst80 + 5 * (size_t)t.curInst;
    float4 hdr = PC_LDG(ip);
    t.curRank = f2u(hdr.z);
    const uint32_t iflags = f2u(hdr.y);
    if (iflags & INST_FLAG_TRANSLATION) {
        // translation only: mul4x1 is x*1 + y*0 + z*0 + t == x + t and mul3x1 leaves the direction (and
        // therefore 1/d) untouched, exactly, for finite inputs -- one 16 B load instead of four and no
        // reciprocals on the way in or out
        float4 m3 = PC_LDG(ip + 4);
        t.o = f3(t.o.x + m3.x, t.o.y + m3.y, t.o.z + m3.z);
        travPush<ANY_HIT>(t, stack, REF_POP_TRANSLATED, 0.0f);
    } else if (!(iflags & INST_FLAG_IDENTITY)) {
        // identity matrices are skipped: x*1 + y*0 + z*0 + 0 == x exactly for finite inputs
        float4 m0 = PC_LDG(ip + 1), m1 = PC_LDG(ip + 2), m2 = PC_LDG(ip + 3), m3 = PC_LDG(ip + 4);
        t.o = mul4x1(t.o, m0, m1, m2, m3);
        t.d = mul3x1(t.d, m0, m1, m2);
        t.invDir = f3(1.0f / t.d.x, 1.0f / t.d.y, 1.0f / t.d.z);
        travPush<ANY_HIT>(t, stack, REF_POP_INSTANCE, 0.0f);
    }
    t.cur = f2u(hdr.x);
    return 0;
}

// One triangle leaf (:251-291).  Returns 0 to continue, 1 when the walk is over, 2 when an any-hit
// ray found its occluder.  Requires refIsTriLeaf(t.cur).
template <bool ANY_HIT, bool COUNT, class Stack>
PC_HD int travTris(Trav &t, const DScene &sc, Stack stack, TravStats &st) {
    uint32_t tri = t.cur & 0x3FFFFFFFu;
    const float4 *tp = sc.tri48 + 3 * (size_t)tri;
    float4 a = PC_LDG(tp);
    uint32_t count = f2u(a.w);
    for (;;) {
        float4 b = PC_LDG(tp + 1), c = PC_LDG(tp + 2);
        if (COUNT) st.tris++;
        float u, v, tt;
        if (triTest(xyz(a), xyz(b), xyz(c), t.o, t.d, u, v, tt)) {
            if (ANY_HIT) {
                if (tt > PC_EPS && tt < t.tmaxRay) return 2;  // (:120-124)
            } else if (tt > PC_EPS) {                         // (:281)
                float bt = t.best.wuvt.w;
                bool closer = tt < bt;
                // equal t: keep what the reference's visiting order would have kept
                bool tie = (tt == bt) && bt < t.tmaxRay &&
                           (t.curRank < t.best.rank || (t.curRank == t.best.rank && tri < t.best.tri));
                if (closer || tie) {
                    t.best.wuvt = make_float4(1.0f - (u + v), u, v, tt);
                    t.best.tri = tri;
                    t.best.inst = t.curInst;
                    t.best.rank = t.curRank;
                }
            }
        }
        if (--count == 0) break;
        tri++;
        tp += 3;
        a = PC_LDG(tp);
    }
    t.cur = travPop<ANY_HIT>(t, stack);
    return t.cur == REF_DONE ? 1 : 0;
}

// Any leaf-type reference.  Requires t.cur & REF_LEAF.
template <bool ANY_HIT, bool COUNT, class Stack>
PC_HD int travLeaf(Trav &t, const DScene &sc, Stack stack, TravStats &st) {
    if (t.cur == REF_DONE) return 1;
    if (refIsTriLeaf(t.cur)) return travTris<ANY_HIT, COUNT>(t, sc, stack, st);
    return travOther<ANY_HIT, COUNT>(t, sc, stack, st);
}

// The plain loop: inner-node steps and leaf work in separate loops ("while-while"), so a warp
// reconverges on "all lanes test boxes" / "all lanes test triangles".  Returns 1 on hit.
template <bool ANY_HIT, bool COUNT, class Stack>
PC_HD int traverseWith(const DScene &sc, Stack stack, float3 o0, float3 d0, float tmaxRay, Hit &best, TravStats &st) {
    Trav t;
    travInit(t, sc, o0, d0, tmaxRay);
    int r;
    for (;;) {
#ifdef PC_WIDE_BVH
        while (!(t.cur & REF_LEAF)) travInnerWide<ANY_HIT, COUNT>(t, sc, stack, st);
#else
        while (!(t.cur & REF_LEAF)) travInner<ANY_HIT, COUNT>(t, sc, stack, st);
#endif
        r = travLeaf<ANY_HIT, COUNT>(t, sc, stack, st);
        if (r) break;
    }
    best = t.best;
    if (ANY_HIT) return r == 2 ? 1 : 0;
    return best.wuvt.w < tmaxRay ? 1 : 0;  // (:345)
}
template <bool ANY_HIT, bool COUNT>
PC_HD int traverse(const DScene &sc, float3 o0, float3 d0, float tmaxRay, Hit &best, TravStats &st) {
    uint32_t stack[PC_STACK_SIZE];
    return traverseWith<ANY_HIT, COUNT>(sc, &stack[0], o0, d0, tmaxRay, best, st);
}

// Literal restatement of intersect.cl:184-347 / :26-180 on the reference's own 32 B nodes:
// left child first, boxes culled against the ray's max distance only.  PC_OPT_REFERENCE_ORDER.
template <bool ANY_HIT>
PC_HD int traverseReference(const DScene &sc, float3 o0, float3 d0, float tmaxRay, Hit &best) {
    uint32_t stack[PC_STACK_SIZE];
    int stackIndex = 0, meshStart = -1;
    float3 o = o0, d = d0;
    uint32_t instId = 0;
    best.wuvt = make_float4(0.0f, 0.0f, 0.0f, tmaxRay);
    best.inst = 0; best.tri = 0; best.rank = 0;
    float4 c0 = PC_LDG(sc.bvhNodes), c1 = PC_LDG(sc.bvhNodes + 1);  // curNode = bvhNodes[0]
    while (stackIndex > -1) {
        int L = (int)f2u(c0.w), R = (int)f2u(c1.w);
        bool wantLeft = false, wantRight = false;
        float4 l0, l1, r0, r1;
        if (L <= 0) {
            if (R == 0) {
                instId = (uint32_t)(-L);
                const float4 *ip = sc.meshInstances + 5 * (size_t)instId;
                float4 hdr = PC_LDG(ip), m0 = PC_LDG(ip + 1), m1 = PC_LDG(ip + 2), m2 = PC_LDG(ip + 3), m3 = PC_LDG(ip + 4);
                meshStart = stackIndex;
                if (stackIndex < PC_STACK_SIZE) stack[stackIndex] = f2u(hdr.y);  // bvhRoot
                stackIndex++;
                o = mul4x1(o, m0, m1, m2, m3);
                d = mul3x1(d, m0, m1, m2);
            } else {
                int triStart = -L;
                for (int vIndex = triStart * 3; vIndex < (triStart + R) * 3; vIndex += 3) {
                    float3 v0 = xyz(PC_LDG(sc.vertices + vIndex));
                    float3 e1 = xyz(PC_LDG(sc.vertices + vIndex + 1)) - v0;
                    float3 e2 = xyz(PC_LDG(sc.vertices + vIndex + 2)) - v0;
                    float u, v, t;
                    if (!triTest(v0, e1, e2, o, d, u, v, t)) continue;
                    if (ANY_HIT) {
                        if (t > PC_EPS && t < tmaxRay) return 1;
                    } else if (t > PC_EPS && t < best.wuvt.w) {
                        best.wuvt = make_float4(1.0f - (u + v), u, v, t);
                        best.tri = (uint32_t)(vIndex / 3);
                        best.inst = instId;
                    }
                }
            }
        } else {
            l0 = PC_LDG(sc.bvhNodes + 2 * (size_t)L); l1 = PC_LDG(sc.bvhNodes + 2 * (size_t)L + 1);
            r0 = PC_LDG(sc.bvhNodes + 2 * (size_t)R); r1 = PC_LDG(sc.bvhNodes + 2 * (size_t)R + 1);
            float3 invDir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
            wantLeft = slabEntry(xyz(l0), xyz(l1), o, invDir, tmaxRay) < FLT_MAX;
            wantRight = slabEntry(xyz(r0), xyz(r1), o, invDir, tmaxRay) < FLT_MAX;
        }
        if (wantLeft && wantRight) {
            if (stackIndex < PC_STACK_SIZE) stack[stackIndex] = (uint32_t)R;
            stackIndex++;
            c0 = l0; c1 = l1;
        } else if (wantLeft || wantRight) {
            c0 = wantLeft ? l0 : r0;
            c1 = wantLeft ? l1 : r1;
        } else {
            if (stackIndex == meshStart) {
                o = o0; d = d0;
                meshStart = -1;
            }
            if (--stackIndex >= 0) {
                uint32_t n = stack[stackIndex < PC_STACK_SIZE ? stackIndex : PC_STACK_SIZE - 1];
                c0 = PC_LDG(sc.bvhNodes + 2 * (size_t)n);
                c1 = PC_LDG(sc.bvhNodes + 2 * (size_t)n + 1);
            }
        }
    }
    if (ANY_HIT) return 0;
    return best.wuvt.w < tmaxRay ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// kernels/pt_integrator.cl:17-211 -- the per-ray body of shadeHits (everything between the two
// work-group barriers).  The caller owns compaction and the stores.
// ------------------------------------------------------------------------------------------------
struct ShadeOut {
    float3 accumAdd;      // implicit light for an emissive hit (:103-107)
    bool accum;
    bool wantOcc, wantInd;
    float3 occOrigin, occDir, occSample;
    float occMaxDist;
    float3 indOrigin, indDir;
    float3 newThroughput;  // valid when wantInd (pathSetThroughput, :175)
    uint32_t pathFlags;    // possibly updated by a disperse node
    bool flagsChanged;
};

PC_HD void shadeHit(const DScene &sc, float3 rayDir, float3 curPathThroughput, uint32_t pathFlagsIn, float4 wuvt,
                    uint32_t triIndex, uint32_t rayIndex, uint32_t bounce, uint32_t minBouncesForRR, uint32_t randSeed,
                    ShadeOut &out) {
    out.accum = false; out.wantOcc = false; out.wantInd = false; out.flagsChanged = false;
    out.pathFlags = pathFlagsIn;
    float3 bxdfTint = f3(1.0f, 1.0f, 1.0f);
    float3 bxdfOutRayDir = f3s(0.0f), emissiveOutRayDir = f3s(0.0f), emissiveSample = f3s(0.0f);
    float bxdfPdf = 1.0f, bxdfWeight = 1.0f;
    float emissivePdf = 0.0f, emissiveSelectionPdf = 0.0f, emissiveWeight = 0.0f, distToEmissive = 0.0f;

    uint2 rnd = make_uint2(randSeed, rayIndex);  // (:81)
    float2 sample0 = randomGetSample2f(rnd);
    float2 sample1 = randomGetSample2f(rnd);
    float2 sample2 = randomGetSample2f(rnd);

    float3 inRayDir = -rayDir;
    Surface surface;
    surfaceInit(surface, wuvt, triIndex, sc);
    uint32_t flags = pathFlagsIn;
    MatNode materialNode = matSelectNode(flags, surface, bxdfTint, sc, rnd);
    if (flags != pathFlagsIn) {
        out.pathFlags = flags;
        out.flagsChanged = true;
    }
    float inRayDotNormal = dot(inRayDir, surface.normal);
    if (materialNode.type == BXDF_EMISSIVE) {
        if (inRayDotNormal > 0.0f) {
            out.accum = true;
            out.accumAdd = curPathThroughput * materialNode.scaleOrRoughness * matGetSample3f(surface.uv, materialNode.u2, materialNode.tex, sc);
        }
        return;
    }
    bool rejectSample = materialNode.type == BXDF_INVALID;
    if (bounce >= minBouncesForRR) {  // Russian roulette (:113-124)
        float rrProbability = cl_max(cl_min(0.5f, 0.2126f * curPathThroughput.x + 0.7152f * curPathThroughput.y + 0.0722f * curPathThroughput.z), 0.01f);
        if (rrProbability < sample2.x) rejectSample = true;
        else curPathThroughput = curPathThroughput / rrProbability;
    }
    if (rejectSample) return;

    float3 bxdfSample = bxdfGetSample(surface, materialNode, sc, sample0, inRayDir, bxdfOutRayDir, bxdfPdf);
    float displaceDir = cl_sign(dot(surface.normal, bxdfOutRayDir));
    float3 outBxdfRayOrigin = surface.point + surface.normal * displaceDir * PC_EPS;  // (:135-136)
    float3 outEmissiveRayOrigin = surface.point + surface.normal * PC_EPS;            // (:138)

    if (sc.numEmissives > 0) {  // (:141-155)
        emissiveSelectionPdf = 1.0f / (float)(int)sc.numEmissives;  // emissiveSelect (:227-237)
        int emissiveIndex = cl_clampi((int)(sample1.x * (int)sc.numEmissives), 0, (int)sc.numEmissives - 1);
        Emissive em = loadEmissive(sc, (uint32_t)emissiveIndex);
        emissiveSample = emissiveGetSample(surface, em, sc, sample1, emissiveOutRayDir, emissivePdf, distToEmissive);
        float bxdfEmissivePdf = bxdfGetPdf(surface, materialNode, sc, inRayDir, emissiveOutRayDir);
        emissiveWeight = (emissivePdf * emissivePdf) / (emissivePdf * emissivePdf + bxdfEmissivePdf * bxdfEmissivePdf);
        float emissiveBxdfPdf = emissiveGetPdf(surface, em, sc, bxdfOutRayDir);
        bxdfWeight = (bxdfPdf * bxdfPdf) / (bxdfPdf * bxdfPdf + emissiveBxdfPdf * emissiveBxdfPdf);
    }
    float nDotEmissiveOutRay = cl_max(0.0f, dot(surface.normal, emissiveOutRayDir));
    if (max3(emissiveSample) > 0.0f && emissivePdf > 0.0f && nDotEmissiveOutRay > 0.0f) {  // (:159-163)
        float3 bxdfEmissiveSample = bxdfEval(surface, materialNode, sc, inRayDir, emissiveOutRayDir);
        emissiveSample = emissiveSample * (emissiveWeight * bxdfEmissiveSample * curPathThroughput * nDotEmissiveOutRay / (emissivePdf * emissiveSelectionPdf));
        if (max3(emissiveSample) > 0.0f) {
            out.wantOcc = true;
            out.occSample = emissiveSample;
            out.occOrigin = outEmissiveRayOrigin;
            out.occDir = emissiveOutRayDir;
            out.occMaxDist = distToEmissive - PC_LIGHT_EPS;  // (:203)
        }
    }
    if ((materialNode.type & (BXDF_CONDUCTOR | BXDF_DIELECTRIC)) != 0) bxdfWeight = 1.0f;  // (:166-168)
    float3 throughput = bxdfWeight * bxdfSample * bxdfTint * fabsf(dot(surface.normal, bxdfOutRayDir));
    if (max3(throughput) > 0.0f && bxdfPdf > 0.0f) {  // (:174-177)
        out.wantInd = true;
        out.newThroughput = curPathThroughput * throughput / bxdfPdf;
        out.indOrigin = outBxdfRayOrigin;
        out.indDir = bxdfOutRayDir;
    }
}

// shadePrimaryRayMisses / shadeIndirectRayMisses (pt_integrator.cl:214-275): background sample
PC_HD float3 shadeMiss(const DScene &sc, float3 rayDir) {
    MatNode m = loadMatNode(sc, (uint32_t)sc.sceneDiffuseMat);
    float2 uv = rayToLatLongUV(rayDir);
    return matGetSample3f(uv, m.u2, m.tex, sc);
}

}  // namespace pc
