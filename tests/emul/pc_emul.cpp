// pc_emul.cpp -- TEST-ONLY host build of the CUDA tracer's per-ray device functions.
//
// polaris_b200/csrc/pc_device.cuh is written host+device portable.  This harness compiles it
// with g++ and drives it with plain loops so the CPU-only test tier can compare the CUDA
// side's traversal (derived node64/tri48 layout, near-first order, culling, tie-break) and
// shading code with the oracle without a GPU.  With the same libm on both sides the results
// are expected to be bit-identical.  It is not a fallback: nothing in polaris_b200/ loads it.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/polaris_cuda.h"
#include "../../polaris_b200/csrc/pc_device.cuh"
#include "../../polaris_b200/csrc/pc_layout.hpp"

using namespace pc;

namespace {
struct Ray { float4 origin, dir; };
struct PathRec { float4 throughput; uint32_t pixelIndex, flags, r1, r2; };
struct HitRec { float4 wuvt; uint32_t inst, tri, r1, r2; };

struct Emul {
    pc_layout::Layout layout;
    DScene sc{};
    uint32_t W = 0, H = 0;
    CameraParams cam{};
    std::vector<Ray> rays[3];
    std::vector<PathRec> paths;
    std::vector<uint32_t> flags;
    std::vector<HitRec> hits;
    std::vector<float4> emissiveSamples, traceAcc;
    int numRays[3] = {0, 0, 0};
    std::string error;
};
}  // namespace

extern "C" {

void *pe_create(const pc_scene_view *v) {
    auto *e = new Emul();
    pc_layout::Builder b((const pc_layout::RefNode *)v->bvh_nodes, v->bvh_nodes_bytes / 32,
                         (const pc_layout::RefInstance *)v->mesh_instances, v->mesh_instances_bytes / 80,
                         (const pc_layout::Q *)v->vertices, v->vertices_bytes / 16);
    e->layout = b.build();
    e->error = e->layout.error;
#ifdef PC_WIDE_BVH
    if (e->error.empty()) pc_layout::Builder::build_wide(e->layout);
#endif
    DScene &s = e->sc;
    s.node64 = (const float4 *)e->layout.node64.data();
#ifdef PC_WIDE_BVH
    s.node128 = (const float4 *)e->layout.node128.data();
#endif
    s.tri48 = (const float4 *)e->layout.tri48.data();
    s.inst80 = (const float4 *)e->layout.inst80.data();
    s.rootRef = e->layout.root_ref;
    s.bvhNodes = (const float4 *)v->bvh_nodes;
    s.meshInstances = (const float4 *)v->mesh_instances;
    s.vertices = (const float4 *)v->vertices;
    s.normals = (const float4 *)v->normals;
    s.uvs = (const float2 *)v->uvs;
    s.matIndex = (const uint32_t *)v->material_indices;
    s.matNodes = (const float4 *)v->material_nodes;
    s.emissives = (const float4 *)v->emissives;
    s.texMeta = (const uint4 *)v->texture_metadata;
    s.texData = (const uint8_t *)v->texture_data;
    s.numEmissives = (uint32_t)(v->emissives_bytes / 80);
    s.sceneDiffuseMat = v->scene_diffuse_mat_index;
    return e;
}
void pe_destroy(void *h) { delete (Emul *)h; }
const char *pe_error(void *h) { return ((Emul *)h)->error.empty() ? nullptr : ((Emul *)h)->error.c_str(); }
void pe_layout_info(void *h, int *top_depth, int *mesh_depth, int *stack_need, uint32_t *inner_nodes) {
    Emul &e = *(Emul *)h;
    *top_depth = e.layout.top_depth; *mesh_depth = e.layout.mesh_depth; *stack_need = e.layout.stack_need;
    *inner_nodes = (uint32_t)(e.layout.node64.size() / 4);
}

// mode 0: closest hit (derived layout), 1: any hit, 2: closest hit reference order, 3: any hit reference order
int pe_intersect(void *h, const void *rays_in, uint32_t n, int mode, uint32_t *out_flags, void *out_hits, uint64_t *counters) {
    Emul &e = *(Emul *)h;
    if (e.layout.stack_need > PC_STACK_SIZE) return PC_ERR_STACK_DEPTH;
    const Ray *rays = (const Ray *)rays_in;
    HitRec *hits = (HitRec *)out_hits;
    uint64_t nodes = 0, tris = 0, inst = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : nodes, tris, inst)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Hit best;
        TravStats st{0, 0, 0};
        float3 o = xyz(rays[i].origin), d = xyz(rays[i].dir);
        float tmax = rays[i].origin.w;
        int f;
        switch (mode) {
            case 0: f = traverse<false, true>(e.sc, o, d, tmax, best, st); break;
            case 1: f = traverse<true, true>(e.sc, o, d, tmax, best, st); break;
            case 2: f = traverseReference<false>(e.sc, o, d, tmax, best); break;
            default: f = traverseReference<true>(e.sc, o, d, tmax, best); break;
        }
        out_flags[i] = (uint32_t)f;
        if (hits && (mode == 0 || mode == 2)) hits[i] = HitRec{best.wuvt, best.inst, best.tri, 0, 0};
        nodes += st.nodes; tris += st.tris; inst += st.instances;
    }
    if (counters) { counters[0] = nodes; counters[1] = tris; counters[2] = inst; }
    return 0;
}

int pe_resize(void *h, uint32_t w, uint32_t hh) {
    Emul &e = *(Emul *)h;
    e.W = w; e.H = hh;
    size_t px = (size_t)w * hh;
    for (auto &r : e.rays) r.assign(px, Ray{});
    e.paths.assign(px, PathRec{});
    e.flags.assign(px, 0);
    e.hits.assign(px, HitRec{});
    e.emissiveSamples.assign(px, make_float4(0, 0, 0, 0));
    e.traceAcc.assign(px, make_float4(0, 0, 0, 0));
    return 0;
}
int pe_set_camera(void *h, const float eye[3], const float fr[16]) {
    Emul &e = *(Emul *)h;
    e.cam.eye = f3(eye[0], eye[1], eye[2]);
    memcpy(&e.cam.frustrumTL, fr, 64);
    return 0;
}

// The sample / bounce loops of tracer.go:194-247 + pipeline.go:94-213 over the device functions,
// in the fused stage order the CUDA kernels use (miss+hit shading, occlusion+accumulate).
int pe_trace(void *h, pc_block_request *req, const uint32_t *seeds, size_t n_seeds) {
    Emul &e = *(Emul *)h;
    const DScene &sc = e.sc;
    const size_t per = 1 + (size_t)req->num_bounces;
    if (n_seeds < per * req->samples_per_pixel) return PC_ERR_INVALID_ARGUMENT;
    std::fill(e.traceAcc.begin(), e.traceAcc.end(), make_float4(0, 0, 0, 0));
    e.cam.texelDims = make_float2(1.0f / (float)req->frame_w, 1.0f / (float)req->frame_h);
    const uint32_t W = req->frame_w, BH = req->block_h, BY = req->block_y;
    std::vector<ShadeOut> so((size_t)W * BH);
    std::vector<uint8_t> shaded((size_t)W * BH);
    for (uint32_t s = 0; s < req->samples_per_pixel; s++) {
        const uint32_t *ss = seeds + per * s;
        req->seed = ss[0];
        int n = (int)(W * BH);
        e.numRays[0] = n;
#pragma omp parallel for schedule(dynamic, 256)
        for (int i = 0; i < n; i++) {
            uint32_t gx = (uint32_t)i % W, gy = (uint32_t)i / W;
            float3 d = primaryRayDir(e.cam, gx, gy, BY, ss[0]);
            e.rays[0][i].origin = f4(e.cam.eye, FLT_MAX);
            e.rays[0][i].dir = f4(d, (float)i);
            e.paths[i] = PathRec{make_float4(1.0f, 1.0f, 1.0f, 0.0f), (gy + BY) * W + gx, 0, 0, 0};
            Hit best; TravStats st{0, 0, 0};
            e.flags[i] = (uint32_t)traverse<false, false>(sc, e.cam.eye, d, FLT_MAX, best, st);
            e.hits[i] = HitRec{best.wuvt, best.inst, best.tri, 0, 0};
        }
        int a = 0;
        for (uint32_t bounce = 0; bounce < req->num_bounces; bounce++) {
            n = e.numRays[a];
#pragma omp parallel for schedule(dynamic, 256)
            for (int i = 0; i < n; i++) {
                shaded[i] = 0;
                const Ray &r = e.rays[a][i];
                uint32_t pathIdx = (uint32_t)r.dir.w;
                PathRec &p = e.paths[pathIdx];
                if (!e.flags[i]) {
                    if (sc.sceneDiffuseMat != -1) {
                        float3 kd = shadeMiss(sc, xyz(r.dir));
                        float3 add = bounce == 0 ? kd : xyz(p.throughput) * kd;
                        float4 &acc = e.traceAcc[p.pixelIndex];
                        acc.x += add.x; acc.y += add.y; acc.z += add.z;
                    }
                    continue;
                }
                shaded[i] = 1;
                shadeHit(sc, xyz(r.dir), xyz(p.throughput), p.flags, e.hits[i].wuvt, e.hits[i].tri, (uint32_t)i, bounce,
                         req->min_bounces_for_rr, ss[1 + bounce], so[i]);
                if (so[i].flagsChanged) p.flags = so[i].pathFlags;
                if (so[i].accum) {
                    float4 &acc = e.traceAcc[p.pixelIndex];
                    acc.x += so[i].accumAdd.x; acc.y += so[i].accumAdd.y; acc.z += so[i].accumAdd.z;
                }
                if (so[i].wantInd) p.throughput = f4(so[i].newThroughput, p.throughput.w);
            }
            int nOcc = 0, nInd = 0;
            for (int i = 0; i < n; i++) {
                if (!shaded[i]) continue;
                float pathIdx = e.rays[a][i].dir.w;
                if (so[i].wantOcc) {
                    e.emissiveSamples[nOcc] = f4(so[i].occSample, 0.0f);
                    e.rays[2][nOcc] = Ray{f4(so[i].occOrigin, so[i].occMaxDist), f4(so[i].occDir, pathIdx)};
                    nOcc++;
                }
                if (so[i].wantInd) e.rays[1 - a][nInd++] = Ray{f4(so[i].indOrigin, FLT_MAX), f4(so[i].indDir, pathIdx)};
            }
            e.numRays[2] = nOcc;
            e.numRays[1 - a] = nInd;
            std::vector<uint8_t> occl(nOcc);
#pragma omp parallel for schedule(dynamic, 256)
            for (int i = 0; i < nOcc; i++) {
                Hit best; TravStats st{0, 0, 0};
                const Ray &r = e.rays[2][i];
                occl[i] = (uint8_t)traverse<true, false>(sc, xyz(r.origin), xyz(r.dir), r.origin.w, best, st);
            }
            for (int i = 0; i < nOcc; i++) {
                if (occl[i]) continue;
                uint32_t pathIdx = (uint32_t)e.rays[2][i].dir.w;
                float4 &acc = e.traceAcc[e.paths[pathIdx].pixelIndex];
                acc.x += e.emissiveSamples[i].x; acc.y += e.emissiveSamples[i].y; acc.z += e.emissiveSamples[i].z;
            }
            if (bounce + 1 < req->num_bounces) {
                a = 1 - a;
                n = e.numRays[a];
#pragma omp parallel for schedule(dynamic, 256)
                for (int i = 0; i < n; i++) {
                    Hit best; TravStats st{0, 0, 0};
                    const Ray &r = e.rays[a][i];
                    e.flags[i] = (uint32_t)traverse<false, false>(sc, xyz(r.origin), xyz(r.dir), r.origin.w, best, st);
                    e.hits[i] = HitRec{best.wuvt, best.inst, best.tri, 0, 0};
                }
            }
        }
        req->accumulated_samples++;
    }
    return 0;
}

int pe_read_buffer(void *h, int which, void *dst, uint64_t bytes) {
    Emul &e = *(Emul *)h;
    const void *src = nullptr;
    switch (which) {
        case PC_BUF_RAYS0: case PC_BUF_RAYS1: case PC_BUF_RAYS2: src = e.rays[which].data(); break;
        case PC_BUF_PATHS: src = e.paths.data(); break;
        case PC_BUF_HIT_FLAGS: src = e.flags.data(); break;
        case PC_BUF_INTERSECTIONS: src = e.hits.data(); break;
        case PC_BUF_EMISSIVE_SAMPLES: src = e.emissiveSamples.data(); break;
        case PC_BUF_TRACE_ACCUMULATOR: src = e.traceAcc.data(); break;
        case PC_BUF_RAY_COUNTERS: src = e.numRays; break;
        default: return PC_ERR_INVALID_ARGUMENT;
    }
    memcpy(dst, src, bytes);
    return 0;
}

struct BxdfIn { float n[3]; uint32_t matNode; float in[3], p0; float out[3], p1; float rnd[2], uv[2]; };
struct BxdfOut { float sample[3], samplePdf; float dir[3], pdf; float eval[3], p; };
int pe_bxdf(void *h, const void *in_records, uint32_t n, void *out_records) {
    Emul &e = *(Emul *)h;
    const BxdfIn *in = (const BxdfIn *)in_records;
    BxdfOut *out = (BxdfOut *)out_records;
    for (uint32_t i = 0; i < n; i++) {
        Surface s;
        s.point = f3s(0.0f);
        s.normal = f3(in[i].n[0], in[i].n[1], in[i].n[2]);
        s.uv = make_float2(in[i].uv[0], in[i].uv[1]);
        s.matNodeIndex = in[i].matNode;
        MatNode m = loadMatNode(e.sc, in[i].matNode);
        float3 inDir = f3(in[i].in[0], in[i].in[1], in[i].in[2]), outDir = f3(in[i].out[0], in[i].out[1], in[i].out[2]);
        float3 dir = f3s(0.0f);
        float pdf = 1.0f;
        float3 smp = bxdfGetSample(s, m, e.sc, make_float2(in[i].rnd[0], in[i].rnd[1]), inDir, dir, pdf);
        float p = bxdfGetPdf(s, m, e.sc, inDir, outDir);
        float3 ev = bxdfEval(s, m, e.sc, inDir, outDir);
        out[i] = BxdfOut{{smp.x, smp.y, smp.z}, pdf, {dir.x, dir.y, dir.z}, p, {ev.x, ev.y, ev.z}, 0.0f};
    }
    return 0;
}

int pe_tonemap(const float *acc, uint32_t n, float w, float exposure, uint8_t *rgba) {
    for (uint32_t i = 0; i < n; i++) {
        uchar4 c = tonemapReinhard(make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], 0.0f), w, exposure);
        rgba[4 * i] = c.x; rgba[4 * i + 1] = c.y; rgba[4 * i + 2] = c.z; rgba[4 * i + 3] = c.w;
    }
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// SIMT schedule model (a design TOOL, not a test): runs 32 consecutive rays in lockstep through the
// Trav state machine under a given control structure and counts warp-level iterations against
// lane-level useful iterations, so traversal-loop variants can be compared without GPU time.
//   variant 0: while-while, leaf-type references (instance entry / exit marker / triangles) in phase 2
//   variant 1: instance entry / exit handled inside the phase-1 loop
//   variant 2: variant 1 with identity-instance entries free (flattened two-level tree)
//   triCap   : at most this many triangles per lane per round (0 = whole leaf), resuming mid-leaf
//   refill   : refill threshold (0 = fixed 32-ray units)
//   innerMin : leave phase 1 when fewer than this many lanes still search while others hold a leaf
// out[0..7] = warpInner, laneInner, warpTri, laneTri, warpOther, laneOther, rounds, rays
// ------------------------------------------------------------------------------------------------
extern "C" int pe_simt(void *h, const void *rays_in, uint32_t n, int anyHit, int variant, int triCap, int refill, int innerMin, double *out) {
    Emul &e = *(Emul *)h;
    const DScene &sc = e.sc;
    const Ray *rays = (const Ray *)rays_in;
    double warpInner = 0, laneInner = 0, warpTri = 0, laneTri = 0, warpOther = 0, laneOther = 0, rounds = 0;
    struct Lane { Trav t; bool busy; uint32_t stack[PC_STACK_SIZE]; uint32_t triLeft; };
    std::vector<Lane> L(32);
    uint32_t next = 0;
    TravStats st{0, 0, 0};
    auto isTri = [](uint32_t c) { return (c & REF_LEAF) && !(c & REF_TOP) && c != REF_POP_INSTANCE && c != REF_DONE; };
    auto isIdentityInst = [&](uint32_t c) {
        if (!((c & REF_LEAF) && (c & REF_TOP)) || c >= REF_POP_TRANSLATED) return false;
        uint32_t id = c & 0x3FFFFFFFu;
        return (f2u(sc.inst80[5 * (size_t)id].y) & INST_FLAG_IDENTITY) != 0;
    };
    for (auto &l : L) l.busy = false;
    bool exhausted = false;
    for (;;) {
        int idle = 0;
        for (auto &l : L) idle += !l.busy;
        if (idle && !exhausted) {
            for (auto &l : L) {
                if (l.busy) continue;
                if (next >= n) { exhausted = true; break; }
                const Ray &r = rays[next++];
                travInit(l.t, sc, xyz(r.origin), xyz(r.dir), r.origin.w);
                l.busy = true;
            }
            if (next >= n) exhausted = true;
        }
        int live = 0;
        for (auto &l : L) live += l.busy;
        if (!live) break;
        for (;;) {
            rounds++;
            // ---- phase 1
            for (;;) {
                bool anyInner = false, anyOther = false;
                int nInner = 0, nOther = 0;
                for (auto &l : L) {
                    if (!l.busy) continue;
                    uint32_t c = l.t.cur;
                    if (!(c & REF_LEAF)) { anyInner = true; nInner++; }
                    else if (variant >= 1 && c != REF_DONE && !isTri(c)) { anyOther = true; nOther++; }
                }
                if (!anyInner && !anyOther) break;
                if (innerMin > 0) {  // bounded phase 1: stop when only a few lanes are still searching and others wait with a leaf
                    int waiting = 0, searching = nInner + nOther;
                    for (auto &l : L) if (l.busy && isTri(l.t.cur)) waiting++;
                    if (waiting > 0 && searching < innerMin) break;
                }
                for (auto &l : L) {
                    if (!l.busy) continue;
                    uint32_t c = l.t.cur;
                    if (!(c & REF_LEAF)) {
                        if (anyHit) travInner<true, true>(l.t, sc, l.stack, st); else travInner<false, true>(l.t, sc, l.stack, st);
                    } else if (variant >= 1 && c != REF_DONE && !isTri(c)) {
                        bool free_ = variant >= 2 && isIdentityInst(c);
                        int r = anyHit ? travLeaf<true, true>(l.t, sc, l.stack, st) : travLeaf<false, true>(l.t, sc, l.stack, st);
                        if (r) l.t.cur = REF_DONE;
                        if (free_) nOther--;
                    }
                }
                if (anyInner) { warpInner++; laneInner += nInner; }
                if (nOther > 0) { warpOther++; laneOther += nOther; }
            }
            // ---- phase 2
            uint32_t maxTri = 0;
            int nOther = 0;
            for (auto &l : L) {
                if (!l.busy) continue;
                uint32_t c = l.t.cur;
                if (c == REF_DONE) { l.busy = false; continue; }
                if (isTri(c)) {
                    uint32_t before = st.tris;
                    uint32_t tri = c & 0x3FFFFFFFu;
                    uint32_t count = f2u(sc.tri48[3 * (size_t)tri].w);
                    int r;
                    if (triCap > 0 && count > (uint32_t)triCap) {
                        // process triCap triangles by temporarily shortening the leaf: emulate with a patched count
                        // (model only: run the whole leaf but charge triCap now and the rest next round)
                        r = anyHit ? travLeaf<true, true>(l.t, sc, l.stack, st) : travLeaf<false, true>(l.t, sc, l.stack, st);
                        uint32_t done = st.tris - before;
                        // charge in slices of triCap across rounds: approximate as extra rounds of triCap
                        uint32_t slices = (done + triCap - 1) / triCap;
                        laneTri += done;
                        // first slice is part of this round's max; remaining slices cost their own warp iterations at low occupancy
                        if ((uint32_t)triCap > maxTri) maxTri = std::min<uint32_t>(done, triCap) > maxTri ? std::min<uint32_t>(done, triCap) : maxTri;
                        warpTri += (slices - 1) * 0.0;  // accounted below through l.triLeft
                        l.triLeft = done > (uint32_t)triCap ? done - triCap : 0;
                    } else {
                        r = anyHit ? travLeaf<true, true>(l.t, sc, l.stack, st) : travLeaf<false, true>(l.t, sc, l.stack, st);
                        uint32_t done = st.tris - before;
                        laneTri += done;
                        if (done > maxTri) maxTri = done;
                        l.triLeft = 0;
                    }
                    if (r) l.busy = false;
                } else if (c & REF_LEAF) {
                    nOther++;
                    int r = anyHit ? travLeaf<true, true>(l.t, sc, l.stack, st) : travLeaf<false, true>(l.t, sc, l.stack, st);
                    if (r) l.busy = false;
                }  // else: still searching (bounded phase 1), continues next round
            }
            warpTri += maxTri;
            if (nOther) { warpOther++; laneOther += nOther; }
            live = 0;
            for (auto &l : L) live += l.busy;
            if (!live) break;
            if (refill > 0 && !exhausted && live < refill) break;
        }
    }
    out[0] = warpInner; out[1] = laneInner; out[2] = warpTri; out[3] = laneTri; out[4] = warpOther; out[5] = laneOther;
    out[6] = rounds; out[7] = n;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Wide (4-ary) collapse of the same tree -- a DESIGN TOOL like pe_simt, not product code and not a test oracle.
//
// Every binary inner node i of the derived layout gets a wide record holding the boxes and references of its
// GRANDCHILDREN (a child that is a leaf is kept as it is): 2 to 4 children, same reference numbering, so instance roots
// and stack contents need no translation and a walk simply visits every other level.  The boxes tested are the
// reference's own boxes of those nodes; the skipped intermediate boxes contain them, and the float slab test is monotone
// in the box bounds, so a child that passes would have passed its parent too: the set of leaves visited -- and with the
// tie-break key the hit -- is unchanged.  pe_simt_wide runs 32 consecutive rays in lockstep (while-while) through it,
// counts warp-level and lane-level iterations like pe_simt, and returns the hit records so the claim can be checked
// against the binary walk bit for bit.
// out[0..7] = warpWide, laneWide, warpTri, laneTri, warpOther, laneOther, rounds, rays
// ------------------------------------------------------------------------------------------------
namespace {
struct WideNode {
    float3 bmin[4], bmax[4];
    uint32_t ref[4];
    int n;
};
std::vector<WideNode> build_wide(const pc_layout::Layout &L) {
    const size_t inner = L.node64.size() / 4;
    std::vector<WideNode> W(inner);
    auto child = [&](size_t node, int side, float3 &mn, float3 &mx, uint32_t &ref) {
        const pc_layout::Q *q = &L.node64[4 * node];
        const pc_layout::Q &a = q[side ? 2 : 0], &b = q[side ? 3 : 1];
        mn = f3(a.x, a.y, a.z);
        mx = f3(b.x, b.y, b.z);
        ref = f2u(side ? q[1].w : q[0].w);
    };
    for (size_t i = 0; i < inner; i++) {
        WideNode w{};
        w.n = 0;
        for (int side = 0; side < 2; side++) {
            float3 mn, mx;
            uint32_t ref;
            child(i, side, mn, mx, ref);
            if (!(ref & REF_LEAF)) {  // inner child: take ITS children instead
                for (int s2 = 0; s2 < 2; s2++) {
                    child(ref, s2, w.bmin[w.n], w.bmax[w.n], w.ref[w.n]);
                    w.n++;
                }
            } else {
                w.bmin[w.n] = mn; w.bmax[w.n] = mx; w.ref[w.n] = ref;
                w.n++;
            }
        }
        W[i] = w;
    }
    return W;
}

struct WideLane {
    Trav t;
    bool busy = false;
    std::vector<uint32_t> stack;
};

// one wide step of a lane: nearest accepted child next, the others pushed far to near
template <bool ANY_HIT>
void wide_step(WideLane &l, const WideNode &w) {
    Trav &t = l.t;
    float e[4];
    int order[4], m = 0;
    for (int k = 0; k < w.n; k++) {
        e[k] = slabEntry(w.bmin[k], w.bmax[k], t.o, t.invDir, t.tmaxRay);
        bool want = e[k] < FLT_MAX;
        if (!ANY_HIT) want = want && !(e[k] > t.best.wuvt.w * PC_CULL_SLACK);
        if (want) order[m++] = k;
    }
    for (int a = 1; a < m; a++)  // insertion sort by entry distance (stable)
        for (int b = a; b > 0 && e[order[b]] < e[order[b - 1]]; b--) std::swap(order[b], order[b - 1]);
    if (m == 0) {
        if (l.stack.empty()) t.cur = REF_DONE;
        else { t.cur = l.stack.back(); l.stack.pop_back(); }
        return;
    }
    for (int a = m - 1; a >= 1; a--) l.stack.push_back(w.ref[order[a]]);
    t.cur = w.ref[order[0]];
}
}  // namespace

extern "C" int pe_simt_wide(void *h, const void *rays_in, uint32_t n, int anyHit, double *out, uint32_t *out_flags, void *out_hits) {
    Emul &e = *(Emul *)h;
    const DScene &sc = e.sc;
    const Ray *rays = (const Ray *)rays_in;
    HitRec *hits = (HitRec *)out_hits;
    static std::vector<WideNode> W;
    static const void *builtFor = nullptr;
    if (builtFor != h) { W = build_wide(e.layout); builtFor = h; }
    double warpWide = 0, laneWide = 0, warpTri = 0, laneTri = 0, warpOther = 0, laneOther = 0, rounds = 0;
    TravStats st{0, 0, 0};
    auto isTri = [](uint32_t c) { return refIsTriLeaf(c); };
    for (uint32_t base = 0; base < n; base += 32) {
        std::vector<WideLane> L(32);
        uint32_t tmpStack[PC_STACK_SIZE];
        for (uint32_t k = 0; k < 32 && base + k < n; k++) {
            const Ray &r = rays[base + k];
            travInit(L[k].t, sc, xyz(r.origin), xyz(r.dir), r.origin.w);
            L[k].busy = true;
        }
        auto finish = [&](uint32_t k, int res) {
            WideLane &l = L[k];
            l.busy = false;
            const int hit = anyHit ? (res == 2 ? 1 : 0) : (l.t.best.wuvt.w < l.t.tmaxRay ? 1 : 0);
            if (out_flags) out_flags[base + k] = (uint32_t)hit;
            if (hits) { hits[base + k].wuvt = l.t.best.wuvt; hits[base + k].inst = l.t.best.inst; hits[base + k].tri = l.t.best.tri; hits[base + k].r1 = hits[base + k].r2 = 0; }
        };
        for (;;) {
            int live = 0;
            for (auto &l : L) live += l.busy;
            if (!live) break;
            rounds++;
            for (;;) {  // phase 1: wide inner steps until every busy lane holds a leaf-type reference
                int nInner = 0;
                for (auto &l : L) nInner += l.busy && !(l.t.cur & REF_LEAF);
                if (!nInner) break;
                warpWide++;
                laneWide += nInner;
                for (auto &l : L)
                    if (l.busy && !(l.t.cur & REF_LEAF)) {
                        st.nodes++;
                        if (anyHit) wide_step<true>(l, W[l.t.cur]); else wide_step<false>(l, W[l.t.cur]);
                    }
            }
            uint32_t maxTri = 0;
            int nOther = 0;
            for (uint32_t k = 0; k < 32; k++) {  // phase 2: leaves through the product code's own leaf handlers
                WideLane &l = L[k];
                if (!l.busy) continue;
                if (l.t.cur == REF_DONE) { finish(k, 1); continue; }
                // hand the lane's stack to travLeaf (it pops the next reference itself)
                const size_t keep = l.stack.size();
                l.t.sp = 0;
                if (keep) { tmpStack[0] = l.stack.back(); l.t.sp = 1; }
                const uint32_t before = st.tris;
                const bool tri = isTri(l.t.cur);
                int r = anyHit ? travLeaf<true, true>(l.t, sc, tmpStack, st) : travLeaf<false, true>(l.t, sc, tmpStack, st);
                // write back what travLeaf did to the one-entry window: it either popped it (sp == 0) or pushed a marker (sp == 2)
                if (keep) l.stack.pop_back();
                for (int q = 0; q < l.t.sp; q++) l.stack.push_back(tmpStack[q]);
                if (!keep && r == 1 && l.t.sp == 0 && !l.stack.empty()) r = 0;  // (cannot happen: keep == 0 means the stack was empty)
                if (tri) { const uint32_t done = st.tris - before; laneTri += done; if (done > maxTri) maxTri = done; }
                else nOther++;
                if (r) finish(k, r);
            }
            warpTri += maxTri;
            if (nOther) { warpOther++; laneOther += nOther; }
        }
    }
    out[0] = warpWide; out[1] = laneWide; out[2] = warpTri; out[3] = laneTri; out[4] = warpOther; out[5] = laneOther; out[6] = rounds; out[7] = n;
    return 0;
}
