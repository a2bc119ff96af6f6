// scene_compile.hpp -- the geometry half of the reference's scene compiler (asset/compiler/compiler.go:81-231) around a
// pluggable BVH builder.
//
// Two builders produce the SAME tree (bvh_builder.go:124-224 decisions, bit for bit): the host one in scene_compiler.cpp
// (exact binned SAH under OpenMP, libpolaris_scene.so) and the device one in pc_bvh_build.cu (the same binned SAH as
// level-synchronous CUDA kernels, libpolaris_cuda.so).  Everything else -- pre-order flattening, leaf-order triangle layout,
// instance and emissive records -- is this header, shared by both.
#pragma once
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace scenec {

struct BvhNode {  // asset/scene/optimized_scene.go:25-31
    float min[3];
    int32_t ldata;
    float max[3];
    int32_t rdata;
};
static_assert(sizeof(BvhNode) == 32, "BvhNode must be 32 bytes");

struct MeshInstance {  // optimized_scene.go:141-152
    uint32_t mesh_index, bvh_root, pad[2];
    float transform[16];
};
static_assert(sizeof(MeshInstance) == 80, "MeshInstance must be 80 bytes");

struct Emissive {  // optimized_scene.go:121-137
    float transform[16];
    float area;
    uint32_t prim_index, mat_node_index, type;
};
static_assert(sizeof(Emissive) == 80, "EmissivePrimitive must be 80 bytes");

constexpr float kMinSideLength = 1e-3f;  // bvh_builder.go:21
constexpr float kMinSplitStep = 1e-5f;   // bvh_builder.go:26

// Bounded volumes to partition: SoA views owned by the caller.
struct Volumes {
    const float *bmin, *bmax, *center;  // n x 3 each
};

struct TreeNode {
    float min[3], max[3];
    std::unique_ptr<TreeNode> left, right;
    uint32_t first = 0, count = 0;  // leaf: range in the (stably partitioned) work list
    bool leaf = false;
};

struct BuildError {
    std::string msg;
};

// Flatten in the reference's order: inner node appended first, then the whole left
// subtree, then the right one (bvh_builder.go:214-221); leaves fire the callback.
template <class LeafFn>
uint32_t flatten(const TreeNode *t, std::vector<BvhNode> &out, const uint32_t *work, LeafFn &&leaf_fn) {
    BvhNode n;
    std::memcpy(n.min, t->min, 12);
    std::memcpy(n.max, t->max, 12);
    n.ldata = n.rdata = 0;
    if (t->leaf) {
        leaf_fn(n, work + t->first, t->count);
        out.push_back(n);
        return (uint32_t)out.size() - 1;
    }
    uint32_t idx = (uint32_t)out.size();
    out.push_back(n);
    uint32_t l = flatten(t->left.get(), out, work, leaf_fn);
    uint32_t r = flatten(t->right.get(), out, work, leaf_fn);
    out[idx].ldata = (int32_t)l;
    out[idx].rdata = (int32_t)r;
    return idx;
}

// std::vector without the zero fill: the big output arrays (1.3 GB for the 10 M-triangle terrain) are written exactly once,
// every lane of them, by the leaf-order gather
template <class T>
struct NoInitAlloc {
    using value_type = T;
    NoInitAlloc() = default;
    template <class U> NoInitAlloc(const NoInitAlloc<U> &) {}
    T *allocate(size_t n) { return static_cast<T *>(::operator new(n * sizeof(T))); }
    void deallocate(T *p, size_t) { ::operator delete(p); }
    template <class U, class... A> void construct(U *p, A &&...a) {
        if constexpr (sizeof...(A) > 0) ::new ((void *)p) U(std::forward<A>(a)...);  // default construction: leave as is
    }
    template <class U> bool operator==(const NoInitAlloc<U> &) const { return true; }
    template <class U> bool operator!=(const NoInitAlloc<U> &) const { return false; }
};

enum Timing { T_BOUNDS = 0, T_BUILD = 1, T_FLATTEN = 2, T_GATHER = 3, T_TOTAL = 4, T_BUILD_DEVICE = 5, T_COUNT = 8 };

struct Compiled {
    std::vector<BvhNode> nodes;
    std::vector<MeshInstance> instances;
    std::vector<Emissive> emissives;
    std::vector<float, NoInitAlloc<float>> vertices, normals;  // float4 per vertex
    std::vector<float, NoInitAlloc<float>> uvs;                // float2 per vertex
    std::vector<uint32_t, NoInitAlloc<uint32_t>> material_index;
    int top_depth = 0, mesh_depth = 0;
    double timing[T_COUNT] = {};  // seconds: triangle bounds, BVH builds (wall), flatten, leaf-order gather, total; [5] = the
                                  // device builder's own time (uploads + kernels + read-backs), 0 for the host builder
    std::string error;
};

inline double now_seconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
template <class B>
auto builder_device_seconds(const B &b, int) -> decltype(b.device_seconds) { return b.device_seconds; }
template <class B>
double builder_device_seconds(const B &, long) { return 0.0; }



// One triangle soup.  vertices/normals: ntris*9 floats, uvs: ntris*6 floats,
// material: ntris ints (index into mat_root / mat_emissive).
struct RawMesh {
    const float *vertices;
    const float *normals;
    const float *uvs;
    const int32_t *material;
    uint32_t ntris;
};

// One mesh instance as the wavefront reader leaves it (wavefront.go:505-523):
// inverse transform (what compiler.go:191 stores), world AABB and its midpoint.
struct RawInstance {
    uint32_t mesh_index;
    float inv_transform[16];
    float bbox_min[3], bbox_max[3], center[3];
};

enum Buffer { BUF_BVH_NODES = 0, BUF_MESH_INSTANCES = 1, BUF_EMISSIVES = 2, BUF_VERTICES = 3, BUF_NORMALS = 4, BUF_UVS = 5, BUF_MATERIAL_INDEX = 6 };

// BVH only: n volumes (bmin/bmax/center n x 3), leaf size, returns node array + leaf item lists; used directly by the
// tests that pin bvh_builder_test.go's known answers.  leaf.ldata = -(first index into out_order), rdata = count.
// MakeBuilder(volumes, min_leaf_items) -> object with build(std::vector<uint32_t> &work) (partitions `work` in place),
// flatten_into(nodes, work, leaf_fn) (the reference's pre-order, leaves fire the callback) and max_depth().
template <class MakeBuilder>
Compiled *build_bvh_only(const float *bmin, const float *bmax, const float *center, uint32_t n, int min_leaf_items,
                         uint32_t *out_order, MakeBuilder make_builder) {
    auto *c = new Compiled();
    try {
        Volumes v{bmin, bmax, center};
        std::vector<uint32_t> work(n);
        for (uint32_t i = 0; i < n; i++) work[i] = i;
        auto b = make_builder(v, min_leaf_items);
        b.build(work);
        uint32_t off = 0;
        b.flatten_into(c->nodes, work.data(), [&](BvhNode &leaf, const uint32_t *items, uint32_t cnt) {
            leaf.ldata = -(int32_t)off;
            leaf.rdata = (int32_t)cnt;
            for (uint32_t i = 0; i < cnt; i++) out_order[off + i] = items[i];
            off += cnt;
        });
        c->top_depth = b.max_depth();
    } catch (BuildError &e) {
        c->error = e.msg;
    }
    return c;
}

// compiler.go:81-231 (partitionGeometry).  mat_root[m] = root material node of material m (matIndexToMatRoot),
// mat_emissive[m] = emissive leaf of its tree or -1 (emissiveIndexCache).  env_emissive_node: material node of the
// environment light or -1 (compiler.go:214-220).
template <class MakeBuilder>
Compiled *compile_geometry(const RawMesh *meshes, uint32_t n_meshes, const RawInstance *insts, uint32_t n_insts,
                           const int32_t *mat_root, const int32_t *mat_emissive, uint32_t n_materials,
                           int32_t env_emissive_node, MakeBuilder make_builder) {
    auto *c = new Compiled();
    const double t_begin = now_seconds();
    try {
        // --- top-level BVH over instances, one instance per leaf (compiler.go:87-101)
        {
            std::vector<float> bmin(3 * (size_t)n_insts), bmax(3 * (size_t)n_insts), cen(3 * (size_t)n_insts);
            for (uint32_t i = 0; i < n_insts; i++) {
                std::memcpy(&bmin[3 * i], insts[i].bbox_min, 12);
                std::memcpy(&bmax[3 * i], insts[i].bbox_max, 12);
                std::memcpy(&cen[3 * i], insts[i].center, 12);
            }
            Volumes v{bmin.data(), bmax.data(), cen.data()};
            std::vector<uint32_t> work(n_insts);
            for (uint32_t i = 0; i < n_insts; i++) work[i] = i;
            auto b = make_builder(v, 1);
            const double tb = now_seconds();
            b.build(work);
            c->timing[T_BUILD] += now_seconds() - tb;
            c->timing[T_BUILD_DEVICE] += builder_device_seconds(b, 0);
            b.flatten_into(c->nodes, work.data(), [&](BvhNode &leaf, const uint32_t *items, uint32_t) {
                leaf.ldata = -(int32_t)items[0];  // SetMeshIndex(workList[0]) only (:92-99)
                leaf.rdata = 0;
            });
            c->top_depth = b.max_depth();
        }
        size_t total_tris = 0;
        for (uint32_t m = 0; m < n_meshes; m++) total_tris += meshes[m].ntris;
        c->vertices.resize(total_tris * 12);
        c->normals.resize(total_tris * 12);
        c->uvs.resize(total_tris * 6);
        c->material_index.resize(total_tris);

        // --- one BVH per mesh, triangles re-ordered into leaf order (compiler.go:121-179)
        uint32_t prim_offset = 0;
        std::vector<uint32_t> mesh_roots(n_meshes);
        std::vector<Emissive> mesh_emissives;
        std::vector<uint32_t> mesh_emissive_mesh;
        for (uint32_t m = 0; m < n_meshes; m++) {
            const RawMesh &pm = meshes[m];
            const double t0 = now_seconds();
            std::vector<float, NoInitAlloc<float>> bmin(3 * (size_t)pm.ntris), bmax(3 * (size_t)pm.ntris), cen(3 * (size_t)pm.ntris);
#pragma omp parallel for schedule(static)
            for (int64_t t = 0; t < (int64_t)pm.ntris; t++) {
                const float *v = pm.vertices + 9 * t;
                for (int k = 0; k < 3; k++) {
                    // wavefront.go:637-643: AABB of the 3 vertices, centre = vertex centroid
                    float lo = v[3 + k] < v[6 + k] ? v[3 + k] : v[6 + k];  // MinVec3(v1, v2): out=v1; if v2<out
                    lo = lo < v[k] ? lo : v[k];
                    float hi = v[3 + k] > v[6 + k] ? v[3 + k] : v[6 + k];
                    hi = hi > v[k] ? hi : v[k];
                    bmin[3 * t + k] = lo;
                    bmax[3 * t + k] = hi;
                    float s = v[k] + v[3 + k];
                    s = s + v[6 + k];
                    cen[3 * t + k] = s * (float)(1.0 / 3.0);
                }
            }
            Volumes v{bmin.data(), bmax.data(), cen.data()};
            std::vector<uint32_t> work(pm.ntris);
            for (uint32_t i = 0; i < pm.ntris; i++) work[i] = i;
            const double t1 = now_seconds();
            c->timing[T_BOUNDS] += t1 - t0;
            auto b = make_builder(v, 10);  // minPrimitivesPerLeaf (compiler.go:19)
            b.build(work);
            if (b.max_depth() > c->mesh_depth) c->mesh_depth = b.max_depth();
            const double t2 = now_seconds();
            c->timing[T_BUILD] += t2 - t1;
            c->timing[T_BUILD_DEVICE] += builder_device_seconds(b, 0);

            // pre-order flatten (bvh_builder.go:214-221); the leaf callbacks of compiler.go:128-170 are split in two: the
            // callback proper only records where the leaf's triangles go (SetPrimitives), the copies run afterwards, one
            // leaf per OpenMP iteration (every destination range is known and disjoint)
            struct LeafRec { const uint32_t *items; uint32_t cnt, prim; };
            std::vector<LeafRec> leaves;
            std::vector<BvhNode> mesh_nodes;
            b.flatten_into(mesh_nodes, work.data(), [&](BvhNode &leaf, const uint32_t *items, uint32_t cnt) {
                leaf.ldata = -(int32_t)prim_offset;  // SetPrimitives(primOffset, len) (:129)
                leaf.rdata = (int32_t)cnt;
                leaves.push_back(LeafRec{items, cnt, prim_offset});
                prim_offset += cnt;
            });
            const double t3 = now_seconds();
            c->timing[T_FLATTEN] += t3 - t2;
            int bad_material = 0;
#pragma omp parallel for schedule(dynamic, 512)
            for (int64_t li = 0; li < (int64_t)leaves.size(); li++) {
                const LeafRec &lf = leaves[li];
                for (uint32_t i = 0; i < lf.cnt; i++) {
                    const uint32_t t = lf.items[i];
                    const size_t dst = (size_t)lf.prim + i;
                    const float *pv = pm.vertices + 9 * (size_t)t;
                    const float *pn = pm.normals + 9 * (size_t)t;
                    const float *pu = pm.uvs + 6 * (size_t)t;
                    float *ov = &c->vertices[12 * dst];
                    float *on = &c->normals[12 * dst];
                    float *ou = &c->uvs[6 * dst];
                    for (int k = 0; k < 3; k++) {
                        std::memcpy(ov + 4 * k, pv + 3 * k, 12);  // Vec4(0)
                        ov[4 * k + 3] = 0.f;
                        std::memcpy(on + 4 * k, pn + 3 * k, 12);
                        on[4 * k + 3] = 0.f;
                        std::memcpy(ou + 2 * k, pu + 2 * k, 8);
                    }
                    const int32_t mat = pm.material[t];
                    if (mat < 0 || (uint32_t)mat >= n_materials) {
                        bad_material = 1;
                        c->material_index[dst] = 0;
                    } else {
                        c->material_index[dst] = (uint32_t)mat_root[mat];
                    }
                }
            }
            if (bad_material) throw BuildError{"material index out of range"};
            // emissive triangles, in leaf order (:155-165)
            bool any_emissive = false;
            for (uint32_t k = 0; k < n_materials; k++) any_emissive = any_emissive || mat_emissive[k] != -1;
            for (size_t li = 0; any_emissive && li < leaves.size(); li++) {
                const LeafRec &lf = leaves[li];
                for (uint32_t i = 0; i < lf.cnt; i++) {
                    const uint32_t t = lf.items[i];
                    const int32_t em = mat_emissive[pm.material[t]];
                    if (em == -1) continue;
                    const float *pv = pm.vertices + 9 * (size_t)t;
                    float a[3] = {pv[6] - pv[0], pv[7] - pv[1], pv[8] - pv[2]};  // v2-v0
                    float bb[3] = {pv[6] - pv[3], pv[7] - pv[4], pv[8] - pv[5]}; // v2-v1
                    float cx = a[1] * bb[2] - a[2] * bb[1];
                    float cy = a[2] * bb[0] - a[0] * bb[2];
                    float cz = a[0] * bb[1] - a[1] * bb[0];
                    float l2 = cx * cx + cy * cy;
                    l2 = l2 + cz * cz;
                    float len = (float)std::sqrt((double)l2);  // Vec3.Len (vector.go)
                    Emissive e;
                    std::memset(&e, 0, sizeof(e));
                    e.area = 0.5f * len;
                    e.prim_index = lf.prim + i;
                    e.mat_node_index = (uint32_t)em;
                    e.type = 0;  // AreaLight
                    mesh_emissives.push_back(e);
                    mesh_emissive_mesh.push_back(m);
                }
            }
            c->timing[T_GATHER] += now_seconds() - t3;
            int32_t offset = (int32_t)c->nodes.size();  // :173-178
            mesh_roots[m] = (uint32_t)offset;
            for (auto &n : mesh_nodes) {
                if (n.ldata > 0) {  // OffsetChildNodes ignores leaves
                    n.ldata += offset;
                    n.rdata += offset;
                }
            }
            c->nodes.insert(c->nodes.end(), mesh_nodes.begin(), mesh_nodes.end());
        }
        // --- instances (compiler.go:184-192)
        c->instances.resize(n_insts);
        for (uint32_t i = 0; i < n_insts; i++) {
            MeshInstance &mi = c->instances[i];
            std::memset(&mi, 0, sizeof(mi));
            if (insts[i].mesh_index >= n_meshes) throw BuildError{"instance references unknown mesh"};
            mi.mesh_index = insts[i].mesh_index;
            mi.bvh_root = mesh_roots[insts[i].mesh_index];
            std::memcpy(mi.transform, insts[i].inv_transform, 64);
        }
        // --- one emissive per (instance x emissive triangle of its mesh) (:199-211);
        // the reference iterates a Go map here (random order), we use ascending index.
        for (uint32_t i = 0; i < n_insts; i++) {
            for (size_t e = 0; e < mesh_emissives.size(); e++) {
                if (c->instances[i].mesh_index != mesh_emissive_mesh[e]) continue;
                Emissive emp = mesh_emissives[e];
                std::memcpy(emp.transform, c->instances[i].transform, 64);  // the inverse
                c->emissives.push_back(emp);
            }
        }
        if (env_emissive_node != -1) {  // :214-220
            Emissive emp;
            std::memset(&emp, 0, sizeof(emp));
            emp.mat_node_index = (uint32_t)env_emissive_node;
            emp.type = 1;  // EnvironmentLight
            c->emissives.push_back(emp);
        }
    } catch (BuildError &e) {
        c->error = e.msg;
    }
    c->timing[T_TOTAL] = now_seconds() - t_begin;
    return c;
}

inline int compiled_get(Compiled *c, int which, const void **ptr, uint64_t *bytes) {
    switch (which) {
        case BUF_BVH_NODES: *ptr = c->nodes.data(); *bytes = c->nodes.size() * sizeof(BvhNode); return 0;
        case BUF_MESH_INSTANCES: *ptr = c->instances.data(); *bytes = c->instances.size() * sizeof(MeshInstance); return 0;
        case BUF_EMISSIVES: *ptr = c->emissives.data(); *bytes = c->emissives.size() * sizeof(Emissive); return 0;
        case BUF_VERTICES: *ptr = c->vertices.data(); *bytes = c->vertices.size() * 4; return 0;
        case BUF_NORMALS: *ptr = c->normals.data(); *bytes = c->normals.size() * 4; return 0;
        case BUF_UVS: *ptr = c->uvs.data(); *bytes = c->uvs.size() * 4; return 0;
        case BUF_MATERIAL_INDEX: *ptr = c->material_index.data(); *bytes = c->material_index.size() * 4; return 0;
    }
    return 1;
}
}  // namespace scenec
