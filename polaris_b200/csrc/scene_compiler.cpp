// scene_compiler.cpp -- host-side producer of the flat scene buffers the tracer uploads.
//
// polaris compiles a parsed scene into GPU-friendly flat arrays with
// asset/compiler/compiler.go (partitionGeometry, :81-231) and the SAH builder
// asset/compiler/bvh/bvh_builder.go (:124-308).  The CUDA tracer consumes exactly those
// arrays, so to produce "the same compiled scene" without a Go toolchain this file
// re-implements the geometry half of the compiler in C++ (materials, textures and the
// camera are handled by polaris_b200/scene.py, they are tiny).
//
// The algorithm is the reference's; the implementation is not:
//  * the reference scores every candidate split plane with a full pass over the node's
//    items, one goroutine per plane (bvh_builder.go:167-181, O(planes x items)).  Here
//    each item is dropped into the bin between two consecutive planes and the left/right
//    boxes of every plane come from prefix/suffix merges of the bins: O(items + planes).
//    min/max/count are order independent and each score is evaluated once with the
//    reference's float32 expression, so every plane gets the bit-identical score;
//  * ties between equal scores are resolved in iteration order (axis-major, ascending
//    plane) -- one of the orders the reference's channel can deliver;
//  * subtrees are built as OpenMP tasks into a pointer tree and then flattened in the
//    reference's pre-order (node, left subtree, right subtree; bvh_builder.go:214-221),
//    which is also the order in which leaf callbacks fire and triangles are laid out
//    (compiler.go:128-170).
//
// Built as libpolaris_scene.so (plain C ABI, loaded with ctypes).  CPU only.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

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

class Builder {
  public:
    Builder(const Volumes &v, int min_leaf_items) : v_(v), min_leaf_(min_leaf_items) {}

    // work: item indices; scratch: same length. Returns the pointer tree.
    std::unique_ptr<TreeNode> build(std::vector<uint32_t> &work) {
        scratch_.resize(work.size());
        std::unique_ptr<TreeNode> root;
#pragma omp parallel
#pragma omp single
        root = partition(work.data(), scratch_.data(), 0, (uint32_t)work.size(), 0);
        if (!error_.empty()) throw BuildError{error_};
        return root;
    }
    int max_depth() const { return max_depth_; }

  private:
    const Volumes &v_;
    int min_leaf_;
    std::vector<uint32_t> scratch_;
    std::string error_;
    int max_depth_ = 0;

    static float half_area(const float lo[3], const float hi[3]) {
        // side := max.Sub(min); side[0]*side[1] + side[1]*side[2] + side[0]*side[2]
        float s0 = hi[0] - lo[0], s1 = hi[1] - lo[1], s2 = hi[2] - lo[2];
        float a = s0 * s1;
        float b = s1 * s2;
        float c = s0 * s2;
        float ab = a + b;
        return ab + c;
    }

    struct Bin {
        float lo[3], hi[3];
        uint32_t n;
    };
    static void bin_reset(Bin &b) {
        b.lo[0] = b.lo[1] = b.lo[2] = FLT_MAX;
        b.hi[0] = b.hi[1] = b.hi[2] = -FLT_MAX;
        b.n = 0;
    }
    static void bin_merge(Bin &d, const Bin &s) {
        for (int k = 0; k < 3; k++) {
            if (s.lo[k] < d.lo[k]) d.lo[k] = s.lo[k];
            if (s.hi[k] > d.hi[k]) d.hi[k] = s.hi[k];
        }
        d.n += s.n;
    }

    std::unique_ptr<TreeNode> partition(uint32_t *work, uint32_t *scratch, uint32_t first,
                                        uint32_t n, int depth) {
        if (depth > max_depth_) {
#pragma omp critical(pb_depth)
            if (depth > max_depth_) max_depth_ = depth;
        }
        auto node = std::make_unique<TreeNode>();
        for (int k = 0; k < 3; k++) {
            node->min[k] = FLT_MAX;
            node->max[k] = -FLT_MAX;
        }
        for (uint32_t i = 0; i < n; i++) {  // bvh_builder.go:134-139
            const float *lo = v_.bmin + 3 * (size_t)work[first + i];
            const float *hi = v_.bmax + 3 * (size_t)work[first + i];
            for (int k = 0; k < 3; k++) {
                if (lo[k] < node->min[k]) node->min[k] = lo[k];
                if (hi[k] > node->max[k]) node->max[k] = hi[k];
            }
        }
        auto make_leaf = [&]() {
            node->leaf = true;
            node->first = first;
            node->count = n;
            return std::move(node);
        };
        if ((int)n <= min_leaf_) return make_leaf();  // :142-144

        // ScorePartition(workList) (:292-308): the node box is the union just computed.
        float best_score = (float)n * half_area(node->min, node->max);
        int best_axis = -1;
        float best_split = 0.f;
        uint32_t best_left = 0;

        std::vector<float> cand;
        std::vector<Bin> bins;
        for (int axis = 0; axis < 3; axis++) {
            float side = node->max[axis] - node->min[axis];
            if (side < kMinSideLength) continue;  // :157
            float split_step = side / (1024.0f / (float)(depth + 1));  // :162
            if (split_step < kMinSplitStep) continue;                  // :163
            cand.clear();
            for (float p = node->min[axis]; p < node->max[axis];) {  // :167
                cand.push_back(p);
                float next = p + split_step;
                if (!(next > p)) {  // the reference would spin forever here
#pragma omp critical(pb_err)
                    error_ = "bvh: split step below float resolution (reference never terminates); "
                             "keep scene coordinates smaller";
                    return make_leaf();
                }
                p = next;
            }
            const size_t K = cand.size();
            bins.resize(K + 1);
            for (auto &b : bins) bin_reset(b);
            // item -> bin b = number of planes p_k with p_k <= center, so that the item is
            // on the left (center < p_k, :263) of exactly the planes k >= b.
            for (uint32_t i = 0; i < n; i++) {
                uint32_t it = work[first + i];
                float c = v_.center[3 * (size_t)it + axis];
                size_t b = std::upper_bound(cand.begin(), cand.end(), c) - cand.begin();
                Bin &bn = bins[b];
                const float *lo = v_.bmin + 3 * (size_t)it;
                const float *hi = v_.bmax + 3 * (size_t)it;
                for (int k = 0; k < 3; k++) {
                    if (lo[k] < bn.lo[k]) bn.lo[k] = lo[k];
                    if (hi[k] > bn.hi[k]) bn.hi[k] = hi[k];
                }
                bn.n++;
            }
            // suffix[k] = union of bins k..K  (right side of plane k-1)
            std::vector<Bin> suffix(K + 2);
            bin_reset(suffix[K + 1]);
            for (size_t k = K + 1; k-- > 0;) {
                suffix[k] = suffix[k + 1];
                bin_merge(suffix[k], bins[k]);
            }
            Bin left;
            bin_reset(left);
            for (size_t k = 0; k < K; k++) {
                bin_merge(left, bins[k]);  // items with center < cand[k]
                const Bin &right = suffix[k + 1];
                float score;
                if (left.n == 0 || right.n == 0) {
                    score = FLT_MAX;  // :275-277
                } else {
                    float ls = (float)left.n * half_area(left.lo, left.hi);
                    float rs = (float)right.n * half_area(right.lo, right.hi);
                    score = ls + rs;  // :281-282
                }
                if (score < best_score) {  // :186, first seen wins ties
                    best_score = score;
                    best_axis = axis;
                    best_split = cand[k];
                    best_left = left.n;
                }
            }
        }
        if (best_axis < 0) return make_leaf();  // :193-195

        // stable split of the work list (:198-211)
        uint32_t li = 0, ri = best_left;
        for (uint32_t i = 0; i < n; i++) {
            uint32_t it = work[first + i];
            if (v_.center[3 * (size_t)it + best_axis] < best_split)
                scratch[first + li++] = it;
            else
                scratch[first + ri++] = it;
        }
        std::memcpy(work + first, scratch + first, sizeof(uint32_t) * n);
        uint32_t nl = best_left, nr = n - best_left;
        TreeNode *raw = node.get();
        if (n > 4096) {
#pragma omp task shared(raw) firstprivate(work, scratch, first, nl, depth)
            raw->left = partition(work, scratch, first, nl, depth + 1);
#pragma omp task shared(raw) firstprivate(work, scratch, first, nl, nr, depth)
            raw->right = partition(work, scratch, first + nl, nr, depth + 1);
#pragma omp taskwait
        } else {
            raw->left = partition(work, scratch, first, nl, depth + 1);
            raw->right = partition(work, scratch, first + nl, nr, depth + 1);
        }
        return node;
    }
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

struct Compiled {
    std::vector<BvhNode> nodes;
    std::vector<MeshInstance> instances;
    std::vector<Emissive> emissives;
    std::vector<float> vertices, normals;  // float4 per vertex
    std::vector<float> uvs;                // float2 per vertex
    std::vector<uint32_t> material_index;
    int top_depth = 0, mesh_depth = 0;
    std::string error;
};

}  // namespace

extern "C" {

// One triangle soup.  vertices/normals: ntris*9 floats, uvs: ntris*6 floats,
// material: ntris ints (index into mat_root / mat_emissive).
struct ps_mesh {
    const float *vertices;
    const float *normals;
    const float *uvs;
    const int32_t *material;
    uint32_t ntris;
};

// One mesh instance as the wavefront reader leaves it (wavefront.go:505-523):
// inverse transform (what compiler.go:191 stores), world AABB and its midpoint.
struct ps_instance {
    uint32_t mesh_index;
    float inv_transform[16];
    float bbox_min[3], bbox_max[3], center[3];
};

enum ps_buffer {
    PS_BVH_NODES = 0,
    PS_MESH_INSTANCES = 1,
    PS_EMISSIVES = 2,
    PS_VERTICES = 3,
    PS_NORMALS = 4,
    PS_UVS = 5,
    PS_MATERIAL_INDEX = 6
};

// BVH only: n volumes (bmin/bmax/center n x 3), leaf size, returns node array + leaf item
// lists; used directly by the tests that pin bvh_builder_test.go's known answers.
// leaf callback semantics: leaf.ldata = -(first index into out_order), rdata = count.
void *ps_build_bvh(const float *bmin, const float *bmax, const float *center, uint32_t n,
                   int min_leaf_items, uint32_t *out_order /* n */) {
    auto *c = new Compiled();
    try {
        Volumes v{bmin, bmax, center};
        std::vector<uint32_t> work(n);
        for (uint32_t i = 0; i < n; i++) work[i] = i;
        Builder b(v, min_leaf_items);
        auto root = b.build(work);
        uint32_t off = 0;
        flatten(root.get(), c->nodes, work.data(), [&](BvhNode &leaf, const uint32_t *items, uint32_t cnt) {
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

// compiler.go:81-231 (partitionGeometry).  mat_root[m] = root material node of material m
// (matIndexToMatRoot), mat_emissive[m] = emissive leaf of its tree or -1 (emissiveIndexCache).
// env_emissive_node: material node of the environment light or -1 (compiler.go:214-220).
void *ps_compile(const ps_mesh *meshes, uint32_t n_meshes, const ps_instance *insts, uint32_t n_insts,
                 const int32_t *mat_root, const int32_t *mat_emissive, uint32_t n_materials,
                 int32_t env_emissive_node) {
    auto *c = new Compiled();
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
            Builder b(v, 1);
            auto root = b.build(work);
            flatten(root.get(), c->nodes, work.data(), [&](BvhNode &leaf, const uint32_t *items, uint32_t) {
                leaf.ldata = -(int32_t)items[0];  // SetMeshIndex(workList[0]) only (:92-99)
                leaf.rdata = 0;
            });
            c->top_depth = b.max_depth();
        }
        size_t total_tris = 0;
        for (uint32_t m = 0; m < n_meshes; m++) total_tris += meshes[m].ntris;
        c->vertices.assign(total_tris * 12, 0.f);
        c->normals.assign(total_tris * 12, 0.f);
        c->uvs.assign(total_tris * 6, 0.f);
        c->material_index.assign(total_tris, 0);

        // --- one BVH per mesh, triangles re-ordered into leaf order (compiler.go:121-179)
        uint32_t prim_offset = 0;
        std::vector<uint32_t> mesh_roots(n_meshes);
        std::vector<Emissive> mesh_emissives;
        std::vector<uint32_t> mesh_emissive_mesh;
        for (uint32_t m = 0; m < n_meshes; m++) {
            const ps_mesh &pm = meshes[m];
            std::vector<float> bmin(3 * (size_t)pm.ntris), bmax(3 * (size_t)pm.ntris), cen(3 * (size_t)pm.ntris);
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
            Builder b(v, 10);  // minPrimitivesPerLeaf (compiler.go:19)
            auto root = b.build(work);
            if (b.max_depth() > c->mesh_depth) c->mesh_depth = b.max_depth();

            std::vector<BvhNode> mesh_nodes;
            flatten(root.get(), mesh_nodes, work.data(), [&](BvhNode &leaf, const uint32_t *items, uint32_t cnt) {
                leaf.ldata = -(int32_t)prim_offset;  // SetPrimitives(primOffset, len) (:129)
                leaf.rdata = (int32_t)cnt;
                for (uint32_t i = 0; i < cnt; i++) {
                    uint32_t t = items[i];
                    const float *pv = pm.vertices + 9 * (size_t)t;
                    const float *pn = pm.normals + 9 * (size_t)t;
                    const float *pu = pm.uvs + 6 * (size_t)t;
                    float *ov = &c->vertices[12 * (size_t)prim_offset];
                    float *on = &c->normals[12 * (size_t)prim_offset];
                    float *ou = &c->uvs[6 * (size_t)prim_offset];
                    for (int k = 0; k < 3; k++) {
                        std::memcpy(ov + 4 * k, pv + 3 * k, 12);  // Vec4(0)
                        std::memcpy(on + 4 * k, pn + 3 * k, 12);
                        std::memcpy(ou + 2 * k, pu + 2 * k, 8);
                    }
                    int32_t mat = pm.material[t];
                    if (mat < 0 || (uint32_t)mat >= n_materials) throw BuildError{"material index out of range"};
                    c->material_index[prim_offset] = (uint32_t)mat_root[mat];
                    int32_t em = mat_emissive[mat];
                    if (em != -1) {  // :155-165
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
                        e.prim_index = prim_offset;
                        e.mat_node_index = (uint32_t)em;
                        e.type = 0;  // AreaLight
                        mesh_emissives.push_back(e);
                        mesh_emissive_mesh.push_back(m);
                    }
                    prim_offset++;
                }
            });
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
    return c;
}

const char *ps_error(void *h) {
    auto *c = (Compiled *)h;
    return c->error.empty() ? nullptr : c->error.c_str();
}

int ps_get(void *h, int which, const void **ptr, uint64_t *bytes) {
    auto *c = (Compiled *)h;
    switch (which) {
        case PS_BVH_NODES: *ptr = c->nodes.data(); *bytes = c->nodes.size() * sizeof(BvhNode); return 0;
        case PS_MESH_INSTANCES: *ptr = c->instances.data(); *bytes = c->instances.size() * sizeof(MeshInstance); return 0;
        case PS_EMISSIVES: *ptr = c->emissives.data(); *bytes = c->emissives.size() * sizeof(Emissive); return 0;
        case PS_VERTICES: *ptr = c->vertices.data(); *bytes = c->vertices.size() * 4; return 0;
        case PS_NORMALS: *ptr = c->normals.data(); *bytes = c->normals.size() * 4; return 0;
        case PS_UVS: *ptr = c->uvs.data(); *bytes = c->uvs.size() * 4; return 0;
        case PS_MATERIAL_INDEX: *ptr = c->material_index.data(); *bytes = c->material_index.size() * 4; return 0;
    }
    return 1;
}

void ps_depths(void *h, int *top_depth, int *mesh_depth) {
    auto *c = (Compiled *)h;
    *top_depth = c->top_depth;
    *mesh_depth = c->mesh_depth;
}

void ps_free(void *h) { delete (Compiled *)h; }

}  // extern "C"
