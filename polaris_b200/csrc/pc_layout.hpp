// pc_layout.hpp -- host-side derivation of the traversal layout from the reference's scene buffers.
//
// The reference walks 32-byte BvhNodes ({min, L}{max, R}, asset/scene/optimized_scene.go:25-31) and
// fetches BOTH children of every inner node it visits (two scattered 32 B loads,
// CL/kernels/intersect.cl:298-299), then float4 vertices per triangle, computing the two edges on
// the fly (:254-256).  At upload we re-pack the same data, bit for bit, into 16-byte-aligned
// records sized for one vector-load burst each:
//
//   node64   one 64 B record per INNER node holding the boxes and references of its two children
//            q0 = {left.min.xyz , leftRef } q1 = {left.max.xyz , rightRef}
//            q2 = {right.min.xyz, 0       } q3 = {right.max.xyz, 0       }
//            -> one aligned 64 B fetch per traversal step instead of two unrelated 32 B ones;
//   tri48    per triangle {v0.xyz, remaining triangles in this leaf}{v1-v0, 0}{v2-v0, 0}:
//            the edges are the same float32 subtractions the kernel would do, done once;
//            the leaf's triangle count rides in the spare w lane, so a leaf needs no node fetch;
//   inst80   per instance {rootRef, flags, dfsRank, meshIndex} + the 4 matrix columns.
//
// References (uint32): inner node -> index into node64; 0x80000000|firstTri -> mesh leaf;
// 0xC0000000|instance -> top-level leaf.  Inner nodes are renumbered in depth-first order so a
// node's left child is the next record.
//
// dfsRank is the position of the instance's leaf in a left-first walk of the top-level tree:
// together with the triangle index (leaf order == left-first order inside a mesh,
// compiler.go:128-170) it reproduces which of two equal-t hits the reference's left-first,
// strictly-closer traversal keeps, independent of our traversal order.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace pc_layout {

struct RefNode {  // reference BvhNode
    float min[3];
    int32_t l;
    float max[3];
    int32_t r;
};
struct RefInstance {  // reference MeshInstance
    uint32_t mesh_index, bvh_root, pad[2];
    float m[16];
};
struct Q {
    float x, y, z, w;
};
static_assert(sizeof(RefNode) == 32 && sizeof(RefInstance) == 80 && sizeof(Q) == 16, "layout");

constexpr uint32_t REF_LEAF = 0x80000000u;
constexpr uint32_t REF_TOP = 0x40000000u;
constexpr uint32_t REF_POP_INSTANCE = 0xFFFFFFFFu;  // stack marker, never stored in a node
constexpr uint32_t INST_FLAG_IDENTITY = 1u;
constexpr uint32_t INST_FLAG_TRANSLATION = 2u;  // linear part is the identity: only the origin moves

inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

struct Layout {
    std::vector<Q> node64;  // 4 per inner node
    std::vector<Q> node128;  // 8 per inner node: the 4-ary collapse (PC_WIDE_BVH builds only), see build_wide()
    std::vector<Q> tri48;   // 3 per triangle
    std::vector<Q> inst80;  // 5 per instance
    uint32_t root_ref = 0;
    int top_depth = 0;   // max inner nodes on a root->leaf path of the top-level tree
    int mesh_depth = 0;  // same for the deepest mesh tree
    int stack_need = 0;  // entries a near-first traversal can have on its stack
    std::string error;
};

class Builder {
  public:
    // derive_tris = false: the caller builds tri48 itself (the CUDA library does it on the device from the uploaded
    // vertices and BVH nodes, saving a 48 B/triangle host pass and upload); everything else, including the validation of
    // every leaf's triangle range, still happens here.
    Builder(const RefNode *nodes, size_t n_nodes, const RefInstance *inst, size_t n_inst, const Q *vertices,
            size_t n_vertices, bool derive_tris = true)
        : nodes_(nodes), n_nodes_(n_nodes), inst_(inst), n_inst_(n_inst), verts_(vertices), n_tris_(n_vertices / 3),
          derive_tris_(derive_tris) {}

    Layout build() {
        Layout L;
        out_ = &L;
        if (n_nodes_ == 0) {
            L.error = "scene has no BVH nodes";
            return L;
        }
        if (derive_tris_) L.tri48.assign(3 * n_tris_, Q{0, 0, 0, 0});
        L.node64.reserve(2 * n_nodes_ + 4);  // at most (n_nodes - 1) / 2 inner nodes, 4 records each
        for (size_t t = 0; derive_tris_ && t < n_tris_; t++) {
            const Q &v0 = verts_[3 * t], &v1 = verts_[3 * t + 1], &v2 = verts_[3 * t + 2];
            L.tri48[3 * t] = Q{v0.x, v0.y, v0.z, u2f(0)};
            L.tri48[3 * t + 1] = Q{v1.x - v0.x, v1.y - v0.y, v1.z - v0.z, 0.f};  // edge01 (intersect.cl:255)
            L.tri48[3 * t + 2] = Q{v2.x - v0.x, v2.y - v0.y, v2.z - v0.z, 0.f};  // edge02 (:256)
        }
        L.inst80.assign(5 * n_inst_, Q{0, 0, 0, 0});
        inst_seen_.assign(n_inst_, 0);
        // top-level tree first (node 0 is the scene root, compiler.go:91), then every mesh tree
        rank_ = 0;
        L.root_ref = convert(0, /*top=*/true, 0, &L.top_depth);
        if (!L.error.empty()) return L;
        for (size_t i = 0; i < n_inst_; i++) {
            const RefInstance &ri = inst_[i];
            uint32_t root_ref;
            auto it = mesh_root_ref_.find(ri.bvh_root);
            if (it != mesh_root_ref_.end()) {
                root_ref = it->second;
            } else {
                if (ri.bvh_root >= n_nodes_) {
                    L.error = "instance bvhRoot out of range";
                    return L;
                }
                int depth = 0;
                root_ref = convert(ri.bvh_root, /*top=*/false, 0, &depth);
                if (!L.error.empty()) return L;
                if (depth > L.mesh_depth) L.mesh_depth = depth;
                mesh_root_ref_[ri.bvh_root] = root_ref;
            }
            bool ident = true;
            static const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
            for (int k = 0; k < 16; k++) ident = ident && (ri.m[k] == I[k]);
            bool transl = true;  // columns 0-2 and the last row are the identity's (column-major, matrix.go:61-69)
            for (int k = 0; k < 12; k++) transl = transl && (ri.m[k] == I[k]);
            transl = transl && ri.m[15] == 1.0f;
            Q *q = &L.inst80[5 * i];
            q[0].x = u2f(root_ref);
            q[0].y = u2f(ident ? INST_FLAG_IDENTITY : (transl ? INST_FLAG_TRANSLATION : 0u));
            q[0].z = u2f(inst_rank_.count((uint32_t)i) ? inst_rank_[(uint32_t)i] : 0xFFFFFFu);
            q[0].w = u2f(ri.mesh_index);
            for (int c = 0; c < 4; c++) q[1 + c] = Q{ri.m[4 * c], ri.m[4 * c + 1], ri.m[4 * c + 2], ri.m[4 * c + 3]};
        }
        // near-first traversal pushes at most one entry per inner node on the current path, plus
        // one restore marker per instance entered
        L.stack_need = L.top_depth + 1 + L.mesh_depth + 1;
        return L;
    }

    // The 4-ary collapse (experiment, PC_WIDE_BVH): record i holds the boxes and references of binary inner node i's
    // GRANDCHILDREN (a child that is a leaf is kept as it is) -- 2 to 4 children, same reference numbering as node64, so
    // instance roots and stack contents need no translation and a walk visits every other level.
    //   q[2k] = {child k min.xyz, ref k}   q[2k+1] = {child k max.xyz, -}   q[1].w = number of children
    // The boxes are the reference's own; the skipped intermediate boxes contain them and the float slab test is monotone
    // in the bounds, so the set of leaves a ray visits -- and with the tie-break key its hit -- is the binary walk's
    // (tests/test_cpu_golden.py::test_wide_collapse_preserves_hit_records).  A walk pushes up to 3 entries per wide level.
    static void build_wide(Layout &L) {
        const size_t inner = L.node64.size() / 4;
        L.node128.assign(8 * inner, Q{0, 0, 0, 0});
        auto child = [&](size_t node, int side, Q &mn, Q &mx) {
            const Q *q = &L.node64[4 * node];
            mn = q[side ? 2 : 0];
            mx = q[side ? 3 : 1];
            mn.w = side ? q[1].w : q[0].w;  // the child's reference
            mx.w = 0.f;
        };
        for (size_t i = 0; i < inner; i++) {
            Q *w = &L.node128[8 * i];
            uint32_t n = 0;
            for (int side = 0; side < 2; side++) {
                Q mn, mx;
                child(i, side, mn, mx);
                uint32_t ref;
                memcpy(&ref, &mn.w, 4);
                if (!(ref & REF_LEAF)) {
                    for (int s2 = 0; s2 < 2; s2++) {
                        child(ref, s2, w[2 * n], w[2 * n + 1]);
                        n++;
                    }
                } else {
                    w[2 * n] = mn;
                    w[2 * n + 1] = mx;
                    n++;
                }
            }
            w[1].w = u2f(n);
        }
        L.stack_need = 3 * ((L.top_depth + 1) / 2) + 1 + 3 * ((L.mesh_depth + 1) / 2) + 1;
    }

  private:
    const RefNode *nodes_;
    size_t n_nodes_;
    const RefInstance *inst_;
    size_t n_inst_;
    const Q *verts_;
    size_t n_tris_;
    bool derive_tris_ = true;
    Layout *out_ = nullptr;
    uint32_t rank_ = 0;
    std::vector<uint8_t> inst_seen_;
    std::unordered_map<uint32_t, uint32_t> mesh_root_ref_;
    std::unordered_map<uint32_t, uint32_t> inst_rank_;

    // Returns the reference for reference-node `idx`; *depth = inner nodes below (inclusive).
    uint32_t convert(uint32_t idx, bool top, int level, int *depth) {
        Layout &L = *out_;
        if (!L.error.empty()) return 0;
        if (idx >= n_nodes_) {
            L.error = "BVH child index out of range";
            return 0;
        }
        if (level > 4096) {
            L.error = "BVH deeper than 4096 levels (cycle?)";
            return 0;
        }
        const RefNode &n = nodes_[idx];
        if (n.l <= 0) {  // BVH_IS_LEAF (intersect.cl:6)
            *depth = 0;
            if (n.r == 0) {  // top-level leaf -> instance -n.l (intersect.cl:237-239)
                if (!top) {
                    L.error = "mesh BVH contains a top-level (instance) leaf or an empty leaf";
                    return 0;
                }
                uint32_t id = (uint32_t)(-n.l);
                if (id >= n_inst_) {
                    L.error = "top-level leaf references unknown instance";
                    return 0;
                }
                if (!inst_seen_[id]) {
                    inst_seen_[id] = 1;
                    inst_rank_[id] = rank_;
                }
                rank_++;
                return REF_LEAF | REF_TOP | id;
            }
            if (top) {
                L.error = "top-level BVH contains a triangle leaf";
                return 0;
            }
            uint32_t first = (uint32_t)(-n.l), count = (uint32_t)n.r;
            if (n.r < 0 || (size_t)first + count > n_tris_ || first >= REF_TOP) {
                L.error = "leaf triangle range out of bounds";
                return 0;
            }
            for (uint32_t j = 0; derive_tris_ && j < count; j++) L.tri48[3 * (size_t)(first + j)].w = u2f(count - j);
            return REF_LEAF | first;
        }
        if (n.r <= 0) {
            L.error = "inner BVH node with a non-positive right child";
            return 0;
        }
        uint32_t my = (uint32_t)(L.node64.size() / 4);
        if (my >= REF_TOP) {
            L.error = "too many BVH nodes";
            return 0;
        }
        L.node64.resize(L.node64.size() + 4);
        int dl = 0, dr = 0;
        uint32_t lref = convert((uint32_t)n.l, top, level + 1, &dl);
        uint32_t rref = convert((uint32_t)n.r, top, level + 1, &dr);
        if (!L.error.empty()) return 0;
        const RefNode &cl = nodes_[n.l], &cr = nodes_[n.r];
        Q *q = &L.node64[4 * (size_t)my];
        q[0] = Q{cl.min[0], cl.min[1], cl.min[2], u2f(lref)};
        q[1] = Q{cl.max[0], cl.max[1], cl.max[2], u2f(rref)};
        q[2] = Q{cr.min[0], cr.min[1], cr.min[2], 0.f};
        q[3] = Q{cr.max[0], cr.max[1], cr.max[2], 0.f};
        *depth = 1 + (dl > dr ? dl : dr);
        return my;
    }
};

}  // namespace pc_layout
