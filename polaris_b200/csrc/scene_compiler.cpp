// scene_compiler.cpp -- host-side producer of the flat scene buffers the tracer uploads (libpolaris_scene.so).
//
// polaris compiles a parsed scene into GPU-friendly flat arrays with
// asset/compiler/compiler.go (partitionGeometry, :81-231) and the SAH builder
// asset/compiler/bvh/bvh_builder.go (:124-308).  The CUDA tracer consumes exactly those
// arrays, so to produce "the same compiled scene" without a Go toolchain this file
// re-implements the geometry half of the compiler in C++ (materials, textures and the
// camera are handled by polaris_b200/scene.py, they are tiny).  The compiler proper is scene_compile.hpp (shared with
// the device builder of libpolaris_cuda.so, pc_bvh_build.cu); this file holds the HOST BVH builder.
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
// Plain C ABI, loaded with ctypes.  CPU only.
#include "scene_compile.hpp"

using namespace scenec;

namespace {

class Builder {
  public:
    Builder(const Volumes &v, int min_leaf_items) : v_(v), min_leaf_(min_leaf_items) {}

    // work: item indices, partitioned in place into leaf order
    void build(std::vector<uint32_t> &work) {
        scratch_.resize(work.size());
#pragma omp parallel
#pragma omp single
        root_ = partition(work.data(), scratch_.data(), 0, (uint32_t)work.size(), 0);
        if (!error_.empty()) throw BuildError{error_};
    }
    template <class LeafFn>
    void flatten_into(std::vector<BvhNode> &out, const uint32_t *work, LeafFn &&leaf_fn) {
        flatten(root_.get(), out, work, leaf_fn);
        root_.reset();
    }
    int max_depth() const { return max_depth_; }

  private:
    const Volumes &v_;
    int min_leaf_;
    std::vector<uint32_t> scratch_;
    std::unique_ptr<TreeNode> root_;
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


}  // namespace

extern "C" {

typedef scenec::RawMesh ps_mesh;
typedef scenec::RawInstance ps_instance;

void *ps_build_bvh(const float *bmin, const float *bmax, const float *center, uint32_t n,
                   int min_leaf_items, uint32_t *out_order /* n */) {
    return build_bvh_only(bmin, bmax, center, n, min_leaf_items, out_order,
                          [](const Volumes &v, int min_leaf) { return Builder(v, min_leaf); });
}

void *ps_compile(const ps_mesh *meshes, uint32_t n_meshes, const ps_instance *insts, uint32_t n_insts,
                 const int32_t *mat_root, const int32_t *mat_emissive, uint32_t n_materials,
                 int32_t env_emissive_node) {
    return compile_geometry(meshes, n_meshes, insts, n_insts, mat_root, mat_emissive, n_materials, env_emissive_node,
                            [](const Volumes &v, int min_leaf) { return Builder(v, min_leaf); });
}

const char *ps_error(void *h) {
    auto *c = (Compiled *)h;
    return c->error.empty() ? nullptr : c->error.c_str();
}

int ps_get(void *h, int which, const void **ptr, uint64_t *bytes) { return compiled_get((Compiled *)h, which, ptr, bytes); }

void ps_depths(void *h, int *top_depth, int *mesh_depth) {
    auto *c = (Compiled *)h;
    *top_depth = c->top_depth;
    *mesh_depth = c->mesh_depth;
}

void ps_timing(void *h, double *out8) { memcpy(out8, ((Compiled *)h)->timing, sizeof(double) * scenec::T_COUNT); }

void ps_free(void *h) { delete (Compiled *)h; }

}  // extern "C"
