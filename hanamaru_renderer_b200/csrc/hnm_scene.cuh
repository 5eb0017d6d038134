// hnm_scene.cuh -- hnm_scene: validation of the host description, GPU re-layout of the two BVH
// levels into ONE tree of 64-byte two-child nodes with conservative f32 boxes, triangles gathered
// in the reference's leaf order with precomputed edges, textures as CUDA texture objects.
#ifndef HNM_SCENE_CUH
#define HNM_SCENE_CUH

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "hnm_device.cuh"

namespace hnm {

extern thread_local std::string g_last_error;
inline int set_error(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define HNM_CUDA(call)                                                                                         \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess)                                                                                \
            return hnm::set_error(HNM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
    } while (0)

}  // namespace hnm

struct hnm_scene {
    int device = 0;
    int sm_count = 148;
    hnm::DScene d;
    hnm_config config;
    std::vector<void*> allocs;
    std::vector<cudaArray_t> arrays;
    std::vector<cudaTextureObject_t> texs;
    uint32_t num_nodes = 0, num_tris = 0, num_elements = 0, num_emissions = 0, num_images = 0;
    uint32_t tree_depth = 0;
    std::vector<int32_t> elem_surface;  // host copy, per element
};

namespace hnm {

struct BoxD {
    double lo[3], hi[3];
    bool empty = true;
    void grow(const BoxD& o) {
        if (o.empty) return;
        if (empty) { *this = o; return; }
        for (int k = 0; k < 3; k++) { lo[k] = o.lo[k] < lo[k] ? o.lo[k] : lo[k]; hi[k] = o.hi[k] > hi[k] ? o.hi[k] : hi[k]; }
    }
};
inline BoxD box_of(const double* mn, const double* mx) {
    BoxD b;
    b.empty = !(mn[0] <= mx[0] && mn[1] <= mx[1] && mn[2] <= mx[2]);  // the builder's INF/-INF box of an empty node
    for (int k = 0; k < 3; k++) { b.lo[k] = mn[k]; b.hi[k] = mx[k]; }
    return b;
}
inline float f32_down(double v) {
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -INFINITY);
    return f;
}
inline float f32_up(double v) {
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, INFINITY);
    return f;
}

class SceneBuilder {
  public:
    SceneBuilder(const hnm_scene_desc* d) : d_(d) {}
    std::vector<DNode> nodes;
    std::vector<DTri> tris;
    std::vector<uint32_t> tri_elem, tri_face;
    std::vector<uint32_t> elem_seq;
    // the reference's box chain (hnm_device.cuh: chain_pass)
    std::vector<RefNode> ref_nodes;
    std::vector<uint32_t> tri_leaf, elem_leaf;
    std::vector<double> tri_box, elem_box;
    bool chain_full = false;
    BoxD bounds;
    double pad = 0.0;
    std::string error;

    bool validate() {
        const hnm_scene_desc* d = d_;
        auto fail = [&](const char* m) { error = m; return false; };
        if (!d) return fail("null scene description");
        if (d->abi_version != HNM_ABI_VERSION) return fail("abi_version mismatch");
        if (d->num_elements == 0 || !d->elements) return fail("scene has no elements");
        if (d->num_top_nodes == 0 || !d->top_nodes) return fail("scene has no top-level BVH");
        if (d->config.supersampling == 0 || d->config.supersampling > 8) return fail("supersampling out of range");
        // every path consumes one pair of the random stream per lens iteration and per bounce (src/camera.rs:68,
        // src/renderer.rs:175); HNM_RNG_TAIL words per path are stored, so 2 * (bounce_limit - 1) must fit
        if (d->config.bounce_limit < 2 || 2 * (d->config.bounce_limit - 1) > HNM_RNG_TAIL)
            return fail("bounce_limit out of range (2 .. 17: 2 * (bounce_limit - 1) words of the per-path random stream must fit HNM_RNG_TAIL)");
        for (uint32_t i = 0; i < d->num_elements; i++) {
            const hnm_element& e = d->elements[i];
            if (e.kind < 0 || e.kind > 2) return fail("bad element kind");
            if (e.material < 0 || (uint32_t)e.material >= d->num_materials) return fail("bad material index");
            if (e.kind == HNM_ELEM_MESH && (e.mesh < 0 || (uint32_t)e.mesh >= d->num_meshes)) return fail("bad mesh index");
        }
        for (uint32_t i = 0; i < d->num_materials; i++) {
            const hnm_material& m = d->materials[i];
            if (m.surface < 0 || m.surface > 4) return fail("bad surface type");
            const hnm_texture* ts[3] = {&m.albedo, &m.emission, &m.roughness};
            for (auto t : ts)
                if (t->image >= (int32_t)d->num_images) return fail("bad image index");
        }
        for (uint32_t i = 0; i < d->num_images; i++)
            if (!d->images[i].rgba || d->images[i].width == 0 || d->images[i].height == 0) return fail("bad image");
        for (int k = 0; k < 6; k++)
            if (d->skybox_images[k] < 0 || (uint32_t)d->skybox_images[k] >= d->num_images) return fail("bad skybox image index");
        for (uint32_t i = 0; i < d->num_meshes; i++) {
            const hnm_mesh& m = d->meshes[i];
            if ((uint64_t)m.vertex_offset + m.vertex_count > d->num_vertices) return fail("mesh vertices out of range");
            if ((uint64_t)m.face_offset + m.face_count > d->num_faces) return fail("mesh faces out of range");
            if ((uint64_t)m.node_offset + m.node_count > d->num_mesh_nodes || m.node_count == 0) return fail("mesh nodes out of range");
            if ((uint64_t)m.index_offset + m.index_count > d->num_mesh_indices) return fail("mesh indices out of range");
            for (uint32_t f = 0; f < m.face_count * 3; f++)
                if (d->faces[(size_t)m.face_offset * 3 + f] >= m.vertex_count) return fail("face vertex index out of range");
            for (uint32_t n = 0; n < m.node_count; n++) {
                const hnm_bvh_node& nd = d->mesh_nodes[m.node_offset + n];
                if (nd.child0 < 0) {
                    if ((uint64_t)nd.first + nd.count > m.index_count) return fail("mesh leaf range out of range");
                } else if ((uint32_t)nd.child0 >= m.node_count || nd.child1 < 0 || (uint32_t)nd.child1 >= m.node_count ||
                           (uint32_t)nd.child0 <= n || (uint32_t)nd.child1 <= n) {
                    return fail("mesh node links must point forward (DFS pre-order)");
                }
            }
            for (uint32_t k = 0; k < m.index_count; k++)
                if (d->mesh_indices[m.index_offset + k] >= m.face_count) return fail("mesh index entry out of range");
        }
        for (uint32_t n = 0; n < d->num_top_nodes; n++) {
            const hnm_bvh_node& nd = d->top_nodes[n];
            if (nd.child0 < 0) {
                if ((uint64_t)nd.first + nd.count > d->num_top_indices) return fail("top leaf range out of range");
            } else if ((uint32_t)nd.child0 >= d->num_top_nodes || nd.child1 < 0 || (uint32_t)nd.child1 >= d->num_top_nodes ||
                       (uint32_t)nd.child0 <= n || (uint32_t)nd.child1 <= n) {
                return fail("top node links must point forward (DFS pre-order)");
            }
        }
        for (uint32_t k = 0; k < d->num_top_indices; k++)
            if (d->top_indices[k] >= d->num_elements) return fail("top index entry out of range");
        for (uint32_t k = 0; k < d->num_emissions; k++)
            if (d->emissions[k] >= d->num_elements || d->elements[d->emissions[k]].kind != HNM_ELEM_SPHERE)
                return fail("emissions must be spheres (only Sphere::sample_on_surface exists, src/scene.rs:92)");
        return true;
    }

    bool build() {
        const hnm_scene_desc* d = d_;
        // 1. the reference's DFS order of elements, and triangles gathered in (element, leaf) order
        elem_seq.assign(d->num_elements, 0xFFFFFFFFu);
        std::vector<uint32_t> order;
        collect_top(0, order);
        for (uint32_t s = 0; s < order.size(); s++) {
            if (elem_seq[order[s]] != 0xFFFFFFFFu) { error = "element listed twice in the top-level BVH"; return false; }
            elem_seq[order[s]] = s;
        }
        // the host's trees as parent-linked nodes: top level first, then each mesh below the top-level leaf of its element
        ref_nodes.clear();
        ref_nodes.resize(d->num_top_nodes);
        elem_leaf.assign(d->num_elements, REF_NONE);
        auto set_ref = [&](uint32_t at, const hnm_bvh_node& n, uint32_t parent) {
            for (int k = 0; k < 3; k++) { ref_nodes[at].box[k] = n.aabb_min[k]; ref_nodes[at].box[3 + k] = n.aabb_max[k]; }
            ref_nodes[at].parent = parent;
            ref_nodes[at]._pad = 0;
        };
        for (uint32_t n = 0; n < d->num_top_nodes; n++) set_ref(n, d->top_nodes[n], REF_NONE);
        for (uint32_t n = 0; n < d->num_top_nodes; n++) {
            const hnm_bvh_node& nd = d->top_nodes[n];
            if (nd.child0 >= 0) { ref_nodes[nd.child0].parent = n; ref_nodes[nd.child1].parent = n; }
            else for (uint32_t k = 0; k < nd.count; k++) elem_leaf[d->top_indices[nd.first + k]] = n;
        }
        elem_box.assign((size_t)d->num_elements * 6, 0.0);
        for (uint32_t el = 0; el < d->num_elements; el++) {
            if (elem_leaf[el] == REF_NONE) continue;  // not listed: unreachable for the reference too (never a candidate)
            for (int k = 0; k < 6; k++) elem_box[(size_t)el * 6 + k] = ref_nodes[elem_leaf[el]].box[k];
        }
        mesh_tri_base_.assign(d->num_meshes, 0);
        std::vector<char> mesh_used(d->num_meshes, 0);
        for (uint32_t el : order) {
            const hnm_element& e = d->elements[el];
            BoxD eb = element_box(e);
            bounds.grow(eb);
            if (e.kind != HNM_ELEM_MESH) continue;
            if (mesh_used[e.mesh]) { error = "mesh shared by two elements"; return false; }
            mesh_used[e.mesh] = 1;
            const hnm_mesh& m = d->meshes[e.mesh];
            mesh_tri_base_[e.mesh] = (uint32_t)tris.size();
            {
                const uint32_t nbase = (uint32_t)ref_nodes.size(), tbase = (uint32_t)tris.size();
                ref_nodes.resize(nbase + m.node_count);
                tri_leaf.resize(tbase + m.index_count, REF_NONE);
                tri_box.resize((size_t)(tbase + m.index_count) * 6, 0.0);
                for (uint32_t n = 0; n < m.node_count; n++) set_ref(nbase + n, d->mesh_nodes[m.node_offset + n], n == 0 ? elem_leaf[el] : REF_NONE);
                for (uint32_t n = 0; n < m.node_count; n++) {
                    const hnm_bvh_node& nd = d->mesh_nodes[m.node_offset + n];
                    if (nd.child0 >= 0) { ref_nodes[nbase + nd.child0].parent = nbase + n; ref_nodes[nbase + nd.child1].parent = nbase + n; }
                    else for (uint32_t k = 0; k < nd.count; k++) {
                        tri_leaf[tbase + nd.first + k] = nbase + n;
                        for (int c = 0; c < 6; c++) tri_box[(size_t)(tbase + nd.first + k) * 6 + c] = ref_nodes[nbase + n].box[c];
                    }
                }
                for (uint32_t k = 0; k < m.index_count; k++)
                    if (tri_leaf[tbase + k] == REF_NONE) { error = "mesh index entry not owned by any leaf"; return false; }
            }
            const double* vb = d->vertices + 3 * (size_t)m.vertex_offset;
            for (uint32_t k = 0; k < m.index_count; k++) {
                uint32_t face = d->mesh_indices[m.index_offset + k];
                const uint32_t* f = d->faces + 3 * (size_t)(m.face_offset + face);
                DTri t;
                t._pad = 0.0;
                t.v0x = vb[3 * f[0]]; t.v0y = vb[3 * f[0] + 1]; t.v0z = vb[3 * f[0] + 2];
                // edge1 = v1 - v0, edge2 = v2 - v0 (src/bvh.rs:268-269)
                t.e1x = vb[3 * f[1]] - t.v0x; t.e1y = vb[3 * f[1] + 1] - t.v0y; t.e1z = vb[3 * f[1] + 2] - t.v0z;
                t.e2x = vb[3 * f[2]] - t.v0x; t.e2y = vb[3 * f[2] + 1] - t.v0y; t.e2z = vb[3 * f[2] + 2] - t.v0z;
                tris.push_back(t);
                tri_elem.push_back(el);
                tri_face.push_back(face);
            }
        }
        if (tris.size() >= (1u << 26)) { error = "too many triangles for the 26-bit leaf encoding"; return false; }
        if (d->num_elements >= (1u << 26)) { error = "too many elements"; return false; }
        double R = 0.0;
        if (!bounds.empty)
            for (int k = 0; k < 3; k++) R = std::fmax(R, std::fmax(std::fabs(bounds.lo[k]), std::fabs(bounds.hi[k])));
        pad = std::fmax(R, 1e-30) * 0x1p-19;  // absolute slack for the f32 rounding of (box - origin)
        // 2. one tree: node 0 is reserved for the root.  Default: a fresh SAH tree over ALL primitives (results do not
        //    depend on the tree -- DESIGN.md section 2 -- only the number of visits does; the host's median-split
        //    trees cost ~1.5x more node visits).  HNM_BVH=ref keeps the host's topology (A/B and debugging).
        nodes.clear();
        nodes.push_back(DNode{});
        BoxD rb;
        const char* mode = getenv("HNM_BVH");
        const bool use_ref = mode && std::string(mode) == "ref";
        tri_order.resize(tris.size());
        for (size_t i = 0; i < tris.size(); i++) tri_order[i] = (uint32_t)i;
        int32_t root = use_ref ? build_top(0, rb) : build_sah(rb);
        if (root >= 0) {
            nodes[0] = nodes[root];
        } else {
            DNode n;
            set_child(n, 0, root, rb);
            BoxD none;
            set_child(n, 1, leaf_link(LEAF_NONE, 0, 0), none);
            nodes[0] = n;
        }
        // the traversal kernels keep the far child of every level on a stack of HNM_STACK entries: a deeper tree
        // (a degenerate caller-supplied topology under HNM_BVH=ref, or pathological geometry) is refused, not truncated
        depth = 0;
        {
            std::vector<std::pair<int32_t, uint32_t>> st;
            st.push_back({0, 1u});
            while (!st.empty()) {
                auto [idx, dep] = st.back();
                st.pop_back();
                depth = std::max(depth, dep);
                const DNode& n = nodes[idx];
                if (n.c0 >= 0) st.push_back({n.c0, dep + 1});
                if (n.c1 >= 0) st.push_back({n.c1, dep + 1});
            }
        }
        if (depth > HNM_STACK) { error = "BVH deeper than the traversal stack (" + std::to_string(depth) + " > " + std::to_string(HNM_STACK) + " levels)"; return false; }
        // chain_pass relies on nested boxes (child inside parent, as the reference's builder produces them by merging);
        // a description whose boxes are not nested still renders exactly, on the slow path that walks every chain
        chain_full = false;
        for (const RefNode& n : ref_nodes) {
            if (n.parent == REF_NONE) continue;
            const RefNode& p = ref_nodes[n.parent];
            const bool empty = !(n.box[0] <= n.box[3] && n.box[1] <= n.box[4] && n.box[2] <= n.box[5]);
            if (empty) continue;
            for (int k = 0; k < 3; k++)
                if (!(n.box[k] >= p.box[k] && n.box[3 + k] <= p.box[3 + k])) chain_full = true;
        }
        if (getenv("HNM_CHAIN_FULL")) chain_full = atoi(getenv("HNM_CHAIN_FULL")) != 0;  // A/B and tests
        return true;
    }
    uint32_t depth = 0;

    std::vector<uint32_t> tri_order;  // traversal (leaf) position -> triangle index in the reference's leaf order
    void detach() { d_ = nullptr; }   // a cached build outlives the caller's description: only the result vectors remain valid

  private:
    const hnm_scene_desc* d_;
    std::vector<uint32_t> mesh_tri_base_;

    // ---- binned SAH build over every primitive of the scene (triangles, spheres, cuboids) ------------------
    struct Prim {
        BoxD box;
        double c[3];
        int kind;      // LEAF_TRI / LEAF_SPHERE / LEAF_CUBOID
        uint32_t id;   // triangle index (reference leaf order) or element id
    };
    static double half_area(const BoxD& b) {
        if (b.empty) return 0.0;
        double dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
    int32_t sah_rec(std::vector<Prim>& prims, size_t lo, size_t hi, std::vector<uint32_t>& order, BoxD& out) {
        out = BoxD();
        BoxD cb;  // centroid bounds
        bool all_tri = true;
        {
            double ol[3] = {INFINITY, INFINITY, INFINITY}, oh[3] = {-INFINITY, -INFINITY, -INFINITY};
            double cl[3] = {INFINITY, INFINITY, INFINITY}, ch[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (size_t i = lo; i < hi; i++) {
                const Prim& p = prims[i];
                for (int k = 0; k < 3; k++) {
                    ol[k] = p.box.lo[k] < ol[k] ? p.box.lo[k] : ol[k]; oh[k] = p.box.hi[k] > oh[k] ? p.box.hi[k] : oh[k];
                    cl[k] = p.c[k] < cl[k] ? p.c[k] : cl[k]; ch[k] = p.c[k] > ch[k] ? p.c[k] : ch[k];
                }
                all_tri = all_tri && p.kind == LEAF_TRI;
            }
            out.empty = cb.empty = hi <= lo;
            for (int k = 0; k < 3; k++) { out.lo[k] = ol[k]; out.hi[k] = oh[k]; cb.lo[k] = cl[k]; cb.hi[k] = ch[k]; }
        }
        const size_t n = hi - lo;
        auto make_leaf = [&]() -> int32_t {
            if (prims[lo].kind != LEAF_TRI) return leaf_link(prims[lo].kind, 1, prims[lo].id);
            uint32_t first = (uint32_t)order.size();
            for (size_t i = lo; i < hi; i++) order.push_back(prims[i].id);
            return leaf_link(LEAF_TRI, (uint32_t)n, first);
        };
        if (n == 1) return make_leaf();
        // best binned split over the three axes: ONE pass over the primitives fills the bins of all three axes (plain
        // min / max on +-inf-initialised boxes; the bin index is a multiply, and std::partition below uses the same
        // expression).  This loop is the host-side cost of hnm_scene_create: 44 -> ~15 ms for the 12 k-triangle default scene.
        const int NBMAX = 32;
        const int NB = n >= 64 ? 32 : (n >= 16 ? 16 : 8);  // the fixed cost per node (bin set-up + sweeps) dominates small nodes
        double best_cost = 1e300;
        int best_axis = -1, best_bin = -1;
        const double parent_area = std::fmax(half_area(out), 1e-300);
        double scale[3];
        bool use_axis[3];
        for (int axis = 0; axis < 3; axis++) {
            const double ext = cb.hi[axis] - cb.lo[axis];
            use_axis[axis] = ext > 0.0;
            scale[axis] = use_axis[axis] ? (double)NB / ext : 0.0;
        }
        struct Bin { double lo[3], hi[3]; size_t cnt; };
        Bin bins[3][NBMAX];
        for (int axis = 0; axis < 3; axis++)
            for (int k = 0; k < NB; k++) {
                Bin& bn = bins[axis][k];
                bn.lo[0] = bn.lo[1] = bn.lo[2] = INFINITY; bn.hi[0] = bn.hi[1] = bn.hi[2] = -INFINITY; bn.cnt = 0;
            }
        for (size_t i = lo; i < hi; i++) {
            const Prim& p = prims[i];
            for (int axis = 0; axis < 3; axis++) {
                if (!use_axis[axis]) continue;
                int k = (int)((p.c[axis] - cb.lo[axis]) * scale[axis]);
                k = k < 0 ? 0 : (k >= NB ? NB - 1 : k);
                Bin& bn = bins[axis][k];
                for (int c = 0; c < 3; c++) {
                    bn.lo[c] = p.box.lo[c] < bn.lo[c] ? p.box.lo[c] : bn.lo[c];
                    bn.hi[c] = p.box.hi[c] > bn.hi[c] ? p.box.hi[c] : bn.hi[c];
                }
                bn.cnt++;
            }
        }
        auto area_of = [](const double* l, const double* h) {
            const double dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
            return dx * dy + dy * dz + dz * dx;
        };
        for (int axis = 0; axis < 3; axis++) {
            if (!use_axis[axis]) continue;
            double right_area[NBMAX];
            size_t right_cnt[NBMAX];
            double l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
            size_t c = 0;
            for (int k = NB - 1; k > 0; k--) {
                const Bin& bn = bins[axis][k];
                for (int t = 0; t < 3; t++) { l[t] = bn.lo[t] < l[t] ? bn.lo[t] : l[t]; h[t] = bn.hi[t] > h[t] ? bn.hi[t] : h[t]; }
                c += bn.cnt;
                right_area[k] = c ? area_of(l, h) : 0.0;
                right_cnt[k] = c;
            }
            l[0] = l[1] = l[2] = INFINITY; h[0] = h[1] = h[2] = -INFINITY;
            c = 0;
            for (int k = 0; k < NB - 1; k++) {
                const Bin& bn = bins[axis][k];
                for (int t = 0; t < 3; t++) { l[t] = bn.lo[t] < l[t] ? bn.lo[t] : l[t]; h[t] = bn.hi[t] > h[t] ? bn.hi[t] : h[t]; }
                c += bn.cnt;
                if (c == 0 || right_cnt[k + 1] == 0) continue;
                const double cost = 1.0 + sah_ct * (area_of(l, h) * (double)c + right_area[k + 1] * (double)right_cnt[k + 1]) / parent_area;
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = k; }
            }
        }
        const double leaf_cost = sah_ct * (double)n;  // cost of one triangle pre-test in node steps
        if (all_tri && n <= sah_max_leaf && (best_axis < 0 || leaf_cost <= best_cost)) return make_leaf();
        size_t mid;
        if (best_axis >= 0) {
            const double base = cb.lo[best_axis], sc_ = scale[best_axis];
            auto it = std::partition(prims.begin() + lo, prims.begin() + hi, [&](const Prim& p) {
                int b = (int)((p.c[best_axis] - base) * sc_);
                b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                return b <= best_bin;
            });
            mid = (size_t)(it - prims.begin());
        } else {
            mid = lo + n / 2;  // identical centroids: split by count
        }
        if (mid == lo || mid == hi) mid = lo + n / 2;
        int32_t slot = (int32_t)nodes.size();
        nodes.push_back(DNode{});
        BoxD b0, b1;
        int32_t l0 = sah_rec(prims, lo, mid, order, b0);
        int32_t l1 = sah_rec(prims, mid, hi, order, b1);
        return make_inner(l0, b0, l1, b1, slot);
    }
    double sah_ct = 1.0;
    size_t sah_max_leaf = 7;
    int32_t build_sah(BoxD& out) {
        if (const char* e = getenv("HNM_SAH_CT")) { double v = atof(e); if (v > 0.0) sah_ct = v; }
        if (const char* e = getenv("HNM_SAH_MAXLEAF")) { int v = atoi(e); if (v >= 1 && v <= 7) sah_max_leaf = (size_t)v; }
        std::vector<Prim> prims;
        prims.reserve(tris.size() + d_->num_elements);
        for (size_t g = 0; g < tris.size(); g++) {
            const DTri& t = tris[g];
            // the vertices as the host stored them: v1 = v0 + e1 is not bit-exact, so box the three points generously
            double vx[3] = {t.v0x, t.v0x + t.e1x, t.v0x + t.e2x}, vy[3] = {t.v0y, t.v0y + t.e1y, t.v0y + t.e2y}, vz[3] = {t.v0z, t.v0z + t.e1z, t.v0z + t.e2z};
            Prim p;
            p.box.empty = false;
            p.box.lo[0] = std::fmin(vx[0], std::fmin(vx[1], vx[2])); p.box.hi[0] = std::fmax(vx[0], std::fmax(vx[1], vx[2]));
            p.box.lo[1] = std::fmin(vy[0], std::fmin(vy[1], vy[2])); p.box.hi[1] = std::fmax(vy[0], std::fmax(vy[1], vy[2]));
            p.box.lo[2] = std::fmin(vz[0], std::fmin(vz[1], vz[2])); p.box.hi[2] = std::fmax(vz[0], std::fmax(vz[1], vz[2]));
            for (int k = 0; k < 3; k++) {
                // v0 + e can differ from the stored vertex by one rounding: widen by 2 ulp of the magnitude
                double m = std::fmax(std::fabs(p.box.lo[k]), std::fabs(p.box.hi[k])) * 0x1p-51;
                p.box.lo[k] -= m; p.box.hi[k] += m;
                p.c[k] = 0.5 * (p.box.lo[k] + p.box.hi[k]);
            }
            p.kind = LEAF_TRI;
            p.id = (uint32_t)g;
            prims.push_back(p);
        }
        for (uint32_t el = 0; el < d_->num_elements; el++) {
            const hnm_element& e = d_->elements[el];
            if (e.kind == HNM_ELEM_MESH) continue;
            if (elem_seq[el] == 0xFFFFFFFFu) continue;  // not listed in the top-level tree: the reference never tests it
            Prim p;
            p.box = element_box(e);
            if (p.box.empty) continue;
            for (int k = 0; k < 3; k++) p.c[k] = 0.5 * (p.box.lo[k] + p.box.hi[k]);
            p.kind = e.kind == HNM_ELEM_SPHERE ? LEAF_SPHERE : LEAF_CUBOID;
            p.id = el;
            prims.push_back(p);
        }
        tri_order.clear();
        if (prims.empty()) { out = BoxD(); return leaf_link(LEAF_NONE, 0, 0); }
        return sah_rec(prims, 0, prims.size(), tri_order, out);
    }

    static int32_t leaf_link(int kind, uint32_t count, uint32_t first) { return ~(int32_t)(((uint32_t)kind << 29) | (count << 26) | first); }

    void collect_top(uint32_t node, std::vector<uint32_t>& order) {
        const hnm_bvh_node& n = d_->top_nodes[node];
        if (n.child0 < 0) {
            for (uint32_t k = 0; k < n.count; k++) order.push_back(d_->top_indices[n.first + k]);
        } else {
            collect_top((uint32_t)n.child0, order);
            collect_top((uint32_t)n.child1, order);
        }
    }
    BoxD element_box(const hnm_element& e) {
        if (e.kind == HNM_ELEM_SPHERE) {  // src/scene.rs:82-87
            double mn[3] = {e.a.x - e.radius, e.a.y - e.radius, e.a.z - e.radius};
            double mx[3] = {e.a.x + e.radius, e.a.y + e.radius, e.a.z + e.radius};
            return box_of(mn, mx);
        }
        if (e.kind == HNM_ELEM_CUBOID) {
            double mn[3] = {e.a.x, e.a.y, e.a.z}, mx[3] = {e.b.x, e.b.y, e.b.z};
            return box_of(mn, mx);
        }
        const hnm_bvh_node& r = d_->mesh_nodes[d_->meshes[e.mesh].node_offset];
        return box_of(r.aabb_min, r.aabb_max);
    }
    void set_child(DNode& n, int which, int32_t link, const BoxD& b) {
        float lo[3], hi[3];
        for (int k = 0; k < 3; k++) {
            if (b.empty) { lo[k] = NAN; hi[k] = NAN; }  // NaN box: every comparison fails, never entered
            else { lo[k] = f32_down(b.lo[k] - pad); hi[k] = f32_up(b.hi[k] + pad); }
        }
        if (which == 0) {
            n.lo0x = lo[0]; n.lo0y = lo[1]; n.lo0z = lo[2]; n.hi0x = hi[0]; n.hi0y = hi[1]; n.hi0z = hi[2];
            n.c0 = link;
        } else {
            n.lo1x = lo[0]; n.lo1y = lo[1]; n.lo1z = lo[2]; n.hi1x = hi[0]; n.hi1y = hi[1]; n.hi1z = hi[2];
            n.c1 = link;
        }
    }
    int32_t make_inner(int32_t l0, const BoxD& b0, int32_t l1, const BoxD& b1, int32_t slot = -1) {
        DNode n;
        memset(&n, 0, sizeof(n));
        set_child(n, 0, l0, b0);
        set_child(n, 1, l1, b1);
        if (slot < 0) { nodes.push_back(n); return (int32_t)nodes.size() - 1; }
        nodes[slot] = n;
        return slot;
    }
    // a run of leaf-order triangles [first, first+count) -> leaf link (count <= 7) or a small subtree
    int32_t tri_run(uint32_t first, uint32_t count, const BoxD& box) {
        if (count == 0) return leaf_link(LEAF_NONE, 0, 0);
        if (count <= 7) return leaf_link(LEAF_TRI, count, first);
        uint32_t h = count / 2;
        int32_t a = tri_run(first, h, box), b = tri_run(first + h, count - h, box);
        return make_inner(a, box, b, box);
    }
    int32_t build_mesh(const hnm_mesh& m, uint32_t base, uint32_t rel, BoxD& out) {
        const hnm_bvh_node& n = d_->mesh_nodes[m.node_offset + rel];
        out = box_of(n.aabb_min, n.aabb_max);
        if (n.child0 < 0) return tri_run(base + n.first, n.count, out);
        int32_t slot = (int32_t)nodes.size();
        nodes.push_back(DNode{});
        BoxD b0, b1;
        int32_t l0 = build_mesh(m, base, (uint32_t)n.child0, b0);
        int32_t l1 = build_mesh(m, base, (uint32_t)n.child1, b1);
        return make_inner(l0, b0, l1, b1, slot);
    }
    int32_t element_ref(uint32_t el, BoxD& out) {
        const hnm_element& e = d_->elements[el];
        if (e.kind == HNM_ELEM_MESH) return build_mesh(d_->meshes[e.mesh], mesh_tri_base_[e.mesh], 0, out);
        out = element_box(e);
        return leaf_link(e.kind == HNM_ELEM_SPHERE ? LEAF_SPHERE : LEAF_CUBOID, 1, el);
    }
    int32_t combine(const uint32_t* els, uint32_t count, BoxD& out) {
        if (count == 0) { out = BoxD(); return leaf_link(LEAF_NONE, 0, 0); }
        if (count == 1) return element_ref(els[0], out);
        uint32_t h = count / 2;
        BoxD b0, b1;
        int32_t l0 = combine(els, h, b0), l1 = combine(els + h, count - h, b1);
        out = BoxD();
        out.grow(b0); out.grow(b1);
        return make_inner(l0, b0, l1, b1);
    }
    int32_t build_top(uint32_t node, BoxD& out) {
        const hnm_bvh_node& n = d_->top_nodes[node];
        if (n.child0 < 0) return combine(d_->top_indices + n.first, n.count, out);
        BoxD b0, b1;
        int32_t l0 = build_top((uint32_t)n.child0, b0), l1 = build_top((uint32_t)n.child1, b1);
        out = BoxD();
        out.grow(b0); out.grow(b1);
        return make_inner(l0, b0, l1, b1);
    }
};

template <typename T>
inline int upload(hnm_scene* s, const std::vector<T>& v, const T** out) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    HNM_CUDA(cudaMalloc(&p, bytes));
    s->allocs.push_back(p);
    if (!v.empty()) HNM_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)p;
    return 0;
}

inline void scene_free(hnm_scene* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    for (auto t : s->texs) cudaDestroyTextureObject(t);
    for (auto a : s->arrays) cudaFreeArray(a);
    for (auto p : s->allocs) cudaFree(p);
    delete s;
}

// host part of hnm_scene_create: validation + re-layout (no CUDA call); one build serves every device of a group
inline int scene_build_host(const hnm_scene_desc* desc, SceneBuilder& b) {
    const bool timing = getenv("HNM_BUILD_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    if (!b.validate()) return set_error(HNM_ERR_INVALID, "scene description: " + b.error);
    const double t1 = now();
    if (!b.build()) return set_error(HNM_ERR_INVALID, "scene description: " + b.error);
    if (timing) fprintf(stderr, "hnm_scene_create host part: validate %.2f ms, build %.2f ms (%zu nodes, depth %u)\n", t1 - t0, now() - t1, b.nodes.size(), b.depth);
    (void)desc;
    return 0;
}
inline int scene_upload(const hnm_scene_desc* desc, const SceneBuilder& b, int device, hnm_scene** out);

// The re-layout depends on the geometry arrays and the two trees only.  A host that calls Renderer::render again with the
// same scene (progress renders, the debug pass after the main pass, one scene copy per device) gets the previous build
// back: a 64-bit content hash of those arrays (and of the environment knobs the builder reads) keys a small cache.
inline uint64_t hash_bytes(uint64_t h, const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t v;
        memcpy(&v, b + i, 8);
        h = (h ^ v) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
    }
    for (; i < n; i++) h = (h ^ b[i]) * 0x100000001B3ull;
    return h;
}
inline uint64_t scene_geometry_hash(const hnm_scene_desc* d) {
    uint64_t h = 0xCBF29CE484222325ull;
    auto arr = [&](const void* p, size_t n) { uint64_t sz = n; h = hash_bytes(h, &sz, 8); if (p && n) h = hash_bytes(h, p, n); };
    arr(d->elements, (size_t)d->num_elements * sizeof(hnm_element));
    arr(d->meshes, (size_t)d->num_meshes * sizeof(hnm_mesh));
    arr(d->vertices, (size_t)d->num_vertices * 24);
    arr(d->faces, (size_t)d->num_faces * 12);
    arr(d->mesh_nodes, (size_t)d->num_mesh_nodes * sizeof(hnm_bvh_node));
    arr(d->mesh_indices, (size_t)d->num_mesh_indices * 4);
    arr(d->top_nodes, (size_t)d->num_top_nodes * sizeof(hnm_bvh_node));
    arr(d->top_indices, (size_t)d->num_top_indices * 4);
    for (const char* k : {"HNM_BVH", "HNM_SAH_CT", "HNM_SAH_MAXLEAF", "HNM_CHAIN_FULL"}) {
        const char* v = getenv(k);
        arr(v ? v : "", v ? strlen(v) : 0);
    }
    return h;
}
struct BuildCache {
    std::mutex m;
    std::vector<std::pair<uint64_t, std::shared_ptr<SceneBuilder>>> entries;  // most recent last, at most 4
};
inline BuildCache& build_cache() { static BuildCache c; return c; }
// validated + built description, from the cache when the same geometry was built before
inline int scene_build_cached(const hnm_scene_desc* desc, std::shared_ptr<SceneBuilder>& out) {
    if (!desc) return set_error(HNM_ERR_INVALID, "scene description: null scene description");
    {
        // validation always runs (it covers the parts of the description that are not hashed: materials, images, config)
        SceneBuilder v(desc);
        if (!v.validate()) return set_error(HNM_ERR_INVALID, "scene description: " + v.error);
    }
    const bool use_cache = !getenv("HNM_BUILD_CACHE") || atoi(getenv("HNM_BUILD_CACHE")) != 0;
    const uint64_t key = scene_geometry_hash(desc);
    BuildCache& c = build_cache();
    if (use_cache) {
        std::lock_guard<std::mutex> g(c.m);
        for (auto& e : c.entries)
            if (e.first == key) { out = e.second; return 0; }
    }
    auto b = std::make_shared<SceneBuilder>(desc);
    int rc = scene_build_host(desc, *b);
    if (rc) return rc;
    b->detach();
    if (use_cache) {
        std::lock_guard<std::mutex> g(c.m);
        c.entries.emplace_back(key, b);
        if (c.entries.size() > 4) c.entries.erase(c.entries.begin());
    }
    out = b;
    return 0;
}
inline int scene_create(const hnm_scene_desc* desc, int device, hnm_scene** out) {
    if (!out) return set_error(HNM_ERR_INVALID, "null output pointer");
    *out = nullptr;
    std::shared_ptr<SceneBuilder> b;
    int rc = scene_build_cached(desc, b);
    if (rc) return rc;
    return scene_upload(desc, *b, device, out);
}
inline int scene_upload(const hnm_scene_desc* desc, const SceneBuilder& b, int device, hnm_scene** out) {
    *out = nullptr;
    HNM_CUDA(cudaSetDevice(device));
    int sm_count = 0;
    HNM_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    hnm_scene* s = new hnm_scene();
    s->device = device;
    s->sm_count = sm_count;
    s->config = desc->config;
    memset(&s->d, 0, sizeof(s->d));
    int rc = 0;
    auto bail = [&](int code) { scene_free(s); return code; };
    if ((rc = upload(s, b.nodes, &s->d.nodes))) return bail(rc);
    if ((rc = upload(s, b.tris, &s->d.tris))) return bail(rc);
    {
        // f32 copies (round to nearest: relative error 2^-24 per component, covered by the pre-test's margins);
        // the normal e1 x e2 is formed in f64 first
        // ... in TRAVERSAL order (tri_order); candidates map back to the reference order through tri_perm
        std::vector<float4> tf(b.tri_order.size() * 3);
        for (size_t i = 0; i < b.tri_order.size(); i++) {
            const DTri& t = b.tris[b.tri_order[i]];
            double nx = t.e1y * t.e2z - t.e1z * t.e2y, ny = t.e1z * t.e2x - t.e1x * t.e2z, nz = t.e1x * t.e2y - t.e1y * t.e2x;
            tf[3 * i] = make_float4((float)t.v0x, (float)t.v0y, (float)t.v0z, (float)t.e1x);
            tf[3 * i + 1] = make_float4((float)t.e1y, (float)t.e1z, (float)t.e2x, (float)t.e2y);
            tf[3 * i + 2] = make_float4((float)t.e2z, (float)nx, (float)ny, (float)nz);
        }
        if ((rc = upload(s, tf, &s->d.trif))) return bail(rc);
        if ((rc = upload(s, b.tri_order, &s->d.tri_perm))) return bail(rc);
    }
    if ((rc = upload(s, b.ref_nodes, &s->d.ref_nodes))) return bail(rc);
    if ((rc = upload(s, b.tri_leaf, &s->d.tri_leaf))) return bail(rc);
    if ((rc = upload(s, b.tri_box, &s->d.tri_box))) return bail(rc);
    if ((rc = upload(s, b.elem_leaf, &s->d.elem_leaf))) return bail(rc);
    if ((rc = upload(s, b.elem_box, &s->d.elem_box))) return bail(rc);
    s->d.chain_full = b.chain_full ? 1u : 0u;
    if ((rc = upload(s, b.tri_elem, &s->d.tri_elem))) return bail(rc);
    if ((rc = upload(s, b.tri_face, &s->d.tri_face))) return bail(rc);
    std::vector<DElement> els(desc->num_elements);
    s->elem_surface.resize(desc->num_elements);
    for (uint32_t i = 0; i < desc->num_elements; i++) {
        const hnm_element& e = desc->elements[i];
        DElement& de = els[i];
        memset(&de, 0, sizeof(de));
        de.ax = e.a.x; de.ay = e.a.y; de.az = e.a.z; de.bx = e.b.x; de.by = e.b.y; de.bz = e.b.z;
        de.radius = e.radius; de.kind = e.kind; de.material = e.material; de.seq = b.elem_seq[i]; de.mesh = e.mesh;
        s->elem_surface[i] = desc->materials[e.material].surface;
    }
    if ((rc = upload(s, els, &s->d.elements))) return bail(rc);
    {
        std::vector<float4> ef((size_t)desc->num_elements * 4, make_float4(NAN, NAN, NAN, NAN));
        for (uint32_t i = 0; i < desc->num_elements; i++) {
            const hnm_element& e = desc->elements[i];
            if (e.kind == HNM_ELEM_SPHERE) {
                ef[4 * i] = make_float4((float)e.a.x, (float)e.a.y, (float)e.a.z, (float)e.radius);
            } else if (e.kind == HNM_ELEM_CUBOID) {
                const double lo[3] = {e.a.x, e.a.y, e.a.z}, hi[3] = {e.b.x, e.b.y, e.b.z};
                float ol[3], oh[3], il[3], ih[3];
                bool inner_ok = true;
                for (int k = 0; k < 3; k++) {
                    ol[k] = f32_down(lo[k] - b.pad); oh[k] = f32_up(hi[k] + b.pad);
                    il[k] = f32_up(lo[k] + b.pad); ih[k] = f32_down(hi[k] - b.pad);
                    inner_ok = inner_ok && il[k] <= ih[k];
                }
                ef[4 * i] = make_float4(ol[0], ol[1], ol[2], 0.f);
                ef[4 * i + 1] = make_float4(oh[0], oh[1], oh[2], 0.f);
                if (inner_ok) {  // else: stays NaN, the certain-hit test never passes
                    ef[4 * i + 2] = make_float4(il[0], il[1], il[2], 0.f);
                    ef[4 * i + 3] = make_float4(ih[0], ih[1], ih[2], 0.f);
                }
            }
        }
        if ((rc = upload(s, ef, &s->d.elemf))) return bail(rc);
    }
    std::vector<DMaterial> mats(desc->num_materials);
    for (uint32_t i = 0; i < desc->num_materials; i++) {
        const hnm_material& m = desc->materials[i];
        auto tex = [](const hnm_texture& t) { DTexture d; d.r = t.color.x; d.g = t.color.y; d.b = t.color.z; d.image = t.image < 0 ? -1 : t.image; d._pad = 0; return d; };
        mats[i].albedo = tex(m.albedo); mats[i].emission = tex(m.emission); mats[i].roughness = tex(m.roughness);
        mats[i].param = m.param; mats[i].surface = m.surface;
        mats[i].has_image = (m.albedo.image >= 0 || m.emission.image >= 0 || m.roughness.image >= 0) ? 1 : 0;
    }
    if ((rc = upload(s, mats, &s->d.materials))) return bail(rc);
    {
        std::vector<double> lut(256);
        for (int i = 0; i < 256; i++) lut[i] = (double)i / 255.0;  // src/color.rs:18-24, one correctly rounded division each
        if ((rc = upload(s, lut, &s->d.unorm8))) return bail(rc);
    }
    std::vector<DImage> imgs(desc->num_images);
    for (uint32_t i = 0; i < desc->num_images; i++) {
        const hnm_image& im = desc->images[i];
        cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
        cudaArray_t arr = nullptr;
        cudaError_t e = cudaMallocArray(&arr, &cd, im.width, im.height);
        if (e != cudaSuccess) { set_error(HNM_ERR_CUDA, std::string("cudaMallocArray: ") + cudaGetErrorString(e)); return bail(HNM_ERR_CUDA); }
        s->arrays.push_back(arr);
        e = cudaMemcpy2DToArray(arr, 0, 0, im.rgba, (size_t)im.width * 4, (size_t)im.width * 4, im.height, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { set_error(HNM_ERR_CUDA, std::string("cudaMemcpy2DToArray: ") + cudaGetErrorString(e)); return bail(HNM_ERR_CUDA); }
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = arr;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;       // exact texels; the reference's f64 bilinear is done in the kernel
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        cudaTextureObject_t tex = 0;
        e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
        if (e != cudaSuccess) { set_error(HNM_ERR_CUDA, std::string("cudaCreateTextureObject: ") + cudaGetErrorString(e)); return bail(HNM_ERR_CUDA); }
        s->texs.push_back(tex);
        imgs[i].tex = tex; imgs[i].width = im.width; imgs[i].height = im.height;
    }
    if ((rc = upload(s, imgs, &s->d.images))) return bail(rc);
    std::vector<uint32_t> em(desc->emissions, desc->emissions + desc->num_emissions);
    if ((rc = upload(s, em, &s->d.emissions))) return bail(rc);
    s->d.num_emissions = desc->num_emissions;
    s->d.num_elements = desc->num_elements;
    std::vector<DImage> faces(6);
    for (int k = 0; k < 6; k++) faces[k] = imgs[desc->skybox_images[k]];
    if ((rc = upload(s, faces, &s->d.sky_faces))) return bail(rc);
    s->d.sky_r = desc->skybox_intensity.x; s->d.sky_g = desc->skybox_intensity.y; s->d.sky_b = desc->skybox_intensity.z;
    s->d.eps = desc->config.eps; s->d.offset = desc->config.offset; s->d.inf = desc->config.inf; s->d.gamma = desc->config.gamma_factor;
    s->d.bounce_limit = desc->config.bounce_limit; s->d.supersampling = desc->config.supersampling;
    double R = 0.0;
    for (int k = 0; k < 3; k++) {
        s->d.bounds_lo[k] = b.bounds.empty ? 0.0 : b.bounds.lo[k] - b.pad;
        s->d.bounds_hi[k] = b.bounds.empty ? 0.0 : b.bounds.hi[k] + b.pad;
        R = std::fmax(R, std::fmax(std::fabs(s->d.bounds_lo[k]), std::fabs(s->d.bounds_hi[k])));
    }
    s->d.far_limit = (float)(4.0 * R);
    s->d.scene_r = f32_up(R);
    s->num_nodes = (uint32_t)b.nodes.size(); s->num_tris = (uint32_t)b.tris.size();
    s->num_elements = desc->num_elements; s->num_emissions = desc->num_emissions; s->num_images = desc->num_images;
    s->tree_depth = b.depth;
    *out = s;
    return 0;
}

}  // namespace hnm
#endif
