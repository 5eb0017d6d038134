// hnm_device.cuh -- device-side arithmetic of the radiance loop.
//
// Every function here evaluates the SAME sequence of IEEE f64 operations as the
// reference function it cites (and as oracle/oracle.cpp), so that with
// -fmad=false the device result is bit-identical to the oracle's "det" flavour.
// Transcendentals come from hnm_detmath.h (deterministic, shared with the oracle).
// What is deliberately NOT like the reference is everything that cannot change a
// result: data layout, traversal order, culling, f32 conservative box tests.
#ifndef HNM_DEVICE_CUH
#define HNM_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "hanamaru_b200.h"
#include "hnm_detmath.h"

namespace hnm {

#define HNM_D __device__ __forceinline__

// ---------------------------------------------------------------- src/vector.rs
struct D3 {
    double x, y, z;
};
HNM_D D3 d3(double x, double y, double z) { return D3{x, y, z}; }
HNM_D D3 d3(const hnm_vec3& a) { return D3{a.x, a.y, a.z}; }
HNM_D D3 splat(double v) { return D3{v, v, v}; }
HNM_D D3 operator+(D3 a, D3 b) { return D3{a.x + b.x, a.y + b.y, a.z + b.z}; }
HNM_D D3 operator-(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
HNM_D D3 operator*(D3 a, D3 b) { return D3{a.x * b.x, a.y * b.y, a.z * b.z}; }
HNM_D D3 operator/(D3 a, D3 b) { return D3{a.x / b.x, a.y / b.y, a.z / b.z}; }
HNM_D D3 operator*(D3 a, double s) { return D3{a.x * s, a.y * s, a.z * s}; }
HNM_D D3 operator*(double s, D3 a) { return a * s; }
HNM_D D3 operator/(D3 a, double s) { return D3{a.x / s, a.y / s, a.z / s}; }
HNM_D D3 operator-(D3 a) { return D3{-a.x, -a.y, -a.z}; }
HNM_D bool all_zero(D3 a) { return a.x == 0.0 && a.y == 0.0 && a.z == 0.0; }  // `== Vector3::zero()`
HNM_D double norm(D3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
HNM_D double length(D3 a) { return __dsqrt_rn(norm(a)); }
HNM_D D3 normalize(D3 a) {
    double inv_len = 1.0 / length(a);
    return D3{a.x * inv_len, a.y * inv_len, a.z * inv_len};
}
HNM_D double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HNM_D D3 cross(D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
HNM_D D3 reflect(D3 v, D3 n) { return v - 2.0 * dot(v, n) * n; }
HNM_D D3 refract(D3 v, D3 n, double ri) {
    double k = 1.0 - ri * ri * (1.0 - dot(n, v) * dot(v, n));
    if (k < 0.0) return splat(0.0);
    return ri * v - (ri * dot(v, n) + __dsqrt_rn(k)) * n;
}

// ---------------------------------------------------------------- src/math.rs
HNM_D bool signbit_(double v) { return __double2hiint(v) < 0; }
HNM_D double saturate(double v) { return fmin(fmax(v, 0.0), 1.0); }  // f64::max/min: NaN-ignoring, like fmax/fmin
HNM_D D3 saturate(D3 a) { return D3{saturate(a.x), saturate(a.y), saturate(a.z)}; }
HNM_D double det(D3 a, D3 b, D3 c) {
    return (a.x * b.y * c.z) + (a.y * b.z * c.x) + (a.z * b.x * c.y) - (a.x * b.z * c.y) - (a.y * b.x * c.z) - (a.z * b.y * c.x);
}
HNM_D double signum(double v) {
    if (v != v) return v;
    return signbit_(v) ? -1.0 : 1.0;
}
HNM_D uint32_t f64_as_u32(double v) {  // Rust `as u32`: saturating, NaN -> 0 (cvt.rzi.u32.f64 saturates; NaN -> 0)
    return __double2uint_rz(v);
}
HNM_D uint64_t f64_as_u64(double v) { return __double2ull_rz(v); }
HNM_D uint32_t clamp_u32(uint32_t x, uint32_t mn, uint32_t mx) { return x < mn ? mn : (x > mx ? mx : x); }

#define HNM_PI 3.14159265358979323846
#define HNM_PI2 (2.0 * HNM_PI)

// ---------------------------------------------------------------- device scene
struct DTexture {
    double r, g, b;
    int32_t image;
    int32_t _pad;
};
struct DMaterial {
    DTexture albedo, emission, roughness;
    double param;
    int32_t surface;
    int32_t has_image;  // any of the three textures is image-backed
};
struct DImage {
    cudaTextureObject_t tex;  // uchar4, point sampled, unnormalised coordinates
    uint32_t width, height;
};
struct DElement {
    double ax, ay, az, bx, by, bz, radius;
    int32_t kind, material;
    uint32_t seq;  // position in the reference's top-level DFS order (tie breaking)
    int32_t mesh;
};
// 64-byte BVH node: both children's boxes (f32, rounded outward and padded) + child links.
// link >= 0: inner node; link < 0: ~link = kind<<29 | count<<26 | first
struct __align__(16) DNode {
    float lo0x, lo0y, lo0z, hi0x;
    float hi0y, hi0z, lo1x, lo1y;
    float lo1z, hi1x, hi1y, hi1z;
    int32_t c0, c1;
    uint32_t _pad0, _pad1;
};
enum { LEAF_TRI = 0, LEAF_SPHERE = 1, LEAF_CUBOID = 2, LEAF_NONE = 3 };
HNM_D int leaf_kind(int32_t link) { return (int)((uint32_t)(~link) >> 29); }
HNM_D uint32_t leaf_count(int32_t link) { return ((uint32_t)(~link) >> 26) & 7u; }
HNM_D uint32_t leaf_first(int32_t link) { return (uint32_t)(~link) & 0x3FFFFFFu; }

// triangle in leaf order: v0, edge1 = v1 - v0, edge2 = v2 - v0 (the reference's own subtractions, done once)
struct __align__(16) DTri {  // 80 bytes: five 16-byte loads
    double v0x, v0y, v0z, e1x, e1y, e1z, e2x, e2y, e2z, _pad;
};
HNM_D DTri load_tri(const DTri* p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4);
    DTri t;
    t.v0x = a.x; t.v0y = a.y; t.v0z = b.x; t.e1x = b.y; t.e1y = c.x; t.e1z = c.y; t.e2x = d.x; t.e2y = d.y; t.e2z = e.x; t._pad = 0.0;
    return t;
}

// One node of the HOST's (reference-topology) trees, top level first, then every mesh: f64 box as the reference built
// it + link to the parent.  The reference tests a primitive only if every box from the root down to its leaf passes
// Aabb::intersect_ray (src/bvh.rs:214,240); see chain_pass below.
struct RefNode {
    double box[6];    // min xyz, max xyz
    uint32_t parent;  // REF_NONE at the top-level root; a mesh root's parent is the top-level leaf that lists its element
    uint32_t _pad;
};
constexpr uint32_t REF_NONE = 0xFFFFFFFFu;

struct DScene {
    const DNode* nodes;
    const DTri* tris;
    const float4* trif;        // f32 copy for the conservative pre-test: 3 x float4 per triangle = v0, e1, e2, e1 x e2,
                               // in TRAVERSAL order (the leaves of the GPU-side tree index this array)
    const uint32_t* tri_perm;  // traversal position -> index into tris[] (the reference's leaf order)
    float scene_r;             // max |coordinate| of the scene box (error bound of the f32 origin)
    const uint32_t* tri_elem;  // element id per triangle
    const uint32_t* tri_face;  // face index inside its mesh
    const DElement* elements;
    const float4* elemf;       // 4 x float4 per element for the f32 pre-tests (hnm_trace.cuh): sphere = (centre, radius);
                               // cuboid = box rounded outward (lo, hi), box rounded inward (lo, hi)
    const DMaterial* materials;
    const double* unorm8;      // [256]: i / 255.0
    const DImage* images;
    const DImage* sky_faces;  // px nx py ny pz nz (device memory: indexed at run time)
    const uint32_t* emissions;
    uint32_t num_emissions, num_elements;
    double sky_r, sky_g, sky_b;
    double eps, offset, inf, gamma;
    uint32_t bounce_limit, supersampling;
    float far_limit;  // |origin| beyond this: advance the ray to the scene box before the f32 traversal
    double bounds_lo[3], bounds_hi[3];
    // the reference's box chain (chain_pass)
    const double* tri_box;      // [6] per triangle (reference leaf order): the box of the mesh leaf that lists it
    const uint32_t* tri_leaf;   // that leaf's index in ref_nodes
    const double* elem_box;     // [6] per element: the box of the top-level leaf that lists it
    const uint32_t* elem_leaf;
    const RefNode* ref_nodes;
    uint32_t chain_full;        // the caller's boxes are not nested (child inside parent): always walk the whole chain
};

// closest hit, before material resolution
struct Hit {
    double t, u, v;
    uint32_t kind;  // LEAF_TRI / LEAF_SPHERE / LEAF_CUBOID / LEAF_NONE
    uint32_t id;    // triangle index (leaf order) or element id
};

// f64::min / f64::max of the reference's toolchain (Rust 1.20 .. 1.36 libcore: min = if other.is_nan() || self < other
// { self } else { other }, max = if self.is_nan() || self < other { other } else { self }): NaN-ignoring like fmin / fmax,
// but for operands that compare equal (+0.0 vs -0.0) min keeps the SECOND and max the FIRST operand, where the hardware
// min / max order -0.0 below +0.0.  Only the slab test's sign-of-tmax check can see the difference.
HNM_D double rs_min(double a, double b) { return (b != b || a < b) ? a : b; }
HNM_D double rs_max(double a, double b) { return (a != a || a < b) ? b : a; }

// ---------------------------------------------------------------- the reference's box chain
// src/bvh.rs:20-39 on a stored box with the ray's reciprocal direction; also returns tmax
HNM_D bool ref_box_hit(const double* __restrict__ b, D3 o, double ix, double iy, double iz, double& tmax) {
    const double t1 = (__ldg(b + 0) - o.x) * ix, t2 = (__ldg(b + 3) - o.x) * ix;
    const double t3 = (__ldg(b + 1) - o.y) * iy, t4 = (__ldg(b + 4) - o.y) * iy;
    const double t5 = (__ldg(b + 2) - o.z) * iz, t6 = (__ldg(b + 5) - o.z) * iz;
    const double tmin = rs_max(rs_max(rs_min(t1, t2), rs_min(t3, t4)), rs_min(t5, t6));
    tmax = rs_min(rs_min(rs_max(t1, t2), rs_max(t3, t4)), rs_max(t5, t6));
    return tmin <= tmax && !(__double2hiint(tmax) < 0);
}
HNM_D bool ref_chain_full(const RefNode* __restrict__ nodes, uint32_t node, D3 o, double ix, double iy, double iz) {
    while (node != REF_NONE) {
        double tmax;
        if (!ref_box_hit(nodes[node].box, o, ix, iy, iz, tmax)) return false;
        node = nodes[node].parent;
    }
    return true;
}
// Would the reference have reached this primitive?  It tests a primitive only after EVERY box from the root of the top-level
// tree down to the primitive's leaf passed Aabb::intersect_ray for this ray (src/bvh.rs:214,240), and that test is
// not the geometric one: 0 * inf = NaN when a direction component is zero and the origin lies on a box plane, and the
// roundings of a grazing ray can make tmin > tmax for a box the ray does touch.  The GPU traversal (another tree,
// conservative f32 boxes) only proposes candidates, so the chain is re-checked here for every primitive whose exact test
// passed.  `box6` = the primitive's LEAF box: boxes are nested (the builder merges children, src/bvh.rs:79-105) and
// fl((plane - o) * inv) is monotone in `plane`, so with finite reciprocals tmin(ancestor) <= tmin(leaf) and
// tmax(ancestor) >= tmax(leaf): the leaf passing implies every ancestor passing -- except for the sign of a zero tmax,
// and except when a reciprocal is infinite (NaN terms are dropped by min / max): those walk the whole chain.
// (Out of line: it runs once per ACCEPTED candidate, about once per ray; inlined into the candidate loops it cost the
// confirming kernels registers they do not have.)
__device__ __noinline__ bool chain_pass_impl(const RefNode* __restrict__ nodes, uint32_t chain_full, const double* __restrict__ box6,
                                             const uint32_t* __restrict__ leaf, double ox, double oy, double oz, double dx, double dy, double dz) {
    const D3 o = D3{ox, oy, oz};
    const double ix = 1.0 / dx, iy = 1.0 / dy, iz = 1.0 / dz;
    const bool finite = fabs(ix) <= 1.79769313486231570815e308 && fabs(iy) <= 1.79769313486231570815e308 && fabs(iz) <= 1.79769313486231570815e308;
    if (finite && !chain_full) {
        double tmax;
        if (!ref_box_hit(box6, o, ix, iy, iz, tmax)) return false;
        if (tmax != 0.0) return true;
    }
    return ref_chain_full(nodes, __ldg(leaf), o, ix, iy, iz);
}
HNM_D bool chain_pass(const DScene& sc, const double* __restrict__ box6, const uint32_t* __restrict__ leaf, D3 o, D3 dir) {
    return chain_pass_impl(sc.ref_nodes, sc.chain_full, box6, leaf, o.x, o.y, o.z, dir.x, dir.y, dir.z);
}
// a ray for which the f32 candidate search must not cull by "certain" hits: its chain tests may produce NaN terms
HNM_D bool ray_needs_exact_path(const DScene& sc, D3 dir) {
    const double ix = 1.0 / dir.x, iy = 1.0 / dir.y, iz = 1.0 / dir.z;
    const bool finite = fabs(ix) <= 1.79769313486231570815e308 && fabs(iy) <= 1.79769313486231570815e308 && fabs(iz) <= 1.79769313486231570815e308;
    return !finite || sc.chain_full != 0u;
}

// ---------------------------------------------------------------- primitive tests (f64, reference arithmetic)
// src/bvh.rs:266-290 with the running `intersection.distance` = best.t.  The reference keeps the LAST
// candidate in DFS order among exact ties (`t > distance` rejects); with any visiting order that is
// "larger leaf-order index wins" (SURVEY 7.2 hard part 3).
HNM_D void tri_test(const DScene& sc, const DTri& tr, uint32_t g, D3 o, D3 dir, Hit& best) {
    D3 ray_inv = -dir;
    D3 edge1 = d3(tr.e1x, tr.e1y, tr.e1z);
    D3 edge2 = d3(tr.e2x, tr.e2y, tr.e2z);
    double denominator = det(edge1, edge2, ray_inv);
    if (denominator == 0.0) return;
    double denominator_inv = 1.0 / denominator;
    D3 d = o - d3(tr.v0x, tr.v0y, tr.v0z);
    double u = det(d, edge2, ray_inv) * denominator_inv;
    if (u < 0.0 || u > 1.0) return;
    double v = det(edge1, d, ray_inv) * denominator_inv;
    if (v < 0.0 || u + v > 1.0) return;
    double t = det(edge1, edge2, d) * denominator_inv;
    if (t < 0.0 || t > best.t) return;
    if (t == best.t && best.kind == LEAF_TRI && g < best.id) return;  // an earlier triangle loses an exact tie
    if (!chain_pass(sc, sc.tri_box + 6 * (size_t)g, sc.tri_leaf + g, o, dir)) return;  // the reference never got to this triangle
    best.t = t; best.u = u; best.v = v; best.kind = LEAF_TRI; best.id = g;
}

// src/scene.rs:58-78 (acceptance only; normal/uv are recomputed for the winner)
HNM_D void sphere_test(const DScene& sc, const DElement& e, uint32_t elem, const DElement* elements, D3 o, D3 dir, Hit& best) {
    D3 a = o - d3(e.ax, e.ay, e.az);
    double b = dot(a, dir);
    double c = dot(a, a) - e.radius * e.radius;
    double d = b * b - c;
    double t = -b - __dsqrt_rn(d);
    if (d > 0.0 && t > 0.0) {
        bool take = t < best.t;
        // exact tie with a later non-triangle element: the reference would have kept this (earlier) one
        if (!take && t == best.t && best.kind != LEAF_TRI && best.kind != LEAF_NONE && e.seq < elements[best.id].seq) take = true;
        if (take && chain_pass(sc, sc.elem_box + 6 * (size_t)elem, sc.elem_leaf + elem, o, dir)) {
            best.t = t; best.u = 0.0; best.v = 0.0; best.kind = LEAF_SPHERE; best.id = elem;
        }
    }
}

// src/bvh.rs:20-39 in f64 (used by Cuboid::intersect, src/scene.rs:152-183)
HNM_D bool aabb_intersect_ray(double mnx, double mny, double mnz, double mxx, double mxy, double mxz, D3 o, D3 dir, double* distance) {
    double ix = 1.0 / dir.x, iy = 1.0 / dir.y, iz = 1.0 / dir.z;
    double t1 = (mnx - o.x) * ix;
    double t2 = (mxx - o.x) * ix;
    double t3 = (mny - o.y) * iy;
    double t4 = (mxy - o.y) * iy;
    double t5 = (mnz - o.z) * iz;
    double t6 = (mxz - o.z) * iz;
    double tmin = rs_max(rs_max(rs_min(t1, t2), rs_min(t3, t4)), rs_min(t5, t6));
    double tmax = rs_min(rs_min(rs_max(t1, t2), rs_max(t3, t4)), rs_max(t5, t6));
    bool hit = tmin <= tmax && !signbit_(tmax);
    *distance = !signbit_(tmin) ? tmin : tmax;
    return hit;
}
HNM_D void cuboid_test(const DScene& sc, const DElement& e, uint32_t elem, const DElement* elements, D3 o, D3 dir, Hit& best) {
    double distance;
    bool hit = aabb_intersect_ray(e.ax, e.ay, e.az, e.bx, e.by, e.bz, o, dir, &distance);
    if (hit) {
        bool take = distance < best.t;
        if (!take && distance == best.t && best.kind != LEAF_TRI && best.kind != LEAF_NONE && e.seq < elements[best.id].seq) take = true;
        if (take && chain_pass(sc, sc.elem_box + 6 * (size_t)elem, sc.elem_leaf + elem, o, dir)) {
            best.t = distance; best.u = 0.0; best.v = 0.0; best.kind = LEAF_CUBOID; best.id = elem;
        }
    }
}

// ---------------------------------------------------------------- two-level BVH traversal
// f32 conservative slab tests decide only WHICH primitives get the exact f64 test: every primitive
// whose reference box chain the ray passes is still tested, so the closest hit is the reference's.
#define HNM_STACK 48
struct TraceStats {
    uint32_t nodes, prims;
};

template <bool STATS>
HNM_D Hit trace(const DScene& sc, D3 o, D3 dir, TraceStats* st) {
    Hit best;
    best.t = sc.inf; best.u = 0.0; best.v = 0.0; best.kind = LEAF_NONE; best.id = 0;

    // f32 copy of the ray for the box tests.  If the origin is far outside the scene box the f32
    // origin would lose too many bits: advance it to the box first (box tests only).
    double t0 = 0.0;
    float fmaxo = fmaxf(fmaxf(fabsf((float)o.x), fabsf((float)o.y)), fabsf((float)o.z));
    if (fmaxo > sc.far_limit) {
        double dist;
        bool h = aabb_intersect_ray(sc.bounds_lo[0], sc.bounds_lo[1], sc.bounds_lo[2], sc.bounds_hi[0], sc.bounds_hi[1], sc.bounds_hi[2], o, dir, &dist);
        if (!h) return best;  // cannot hit anything: every primitive lies inside the scene box
        if (dist > 0.0 && dist < sc.inf) t0 = dist * (1.0 - 1e-6);
    }
    float ox = (float)(o.x + dir.x * t0), oy = (float)(o.y + dir.y * t0), oz = (float)(o.z + dir.z * t0);
    float ix = (float)(1.0 / dir.x), iy = (float)(1.0 / dir.y), iz = (float)(1.0 / dir.z);
    const float W = 4.76837158203125e-07f;  // 2^-21: covers the rounding of (lo-o)*inv in f32

    int32_t stack[HNM_STACK];
    int sp = 0;
    int32_t cur = 0;  // root is always an inner node
    for (;;) {
        if (cur >= 0) {
            const float4* np = reinterpret_cast<const float4*>(sc.nodes + cur);
            float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
            int4 n3 = __ldg(reinterpret_cast<const int4*>(np + 3));
            if (STATS) st->nodes++;
            float bestf = __double2float_ru(best.t - t0);
            // child 0: lo = (n0.x n0.y n0.z) hi = (n0.w n1.x n1.y)
            float a0 = (n0.x - ox) * ix, b0 = (n0.w - ox) * ix;
            float a1 = (n0.y - oy) * iy, b1 = (n1.x - oy) * iy;
            float a2 = (n0.z - oz) * iz, b2 = (n1.y - oz) * iz;
            float tmin0 = fmaxf(fmaxf(fminf(a0, b0), fminf(a1, b1)), fminf(a2, b2));
            float tmax0 = fminf(fminf(fmaxf(a0, b0), fmaxf(a1, b1)), fmaxf(a2, b2));
            // child 1: lo = (n1.z n1.w n2.x) hi = (n2.y n2.z n2.w)
            float c0 = (n1.z - ox) * ix, e0 = (n2.y - ox) * ix;
            float c1 = (n1.w - oy) * iy, e1 = (n2.z - oy) * iy;
            float c2 = (n2.x - oz) * iz, e2 = (n2.w - oz) * iz;
            float tmin1 = fmaxf(fmaxf(fminf(c0, e0), fminf(c1, e1)), fminf(c2, e2));
            float tmax1 = fminf(fminf(fmaxf(c0, e0), fmaxf(c1, e1)), fmaxf(c2, e2));
            float lo0 = tmin0 - fabsf(tmin0) * W, up0 = tmax0 + fabsf(tmax0) * W;
            float lo1 = tmin1 - fabsf(tmin1) * W, up1 = tmax1 + fabsf(tmax1) * W;
            bool h0 = (lo0 <= up0) && (up0 >= 0.0f) && (lo0 <= bestf);
            bool h1 = (lo1 <= up1) && (up1 >= 0.0f) && (lo1 <= bestf);
            if (h0 && h1) {
                bool swap = lo1 < lo0;
                int32_t nearc = swap ? n3.y : n3.x;
                int32_t farc = swap ? n3.x : n3.y;
                if (sp < HNM_STACK) stack[sp++] = farc;
                cur = nearc;
                continue;
            } else if (h0) {
                cur = n3.x;
                continue;
            } else if (h1) {
                cur = n3.y;
                continue;
            }
        } else {
            int kind = leaf_kind(cur);
            uint32_t first = leaf_first(cur);
            if (kind == LEAF_TRI) {
                uint32_t cnt = leaf_count(cur);
                for (uint32_t k = 0; k < cnt; k++) {
                    uint32_t g = __ldg(sc.tri_perm + first + k);
                    DTri tr = load_tri(sc.tris + g);
                    if (STATS) st->prims++;
                    tri_test(sc, tr, g, o, dir, best);
                }
            } else if (kind == LEAF_SPHERE) {
                if (STATS) st->prims++;
                sphere_test(sc, sc.elements[first], first, sc.elements, o, dir, best);
            } else if (kind == LEAF_CUBOID) {
                if (STATS) st->prims++;
                cuboid_test(sc, sc.elements[first], first, sc.elements, o, dir, best);
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return best;
}

// ---------------------------------------------------------------- src/texture.rs, src/color.rs
// `(texel as f64) / 255.0` (src/color.rs:18-24) for the 256 possible texel values, computed on the host with the same
// correctly rounded IEEE division: a 2 KB table in global memory (L1-resident, DScene::unorm8) instead of twelve f64
// divisions per bilinear sample (ncu, round 1: 9 % of the instructions of k_shade_surf<NEE>)
HNM_D uchar4 texel_screen(const DImage& im, uint32_t x, uint32_t y) {  // src/texture.rs:59-63
    x = clamp_u32(x, 0u, im.width - 1u);
    y = clamp_u32(im.height - y - 1u, 0u, im.height - 1u);  // wraps at y == height, then clamps
    return tex2D<uchar4>(im.tex, (float)x + 0.5f, (float)y + 0.5f);
}
// One out-of-line copy per module: it is called from every shading kernel.  ptxas 12.9 (sm_100a, -O3) has produced a
// wrong THIRD component of this function's result in some kernels for some shapes of this code (PTX correct, SASS
// wrong); tests/test_gpu_parity.py pins every kernel that calls it bit-for-bit.
#ifndef HNM_BILINEAR_ATTR
#define HNM_BILINEAR_ATTR __device__ __noinline__
#endif
HNM_BILINEAR_ATTR void sample_bilinear_to(const double* __restrict__ unorm8, double gamma, DImage im, double u, double v,
                                                double* out) {  // src/texture.rs:29-49
    const double x = u * (double)im.width;
    const double y = v * (double)im.height;
    const double x1 = floor(x), y1 = floor(y);
    const double x2 = x1 + 1.0, y2 = y1 + 1.0;
    const uchar4 t11 = texel_screen(im, f64_as_u32(x1), f64_as_u32(y1));
    const uchar4 t12 = texel_screen(im, f64_as_u32(x1), f64_as_u32(y2));
    const uchar4 t21 = texel_screen(im, f64_as_u32(x2), f64_as_u32(y1));
    const uchar4 t22 = texel_screen(im, f64_as_u32(x2), f64_as_u32(y2));
    const double ax = x2 - x, bx = x - x1, ay = y2 - y, by = y - y1;
    const double den = (x2 - x1) * (y2 - y1);
    // per channel, in the reference's order: ((p11*(x2-x))*(y2-y) + (p21*(x-x1))*(y2-y) + (p12*(x2-x))*(y-y1) + (p22*(x-x1))*(y-y1)) / den
    const unsigned char c11[3] = {t11.x, t11.y, t11.z}, c12[3] = {t12.x, t12.y, t12.z};
    const unsigned char c21[3] = {t21.x, t21.y, t21.z}, c22[3] = {t22.x, t22.y, t22.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double p11 = __ldg(unorm8 + c11[k]), p12 = __ldg(unorm8 + c12[k]);
        const double p21 = __ldg(unorm8 + c21[k]), p22 = __ldg(unorm8 + c22[k]);
        const double num = p11 * ax * ay + p21 * bx * ay + p12 * ax * by + p22 * bx * by;
        const double g = den == 1.0 ? num : num / den;  // den = ((x1 + 1) - x1) * ((y1 + 1) - y1) is exactly 1 below 2^53
        out[k] = dm::pow(g, gamma);  // gamma_to_linear
    }
}
// FAST (opt-in perf mode, hnm_set_precision): the texture path of the shading kernels in f32 -- bilinear blend, powf --
// plus sincosf per BSDF sample and acosf for sphere uv, instead of the f64 blend and the deterministic double-double
// library.  Results are then statistically, not bit-wise, equal to the reference's.  A function of its own: the exact one
// above must keep the code shape the parity tests pinned.
#ifndef HNM_FAST_POW
#define HNM_FAST_POW 1
#endif
#ifndef HNM_FAST_ACOS
#define HNM_FAST_ACOS 1
#endif
#ifndef HNM_FAST_SINCOS
#define HNM_FAST_SINCOS 1
#endif
__device__ __noinline__ float3 sample_bilinear_fast(float gamma, DImage im, float u, float v) {
    const float x = u * (float)im.width, y = v * (float)im.height;
    const float x1 = floorf(x), y1 = floorf(y);
    const uint32_t ix = (uint32_t)fmaxf(x1, 0.0f), iy = (uint32_t)fmaxf(y1, 0.0f);
    const uchar4 t11 = texel_screen(im, ix, iy), t12 = texel_screen(im, ix, iy + 1u);
    const uchar4 t21 = texel_screen(im, ix + 1u, iy), t22 = texel_screen(im, ix + 1u, iy + 1u);
    const float bx = x - x1, by = y - y1, ax = 1.0f - bx, ay = 1.0f - by;
    const float w11 = ax * ay * (1.0f / 255.0f), w21 = bx * ay * (1.0f / 255.0f), w12 = ax * by * (1.0f / 255.0f), w22 = bx * by * (1.0f / 255.0f);
    float3 c;
    c.x = powf(t11.x * w11 + t21.x * w21 + t12.x * w12 + t22.x * w22, gamma);
    c.y = powf(t11.y * w11 + t21.y * w21 + t12.y * w12 + t22.y * w22, gamma);
    c.z = powf(t11.z * w11 + t21.z * w21 + t12.z * w12 + t22.z * w22, gamma);
    return c;
}
template <bool FAST = false>
HNM_D D3 sample_bilinear(const double* unorm8, double gamma, DImage im, double u, double v) {
    if (FAST && HNM_FAST_POW) {
        const float3 c = sample_bilinear_fast((float)gamma, im, (float)u, (float)v);
        return d3((double)c.x, (double)c.y, (double)c.z);
    }
    double r[3];
    sample_bilinear_to(unorm8, gamma, im, u, v, r);
    return d3(r[0], r[1], r[2]);
}
template <bool FAST = false>
HNM_D D3 texture_sample(const DScene& sc, const DTexture& t, double u, double v) {  // src/texture.rs:108-114
    if (t.image >= 0) return sample_bilinear<FAST>(sc.unorm8, sc.gamma, sc.images[t.image], u, v) * d3(t.r, t.g, t.b);
    return d3(t.r, t.g, t.b);
}
// src/scene.rs:295-319
template <bool FAST = false>
HNM_D D3 skybox_sample(const DScene& sc, D3 direction) {
    double abs_x = fabs(direction.x), abs_y = fabs(direction.y), abs_z = fabs(direction.z);
    int face;
    double u, v;
    if (abs_x > abs_y && abs_x > abs_z) {
        if (!signbit_(direction.x)) { face = 0; u = -direction.z / direction.x; v = direction.y / direction.x; }
        else { face = 1; u = -direction.z / direction.x; v = -direction.y / direction.x; }
    } else if (abs_y > abs_x && abs_y > abs_z) {
        if (!signbit_(direction.y)) { face = 2; u = direction.x / direction.y; v = -direction.z / direction.y; }
        else { face = 3; u = -direction.x / direction.y; v = -direction.z / direction.y; }
    } else {
        if (!signbit_(direction.z)) { face = 4; u = direction.x / direction.z; v = direction.y / direction.z; }
        else { face = 5; u = direction.x / direction.z; v = -direction.y / direction.z; }
    }
    // sample_bilinear_0center (src/texture.rs:22-26)
    // the six faces live in device memory: a run-time index into a kernel-PARAMETER array makes nvcc 12.9 spill the
    // parameter struct to local memory, and the copy it generated in one kernel was wrong (blue intensity garbage)
    D3 c = sample_bilinear<FAST>(sc.unorm8, sc.gamma, sc.sky_faces[face], 0.5 * (u + 1.0), 0.5 * (v + 1.0));
    return d3(sc.sky_r, sc.sky_g, sc.sky_b) * c;
}

// ---------------------------------------------------------------- surface point of a hit
struct PointMaterial {  // src/material.rs:25-31
    D3 albedo, emission;
    double roughness, param;
    int32_t surface;
};
struct SurfacePoint {
    D3 position, normal;
    double u, v;
    int32_t element, face;
};
// recompute what the winning `intersect` call wrote into the Intersection (same formulas, same bits)
template <bool FAST = false>
HNM_D SurfacePoint surface_point(const DScene& sc, const Hit& h, D3 o, D3 dir, bool need_uv) {
    SurfacePoint s;
    s.position = o + dir * h.t;
    s.u = h.u; s.v = h.v; s.face = -1;
    if (h.kind == LEAF_TRI) {
        const DTri& tr = sc.tris[h.id];
        s.normal = normalize(cross(d3(tr.e1x, tr.e1y, tr.e1z), d3(tr.e2x, tr.e2y, tr.e2z)));  // src/bvh.rs:286
        s.element = (int32_t)sc.tri_elem[h.id];
        s.face = (int32_t)sc.tri_face[h.id];
    } else if (h.kind == LEAF_SPHERE) {
        const DElement& e = sc.elements[h.id];
        s.element = (int32_t)h.id;
        s.normal = normalize(s.position - d3(e.ax, e.ay, e.az));  // src/scene.rs:67
        if (need_uv) {  // src/scene.rs:69-73 (only observable through image textures)
            s.v = 1.0 - ((FAST && HNM_FAST_ACOS) ? (double)acosf((float)s.normal.y) : dm::acos(s.normal.y)) / HNM_PI;
            double xz_len = __dsqrt_rn(s.normal.x * s.normal.x + s.normal.z * s.normal.z);
            s.u = 0.5 - signum(s.normal.z) * ((FAST && HNM_FAST_ACOS) ? (double)acosf((float)(s.normal.x / xz_len)) : dm::acos(s.normal.x / xz_len)) / HNM_PI2;
        }
    } else {  // cuboid, src/scene.rs:156-181
        const DElement& e = sc.elements[h.id];
        s.element = (int32_t)h.id;
        D3 uvw = (s.position - d3(e.ax, e.ay, e.az)) / (d3(e.bx, e.by, e.bz) - d3(e.ax, e.ay, e.az));
        s.normal = splat(0.0);  // the reference leaves a stale normal if no face matches (unreachable: the hit lies on a face)
        if (fabs(s.position.y - e.by) < sc.eps) { s.normal = d3(0.0, 1.0, 0.0); s.u = uvw.x; s.v = 1.0 - uvw.z; }
        else if (fabs(s.position.y - e.ay) < sc.eps) { s.normal = d3(0.0, -1.0, 0.0); s.u = uvw.x; s.v = 1.0 - uvw.z; }
        else if (fabs(s.position.x - e.ax) < sc.eps) { s.normal = d3(-1.0, 0.0, 0.0); s.u = uvw.z; s.v = uvw.y; }
        else if (fabs(s.position.x - e.bx) < sc.eps) { s.normal = d3(1.0, 0.0, 0.0); s.u = uvw.z; s.v = uvw.y; }
        else if (fabs(s.position.z - e.az) < sc.eps) { s.normal = d3(0.0, 0.0, -1.0); s.u = uvw.x; s.v = uvw.y; }
        else if (fabs(s.position.z - e.bz) < sc.eps) { s.normal = d3(0.0, 0.0, 1.0); s.u = uvw.x; s.v = uvw.y; }
    }
    return s;
}
// src/scene.rs:389-395
template <bool FAST = false>
HNM_D PointMaterial resolve_material(const DScene& sc, const DMaterial& m, double u, double v) {
    PointMaterial pm;
    pm.surface = m.surface; pm.param = m.param;
    pm.albedo = texture_sample<FAST>(sc, m.albedo, u, v);
    pm.emission = texture_sample<FAST>(sc, m.emission, u, v);
    pm.roughness = texture_sample<FAST>(sc, m.roughness, u, v).x;
    return pm;
}

// ---------------------------------------------------------------- src/material.rs
struct Rand2 { double r0, r1; };
HNM_D bool nee_available(int32_t surface) { return surface == HNM_SURFACE_DIFFUSE || surface == HNM_SURFACE_GGX; }
HNM_D void tangent_space_basis(const DScene& sc, D3 normal, D3& tangent, D3& binormal) {  // :202-211
    D3 up = fabs(normal.x) > sc.eps ? d3(0.0, 1.0, 0.0) : d3(1.0, 0.0, 0.0);
    tangent = normalize(cross(up, normal));
    binormal = cross(normal, tangent);
}
HNM_D D3 importance_sample_diffuse(const DScene& sc, Rand2 random, double cos_phi, double sin_phi, D3 normal) {  // :227-248
    D3 tangent, binormal;
    tangent_space_basis(sc, normal, tangent, binormal);
    return (tangent * cos_phi + binormal * sin_phi) * __dsqrt_rn(random.r1) + normal * __dsqrt_rn(1.0 - random.r1);
}
HNM_D D3 importance_sample_ggx_half(const DScene& sc, Rand2 random, double cos_phi, double sin_phi, D3 normal, double alpha2) {  // :260-269
    D3 tangent, binormal;
    tangent_space_basis(sc, normal, tangent, binormal);
    double cos_theta = __dsqrt_rn((1.0 - random.r1) / (1.0 + (alpha2 - 1.0) * random.r1));
    double sin_theta = __dsqrt_rn(1.0 - cos_theta * cos_theta);
    D3 h = d3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta);
    return tangent * h.x + binormal * h.y + normal * h.z;
}
HNM_D double g_smith_joint_lambda(double x_dot_n, double alpha2) {
    double a = 1.0 / (x_dot_n * x_dot_n) - 1.0;
    return 0.5 * __dsqrt_rn(1.0 + alpha2 * a) - 0.5;
}
HNM_D double g_smith_joint(double l_dot_n, double v_dot_n, double alpha2) {
    double lambda_l = g_smith_joint_lambda(l_dot_n, alpha2);
    double lambda_v = g_smith_joint_lambda(v_dot_n, alpha2);
    return 1.0 / (1.0 + lambda_l + lambda_v);
}
HNM_D double powi5(double x) {
    double x2 = x * x;
    double x4 = x2 * x2;
    return x * x4;
}
HNM_D double f_schlick(double v_dot_h, double f0) { return f0 + (1.0 - f0) * powi5(1.0 - v_dot_h); }

HNM_D double bsdf(const PointMaterial& m, D3 view, D3 normal, D3 light) {  // :53-89
    if (m.surface == HNM_SURFACE_DIFFUSE) return 1.0 / HNM_PI;
    double f0 = m.param;
    double alpha2 = m.roughness * m.roughness;
    D3 half = normalize(light + view);
    double l_dot_n = dot(light, normal);
    if (signbit_(l_dot_n)) return 0.0;
    double v_dot_n = dot(view, normal);
    double v_dot_h = dot(view, half);
    double h_dot_n = dot(half, normal);
    double tmp = 1.0 - (1.0 - alpha2) * h_dot_n * h_dot_n;
    double d = alpha2 / (HNM_PI * tmp * tmp);
    double g = g_smith_joint(l_dot_n, v_dot_n, alpha2);
    double f = f_schlick(v_dot_h, f0);
    return d * g * f / (4.0 * l_dot_n * v_dot_n);
}

struct SampleResult {
    D3 origin, direction;
    double reflectance;
};
HNM_D bool sample_refraction(const DScene& sc, Rand2 random, D3 position, D3 view, D3 normal, double refractive_index, SampleResult& out) {  // :154-199
    bool is_incoming = signbit_(dot(view, normal));
    D3 oriented_normal = is_incoming ? normal : -normal;
    double nnt = is_incoming ? 1.0 / refractive_index : refractive_index;
    D3 reflect_direction = reflect(view, oriented_normal);
    D3 refract_direction = refract(view, oriented_normal, nnt);
    if (all_zero(refract_direction)) {
        out.origin = position + sc.offset * oriented_normal;
        out.direction = reflect_direction;
        out.reflectance = 1.0;
        return true;
    }
    double cos_i = dot(view, -oriented_normal);
    double cos_t = dot(refract_direction, -oriented_normal);
    double r_s = (nnt * cos_i - cos_t) * (nnt * cos_i - cos_t) / ((nnt * cos_i + cos_t) * (nnt * cos_i + cos_t));
    double r_p = (nnt * cos_t - cos_i) * (nnt * cos_t - cos_i) / ((nnt * cos_t + cos_i) * (nnt * cos_t + cos_i));
    double fr = 0.5 * (r_s + r_p);
    if (random.r0 <= fr) {
        out.origin = position + sc.offset * oriented_normal;
        out.direction = reflect_direction;
        out.reflectance = 1.0;
    } else {
        out.origin = position - sc.offset * oriented_normal;
        out.direction = refract_direction;
        out.reflectance = nnt * nnt;
    }
    return true;
}
// `PointMaterial::sample` (:91-151).  cos_phi/sin_phi = cos/sin(PI2 * random.0), shared with the light sample.
HNM_D bool material_sample(const DScene& sc, const PointMaterial& m, Rand2 random, double cos_phi, double sin_phi, D3 position, D3 view,
                           D3 normal, SampleResult& out) {
    D3 ray = -view;
    switch (m.surface) {
        case HNM_SURFACE_DIFFUSE:
            out.origin = position + normal * sc.offset;
            out.direction = importance_sample_diffuse(sc, random, cos_phi, sin_phi, normal);
            out.reflectance = 1.0;
            return true;
        case HNM_SURFACE_SPECULAR:
            out.origin = position + normal * sc.offset;
            out.direction = reflect(ray, normal);
            out.reflectance = 1.0;
            return true;
        case HNM_SURFACE_REFRACTION:
            return sample_refraction(sc, random, position, ray, normal, m.param, out);
        case HNM_SURFACE_GGX: {
            double f0 = m.param;
            double alpha2 = m.roughness * m.roughness;
            D3 half = importance_sample_ggx_half(sc, random, cos_phi, sin_phi, normal, alpha2);
            D3 next_direction = reflect(ray, half);
            double l_dot_n = dot(next_direction, normal);
            if (signbit_(l_dot_n)) return false;
            double v_dot_n = dot(view, normal);
            double v_dot_h = dot(view, half);
            double h_dot_n = dot(half, normal);
            double g = g_smith_joint(l_dot_n, v_dot_n, alpha2);
            double f = f_schlick(v_dot_h, f0);
            out.origin = position + normal * sc.offset;
            out.direction = next_direction;
            out.reflectance = f * saturate(g * v_dot_h / (h_dot_n * v_dot_n));
            return true;
        }
        default: {
            double alpha2 = m.roughness * m.roughness;
            D3 half = importance_sample_ggx_half(sc, random, cos_phi, sin_phi, normal, alpha2);
            return sample_refraction(sc, random, position, ray, half, m.param, out);
        }
    }
}

// ---------------------------------------------------------------- rand 0.4: u64 -> f64 in [0,1)
HNM_D double u64_to_f64(uint64_t w) {
    return __longlong_as_double((long long)(0x3FF0000000000000ull | (w & 0xFFFFFFFFFFFFFull))) - 1.0;
}

}  // namespace hnm
#endif
