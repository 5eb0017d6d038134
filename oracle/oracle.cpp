// oracle.cpp -- CPU restatement of hanamaru-renderer's radiance loop.
//
// TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library.  The
// product path (libhanamaru_b200.so) never calls into it.
//
// What it is: a function-by-function f64 restatement of the reference's hot
// path, each function citing the file:line it follows (reference checkout at
// /root/reference, commit f292ed36).  The reference is Rust and cannot be built
// here (no cargo/rustc, no vendored crates: SURVEY F1/F3), so this is a "port"
// oracle, pinned by
//   * rand's own ISAAC-64 known-answer vectors (tests/test_oracle.py),
//   * the reference's golden image rtcamp6_1000x4spp.png (tests/golden/, made by
//     tools/make_golden.py), statistically,
//   * the reference's other published image rtcamp5.png: its 42 diamonds are placed by StdRng::gen_range, and the
//     host's restatement of rand 0.4.3 reproduces the layout diamond for diamond -- that pins the u64->f64 mapping
//     (tests/test_oracle.py::test_rtcamp5_layout_matches_published_image),
//   * hand-checked vectors for the pure functions (tests/test_oracle.py).
// Parity UNPINNED at two steps only, both third-party code absent from the tree:
// the order of the two draws of rand 0.4.3's `gen::<(f64, f64)>()` (its tuple_impl! macro builds the tuple
// expression `(rng.gen(), rng.gen())`, which Rust evaluates left to right; restated, not testable with an image),
// and the image crate's JPEG decoder (decoded texels are inputs here).
//
// Input is the same flat hnm_scene_desc the CUDA core consumes (the host builds
// the BVH with the reference's algorithm and flattens it in DFS order, so the
// recursion below visits nodes, faces and elements in the reference's order).
//
// Two build flavours (oracle/Makefile):
//   liboracle.so      transcendental functions from glibc, as Rust's std does;
//   liboracle_det.so  -DORACLE_DETMATH: the device code's deterministic
//                     sin/cos/exp/pow/acos (hnm_detmath.h), which makes the GPU
//                     comparison bit-exact; tests bound glibc-vs-det separately.
// Build: g++ -O2 -ffp-contract=off (no FMA contraction, no fast-math), OpenMP
// over pixels where the reference uses rayon (src/renderer.rs:33).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "hanamaru_b200.h"

#ifdef ORACLE_DETMATH
#include "hnm_detmath.h"
namespace om {
static inline double sin(double x) { return hnm::dm::sin(x); }
static inline double cos(double x) { return hnm::dm::cos(x); }
static inline double exp(double x) { return hnm::dm::exp(x); }
static inline double pow(double x, double y) { return hnm::dm::pow(x, y); }
static inline double acos(double x) { return hnm::dm::acos(x); }
}  // namespace om
#define ORACLE_FLAVOR "detmath"
#else
namespace om {
static inline double sin(double x) { return std::sin(x); }
static inline double cos(double x) { return std::cos(x); }
static inline double exp(double x) { return std::exp(x); }
static inline double pow(double x, double y) { return std::pow(x, y); }
static inline double acos(double x) { return std::acos(x); }
}  // namespace om
#define ORACLE_FLAVOR "glibc"
#endif

namespace {

// ---------------------------------------------------------------- src/vector.rs
struct V3 {
    double x, y, z;
};
inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 v3(const hnm_vec3& a) { return V3{a.x, a.y, a.z}; }
inline V3 from_one(double v) { return V3{v, v, v}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 operator*(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(double s, V3 a) { return a * s; }  // src/vector.rs:176-182
inline V3 operator/(V3 a, double s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline double norm(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }       // :35-37 (squared)
inline double length(V3 a) { return std::sqrt(norm(a)); }                    // :31-33
inline V3 normalize(V3 a) {                                                  // :39-46
    double inv_len = 1.0 / length(a);
    return V3{a.x * inv_len, a.y * inv_len, a.z * inv_len};
}
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // :48-50
inline V3 cross(V3 a, V3 b) {                                                // :52-58
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline V3 reflect(V3 v, V3 n) { return v - 2.0 * dot(v, n) * n; }            // :60-62
inline V3 refract(V3 v, V3 n, double ri) {                                   // :64-71
    double k = 1.0 - ri * ri * (1.0 - dot(n, v) * dot(v, n));
    if (k < 0.0) return from_one(0.0);
    return ri * v - (ri * dot(v, n) + std::sqrt(k)) * n;
}
inline bool approximately(V3 a, V3 b, double offset) { return norm(a - b) < offset * 4.0; }  // :89-91

// ---------------------------------------------------------------- src/math.rs
inline double clampf(double v, double mn, double mx) { return std::fmin(std::fmax(v, mn), mx); }  // :9-11 (f64::max/min ignore NaN)
inline uint32_t clamp_u32(uint32_t x, uint32_t mn, uint32_t mx) { return x < mn ? mn : (x > mx ? mx : x); }  // :13-15
inline double saturate(double v) { return clampf(v, 0.0, 1.0); }                                  // :17-19
inline V3 saturate(V3 a) { return V3{saturate(a.x), saturate(a.y), saturate(a.z)}; }
inline double det(V3 a, V3 b, V3 c) {                                                             // :25-32
    return (a.x * b.y * c.z) + (a.y * b.z * c.x) + (a.z * b.x * c.y) - (a.x * b.z * c.y) - (a.y * b.x * c.z) - (a.z * b.y * c.x);
}
inline double signum(double v) {  // f64::signum: NaN -> NaN, else copysign(1, v)
    if (v != v) return v;
    return std::signbit(v) ? -1.0 : 1.0;
}
inline uint32_t f64_as_u32(double v) {  // Rust `as u32`: saturating, NaN -> 0
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}
inline uint64_t f64_as_u64(double v) {  // Rust `as usize`
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 18446744073709551615.0) return ~0ull;
    return (uint64_t)v;
}
inline uint8_t f64_as_u8(double v) {
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}

// ---------------------------------------------------------------- rand 0.4.3 StdRng = Isaac64Rng
// (third-party; restated from the crate's published src/prng/isaac64.rs and
// src/lib.rs -- Cargo.lock pins rand 0.3.22 -> 0.4.3).
struct Isaac64 {
    uint64_t rsl[256], mem[256];
    uint64_t a, b, c;
    uint32_t cnt;

    // SeedableRng<&[u64]>::from_seed: rsl = seed ++ zeros, a=b=c=0, init(true)
    void from_seed(const uint64_t* seed, int n) {
        for (int i = 0; i < 256; i++) rsl[i] = i < n ? seed[i] : 0;
        cnt = 0;
        a = b = c = 0;
        init();
    }
    static inline void mix(uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d, uint64_t& e, uint64_t& f, uint64_t& g, uint64_t& h) {
        a -= e; f ^= h >> 9;  h += a;
        b -= f; g ^= a << 9;  a += b;
        c -= g; h ^= b >> 23; b += c;
        d -= h; a ^= c << 15; c += d;
        e -= a; b ^= d >> 14; d += e;
        f -= b; c ^= e << 20; e += f;
        g -= c; d ^= f >> 17; f += g;
        h -= d; e ^= g << 14; g += h;
    }
    void init() {
        uint64_t a_, b_, c_, d_, e_, f_, g_, h_;
        a_ = b_ = c_ = d_ = e_ = f_ = g_ = h_ = 0x9e3779b97f4a7c13ull;
        for (int i = 0; i < 4; i++) mix(a_, b_, c_, d_, e_, f_, g_, h_);
        for (int pass = 0; pass < 2; pass++) {
            const uint64_t* src = pass == 0 ? rsl : mem;
            for (int i = 0; i < 256; i += 8) {
                a_ += src[i]; b_ += src[i + 1]; c_ += src[i + 2]; d_ += src[i + 3];
                e_ += src[i + 4]; f_ += src[i + 5]; g_ += src[i + 6]; h_ += src[i + 7];
                mix(a_, b_, c_, d_, e_, f_, g_, h_);
                mem[i] = a_; mem[i + 1] = b_; mem[i + 2] = c_; mem[i + 3] = d_;
                mem[i + 4] = e_; mem[i + 5] = f_; mem[i + 6] = g_; mem[i + 7] = h_;
            }
        }
        isaac64();
    }
    void isaac64() {
        c += 1;
        uint64_t aa = a, bb = b + c;
        for (int half = 0; half < 2; half++) {
            int mr = half == 0 ? 0 : 128, m2 = half == 0 ? 128 : 0;
            for (int base = 0; base < 128; base += 4) {
                for (int j = 0; j < 4; j++) {
                    uint64_t mixv;
                    switch (j) {
                        case 0: mixv = ~(aa ^ (aa << 21)); break;
                        case 1: mixv = aa ^ (aa >> 5); break;
                        case 2: mixv = aa ^ (aa << 12); break;
                        default: mixv = aa ^ (aa >> 33); break;
                    }
                    uint64_t x = mem[base + j + mr];
                    aa = mixv + mem[base + j + m2];
                    uint64_t y = mem[(x >> 3) & 255] + aa + bb;
                    mem[base + j + mr] = y;
                    bb = mem[(y >> 11) & 255] + x;
                    rsl[base + j + mr] = bb;
                }
            }
        }
        a = aa;
        b = bb;
        cnt = 256;
    }
    uint64_t next_u64() {
        if (cnt == 0) isaac64();
        cnt -= 1;
        return rsl[cnt & 255];
    }
    // Rng::next_f64 (rand 0.4): mantissa bits under exponent 0 -> [1,2) - 1
    double next_f64() {
        uint64_t tmp = 0x3FF0000000000000ull | (next_u64() & 0xFFFFFFFFFFFFFull);
        double r;
        memcpy(&r, &tmp, 8);
        return r - 1.0;
    }
};
struct Rand2 { double r0, r1; };
inline Rand2 gen_pair(Isaac64& rng) {  // `rng.gen::<(f64, f64)>()`: left then right
    Rand2 r;
    r.r0 = rng.next_f64();
    r.r1 = rng.next_f64();
    return r;
}

// ---------------------------------------------------------------- counters
struct Counters {
    uint64_t paths = 0, segments = 0, shadow_rays = 0, node_visits = 0, tri_tests = 0, elem_tests = 0, lens_iters = 0;
};
thread_local Counters* tl_counters = nullptr;
#define COUNT(field) do { if (tl_counters) tl_counters->field++; } while (0)

// ---------------------------------------------------------------- scene view
struct Ray { V3 origin, direction; };

struct PointMaterial {  // src/material.rs:25-31
    int32_t surface; double param;
    V3 albedo, emission; double roughness;
};
struct Intersection {   // src/scene.rs:10-17
    V3 position; double distance; V3 normal; double u, v;
    PointMaterial material;
    int32_t element, face;  // bookkeeping only
};
struct Ctx {
    const hnm_scene_desc* d;
    double EPS, OFFSET, INF, GAMMA;
};
const double PI = 3.14159265358979323846;  // f64::consts::PI
const double PI2 = 2.0 * PI;               // src/config.rs:5

Intersection empty_intersection(const Ctx& c) {  // src/scene.rs:26-39
    Intersection i;
    i.position = from_one(0.0); i.distance = c.INF; i.normal = from_one(0.0); i.u = 0.0; i.v = 0.0;
    i.material.surface = HNM_SURFACE_DIFFUSE; i.material.param = 0.0;
    i.material.albedo = from_one(1.0); i.material.emission = from_one(0.0); i.material.roughness = 0.2;
    i.element = -1; i.face = -1;
    return i;
}

// f64::min / f64::max as the reference's toolchain defined them (Rust 1.20 .. 1.36, libcore/num/f64.rs, restated from
// memory -- std is not in the tree):  min = (if other.is_nan() || self < other { self } else { other }) * 1.0,
// max = (if self.is_nan() || self < other { other } else { self }) * 1.0.  Both ignore a NaN operand like C's fmin / fmax;
// they differ from glibc only for operands that compare equal, i.e. +0.0 against -0.0: min keeps the SECOND operand,
// max the FIRST.  Observable only in the slab test below (`tmax.is_sign_positive()`), and only when the ray origin lies
// exactly on box planes of two axes.  Parity UNPINNED at this step (no artefact of the reference exercises it).
inline double rs_min(double a, double b) { return (b != b || a < b) ? a : b; }
inline double rs_max(double a, double b) { return (a != a || a < b) ? b : a; }

// ---------------------------------------------------------------- src/bvh.rs:20-39
inline bool aabb_intersect_ray(const double* mn, const double* mx, const Ray& ray, double* distance) {
    V3 dir_inv = v3(1.0 / ray.direction.x, 1.0 / ray.direction.y, 1.0 / ray.direction.z);
    double t1 = (mn[0] - ray.origin.x) * dir_inv.x;
    double t2 = (mx[0] - ray.origin.x) * dir_inv.x;
    double t3 = (mn[1] - ray.origin.y) * dir_inv.y;
    double t4 = (mx[1] - ray.origin.y) * dir_inv.y;
    double t5 = (mn[2] - ray.origin.z) * dir_inv.z;
    double t6 = (mx[2] - ray.origin.z) * dir_inv.z;
    double tmin = rs_max(rs_max(rs_min(t1, t2), rs_min(t3, t4)), rs_min(t5, t6));  // (t1.min(t2).max(t3.min(t4))).max(t5.min(t6))
    double tmax = rs_min(rs_min(rs_max(t1, t2), rs_max(t3, t4)), rs_max(t5, t6));
    bool hit = tmin <= tmax && !std::signbit(tmax);
    *distance = !std::signbit(tmin) ? tmin : tmax;
    return hit;
}

// ---------------------------------------------------------------- src/bvh.rs:266-290
inline bool intersect_polygon(V3 v0, V3 v1, V3 v2, const Ray& ray, Intersection& isect) {
    COUNT(tri_tests);
    V3 ray_inv = -ray.direction;
    V3 edge1 = v1 - v0;
    V3 edge2 = v2 - v0;
    double denominator = det(edge1, edge2, ray_inv);
    if (denominator == 0.0) return false;
    double denominator_inv = 1.0 / denominator;
    V3 d = ray.origin - v0;
    double u = det(d, edge2, ray_inv) * denominator_inv;
    if (u < 0.0 || u > 1.0) return false;
    double v = det(edge1, d, ray_inv) * denominator_inv;
    if (v < 0.0 || u + v > 1.0) return false;
    double t = det(edge1, edge2, d) * denominator_inv;
    if (t < 0.0 || t > isect.distance) return false;
    isect.position = ray.origin + ray.direction * t;
    isect.normal = normalize(cross(edge1, edge2));
    isect.distance = t;
    isect.u = u; isect.v = v;
    return true;
}

// ---------------------------------------------------------------- src/bvh.rs:213-237
bool intersect_for_mesh(const Ctx& c, const hnm_mesh& m, uint32_t node_rel, const Ray& ray, Intersection& isect) {
    const hnm_bvh_node& n = c.d->mesh_nodes[m.node_offset + node_rel];
    COUNT(node_visits);
    double dist;
    if (!aabb_intersect_ray(n.aabb_min, n.aabb_max, ray, &dist)) return false;
    bool any_hit = false;
    if (n.child0 < 0) {
        for (uint32_t k = 0; k < n.count; k++) {
            uint32_t face = c.d->mesh_indices[m.index_offset + n.first + k];
            const uint32_t* f = c.d->faces + 3 * (size_t)(m.face_offset + face);
            const double* vb = c.d->vertices + 3 * (size_t)m.vertex_offset;
            V3 v0 = v3(vb[3 * f[0]], vb[3 * f[0] + 1], vb[3 * f[0] + 2]);
            V3 v1 = v3(vb[3 * f[1]], vb[3 * f[1] + 1], vb[3 * f[1] + 2]);
            V3 v2 = v3(vb[3 * f[2]], vb[3 * f[2] + 1], vb[3 * f[2] + 2]);
            if (intersect_polygon(v0, v1, v2, ray, isect)) {
                any_hit = true;
                isect.face = (int32_t)face;
            }
        }
    } else {
        if (intersect_for_mesh(c, m, (uint32_t)n.child0, ray, isect)) any_hit = true;
        if (intersect_for_mesh(c, m, (uint32_t)n.child1, ray, isect)) any_hit = true;
    }
    return any_hit;
}

// ---------------------------------------------------------------- src/scene.rs:58-78
bool sphere_intersect(const hnm_element& e, const Ray& ray, Intersection& isect) {
    V3 center = v3(e.a);
    V3 a = ray.origin - center;
    double b = dot(a, ray.direction);
    double cc = dot(a, a) - e.radius * e.radius;
    double d = b * b - cc;
    double t = -b - std::sqrt(d);
    if (d > 0.0 && t > 0.0 && t < isect.distance) {
        isect.position = ray.origin + ray.direction * t;
        isect.distance = t;
        isect.normal = normalize(isect.position - center);
        isect.v = 1.0 - om::acos(isect.normal.y) / PI;
        double xz_len = std::sqrt(isect.normal.x * isect.normal.x + isect.normal.z * isect.normal.z);  // Vector2::length
        isect.u = 0.5 - signum(isect.normal.z) * om::acos(isect.normal.x / xz_len) / PI2;
        return true;
    }
    return false;
}

// ---------------------------------------------------------------- src/scene.rs:152-183
bool cuboid_intersect(const Ctx& c, const hnm_element& e, const Ray& ray, Intersection& isect) {
    double mn[3] = {e.a.x, e.a.y, e.a.z}, mx[3] = {e.b.x, e.b.y, e.b.z};
    double distance;
    bool hit = aabb_intersect_ray(mn, mx, ray, &distance);
    if (hit && distance < isect.distance) {
        isect.position = ray.origin + ray.direction * distance;
        isect.distance = distance;
        V3 uvw = (isect.position - v3(e.a)) / (v3(e.b) - v3(e.a));
        auto equals_eps = [&](double a, double b) { return std::fabs(a - b) < c.EPS; };  // src/math.rs:21-23
        if (equals_eps(isect.position.y, e.b.y)) {
            isect.normal = v3(0.0, 1.0, 0.0); isect.u = uvw.x; isect.v = 1.0 - uvw.z;  // xiz
        } else if (equals_eps(isect.position.y, e.a.y)) {
            isect.normal = v3(0.0, -1.0, 0.0); isect.u = uvw.x; isect.v = 1.0 - uvw.z;
        } else if (equals_eps(isect.position.x, e.a.x)) {
            isect.normal = v3(-1.0, 0.0, 0.0); isect.u = uvw.z; isect.v = uvw.y;       // zy
        } else if (equals_eps(isect.position.x, e.b.x)) {
            isect.normal = v3(1.0, 0.0, 0.0); isect.u = uvw.z; isect.v = uvw.y;
        } else if (equals_eps(isect.position.z, e.a.z)) {
            isect.normal = v3(0.0, 0.0, -1.0); isect.u = uvw.x; isect.v = uvw.y;       // xy
        } else if (equals_eps(isect.position.z, e.b.z)) {
            isect.normal = v3(0.0, 0.0, 1.0); isect.u = uvw.x; isect.v = uvw.y;
        }
        return true;
    }
    return false;
}

bool element_intersect(const Ctx& c, uint32_t index, const Ray& ray, Intersection& isect) {
    const hnm_element& e = c.d->elements[index];
    COUNT(elem_tests);
    switch (e.kind) {
        case HNM_ELEM_SPHERE: return sphere_intersect(e, ray, isect);
        case HNM_ELEM_CUBOID: return cuboid_intersect(c, e, ray, isect);
        default: return intersect_for_mesh(c, c.d->meshes[e.mesh], 0, ray, isect);  // src/scene.rs:242-244
    }
}

// ---------------------------------------------------------------- src/bvh.rs:239-263
int32_t intersect_for_scene(const Ctx& c, uint32_t node, const Ray& ray, Intersection& isect) {
    const hnm_bvh_node& n = c.d->top_nodes[node];
    COUNT(node_visits);
    double dist;
    if (!aabb_intersect_ray(n.aabb_min, n.aabb_max, ray, &dist)) return -1;
    int32_t nearest = -1;
    if (n.child0 < 0) {
        for (uint32_t k = 0; k < n.count; k++) {
            uint32_t index = c.d->top_indices[n.first + k];
            if (element_intersect(c, index, ray, isect)) nearest = (int32_t)index;
        }
    } else {
        int32_t r0 = intersect_for_scene(c, (uint32_t)n.child0, ray, isect);
        if (r0 >= 0) nearest = r0;
        int32_t r1 = intersect_for_scene(c, (uint32_t)n.child1, ray, isect);
        if (r1 >= 0) nearest = r1;
    }
    return nearest;
}

// ---------------------------------------------------------------- src/texture.rs, src/color.rs
inline V3 rgba_to_color(const uint8_t* p) { return v3(p[0] / 255.0, p[1] / 255.0, p[2] / 255.0); }  // src/color.rs:18-24
inline V3 gamma_to_linear(const Ctx& c, V3 g) { return v3(om::pow(g.x, c.GAMMA), om::pow(g.y, c.GAMMA), om::pow(g.z, c.GAMMA)); }  // :26-36
V3 sample_nearest_screen(const hnm_image& im, uint32_t x, uint32_t y) {  // src/texture.rs:59-63
    x = clamp_u32(x, 0, im.width - 1);
    y = clamp_u32(im.height - y - 1u, 0, im.height - 1);  // u32 wrap at y == height, then clamp
    return rgba_to_color(im.rgba + 4 * ((size_t)y * im.width + x));
}
V3 sample_bilinear(const Ctx& c, const hnm_image& im, double u, double v) {  // src/texture.rs:29-49
    double x = u * (double)im.width;
    double y = v * (double)im.height;
    double x1 = std::floor(x), y1 = std::floor(y);
    double x2 = x1 + 1.0, y2 = y1 + 1.0;
    V3 p11 = sample_nearest_screen(im, f64_as_u32(x1), f64_as_u32(y1));
    V3 p12 = sample_nearest_screen(im, f64_as_u32(x1), f64_as_u32(y2));
    V3 p21 = sample_nearest_screen(im, f64_as_u32(x2), f64_as_u32(y1));
    V3 p22 = sample_nearest_screen(im, f64_as_u32(x2), f64_as_u32(y2));
    V3 gamma = (p11 * (x2 - x) * (y2 - y) + p21 * (x - x1) * (y2 - y) + p12 * (x2 - x) * (y - y1) + p22 * (x - x1) * (y - y1)) /
               ((x2 - x1) * (y2 - y1));
    return gamma_to_linear(c, gamma);
}
V3 sample_bilinear_0center(const Ctx& c, const hnm_image& im, double u, double v) {  // src/texture.rs:22-26
    return sample_bilinear(c, im, 0.5 * (u + 1.0), 0.5 * (v + 1.0));
}
V3 texture_sample(const Ctx& c, const hnm_texture& t, double u, double v) {  // src/texture.rs:108-114
    if (t.image >= 0) return sample_bilinear(c, c.d->images[t.image], u, v) * v3(t.color);
    return v3(t.color);
}

// ---------------------------------------------------------------- src/scene.rs:295-319
V3 skybox_sample(const Ctx& c, V3 direction) {
    const hnm_scene_desc* d = c.d;
    V3 intensity = v3(d->skybox_intensity);
    double abs_x = std::fabs(direction.x), abs_y = std::fabs(direction.y), abs_z = std::fabs(direction.z);
    auto img = [&](int k) -> const hnm_image& { return d->images[d->skybox_images[k]]; };
    if (abs_x > abs_y && abs_x > abs_z) {
        if (!std::signbit(direction.x)) return intensity * sample_bilinear_0center(c, img(0), -direction.z / direction.x, direction.y / direction.x);
        return intensity * sample_bilinear_0center(c, img(1), -direction.z / direction.x, -direction.y / direction.x);
    } else if (abs_y > abs_x && abs_y > abs_z) {
        if (!std::signbit(direction.y)) return intensity * sample_bilinear_0center(c, img(2), direction.x / direction.y, -direction.z / direction.y);
        return intensity * sample_bilinear_0center(c, img(3), -direction.x / direction.y, -direction.z / direction.y);
    } else {
        if (!std::signbit(direction.z)) return intensity * sample_bilinear_0center(c, img(4), direction.x / direction.z, direction.y / direction.z);
        return intensity * sample_bilinear_0center(c, img(5), direction.x / direction.z, -direction.y / direction.z);
    }
}

// ---------------------------------------------------------------- src/scene.rs:385-401
bool scene_intersect(const Ctx& c, const Ray& ray, Intersection& isect) {
    isect = empty_intersection(c);
    int32_t nearest = c.d->num_top_nodes ? intersect_for_scene(c, 0, ray, isect) : -1;
    if (nearest >= 0) {
        const hnm_material& m = c.d->materials[c.d->elements[nearest].material];
        isect.element = nearest;
        if (c.d->elements[nearest].kind != HNM_ELEM_MESH) isect.face = -1;
        isect.material.surface = m.surface;
        isect.material.param = m.param;
        isect.material.albedo = texture_sample(c, m.albedo, isect.u, isect.v);
        isect.material.emission = texture_sample(c, m.emission, isect.u, isect.v);
        isect.material.roughness = texture_sample(c, m.roughness, isect.u, isect.v).x;
        return true;
    }
    isect.element = -1; isect.face = -1;
    isect.material.emission = skybox_sample(c, ray.direction);
    return false;
}

// ---------------------------------------------------------------- src/material.rs
inline bool nee_available(const PointMaterial& m) { return m.surface == HNM_SURFACE_DIFFUSE || m.surface == HNM_SURFACE_GGX; }  // :42-51
inline double roughness_to_alpha2(double roughness) { double alpha = roughness; return alpha * alpha; }  // :250-255
inline void tangent_space_basis(const Ctx& c, V3 normal, V3& tangent, V3& binormal) {  // :202-211
    V3 up = std::fabs(normal.x) > c.EPS ? v3(0.0, 1.0, 0.0) : v3(1.0, 0.0, 0.0);
    tangent = normalize(cross(up, normal));
    binormal = cross(normal, tangent);
}
inline V3 importance_sample_diffuse(const Ctx& c, Rand2 random, V3 normal) {  // :227-248
    V3 tangent, binormal;
    tangent_space_basis(c, normal, tangent, binormal);
    double phi = PI2 * random.r0;
    return (tangent * om::cos(phi) + binormal * om::sin(phi)) * std::sqrt(random.r1) + normal * std::sqrt(1.0 - random.r1);
}
inline V3 importance_sample_ggx_half(const Ctx& c, Rand2 random, V3 normal, double alpha2) {  // :260-269
    V3 tangent, binormal;
    tangent_space_basis(c, normal, tangent, binormal);
    double phi = PI2 * random.r0;
    double cos_theta = std::sqrt((1.0 - random.r1) / (1.0 + (alpha2 - 1.0) * random.r1));
    double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
    V3 h = v3(sin_theta * om::cos(phi), sin_theta * om::sin(phi), cos_theta);
    return tangent * h.x + binormal * h.y + normal * h.z;
}
inline double g_smith_joint_lambda(double x_dot_n, double alpha2) {  // :271-274
    double a = 1.0 / (x_dot_n * x_dot_n) - 1.0;
    return 0.5 * std::sqrt(1.0 + alpha2 * a) - 0.5;
}
inline double g_smith_joint(double l_dot_n, double v_dot_n, double alpha2) {  // :276-280
    double lambda_l = g_smith_joint_lambda(l_dot_n, alpha2);
    double lambda_v = g_smith_joint_lambda(v_dot_n, alpha2);
    return 1.0 / (1.0 + lambda_l + lambda_v);
}
inline double powi5(double x) {  // f64::powi(5) = llvm.powi: x * ((x*x)*(x*x))
    double x2 = x * x;
    double x4 = x2 * x2;
    return x * x4;
}
inline double f_schlick(double v_dot_h, double f0) { return f0 + (1.0 - f0) * powi5(1.0 - v_dot_h); }  // :282-284

double bsdf(const PointMaterial& m, V3 view, V3 normal, V3 light) {  // :53-89
    if (m.surface == HNM_SURFACE_DIFFUSE) return 1.0 / PI;
    if (m.surface == HNM_SURFACE_GGX) {
        double f0 = m.param;
        double alpha2 = roughness_to_alpha2(m.roughness);
        V3 half = normalize(light + view);
        double l_dot_n = dot(light, normal);
        if (std::signbit(l_dot_n)) return 0.0;
        double v_dot_n = dot(view, normal);
        double v_dot_h = dot(view, half);
        double h_dot_n = dot(half, normal);
        double tmp = 1.0 - (1.0 - alpha2) * h_dot_n * h_dot_n;
        double d = alpha2 / (PI * tmp * tmp);
        double g = g_smith_joint(l_dot_n, v_dot_n, alpha2);
        double f = f_schlick(v_dot_h, f0);
        return d * g * f / (4.0 * l_dot_n * v_dot_n);
    }
    return std::nan("");  // unimplemented!() in the reference; unreachable behind nee_available
}

struct SampleResult { Ray ray; double reflectance; };

bool sample_refraction(const Ctx& c, Rand2 random, V3 position, V3 view, V3 normal, double refractive_index, SampleResult& out) {  // :154-199
    bool is_incoming = std::signbit(dot(view, normal));
    V3 oriented_normal = is_incoming ? normal : -normal;
    double nnt = is_incoming ? 1.0 / refractive_index : refractive_index;
    V3 reflect_direction = reflect(view, oriented_normal);
    V3 refract_direction = refract(view, oriented_normal, nnt);
    if (refract_direction == from_one(0.0)) {
        out.ray.origin = position + c.OFFSET * oriented_normal;
        out.ray.direction = reflect_direction;
        out.reflectance = 1.0;
        return true;
    }
    double cos_i = dot(view, -oriented_normal);
    double cos_t = dot(refract_direction, -oriented_normal);
    double r_s = (nnt * cos_i - cos_t) * (nnt * cos_i - cos_t) / ((nnt * cos_i + cos_t) * (nnt * cos_i + cos_t));
    double r_p = (nnt * cos_t - cos_i) * (nnt * cos_t - cos_i) / ((nnt * cos_t + cos_i) * (nnt * cos_t + cos_i));
    double fr = 0.5 * (r_s + r_p);
    if (random.r0 <= fr) {
        out.ray.origin = position + c.OFFSET * oriented_normal;
        out.ray.direction = reflect_direction;
        out.reflectance = 1.0;
    } else {
        out.ray.origin = position - c.OFFSET * oriented_normal;
        out.ray.direction = refract_direction;
        out.reflectance = nnt * nnt;
    }
    return true;
}

bool material_sample(const Ctx& c, const PointMaterial& m, Rand2 random, V3 position, V3 view, V3 normal, SampleResult& out) {  // :91-151
    V3 ray = -view;
    switch (m.surface) {
        case HNM_SURFACE_DIFFUSE:
            out.ray.origin = position + normal * c.OFFSET;
            out.ray.direction = importance_sample_diffuse(c, random, normal);
            out.reflectance = 1.0;
            return true;
        case HNM_SURFACE_SPECULAR:
            out.ray.origin = position + normal * c.OFFSET;
            out.ray.direction = reflect(ray, normal);
            out.reflectance = 1.0;
            return true;
        case HNM_SURFACE_REFRACTION:
            return sample_refraction(c, random, position, ray, normal, m.param, out);
        case HNM_SURFACE_GGX: {
            double f0 = m.param;
            double alpha2 = roughness_to_alpha2(m.roughness);
            V3 half = importance_sample_ggx_half(c, random, normal, alpha2);
            V3 next_direction = reflect(ray, half);
            double l_dot_n = dot(next_direction, normal);
            if (std::signbit(l_dot_n)) return false;
            double v_dot_n = dot(view, normal);
            double v_dot_h = dot(view, half);
            double h_dot_n = dot(half, normal);
            double g = g_smith_joint(l_dot_n, v_dot_n, alpha2);
            double f = f_schlick(v_dot_h, f0);
            out.ray.origin = position + normal * c.OFFSET;
            out.ray.direction = next_direction;
            out.reflectance = f * saturate(g * v_dot_h / (h_dot_n * v_dot_n));
            return true;
        }
        default: {  // GGXRefraction
            double alpha2 = roughness_to_alpha2(m.roughness);
            V3 half = importance_sample_ggx_half(c, random, normal, alpha2);
            return sample_refraction(c, random, position, ray, half, m.param, out);
        }
    }
}

// ---------------------------------------------------------------- src/scene.rs:92-101
struct Surface { V3 position, normal; double pdf; };
Surface sphere_sample_on_surface(const Ctx& c, const hnm_element& e, Rand2 random) {
    double theta = PI2 * random.r0;
    double unit_z = 1.0 - 2.0 * random.r1;
    double a = std::sqrt(1.0 - unit_z * unit_z);
    Surface s;
    s.normal = v3(a * om::cos(theta), a * om::sin(theta), unit_z);
    s.position = v3(e.a) + (e.radius + c.OFFSET) * s.normal;
    s.pdf = 1.0 / (4.0 * PI * e.radius * e.radius);
    return s;
}

// ---------------------------------------------------------------- src/camera.rs
V3 cam(const hnm_vec3& a) { return v3(a); }
Ray camera_ray_with_dof(const hnm_camera& cm, double ncx, double ncy, Isaac64& rng) {  // :66-96
    double sqx, sqy;
    for (;;) {
        COUNT(lens_iters);
        Rand2 r = gen_pair(rng);
        sqx = 2.0 * r.r0 - 1.0;
        sqy = 2.0 * r.r1 - 1.0;
        if (cm.lens_shape == 0) break;                // Square
        if (sqx * sqx + sqy * sqy < 1.0) break;       // Vector2::norm() is the SQUARED length
    }
    double lx = sqx * cm.lens_radius, ly = sqy * cm.lens_radius;
    V3 lens_pos = cam(cm.right) * lx + cam(cm.up) * ly;
    Ray r;
    r.origin = cam(cm.eye) + lens_pos;
    r.direction = normalize(ncx * cam(cm.plane_half_right) + ncy * cam(cm.plane_half_up) + cm.focus_distance * cam(cm.forward) - lens_pos);
    return r;
}
Ray camera_ray(const hnm_camera& cm, double ncx, double ncy) {  // :98-107
    Ray r;
    r.origin = cam(cm.eye);
    r.direction = normalize(ncx * cam(cm.plane_half_right) + ncy * cam(cm.plane_half_up) + cm.focus_distance * cam(cm.forward));
    return r;
}

// ---------------------------------------------------------------- src/renderer.rs:269-296
V3 next_event_estimation(const Ctx& c, Rand2 random, V3 position, V3 view, V3 normal, const PointMaterial& material) {
    V3 accumulation = from_one(0.0);
    for (uint32_t k = 0; k < c.d->num_emissions; k++) {
        const hnm_element& e = c.d->elements[c.d->emissions[k]];
        Surface surface = sphere_sample_on_surface(c, e, random);
        V3 shadow_vec = surface.position - position;
        V3 shadow_dir = normalize(shadow_vec);
        Ray shadow_ray{position, shadow_dir};
        Intersection si;
        COUNT(shadow_rays);
        bool shadow_hit = scene_intersect(c, shadow_ray, si);
        if (shadow_hit && approximately(si.position, surface.position, c.OFFSET)) {
            double dot_0 = std::fabs(dot(normal, shadow_dir));
            double dot_l = std::fabs(dot(surface.normal, shadow_dir));
            double distance_pow2 = dot(shadow_vec, shadow_vec);
            double g = (dot_0 * dot_l) / distance_pow2;
            double pdf = surface.pdf;
            accumulation = accumulation + si.material.emission * bsdf(material, view, normal, shadow_dir) * g / pdf;
        }
    }
    return accumulation * material.albedo;
}

// ---------------------------------------------------------------- src/renderer.rs:163-203
V3 pathtracing_calc_pixel(const Ctx& c, const hnm_camera& camera, double ncx, double ncy, uint32_t sampling) {
    COUNT(paths);
    uint64_t s = f64_as_u64((4.0 + ncx) * 100870.0);
    uint64_t t = f64_as_u64((4.0 + ncy) * 100304.0);
    uint64_t seed[4] = {8700304ull, (uint64_t)sampling, s, t};
    Isaac64 rng;
    rng.from_seed(seed, 4);
    Ray ray = camera_ray_with_dof(camera, ncx, ncy, rng);
    V3 accumulation = from_one(0.0);
    V3 reflectance = from_one(1.0);
    for (uint32_t b = 1; b < c.d->config.bounce_limit; b++) {
        Rand2 random = gen_pair(rng);
        Intersection isect;
        COUNT(segments);
        bool hit = scene_intersect(c, ray, isect);
        double current_reflectance = 1.0;
        if (hit) {
            V3 view = -ray.direction;
            SampleResult result;
            if (material_sample(c, isect.material, random, isect.position, view, isect.normal, result)) {
                if (nee_available(isect.material)) {
                    accumulation = accumulation + reflectance * next_event_estimation(c, random, result.ray.origin, view, isect.normal, isect.material);
                }
                ray = result.ray;
                current_reflectance = result.reflectance;
            } else {
                break;
            }
        }
        accumulation = accumulation + reflectance * isect.material.emission;
        reflectance = reflectance * (isect.material.albedo * current_reflectance);
        if (!hit || reflectance == from_one(0.0)) break;
    }
    return accumulation;
}

// ---------------------------------------------------------------- src/renderer.rs:116-139
V3 debug_calc_pixel(const Ctx& c, const hnm_camera& camera, double ncx, double ncy, int mode) {
    COUNT(paths);
    Ray ray = camera_ray(camera, ncx, ncy);
    V3 light_direction = normalize(v3(1.0, 2.0, -1.0));
    Intersection isect;
    COUNT(segments);
    bool hit = scene_intersect(c, ray, isect);
    if (!hit) return isect.material.emission;
    switch (mode) {
        case HNM_MODE_DEBUG_SHADING: {
            Ray shadow_ray{isect.position + isect.normal * c.OFFSET, light_direction};
            Intersection si;
            COUNT(shadow_rays);
            bool shadow_hit = scene_intersect(c, shadow_ray, si);
            double shadow = shadow_hit ? 0.5 : 1.0;
            double diffuse = std::fmax(dot(isect.normal, light_direction), 0.0);
            return isect.material.emission + isect.material.albedo * diffuse * shadow;
        }
        case HNM_MODE_DEBUG_NORMAL: return isect.normal;
        case HNM_MODE_DEBUG_DEPTH: return from_one(0.5 * isect.distance / camera.focus_distance);
        default: return from_one(std::fabs(isect.distance - camera.focus_distance));
    }
}

// ---------------------------------------------------------------- src/renderer.rs:48-60 (one sub-pixel)
inline void normalized_coord(uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t ss, uint32_t sx, uint32_t sy, double* ncx, double* ncy) {
    // frag_coord = (x, height - y)  (src/renderer.rs:36): y runs over 1..=H
    double fx = (double)x, fy = (double)(h - y);
    double offx = (double)sx / (double)ss - 0.5, offy = (double)sy / (double)ss - 0.5;
    double rx = (double)w, ry = (double)h;
    double m = std::fmin(rx, ry);
    *ncx = ((fx + offx) * 2.0 - rx) / m;
    *ncy = ((fy + offy) * 2.0 - ry) / m;
}
V3 calc_pixel(const Ctx& c, const hnm_camera& camera, int mode, double ncx, double ncy, uint32_t sampling) {
    if (mode == HNM_MODE_PATHTRACING) return pathtracing_calc_pixel(c, camera, ncx, ncy, sampling);
    return debug_calc_pixel(c, camera, ncx, ncy, mode);
}

Ctx make_ctx(const hnm_scene_desc* d) {
    Ctx c;
    c.d = d; c.EPS = d->config.eps; c.OFFSET = d->config.offset; c.INF = d->config.inf; c.GAMMA = d->config.gamma_factor;
    return c;
}

// ---------------------------------------------------------------- src/tonemap.rs, src/filter.rs, src/color.rs
inline V3 reinhard(V3 color, double exposure, double white_point) {  // src/tonemap.rs:22-27
    color = color * exposure;
    double luminance = 0.22 * color.x + 0.707 * color.y + 0.071 * color.z;  // src/color.rs:63-65
    white_point = white_point * exposure;
    return saturate(color * (luminance / (white_point * white_point) + 1.0) / (luminance + 1.0));
}
inline double gaussian(double x, double sigma) {  // src/filter.rs:13-15
    return om::exp(-(x * x) / (2.0 * sigma * sigma)) / (2.0 * PI * sigma * sigma);
}
inline double px_distance(uint32_t x, uint32_t y, uint32_t i, uint32_t j) {  // src/filter.rs:7-11 (wrapping u32, release build)
    uint32_t dx = x - i, dy = y - j;
    return std::sqrt((double)(uint32_t)(dx * dx + dy * dy));
}
V3 bilateral(const hnm_config& cfg, const std::vector<V3>& img, uint32_t current, uint32_t width, uint32_t height) {  // src/filter.rs:32-58
    V3 pixel = img[current];
    uint32_t x = current % width, y = current / width;
    double current_sum = pixel.x + pixel.y + pixel.z;
    double sum_scale = 1.0 / 3.0;
    V3 filtered = from_one(0.0);
    double w_p = 0.0;
    uint32_t diameter = cfg.bilateral_diameter;
    uint32_t half = diameter / 2;
    for (uint32_t i = 0; i < diameter; i++) {
        for (uint32_t j = 0; j < diameter; j++) {
            uint32_t neighbor_x = clamp_u32(x - (half - i), 0, width - 1);   // wrapping u32 arithmetic
            uint32_t neighbor_y = clamp_u32(y - (half - j), 0, height - 1);
            V3 neighbor = img[(size_t)neighbor_y * width + neighbor_x];
            double neighbor_sum = neighbor.x + neighbor.y + neighbor.z;
            double g_i = gaussian(sum_scale * (neighbor_sum - current_sum), cfg.bilateral_sigma_i);
            double g_s = gaussian(px_distance(x, y, neighbor_x, neighbor_y), cfg.bilateral_sigma_s);
            double w = g_i * g_s;
            filtered = filtered + neighbor * w;
            w_p += w;
        }
    }
    return filtered / w_p;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* oracle_flavor() { return ORACLE_FLAVOR; }

struct oracle_counters {
    uint64_t paths, segments, shadow_rays, node_visits, tri_tests, elem_tests, lens_iters;
};

// first `count` outputs of StdRng::from_seed(&seed[0..n])
void oracle_isaac64(const uint64_t* seed, uint32_t n, uint32_t count, uint64_t* out) {
    Isaac64 r;
    r.from_seed(seed, (int)n);
    for (uint32_t i = 0; i < count; i++) out[i] = r.next_u64();
}
void oracle_isaac64_f64(const uint64_t* seed, uint32_t n, uint32_t count, double* out) {
    Isaac64 r;
    r.from_seed(seed, (int)n);
    for (uint32_t i = 0; i < count; i++) out[i] = r.next_f64();
}
// skip `skip` outputs first (rand's second KAT)
void oracle_isaac64_skip(const uint64_t* seed, uint32_t n, uint32_t skip, uint32_t count, uint64_t* out) {
    Isaac64 r;
    r.from_seed(seed, (int)n);
    for (uint32_t i = 0; i < skip; i++) r.next_u64();
    for (uint32_t i = 0; i < count; i++) out[i] = r.next_u64();
}

void oracle_math(int fn, const double* x, const double* y, uint32_t n, double* out) {
    for (uint32_t i = 0; i < n; i++) {
        switch (fn) {
            case 0: out[i] = om::sin(x[i]); break;
            case 1: out[i] = om::cos(x[i]); break;
            case 2: out[i] = om::exp(x[i]); break;
            case 3: out[i] = om::pow(x[i], y[i]); break;
            default: out[i] = om::acos(x[i]); break;
        }
    }
}

void oracle_intersect_batch(const hnm_scene_desc* d, const hnm_ray* rays, uint32_t n, hnm_hit* hits) {
    Ctx c = make_ctx(d);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Ray r{v3(rays[i].origin), v3(rays[i].direction)};
        Intersection is;
        bool hit = scene_intersect(c, r, is);
        hnm_hit& h = hits[i];
        memset(&h, 0, sizeof(h));
        h.position = hnm_vec3{is.position.x, is.position.y, is.position.z};
        h.normal = hnm_vec3{is.normal.x, is.normal.y, is.normal.z};
        h.albedo = hnm_vec3{is.material.albedo.x, is.material.albedo.y, is.material.albedo.z};
        h.emission = hnm_vec3{is.material.emission.x, is.material.emission.y, is.material.emission.z};
        h.distance = is.distance; h.u = is.u; h.v = is.v;
        h.roughness = is.material.roughness; h.param = is.material.param;
        h.hit = hit ? 1 : 0; h.element = is.element; h.face = is.face; h.surface = is.material.surface;
    }
}

// layouts as hnm_material_sample_batch / hnm_material_bsdf_batch in include/hanamaru_b200.h
void oracle_material_sample_batch(const hnm_config* cfg, const double* in, uint32_t n, double* out) {
    hnm_scene_desc d;
    memset(&d, 0, sizeof(d));
    d.config = *cfg;
    Ctx c = make_ctx(&d);
    for (uint32_t i = 0; i < n; i++) {
        const double* p = in + 14 * (size_t)i;
        PointMaterial m;
        m.surface = (int32_t)p[0]; m.param = p[1]; m.roughness = p[2];
        m.albedo = from_one(1.0); m.emission = from_one(0.0);
        Rand2 rnd{p[3], p[4]};
        SampleResult r;
        memset(&r, 0, sizeof(r));
        bool some = material_sample(c, m, rnd, v3(p[5], p[6], p[7]), v3(p[8], p[9], p[10]), v3(p[11], p[12], p[13]), r);
        double* o = out + 8 * (size_t)i;
        o[0] = some ? 1.0 : 0.0;
        o[1] = some ? r.ray.origin.x : 0.0; o[2] = some ? r.ray.origin.y : 0.0; o[3] = some ? r.ray.origin.z : 0.0;
        o[4] = some ? r.ray.direction.x : 0.0; o[5] = some ? r.ray.direction.y : 0.0; o[6] = some ? r.ray.direction.z : 0.0;
        o[7] = some ? r.reflectance : 0.0;
    }
}
void oracle_material_bsdf_batch(const double* in, uint32_t n, double* out) {
    for (uint32_t i = 0; i < n; i++) {
        const double* p = in + 12 * (size_t)i;
        PointMaterial m;
        m.surface = (int32_t)p[0]; m.param = p[1]; m.roughness = p[2];
        m.albedo = from_one(1.0); m.emission = from_one(0.0);
        out[i] = bsdf(m, v3(p[3], p[4], p[5]), v3(p[6], p[7], p[8]), v3(p[9], p[10], p[11]));
    }
}

// Radiance of every camera path of one pass, [H][W][ss*ss][3] (sy-major), for per-path parity.
void oracle_render_paths(const hnm_scene_desc* d, const hnm_camera* camera, uint32_t w, uint32_t h, int mode, uint32_t sampling,
                         double* out, oracle_counters* counters) {
    Ctx c = make_ctx(d);
    uint32_t ss = d->config.supersampling;
    Counters total;
#pragma omp parallel
    {
        Counters local;
        tl_counters = &local;
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)w * h; i++) {
            uint32_t y = (uint32_t)(i / w), x = (uint32_t)(i - (int64_t)y * w);
            for (uint32_t sy = 0; sy < ss; sy++)
                for (uint32_t sx = 0; sx < ss; sx++) {
                    double ncx, ncy;
                    normalized_coord(x, y, w, h, ss, sx, sy, &ncx, &ncy);
                    V3 r = calc_pixel(c, *camera, mode, ncx, ncy, sampling);
                    double* o = out + ((size_t)i * ss * ss + sy * ss + sx) * 3;
                    o[0] = r.x; o[1] = r.y; o[2] = r.z;
                }
        }
        tl_counters = nullptr;
#pragma omp critical
        {
            total.paths += local.paths; total.segments += local.segments; total.shadow_rays += local.shadow_rays;
            total.node_visits += local.node_visits; total.tri_tests += local.tri_tests; total.elem_tests += local.elem_tests;
            total.lens_iters += local.lens_iters;
        }
    }
    if (counters) {
        counters->paths = total.paths; counters->segments = total.segments; counters->shadow_rays = total.shadow_rays;
        counters->node_visits = total.node_visits; counters->tri_tests = total.tri_tests; counters->elem_tests = total.elem_tests;
        counters->lens_iters = total.lens_iters;
    }
}

// `Renderer::render` (src/renderer.rs:25-46) for passes sampling_first .. +count-1 over image rows
// [row_begin, row_end): accum (f64 rgb, full image, row 0 = top) += supersampling(...)
// src/renderer.rs:25-46 restricted to the pixel rectangle [col_begin, col_end) x [row_begin, row_end) of the w x h image
// (every pixel is independent: the rectangle's pixels get exactly the values a full render gives them)
void oracle_render_rect(const hnm_scene_desc* d, const hnm_camera* camera, uint32_t w, uint32_t h, int mode, uint32_t sampling_first,
                        uint32_t count, uint32_t col_begin, uint32_t col_end, uint32_t row_begin, uint32_t row_end, double* accum,
                        oracle_counters* counters);
void oracle_render(const hnm_scene_desc* d, const hnm_camera* camera, uint32_t w, uint32_t h, int mode, uint32_t sampling_first,
                   uint32_t count, uint32_t row_begin, uint32_t row_end, double* accum, oracle_counters* counters) {
    oracle_render_rect(d, camera, w, h, mode, sampling_first, count, 0, w, row_begin, row_end, accum, counters);
}
void oracle_render_rect(const hnm_scene_desc* d, const hnm_camera* camera, uint32_t w, uint32_t h, int mode, uint32_t sampling_first,
                        uint32_t count, uint32_t col_begin, uint32_t col_end, uint32_t row_begin, uint32_t row_end, double* accum,
                        oracle_counters* counters) {
    Ctx c = make_ctx(d);
    const uint32_t cw = col_end - col_begin;
    uint32_t ss = d->config.supersampling;
    Counters total;
    for (uint32_t sampling = sampling_first; sampling < sampling_first + count; sampling++) {
#pragma omp parallel
        {
            Counters local;
            tl_counters = counters ? &local : nullptr;
#pragma omp for schedule(dynamic, 16)
            for (int64_t k = 0; k < (int64_t)(row_end - row_begin) * cw; k++) {
                uint32_t y = row_begin + (uint32_t)(k / cw), x = col_begin + (uint32_t)(k % cw);
                const int64_t i = (int64_t)y * w + x;
                V3 accumulation = from_one(0.0);  // src/renderer.rs:49
                for (uint32_t sy = 0; sy < ss; sy++)
                    for (uint32_t sx = 0; sx < ss; sx++) {
                        double ncx, ncy;
                        normalized_coord(x, y, w, h, ss, sx, sy, &ncx, &ncy);
                        accumulation = accumulation + calc_pixel(c, *camera, mode, ncx, ncy, sampling);
                    }
                accum[3 * i] += accumulation.x; accum[3 * i + 1] += accumulation.y; accum[3 * i + 2] += accumulation.z;
            }
            tl_counters = nullptr;
            if (counters) {
#pragma omp critical
                {
                    total.paths += local.paths; total.segments += local.segments; total.shadow_rays += local.shadow_rays;
                    total.node_visits += local.node_visits; total.tri_tests += local.tri_tests; total.elem_tests += local.elem_tests;
                    total.lens_iters += local.lens_iters;
                }
            }
        }
    }
    if (counters) {
        counters->paths = total.paths; counters->segments = total.segments; counters->shadow_rays = total.shadow_rays;
        counters->node_visits = total.node_visits; counters->tri_tests = total.tri_tests; counters->elem_tests = total.elem_tests;
        counters->lens_iters = total.lens_iters;
    }
}

// `update_imgbuf` (src/renderer.rs:64-90)
void oracle_resolve(const hnm_config* cfg, const double* accum, uint32_t width, uint32_t height, uint32_t sampling, uint8_t* rgb8) {
    double scale = 1.0 / (double)(uint32_t)(sampling * cfg->supersampling * cfg->supersampling);
    size_t n = (size_t)width * height;
    std::vector<V3> tmp(n);
    double inv_gamma = 1.0 / cfg->gamma_factor;
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        V3 hdr = v3(accum[3 * i], accum[3 * i + 1], accum[3 * i + 2]) * scale;
        V3 ldr = cfg->tone_mapping_mode == 1 ? reinhard(hdr, cfg->tone_exposure, cfg->tone_white_point) : hdr;
        tmp[i] = v3(om::pow(ldr.x, inv_gamma), om::pow(ldr.y, inv_gamma), om::pow(ldr.z, inv_gamma));  // src/color.rs:38-48
    }
    for (uint32_t it = 0; it < cfg->bilateral_iteration; it++) {
        std::vector<V3> next(n);
#pragma omp parallel for
        for (int64_t i = 0; i < (int64_t)n; i++) next[i] = bilateral(*cfg, tmp, (uint32_t)i, width, height);
        tmp.swap(next);
    }
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {  // src/color.rs:10-16
        rgb8[3 * i] = f64_as_u8(255.0 * saturate(tmp[i].x));
        rgb8[3 * i + 1] = f64_as_u8(255.0 * saturate(tmp[i].y));
        rgb8[3 * i + 2] = f64_as_u8(255.0 * saturate(tmp[i].z));
    }
}

// single functions for unit tests
void oracle_texture_sample(const hnm_scene_desc* d, int32_t image, double tint_r, double tint_g, double tint_b, const double* uv, uint32_t n,
                           double* out) {
    Ctx c = make_ctx(d);
    hnm_texture t;
    t.color = hnm_vec3{tint_r, tint_g, tint_b};
    t.image = image;
    for (uint32_t i = 0; i < n; i++) {
        V3 r = texture_sample(c, t, uv[2 * i], uv[2 * i + 1]);
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
void oracle_skybox_sample(const hnm_scene_desc* d, const double* dirs, uint32_t n, double* out) {
    Ctx c = make_ctx(d);
    for (uint32_t i = 0; i < n; i++) {
        V3 r = skybox_sample(c, v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
// the camera ray of sub-pixel (x, y, sx, sy) of pass `sampling`: out = origin xyz, direction xyz, lens iterations
void oracle_camera_ray(const hnm_scene_desc* d, const hnm_camera* camera, uint32_t w, uint32_t h, uint32_t x, uint32_t y, uint32_t sx,
                       uint32_t sy, uint32_t sampling, int dof, double* out) {
    double ncx, ncy;
    normalized_coord(x, y, w, h, d->config.supersampling, sx, sy, &ncx, &ncy);
    Ray r;
    if (dof) {
        uint64_t seed[4] = {8700304ull, (uint64_t)sampling, f64_as_u64((4.0 + ncx) * 100870.0), f64_as_u64((4.0 + ncy) * 100304.0)};
        Isaac64 rng;
        rng.from_seed(seed, 4);
        r = camera_ray_with_dof(*camera, ncx, ncy, rng);
        out[6] = (double)((256 - rng.cnt) / 2);
    } else {
        r = camera_ray(*camera, ncx, ncy);
        out[6] = 0.0;
    }
    out[0] = r.origin.x; out[1] = r.origin.y; out[2] = r.origin.z;
    out[3] = r.direction.x; out[4] = r.direction.y; out[5] = r.direction.z;
}

}  // extern "C"
