// hanamaru_host.cpp -- host-side mirror of the reference (see hanamaru_host.h).
// Scene authoring, OBJ parsing, BVH build and flattening.  Nothing here is on
// the hot path and nothing here touches CUDA.
#include "hanamaru_host.h"

#include <dlfcn.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace hanamaru {

hnm_config config::to_abi() {
    hnm_config c;
    memset(&c, 0, sizeof(c));
    c.eps = EPS; c.offset = OFFSET; c.inf = INF; c.gamma_factor = GAMMA_FACTOR;
    c.tone_exposure = TONE_MAPPING_EXPOSURE; c.tone_white_point = TONE_MAPPING_WHITE_POINT;
    c.bilateral_sigma_i = BILATERAL_FILTER_SIGMA_I; c.bilateral_sigma_s = BILATERAL_FILTER_SIGMA_S;
    c.supersampling = SUPERSAMPLING; c.bounce_limit = PATHTRACING_BOUNCE_LIMIT;
    c.tone_mapping_mode = TONE_MAPPING_MODE; c.bilateral_iteration = BILATERAL_FILTER_ITERATION;
    c.bilateral_diameter = BILATERAL_FILTER_DIAMETER;
    return c;
}

// ---- src/color.rs:51-61 --------------------------------------------------------
static double saturate(double v) { return std::fmin(std::fmax(v, 0.0), 1.0); }
static Color hue(double h) {
    return Color(saturate(std::fabs(h * 6.0 - 3.0) - 1.0), saturate(2.0 - std::fabs(h * 6.0 - 2.0)),
                 saturate(2.0 - std::fabs(h * 6.0 - 4.0)));
}
Color hsv_to_rgb(Color c) { return ((hue(c.x) - 1.0) * c.y + 1.0) * c.z; }

// ---- src/matrix.rs ---------------------------------------------------------------
Matrix44 Matrix44::identity() {
    Matrix44 m;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) m.e[i][j] = (i == j) ? 1.0 : 0.0;
    return m;
}
Matrix44 Matrix44::scale(double sx, double sy, double sz) {
    Matrix44 m = identity();
    m.e[0][0] = sx; m.e[1][1] = sy; m.e[2][2] = sz;
    return m;
}
Matrix44 Matrix44::rotate_x(double t) {
    double s = std::sin(t), c = std::cos(t);
    Matrix44 m = identity();
    m.e[1][1] = c; m.e[1][2] = -s; m.e[2][1] = s; m.e[2][2] = c;
    return m;
}
Matrix44 Matrix44::rotate_y(double t) {
    double s = std::sin(t), c = std::cos(t);
    Matrix44 m = identity();
    m.e[0][0] = c; m.e[0][2] = s; m.e[2][0] = -s; m.e[2][2] = c;
    return m;
}
Matrix44 Matrix44::rotate_z(double t) {
    double s = std::sin(t), c = std::cos(t);
    Matrix44 m = identity();
    m.e[0][0] = c; m.e[0][1] = -s; m.e[1][0] = s; m.e[1][1] = c;
    return m;
}
Matrix44 Matrix44::translate(double tx, double ty, double tz) {
    Matrix44 m = identity();
    m.e[0][3] = tx; m.e[1][3] = ty; m.e[2][3] = tz;
    return m;
}
Matrix44 Matrix44::operator*(const Matrix44& o) const {  // src/matrix.rs:163-176
    Matrix44 r = identity();
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            r.e[i][j] = e[i][0] * o.e[0][j] + e[i][1] * o.e[1][j] + e[i][2] * o.e[2][j] + e[i][3] * o.e[3][j];
    return r;
}
Vector3 Matrix44::operator*(const Vector3& v) const {  // src/matrix.rs:180-190
    return Vector3(v.x * e[0][0] + v.y * e[0][1] + v.z * e[0][2] + e[0][3],
                   v.x * e[1][0] + v.y * e[1][1] + v.z * e[1][2] + e[1][3],
                   v.x * e[2][0] + v.y * e[2][1] + v.z * e[2][2] + e[2][3]);
}

// ---- src/camera.rs:45-64 ------------------------------------------------------------
Camera::Camera(Vector3 eye_, Vector3 target, Vector3 y_up, double v_fov, LensShape shape, double aperture,
               double focus) {
    lens_radius = 0.5 * aperture;
    // f64::to_radians is `self * (PI / 180.0)`; note the FULL fov is used as the half angle
    double plane_half_height = std::tan(v_fov * (config::PI / 180.0));
    forward = (target - eye_).normalize();
    right = forward.cross(y_up).normalize();
    up = right.cross(forward).normalize();
    eye = eye_;
    lens_shape = shape;
    focus_distance = focus;
    plane_half_right = right * plane_half_height * focus;
    plane_half_up = up * plane_half_height * focus;
}
hnm_camera Camera::abi() const {
    hnm_camera c;
    memset(&c, 0, sizeof(c));
    c.eye = eye.abi(); c.right = right.abi(); c.up = up.abi(); c.forward = forward.abi();
    c.plane_half_right = plane_half_right.abi(); c.plane_half_up = plane_half_up.abi();
    c.lens_radius = lens_radius; c.focus_distance = focus_distance;
    c.lens_shape = (int32_t)lens_shape;
    return c;
}

// ---- src/loader.rs:12-59 ---------------------------------------------------------------
static std::vector<std::string> split_char(const std::string& s, char sep) {  // str::split(" ")
    std::vector<std::string> out;
    size_t start = 0;
    for (;;) {
        size_t p = s.find(sep, start);
        if (p == std::string::npos) { out.push_back(s.substr(start)); break; }
        out.push_back(s.substr(start, p - start));
        start = p + 1;
    }
    return out;
}
static double parse_f64(const std::string& s) {
    // Rust `parse::<f64>()`: whole string must be a number, correctly rounded (as is strtod)
    if (s.empty()) throw std::runtime_error("obj: empty float field (the reference would panic here)");
    char* end = nullptr;
    double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end != '\0') throw std::runtime_error("obj: bad float '" + s + "'");
    return v;
}
static size_t parse_usize(const std::string& s) {
    if (s.empty()) throw std::runtime_error("obj: empty index field (the reference would panic here)");
    char* end = nullptr;
    unsigned long long v = std::strtoull(s.c_str(), &end, 10);
    if (end == s.c_str() || *end != '\0' || s[0] == '-' || s[0] == ' ') throw std::runtime_error("obj: bad index '" + s + "'");
    return (size_t)v;
}
ObjGeometry parse_obj(const std::string& text) {
    ObjGeometry g;
    size_t pos = 0;
    while (pos < text.size()) {
        size_t nl = text.find('\n', pos);
        std::string l = text.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
        pos = (nl == std::string::npos) ? text.size() : nl + 1;
        if (!l.empty() && l.back() == '\r') l.pop_back();  // BufRead::lines strips "\r\n"
        std::vector<std::string> sp = split_char(l, ' ');
        if (sp[0] == "v") {
            if (sp.size() < 4) throw std::runtime_error("obj: short v line");
            g.vertexes.push_back(Vector3(parse_f64(sp[1]), parse_f64(sp[2]), parse_f64(sp[3])));
        } else if (sp[0] == "f") {
            if (sp.size() < 4) throw std::runtime_error("obj: short f line");
            size_t a = parse_usize(split_char(sp[1], '/')[0]) - 1;
            size_t b = parse_usize(split_char(sp[2], '/')[0]) - 1;
            size_t c = parse_usize(split_char(sp[3], '/')[0]) - 1;
            g.faces.push_back((uint32_t)a); g.faces.push_back((uint32_t)b); g.faces.push_back((uint32_t)c);
            if (sp.size() == 5) {  // quad -> (a, c, d)
                size_t d = parse_usize(split_char(sp[4], '/')[0]) - 1;
                g.faces.push_back((uint32_t)a); g.faces.push_back((uint32_t)c); g.faces.push_back((uint32_t)d);
            }
        }
    }
    return g;
}

Mesh ObjLoader::load(const AssetStore& assets, const std::string& path, const Matrix44& matrix, Material material) {
    std::shared_ptr<ObjGeometry> g = assets.obj(path);
    Mesh mesh;
    mesh.material = std::move(material);
    mesh.vertexes.reserve(g->vertexes.size());
    for (const Vector3& v : g->vertexes) mesh.vertexes.push_back(matrix * v);
    for (size_t i = 0; i + 2 < g->faces.size(); i += 3) mesh.faces.push_back(Face{g->faces[i], g->faces[i + 1], g->faces[i + 2]});
    return mesh;
}

// ---- asset store ----------------------------------------------------------------------------
static bool read_all(const std::string& path, std::string& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

std::shared_ptr<ObjGeometry> AssetStore::obj(const std::string& path) const {
    auto it = objs_.find(path);
    if (it != objs_.end()) return it->second;
    if (!root_.empty()) {
        std::string text;
        if (read_all(root_ + "/" + path, text)) {
            auto g = std::make_shared<ObjGeometry>(parse_obj(text));
            objs_[path] = g;
            return g;
        }
    }
    throw std::runtime_error("asset not found: " + path);
}
std::shared_ptr<Image> AssetStore::image(const std::string& path) const {
    auto it = images_.find(path);
    if (it != images_.end()) return it->second;
    auto en = encoded_.find(path);
    if (en != encoded_.end()) {
        auto img = std::make_shared<Image>();
        std::string err;
        if (!image_decode(en->second->data(), en->second->size(), *img, &err)) throw std::runtime_error(path + ": " + err);
        images_[path] = img;
        return img;
    }
    if (!root_.empty()) {
        std::string bytes;
        if (read_all(root_ + "/" + path, bytes)) {  // `image::open` (src/texture.rs:18)
            auto img = std::make_shared<Image>();
            std::string err;
            if (!image_decode((const uint8_t*)bytes.data(), bytes.size(), *img, &err)) throw std::runtime_error(path + ": " + err);
            images_[path] = img;
            return img;
        }
    }
    throw std::runtime_error("image asset not found (set a root directory, register it, or load a pack): " + path);
}

// pack layout (little endian), written by tools/make_asset_pack.py:
//   "HNMPACK1" u32 count { u32 name_len, name, u32 kind(1 mesh | 2 image), u32 a, u32 b, u64 raw, u64 comp, zlib bytes }
//   mesh : a = vertex count, b = face count; payload = f64 xyz[a] then u32 v0v1v2[b]
//   image: a = width, b = height;           payload = RGBA8 rows, top row first
//   kind 3 = image file bytes (PNG / JPEG) as shipped by the reference, decoded on first use (a = b = 0)
bool AssetStore::load_pack(const std::string& path, std::string* err) {
    std::string buf;
    if (!read_all(path, buf)) { if (err) *err = "cannot read " + path; return false; }
    const uint8_t* p = (const uint8_t*)buf.data();
    const uint8_t* end = p + buf.size();
    auto need = [&](size_t n) { if ((size_t)(end - p) < n) throw std::runtime_error("truncated pack"); };
    auto rd32 = [&]() { need(4); uint32_t v; memcpy(&v, p, 4); p += 4; return v; };
    auto rd64 = [&]() { need(8); uint64_t v; memcpy(&v, p, 8); p += 8; return v; };
    try {
        need(8);
        if (memcmp(p, "HNMPACK1", 8) != 0) throw std::runtime_error("bad pack magic");
        p += 8;
        uint32_t count = rd32();
        for (uint32_t i = 0; i < count; i++) {
            uint32_t nl = rd32();
            need(nl);
            std::string name((const char*)p, nl);
            p += nl;
            uint32_t kind = rd32(), a = rd32(), b = rd32();
            uint64_t raw = rd64(), comp = rd64();
            need(comp);
            std::vector<uint8_t> data(raw);
            uLongf dl = (uLongf)raw;
            if (uncompress(data.data(), &dl, p, (uLong)comp) != Z_OK || dl != raw) throw std::runtime_error("zlib failure in " + name);
            p += comp;
            if (kind == 1) {
                if (raw != (uint64_t)a * 24 + (uint64_t)b * 12) throw std::runtime_error("bad mesh size " + name);
                auto g = std::make_shared<ObjGeometry>();
                g->vertexes.resize(a);
                for (uint32_t k = 0; k < a; k++) {
                    double xyz[3];
                    memcpy(xyz, data.data() + (size_t)k * 24, 24);
                    g->vertexes[k] = Vector3(xyz[0], xyz[1], xyz[2]);
                }
                g->faces.resize((size_t)b * 3);
                memcpy(g->faces.data(), data.data() + (size_t)a * 24, (size_t)b * 12);
                objs_[name] = g;
            } else if (kind == 2) {
                if (raw != (uint64_t)a * b * 4) throw std::runtime_error("bad image size " + name);
                auto img = std::make_shared<Image>();
                img->width = a; img->height = b;
                img->rgba = std::move(data);
                images_[name] = img;
            } else if (kind == 3) {
                // an image FILE as the reference ships it (PNG / JPEG bytes): decoded by image_decode on first use
                encoded_[name] = std::make_shared<std::vector<uint8_t>>(std::move(data));
            } else {
                throw std::runtime_error("unknown pack entry kind");
            }
        }
    } catch (const std::exception& e) {
        if (err) *err = e.what();
        return false;
    }
    return true;
}

// ---- src/bvh.rs -------------------------------------------------------------------------------
bool Aabb::intersect_aabb(const Aabb& o) const {
    return min.x < o.max.x && max.x > o.min.x && min.y < o.max.y && max.y > o.min.y && min.z < o.max.z && max.z > o.min.z;
}
void Aabb::merge(const Aabb& o) {  // f64::min / f64::max
    min.x = std::fmin(min.x, o.min.x); min.y = std::fmin(min.y, o.min.y); min.z = std::fmin(min.z, o.min.z);
    max.x = std::fmax(max.x, o.max.x); max.y = std::fmax(max.y, o.max.y); max.z = std::fmax(max.z, o.max.z);
}
static Aabb aabb_from_triangle(const Vector3& v0, const Vector3& v1, const Vector3& v2) {
    Aabb a;
    a.min = Vector3(std::fmin(std::fmin(v0.x, v1.x), v2.x), std::fmin(std::fmin(v0.y, v1.y), v2.y), std::fmin(std::fmin(v0.z, v1.z), v2.z));
    a.max = Vector3(std::fmax(std::fmax(v0.x, v1.x), v2.x), std::fmax(std::fmax(v0.y, v1.y), v2.y), std::fmax(std::fmax(v0.z, v1.z), v2.z));
    return a;
}
static std::unique_ptr<BvhNode> empty_node() {
    auto n = std::make_unique<BvhNode>();
    n->aabb.min = Vector3::from_one(config::INF);
    n->aabb.max = Vector3::from_one(-config::INF);
    return n;
}
// 0 = x, 1 = y, 2 = z (src/bvh.rs:125,134,143: strict comparisons, ties fall through to z)
static int split_axis(const Aabb& a) {
    double lx = a.max.x - a.min.x, ly = a.max.y - a.min.y, lz = a.max.z - a.min.z;
    if (lx > ly && lx > lz) return 0;
    if (ly > lx && ly > lz) return 1;
    return 2;
}
static double comp(const Vector3& v, int axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

// src/bvh.rs:107-153
static std::unique_ptr<BvhNode> build_mesh_rec(const Mesh& mesh, std::vector<size_t>& idx) {
    auto node = empty_node();
    for (size_t fi : idx) {
        const Face& f = mesh.faces[fi];
        node->aabb.merge(aabb_from_triangle(mesh.vertexes[f.v0], mesh.vertexes[f.v1], mesh.vertexes[f.v2]));
    }
    size_t mid = idx.size() / 2;
    if (mid <= 2) {
        node->indexes = idx;
    } else {
        int axis = split_axis(node->aabb);
        // slice::sort_by is a stable sort; partial_cmp(..).unwrap() would panic on NaN
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) {
            const Face& fa = mesh.faces[a];
            const Face& fb = mesh.faces[b];
            double sa = comp(mesh.vertexes[fa.v0], axis) + comp(mesh.vertexes[fa.v1], axis) + comp(mesh.vertexes[fa.v2], axis);
            double sb = comp(mesh.vertexes[fb.v0], axis) + comp(mesh.vertexes[fb.v1], axis) + comp(mesh.vertexes[fb.v2], axis);
            return sa < sb;
        });
        std::vector<size_t> second(idx.begin() + mid, idx.end());  // split_off(mid)
        idx.resize(mid);
        node->children.push_back(build_mesh_rec(mesh, idx));
        node->children.push_back(build_mesh_rec(mesh, second));
    }
    return node;
}
std::unique_ptr<BvhNode> build_from_mesh(const Mesh& mesh) {
    std::vector<size_t> idx(mesh.faces.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = i;
    return build_mesh_rec(mesh, idx);
}
// src/bvh.rs:155-201
static std::unique_ptr<BvhNode> build_scene_rec(const Scene& scene, std::vector<size_t>& idx) {
    auto node = empty_node();
    for (size_t i : idx) node->aabb.merge(scene.elements[i]->aabb());
    size_t mid = idx.size() / 2;
    if (mid <= 2) {
        node->indexes = idx;
    } else {
        int axis = split_axis(node->aabb);
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) {
            Aabb aa = scene.elements[a]->aabb(), ab = scene.elements[b]->aabb();
            double sa = comp(aa.min, axis) + comp(aa.max, axis);
            double sb = comp(ab.min, axis) + comp(ab.max, axis);
            return sa < sb;
        });
        std::vector<size_t> second(idx.begin() + mid, idx.end());
        idx.resize(mid);
        node->children.push_back(build_scene_rec(scene, idx));
        node->children.push_back(build_scene_rec(scene, second));
    }
    return node;
}
std::unique_ptr<BvhNode> build_from_scene(const Scene& scene) {
    std::vector<size_t> idx(scene.elements.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = i;
    return build_scene_rec(scene, idx);
}

// ---- src/scene.rs ----------------------------------------------------------------------------------
Aabb Sphere::aabb() const { return Aabb{center - Vector3::from_one(radius), center + Vector3::from_one(radius)}; }
std::unique_ptr<BvhMesh> BvhMesh::from_mesh(Mesh mesh) {
    auto m = std::make_unique<BvhMesh>();
    m->bvh = build_from_mesh(mesh);
    m->mesh = std::move(mesh);
    return m;
}
bool Scene::add_with_check_collisions(std::unique_ptr<Intersectable> e) {
    Aabb a = e->aabb();
    for (const auto& o : elements)
        if (o->aabb().intersect_aabb(a)) return false;
    elements.push_back(std::move(e));
    return true;
}
std::vector<uint32_t> Scene::emissions() const {
    std::vector<uint32_t> out;
    for (size_t i = 0; i < elements.size(); i++)
        if (elements[i]->nee_available() && elements[i]->material().emission.color != Color::zero()) out.push_back((uint32_t)i);
    return out;
}

// ---- flattening ---------------------------------------------------------------------------------------
int32_t FlatSceneBuilder::add_image(const std::shared_ptr<Image>& img) {
    if (!img) return -1;
    auto it = image_ids_.find(img.get());
    if (it != image_ids_.end()) return it->second;
    int32_t id = (int32_t)out_.images.size();
    out_.images.push_back(hnm_image{img->rgba.data(), img->width, img->height});
    out_.image_refs.push_back(img);
    image_ids_[img.get()] = id;
    return id;
}
int32_t FlatSceneBuilder::add_material(const Material& m) {
    hnm_material hm;
    memset(&hm, 0, sizeof(hm));
    auto tex = [&](const Texture& t) {
        hnm_texture ht;
        memset(&ht, 0, sizeof(ht));
        ht.color = t.color.abi();
        ht.image = add_image(t.image_texture);
        return ht;
    };
    hm.albedo = tex(m.albedo); hm.emission = tex(m.emission); hm.roughness = tex(m.roughness);
    hm.param = m.surface.param; hm.surface = m.surface.tag;
    out_.materials.push_back(hm);
    return (int32_t)out_.materials.size() - 1;
}
void FlatSceneBuilder::flatten_tree(const BvhNode& root, std::vector<hnm_bvh_node>& nodes, std::vector<uint32_t>& indices,
                                    uint32_t index_base) {
    // DFS pre-order == the reference's recursion order (src/bvh.rs:213-263)
    struct Rec {
        static uint32_t go(const BvhNode& n, std::vector<hnm_bvh_node>& nodes, std::vector<uint32_t>& indices, uint32_t index_base) {
            uint32_t me = (uint32_t)nodes.size();
            nodes.push_back(hnm_bvh_node{});
            hnm_bvh_node h;
            memset(&h, 0, sizeof(h));
            h.aabb_min[0] = n.aabb.min.x; h.aabb_min[1] = n.aabb.min.y; h.aabb_min[2] = n.aabb.min.z;
            h.aabb_max[0] = n.aabb.max.x; h.aabb_max[1] = n.aabb.max.y; h.aabb_max[2] = n.aabb.max.z;
            if (n.children.empty()) {
                h.child0 = h.child1 = -1;
                h.first = (uint32_t)indices.size() - index_base;
                h.count = (uint32_t)n.indexes.size();
                for (size_t i : n.indexes) indices.push_back((uint32_t)i);
            } else {
                h.child0 = (int32_t)go(*n.children[0], nodes, indices, index_base);
                h.child1 = (int32_t)go(*n.children[1], nodes, indices, index_base);
            }
            nodes[me] = h;
            return me;
        }
    };
    Rec::go(root, nodes, indices, index_base);
}
void FlatSceneBuilder::add_sphere(const Sphere& s) {
    hnm_element e;
    memset(&e, 0, sizeof(e));
    e.kind = HNM_ELEM_SPHERE; e.a = s.center.abi(); e.radius = s.radius; e.mesh = -1;
    e.material = add_material(s.mat);
    out_.elements.push_back(e);
}
void FlatSceneBuilder::add_cuboid(const Cuboid& c) {
    hnm_element e;
    memset(&e, 0, sizeof(e));
    e.kind = HNM_ELEM_CUBOID; e.a = c.box.min.abi(); e.b = c.box.max.abi(); e.mesh = -1;
    e.material = add_material(c.mat);
    out_.elements.push_back(e);
}
void FlatSceneBuilder::add_mesh(const BvhMesh& m) {
    hnm_mesh hm;
    memset(&hm, 0, sizeof(hm));
    hm.vertex_offset = (uint32_t)(out_.vertices.size() / 3);
    hm.vertex_count = (uint32_t)m.mesh.vertexes.size();
    for (const Vector3& v : m.mesh.vertexes) { out_.vertices.push_back(v.x); out_.vertices.push_back(v.y); out_.vertices.push_back(v.z); }
    hm.face_offset = (uint32_t)(out_.faces.size() / 3);
    hm.face_count = (uint32_t)m.mesh.faces.size();
    for (const Face& f : m.mesh.faces) { out_.faces.push_back((uint32_t)f.v0); out_.faces.push_back((uint32_t)f.v1); out_.faces.push_back((uint32_t)f.v2); }
    hm.node_offset = (uint32_t)out_.mesh_nodes.size();
    hm.index_offset = (uint32_t)out_.mesh_indices.size();
    std::vector<hnm_bvh_node> nodes;
    flatten_tree(*m.bvh, nodes, out_.mesh_indices, hm.index_offset);
    // child links are relative to this mesh's node_offset
    out_.mesh_nodes.insert(out_.mesh_nodes.end(), nodes.begin(), nodes.end());
    hm.node_count = (uint32_t)nodes.size();
    hm.index_count = (uint32_t)out_.mesh_indices.size() - hm.index_offset;
    hnm_element e;
    memset(&e, 0, sizeof(e));
    e.kind = HNM_ELEM_MESH; e.mesh = (int32_t)out_.meshes.size();
    e.a = m.bvh->aabb.min.abi(); e.b = m.bvh->aabb.max.abi();
    e.material = add_material(m.mesh.material);
    out_.meshes.push_back(hm);
    out_.elements.push_back(e);
}
void Sphere::flatten(FlatSceneBuilder& b) const { b.add_sphere(*this); }
void Cuboid::flatten(FlatSceneBuilder& b) const { b.add_cuboid(*this); }
void BvhMesh::flatten(FlatSceneBuilder& b) const { b.add_mesh(*this); }

void FlatScene::finalize() {
    memset(&desc, 0, sizeof(desc));
    desc.abi_version = HNM_ABI_VERSION;
    desc.elements = elements.data(); desc.num_elements = (uint32_t)elements.size();
    desc.materials = materials.data(); desc.num_materials = (uint32_t)materials.size();
    desc.images = images.data(); desc.num_images = (uint32_t)images.size();
    desc.meshes = meshes.data(); desc.num_meshes = (uint32_t)meshes.size();
    desc.vertices = vertices.data(); desc.num_vertices = (uint32_t)(vertices.size() / 3);
    desc.faces = faces.data(); desc.num_faces = (uint32_t)(faces.size() / 3);
    desc.mesh_nodes = mesh_nodes.data(); desc.num_mesh_nodes = (uint32_t)mesh_nodes.size();
    desc.mesh_indices = mesh_indices.data(); desc.num_mesh_indices = (uint32_t)mesh_indices.size();
    desc.top_nodes = top_nodes.data(); desc.num_top_nodes = (uint32_t)top_nodes.size();
    desc.top_indices = top_indices.data(); desc.num_top_indices = (uint32_t)top_indices.size();
    desc.emissions = emissions.data(); desc.num_emissions = (uint32_t)emissions.size();
    desc.config = config::to_abi();
}

std::unique_ptr<BvhScene> BvhScene::from_scene(Scene scene) {
    auto bs = std::make_unique<BvhScene>();
    bs->bvh = build_from_scene(scene);
    bs->scene = std::move(scene);
    FlatSceneBuilder b(bs->flat);
    for (const auto& e : bs->scene.elements) e->flatten(b);
    FlatSceneBuilder::flatten_tree(*bs->bvh, bs->flat.top_nodes, bs->flat.top_indices, 0);
    bs->flat.emissions = bs->scene.emissions();
    const Skybox& sb = bs->scene.skybox;
    bs->flat.finalize();
    const std::shared_ptr<Image> faces[6] = {sb.px, sb.nx, sb.py, sb.ny, sb.pz, sb.nz};
    for (int i = 0; i < 6; i++) bs->flat.desc.skybox_images[i] = b.add_image(faces[i]);
    bs->flat.desc.skybox_intensity = sb.intensity.abi();
    // add_image may have grown the image table
    bs->flat.desc.images = bs->flat.images.data();
    bs->flat.desc.num_images = (uint32_t)bs->flat.images.size();
    return bs;
}

// ---- rand 0.4.3 src/prng/isaac64.rs + src/lib.rs (Rng::next_f64, gen_range) ------------------------------------
namespace {
inline void isaac_mix(uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d, uint64_t& e, uint64_t& f, uint64_t& g, uint64_t& h) {
    a -= e; f ^= h >> 9;  h += a;
    b -= f; g ^= a << 9;  a += b;
    c -= g; h ^= b >> 23; b += c;
    d -= h; a ^= c << 15; c += d;
    e -= a; b ^= d >> 14; d += e;
    f -= b; c ^= e << 20; e += f;
    g -= c; d ^= f >> 17; f += g;
    h -= d; e ^= g << 14; g += h;
}
}  // namespace
StdRng::StdRng(const std::vector<uint64_t>& seed) {
    // from_seed: the seed words, zero padded, become rsl; a = b = c = 0; init(true)
    for (int i = 0; i < 256; i++) rsl_[i] = i < (int)seed.size() ? seed[i] : 0;
    uint64_t a, b, c, d, e, f, g, h;
    a = b = c = d = e = f = g = h = 0x9e3779b97f4a7c13ull;
    for (int i = 0; i < 4; i++) isaac_mix(a, b, c, d, e, f, g, h);
    for (int pass = 0; pass < 2; pass++) {
        const uint64_t* src = pass == 0 ? rsl_ : mem_;
        for (int i = 0; i < 256; i += 8) {
            a += src[i]; b += src[i + 1]; c += src[i + 2]; d += src[i + 3];
            e += src[i + 4]; f += src[i + 5]; g += src[i + 6]; h += src[i + 7];
            isaac_mix(a, b, c, d, e, f, g, h);
            mem_[i] = a; mem_[i + 1] = b; mem_[i + 2] = c; mem_[i + 3] = d;
            mem_[i + 4] = e; mem_[i + 5] = f; mem_[i + 6] = g; mem_[i + 7] = h;
        }
    }
    isaac64();
}
void StdRng::isaac64() {
    c_ += 1;
    uint64_t a = a_, b = b_ + c_;
    for (int i = 0; i < 256; i++) {
        uint64_t mixed;
        switch (i & 3) {
            case 0: mixed = ~(a ^ (a << 21)); break;
            case 1: mixed = a ^ (a >> 5); break;
            case 2: mixed = a ^ (a << 12); break;
            default: mixed = a ^ (a >> 33); break;
        }
        uint64_t x = mem_[i];
        a = mixed + mem_[(i + 128) & 255];
        uint64_t y = mem_[(x >> 3) & 255] + a + b;
        mem_[i] = y;
        b = mem_[(y >> 11) & 255] + x;
        rsl_[i] = b;
    }
    a_ = a; b_ = b; cnt_ = 256;
}
uint64_t StdRng::next_u64() {
    if (cnt_ == 0) isaac64();
    cnt_ -= 1;
    return rsl_[cnt_ & 255];
}
double StdRng::next_f64() {
    uint64_t bits = 0x3FF0000000000000ull | (next_u64() & 0xFFFFFFFFFFFFFull);
    double v;
    memcpy(&v, &bits, sizeof(v));
    return v - 1.0;
}
double StdRng::gen_range(double low, double high) { return low + (high - low) * next_f64(); }

// ---- scene authoring (src/main.rs) ---------------------------------------------------------------------
static Skybox make_skybox(const AssetStore& a, const std::string& dir, Vector3 intensity) {
    Skybox s;
    s.px = a.image(dir + "/posx.jpg"); s.nx = a.image(dir + "/negx.jpg");
    s.py = a.image(dir + "/posy.jpg"); s.ny = a.image(dir + "/negy.jpg");
    s.pz = a.image(dir + "/posz.jpg"); s.nz = a.image(dir + "/negz.jpg");
    s.intensity = intensity;
    return s;
}
static Texture tex_path(const AssetStore& a, const std::string& p) { return Texture::from_image(a.image(p)); }
static double fract(double v) { return v - std::trunc(v); }

// camera + light + mirror + frame + floor of the submitted scene; shared by the
// default scene and the two builder-defined benchmark scenes
static SceneAndCamera rtcamp6_stage(const AssetStore& a, double aperture, double focus, bool with_bunny_and_mirror) {
    const double scene_scale = 1.0;
    double theta = config::PI2 * 0.03;
    double r = 6.5 * scene_scale;
    SceneAndCamera sc;
    sc.camera = Camera(Vector3(r * std::sin(theta), 2.0 * scene_scale, r * std::cos(theta)), Vector3(0.0, 1.0 * scene_scale, 0.0),
                       Vector3(0.0, 1.0, 0.0).normalize(), 20.0, LensShape::Circle, aperture, focus * scene_scale);
    double radius = 0.2;
    double floor_s = 9.0 * scene_scale;
    Scene& scene = sc.scene;
    scene.add(std::make_unique<Sphere>(Vector3(-0.3, 0.5 + radius, 0.0) * scene_scale, radius * scene_scale,
                                       Material{SurfaceType::Diffuse(), Texture::black(), Texture::from_color(Color(30.0, 20.0, 4.0)), Texture::black()}));
    if (with_bunny_and_mirror) {
        scene.add(BvhMesh::from_mesh(ObjLoader::load(
            a, "models/bunny/bunny_wired_300.obj",
            Matrix44::scale_linear(1.5 * scene_scale) * Matrix44::translate(0.0, 0.0, 0.0) * Matrix44::rotate_y(0.3),
            Material{SurfaceType::GGX(0.8), Texture::from_color(Color(1.0, 0.01, 0.01)), Texture::black(), Texture::from_color(Color::from_one(0.05))})));
        scene.add(BvhMesh::from_mesh(ObjLoader::load(
            a, "models/box.obj",
            Matrix44::translate(1.0 * scene_scale, 0.0, -3.0 * scene_scale) * Matrix44::rotate_y(-config::PI / 8.0) *
                Matrix44::scale(4.0 * 0.9 * scene_scale, 3.0 * 0.9 * scene_scale, 0.1 * 0.9 * scene_scale),
            Material{SurfaceType::Specular(), Texture::white(), Texture::black(), Texture::black()})));
        scene.add(BvhMesh::from_mesh(ObjLoader::load(
            a, "models/picture_frame.obj",
            Matrix44::translate(1.0 * scene_scale, 0.0, -3.0 * scene_scale) * Matrix44::rotate_y(-config::PI / 8.0) *
                Matrix44::scale(4.0 * scene_scale, 3.0 * scene_scale, scene_scale),
            Material{SurfaceType::GGX(0.9), Texture::from_color(Color(0.33, 0.27, 0.22)), Texture::black(), Texture::from_color(Color::from_one(0.3))})));
    }
    scene.add(std::make_unique<Cuboid>(Aabb{Vector3(-floor_s, -1.0, -floor_s), Vector3(floor_s, 0.0, floor_s)},
                                       Material{SurfaceType::Diffuse(), tex_path(a, "textures/2d/magic-circle3.png"), Texture::black(), Texture::white()}));
    scene.skybox = make_skybox(a, "textures/cube/Powerlines", Vector3::from_one(1.0));
    return sc;
}

// src/main.rs:1020-1153
SceneAndCamera init_scene_rtcamp6_v3_1(const AssetStore& a) {
    const double scene_scale = 1.0;
    SceneAndCamera sc = rtcamp6_stage(a, 0.03, 5.0, true);
    int count = 6;
    for (int i = 0; i < count; i++) {
        double r = 2.2 * scene_scale;
        double dr = (double)i / (double)count;
        double theta = config::PI2 * dr;
        double px = r * std::sin(theta), py = 0.0, pz = r * std::cos(theta);
        double s = scene_scale;
        double offset = 0.45;
        Material m = (i % 2 == 0)
                         ? Material{SurfaceType::Refraction(1.5), Texture::from_color(hsv_to_rgb(Color(fract(offset + dr), 0.2, 1.0))),
                                    Texture::black(), Texture::from_color(Color::from_one(0.1))}
                         : Material{SurfaceType::GGX(0.8), Texture::from_color(hsv_to_rgb(Color(fract(offset + dr), 1.0, 1.0))),
                                    Texture::black(), Texture::from_color(Color::from_one(0.05 * (double)i))};
        sc.scene.add(BvhMesh::from_mesh(ObjLoader::load(
            a, "models/armadilo_1000.obj", Matrix44::translate(px, py, pz) * Matrix44::rotate_y(theta) * Matrix44::scale_linear(s), std::move(m))));
    }
    return sc;
}

// src/main.rs:1155-1212
SceneAndCamera init_scene_rtcamp6_v4(const AssetStore& a) {
    SceneAndCamera sc;
    sc.camera = Camera(Vector3(0.0, 1.0, 6.0), Vector3(0.0, 0.0, 0.0), Vector3(0.0, 1.0, 0.0).normalize(), 30.0, LensShape::Circle,
                       0.2 * 0.0, 4.9);
    sc.scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/fractal_icosahedron.obj", Matrix44::scale_linear(1.0) * Matrix44::translate(0.0, 0.0, 0.0) * Matrix44::rotate_y(0.3),
        Material{SurfaceType::GGX(0.8), Texture::from_color(Color(1.0, 1.0, 1.0)), Texture::black(), Texture::from_color(Color::from_one(0.05))})));
    sc.scene.add(std::make_unique<Sphere>(sc.camera.eye - sc.camera.forward, 0.001,
                                          Material{SurfaceType::Diffuse(), Texture::black(), Texture::from_color(Color::from_one(1000.0)), Texture::black()}));
    sc.scene.skybox = make_skybox(a, "textures/cube/Ryfjallet", Vector3::from_one(1.0));
    return sc;
}

// src/main.rs:54-131 (init_scene_simple) and :133-250 (init_scene_material_examples)
static SceneAndCamera simple_stage(const AssetStore& a, double aperture, SurfaceType floor_surface) {
    SceneAndCamera sc;
    sc.camera = Camera(Vector3(0.0, 2.0, 9.0), Vector3(0.0, 1.0, 0.0), Vector3(0.0, 1.0, 0.0).normalize(), 10.0, LensShape::Circle, aperture, 8.8);
    sc.scene.add(std::make_unique<Cuboid>(
        Aabb{Vector3(-5.0, -1.0, -5.0), Vector3(5.0, 0.0, 5.0)},
        Material{floor_surface, tex_path(a, "textures/2d/checkered_diagonal_10_0.5_1.0_512.png"), Texture::black(),
                 tex_path(a, "textures/2d/checkered_diagonal_10_0.1_0.6_512.png")}));
    return sc;
}
static SceneAndCamera scene_simple(const AssetStore& a, const std::string& sky) {
    double radius = 0.6;
    SceneAndCamera st = simple_stage(a, 0.2 * 0.0, SurfaceType::GGX(0.8));
    SceneAndCamera sc;
    sc.camera = st.camera;
    sc.scene.add(std::make_unique<Sphere>(Vector3(0.0, radius, 0.0), radius,
                                          Material{SurfaceType::Diffuse(), Texture::white(), Texture::black(), Texture::from_color(Color::from_one(0.99))}));
    sc.scene.add(std::make_unique<Sphere>(Vector3(3.0, 2.0 + radius, -2.0), radius * 0.2,
                                          Material{SurfaceType::Diffuse(), Texture::black(), Texture::from_color(Color(200.0, 10.0, 10.0)), Texture::from_color(Color::from_one(0.05))}));
    sc.scene.add(std::make_unique<Sphere>(Vector3(-3.0, 2.0 + radius, -2.0), radius * 0.2,
                                          Material{SurfaceType::Diffuse(), Texture::black(), Texture::from_color(Color(10.0, 200.0, 10.0)), Texture::from_color(Color::from_one(0.05))}));
    sc.scene.add(std::move(st.scene.elements[0]));
    sc.scene.skybox = make_skybox(a, sky, Vector3::zero());
    return sc;
}
SceneAndCamera init_scene_simple(const AssetStore& a) { return scene_simple(a, "textures/cube/LancellottiChapel"); }
static SceneAndCamera scene_material_examples(const AssetStore& a, const std::string& sky) {
    double radius = 0.4;
    SceneAndCamera st = simple_stage(a, 0.2, SurfaceType::Diffuse());
    SceneAndCamera sc;
    sc.camera = st.camera;
    Texture rough = Texture::from_color(Color::from_one(0.05));
    const SurfaceType kinds[5] = {SurfaceType::Diffuse(), SurfaceType::GGX(0.8), SurfaceType::Specular(), SurfaceType::Refraction(1.5),
                                  SurfaceType::GGXRefraction(1.5)};
    for (int i = 0; i < 5; i++)
        sc.scene.add(std::make_unique<Sphere>(Vector3(-2.0 + (double)i, radius, 0.0), radius, Material{kinds[i], Texture::white(), Texture::black(), rough}));
    sc.scene.add(std::make_unique<Sphere>(Vector3(0.0, 2.0 + radius, -2.0), radius,
                                          Material{SurfaceType::Diffuse(), Texture::black(), Texture::from_color(Color::from_one(20.0)), rough}));
    sc.scene.add(std::move(st.scene.elements[0]));
    sc.scene.skybox = make_skybox(a, sky, Vector3::one());
    return sc;
}
SceneAndCamera init_scene_material_examples(const AssetStore& a) { return scene_material_examples(a, "textures/cube/LancellottiChapel"); }


// src/main.rs:252-499.  `sky` lets the tests swap the 2048^2 LancellottiChapel cubemap for the one in the asset pack.
static const char* MARBLE_DIFFUSE = "textures/2d/MarbleFloorTiles2/TexturesCom_MarbleFloorTiles2_1024_c_diffuse.tiff";
static const char* MARBLE_ROUGHNESS = "textures/2d/MarbleFloorTiles2/TexturesCom_MarbleFloorTiles2_1024_roughness.png";
static const char* EARTH = "textures/2d/earth_inverse_2048.jpg";
static double to_radians(double deg) { return deg * (config::PI / 180.0); }
static Material diamond_material() { return Material{SurfaceType::Refraction(2.42), Texture::white(), Texture::black(), Texture::black()}; }
static SceneAndCamera scene_rtcamp5(const AssetStore& a, const std::string& sky) {
    StdRng rng({870, 2000, 304, 2});
    SceneAndCamera sc;
    sc.camera = Camera(Vector3(0.0, 2.5, 9.0), Vector3(0.0, 1.0, 0.0), Vector3(0.0, 1.0, 0.0).normalize(), 17.0, LensShape::Circle, 0.15, 8.5);
    Scene& scene = sc.scene;
    scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/bunny/bunny_face1000.obj", Matrix44::scale_linear(1.5) * Matrix44::translate(1.2, 0.0, 0.0) * Matrix44::rotate_y(0.2),
        Material{SurfaceType::Refraction(1.5), Texture::from_color(Color(0.7, 0.7, 1.0)), Texture::black(), Texture::from_color(Color::from_one(0.1))})));
    scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/bunny/bunny_face1000_flip.obj", Matrix44::scale(1.5, 1.5, 1.5) * Matrix44::translate(-1.2, 0.0, 0.0) * Matrix44::rotate_y(-0.2),
        Material{SurfaceType::GGX(0.8), Texture::from_color(Color(1.0, 0.04, 0.04)), Texture::black(), Texture::from_color(Color::from_one(0.1))})));
    scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/dia/dia.obj",
        Matrix44::translate(3.1, 0.0, 0.8) * Matrix44::scale_linear(1.0) * Matrix44::rotate_y(-0.5) * Matrix44::rotate_x(to_radians(40.35)), diamond_material())));
    scene.add(std::make_unique<Sphere>(Vector3(0.0, 0.5, -0.5), 0.5,
                                       Material{SurfaceType::GGX(0.8), Texture::white(), Texture::from_image(a.image(EARTH), Color(5.0, 5.0, 2.0)),
                                                Texture::from_color(Color::from_one(0.05))}));
    scene.add(std::make_unique<Sphere>(Vector3(-3.5, 0.5, 0.0), 0.5,
                                       Material{SurfaceType::GGX(0.8), Texture::from_color(Color(1.0, 1.0, 1.0)), Texture::black(), tex_path(a, EARTH)}));
    struct Ball { double x, y, z, r, hue, rough; };
    const Ball balls[5] = {{0.5018854352719382, 0.3899602675366644, 1.8484239850862165, 0.3899602675366644, 0.2, 0.01},
                           {-0.5748933256792994, 0.2951263257801348, 2.266298272012876, 0.2951263257801348, 0.4, 0.05},
                           {-0.9865234498515534, 0.3386858117447873, 2.9809338871934585, 0.3386858117447873, 0.6, 0.02},
                           {0.6946459502665004, 0.2764689077971783, 2.7455446851003025, 0.2764689077971783, 0.05, 0.0},
                           {3.7027464198816952, 0.3917608374245498, -0.40505849281451556, 0.3917608374245498, 0.8, 0.1}};
    for (const Ball& b : balls)
        scene.add(std::make_unique<Sphere>(Vector3(b.x, b.y, b.z), b.r,
                                           Material{SurfaceType::GGX(0.8), Texture::from_color(hsv_to_rgb(Color(b.hue, 1.0, 1.0))), Texture::black(),
                                                    Texture::from_color(Color::from_one(b.rough))}));
    scene.add(std::make_unique<Cuboid>(Aabb{Vector3(-5.0, -1.0, -5.0), Vector3(5.0, 0.0, 5.0)},
                                       Material{SurfaceType::GGX(0.8), tex_path(a, MARBLE_DIFFUSE), Texture::black(), tex_path(a, MARBLE_ROUGHNESS)}));
    scene.skybox = make_skybox(a, sky, Vector3::one());
    // (the `while count < 0` loop of metal spheres draws nothing)
    int count = 0;
    while (count < 12) {  // diamonds lying on the floor
        double px = rng.gen_range(-4.5, 4.5);
        double py = 0.0;
        double pz = rng.gen_range(-2.5, 4.5);
        double s = rng.gen_range(0.7, 1.1);
        double ry = rng.gen_range(-to_radians(180.0), to_radians(180.0));
        if (scene.add_with_check_collisions(BvhMesh::from_mesh(ObjLoader::load(
                a, "models/dia/dia.obj",
                Matrix44::translate(px, py, pz) * Matrix44::scale_linear(s) * Matrix44::rotate_y(ry) * Matrix44::rotate_x(to_radians(40.35)), diamond_material()))))
            count += 1;
    }
    count = 0;
    while (count < 30) {  // diamonds floating in the air
        double px = rng.gen_range(-4.5, 4.5);
        double py = rng.gen_range(0.0, 4.0);
        double pz = rng.gen_range(-4.5, 3.5);
        double s = rng.gen_range(0.6, 1.1);
        double ry = rng.gen_range(-to_radians(180.0), to_radians(180.0));
        double rx = rng.gen_range(-to_radians(180.0), to_radians(180.0));
        if (scene.add_with_check_collisions(BvhMesh::from_mesh(ObjLoader::load(
                a, "models/dia/dia.obj", Matrix44::translate(px, py, pz) * Matrix44::scale_linear(s) * Matrix44::rotate_y(ry) * Matrix44::rotate_x(rx),
                diamond_material()))))
            count += 1;
    }
    return sc;
}
SceneAndCamera init_scene_rtcamp5(const AssetStore& a) { return scene_rtcamp5(a, "textures/cube/LancellottiChapel"); }

// src/main.rs:502-722: four emissive spheres with a textured emission (multi-light NEE), 8 metal spheres and 20
// diamonds placed by StdRng with add_with_check_collisions
static SceneAndCamera scene_tbf3(const AssetStore& a, const std::string& sky) {
    StdRng rng({870, 2000, 304, 1});
    SceneAndCamera sc;
    sc.camera = Camera(Vector3(0.0, 2.5, 9.0), Vector3(0.0, 1.5, 0.0), Vector3(0.0, 1.0, 0.0).normalize(), 19.0, LensShape::Circle, 0.18, 7.0);
    Scene& scene = sc.scene;
    scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/klab_logo/klab_logo_triangle.obj", Matrix44::scale_linear(0.4) * Matrix44::translate(0.0, 3.1782, 2.0) * Matrix44::rotate_y(-0.5),
        Material{SurfaceType::GGX(0.8), Texture::from_color(Color(0.4, 0.4, 1.0)), Texture::black(), Texture::from_color(Color::from_one(0.05))})));
    scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/dia/dia.obj",
        Matrix44::translate(1.3, 0.0, 2.2) * Matrix44::scale_linear(1.0) * Matrix44::rotate_y(-0.4) * Matrix44::rotate_x(to_radians(40.35)), diamond_material())));
    scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/dia/dia.obj",
        Matrix44::translate(-0.1, 0.0, 2.4) * Matrix44::scale_linear(1.0) * Matrix44::rotate_y(-1.4) * Matrix44::rotate_x(to_radians(40.35)), diamond_material())));
    struct Light { double x, y, z, r; Color albedo, emission; };
    const Light lights[4] = {{-1.0, 0.4, 4.0, 0.4, Color::one(), Color(3.0, 3.0, 1.1)},
                             {-3.0, 0.4, -3.5, 0.4, Color(0.5, 1.0, 1.0), Color(1.0, 3.0, 3.5)},
                             {4.0, 0.2, -4.5, 0.2, Color(0.3, 0.7, 1.0), Color(3.0, 3.0, 1.1)},
                             {3.0, 0.2, -4.2, 0.2, Color(1.0, 0.7, 0.9), Color(2.0, 3.0, 1.0)}};
    for (const Light& l : lights)
        scene.add(std::make_unique<Sphere>(Vector3(l.x, l.y, l.z), l.r,
                                           Material{SurfaceType::GGX(0.8), Texture::from_color(l.albedo), Texture::from_image(a.image(EARTH), l.emission),
                                                    Texture::from_color(Color::from_one(0.01))}));
    scene.add(std::make_unique<Cuboid>(Aabb{Vector3(-5.0, -1.0, -5.0), Vector3(5.0, 0.0, 5.0)},
                                       Material{SurfaceType::GGX(0.8), tex_path(a, MARBLE_DIFFUSE), Texture::black(), tex_path(a, MARBLE_ROUGHNESS)}));
    scene.skybox = make_skybox(a, sky, Vector3(2.0, 2.0, 3.0));
    int count = 0;
    while (count < 8) {  // metal spheres
        double px = rng.gen_range(-3.0, 3.0);
        double py = 0.0;
        double pz = rng.gen_range(-5.0, 5.0);
        double r = rng.gen_range(0.2, 0.4);
        // the struct literal draws the roughness before the collision test decides (src/main.rs:659-667)
        double rough = rng.gen_range(0.0, 0.2);
        if (scene.add_with_check_collisions(std::make_unique<Sphere>(
                Vector3(px, r + py, pz), r,
                Material{SurfaceType::GGX(0.8), Texture::from_color(hsv_to_rgb(Color(0.2 + 0.1 * (double)count, 1.0, 1.0))), Texture::black(),
                         Texture::from_color(Color::from_one(rough))})))
            count += 1;
    }
    count = 0;
    while (count < 20) {  // diamonds lying on the floor
        double px = rng.gen_range(-4.0, 4.0);
        double py = 0.0;
        double pz = rng.gen_range(-5.0, 5.0);
        double s = rng.gen_range(0.7, 1.1);
        double ry = rng.gen_range(-to_radians(180.0), to_radians(180.0));
        if (scene.add_with_check_collisions(BvhMesh::from_mesh(ObjLoader::load(
                a, "models/dia/dia.obj",
                Matrix44::translate(px, py, pz) * Matrix44::scale_linear(s) * Matrix44::rotate_y(ry) * Matrix44::rotate_x(to_radians(40.35)), diamond_material()))))
            count += 1;
    }
    return sc;
}
SceneAndCamera init_scene_tbf3(const AssetStore& a) { return scene_tbf3(a, "textures/cube/LancellottiChapel"); }

// src/main.rs:725-802
static SceneAndCamera scene_rtcamp6_v1(const AssetStore& a, const std::string& sky) {
    SceneAndCamera sc;
    sc.camera = Camera(Vector3(0.0, 2.0, 10.0), Vector3(0.0, 1.0, 0.0), Vector3(0.0, 1.0, 0.0).normalize(), 10.0, LensShape::Circle, 0.2 * 0.0, 8.8);
    double radius = 0.6;
    Scene& scene = sc.scene;
    scene.add(std::make_unique<Sphere>(Vector3(0.0, 3.1782 * 0.4, 0.0), radius,
                                       Material{SurfaceType::Diffuse(), Texture::white(), Texture::from_color(Color::from_one(10.0)), Texture::from_color(Color::from_one(0.05))}));
    scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/houdini_boss.obj", Matrix44::scale_linear(0.4) * Matrix44::translate(0.0, 3.1782, 2.0) * Matrix44::rotate_y(-0.5),
        Material{SurfaceType::Refraction(1.5), Texture::from_color(Color(0.7, 0.7, 1.0)), Texture::black(), Texture::from_color(Color::from_one(0.1))})));
    scene.add(std::make_unique<Cuboid>(Aabb{Vector3(-5.0, -1.0, -5.0), Vector3(5.0, 0.0, 5.0)},
                                       Material{SurfaceType::Diffuse(), tex_path(a, "textures/2d/checkered_diagonal_10_0.5_1.0_512.png"), Texture::black(),
                                                tex_path(a, "textures/2d/checkered_diagonal_10_0.1_0.6_512.png")}));
    scene.skybox = make_skybox(a, sky, Vector3::from_one(0.5));
    return sc;
}
SceneAndCamera init_scene_rtcamp6_v1(const AssetStore& a) { return scene_rtcamp6_v1(a, "textures/cube/LancellottiChapel"); }

// src/main.rs:928-1018: the second emitter is a 1 mm sphere one unit BEHIND the camera (smaller than the 0.02 window
// of the NEE visibility test)
SceneAndCamera init_scene_rtcamp6_v3(const AssetStore& a) {
    SceneAndCamera sc;
    sc.camera = Camera(Vector3(0.0, 2.0, 6.0), Vector3(0.0, 1.0, 0.0), Vector3(0.0, 1.0, 0.0).normalize(), 20.0, LensShape::Circle, 0.2, 4.9);
    double radius = 0.2;
    Scene& scene = sc.scene;
    scene.add(std::make_unique<Sphere>(Vector3(-0.3, 0.5 + radius, 0.0), radius,
                                       Material{SurfaceType::Diffuse(), Texture::black(), Texture::from_color(Color::from_one(10.0)), Texture::black()}));
    scene.add(std::make_unique<Sphere>(sc.camera.eye - sc.camera.forward, 0.001,
                                       Material{SurfaceType::Diffuse(), Texture::black(), Texture::from_color(Color::from_one(1000.0)), Texture::black()}));
    scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/bunny/bunny_wired_300.obj", Matrix44::scale_linear(1.5) * Matrix44::translate(0.0, 0.0, 0.0) * Matrix44::rotate_y(0.3),
        Material{SurfaceType::GGX(0.8), Texture::from_color(Color(1.0, 0.01, 0.01)), Texture::black(), Texture::from_color(Color::from_one(0.05))})));
    scene.add(std::make_unique<Cuboid>(Aabb{Vector3(-5.0, -1.0, -5.0), Vector3(5.0, 0.0, 5.0)},
                                       Material{SurfaceType::Diffuse(), Texture::white(), Texture::black(), Texture::white()}));
    scene.skybox = make_skybox(a, "textures/cube/Powerlines", Vector3::from_one(1.0));
    return sc;
}

// src/main.rs:804-925: 100 GGX spheres and FIVE emissive ones placed by StdRng (five shadow rays per NEE event)
// around a refractive fractal
static SceneAndCamera scene_rtcamp6_v2(const AssetStore& a, const std::string& sky) {
    StdRng rng({870, 2000, 304, 2});
    SceneAndCamera sc;
    sc.camera = Camera(Vector3(-5.0, -1.0, 0.0), Vector3(0.0, 0.0, 0.0), Vector3(0.0, 1.0, 0.0).normalize(), 10.0, LensShape::Circle, 0.2 * 0.0, 8.8);
    Scene& scene = sc.scene;
    scene.skybox = make_skybox(a, sky, Vector3::from_one(0.5));
    int count = 0;
    while (count < 100) {
        double px = rng.gen_range(-0.5, 2.0);
        double py = rng.gen_range(-2.0, 2.0);
        double pz = rng.gen_range(-2.0, 2.0);
        double s = 0.1;
        // the struct literal draws hue and roughness before the collision test decides (src/main.rs:870-880)
        double hue = rng.gen_range(0.0, 1.0);
        double rough = rng.gen_range(0.0, 1.0);
        if (scene.add_with_check_collisions(std::make_unique<Sphere>(
                Vector3(px, py, pz), s,
                Material{SurfaceType::GGX(0.9), Texture::from_color(hsv_to_rgb(Color(hue, 1.0, 1.0))), Texture::black(), Texture::from_color(Color::from_one(rough))})))
            count += 1;
    }
    count = 0;
    while (count < 5) {
        double px = rng.gen_range(-0.2, 0.5);
        double py = rng.gen_range(-1.0, 1.0);
        double pz = rng.gen_range(-1.0, 1.0);
        double s = 0.1;
        double hue = rng.gen_range(0.0, 1.0);
        double rough = rng.gen_range(0.0, 1.0);
        if (scene.add_with_check_collisions(std::make_unique<Sphere>(
                Vector3(px, py, pz), s,
                Material{SurfaceType::Diffuse(), Texture::black(), Texture::from_color(hsv_to_rgb(Color(hue, 1.0, 1.0)) * 10.0), Texture::from_color(Color::from_one(rough))})))
            count += 1;
    }
    scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/fractal_dodecahedron.obj", Matrix44::scale_linear(1.0) * Matrix44::translate(0.0, 0.0, 0.0) * Matrix44::rotate_y(0.0),
        Material{SurfaceType::Refraction(1.5), Texture::from_color(Color(0.7, 0.7, 1.0)), Texture::black(), Texture::from_color(Color::from_one(0.1))})));
    return sc;
}
SceneAndCamera init_scene_rtcamp6_v2(const AssetStore& a) { return scene_rtcamp6_v2(a, "textures/cube/Ryfjallet"); }

// BASELINE.md config 3 (builder-defined, not a scene of the reference): the
// default scene plus the two fractal meshes with the materials the reference
// gives them (src/main.rs:1171-1186 and :907-922), floating above the ring of
// armadillos.  ~75 k triangles / ~44 k BVH nodes.
SceneAndCamera init_scene_bvh_heavy(const AssetStore& a) {
    SceneAndCamera sc = init_scene_rtcamp6_v3_1(a);
    sc.scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/fractal_icosahedron.obj", Matrix44::translate(-2.2, 2.7, -1.2) * Matrix44::rotate_y(0.3) * Matrix44::scale_linear(0.4),
        Material{SurfaceType::GGX(0.8), Texture::from_color(Color(1.0, 1.0, 1.0)), Texture::black(), Texture::from_color(Color::from_one(0.05))})));
    sc.scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/fractal_dodecahedron.obj", Matrix44::translate(2.3, 2.5, -0.6) * Matrix44::rotate_y(0.0) * Matrix44::scale_linear(0.4),
        Material{SurfaceType::Refraction(1.5), Texture::from_color(Color(0.7, 0.7, 1.0)), Texture::black(), Texture::from_color(Color::from_one(0.1))})));
    return sc;
}

// BASELINE.md config 4 (builder-defined): round_brilliant.obj instances with
// GGXRefraction{2.42} on the textured floor, strong DoF (the commented
// alternative at src/main.rs:1035-1036: aperture 0.3, focus 5.7).
SceneAndCamera init_scene_diamond(const AssetStore& a) {
    SceneAndCamera sc = rtcamp6_stage(a, 0.3, 5.7, false);
    sc.scene.add(BvhMesh::from_mesh(ObjLoader::load(
        a, "models/round_brilliant.obj", Matrix44::translate(0.0, 0.01, 0.0) * Matrix44::rotate_y(0.3) * Matrix44::scale_linear(0.8),
        Material{SurfaceType::GGXRefraction(2.42), Texture::from_color(Color(1.0, 1.0, 1.0)), Texture::black(), Texture::from_color(Color::from_one(0.02))})));
    int count = 6;
    for (int i = 0; i < count; i++) {
        double r = 2.2;
        double dr = (double)i / (double)count;
        double theta = config::PI2 * dr;
        sc.scene.add(BvhMesh::from_mesh(ObjLoader::load(
            a, "models/round_brilliant.obj",
            Matrix44::translate(r * std::sin(theta), 0.01, r * std::cos(theta)) * Matrix44::rotate_y(theta) * Matrix44::scale_linear(0.45),
            Material{SurfaceType::GGXRefraction(2.42), Texture::from_color(hsv_to_rgb(Color(fract(0.45 + dr), 0.2, 1.0))), Texture::black(),
                     Texture::from_color(Color::from_one(0.02 + 0.02 * (double)i))})));
    }
    return sc;
}

SceneAndCamera init_scene_by_name(const std::string& name, const AssetStore& a) {
    if (name == "rtcamp6" || name == "rtcamp6_v3_1") return init_scene_rtcamp6_v3_1(a);
    if (name == "rtcamp6_v4") return init_scene_rtcamp6_v4(a);
    if (name == "simple") return init_scene_simple(a);
    if (name == "material_examples") return init_scene_material_examples(a);
    // the same two scenes under the (much smaller) Powerlines cubemap, so that they fit the committed asset pack
    if (name == "simple_pl") return scene_simple(a, "textures/cube/Powerlines");
    if (name == "material_examples_pl") return scene_material_examples(a, "textures/cube/Powerlines");
    if (name == "rtcamp6_v1") return init_scene_rtcamp6_v1(a);
    if (name == "rtcamp6_v1_pl") return scene_rtcamp6_v1(a, "textures/cube/Powerlines");
    if (name == "rtcamp6_v3") return init_scene_rtcamp6_v3(a);
    if (name == "rtcamp6_v2") return init_scene_rtcamp6_v2(a);
    if (name == "rtcamp6_v2_pl") return scene_rtcamp6_v2(a, "textures/cube/Powerlines");
    if (name == "rtcamp5") return init_scene_rtcamp5(a);
    if (name == "tbf3") return init_scene_tbf3(a);
    if (name == "rtcamp5_pl") return scene_rtcamp5(a, "textures/cube/Powerlines");
    if (name == "tbf3_pl") return scene_tbf3(a, "textures/cube/Powerlines");
    if (name == "bvh_heavy") return init_scene_bvh_heavy(a);
    if (name == "diamond") return init_scene_diamond(a);
    throw std::runtime_error("unknown scene: " + name);
}

std::vector<std::string> scene_asset_paths(const std::string& name, bool images) {
    auto cube = [](const std::string& d) {
        return std::vector<std::string>{d + "/posx.jpg", d + "/negx.jpg", d + "/posy.jpg", d + "/negy.jpg", d + "/posz.jpg", d + "/negz.jpg"};
    };
    std::vector<std::string> out;
    auto add = [&](const std::vector<std::string>& v) { out.insert(out.end(), v.begin(), v.end()); };
    bool rt = (name == "rtcamp6" || name == "rtcamp6_v3_1" || name == "bvh_heavy");
    if (images) {
        if (rt || name == "diamond") { add(cube("textures/cube/Powerlines")); out.push_back("textures/2d/magic-circle3.png"); }
        if (name == "rtcamp6_v4" || name == "rtcamp6_v2") add(cube("textures/cube/Ryfjallet"));
        if (name == "rtcamp6_v2_pl" || name == "rtcamp6_v3") add(cube("textures/cube/Powerlines"));
        if (name == "rtcamp6_v1" || name == "rtcamp6_v1_pl") {
            add(cube(name == "rtcamp6_v1" ? "textures/cube/LancellottiChapel" : "textures/cube/Powerlines"));
            out.push_back("textures/2d/checkered_diagonal_10_0.5_1.0_512.png");
            out.push_back("textures/2d/checkered_diagonal_10_0.1_0.6_512.png");
        }
        if (name == "rtcamp5" || name == "tbf3" || name == "rtcamp5_pl" || name == "tbf3_pl") {
            add(cube(name.size() > 3 && name.substr(name.size() - 3) == "_pl" ? "textures/cube/Powerlines" : "textures/cube/LancellottiChapel"));
            add({EARTH, MARBLE_DIFFUSE, MARBLE_ROUGHNESS});
        }
        if (name == "simple" || name == "material_examples" || name == "simple_pl" || name == "material_examples_pl") {
            add(cube(name.size() > 3 && name.substr(name.size() - 3) == "_pl" ? "textures/cube/Powerlines" : "textures/cube/LancellottiChapel"));
            out.push_back("textures/2d/checkered_diagonal_10_0.5_1.0_512.png");
            out.push_back("textures/2d/checkered_diagonal_10_0.1_0.6_512.png");
        }
    } else {
        if (rt) add({"models/bunny/bunny_wired_300.obj", "models/box.obj", "models/picture_frame.obj", "models/armadilo_1000.obj"});
        if (name == "bvh_heavy") add({"models/fractal_icosahedron.obj", "models/fractal_dodecahedron.obj"});
        if (name == "rtcamp6_v4") out.push_back("models/fractal_icosahedron.obj");
        if (name == "rtcamp6_v2" || name == "rtcamp6_v2_pl") out.push_back("models/fractal_dodecahedron.obj");
        if (name == "rtcamp6_v1" || name == "rtcamp6_v1_pl") out.push_back("models/houdini_boss.obj");
        if (name == "rtcamp6_v3") out.push_back("models/bunny/bunny_wired_300.obj");
        if (name == "diamond") out.push_back("models/round_brilliant.obj");
        if (name == "rtcamp5" || name == "rtcamp5_pl") add({"models/bunny/bunny_face1000.obj", "models/bunny/bunny_face1000_flip.obj", "models/dia/dia.obj"});
        if (name == "tbf3" || name == "tbf3_pl") add({"models/klab_logo/klab_logo_triangle.obj", "models/dia/dia.obj"});
    }
    return out;
}


// ---- Renderer trait over the C ABI (src/renderer.rs:20-267) ---------------------------------------------
// The CUDA core is resolved at run time: this library has no CUDA dependency, and there is no CPU fallback --
// without libhanamaru_b200.so or a device, render() fails with a message.
struct CoreApi {
    void* lib = nullptr;
    const char* (*last_error)(void) = nullptr;
    int (*device_count)(void) = nullptr;
    int (*scene_create)(const hnm_scene_desc*, int, hnm_scene**) = nullptr;
    void (*scene_destroy)(hnm_scene*) = nullptr;
    int (*renderer_create)(hnm_scene*, const hnm_camera*, uint32_t, uint32_t, int, const hnm_shard*, uint32_t, hnm_renderer**) = nullptr;
    void (*renderer_destroy)(hnm_renderer*) = nullptr;
    int (*render_passes)(hnm_renderer*, uint32_t, uint32_t) = nullptr;
    int (*synchronize)(hnm_renderer*) = nullptr;
    int (*resolve)(hnm_renderer*, const void*, uint32_t, uint8_t*) = nullptr;
};
const CoreApi* core_api(std::string* err) {
    static CoreApi api;
    static bool tried = false;
    static std::string load_error;
    if (!tried) {
        tried = true;
        std::string path;
        if (const char* e = getenv("HNM_CORE_LIB")) path = e;
        if (path.empty()) {
            Dl_info info;
            if (dladdr((const void*)&core_api, &info) && info.dli_fname) {
                std::string self = info.dli_fname;
                size_t slash = self.find_last_of('/');
                path = (slash == std::string::npos ? std::string(".") : self.substr(0, slash)) + "/libhanamaru_b200.so";
            } else {
                path = "libhanamaru_b200.so";
            }
        }
        api.lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!api.lib) {
            load_error = std::string("cannot load the CUDA core (") + path + "): " + dlerror();
        } else {
            bool ok = true;
            auto sym = [&](const char* name) { void* p = dlsym(api.lib, name); if (!p) { ok = false; load_error = std::string("missing symbol ") + name; } return p; };
            api.last_error = (const char* (*)(void))sym("hnm_last_error");
            api.device_count = (int (*)(void))sym("hnm_device_count");
            api.scene_create = (int (*)(const hnm_scene_desc*, int, hnm_scene**))sym("hnm_scene_create");
            api.scene_destroy = (void (*)(hnm_scene*))sym("hnm_scene_destroy");
            api.renderer_create = (int (*)(hnm_scene*, const hnm_camera*, uint32_t, uint32_t, int, const hnm_shard*, uint32_t, hnm_renderer**))sym("hnm_renderer_create");
            api.renderer_destroy = (void (*)(hnm_renderer*))sym("hnm_renderer_destroy");
            api.render_passes = (int (*)(hnm_renderer*, uint32_t, uint32_t))sym("hnm_render_passes");
            api.synchronize = (int (*)(hnm_renderer*))sym("hnm_synchronize");
            api.resolve = (int (*)(hnm_renderer*, const void*, uint32_t, uint8_t*))sym("hnm_resolve");
            if (!ok) { dlclose(api.lib); api.lib = nullptr; }
        }
    }
    if (!api.lib) { if (err) *err = load_error; return nullptr; }
    return &api;
}

static double now_sec() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

uint32_t Renderer::render(const BvhScene& scene, const Camera& camera, ImageBuffer& imgbuf) { return render(scene, camera.abi(), imgbuf); }

// src/renderer.rs:25-46: the pass loop; `sampling` is 1-origin; report_progress after every call may stop it
uint32_t Renderer::render(const BvhScene& scene, const hnm_camera& camera, ImageBuffer& imgbuf) {
    error.clear();
    const CoreApi* api = core_api(&error);
    if (!api) return 0;
    if (api->device_count() < 1) { error = std::string("no CUDA device: ") + api->last_error(); return 0; }
    hnm_scene* ds = nullptr;
    if (api->scene_create(&scene.flat.desc, device, &ds) != 0) { error = api->last_error(); return 0; }
    if (api->renderer_create(ds, &camera, imgbuf.width, imgbuf.height, mode(), nullptr, 0, &r_) != 0) {
        error = api->last_error();
        api->scene_destroy(ds);
        return 0;
    }
    uint32_t sampling = 0;
    const uint32_t limit = max_sampling();
    // passes per call: the reference reports after EVERY pass (src/renderer.rs:41); a device call per pass would leave the
    // GPU under-filled at small images (one 480x270 pass is 0.5 M paths; the wavefront wants 16 M in flight) and pay a host
    // synchronisation each time.  `auto` asks for the number of passes that fills the device and lets the renderer shrink it
    // as its time limit approaches (PathTracingRenderer::next_call_passes), so that `-t` still ends on time.
    const uint64_t per_pass = (uint64_t)imgbuf.width * imgbuf.height * 4;
    const uint32_t fill = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, ((16ull << 20) + per_pass - 1) / per_pass));
    while (sampling < limit) {
        uint32_t want = passes_per_call ? passes_per_call : next_call_passes(sampling, fill);
        uint32_t n = std::max(1u, std::min(want, limit - sampling));
        if (api->render_passes(r_, sampling + 1, n) != 0 || api->synchronize(r_) != 0) { error = api->last_error(); break; }
        sampling += n;
        if (report_progress(sampling, imgbuf)) break;
    }
    api->renderer_destroy(r_);
    r_ = nullptr;
    api->scene_destroy(ds);
    return error.empty() ? sampling : 0;
}
void Renderer::update_imgbuf(uint32_t sampling, ImageBuffer& imgbuf) {
    const CoreApi* api = core_api(nullptr);
    if (api && r_ && api->resolve(r_, nullptr, sampling, imgbuf.rgb.data()) != 0) error = api->last_error();
}
// src/renderer.rs:92-98: update_imgbuf, then `image::ImageRgb8(imgbuf.clone()).save(path)` with path = "{:>03}.png"
void Renderer::save_progress_image(uint32_t counter, uint32_t sampling, ImageBuffer& imgbuf) {
    update_imgbuf(sampling, imgbuf);
    if (save_dir.empty() || !error.empty()) return;
    char name[32];
    snprintf(name, sizeof(name), "%03u.png", counter);
    std::string err;
    if (!save_png(save_dir + "/" + name, imgbuf.rgb.data(), imgbuf.width, imgbuf.height, &err)) error = err;
}
bool DebugRenderer::report_progress(uint32_t sampling, ImageBuffer& imgbuf) {  // src/renderer.rs:141-145
    update_imgbuf(sampling, imgbuf);
    return true;
}
PathTracingRenderer::PathTracingRenderer(uint32_t sampling, double time_limit_sec, double report_interval_sec)
    : sampling_(sampling), time_limit_sec_(time_limit_sec), report_interval_sec_(report_interval_sec) {
    begin_ = last_report_progress_ = last_report_image_ = now_sec();
}
// src/renderer.rs:205-251: stop when the time limit would be exceeded by another call like the last one (x1.1) or when
// max sampling is reached; refresh the image every report interval
bool PathTracingRenderer::report_progress(uint32_t sampling, ImageBuffer& imgbuf) {
    const double now = now_sec();
    const double used = now - begin_;
    const double from_last = now - last_report_progress_;
    if (verbose)
        fprintf(stderr, "rendering: %ux4 sampled (last %.3f sec). total: %.3f sec (%.2f %%).\n", sampling, from_last, used, used / time_limit_sec_ * 100.0);
    if (last_call_passes_ > 0) sec_per_pass_ = from_last / last_call_passes_;
    // `offset = from_last_sampling_sec * 1.1` (src/renderer.rs:218): the prediction is for the NEXT call.  With a fixed
    // passes_per_call that is a call like the last one; in auto mode the next call is sized to fit (plan_next), so the
    // loop stops exactly when not even one more pass fits -- the reference's rule at one pass per call.
    const double predicted = (last_call_passes_ ? sec_per_pass_ * plan_next(sampling, now) : from_last) * 1.1;
    if (used + predicted > time_limit_sec_ || sampling >= max_sampling()) {
        save_progress_image(report_image_counter_, sampling, imgbuf);
        return true;
    }
    if (now - last_report_image_ >= report_interval_sec_) {
        save_progress_image(report_image_counter_, sampling, imgbuf);
        report_image_counter_ += 1;
        last_report_image_ = now;
    }
    last_report_progress_ = now_sec();
    return false;
}
// how many passes the next device call should run: the filling batch, shrunk so that (passes x the measured time per
// pass x 1.1) still fits before the time limit and so that the next progress image is not overshot by a whole batch
uint32_t PathTracingRenderer::plan_next(uint32_t sampling, double now) const {
    if (!(sec_per_pass_ > 0.0)) return 1;  // the first call measures one pass
    uint32_t n = std::max(1u, fill_);
    const double left = time_limit_sec_ - (now - begin_);
    const double fit = left / (sec_per_pass_ * 1.1);
    if (fit < (double)n) n = fit >= 1.0 ? (uint32_t)fit : 1u;
    const double to_report = report_interval_sec_ - (now - last_report_image_);
    if (to_report > 0.0 && to_report / sec_per_pass_ < (double)n) n = std::max(1u, (uint32_t)std::ceil(to_report / sec_per_pass_));
    const uint32_t remaining = max_sampling() > sampling ? max_sampling() - sampling : 1u;
    return std::max(1u, std::min(n, remaining));
}
uint32_t PathTracingRenderer::next_call_passes(uint32_t sampling, uint32_t fill) {
    fill_ = fill;
    last_call_passes_ = plan_next(sampling, now_sec());
    return last_call_passes_;
}

bool save_png(const std::string& path, const uint8_t* rgb, uint32_t width, uint32_t height, std::string* err) {
    std::vector<uint8_t> bytes;
    if (!png_encode_rgb8(rgb, width, height, bytes, err)) return false;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { if (err) *err = "cannot write " + path; return false; }
    const bool ok = fwrite(bytes.data(), 1, bytes.size(), f) == bytes.size();
    fclose(f);
    if (!ok && err) *err = "short write to " + path;
    return ok;
}

}  // namespace hanamaru
