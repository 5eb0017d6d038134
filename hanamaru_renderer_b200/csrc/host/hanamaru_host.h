// hanamaru_host.h -- C++ mirror of the reference's HOST side.
//
// The reference's host is Rust (no Rust toolchain in this image), so
// the types a maintainer would keep on the Rust side are restated here in C++
// with the same names and meaning: Vector3 (src/vector.rs), Matrix44
// (src/matrix.rs), Camera::new (src/camera.rs:45-64), Texture / Material
// (src/texture.rs, src/material.rs), Sphere / Cuboid / Mesh / BvhMesh / Scene /
// BvhScene (src/scene.rs), the object-median BVH builder (src/bvh.rs:107-211),
// ObjLoader (src/loader.rs) and the Renderer trait (src/renderer.rs:20-99).
// None of this is on the hot path: it produces the flat hnm_scene_desc that
// crosses the C ABI (include/hanamaru_b200.h) and drives the pass loop.
#ifndef HANAMARU_HOST_H
#define HANAMARU_HOST_H

#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "hanamaru_b200.h"

namespace hanamaru {

namespace config {  // src/config.rs:4-25
constexpr double PI = 3.14159265358979323846;
constexpr double PI2 = 2.0 * PI;
constexpr double EPS = 1e-4;
constexpr double OFFSET = 1e-4;
constexpr double INF = 1e100;
constexpr double GAMMA_FACTOR = 2.2;
constexpr uint32_t SUPERSAMPLING = 2;
constexpr uint32_t PATHTRACING_BOUNCE_LIMIT = 10;
constexpr uint32_t TONE_MAPPING_MODE = 1;  // Reinhard
constexpr double TONE_MAPPING_EXPOSURE = 1.5;
constexpr double TONE_MAPPING_WHITE_POINT = 20.0;
constexpr uint32_t BILATERAL_FILTER_ITERATION = 1;
constexpr uint32_t BILATERAL_FILTER_DIAMETER = 3;
constexpr double BILATERAL_FILTER_SIGMA_I = 1.0;
constexpr double BILATERAL_FILTER_SIGMA_S = 16.0;
hnm_config to_abi();
}  // namespace config

// src/vector.rs (only what the host side uses)
struct Vector3 {
    double x = 0, y = 0, z = 0;
    Vector3() = default;
    Vector3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    static Vector3 zero() { return from_one(0.0); }
    static Vector3 one() { return from_one(1.0); }
    static Vector3 from_one(double v) { return Vector3(v, v, v); }
    double norm() const { return x * x + y * y + z * z; }
    double length() const { return std::sqrt(norm()); }
    Vector3 normalize() const {
        double inv_len = 1.0 / length();
        return Vector3(x * inv_len, y * inv_len, z * inv_len);
    }
    double dot(const Vector3& o) const { return x * o.x + y * o.y + z * o.z; }
    Vector3 cross(const Vector3& o) const {
        return Vector3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x);
    }
    bool operator==(const Vector3& o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator!=(const Vector3& o) const { return !(*this == o); }
    hnm_vec3 abi() const { return hnm_vec3{x, y, z}; }
};
inline Vector3 operator+(Vector3 a, Vector3 b) { return Vector3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vector3 operator-(Vector3 a, Vector3 b) { return Vector3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vector3 operator+(Vector3 a, double b) { return Vector3(a.x + b, a.y + b, a.z + b); }
inline Vector3 operator-(Vector3 a, double b) { return Vector3(a.x - b, a.y - b, a.z - b); }
inline Vector3 operator*(Vector3 a, double b) { return Vector3(a.x * b, a.y * b, a.z * b); }
inline Vector3 operator*(double a, Vector3 b) { return b * a; }
inline Vector3 operator*(Vector3 a, Vector3 b) { return Vector3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vector3 operator/(Vector3 a, double b) { return Vector3(a.x / b, a.y / b, a.z / b); }
inline Vector3 operator-(Vector3 a) { return Vector3(-a.x, -a.y, -a.z); }
using Color = Vector3;

Color hsv_to_rgb(Color c);  // src/color.rs:51-61

// src/matrix.rs
struct Matrix44 {
    double e[4][4];
    static Matrix44 identity();
    static Matrix44 scale_linear(double s) { return scale(s, s, s); }
    static Matrix44 scale(double sx, double sy, double sz);
    static Matrix44 rotate_x(double t);
    static Matrix44 rotate_y(double t);
    static Matrix44 rotate_z(double t);
    static Matrix44 translate(double tx, double ty, double tz);
    Matrix44 operator*(const Matrix44& o) const;
    Vector3 operator*(const Vector3& v) const;
};

// src/camera.rs
enum class LensShape { Square = 0, Circle = 1 };
struct Camera {
    Vector3 eye;
    LensShape lens_shape;
    double lens_radius, focus_distance;
    Vector3 right, up, forward, plane_half_right, plane_half_up;
    Camera() = default;
    Camera(Vector3 eye, Vector3 target, Vector3 y_up, double v_fov, LensShape lens_shape,
           double aperture, double focus_distance);
    hnm_camera abi() const;
};

// decoded image (what `image::open` + get_pixel gives: RGBA8, row 0 = top)
struct Image {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgba;
};

// local-space OBJ geometry as parsed by src/loader.rs before the matrix is applied
struct ObjGeometry {
    std::vector<Vector3> vertexes;
    std::vector<uint32_t> faces;  // v0 v1 v2 triples, 0-based
};

// `image::open` / `DynamicImage::save` of the reference (src/texture.rs:18, src/renderer.rs:97, src/main.rs:1217):
// hanamaru_image.cpp.  PNG (8/16-bit, all colour types, non-interlaced) and baseline JPEG decode; 8-bit RGB PNG encode.
bool png_encode_rgb8(const uint8_t* rgb, uint32_t width, uint32_t height, std::vector<uint8_t>& out, std::string* err);
bool png_decode(const uint8_t* data, size_t n, Image& out, std::string* err);
bool jpeg_decode(const uint8_t* data, size_t n, Image& out, std::string* err);
bool image_decode(const uint8_t* data, size_t n, Image& out, std::string* err);
bool save_png(const std::string& path, const uint8_t* rgb, uint32_t width, uint32_t height, std::string* err);

// Where OBJ files and decoded images come from: a directory laid out like the
// reference checkout (OBJ text and PNG / JPEG files, decoded here), images
// registered by the caller, and/or an asset pack written by tools/make_asset_pack.py.
class AssetStore {
  public:
    void set_root(const std::string& dir) { root_ = dir; }
    bool load_pack(const std::string& path, std::string* err);
    void put_image(const std::string& path, std::shared_ptr<Image> img) { images_[path] = img; }
    void put_obj(const std::string& path, std::shared_ptr<ObjGeometry> g) { objs_[path] = g; }
    std::shared_ptr<Image> image(const std::string& path) const;        // throws std::runtime_error
    std::shared_ptr<ObjGeometry> obj(const std::string& path) const;    // throws std::runtime_error
  private:
    std::string root_;
    mutable std::map<std::string, std::shared_ptr<Image>> images_;
    std::map<std::string, std::shared_ptr<std::vector<uint8_t>>> encoded_;  // image files from a pack, not decoded yet
    mutable std::map<std::string, std::shared_ptr<ObjGeometry>> objs_;
};

// src/loader.rs:12-59 on the text of an OBJ file
ObjGeometry parse_obj(const std::string& text);

// src/texture.rs:72-114
struct Texture {
    std::shared_ptr<Image> image_texture;  // None = nullptr
    Color color;
    static Texture from_image(std::shared_ptr<Image> img, Color c = Color::one()) { return Texture{img, c}; }
    static Texture from_color(Color c) { return Texture{nullptr, c}; }
    static Texture white() { return from_color(Color::one()); }
    static Texture black() { return from_color(Color::zero()); }
};

// src/material.rs:9-23
struct SurfaceType {
    int32_t tag = HNM_SURFACE_DIFFUSE;
    double param = 0.0;
    static SurfaceType Diffuse() { return {HNM_SURFACE_DIFFUSE, 0.0}; }
    static SurfaceType Specular() { return {HNM_SURFACE_SPECULAR, 0.0}; }
    static SurfaceType Refraction(double ri) { return {HNM_SURFACE_REFRACTION, ri}; }
    static SurfaceType GGX(double f0) { return {HNM_SURFACE_GGX, f0}; }
    static SurfaceType GGXRefraction(double ri) { return {HNM_SURFACE_GGX_REFRACTION, ri}; }
};
struct Material {
    SurfaceType surface;
    Texture albedo, emission, roughness;
};

// src/bvh.rs:7-66
struct Aabb {
    Vector3 min, max;
    bool intersect_aabb(const Aabb& o) const;
    void merge(const Aabb& o);
};
// src/bvh.rs:68-77 -- pointer tree exactly as the reference builds it
struct BvhNode {
    Aabb aabb;
    std::vector<std::unique_ptr<BvhNode>> children;  // 0 or 2
    std::vector<size_t> indexes;
};

struct Face { size_t v0, v1, v2; };
struct Mesh {
    std::vector<Vector3> vertexes;
    std::vector<Face> faces;
    Material material;
};

class FlatSceneBuilder;

// src/scene.rs:42-49 plus the one additive method the boundary needs
struct Intersectable {
    virtual ~Intersectable() = default;
    virtual const Material& material() const = 0;
    virtual Aabb aabb() const = 0;
    virtual bool nee_available() const = 0;
    virtual void flatten(FlatSceneBuilder& b) const = 0;
};
struct Sphere : Intersectable {
    Vector3 center; double radius; Material mat;
    Sphere(Vector3 c, double r, Material m) : center(c), radius(r), mat(std::move(m)) {}
    const Material& material() const override { return mat; }
    Aabb aabb() const override;
    bool nee_available() const override { return true; }
    void flatten(FlatSceneBuilder& b) const override;
};
struct Cuboid : Intersectable {
    Aabb box; Material mat;
    Cuboid(Aabb a, Material m) : box(a), mat(std::move(m)) {}
    const Material& material() const override { return mat; }
    Aabb aabb() const override { return box; }
    bool nee_available() const override { return false; }
    void flatten(FlatSceneBuilder& b) const override;
};
struct BvhMesh : Intersectable {
    Mesh mesh; std::unique_ptr<BvhNode> bvh;
    static std::unique_ptr<BvhMesh> from_mesh(Mesh mesh);  // src/scene.rs:257-265
    const Material& material() const override { return mesh.material; }
    Aabb aabb() const override { return bvh->aabb; }
    bool nee_available() const override { return false; }
    void flatten(FlatSceneBuilder& b) const override;
};

struct ObjLoader {  // src/loader.rs
    static Mesh load(const AssetStore& assets, const std::string& path, const Matrix44& matrix, Material material);
};

struct Skybox {  // src/scene.rs:268-293
    std::shared_ptr<Image> px, nx, py, ny, pz, nz;
    Vector3 intensity;
};

struct Scene {  // src/scene.rs:327-377
    std::vector<std::unique_ptr<Intersectable>> elements;
    Skybox skybox;
    void add(std::unique_ptr<Intersectable> e) { elements.push_back(std::move(e)); }
    bool add_with_check_collisions(std::unique_ptr<Intersectable> e);
    std::vector<uint32_t> emissions() const;  // element ids, src/scene.rs:356-358
};

std::unique_ptr<BvhNode> build_from_mesh(const Mesh& mesh);    // src/bvh.rs:203-206
std::unique_ptr<BvhNode> build_from_scene(const Scene& scene); // src/bvh.rs:208-211

// Owns every array an hnm_scene_desc points to.
struct FlatScene {
    std::vector<hnm_element> elements;
    std::vector<hnm_material> materials;
    std::vector<hnm_image> images;
    std::vector<std::shared_ptr<Image>> image_refs;
    std::vector<hnm_mesh> meshes;
    std::vector<double> vertices;
    std::vector<uint32_t> faces;
    std::vector<hnm_bvh_node> mesh_nodes;
    std::vector<uint32_t> mesh_indices;
    std::vector<hnm_bvh_node> top_nodes;
    std::vector<uint32_t> top_indices;
    std::vector<uint32_t> emissions;
    hnm_scene_desc desc;
    void finalize();
};

class FlatSceneBuilder {
  public:
    explicit FlatSceneBuilder(FlatScene& out) : out_(out) {}
    int32_t add_image(const std::shared_ptr<Image>& img);
    int32_t add_material(const Material& m);
    void add_sphere(const Sphere& s);
    void add_cuboid(const Cuboid& c);
    void add_mesh(const BvhMesh& m);
    static void flatten_tree(const BvhNode& root, std::vector<hnm_bvh_node>& nodes, std::vector<uint32_t>& indices,
                             uint32_t index_base);
  private:
    FlatScene& out_;
    std::map<const Image*, int32_t> image_ids_;
};

// BvhScene::from_scene (src/scene.rs:409-415) + flattening for the ABI
struct BvhScene {
    Scene scene;
    std::unique_ptr<BvhNode> bvh;
    FlatScene flat;
    static std::unique_ptr<BvhScene> from_scene(Scene scene);
};

// rand 0.4.3 `StdRng` (= Isaac64Rng on 64-bit targets) as the scene builders use it (src/main.rs:253-254,431-482):
// SeedableRng::from_seed(&[usize]) and Rng::gen_range(low, high) on f64.  Third-party code that is not in the
// reference tree, restated from its published source; the generator core is pinned by rand's own
// known-answer vectors (tests/test_host_and_abi.py), the f64 mapping is not (parity unpinned at that step).
class StdRng {
  public:
    explicit StdRng(const std::vector<uint64_t>& seed);
    uint64_t next_u64();
    double next_f64();                       // [0, 1): 0x3FF0... | (u64 & (2^52 - 1)), minus 1.0
    double gen_range(double low, double high);  // low + (high - low) * next_f64()
  private:
    void isaac64();
    uint64_t rsl_[256], mem_[256], a_ = 0, b_ = 0, c_ = 0;
    uint32_t cnt_ = 0;
};

// Scene authoring (src/main.rs): the default scene and the two builder-defined
// benchmark scenes of BASELINE.md section 3.
struct SceneAndCamera { Camera camera; Scene scene; };
SceneAndCamera init_scene_rtcamp6_v3_1(const AssetStore& a);  // src/main.rs:1020-1153
SceneAndCamera init_scene_rtcamp6_v4(const AssetStore& a);    // src/main.rs:1155-1212
SceneAndCamera init_scene_simple(const AssetStore& a);        // src/main.rs:54-131
SceneAndCamera init_scene_material_examples(const AssetStore& a);  // src/main.rs:133-250
SceneAndCamera init_scene_rtcamp6_v1(const AssetStore& a);    // src/main.rs:725-802
SceneAndCamera init_scene_rtcamp6_v3(const AssetStore& a);    // src/main.rs:928-1018 (1 mm emitter behind the camera)
SceneAndCamera init_scene_rtcamp6_v2(const AssetStore& a);    // src/main.rs:804-925 (105 StdRng-placed spheres, five emitters)
SceneAndCamera init_scene_rtcamp5(const AssetStore& a);       // src/main.rs:252-499 (45 diamonds placed by StdRng)
SceneAndCamera init_scene_tbf3(const AssetStore& a);          // src/main.rs:502-722 (four textured emitters)
SceneAndCamera init_scene_bvh_heavy(const AssetStore& a);     // BASELINE config 3 (builder-defined)
SceneAndCamera init_scene_diamond(const AssetStore& a);       // BASELINE config 4 (builder-defined)
SceneAndCamera init_scene_by_name(const std::string& name, const AssetStore& a);
std::vector<std::string> scene_asset_paths(const std::string& name, bool images);

// ---- Renderer trait (src/renderer.rs:20-99) over the C ABI ------------------------
struct ImageBuffer {  // image::ImageBuffer<Rgb<u8>, Vec<u8>>
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgb;
    ImageBuffer(uint32_t w, uint32_t h) : width(w), height(h), rgb((size_t)w * h * 3, 0) {}
};

// Resolves the entry points of libhanamaru_b200.so at run time (dlopen), so the
// host library itself has no CUDA dependency.
struct CoreApi;
const CoreApi* core_api(std::string* err);

class Renderer {
  public:
    virtual ~Renderer() = default;
    virtual uint32_t max_sampling() const = 0;
    virtual int mode() const = 0;
    // src/renderer.rs:25-46: pass loop on the device, report_progress per batch
    uint32_t render(const BvhScene& scene, const Camera& camera, ImageBuffer& imgbuf);
    // the same with the camera already in ABI form (what hnmh_scene_camera hands out)
    uint32_t render(const BvhScene& scene, const hnm_camera& camera, ImageBuffer& imgbuf);
    // src/renderer.rs:62 -- true = stop
    virtual bool report_progress(uint32_t sampling, ImageBuffer& imgbuf) = 0;
    int device = 0;
    uint32_t passes_per_call = 0;  // 0 = auto: as many passes per call as fill the device, fewer near the time limit
    std::string error;
    // `save_progress_image` (src/renderer.rs:92-98): when non-empty, every image report_progress produces is also written
    // as "<save_dir>/NNN.png" (NNN = the report counter), like the reference writes NNN.png into its cwd
    std::string save_dir;
  protected:
    void update_imgbuf(uint32_t sampling, ImageBuffer& imgbuf);  // src/renderer.rs:64-90 via hnm_resolve
    void save_progress_image(uint32_t counter, uint32_t sampling, ImageBuffer& imgbuf);  // src/renderer.rs:92-98
    virtual uint32_t next_call_passes(uint32_t sampling, uint32_t fill) { (void)sampling; return fill; }
    hnm_renderer* r_ = nullptr;
};

class DebugRenderer : public Renderer {  // src/renderer.rs:109-146
  public:
    explicit DebugRenderer(int debug_mode) : mode_(debug_mode) {}
    uint32_t max_sampling() const override { return 1; }
    int mode() const override { return mode_; }
    bool report_progress(uint32_t sampling, ImageBuffer& imgbuf) override;
  private:
    int mode_;
};

class PathTracingRenderer : public Renderer {  // src/renderer.rs:148-267
  public:
    PathTracingRenderer(uint32_t sampling, double time_limit_sec, double report_interval_sec);
    uint32_t max_sampling() const override { return sampling_; }
    int mode() const override { return HNM_MODE_PATHTRACING; }
    bool report_progress(uint32_t sampling, ImageBuffer& imgbuf) override;
    bool verbose = false;
  protected:
    uint32_t next_call_passes(uint32_t sampling, uint32_t fill) override;
  private:
    uint32_t sampling_;
    double time_limit_sec_, report_interval_sec_;
    double begin_, last_report_progress_, last_report_image_;
    uint32_t report_image_counter_ = 0;
    uint32_t plan_next(uint32_t sampling, double now) const;
    uint32_t last_call_passes_ = 0, fill_ = 1;
    double sec_per_pass_ = 0.0;
};

}  // namespace hanamaru
#endif
