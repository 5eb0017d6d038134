// host_capi.cpp -- C facade over the C++ host mirror, for ctypes (tests, bench,
// tools).  Everything returns 0 / a handle on success; hnmh_last_error() has
// the message otherwise.
#include <cstring>
#include <stdexcept>

#include "hanamaru_host.h"

using namespace hanamaru;

static thread_local std::string g_err;

struct SceneHandle {
    std::unique_ptr<BvhScene> bvh_scene;
    hnm_camera camera;
};
struct BuilderHandle {
    Scene scene;
    Camera camera;
    bool has_camera = false;
};

#define HNMH_TRY try {
#define HNMH_CATCH(ret)                                 \
    }                                                   \
    catch (const std::exception& e) {                   \
        g_err = e.what();                               \
        return ret;                                     \
    }

extern "C" {

const char* hnmh_last_error() { return g_err.c_str(); }

void* hnmh_assets_create() { return new AssetStore(); }
void hnmh_assets_destroy(void* a) { delete (AssetStore*)a; }
int hnmh_assets_set_root(void* a, const char* dir) { ((AssetStore*)a)->set_root(dir); return 0; }
int hnmh_assets_load_pack(void* a, const char* path) {
    std::string err;
    if (!((AssetStore*)a)->load_pack(path, &err)) { g_err = err; return -1; }
    return 0;
}
int hnmh_assets_put_image(void* a, const char* name, const uint8_t* rgba, uint32_t w, uint32_t h) {
    auto img = std::make_shared<Image>();
    img->width = w; img->height = h;
    img->rgba.assign(rgba, rgba + (size_t)w * h * 4);
    ((AssetStore*)a)->put_image(name, img);
    return 0;
}
int hnmh_assets_put_obj_text(void* a, const char* name, const char* text, size_t len) {
    HNMH_TRY
    auto g = std::make_shared<ObjGeometry>(parse_obj(std::string(text, len)));
    ((AssetStore*)a)->put_obj(name, g);
    return 0;
    HNMH_CATCH(-1)
}
// parsed local-space geometry of an OBJ known to the store (src/loader.rs semantics)
int hnmh_assets_obj_counts(void* a, const char* name, uint32_t* nverts, uint32_t* nfaces) {
    HNMH_TRY
    auto g = ((AssetStore*)a)->obj(name);
    *nverts = (uint32_t)g->vertexes.size();
    *nfaces = (uint32_t)(g->faces.size() / 3);
    return 0;
    HNMH_CATCH(-1)
}
int hnmh_assets_obj_copy(void* a, const char* name, double* verts, uint32_t* faces) {
    HNMH_TRY
    auto g = ((AssetStore*)a)->obj(name);
    for (size_t i = 0; i < g->vertexes.size(); i++) { verts[3 * i] = g->vertexes[i].x; verts[3 * i + 1] = g->vertexes[i].y; verts[3 * i + 2] = g->vertexes[i].z; }
    memcpy(faces, g->faces.data(), g->faces.size() * 4);
    return 0;
    HNMH_CATCH(-1)
}

// newline separated asset paths a named scene needs
int hnmh_scene_asset_paths(const char* name, int images, char* buf, size_t buflen) {
    std::string out;
    for (const std::string& p : scene_asset_paths(name, images != 0)) { out += p; out += '\n'; }
    if (out.size() + 1 > buflen) { g_err = "buffer too small"; return -1; }
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

void* hnmh_scene_build(void* assets, const char* name) {
    HNMH_TRY
    SceneAndCamera sc = init_scene_by_name(name, *(AssetStore*)assets);
    auto h = new SceneHandle();
    h->camera = sc.camera.abi();
    h->bvh_scene = BvhScene::from_scene(std::move(sc.scene));
    return h;
    HNMH_CATCH(nullptr)
}
const hnm_scene_desc* hnmh_scene_desc(void* h) { return &((SceneHandle*)h)->bvh_scene->flat.desc; }
const hnm_camera* hnmh_scene_camera(void* h) { return &((SceneHandle*)h)->camera; }
void hnmh_scene_destroy(void* h) { delete (SceneHandle*)h; }

// ---- ad-hoc scenes (tests) -------------------------------------------------------
typedef struct hnmh_material {
    int32_t surface; int32_t _pad;
    double param;
    double albedo[3], emission[3], roughness[3];
    const char* albedo_image;    // asset names or NULL
    const char* emission_image;
    const char* roughness_image;
} hnmh_material;

static Material make_material(const AssetStore* a, const hnmh_material* m) {
    auto tex = [&](const double* c, const char* img) {
        Texture t;
        t.color = Color(c[0], c[1], c[2]);
        if (img) t.image_texture = a->image(img);
        return t;
    };
    Material out;
    out.surface = SurfaceType{m->surface, m->param};
    out.albedo = tex(m->albedo, m->albedo_image);
    out.emission = tex(m->emission, m->emission_image);
    out.roughness = tex(m->roughness, m->roughness_image);
    return out;
}

// The reference-facing call in C++: `renderer.render(&scene, &camera, &mut imgbuf)` (src/main.rs:1216) with the C++
// Renderer classes of hanamaru_host.h over the CUDA core.  mode 0 = PathTracingRenderer(sampling, time_limit, interval),
// 1..4 = DebugRenderer.  rgb8 = width * height * 3 bytes; *passes_done = the return value of render().
int hnmh_render(void* scene_handle, int mode, uint32_t width, uint32_t height, uint32_t sampling, double time_limit_sec,
                double report_interval_sec, uint32_t passes_per_call, int device, uint8_t* rgb8, uint32_t* passes_done) {
    if (!scene_handle || !rgb8 || !passes_done) { g_err = "null argument"; return -1; }
    HNMH_TRY
    SceneHandle* h = (SceneHandle*)scene_handle;
    ImageBuffer img(width, height);
    uint32_t done = 0;
    std::string err;
    if (mode == HNM_MODE_PATHTRACING) {
        PathTracingRenderer r(sampling, time_limit_sec, report_interval_sec);
        r.device = device;
        r.passes_per_call = passes_per_call;
        done = r.render(*h->bvh_scene, h->camera, img);
        err = r.error;
    } else {
        DebugRenderer r(mode);
        r.device = device;
        done = r.render(*h->bvh_scene, h->camera, img);
        err = r.error;
    }
    if (!err.empty()) { g_err = err; return -1; }
    memcpy(rgb8, img.rgb.data(), img.rgb.size());
    *passes_done = done;
    return 0;
    HNMH_CATCH(-1)
}

// `image::open` on bytes in memory: *width / *height always; rgba (width * height * 4 bytes) when non-null.  Two-call
// protocol: first with rgba == NULL to learn the size.
int hnmh_image_decode(const uint8_t* bytes, size_t n, uint32_t* width, uint32_t* height, uint8_t* rgba) {
    if (!bytes || !width || !height) { g_err = "null argument"; return -1; }
    HNMH_TRY
    Image img;
    std::string err;
    if (!image_decode(bytes, n, img, &err)) { g_err = err; return -1; }
    *width = img.width; *height = img.height;
    if (rgba) memcpy(rgba, img.rgba.data(), img.rgba.size());
    return 0;
    HNMH_CATCH(-1)
}
// `DynamicImage::save(path)` for an 8-bit RGB buffer (src/main.rs:1217 result.png, src/renderer.rs:97 NNN.png)
int hnmh_save_png(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height) {
    if (!path || !rgb) { g_err = "null argument"; return -1; }
    HNMH_TRY
    std::string err;
    if (!save_png(path, rgb, width, height, &err)) { g_err = err; return -1; }
    return 0;
    HNMH_CATCH(-1)
}
// hnmh_render + the reference's file outputs: progress / final images as "<out_dir>/NNN.png" (src/renderer.rs:92-98) and
// the final image as "<out_dir>/result.png" (src/main.rs:1217)
int hnmh_render_to_files(void* scene_handle, int mode, uint32_t width, uint32_t height, uint32_t sampling, double time_limit_sec,
                         double report_interval_sec, uint32_t passes_per_call, int device, const char* out_dir, uint8_t* rgb8,
                         uint32_t* passes_done) {
    if (!scene_handle || !out_dir || !passes_done) { g_err = "null argument"; return -1; }
    HNMH_TRY
    SceneHandle* h = (SceneHandle*)scene_handle;
    ImageBuffer img(width, height);
    uint32_t done = 0;
    std::string err;
    if (mode == HNM_MODE_PATHTRACING) {
        PathTracingRenderer r(sampling, time_limit_sec, report_interval_sec);
        r.device = device;
        r.passes_per_call = passes_per_call;
        r.save_dir = out_dir;
        done = r.render(*h->bvh_scene, h->camera, img);
        err = r.error;
    } else {
        DebugRenderer r(mode);
        r.device = device;
        done = r.render(*h->bvh_scene, h->camera, img);
        err = r.error;
    }
    if (!err.empty()) { g_err = err; return -1; }
    if (!save_png(std::string(out_dir) + "/result.png", img.rgb.data(), width, height, &err)) { g_err = err; return -1; }
    if (rgb8) memcpy(rgb8, img.rgb.data(), img.rgb.size());
    *passes_done = done;
    return 0;
    HNMH_CATCH(-1)
}

// rand 0.4.3 StdRng as the scene builders use it: `count` u64 outputs (kind 0) or gen_range(low, high) f64 draws
// (kind 1, written as doubles) after skipping `skip` outputs -- for the known-answer tests
int hnmh_stdrng(const uint64_t* seed, uint32_t nseed, uint32_t skip, uint32_t count, int kind, double low, double high, void* out) {
    if (!seed || !out) { g_err = "null argument"; return -1; }
    StdRng rng(std::vector<uint64_t>(seed, seed + nseed));
    for (uint32_t i = 0; i < skip; i++) rng.next_u64();
    for (uint32_t i = 0; i < count; i++) {
        if (kind == 0) ((uint64_t*)out)[i] = rng.next_u64();
        else ((double*)out)[i] = rng.gen_range(low, high);
    }
    return 0;
}

void* hnmh_builder_create() { return new BuilderHandle(); }
void hnmh_builder_destroy(void* b) { delete (BuilderHandle*)b; }
int hnmh_builder_camera(void* b, const double* eye, const double* target, const double* y_up, double v_fov, int lens_shape,
                        double aperture, double focus_distance) {
    auto* h = (BuilderHandle*)b;
    h->camera = Camera(Vector3(eye[0], eye[1], eye[2]), Vector3(target[0], target[1], target[2]), Vector3(y_up[0], y_up[1], y_up[2]), v_fov,
                       lens_shape ? LensShape::Circle : LensShape::Square, aperture, focus_distance);
    h->has_camera = true;
    return 0;
}
int hnmh_builder_add_sphere(void* b, void* assets, const double* center, double radius, const hnmh_material* m) {
    HNMH_TRY
    ((BuilderHandle*)b)->scene.add(std::make_unique<Sphere>(Vector3(center[0], center[1], center[2]), radius, make_material((AssetStore*)assets, m)));
    return 0;
    HNMH_CATCH(-1)
}
int hnmh_builder_add_cuboid(void* b, void* assets, const double* mn, const double* mx, const hnmh_material* m) {
    HNMH_TRY
    ((BuilderHandle*)b)->scene.add(std::make_unique<Cuboid>(Aabb{Vector3(mn[0], mn[1], mn[2]), Vector3(mx[0], mx[1], mx[2])}, make_material((AssetStore*)assets, m)));
    return 0;
    HNMH_CATCH(-1)
}
// world-space triangle soup; the BVH is built by the reference's algorithm
int hnmh_builder_add_mesh(void* b, void* assets, const double* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces,
                          const hnmh_material* m) {
    HNMH_TRY
    Mesh mesh;
    mesh.material = make_material((AssetStore*)assets, m);
    for (uint32_t i = 0; i < nverts; i++) mesh.vertexes.push_back(Vector3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
    for (uint32_t i = 0; i < nfaces; i++) {
        if (faces[3 * i] >= nverts || faces[3 * i + 1] >= nverts || faces[3 * i + 2] >= nverts) throw std::runtime_error("face index out of range");
        mesh.faces.push_back(Face{faces[3 * i], faces[3 * i + 1], faces[3 * i + 2]});
    }
    ((BuilderHandle*)b)->scene.add(BvhMesh::from_mesh(std::move(mesh)));
    return 0;
    HNMH_CATCH(-1)
}
// OBJ from the store with a row-major 4x4 matrix (src/loader.rs:12)
int hnmh_builder_add_obj(void* b, void* assets, const char* path, const double* m44, const hnmh_material* m) {
    HNMH_TRY
    Matrix44 mat;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) mat.e[i][j] = m44[4 * i + j];
    ((BuilderHandle*)b)->scene.add(BvhMesh::from_mesh(ObjLoader::load(*(AssetStore*)assets, path, mat, make_material((AssetStore*)assets, m))));
    return 0;
    HNMH_CATCH(-1)
}
// six face images (asset names, px nx py ny pz nz) and the intensity
int hnmh_builder_skybox(void* b, void* assets, const char* const* faces, const double* intensity) {
    HNMH_TRY
    auto* a = (AssetStore*)assets;
    Skybox& s = ((BuilderHandle*)b)->scene.skybox;
    s.px = a->image(faces[0]); s.nx = a->image(faces[1]); s.py = a->image(faces[2]);
    s.ny = a->image(faces[3]); s.pz = a->image(faces[4]); s.nz = a->image(faces[5]);
    s.intensity = Vector3(intensity[0], intensity[1], intensity[2]);
    return 0;
    HNMH_CATCH(-1)
}
// consumes the builder's scene
void* hnmh_builder_finish(void* b) {
    HNMH_TRY
    auto* bh = (BuilderHandle*)b;
    if (!bh->has_camera) throw std::runtime_error("builder: camera not set");
    if (!bh->scene.skybox.px) throw std::runtime_error("builder: skybox not set");
    auto h = new SceneHandle();
    h->camera = bh->camera.abi();
    h->bvh_scene = BvhScene::from_scene(std::move(bh->scene));
    bh->scene = Scene();
    return h;
    HNMH_CATCH(nullptr)
}

}  // extern "C"
