/*
 * hanamaru_b200.h -- C ABI of the B200-native radiance-loop core.
 *
 * This is the drop-in boundary underneath hanamaru-renderer's
 * `Renderer::render` (reference src/renderer.rs:25-46) and
 * `Renderer::update_imgbuf` (src/renderer.rs:64-90).  The reference has no
 * FFI of its own: the host (Rust there, the C++ mirror
 * under hanamaru_renderer_b200/csrc/host here) keeps OBJ / texture loading,
 * BVH build (src/bvh.rs:107-211), the CLI and PNG output, flattens its scene
 * into the POD description below and calls these entry points.
 *
 * Conventions
 *   - plain pointers and sizes only; every input array stays owned by the
 *     caller and is deep-copied by hnm_scene_create;
 *   - every function returns 0 on success or a negative hnm_status; the text
 *     of the last error on the calling thread is hnm_last_error();
 *   - nothing aborts or throws across this boundary;
 *   - all floating-point scene data is f64 exactly as in the reference
 *     (`Vector3 {x,y,z: f64}`, src/vector.rs:6-12, #[repr(C)]);
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point returns HNM_ERR_CUDA.
 */
#ifndef HANAMARU_B200_H
#define HANAMARU_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HNM_ABI_VERSION 1u

typedef enum hnm_status {
    HNM_OK = 0,
    HNM_ERR_INVALID = -1,  /* bad argument / malformed scene description */
    HNM_ERR_CUDA = -2,     /* CUDA runtime failure (message has the detail) */
    HNM_ERR_NOMEM = -3,
    HNM_ERR_STATE = -4     /* call order (e.g. resolve before any pass) */
} hnm_status;

/* src/vector.rs:6-12 */
typedef struct hnm_vec3 { double x, y, z; } hnm_vec3;

/* src/camera.rs:7-28 -- the host runs Camera::new (src/camera.rs:45-64) */
typedef struct hnm_camera {
    hnm_vec3 eye;
    hnm_vec3 right, up, forward;
    hnm_vec3 plane_half_right, plane_half_up;
    double lens_radius;
    double focus_distance;
    int32_t lens_shape; /* src/camera.rs:32-35: 0 = Square, 1 = Circle */
    int32_t _pad;
} hnm_camera;

/* src/material.rs:9-15 */
enum {
    HNM_SURFACE_DIFFUSE = 0,
    HNM_SURFACE_SPECULAR = 1,
    HNM_SURFACE_REFRACTION = 2,    /* param = refractive_index */
    HNM_SURFACE_GGX = 3,           /* param = f0 */
    HNM_SURFACE_GGX_REFRACTION = 4 /* param = refractive_index */
};

/* src/texture.rs:72-76: optional image, multiplied by a tint colour */
typedef struct hnm_texture {
    hnm_vec3 color;
    int32_t image; /* index into hnm_scene_desc.images, or -1 */
    int32_t _pad;
} hnm_texture;

/* src/material.rs:17-23 */
typedef struct hnm_material {
    hnm_texture albedo, emission, roughness;
    double param;
    int32_t surface;
    int32_t _pad;
} hnm_material;

/* Decoded texels, RGBA8, row 0 = top row, i.e. exactly what
 * `DynamicImage::get_pixel(x, y)` returns (src/texture.rs:59-63,
 * src/color.rs:18-24).  Decoding (PNG/JPEG) stays on the host. */
typedef struct hnm_image {
    const uint8_t* rgba;
    uint32_t width, height;
} hnm_image;

/* src/scene.rs Intersectable impls that `main::render` can reach */
enum {
    HNM_ELEM_SPHERE = 0, /* src/scene.rs:51-102  : a = center, radius      */
    HNM_ELEM_CUBOID = 1, /* src/scene.rs:146-194 : a = aabb.min, b = aabb.max */
    HNM_ELEM_MESH = 2    /* src/scene.rs:236-266 : BvhMesh, `mesh` indexes meshes[] */
};

typedef struct hnm_element {
    hnm_vec3 a, b;
    double radius;
    int32_t kind;
    int32_t material; /* index into materials[] */
    int32_t mesh;     /* index into meshes[] (HNM_ELEM_MESH), else -1 */
    int32_t _pad;
} hnm_element;

/* One node of a host-built BVH (src/bvh.rs:68-77), flattened in DFS
 * pre-order: the first child of node i is always node i+1.  A leaf has
 * child0 == child1 == -1 and owns index list entries [first, first+count).
 * The order of nodes and of the index list IS the reference's visiting order
 * (src/bvh.rs:213-263); the core uses it only to break exact distance ties
 * the way the reference does. */
typedef struct hnm_bvh_node {
    double aabb_min[3];
    double aabb_max[3];
    int32_t child0, child1;
    uint32_t first, count;
} hnm_bvh_node;

/* src/scene.rs:196-206 + 236-239.  Vertices are world space (the matrix is
 * baked in by the OBJ loader, src/loader.rs:31).  Face vertex indices are
 * relative to vertex_offset.  Nodes index `mesh_indices` relative to
 * index_offset; each entry there is a face index relative to face_offset. */
typedef struct hnm_mesh {
    uint32_t vertex_offset, vertex_count;
    uint32_t face_offset, face_count;
    uint32_t node_offset, node_count; /* node_offset = root */
    uint32_t index_offset, index_count;
} hnm_mesh;

/* src/config.rs:4-25.  Passed across the ABI, not compiled in. */
typedef struct hnm_config {
    double eps;               /* EPS    = 1e-4  */
    double offset;            /* OFFSET = 1e-4  */
    double inf;               /* INF    = 1e100 */
    double gamma_factor;      /* 2.2 */
    double tone_exposure;     /* 1.5 */
    double tone_white_point;  /* 20.0 */
    double bilateral_sigma_i; /* 1.0 */
    double bilateral_sigma_s; /* 16.0 */
    uint32_t supersampling;        /* 2 */
    uint32_t bounce_limit;         /* 10 -> `for _ in 1..10`, 9 segments; accepted range 2 .. 17 */
    uint32_t tone_mapping_mode;    /* 0 None, 1 Reinhard */
    uint32_t bilateral_iteration;  /* 1 */
    uint32_t bilateral_diameter;   /* 3 */
    uint32_t _pad;
} hnm_config;

typedef struct hnm_scene_desc {
    uint32_t abi_version; /* HNM_ABI_VERSION */
    uint32_t _pad0;

    const hnm_element* elements;   uint32_t num_elements;   uint32_t _pad1;
    const hnm_material* materials; uint32_t num_materials;  uint32_t _pad2;
    const hnm_image* images;       uint32_t num_images;     uint32_t _pad3;
    const hnm_mesh* meshes;        uint32_t num_meshes;     uint32_t _pad4;

    const double* vertices;        uint32_t num_vertices;   uint32_t _pad5; /* xyz triples */
    const uint32_t* faces;         uint32_t num_faces;      uint32_t _pad6; /* v0 v1 v2 triples */
    const hnm_bvh_node* mesh_nodes; uint32_t num_mesh_nodes; uint32_t _pad7;
    const uint32_t* mesh_indices;  uint32_t num_mesh_indices; uint32_t _pad8;

    /* BvhScene (src/scene.rs:379-416): nodes over elements, root = 0 */
    const hnm_bvh_node* top_nodes; uint32_t num_top_nodes;  uint32_t _pad9;
    const uint32_t* top_indices;   uint32_t num_top_indices; uint32_t _pad10; /* element ids */

    /* Skybox (src/scene.rs:268-320): image ids px nx py ny pz nz */
    int32_t skybox_images[6];
    hnm_vec3 skybox_intensity;

    /* `scene.emissions()` (src/scene.rs:356-358), in element order */
    const uint32_t* emissions;     uint32_t num_emissions;  uint32_t _pad11;

    hnm_config config;
} hnm_scene_desc;

/* renderer kinds: PathTracingRenderer (src/renderer.rs:148-203) and the four
 * DebugRenderer modes (src/renderer.rs:102-139) */
enum {
    HNM_MODE_PATHTRACING = 0,
    HNM_MODE_DEBUG_SHADING = 1,
    HNM_MODE_DEBUG_NORMAL = 2,
    HNM_MODE_DEBUG_DEPTH = 3,
    HNM_MODE_DEBUG_FOCALPLANE = 4
};

/* Which image rows this renderer owns (multi-GPU sharding).
 * Rows are grouped in tiles of `tile_rows`; tile k belongs to
 * rank k % num_ranks.  {0,1,any} = the whole image. */
typedef struct hnm_shard {
    uint32_t rank, num_ranks, tile_rows, _pad;
} hnm_shard;

typedef struct hnm_counters {
    uint64_t paths;        /* camera paths started (= samples) */
    uint64_t segments;     /* closest-hit rays along camera paths */
    uint64_t shadow_rays;  /* NEE rays (also closest-hit, src/renderer.rs:280) */
    uint64_t rng_fallbacks;/* paths whose lens loop outran the stored ISAAC tail */
    uint64_t kernel_launches;
    uint64_t node_visits;  /* BVH nodes fetched / primitives tested; counted only when */
    uint64_t prim_tests;   /* the environment has HNM_TRACE_STATS=1 (instrumented kernel) */
    uint64_t cand_overflows; /* rays whose candidate list overflowed or that were routed to the exact traversal (HNM_TRACE_STATS=1) */
} hnm_counters;

typedef struct hnm_scene hnm_scene;       /* opaque: device copy of a scene */
typedef struct hnm_renderer hnm_renderer; /* opaque: wavefront state + accumulation buffer */
typedef struct hnm_group hnm_group;       /* opaque: one scene copy + one renderer per device of ONE process */

const char* hnm_last_error(void);
uint32_t hnm_abi_version(void);
/* number of visible CUDA devices, or a negative hnm_status */
int hnm_device_count(void);

/* Deep-copies the description onto `device` (BVH re-laid out for the GPU,
 * textures into CUDA arrays / texture objects). */
int hnm_scene_create(const hnm_scene_desc* desc, int device, hnm_scene** out);
void hnm_scene_destroy(hnm_scene* scene);

/* One renderer = one `Renderer::render` call in the reference: it owns the
 * f64 accumulation buffer (src/renderer.rs:28) for its shard of the image.
 * `max_batch` = how many passes may be in flight in one wavefront (0 = auto). */
int hnm_renderer_create(hnm_scene* scene, const hnm_camera* camera,
                        uint32_t width, uint32_t height, int mode,
                        const hnm_shard* shard /* NULL = whole image */,
                        uint32_t max_batch, hnm_renderer** out);
void hnm_renderer_destroy(hnm_renderer* r);

/* Runs passes sampling_first .. sampling_first+count-1 of the pass loop
 * (src/renderer.rs:32-38; `sampling` is 1-origin and is part of every path's
 * RNG seed, src/renderer.rs:167) and adds them to the accumulation buffer in
 * pass order.  Asynchronous: returns once the work is enqueued.
 * The call is cut into equal-sized batches of at most `max_batch` passes.  The
 * random streams and camera rays of the batch after the current one are
 * generated ahead of time on a second stream; after the last batch of a call
 * that is a guess (an identical call continuing the pass numbering, which is
 * what the reference's pass loop does).  A wrong guess is discarded: results
 * never depend on it.  HNM_RNG_OVERLAP=0 / HNM_RNG_SPECULATE=0 switch it off. */
int hnm_render_passes(hnm_renderer* r, uint32_t sampling_first, uint32_t count);
int hnm_synchronize(hnm_renderer* r);
int hnm_clear(hnm_renderer* r);

/* Number of pixels this renderer owns, and their layout: local row lr maps to
 * image row hnm_local_row_to_global(r, lr). */
uint32_t hnm_owned_rows(const hnm_renderer* r);
uint32_t hnm_local_row_to_global(const hnm_renderer* r, uint32_t local_row);

/* Copies the owned part of the accumulation buffer (f64 rgb triples,
 * row-major, local row order) to host memory: owned_rows*width*3 doubles. */
int hnm_read_accum(hnm_renderer* r, double* rgb);
/* Device address of the same buffer, for the multi-GPU gather (plumbing by
 * the caller: torch.distributed / NCCL). */
int hnm_accum_device_ptr(hnm_renderer* r, void** ptr, size_t* bytes);

/* `update_imgbuf` (src/renderer.rs:64-90): scale by 1/(sampling*ss*ss),
 * Reinhard, gamma, bilateral, quantise.  `accum_full_device` is a DEVICE
 * pointer to a full-image f64 rgb buffer in image row order (for a
 * single-shard renderer pass NULL to use its own buffer).  rgb8 is HOST
 * memory, width*height*3 bytes, row 0 = top. */
int hnm_resolve(hnm_renderer* r, const void* accum_full_device, uint32_t sampling, uint8_t* rgb8);
/* The same in two halves: _begin enqueues update_imgbuf and the copy into a
 * pinned buffer behind the passes enqueued so far and returns at once; _end
 * waits for exactly that work and hands the image out.  A host that reports
 * progress every interval (src/renderer.rs:216-226) enqueues the next passes
 * between the two, so the device never waits for the host's copy.  One image
 * may be pending per renderer. */
int hnm_resolve_begin(hnm_renderer* r, const void* accum_full_device, uint32_t sampling);
int hnm_resolve_end(hnm_renderer* r, uint8_t* rgb8);
/* Scatter gathered per-rank shards ([num_ranks][owned_rows*width*3] f64 on
 * the device) into image row order. */
int hnm_deinterleave(hnm_renderer* r, const void* gathered_device, void* full_device);

/* Precision of the shading kernels.  EXACT (default): every operation is the reference's f64 sequence, transcendental
 * functions from the deterministic library -- bit parity with the oracle.  FAST_MATH (opt-in): pow / sincos / acos of the
 * shading kernels in hardware f32; traversal, exact hit tests, RNG and accumulation are unchanged.  Results are then
 * statistically equal (tests: PSNR >= 45 dB and mean bias < 0.5 level against the oracle at equal spp and seed). */
enum { HNM_PRECISION_EXACT = 0, HNM_PRECISION_FAST_MATH = 1 };
int hnm_set_precision(hnm_renderer* r, int precision);

int hnm_get_counters(hnm_renderer* r, hnm_counters* out);
/* ms of device time spent in the top kernels since the last reset (CUDA
 * events on the renderer's stream); names are static strings. */
int hnm_get_kernel_times(hnm_renderer* r, uint32_t max, const char** names, float* ms, uint32_t* launches, uint32_t* n);
int hnm_set_profiling(hnm_renderer* r, int enabled);
/* Diagnostics (environment HNM_WID_STATS=1, else all zero): bit w of masks[k] = a warp of kernel class k
 * (0 generation, 1 trace, 2 shade) ran in hardware warp slot w of its SM. */
int hnm_debug_warp_slots(hnm_renderer* r, uint64_t* masks, uint32_t n);
/* Diagnostics: the queue counters of the last batch, 16 words per bounce (rays, misses, delta hits, NEE hits, NEE events,
 * shadow rays, work counters). */
int hnm_debug_read_counters(hnm_renderer* r, uint32_t* out, uint32_t n);
/* Device-side stopwatch on the renderer's own stream (torch.cuda.Event only sees torch's stream):
 * hnm_mark records CUDA event `slot` (0..15); hnm_elapsed_ms synchronises on both and returns b - a. */
int hnm_mark(hnm_renderer* r, uint32_t slot);
int hnm_elapsed_ms(hnm_renderer* r, uint32_t slot_a, uint32_t slot_b, float* ms);

/* ---- multi-GPU ---------------------------------------
 * `Renderer::render` is ONE call in ONE process (src/renderer.rs:25, called
 * from src/main.rs:1216).  Every (pixel, sub-pixel, pass) is independent, so
 * the image is sharded by interleaved row tiles (hnm_shard) with no data-path
 * collective; the only exchange is the gather of the f64 accumulation shards
 * before `update_imgbuf` (the 3x3 bilateral filter reads neighbouring rows).
 * Each pixel has exactly one owner and passes are added in order, so the
 * gathered buffer is bit-identical to a single-GPU render.
 *
 * (a) hnm_group_*: one process, one host thread, N devices -- what a
 *     single-process host calls INSTEAD of hnm_scene_create / hnm_renderer_*.
 *     The description is validated and re-laid out once and uploaded to every
 *     device; passes are enqueued asynchronously on all devices; the shards
 *     travel to device 0 as peer copies over NVLink and update_imgbuf runs on
 *     device 0 only.  `devices` may name the same device more than once
 *     (testing on a single GPU).  tile_rows 0 = default (4). */
int hnm_group_create(const hnm_scene_desc* desc, const hnm_camera* camera,
                     uint32_t width, uint32_t height, int mode,
                     uint32_t num_devices, const int* devices,
                     uint32_t tile_rows, uint32_t max_batch, hnm_group** out);
void hnm_group_destroy(hnm_group* g);
uint32_t hnm_group_size(const hnm_group* g);
/* member k's renderer (counters, timing marks, profiling); owned by the group */
hnm_renderer* hnm_group_member(hnm_group* g, uint32_t k);
int hnm_group_render_passes(hnm_group* g, uint32_t sampling_first, uint32_t count);
int hnm_group_synchronize(hnm_group* g);
int hnm_group_clear(hnm_group* g);
/* gather + `update_imgbuf` -> host rgb8 (width*height*3) */
int hnm_group_resolve(hnm_group* g, uint32_t sampling, uint8_t* rgb8);
/* gather -> the full accumulation buffer in image row order (width*height*3 f64, host) */
int hnm_group_read_accum(hnm_group* g, double* rgb);
int hnm_group_get_counters(hnm_group* g, hnm_counters* out); /* summed over the members */

/* (b) hnm_dist_*: one process per device (torchrun / MPI launchers).  The
 *     exchange step is ONE ncclAllGather of the accumulation shards on the
 *     renderer's own stream (libnccl.so.2 is bound at run time; without it
 *     these calls return HNM_ERR_STATE).  Rank 0 makes the id, the launcher
 *     distributes its HNM_DIST_ID_BYTES bytes (any side channel), every rank
 *     calls hnm_dist_init on a renderer created with the matching hnm_shard. */
#define HNM_DIST_ID_BYTES 128u
int hnm_dist_unique_id(uint8_t* id);
int hnm_dist_init(hnm_renderer* r, const uint8_t* id, uint32_t rank, uint32_t num_ranks);
/* A communicator that outlives renderers: the reference's `render` is called once per
 * image (src/main.rs:1216) and a renderer lives for one call, but ncclCommInitRank
 * costs seconds at 8 ranks.  A host process creates ONE hnm_comm after its
 * rendezvous and attaches it to every renderer it makes (instead of
 * hnm_dist_init); the renderer uses it and does not destroy it. */
typedef struct hnm_comm hnm_comm;
int hnm_comm_create(int device, const uint8_t* id, uint32_t rank, uint32_t num_ranks, hnm_comm** out);
void hnm_comm_destroy(hnm_comm* comm);
int hnm_dist_attach(hnm_renderer* r, hnm_comm* comm);
/* Collective (every rank calls it, same order).  rgb8 == NULL: take part in
 * the gather only, asynchronously.  rgb8 != NULL: also run `update_imgbuf`
 * on the gathered image and return it (host, width*height*3). */
int hnm_dist_resolve(hnm_renderer* r, uint32_t sampling, uint8_t* rgb8);
/* Collective and asynchronous: the gather is enqueued on every rank; a rank
 * with want_image != 0 also enqueues update_imgbuf and collects the image
 * later with hnm_resolve_end. */
int hnm_dist_resolve_begin(hnm_renderer* r, uint32_t sampling, int want_image);
/* Collective; rgb != NULL receives the gathered f64 buffer in image row order. */
int hnm_dist_read_accum(hnm_renderer* r, double* rgb);

/* ---- batch entry points (per-function parity) ---------- */

typedef struct hnm_ray { hnm_vec3 origin, direction; } hnm_ray;

/* `BvhScene::intersect` result (src/scene.rs:385-401, Intersection at
 * src/scene.rs:10-17) */
typedef struct hnm_hit {
    hnm_vec3 position;
    hnm_vec3 normal;
    hnm_vec3 albedo;
    hnm_vec3 emission;
    double distance;
    double u, v;
    double roughness;
    double param;
    int32_t hit;     /* 0 / 1 */
    int32_t element; /* element id, -1 on miss */
    int32_t face;    /* face index inside the mesh, -1 otherwise */
    int32_t surface;
} hnm_hit;

/* n closest-hit queries through the full scene (host arrays). */
int hnm_intersect_batch(hnm_scene* scene, const hnm_ray* rays, uint32_t n, hnm_hit* hits);

/* rand 0.4 StdRng (ISAAC-64) seeded with `seeds[4*i..4*i+4]`
 * (src/renderer.rs:165-168): writes the first `count` u64 outputs of each
 * stream (count <= HNM_RNG_TAIL) to out[i*count ..]. */
#define HNM_RNG_TAIL 32u
int hnm_isaac64_batch(int device, const uint64_t* seeds, uint32_t n, uint32_t count, uint64_t* out);

/* `PointMaterial::sample` (src/material.rs:91-151) on n independent inputs:
 * in  = [surface, param, roughness, r0, r1, px,py,pz, vx,vy,vz, nx,ny,nz] (14 f64)
 * out = [some, ox,oy,oz, dx,dy,dz, reflectance] (8 f64) */
int hnm_material_sample_batch(int device, const double* in, uint32_t n, double* out);
/* `PointMaterial::bsdf` (src/material.rs:53-89), Diffuse / GGX only:
 * in = [surface, param, roughness, vx,vy,vz, nx,ny,nz, lx,ly,lz] (12 f64), out = 1 f64 */
int hnm_material_bsdf_batch(int device, const double* in, uint32_t n, double* out);
/* deterministic libm used by the device code (sin, cos, exp, pow, acos):
 * fn 0 sin, 1 cos, 2 exp, 3 pow(x,y), 4 acos ; y ignored unless pow */
int hnm_math_batch(int device, int fn, const double* x, const double* y, uint32_t n, double* out);

/* `Texture::sample` (src/texture.rs:29-63,108-114) of image `image` of the
 * scene (-1: constant colour) times `tint[3]` at n (u, v) pairs -> n rgb triples */
int hnm_texture_sample_batch(hnm_scene* scene, int32_t image, const double* tint, const double* uv, uint32_t n, double* rgb);
/* `Skybox::sample` (src/scene.rs:295-319) for n directions (xyz) -> n rgb triples */
int hnm_skybox_sample_batch(hnm_scene* scene, const double* directions, uint32_t n, double* rgb);

#ifdef __cplusplus
}
#endif
#endif /* HANAMARU_B200_H */
