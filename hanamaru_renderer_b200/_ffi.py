"""ctypes mirror of include/hanamaru_b200.h and of the host facade (csrc/host/host_capi.cpp).

No torch, no numpy requirements beyond array plumbing: this is exactly the
binding a Rust `extern "C"` block would declare.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
CORE_LIB = os.environ.get("HNM_CORE_LIB") or os.path.join(HERE, "libhanamaru_b200.so")  # override: tuning experiments only
HOST_LIB = os.path.join(HERE, "libhanamaru_host.so")


def _point_at_the_bundled_nccl():
    """In a Python process torch loads the libnccl.so.2 of the nvidia-nccl wheel; the core binds NCCL with dlopen at its first
    hnm_dist_* / hnm_comm_* call.  Two different libnccl.so.2 in one process clash by soname, so the core is told to load the
    same file (HNM_NCCL_LIB, read by the C side; a Rust host simply links or sets its own).  No import of torch here."""
    if os.environ.get("HNM_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["HNM_NCCL_LIB"] = cand
                return
    except Exception:
        pass


_point_at_the_bundled_nccl()

HNM_ABI_VERSION = 1
HNM_RNG_TAIL = 32

SURFACE_DIFFUSE, SURFACE_SPECULAR, SURFACE_REFRACTION, SURFACE_GGX, SURFACE_GGX_REFRACTION = range(5)
ELEM_SPHERE, ELEM_CUBOID, ELEM_MESH = range(3)
MODE_PATHTRACING, MODE_DEBUG_SHADING, MODE_DEBUG_NORMAL, MODE_DEBUG_DEPTH, MODE_DEBUG_FOCALPLANE = range(5)


class Vec3(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]

    def tuple(self):
        return (self.x, self.y, self.z)


class Camera(C.Structure):
    _fields_ = [("eye", Vec3), ("right", Vec3), ("up", Vec3), ("forward", Vec3),
                ("plane_half_right", Vec3), ("plane_half_up", Vec3),
                ("lens_radius", C.c_double), ("focus_distance", C.c_double),
                ("lens_shape", C.c_int32), ("_pad", C.c_int32)]


class Texture(C.Structure):
    _fields_ = [("color", Vec3), ("image", C.c_int32), ("_pad", C.c_int32)]


class Material(C.Structure):
    _fields_ = [("albedo", Texture), ("emission", Texture), ("roughness", Texture),
                ("param", C.c_double), ("surface", C.c_int32), ("_pad", C.c_int32)]


class Image(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class Element(C.Structure):
    _fields_ = [("a", Vec3), ("b", Vec3), ("radius", C.c_double), ("kind", C.c_int32),
                ("material", C.c_int32), ("mesh", C.c_int32), ("_pad", C.c_int32)]


class BvhNode(C.Structure):
    _fields_ = [("aabb_min", C.c_double * 3), ("aabb_max", C.c_double * 3),
                ("child0", C.c_int32), ("child1", C.c_int32), ("first", C.c_uint32), ("count", C.c_uint32)]


class Mesh(C.Structure):
    _fields_ = [("vertex_offset", C.c_uint32), ("vertex_count", C.c_uint32),
                ("face_offset", C.c_uint32), ("face_count", C.c_uint32),
                ("node_offset", C.c_uint32), ("node_count", C.c_uint32),
                ("index_offset", C.c_uint32), ("index_count", C.c_uint32)]


class Config(C.Structure):
    _fields_ = [("eps", C.c_double), ("offset", C.c_double), ("inf", C.c_double), ("gamma_factor", C.c_double),
                ("tone_exposure", C.c_double), ("tone_white_point", C.c_double),
                ("bilateral_sigma_i", C.c_double), ("bilateral_sigma_s", C.c_double),
                ("supersampling", C.c_uint32), ("bounce_limit", C.c_uint32), ("tone_mapping_mode", C.c_uint32),
                ("bilateral_iteration", C.c_uint32), ("bilateral_diameter", C.c_uint32), ("_pad", C.c_uint32)]


def _arr(name, typ):
    return [(name, C.POINTER(typ)), ("num_" + name, C.c_uint32), ("_pad_" + name, C.c_uint32)]


class SceneDesc(C.Structure):
    _fields_ = ([("abi_version", C.c_uint32), ("_pad0", C.c_uint32)]
                + _arr("elements", Element) + _arr("materials", Material) + _arr("images", Image) + _arr("meshes", Mesh)
                + _arr("vertices", C.c_double) + _arr("faces", C.c_uint32)
                + _arr("mesh_nodes", BvhNode) + _arr("mesh_indices", C.c_uint32)
                + _arr("top_nodes", BvhNode) + _arr("top_indices", C.c_uint32)
                + [("skybox_images", C.c_int32 * 6), ("skybox_intensity", Vec3)]
                + _arr("emissions", C.c_uint32)
                + [("config", Config)])


class Shard(C.Structure):
    _fields_ = [("rank", C.c_uint32), ("num_ranks", C.c_uint32), ("tile_rows", C.c_uint32), ("_pad", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("rng_fallbacks", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("node_visits", C.c_uint64), ("prim_tests", C.c_uint64), ("cand_overflows", C.c_uint64)]


class Ray(C.Structure):
    _fields_ = [("origin", Vec3), ("direction", Vec3)]


class Hit(C.Structure):
    _fields_ = [("position", Vec3), ("normal", Vec3), ("albedo", Vec3), ("emission", Vec3),
                ("distance", C.c_double), ("u", C.c_double), ("v", C.c_double),
                ("roughness", C.c_double), ("param", C.c_double),
                ("hit", C.c_int32), ("element", C.c_int32), ("face", C.c_int32), ("surface", C.c_int32)]


class HostMaterial(C.Structure):  # hnmh_material
    _fields_ = [("surface", C.c_int32), ("_pad", C.c_int32), ("param", C.c_double),
                ("albedo", C.c_double * 3), ("emission", C.c_double * 3), ("roughness", C.c_double * 3),
                ("albedo_image", C.c_char_p), ("emission_image", C.c_char_p), ("roughness_image", C.c_char_p)]


# every symbol include/hanamaru_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
CORE_SYMBOLS = {
    "hnm_last_error": (C.c_char_p, []),
    "hnm_abi_version": (C.c_uint32, []),
    "hnm_device_count": (C.c_int, []),
    "hnm_scene_create": (C.c_int, [C.POINTER(SceneDesc), C.c_int, C.POINTER(_P)]),
    "hnm_scene_destroy": (None, [_P]),
    "hnm_renderer_create": (C.c_int, [_P, C.POINTER(Camera), C.c_uint32, C.c_uint32, C.c_int, C.POINTER(Shard), C.c_uint32, C.POINTER(_P)]),
    "hnm_renderer_destroy": (None, [_P]),
    "hnm_render_passes": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "hnm_synchronize": (C.c_int, [_P]),
    "hnm_clear": (C.c_int, [_P]),
    "hnm_owned_rows": (C.c_uint32, [_P]),
    "hnm_local_row_to_global": (C.c_uint32, [_P, C.c_uint32]),
    "hnm_read_accum": (C.c_int, [_P, _P]),
    "hnm_accum_device_ptr": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "hnm_resolve": (C.c_int, [_P, _P, C.c_uint32, _P]),
    "hnm_resolve_begin": (C.c_int, [_P, _P, C.c_uint32]),
    "hnm_resolve_end": (C.c_int, [_P, _P]),
    "hnm_deinterleave": (C.c_int, [_P, _P, _P]),
    "hnm_get_counters": (C.c_int, [_P, C.POINTER(Counters)]),
    "hnm_get_kernel_times": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "hnm_set_profiling": (C.c_int, [_P, C.c_int]),
    "hnm_set_precision": (C.c_int, [_P, C.c_int]),
    "hnm_debug_warp_slots": (C.c_int, [_P, C.POINTER(C.c_uint64), C.c_uint32]),
    "hnm_debug_read_counters": (C.c_int, [_P, C.POINTER(C.c_uint32), C.c_uint32]),
    "hnm_mark": (C.c_int, [_P, C.c_uint32]),
    "hnm_elapsed_ms": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]),
    "hnm_intersect_batch": (C.c_int, [_P, _P, C.c_uint32, _P]),
    "hnm_isaac64_batch": (C.c_int, [C.c_int, _P, C.c_uint32, C.c_uint32, _P]),
    "hnm_material_sample_batch": (C.c_int, [C.c_int, _P, C.c_uint32, _P]),
    "hnm_material_bsdf_batch": (C.c_int, [C.c_int, _P, C.c_uint32, _P]),
    "hnm_math_batch": (C.c_int, [C.c_int, C.c_int, _P, _P, C.c_uint32, _P]),
    "hnm_texture_sample_batch": (C.c_int, [_P, C.c_int32, _P, _P, C.c_uint32, _P]),
    "hnm_skybox_sample_batch": (C.c_int, [_P, _P, C.c_uint32, _P]),
    "hnm_group_create": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(Camera), C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.POINTER(C.c_int),
                                   C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "hnm_group_destroy": (None, [_P]),
    "hnm_group_size": (C.c_uint32, [_P]),
    "hnm_group_member": (_P, [_P, C.c_uint32]),
    "hnm_group_render_passes": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "hnm_group_synchronize": (C.c_int, [_P]),
    "hnm_group_clear": (C.c_int, [_P]),
    "hnm_group_resolve": (C.c_int, [_P, C.c_uint32, _P]),
    "hnm_group_read_accum": (C.c_int, [_P, _P]),
    "hnm_group_get_counters": (C.c_int, [_P, C.POINTER(Counters)]),
    "hnm_dist_unique_id": (C.c_int, [_P]),
    "hnm_dist_init": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32]),
    "hnm_comm_create": (C.c_int, [C.c_int, _P, C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "hnm_comm_destroy": (None, [_P]),
    "hnm_dist_attach": (C.c_int, [_P, _P]),
    "hnm_dist_resolve_begin": (C.c_int, [_P, C.c_uint32, C.c_int]),
    "hnm_dist_resolve": (C.c_int, [_P, C.c_uint32, _P]),
    "hnm_dist_read_accum": (C.c_int, [_P, _P]),
}
HNM_DIST_ID_BYTES = 128
PRECISION_EXACT, PRECISION_FAST_MATH = 0, 1

HOST_SYMBOLS = {
    "hnmh_last_error": (C.c_char_p, []),
    "hnmh_assets_create": (_P, []),
    "hnmh_assets_destroy": (None, [_P]),
    "hnmh_assets_set_root": (C.c_int, [_P, C.c_char_p]),
    "hnmh_assets_load_pack": (C.c_int, [_P, C.c_char_p]),
    "hnmh_assets_put_image": (C.c_int, [_P, C.c_char_p, _P, C.c_uint32, C.c_uint32]),
    "hnmh_assets_put_obj_text": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_size_t]),
    "hnmh_assets_obj_counts": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "hnmh_assets_obj_copy": (C.c_int, [_P, C.c_char_p, _P, _P]),
    "hnmh_scene_asset_paths": (C.c_int, [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]),
    "hnmh_scene_build": (_P, [_P, C.c_char_p]),
    "hnmh_scene_desc": (C.POINTER(SceneDesc), [_P]),
    "hnmh_scene_camera": (C.POINTER(Camera), [_P]),
    "hnmh_scene_destroy": (None, [_P]),
    "hnmh_render": (C.c_int, [_P, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_uint32, C.c_int, _P, C.POINTER(C.c_uint32)]),
    "hnmh_image_decode": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), _P]),
    "hnmh_save_png": (C.c_int, [C.c_char_p, _P, C.c_uint32, C.c_uint32]),
    "hnmh_render_to_files": (C.c_int, [_P, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_uint32, C.c_int, C.c_char_p, _P,
                                       C.POINTER(C.c_uint32)]),
    "hnmh_stdrng": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_double, C.c_double, _P]),
    "hnmh_builder_create": (_P, []),
    "hnmh_builder_destroy": (None, [_P]),
    "hnmh_builder_camera": (C.c_int, [_P, _P, _P, _P, C.c_double, C.c_int, C.c_double, C.c_double]),
    "hnmh_builder_add_sphere": (C.c_int, [_P, _P, _P, C.c_double, C.POINTER(HostMaterial)]),
    "hnmh_builder_add_cuboid": (C.c_int, [_P, _P, _P, _P, C.POINTER(HostMaterial)]),
    "hnmh_builder_add_mesh": (C.c_int, [_P, _P, _P, C.c_uint32, _P, C.c_uint32, C.POINTER(HostMaterial)]),
    "hnmh_builder_add_obj": (C.c_int, [_P, _P, C.c_char_p, _P, C.POINTER(HostMaterial)]),
    "hnmh_builder_skybox": (C.c_int, [_P, _P, C.POINTER(C.c_char_p), _P]),
    "hnmh_builder_finish": (_P, [_P]),
}


def _bind(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)  # AttributeError = missing export: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


_core = None
_host = None


class CoreUnavailable(RuntimeError):
    pass


def core():
    """libhanamaru_b200.so (CUDA).  There is no fallback: missing library = error."""
    global _core
    if _core is None:
        if not os.path.exists(CORE_LIB):
            raise CoreUnavailable("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % CORE_LIB)
        _core = _bind(C.CDLL(CORE_LIB), CORE_SYMBOLS)
        if _core.hnm_abi_version() != HNM_ABI_VERSION:
            raise CoreUnavailable("ABI version mismatch")
    return _core


def host():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB):
            raise RuntimeError("%s not built: run __graft_entry__.build()" % HOST_LIB)
        _host = _bind(C.CDLL(HOST_LIB), HOST_SYMBOLS)
    return _host
