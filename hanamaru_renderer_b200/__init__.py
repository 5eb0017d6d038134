"""hanamaru_renderer_b200 -- B200-native radiance-loop core behind hanamaru-renderer's Renderer surface.

Layers (top to bottom):
  Python mirror of the reference's host API (this file): AssetStore, build_scene,
      SceneBuilder, DeviceScene, Renderer / PathTracingRenderer / DebugRenderer
      (src/renderer.rs:20-99, 109-146, 148-267).
  libhanamaru_host.so : C++ mirror of the Rust host (scene authoring, OBJ, BVH build, flattening).
  libhanamaru_b200.so : the C ABI of include/hanamaru_b200.h -- hand-written sm_100a CUDA.
There is no CPU fallback anywhere in this package: compute needs the CUDA library and a GPU.
"""
import ctypes as C
import os
import time

import numpy as np

from . import _ffi
from ._ffi import (MODE_DEBUG_DEPTH, MODE_DEBUG_FOCALPLANE, MODE_DEBUG_NORMAL, MODE_DEBUG_SHADING, MODE_PATHTRACING,  # noqa: F401
                   SURFACE_DIFFUSE, SURFACE_GGX, SURFACE_GGX_REFRACTION, SURFACE_REFRACTION, SURFACE_SPECULAR)

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_PACK = os.path.join(REPO_ROOT, "assets", "hanamaru_assets.hnmpack")
CUBEMAP_PACK = os.path.join(REPO_ROOT, "assets", "hanamaru_cubemaps.hnmpack")  # JPEG files, decoded by the C++ host on first use


class HanamaruError(RuntimeError):
    pass


def _check_host(ok):
    if not ok:
        raise HanamaruError(_ffi.host().hnmh_last_error().decode())


def _check(rc):
    if rc != 0:
        raise HanamaruError("hanamaru_b200 error %d: %s" % (rc, _ffi.core().hnm_last_error().decode()))


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- host side
class AssetStore:
    """OBJ + decoded textures for the scene builders (the reference reads them relative to its cwd)."""

    def __init__(self):
        self._h = C.c_void_p(_ffi.host().hnmh_assets_create())

    def __del__(self):
        try:
            _ffi.host().hnmh_assets_destroy(self._h)
        except Exception:
            pass

    @classmethod
    def from_pack(cls, path=DEFAULT_PACK, extra=(CUBEMAP_PACK,)):
        s = cls()
        _check_host(_ffi.host().hnmh_assets_load_pack(s._h, path.encode()) == 0)
        for p in extra:
            if os.path.exists(p):
                _check_host(_ffi.host().hnmh_assets_load_pack(s._h, p.encode()) == 0)
        return s

    @classmethod
    def from_reference(cls, root, scene_names=("rtcamp6",)):
        """Reads OBJ text from a reference checkout and decodes the images the named scenes need with PIL."""
        from PIL import Image
        s = cls()
        _ffi.host().hnmh_assets_set_root(s._h, root.encode())
        for name in scene_names:
            for p in scene_asset_paths(name, images=True):
                im = np.ascontiguousarray(np.asarray(Image.open(os.path.join(root, p)).convert("RGBA"), dtype=np.uint8))
                s.put_image(p, im)
        return s

    def put_image(self, name, rgba):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        assert rgba.ndim == 3 and rgba.shape[2] == 4
        _check_host(_ffi.host().hnmh_assets_put_image(self._h, name.encode(), _vp(rgba), rgba.shape[1], rgba.shape[0]) == 0)

    def put_obj_text(self, name, text):
        b = text.encode() if isinstance(text, str) else text
        _check_host(_ffi.host().hnmh_assets_put_obj_text(self._h, name.encode(), b, len(b)) == 0)

    def obj_geometry(self, name):
        nv, nf = C.c_uint32(), C.c_uint32()
        _check_host(_ffi.host().hnmh_assets_obj_counts(self._h, name.encode(), C.byref(nv), C.byref(nf)) == 0)
        v = np.empty((nv.value, 3), np.float64)
        f = np.empty((nf.value, 3), np.uint32)
        _check_host(_ffi.host().hnmh_assets_obj_copy(self._h, name.encode(), _vp(v), _vp(f)) == 0)
        return v, f


def scene_asset_paths(name, images):
    buf = C.create_string_buffer(1 << 14)
    _check_host(_ffi.host().hnmh_scene_asset_paths(name.encode(), 1 if images else 0, buf, len(buf)) == 0)
    return [p for p in buf.value.decode().split("\n") if p]


class HostScene:
    """A BvhScene (src/scene.rs:379-416) flattened for the C ABI, plus its camera."""

    def __init__(self, handle, keepalive=None):
        if not handle:
            raise HanamaruError(_ffi.host().hnmh_last_error().decode())
        self._h = C.c_void_p(handle)
        self._keep = keepalive
        self.desc = _ffi.host().hnmh_scene_desc(self._h)      # POINTER(SceneDesc)
        self.camera = _ffi.host().hnmh_scene_camera(self._h)  # POINTER(Camera)

    def __del__(self):
        try:
            _ffi.host().hnmh_scene_destroy(self._h)
        except Exception:
            pass

    def counts(self):
        d = self.desc.contents
        return {"elements": d.num_elements, "triangles": d.num_faces, "mesh_nodes": d.num_mesh_nodes,
                "top_nodes": d.num_top_nodes, "images": d.num_images, "emissions": d.num_emissions}


def build_scene(name, assets):
    """init_scene_* of src/main.rs by name: rtcamp6 (default, :1020-1153), rtcamp6_v4, simple,
    material_examples, bvh_heavy / diamond (BASELINE configs 3 / 4), *_pl variants."""
    return HostScene(_ffi.host().hnmh_scene_build(assets._h, name.encode()), keepalive=assets)


class SceneBuilder:
    """Ad-hoc scenes with the reference's vocabulary (Sphere, Cuboid, BvhMesh, Skybox, Camera::new)."""

    def __init__(self, assets):
        self.assets = assets
        self._h = C.c_void_p(_ffi.host().hnmh_builder_create())

    def __del__(self):
        try:
            _ffi.host().hnmh_builder_destroy(self._h)
        except Exception:
            pass

    @staticmethod
    def material(surface, param=0.0, albedo=(1, 1, 1), emission=(0, 0, 0), roughness=0.2, albedo_image=None,
                 emission_image=None, roughness_image=None):
        m = _ffi.HostMaterial()
        m.surface = surface
        m.param = param
        m.albedo[:] = albedo
        m.emission[:] = emission
        m.roughness[:] = (roughness,) * 3 if np.isscalar(roughness) else roughness
        m.albedo_image = albedo_image.encode() if albedo_image else None
        m.emission_image = emission_image.encode() if emission_image else None
        m.roughness_image = roughness_image.encode() if roughness_image else None
        return m

    def camera(self, eye, target, y_up=(0, 1, 0), v_fov=20.0, circle=True, aperture=0.0, focus_distance=5.0):
        a = [np.asarray(v, np.float64) for v in (eye, target, y_up)]
        _check_host(_ffi.host().hnmh_builder_camera(self._h, _vp(a[0]), _vp(a[1]), _vp(a[2]), v_fov, int(circle), aperture, focus_distance) == 0)

    def add_sphere(self, center, radius, material):
        c = np.asarray(center, np.float64)
        _check_host(_ffi.host().hnmh_builder_add_sphere(self._h, self.assets._h, _vp(c), radius, C.byref(material)) == 0)

    def add_cuboid(self, mn, mx, material):
        a, b = np.asarray(mn, np.float64), np.asarray(mx, np.float64)
        _check_host(_ffi.host().hnmh_builder_add_cuboid(self._h, self.assets._h, _vp(a), _vp(b), C.byref(material)) == 0)

    def add_mesh(self, vertices, faces, material):
        v = np.ascontiguousarray(vertices, np.float64)
        f = np.ascontiguousarray(faces, np.uint32)
        _check_host(_ffi.host().hnmh_builder_add_mesh(self._h, self.assets._h, _vp(v), len(v), _vp(f), len(f), C.byref(material)) == 0)

    def add_obj(self, path, matrix44, material):
        m = np.ascontiguousarray(matrix44, np.float64)
        _check_host(_ffi.host().hnmh_builder_add_obj(self._h, self.assets._h, path.encode(), _vp(m), C.byref(material)) == 0)

    def skybox(self, directory="textures/cube/Powerlines", intensity=(1, 1, 1), ext=".jpg"):
        names = [("%s/%s%s" % (directory, f, ext)).encode() for f in ("posx", "negx", "posy", "negy", "posz", "negz")]
        arr = (C.c_char_p * 6)(*names)
        i = np.asarray(intensity, np.float64)
        _check_host(_ffi.host().hnmh_builder_skybox(self._h, self.assets._h, arr, _vp(i)) == 0)

    def finish(self):
        return HostScene(_ffi.host().hnmh_builder_finish(self._h), keepalive=self.assets)


# --------------------------------------------------------------------------- device side (C ABI)
def device_count_or_zero():
    """device_count() that answers 0 instead of raising when there is no driver / device (tests)."""
    try:
        return device_count()
    except HanamaruError:
        return 0


def device_count():
    n = _ffi.core().hnm_device_count()
    if n < 0:  # a driver error is not "no device"
        raise HanamaruError("hanamaru_b200 error %d: %s" % (n, _ffi.core().hnm_last_error().decode()))
    return n


class DeviceScene:
    """hnm_scene: the scene deep-copied to one GPU."""

    def __init__(self, host_scene, device=0):
        self.host_scene = host_scene
        self.device = device
        h = C.c_void_p()
        _check(_ffi.core().hnm_scene_create(host_scene.desc, device, C.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            _ffi.core().hnm_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def intersect(self, origins, directions):
        """n closest-hit queries (BvhScene::intersect, src/scene.rs:385-401) -> numpy record array of hnm_hit."""
        o = np.ascontiguousarray(origins, np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float64).reshape(-1, 3)
        rays = np.ascontiguousarray(np.concatenate([o, d], axis=1))
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        _check(_ffi.core().hnm_intersect_batch(self._h, _vp(rays), len(rays), _vp(hits)))
        return hits


    def texture_sample(self, image, tint, uv):
        """`Texture::sample` (src/texture.rs:108-114) of scene image `image` (-1: constant) at n (u, v) pairs."""
        uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
        t = np.asarray(tint, np.float64)
        out = np.empty((len(uv), 3), np.float64)
        _check(_ffi.core().hnm_texture_sample_batch(self._h, int(image), _vp(t), _vp(uv), len(uv), _vp(out)))
        return out

    def skybox_sample(self, directions):
        """`Skybox::sample` (src/scene.rs:295-319) for n directions."""
        d = np.ascontiguousarray(directions, np.float64).reshape(-1, 3)
        out = np.empty((len(d), 3), np.float64)
        _check(_ffi.core().hnm_skybox_sample_batch(self._h, _vp(d), len(d), _vp(out)))
        return out


HIT_DTYPE = np.dtype([("position", np.float64, 3), ("normal", np.float64, 3), ("albedo", np.float64, 3), ("emission", np.float64, 3),
                      ("distance", np.float64), ("u", np.float64), ("v", np.float64), ("roughness", np.float64), ("param", np.float64),
                      ("hit", np.int32), ("element", np.int32), ("face", np.int32), ("surface", np.int32)])
assert HIT_DTYPE.itemsize == C.sizeof(_ffi.Hit)


class RenderContext:
    """hnm_renderer: wavefront state + the f64 accumulation buffer of one `Renderer::render` call."""

    def __init__(self, device_scene, camera, width, height, mode=MODE_PATHTRACING, shard=None, max_batch=0):
        self.scene = device_scene
        self.width, self.height, self.mode = width, height, mode
        sh = None
        if shard is not None:
            sh = _ffi.Shard(shard[0], shard[1], shard[2] if len(shard) > 2 else 8, 0)
        h = C.c_void_p()
        _check(_ffi.core().hnm_renderer_create(device_scene._h, camera, width, height, mode, C.byref(sh) if sh else None, max_batch, C.byref(h)))
        self._h = h
        self.owned_rows = _ffi.core().hnm_owned_rows(self._h)

    def close(self):
        if self._h:
            _ffi.core().hnm_renderer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render_passes(self, sampling_first, count):
        _check(_ffi.core().hnm_render_passes(self._h, sampling_first, count))

    def synchronize(self):
        _check(_ffi.core().hnm_synchronize(self._h))

    def clear(self):
        _check(_ffi.core().hnm_clear(self._h))

    def local_rows(self):
        return np.array([_ffi.core().hnm_local_row_to_global(self._h, i) for i in range(self.owned_rows)], dtype=np.int64)

    def read_accum(self):
        out = np.empty((self.owned_rows, self.width, 3), np.float64)
        _check(_ffi.core().hnm_read_accum(self._h, _vp(out)))
        return out

    def accum_device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        _check(_ffi.core().hnm_accum_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def resolve(self, sampling, accum_full_device=None, out=None):
        """update_imgbuf (src/renderer.rs:64-90) -> uint8 [H][W][3]."""
        if out is None:
            out = np.empty((self.height, self.width, 3), np.uint8)
        _check(_ffi.core().hnm_resolve(self._h, C.c_void_p(accum_full_device) if accum_full_device else None, sampling, _vp(out)))
        return out

    def resolve_begin(self, sampling, accum_full_device=None):
        """First half of resolve(): enqueued behind the passes submitted so far, returns at once."""
        _check(_ffi.core().hnm_resolve_begin(self._h, C.c_void_p(accum_full_device) if accum_full_device else None, sampling))

    def resolve_end(self, out=None):
        """Second half: waits for the image of the matching resolve_begin / dist_resolve_begin."""
        if out is None:
            out = np.empty((self.height, self.width, 3), np.uint8)
        _check(_ffi.core().hnm_resolve_end(self._h, _vp(out)))
        return out

    def deinterleave(self, gathered_device, full_device):
        _check(_ffi.core().hnm_deinterleave(self._h, C.c_void_p(gathered_device), C.c_void_p(full_device)))

    def counters(self):
        c = _ffi.Counters()
        _check(_ffi.core().hnm_get_counters(self._h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in c._fields_}

    def warp_slots(self):
        m = (C.c_uint64 * 4)()
        _check(_ffi.core().hnm_debug_warp_slots(self._h, m, 4))
        return [int(x) for x in m]

    def queue_counters(self, bounces=12):
        m = (C.c_uint32 * (16 * bounces))()
        _check(_ffi.core().hnm_debug_read_counters(self._h, m, 16 * bounces))
        return np.array(m, dtype=np.uint32).reshape(bounces, 16)

    def set_precision(self, precision):
        """_ffi.PRECISION_EXACT (default, bit parity) or _ffi.PRECISION_FAST_MATH (opt-in, statistical parity)."""
        _check(_ffi.core().hnm_set_precision(self._h, int(precision)))

    def set_profiling(self, on):
        _check(_ffi.core().hnm_set_profiling(self._h, int(on)))

    def mark(self, slot):
        _check(_ffi.core().hnm_mark(self._h, slot))

    def elapsed_ms(self, a, b):
        ms = C.c_float()
        _check(_ffi.core().hnm_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def kernel_times(self):
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        launches = (C.c_uint32 * 32)()
        n = C.c_uint32()
        _check(_ffi.core().hnm_get_kernel_times(self._h, 32, names, ms, launches, C.byref(n)))
        return {names[i].decode(): (ms[i], launches[i]) for i in range(n.value)}


    # ---- one process per device: the NCCL gather behind the C ABI (hnm_dist_*)
    def dist_init(self, unique_id, rank, num_ranks):
        buf = (C.c_uint8 * _ffi.HNM_DIST_ID_BYTES).from_buffer_copy(bytes(unique_id))
        _check(_ffi.core().hnm_dist_init(self._h, buf, rank, num_ranks))

    def dist_attach(self, comm):
        """Use a process-lifetime communicator (DistComm) instead of creating one for this renderer."""
        _check(_ffi.core().hnm_dist_attach(self._h, comm._h))
        self._comm = comm  # keep it alive

    def dist_resolve_begin(self, sampling, want_image=True):
        """Collective, asynchronous: gather (+ update_imgbuf on the ranks that want the image; collect it with resolve_end)."""
        _check(_ffi.core().hnm_dist_resolve_begin(self._h, sampling, 1 if want_image else 0))

    def dist_resolve(self, sampling, out=None, want_image=True):
        """Collective.  Ranks with want_image=False only take part in the gather (asynchronously) and return None."""
        if not want_image:
            _check(_ffi.core().hnm_dist_resolve(self._h, sampling, None))
            return None
        if out is None:
            out = np.empty((self.height, self.width, 3), np.uint8)
        _check(_ffi.core().hnm_dist_resolve(self._h, sampling, _vp(out)))
        return out

    def dist_read_accum(self, want=True):
        if not want:
            _check(_ffi.core().hnm_dist_read_accum(self._h, None))
            return None
        out = np.empty((self.height, self.width, 3), np.float64)
        _check(_ffi.core().hnm_dist_read_accum(self._h, _vp(out)))
        return out


def dist_unique_id():
    """Rank 0: the NCCL unique id the launcher distributes to every rank (hnm_dist_init)."""
    buf = (C.c_uint8 * _ffi.HNM_DIST_ID_BYTES)()
    _check(_ffi.core().hnm_dist_unique_id(buf))
    return bytes(buf)


class DistComm:
    """hnm_comm: one NCCL communicator per process, created after the launcher's rendezvous and attached to every
    renderer the process makes (ncclCommInitRank takes seconds at 8 ranks; a renderer lives for one `render` call)."""

    def __init__(self, device, unique_id, rank, num_ranks):
        buf = (C.c_uint8 * _ffi.HNM_DIST_ID_BYTES).from_buffer_copy(bytes(unique_id))
        h = C.c_void_p()
        _check(_ffi.core().hnm_comm_create(device, buf, rank, num_ranks, C.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            _ffi.core().hnm_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RenderGroup:
    """hnm_group: ONE process driving N devices -- scene uploaded to each, interleaved row tiles, peer-copy gather to
    device 0, update_imgbuf there.  What a single-process host (the Rust binary) calls instead of RenderContext."""

    def __init__(self, host_scene, camera, width, height, mode=MODE_PATHTRACING, devices=(0,), tile_rows=0, max_batch=0):
        self.width, self.height, self.mode = width, height, mode
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(_ffi.core().hnm_group_create(host_scene.desc, camera, width, height, mode, len(devices), devs, tile_rows, max_batch, C.byref(h)))
        self._h = h
        self._keep = host_scene

    def close(self):
        if self._h:
            _ffi.core().hnm_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render_passes(self, sampling_first, count):
        _check(_ffi.core().hnm_group_render_passes(self._h, sampling_first, count))

    def synchronize(self):
        _check(_ffi.core().hnm_group_synchronize(self._h))

    def clear(self):
        _check(_ffi.core().hnm_group_clear(self._h))

    def resolve(self, sampling, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 3), np.uint8)
        _check(_ffi.core().hnm_group_resolve(self._h, sampling, _vp(out)))
        return out

    def read_accum(self):
        out = np.empty((self.height, self.width, 3), np.float64)
        _check(_ffi.core().hnm_group_read_accum(self._h, _vp(out)))
        return out

    def counters(self):
        c = _ffi.Counters()
        _check(_ffi.core().hnm_group_get_counters(self._h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in c._fields_}


# --------------------------------------------------------------------------- the reference's Renderer surface
class Renderer:
    """`trait Renderer` (src/renderer.rs:20-99): render() owns the pass loop, report_progress() may stop it."""

    mode = MODE_PATHTRACING

    def max_sampling(self):
        raise NotImplementedError

    def report_progress(self, ctx, sampling, imgbuf):
        raise NotImplementedError

    def passes_per_call(self, sampling):
        return 1

    def render(self, device_scene, camera, imgbuf):
        """imgbuf: uint8 [H][W][3], written by update_imgbuf when report_progress decides to.  Returns passes done."""
        h, w, _ = imgbuf.shape
        ctx = RenderContext(device_scene, camera, w, h, self.mode)
        try:
            sampling = 0
            while sampling < self.max_sampling():
                n = max(1, min(self.passes_per_call(sampling), self.max_sampling() - sampling))
                ctx.render_passes(sampling + 1, n)  # NOTICE: sampling is 1 origin (src/renderer.rs:31)
                ctx.synchronize()
                sampling += n
                if self.report_progress(ctx, sampling, imgbuf):
                    return sampling
            return self.max_sampling()
        finally:
            self.last_counters = ctx.counters()
            ctx.close()

    @staticmethod
    def update_imgbuf(ctx, sampling, imgbuf):
        ctx.resolve(sampling, out=imgbuf)


class DebugRenderer(Renderer):
    """src/renderer.rs:109-146"""

    def __init__(self, mode=MODE_DEBUG_FOCALPLANE):
        self.mode = mode

    def max_sampling(self):
        return 1

    def report_progress(self, ctx, sampling, imgbuf):
        self.update_imgbuf(ctx, sampling, imgbuf)
        return True


class PathTracingRenderer(Renderer):
    """src/renderer.rs:148-267: time limit, report interval, progress images."""

    def __init__(self, sampling, time_limit_sec=123.0, report_interval_sec=15.0, batch=1, save_progress=None, verbose=False):
        self.sampling = sampling
        self.time_limit_sec = time_limit_sec
        self.report_interval_sec = report_interval_sec
        self.batch = batch
        self.save_progress = save_progress  # callable(path, imgbuf) or None
        self.verbose = verbose
        now = time.time()
        self.begin = self.last_report_progress = self.last_report_image = now
        self.report_image_counter = 0
        self._last_sampling = 0

    def max_sampling(self):
        return self.sampling

    def passes_per_call(self, sampling):
        return self.batch

    def _save(self, ctx, sampling, imgbuf):
        self.update_imgbuf(ctx, sampling, imgbuf)
        if self.save_progress:
            self.save_progress("%03d.png" % self.report_image_counter, imgbuf)

    def report_progress(self, ctx, sampling, imgbuf):
        now = time.time()
        used = now - self.begin
        from_last = now - self.last_report_progress
        done = sampling - self._last_sampling
        self._last_sampling = sampling
        if self.verbose:
            print("rendering: %dx4 sampled (last %.3f sec). total: %.3f sec (%.2f %%)." % (sampling, from_last, used, used / self.time_limit_sec * 100.0))
        # the reference predicts the next pass from the last one, x1.1 (src/renderer.rs:218);
        # with `done` passes per call the prediction is for the next call
        if used + from_last * 1.1 > self.time_limit_sec or sampling >= self.max_sampling():
            self._save(ctx, sampling, imgbuf)
            return True
        if now - self.last_report_image >= self.report_interval_sec:
            self._save(ctx, sampling, imgbuf)
            self.report_image_counter += 1
            self.last_report_image = now
        self.last_report_progress = now
        del done
        return False


def host_render(scene, mode, width, height, sampling=1, time_limit_sec=1e9, report_interval_sec=1e9, passes_per_call=0, device=0):
    """The reference-facing call through the C++ host mirror (csrc/host: `PathTracingRenderer` / `DebugRenderer` ::render
    over the C ABI), the way a compiled host would drive the core.  Returns (uint8 image [H][W][3], passes done)."""
    img = np.zeros((height, width, 3), np.uint8)
    done = C.c_uint32(0)
    rc = _ffi.host().hnmh_render(scene._h, int(mode), width, height, sampling, float(time_limit_sec), float(report_interval_sec),
                                 passes_per_call, device, _vp(img), C.byref(done))
    if rc != 0:
        raise HanamaruError(_ffi.host().hnmh_last_error().decode())
    return img, done.value


def host_render_to_files(scene, mode, width, height, out_dir, sampling=1, time_limit_sec=1e9, report_interval_sec=1e9, passes_per_call=0, device=0):
    """host_render + the reference's file outputs: NNN.png progress images and result.png, encoded by the C++ host."""
    img = np.zeros((height, width, 3), np.uint8)
    done = C.c_uint32(0)
    rc = _ffi.host().hnmh_render_to_files(scene._h, int(mode), width, height, sampling, float(time_limit_sec), float(report_interval_sec),
                                          passes_per_call, device, str(out_dir).encode(), _vp(img), C.byref(done))
    if rc != 0:
        raise HanamaruError(_ffi.host().hnmh_last_error().decode())
    return img, done.value


def image_decode(data):
    """The C++ host's `image::open` (PNG, baseline JPEG) on bytes -> uint8 [H][W][4]."""
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    w, h = C.c_uint32(), C.c_uint32()
    _check_host(_ffi.host().hnmh_image_decode(buf, len(data), C.byref(w), C.byref(h), None) == 0)
    out = np.empty((h.value, w.value, 4), np.uint8)
    _check_host(_ffi.host().hnmh_image_decode(buf, len(data), C.byref(w), C.byref(h), _vp(out)) == 0)
    return out


def save_png(path, rgb):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    _check_host(_ffi.host().hnmh_save_png(str(path).encode(), _vp(rgb), rgb.shape[1], rgb.shape[0]) == 0)


# --------------------------------------------------------------------------- batch entry points
def isaac64_batch(seeds, count, device=0):
    s = np.ascontiguousarray(seeds, np.uint64).reshape(-1, 4)
    out = np.empty((len(s), count), np.uint64)
    _check(_ffi.core().hnm_isaac64_batch(device, _vp(s), len(s), count, _vp(out)))
    return out


def material_sample_batch(inputs, device=0):
    a = np.ascontiguousarray(inputs, np.float64).reshape(-1, 14)
    out = np.empty((len(a), 8), np.float64)
    _check(_ffi.core().hnm_material_sample_batch(device, _vp(a), len(a), _vp(out)))
    return out


def material_bsdf_batch(inputs, device=0):
    a = np.ascontiguousarray(inputs, np.float64).reshape(-1, 12)
    out = np.empty(len(a), np.float64)
    _check(_ffi.core().hnm_material_bsdf_batch(device, _vp(a), len(a), _vp(out)))
    return out


def math_batch(fn, x, y=None, device=0):
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros_like(x) if y is None else np.ascontiguousarray(y, np.float64)
    out = np.empty_like(x)
    _check(_ffi.core().hnm_math_batch(device, fn, _vp(x), _vp(y), x.size, _vp(out)))
    return out
