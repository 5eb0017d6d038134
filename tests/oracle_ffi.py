"""ctypes binding of oracle/liboracle{,_det}.so.  TEST INFRASTRUCTURE: importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

from hanamaru_renderer_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class OracleCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("paths", "segments", "shadow_rays", "node_visits", "tri_tests", "elem_tests", "lens_iters")]

    def dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


class Oracle:
    def __init__(self, flavor="det"):
        name = "liboracle_det.so" if flavor == "det" else "liboracle.so"
        path = os.path.join(ORACLE_DIR, name)
        if not os.path.exists(path):
            build_oracle()
        self.lib = C.CDLL(path)
        self.lib.oracle_flavor.restype = C.c_char_p
        self.flavor = self.lib.oracle_flavor().decode()

    def isaac64(self, seed, count, skip=0):
        s = np.asarray(seed, np.uint64)
        out = np.empty(count, np.uint64)
        self.lib.oracle_isaac64_skip(_vp(s), C.c_uint32(len(s)), C.c_uint32(skip), C.c_uint32(count), _vp(out))
        return out

    def isaac64_f64(self, seed, count):
        s = np.asarray(seed, np.uint64)
        out = np.empty(count, np.float64)
        self.lib.oracle_isaac64_f64(_vp(s), C.c_uint32(len(s)), C.c_uint32(count), _vp(out))
        return out

    def math(self, fn, x, y=None):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x) if y is None else np.ascontiguousarray(y, np.float64)
        out = np.empty_like(x)
        self.lib.oracle_math(C.c_int(fn), _vp(x), _vp(y), C.c_uint32(x.size), _vp(out))
        return out

    def intersect(self, host_scene, origins, directions):
        from hanamaru_renderer_b200 import HIT_DTYPE
        o = np.ascontiguousarray(origins, np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float64).reshape(-1, 3)
        rays = np.ascontiguousarray(np.concatenate([o, d], axis=1))
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        self.lib.oracle_intersect_batch(host_scene.desc, _vp(rays), C.c_uint32(len(rays)), _vp(hits))
        return hits

    def material_sample(self, host_scene_or_cfg, inputs):
        a = np.ascontiguousarray(inputs, np.float64).reshape(-1, 14)
        out = np.empty((len(a), 8), np.float64)
        cfg = host_scene_or_cfg if isinstance(host_scene_or_cfg, _ffi.Config) else host_scene_or_cfg.desc.contents.config
        self.lib.oracle_material_sample_batch(C.byref(cfg), _vp(a), C.c_uint32(len(a)), _vp(out))
        return out

    def material_bsdf(self, inputs):
        a = np.ascontiguousarray(inputs, np.float64).reshape(-1, 12)
        out = np.empty(len(a), np.float64)
        self.lib.oracle_material_bsdf_batch(_vp(a), C.c_uint32(len(a)), _vp(out))
        return out

    def render_paths(self, host_scene, w, h, mode, sampling, camera=None):
        ss = host_scene.desc.contents.config.supersampling
        out = np.empty((h, w, ss * ss, 3), np.float64)
        cnt = OracleCounters()
        self.lib.oracle_render_paths(host_scene.desc, camera or host_scene.camera, C.c_uint32(w), C.c_uint32(h), C.c_int(mode),
                                     C.c_uint32(sampling), _vp(out), C.byref(cnt))
        return out, cnt.dict()

    def render(self, host_scene, w, h, mode, sampling_first, count, accum=None, rows=None, camera=None, counters=True, cols=None):
        if accum is None:
            accum = np.zeros((h, w, 3), np.float64)
        r0, r1 = rows if rows else (0, h)
        c0, c1 = cols if cols else (0, w)
        cnt = OracleCounters()
        self.lib.oracle_render_rect(host_scene.desc, camera or host_scene.camera, C.c_uint32(w), C.c_uint32(h), C.c_int(mode),
                                    C.c_uint32(sampling_first), C.c_uint32(count), C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(r0),
                                    C.c_uint32(r1), _vp(accum), C.byref(cnt) if counters else None)
        return accum, cnt.dict()

    def resolve(self, cfg, accum, sampling):
        accum = np.ascontiguousarray(accum, np.float64)
        h, w, _ = accum.shape
        out = np.empty((h, w, 3), np.uint8)
        self.lib.oracle_resolve(C.byref(cfg), _vp(accum), C.c_uint32(w), C.c_uint32(h), C.c_uint32(sampling), _vp(out))
        return out

    def texture_sample(self, host_scene, image, tint, uv):
        uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
        out = np.empty((len(uv), 3), np.float64)
        self.lib.oracle_texture_sample(host_scene.desc, C.c_int32(image), C.c_double(tint[0]), C.c_double(tint[1]), C.c_double(tint[2]),
                                       _vp(uv), C.c_uint32(len(uv)), _vp(out))
        return out

    def skybox_sample(self, host_scene, dirs):
        d = np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        out = np.empty((len(d), 3), np.float64)
        self.lib.oracle_skybox_sample(host_scene.desc, _vp(d), C.c_uint32(len(d)), _vp(out))
        return out

    def camera_ray(self, host_scene, w, h, x, y, sx, sy, sampling, dof=True):
        out = np.empty(7, np.float64)
        self.lib.oracle_camera_ray(host_scene.desc, host_scene.camera, C.c_uint32(w), C.c_uint32(h), C.c_uint32(x), C.c_uint32(y),
                                   C.c_uint32(sx), C.c_uint32(sy), C.c_uint32(sampling), C.c_int(int(dof)), _vp(out))
        return out
