"""CPU tests: the host-side mirror (OBJ loader, BVH builder, flattening), the asset pack, and the C-ABI
library (loads, exports every declared symbol, struct layouts agree) -- no compute calls."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


def test_obj_parser_follows_loader_rs(hr):
    a = hr.AssetStore()
    # split on single spaces, only v / f, quads -> (a,b,c),(a,c,d), 1-based, "i/j/k" takes i, CRLF stripped
    a.put_obj_text("t.obj", "# c\r\nv 0 0 0\r\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1/1/1 2/2/2 3/3/3 4/4/4\nf 1 2 3\ng x\n")
    v, f = a.obj_geometry("t.obj")
    assert v.tolist() == [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]]
    assert f.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]
    # double space after `v` makes the reference panic on parse::<f64>("") -- reported, not papered over
    with pytest.raises(hr.HanamaruError):
        a.put_obj_text("bad.obj", "v  0 0 0\n")


def test_bvh_builder_properties(hr, get_scene):
    """src/bvh.rs:107-201: leaf iff len/2 <= 2 (<= 5 faces), first child = first half, DFS pre-order, exact boxes."""
    d = get_scene("rtcamp6").desc.contents
    for mi in range(d.num_meshes):
        m = d.meshes[mi]
        seen = np.zeros(m.face_count, bool)
        verts = np.ctypeslib.as_array(d.vertices, shape=(d.num_vertices, 3))[m.vertex_offset:m.vertex_offset + m.vertex_count]
        faces = np.ctypeslib.as_array(d.faces, shape=(d.num_faces, 3))[m.face_offset:m.face_offset + m.face_count]

        def walk(rel):
            n = d.mesh_nodes[m.node_offset + rel]
            if n.child0 < 0:
                assert 1 <= n.count <= 5
                idx = [d.mesh_indices[m.index_offset + n.first + k] for k in range(n.count)]
                for i in idx:
                    assert not seen[i]
                    seen[i] = True
                tri = verts[faces[idx].reshape(-1)]
                assert np.array_equal(tri.min(0), np.array(n.aabb_min[:])) and np.array_equal(tri.max(0), np.array(n.aabb_max[:]))
                return len(idx), tri.min(0), tri.max(0)
            assert n.child0 == rel + 1 and n.child1 > n.child0
            c0, lo0, hi0 = walk(n.child0)
            c1, lo1, hi1 = walk(n.child1)
            assert c0 == (c0 + c1) // 2 and c0 + c1 > 5          # split_off(len / 2): first child gets the first half
            lo, hi = np.minimum(lo0, lo1), np.maximum(hi0, hi1)
            assert np.array_equal(lo, np.array(n.aabb_min[:])) and np.array_equal(hi, np.array(n.aabb_max[:]))
            return c0 + c1, lo, hi

        total, _, _ = walk(0)
        assert total == m.face_count and seen.all()
    # top level: 11 elements -> root, leaf(5), inner, leaf(3), leaf(3)
    kinds = [(n.child0 < 0, n.count) for n in (d.top_nodes[i] for i in range(d.num_top_nodes))]
    assert kinds == [(False, 0), (True, 5), (False, 0), (True, 3), (True, 3)]
    assert sorted(d.top_indices[i] for i in range(d.num_top_indices)) == list(range(11))


def test_default_scene_authoring(hr, get_scene):
    """src/main.rs:1020-1153 spot checks."""
    s = get_scene("rtcamp6")
    d, cam = s.desc.contents, s.camera.contents
    theta = 2 * np.pi * 0.03
    assert np.allclose(cam.eye.tuple(), (6.5 * np.sin(theta), 2.0, 6.5 * np.cos(theta)), rtol=0, atol=1e-15)
    assert cam.lens_radius == 0.015 and cam.focus_distance == 5.0 and cam.lens_shape == 1
    # Camera::new uses tan(v_fov) with the FULL fov as the half angle (src/camera.rs:48)
    assert abs(np.linalg.norm(cam.plane_half_up.tuple()) - np.tan(np.radians(20.0)) * 5.0) < 1e-14
    e0 = d.elements[0]
    assert e0.kind == 0 and e0.radius == 0.2 and e0.a.tuple() == (-0.3, 0.7, 0.0)
    m0 = d.materials[e0.material]
    assert m0.emission.color.tuple() == (30.0, 20.0, 4.0) and m0.albedo.color.tuple() == (0.0, 0.0, 0.0)
    assert [d.elements[i].kind for i in range(11)] == [0, 2, 2, 2, 1, 2, 2, 2, 2, 2, 2]
    surf = [d.materials[d.elements[i].material].surface for i in range(5, 11)]
    assert surf == [hr.SURFACE_REFRACTION, hr.SURFACE_GGX] * 3
    rough = [d.materials[d.elements[i].material].roughness.color.x for i in range(5, 11)]
    assert np.allclose(rough, [0.1, 0.05, 0.1, 0.15, 0.1, 0.25])
    assert d.emissions[0] == 0 and d.num_emissions == 1
    c = d.config
    assert (c.eps, c.offset, c.inf, c.gamma_factor, c.supersampling, c.bounce_limit) == (1e-4, 1e-4, 1e100, 2.2, 2, 10)


def test_other_scenes_build(hr, get_scene):
    assert get_scene("bvh_heavy").counts()["triangles"] == 12294 + 55888 + 7200
    assert get_scene("diamond").counts()["triangles"] == 7 * 94
    assert get_scene("material_examples_pl").counts()["elements"] == 7
    s = get_scene("simple_pl")
    assert s.counts()["emissions"] == 2 and s.desc.contents.skybox_intensity.tuple() == (0.0, 0.0, 0.0)


def _stdrng(hr, seed, skip, count, kind=0, low=0.0, high=0.0):
    from hanamaru_renderer_b200 import _ffi
    seed = np.asarray(seed, np.uint64)
    out = np.zeros(count, np.uint64 if kind == 0 else np.float64)
    rc = _ffi.host().hnmh_stdrng(seed.ctypes.data_as(C.c_void_p), len(seed), skip, count, kind, low, high, out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def test_host_stdrng_known_answers(hr):
    """The host's StdRng (scene authoring, src/main.rs:253-254) against rand 0.4.3's own test_rng_64_true_values."""
    a = _stdrng(hr, [1, 23, 456, 7890, 12345], 0, 10)
    assert a.tolist() == [547121783600835980, 14377643087320773276, 17351601304698403469, 1238879483818134882, 11952566807690396487,
                          13970131091560099343, 4469761996653280935, 15552757044682284409, 6860251611068737823, 13722198873481261842]
    b = _stdrng(hr, [12345, 67890, 54321, 9876], 10000, 10)
    assert b.tolist() == [18143823860592706164, 8491801882678285927, 2699425367717515619, 17196852593171130876, 2606123525235546165,
                          15790932315217671084, 596345674630742204, 9947027391921273664, 11788097613744130851, 10391409374914919106]
    # gen_range(low, high) = low + (high - low) * next_f64(), next_f64 = (0x3FF0.. | u64 & (2^52-1)) - 1.0
    u = _stdrng(hr, [870, 2000, 304, 2], 0, 6)
    f = ((u & np.uint64(0xFFFFFFFFFFFFF)) | np.uint64(0x3FF0000000000000)).view(np.float64) - 1.0
    g = _stdrng(hr, [870, 2000, 304, 2], 0, 6, kind=1, low=-4.5, high=4.5)
    assert np.array_equal(g, -4.5 + 9.0 * f) and ((g >= -4.5) & (g < 4.5)).all()


def test_rng_authored_scenes(hr, get_scene):
    """init_scene_rtcamp5 / init_scene_tbf3 (src/main.rs:252-722): fixed elements + StdRng-placed ones that passed
    add_with_check_collisions (src/scene.rs:366-376)."""
    s = get_scene("rtcamp5_pl")
    d = s.desc.contents
    assert s.counts()["elements"] == 11 + 12 + 30 and s.counts()["emissions"] == 1
    e = d.elements[d.emissions[0]]
    assert e.kind == 0 and e.a.tuple() == (0.0, 0.5, -0.5) and d.materials[e.material].emission.image >= 0
    assert d.materials[e.material].emission.color.tuple() == (5.0, 5.0, 2.0)
    t = get_scene("tbf3_pl")
    d = t.desc.contents
    assert t.counts()["elements"] == 8 + 8 + 20 and t.counts()["emissions"] == 4
    assert d.skybox_intensity.tuple() == (2.0, 2.0, 3.0)
    # no two element boxes overlap among the RNG-placed elements and everything placed before them
    def box(e):
        if e.kind == 0:
            c, r = np.array(e.a.tuple()), e.radius
            return c - r, c + r
        return np.array(e.a.tuple()), np.array(e.b.tuple())
    for scene, first_random in ((s, 11), (t, 8)):
        d = scene.desc.contents
        boxes = [box(d.elements[i]) for i in range(d.num_elements)]
        for i in range(first_random, d.num_elements):
            for j in range(i):
                lo_i, hi_i = boxes[i]
                lo_j, hi_j = boxes[j]
                assert not ((lo_i < hi_j).all() and (hi_i > lo_j).all()), (i, j)
    v3 = get_scene("rtcamp6_v3")
    d3 = v3.desc.contents
    cam = v3.camera.contents
    assert v3.counts()["emissions"] == 2 and d3.elements[1].radius == 0.001
    # the camera light sits at eye - forward (src/main.rs:957)
    assert np.allclose(d3.elements[1].a.tuple(), np.array(cam.eye.tuple()) - np.array(cam.forward.tuple()), rtol=0, atol=1e-15)
    assert get_scene("rtcamp6_v1_pl").counts() == {"elements": 3, "triangles": 2520, "mesh_nodes": 1023, "top_nodes": 1, "images": 8, "emissions": 1}
    v2 = get_scene("rtcamp6_v2_pl")
    assert v2.counts()["elements"] == 100 + 5 + 1 and v2.counts()["emissions"] == 5
    assert v2.desc.contents.skybox_intensity.tuple() == (0.5, 0.5, 0.5)
    # the metal spheres of tbf3: radius in [0.2, 0.4), resting on the floor, hue 0.2 + 0.1 * k
    d = t.desc.contents
    for k in range(8):
        e = d.elements[8 + k]
        assert e.kind == 0 and 0.2 <= e.radius < 0.4 and e.a.y == e.radius + 0.0
        assert 0.0 <= d.materials[e.material].roughness.color.x < 0.2


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_pack_matches_reference_assets(hr):
    """The committed asset pack is exactly what the host builds from the reference checkout."""
    pack = hr.AssetStore.from_pack()
    ref = hr.AssetStore.from_reference(REFERENCE, ("rtcamp6",))
    for p in hr.scene_asset_paths("bvh_heavy", images=False) + hr.scene_asset_paths("diamond", images=False):
        v0, f0 = pack.obj_geometry(p)
        v1, f1 = ref.obj_geometry(p)
        assert np.array_equal(v0, v1) and np.array_equal(f0, f1), p
    a, b = hr.build_scene("rtcamp6", pack), hr.build_scene("rtcamp6", ref)
    da, db = a.desc.contents, b.desc.contents
    assert a.counts() == b.counts()
    va = np.ctypeslib.as_array(da.vertices, shape=(da.num_vertices * 3,))
    vb = np.ctypeslib.as_array(db.vertices, shape=(db.num_vertices * 3,))
    assert np.array_equal(va, vb)
    for i in range(da.num_images):
        ia, ib = da.images[i], db.images[i]
        assert (ia.width, ia.height) == (ib.width, ib.height)
        pa = np.ctypeslib.as_array(C.cast(ia.rgba, C.POINTER(C.c_uint8)), shape=(ia.height * ia.width * 4,))
        pb = np.ctypeslib.as_array(C.cast(ib.rgba, C.POINTER(C.c_uint8)), shape=(ib.height * ib.width * 4,))
        assert np.array_equal(pa, pb)


# ---- the C ABI ---------------------------------------------------------------------------------
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hanamaru_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hnm_[a-z0-9_]+)\s*\(", text)))


def test_abi_library_exports_every_declared_symbol(hr):
    import __graft_entry__ as g
    path = g.build_core()          # nvcc cross-compiles sm_100a without a GPU
    lib = C.CDLL(path)             # loads without a GPU (no compute call is made)
    declared = _declared_symbols()
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name
    from hanamaru_renderer_b200 import _ffi
    assert sorted(_ffi.CORE_SYMBOLS) == declared, "ctypes table and header disagree"
    lib.hnm_abi_version.restype = C.c_uint32
    assert lib.hnm_abi_version() == _ffi.HNM_ABI_VERSION
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(r"\bT %s\b" % name, out), name


def test_abi_struct_layouts(hr):
    """ctypes mirrors == the C compiler's layout of include/hanamaru_b200.h."""
    from hanamaru_renderer_b200 import _ffi
    names = {"hnm_vec3": _ffi.Vec3, "hnm_camera": _ffi.Camera, "hnm_texture": _ffi.Texture, "hnm_material": _ffi.Material,
             "hnm_image": _ffi.Image, "hnm_element": _ffi.Element, "hnm_bvh_node": _ffi.BvhNode, "hnm_mesh": _ffi.Mesh,
             "hnm_config": _ffi.Config, "hnm_scene_desc": _ffi.SceneDesc, "hnm_shard": _ffi.Shard, "hnm_counters": _ffi.Counters,
             "hnm_ray": _ffi.Ray, "hnm_hit": _ffi.Hit}
    src = '#include <stdio.h>\n#include "hanamaru_b200.h"\nint main(){\n' + "".join(
        'printf("%s %%zu\\n", sizeof(%s));\n' % (n, n) for n in names) + (
        'printf("off_config %zu\\n", offsetof(hnm_scene_desc, config));\n'
        'printf("off_skybox %zu\\n", offsetof(hnm_scene_desc, skybox_images));\n return 0;}\n')
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(td, "t.c"), "-o", os.path.join(td, "t")])
        out = dict(l.split() for l in subprocess.run([os.path.join(td, "t")], capture_output=True, text=True).stdout.splitlines())
    for n, t in names.items():
        assert int(out[n]) == C.sizeof(t), n
    assert int(out["off_config"]) == _ffi.SceneDesc.config.offset
    assert int(out["off_skybox"]) == _ffi.SceneDesc.skybox_images.offset


def test_no_cpu_fallback_without_device(hr):
    """Without a GPU every compute entry point fails loudly; nothing routes to the oracle."""
    import __graft_entry__ as g
    g.build_core()
    if hr.device_count_or_zero() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(hr.HanamaruError):
        hr.device_count()          # a driver error is reported, not read as "0 devices"
    assets = hr.AssetStore.from_pack()
    scene = hr.build_scene("diamond", assets)
    with pytest.raises(hr.HanamaruError):
        hr.DeviceScene(scene, 0)
    with pytest.raises(hr.HanamaruError):
        hr.isaac64_batch([[1, 2, 3, 4]], 8)
    with pytest.raises(hr.HanamaruError):
        hr.DistComm(0, bytes(128), 0, 2)   # no device (or no NCCL): an error, not a silent single-rank mode
    with pytest.raises(hr.HanamaruError):
        hr.DistComm(0, bytes(128), 3, 2)   # rank >= num_ranks
    # and the product package never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "hanamaru_renderer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle_ffi" not in text, f
                assert not re.search(r'#include\s+"[^"]*oracle', text), f
                assert not re.search(r"^\s*(import|from)\s+oracle", text, flags=re.M), f


def test_cpp_host_renderer_needs_the_device(hr, get_scene):
    """The C++ host mirror (csrc/host Renderer::render over the C ABI) has no CPU fallback either."""
    if hr.device_count_or_zero() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(hr.HanamaruError, match="CUDA"):
        hr.host_render(get_scene("rtcamp6"), hr.MODE_PATHTRACING, 32, 18, 1)


def test_scene_validation_rejects_malformed(hr, get_scene):
    """hnm_scene_create validates before touching the device: malformed descriptions -> HNM_ERR_INVALID."""
    import copy
    from hanamaru_renderer_b200 import _ffi
    import __graft_entry__ as g
    g.build_core()
    core = _ffi.core()
    src = get_scene("diamond").desc.contents
    bad = _ffi.SceneDesc()
    C.memmove(C.byref(bad), C.byref(src), C.sizeof(bad))
    bad.abi_version = 99
    h = C.c_void_p()
    assert core.hnm_scene_create(C.byref(bad), 0, C.byref(h)) == -1
    assert b"abi_version" in core.hnm_last_error()
    C.memmove(C.byref(bad), C.byref(src), C.sizeof(bad))
    bad.skybox_images[2] = 1000
    assert core.hnm_scene_create(C.byref(bad), 0, C.byref(h)) == -1
    C.memmove(C.byref(bad), C.byref(src), C.sizeof(bad))
    bad.num_elements = 0
    assert core.hnm_scene_create(C.byref(bad), 0, C.byref(h)) == -1
    assert core.hnm_scene_create(None, 0, C.byref(h)) == -1
    # bounce_limit: 2 * (bounce_limit - 1) words of the per-path random stream must fit the stored tail (HNM_RNG_TAIL = 32)
    for bl, ok in ((1, False), (18, False), (64, False)):
        C.memmove(C.byref(bad), C.byref(src), C.sizeof(bad))
        bad.config.bounce_limit = bl
        assert core.hnm_scene_create(C.byref(bad), 0, C.byref(h)) == -1, bl
        assert b"bounce_limit" in core.hnm_last_error()
    # the group entry point validates the same way, before touching any device
    devs = (C.c_int * 1)(0)
    C.memmove(C.byref(bad), C.byref(src), C.sizeof(bad))
    bad.config.bounce_limit = 18
    assert core.hnm_group_create(C.byref(bad), get_scene("diamond").camera, 16, 16, 0, 1, devs, 0, 0, C.byref(h)) == -1
    del copy
