"""GPU parity tests proper: the CUDA path through the C ABI against the oracle (deterministic-libm flavour,
same arithmetic, so the bar is BIT-EXACT f64 -- not a tolerance), plus size-independent properties at
BASELINE.json's full sizes.  Everything here calls libhanamaru_b200.so; the oracle is only the checker."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def field_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.tobytes() == b.tobytes()


def render_gpu(hr, dev, scene, w, h, mode, first, count, shard=None, max_batch=0):
    ctx = hr.RenderContext(dev, scene.camera, w, h, mode, shard=shard, max_batch=max_batch)
    ctx.render_passes(first, count)
    ctx.synchronize()
    return ctx


# ---------------------------------------------------------------------------------------- unit parity
def test_native_library_is_what_runs(hr, core):
    from hanamaru_renderer_b200 import _ffi
    assert os.path.basename(_ffi.CORE_LIB) == "libhanamaru_b200.so" and os.path.exists(_ffi.CORE_LIB)
    maps = open("/proc/self/maps").read()
    assert "libhanamaru_b200.so" in maps


def test_isaac64_known_answers_on_device(hr, core, oracle):
    # rand's KATs need 5-word seeds / 10k skips: the device seeds 4 words (src/renderer.rs:167), so pin it
    # against the KAT-pinned oracle on random 4-word seeds, both the shared-memory kernel and the exact slow path
    rng = np.random.default_rng(11)
    seeds = rng.integers(0, 2 ** 63, size=(3000, 4), dtype=np.uint64)
    seeds[0] = [8700304, 1, 403480, 401216]
    seeds[1] = [0, 0, 0, 0]
    seeds[2] = [2 ** 64 - 1] * 4
    got = hr.isaac64_batch(seeds, 32)
    want = np.stack([oracle.isaac64(s, 32) for s in seeds])
    assert np.array_equal(got, want)
    # group sizes around the 28-path granule of the TMEM pipeline, partial tails, fewer groups than warp pairs, and a
    # count that covers several groups per pair (20000 > 148 SMs x 4 pairs x 28)
    for n in (1, 27, 28, 29, 113):
        assert np.array_equal(hr.isaac64_batch(seeds[:n], 32), want[:n]), n
    big = rng.integers(0, 2 ** 63, size=(20000, 4), dtype=np.uint64)
    gb = hr.isaac64_batch(big, 5)
    for k in (0, 1, 27, 28, 16575, 16576, 19999):
        assert np.array_equal(gb[k], oracle.isaac64(big[k], 5)), k
    got = hr.isaac64_batch(seeds[:200], 700)  # > 256 outputs: refill
    want = np.stack([oracle.isaac64(s, 700) for s in seeds[:200]])
    assert np.array_equal(got, want)
    v = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    assert np.array_equal(hr.isaac64_batch([[8700304, 1, 223146, 300912]], 32)[0], v["isaac_first_path"])


def test_device_libm_is_bit_identical_to_oracle(hr, core, oracle):
    rng = np.random.default_rng(2)
    n = 1 << 20
    x = rng.random(n)
    assert np.array_equal(bits(hr.math_batch(0, x * 2 * np.pi)), bits(oracle.math(0, x * 2 * np.pi)))
    assert np.array_equal(bits(hr.math_batch(1, x * 2 * np.pi)), bits(oracle.math(1, x * 2 * np.pi)))
    assert np.array_equal(bits(hr.math_batch(2, -x * 745)), bits(oracle.math(2, -x * 745)))
    assert np.array_equal(bits(hr.math_batch(2, (x - 0.5) * 1400)), bits(oracle.math(2, (x - 0.5) * 1400)))
    for y in (2.2, 1 / 2.2):
        yy = np.full(n, y)
        assert np.array_equal(bits(hr.math_batch(3, x, yy)), bits(oracle.math(3, x, yy)))
    t = rng.integers(0, 256, n) / 255.0
    assert np.array_equal(bits(hr.math_batch(3, t, np.full(n, 2.2))), bits(oracle.math(3, t, np.full(n, 2.2))))
    assert np.array_equal(bits(hr.math_batch(4, x * 2 - 1)), bits(oracle.math(4, x * 2 - 1)))
    edge = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 1e-300, np.nan])
    assert np.array_equal(bits(hr.math_batch(4, edge)), bits(oracle.math(4, edge)))


def _material_inputs(rng, n):
    surf = rng.integers(0, 5, n).astype(np.float64)
    param = np.where(surf == 3, rng.random(n), 1.0 + rng.random(n) * 1.5)
    rough = rng.random(n)
    r = rng.random((n, 2))
    pos = rng.normal(size=(n, 3))
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    view = rng.normal(size=(n, 3))
    view /= np.linalg.norm(view, axis=1, keepdims=True)
    return np.concatenate([surf[:, None], param[:, None], rough[:, None], r, pos, view, nrm], axis=1)


def test_material_sample_and_bsdf_batches(hr, core, oracle, get_scene):
    rng = np.random.default_rng(4)
    a = _material_inputs(rng, 1 << 18)
    # axis-aligned normals exercise the |normal.x| > EPS tangent-frame switch (src/material.rs:203)
    a[:6, 11:14] = [[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]
    a[6, 11:14] = [5e-5, 1, 0]
    got = hr.material_sample_batch(a)
    want = oracle.material_sample(get_scene("rtcamp6"), a)
    assert np.array_equal(bits(got), bits(want))
    assert 0.01 < (got[:, 0] == 0).mean() < 0.5  # the GGX `None` branch is exercised
    b = np.concatenate([a[:, 0:3], a[:, 8:11], a[:, 11:14], rng.normal(size=(len(a), 3))], axis=1)
    b[:, 9:12] /= np.linalg.norm(b[:, 9:12], axis=1, keepdims=True)
    b = b[(b[:, 0] == 0) | (b[:, 0] == 3)]  # bsdf() is only implemented (and only reachable) for Diffuse / GGX
    assert np.array_equal(bits(hr.material_bsdf_batch(b)), bits(oracle.material_bsdf(b)))


def _random_rays(rng, scene, n):
    cam = np.array(scene.camera.contents.eye.tuple())
    o = np.concatenate([np.tile(cam, (n // 2, 1)), rng.uniform(-4, 4, size=(n - n // 2, 3)) * [1, 0.5, 1] + [0, 1.0, 0]])
    d = rng.normal(size=(n, 3))
    d[: n // 2] = -cam / np.linalg.norm(cam) + rng.normal(size=(n // 2, 3)) * 0.25
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d


@pytest.mark.parametrize("name", ["rtcamp6", "bvh_heavy", "diamond", "material_examples_pl", "simple_pl", "tbf3_pl", "rtcamp5_pl", "rtcamp6_v2_pl", "rtcamp6_v3"])
def test_intersect_batch_matches_oracle(hr, core, oracle, get_scene, get_device_scene, name):
    """BvhScene::intersect incl. material resolve (textures, skybox) on random rays: every field bit-identical."""
    scene, dev = get_scene(name), get_device_scene(name)
    rng = np.random.default_rng(9)
    o, d = _random_rays(rng, scene, 200000)
    # degenerate directions: axis-parallel (1/0 = inf in the slab tests), and rays starting ON the floor plane
    d[:6] = [[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]
    o[6:12, 1] = 0.0
    got = dev.intersect(o, d)
    want = oracle.intersect(scene, o, d)
    for f in got.dtype.names:
        g, w = got[f], want[f]
        same = np.array_equal(bits(g), bits(w)) if g.dtype == np.float64 else np.array_equal(g, w)
        assert same, (name, f, int((g != w).sum()))
    assert 0.05 < got["hit"].mean() < 1.0


def test_intersect_far_origin_and_tiny_scene(hr, core, oracle, assets):
    """Origins far outside the scene box take the advance-to-box path of the f32 traversal; results unchanged."""
    b = hr.SceneBuilder(assets)
    b.camera((0, 0, 5000), (0, 0, 0))
    mat = hr.SceneBuilder.material(hr.SURFACE_GGX, param=0.8, roughness=0.3, albedo_image="textures/2d/checkered_diagonal_10_0.5_1.0_512.png")
    b.add_sphere((0, 0, 0), 1.0, mat)     # textured sphere: uv through acos (src/scene.rs:69-73)
    b.skybox()
    scene = b.finish()
    dev = hr.DeviceScene(scene, 0)
    rng = np.random.default_rng(1)
    n = 50000
    o = rng.normal(size=(n, 3))
    o = o / np.linalg.norm(o, axis=1, keepdims=True) * rng.uniform(2, 1e6, size=(n, 1))
    tgt = rng.uniform(-1.2, 1.2, size=(n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    got, want = dev.intersect(o, d), oracle.intersect(scene, o, d)
    for f in got.dtype.names:
        assert field_equal(got[f], want[f]), f
    assert got["hit"].mean() > 0.2


def test_exact_ties_follow_reference_order(hr, core, oracle, assets):
    """Duplicated and coincident triangles: the reference keeps the LAST hit in its DFS order among exact
    distance ties (`t > distance` rejects, src/bvh.rs:283).  The GPU traversal visits in another order."""
    rng = np.random.default_rng(6)
    v = rng.uniform(-1, 1, size=(60, 3))
    f = rng.integers(0, 60, size=(150, 3))
    f = f[(f[:, 0] != f[:, 1]) & (f[:, 1] != f[:, 2]) & (f[:, 0] != f[:, 2])]
    f = np.concatenate([f, f[::3], f[::5][:, [1, 2, 0]]])  # exact duplicates and rotated duplicates (same plane, same t? no: same triangle)
    b = hr.SceneBuilder(assets)
    b.camera((0, 0, 5), (0, 0, 0))
    m1 = hr.SceneBuilder.material(hr.SURFACE_DIFFUSE, albedo=(0.2, 0.3, 0.4))
    m2 = hr.SceneBuilder.material(hr.SURFACE_SPECULAR, albedo=(0.9, 0.8, 0.7))
    b.add_mesh(v, f, m1)
    b.add_mesh(v, f[::-1], m2)            # a second element with the same geometry: ties across elements
    b.add_sphere((0, 0, 0), 0.5, m1)
    b.add_sphere((0, 0, 0), 0.5, m2)      # coincident spheres: strict `<` keeps the FIRST
    b.add_cuboid((-0.45, -0.45, -0.3), (0.45, 0.45, 0.3), m1)
    b.add_cuboid((-0.45, -0.45, -0.3), (0.45, 0.45, 0.3), m2)
    b.skybox()
    scene = b.finish()
    dev = hr.DeviceScene(scene, 0)
    n = 100000
    o = rng.normal(size=(n, 3))
    o = o / np.linalg.norm(o, axis=1, keepdims=True) * 4
    d = rng.uniform(-0.9, 0.9, size=(n, 3)) - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    got, want = dev.intersect(o, d), oracle.intersect(scene, o, d)
    assert np.array_equal(got["element"], want["element"])
    assert np.array_equal(got["face"], want["face"])
    assert np.array_equal(bits(got["distance"]), bits(want["distance"]))
    assert np.array_equal(bits(got["normal"]), bits(want["normal"]))
    # mesh ties go to the LATER mesh (element 1), sphere / cuboid ties to the EARLIER one (2 and 4)
    assert set(np.unique(got["element"]).tolist()) >= {-1, 1, 2, 4}


def test_candidate_list_overflow_falls_back_to_exact_traversal(hr, core, oracle, assets):
    """k_trace keeps at most 32 candidates per ray; a ray with more (here: 40 coincident copies of every triangle, all
    tied at the same distance) is re-traced by k_confirm with the exact traversal (trace_pretested).  Same bits either way, for
    single rays and for whole paths (NEE shadow rays through the stack included)."""
    rng = np.random.default_rng(11)
    v = np.array([[-2, 0.5, -2], [2, 0.5, -2], [2, 0.5, 2], [-2, 0.5, 2], [0, 1.5, 0]], np.float64)
    f = np.array([[0, 1, 2], [0, 2, 3], [0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]], np.uint32)
    f = np.concatenate([f] * 40)
    b = hr.SceneBuilder(assets)
    b.camera((0, 3, 6), (0, 0.5, 0), aperture=0.05, focus_distance=6.0)
    b.add_mesh(v, f, hr.SceneBuilder.material(hr.SURFACE_GGX, param=0.8, roughness=0.3, albedo=(0.8, 0.6, 0.4)))
    b.add_sphere((0.5, 3.0, 0.5), 0.3, hr.SceneBuilder.material(hr.SURFACE_DIFFUSE, albedo=(0, 0, 0), emission=(20, 20, 20)))
    b.add_cuboid((-5, -1, -5), (5, 0, 5), hr.SceneBuilder.material(hr.SURFACE_DIFFUSE, albedo=(0.7, 0.7, 0.7)))
    b.skybox()
    scene = b.finish()
    dev = hr.DeviceScene(scene, 0)
    n = 50000
    o = rng.normal(size=(n, 3)) * [1, 0.3, 1] + [0, 4, 0]
    d = rng.uniform(-1.5, 1.5, size=(n, 3)) * [1, 0, 1] + [0, 0.8, 0] - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    got, want = dev.intersect(o, d), oracle.intersect(scene, o, d)
    for fld in got.dtype.names:
        assert field_equal(got[fld], want[fld]), fld
    assert (got["element"] == 0).mean() > 0.3          # many rays do hit the 40-fold mesh
    ctx = render_gpu(hr, dev, scene, 96, 54, hr.MODE_PATHTRACING, 1, 2)
    acc, cnt = oracle.render(scene, 96, 54, hr.MODE_PATHTRACING, 1, 2)
    assert np.array_equal(bits(ctx.read_accum()), bits(acc))
    c = ctx.counters()
    assert (c["segments"], c["shadow_rays"]) == (cnt["segments"], cnt["shadow_rays"])
    ctx.close()


# ---------------------------------------------------------------------------------------- debug passes
@pytest.mark.parametrize("w,h", [(480, 270), (1920, 1080)])
def test_debug_passes_bit_exact(hr, core, oracle, get_scene, get_device_scene, w, h):
    """DebugRenderer Normal / Depth / FocalPlane / Shading: f64 buffer AND the u8 image identical (BASELINE north_star)."""
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    cfg = scene.desc.contents.config
    modes = [hr.MODE_DEBUG_NORMAL, hr.MODE_DEBUG_DEPTH, hr.MODE_DEBUG_FOCALPLANE, hr.MODE_DEBUG_SHADING]
    if w > 1000:
        modes = [hr.MODE_DEBUG_NORMAL, hr.MODE_DEBUG_FOCALPLANE]  # keep the oracle's share of the suite short
    for mode in modes:
        ctx = render_gpu(hr, dev, scene, w, h, mode, 1, 1)
        got = ctx.read_accum()
        img = ctx.resolve(1)
        ctx.close()
        want, _ = oracle.render(scene, w, h, mode, 1, 1, counters=False)
        assert np.array_equal(bits(got), bits(want)), mode
        assert np.array_equal(img, oracle.resolve(cfg, want, 1)), mode


def test_debug_golden_fixture(hr, core, get_scene, get_device_scene):
    v = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    for name, mode in (("normal", hr.MODE_DEBUG_NORMAL), ("depth", hr.MODE_DEBUG_DEPTH), ("focal", hr.MODE_DEBUG_FOCALPLANE),
                       ("shading", hr.MODE_DEBUG_SHADING)):
        ctx = render_gpu(hr, dev, scene, 96, 54, mode, 1, 1)
        assert np.array_equal(ctx.resolve(1), v["debug_%s_96x54_rgb8" % name]), name
        ctx.close()


# ---------------------------------------------------------------------------------------- path tracing
def test_pathtracing_golden_fixture(hr, core, get_scene, get_device_scene):
    v = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    ctx = render_gpu(hr, dev, scene, 64, 36, hr.MODE_PATHTRACING, 1, 2)
    assert np.array_equal(bits(ctx.read_accum()), bits(v["pt_64x36_s2_accum"]))
    assert np.array_equal(ctx.resolve(2), v["pt_64x36_s2_rgb8"])
    c = ctx.counters()
    assert [c["paths"], c["segments"], c["shadow_rays"]] == v["pt_64x36_s2_counters"][:3].tolist()
    ctx.close()


def test_pathtracing_config1_bit_exact(hr, core, oracle, get_scene, get_device_scene):
    """BASELINE config 1: default scene 480x270, `-s 1`: every pixel of the f64 accumulation buffer identical,
    same segment / shadow-ray counts, identical resolved image."""
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    ctx = render_gpu(hr, dev, scene, 480, 270, hr.MODE_PATHTRACING, 1, 1)
    got = ctx.read_accum()
    want, cnt = oracle.render(scene, 480, 270, hr.MODE_PATHTRACING, 1, 1)
    assert np.array_equal(bits(got), bits(want)), int((got != want).any(axis=2).sum())
    c = ctx.counters()
    assert (c["paths"], c["segments"], c["shadow_rays"]) == (cnt["paths"], cnt["segments"], cnt["shadow_rays"])
    assert np.array_equal(ctx.resolve(1), oracle.resolve(scene.desc.contents.config, want, 1))
    ctx.close()


@pytest.mark.parametrize("name,w,h,first,count", [
    ("rtcamp6", 160, 90, 1, 6),                 # several passes accumulate in pass order
    ("rtcamp6", 160, 90, 1000, 2),              # late `sampling` values (part of the seed)
    ("material_examples_pl", 200, 120, 1, 3),   # all five BSDFs on spheres, textured roughness, DoF aperture 0.2
    ("simple_pl", 200, 120, 1, 3),              # two lights (two shadow rays per NEE), GGX floor, black sky
    ("bvh_heavy", 160, 90, 1, 2),               # BASELINE config 3 scene
    ("diamond", 160, 90, 1, 3),                 # BASELINE config 4 scene: GGXRefraction, strong DoF
    ("tbf3_pl", 160, 90, 1, 3),                 # four emitters with a TEXTURED emission (4 shadow rays per NEE), StdRng layout
    ("rtcamp5_pl", 160, 90, 1, 2),              # 45 diamonds (Refraction 2.42), textured emitter, roughness map, marble floor
    ("rtcamp6_v2_pl", 160, 90, 1, 2),           # 105 StdRng-placed spheres, FIVE emitters (5 shadow rays per NEE), no floor
    ("rtcamp6_v3", 160, 90, 1, 2),              # second emitter: a 1 mm sphere behind the camera, smaller than the NEE window
    ("rtcamp6_v1_pl", 160, 90, 1, 2),           # refractive mesh in front of the light, textured albedo AND roughness floor
    ("rtcamp6_v4", 160, 90, 1, 2),              # Ryfjallet cube map (the reference's JPEG files, decoded by the C++ host)
    ("rtcamp5", 160, 90, 1, 2),                 # the same builders under their OWN cube maps (LancellottiChapel), not the
    ("tbf3", 160, 90, 1, 2),                    # Powerlines stand-in of the `_pl` variants
    ("rtcamp6_v2", 160, 90, 1, 2),
    ("rtcamp6_v1", 160, 90, 1, 2),
    ("rtcamp6", 97, 61, 3, 2),                  # odd sizes
    ("rtcamp6", 1, 1, 1, 1),
    ("rtcamp6", 3, 200, 1, 1),                  # W < H: min(res) picks the width
])
def test_pathtracing_scenes_bit_exact(hr, core, oracle, get_scene, get_device_scene, name, w, h, first, count):
    scene, dev = get_scene(name), get_device_scene(name)
    ctx = render_gpu(hr, dev, scene, w, h, hr.MODE_PATHTRACING, first, count)
    got = ctx.read_accum()
    want, cnt = oracle.render(scene, w, h, hr.MODE_PATHTRACING, first, count)
    assert np.array_equal(bits(got), bits(want)), int((got != want).any(axis=2).sum())
    c = ctx.counters()
    assert (c["paths"], c["segments"], c["shadow_rays"]) == (cnt["paths"], cnt["segments"], cnt["shadow_rays"])
    assert np.array_equal(ctx.resolve(first + count - 1), oracle.resolve(scene.desc.contents.config, want, first + count - 1))
    ctx.close()


def test_rng_tail_overflow_path(hr, core, oracle, get_scene, get_device_scene, monkeypatch):
    """Paths whose lens rejection loop outruns the stored ISAAC tail take the exact slow path (full generator in
    local memory).  With the tail shortened to 20 words every path with more than one lens iteration takes it."""
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    monkeypatch.setenv("HNM_RNG_TAIL_K", "20")
    ctx = render_gpu(hr, dev, scene, 120, 68, hr.MODE_PATHTRACING, 1, 2)
    monkeypatch.delenv("HNM_RNG_TAIL_K")
    got = ctx.read_accum()
    c = ctx.counters()
    ctx.close()
    want, cnt = oracle.render(scene, 120, 68, hr.MODE_PATHTRACING, 1, 2)
    assert np.array_equal(bits(got), bits(want))
    assert 0.15 * c["paths"] < c["rng_fallbacks"] < 0.3 * c["paths"]   # P(reject) = 1 - pi/4 = 0.215


def test_batching_and_sharding_do_not_change_bits(hr, core, get_scene, get_device_scene):
    """Property at any size: passes-in-flight and row-tile sharding never change a bit (each pixel has one owner,
    per-pixel accumulation order is the pass order)."""
    from hanamaru_renderer_b200 import dist as hd
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    w, h, passes = 320, 180, 5
    ref = render_gpu(hr, dev, scene, w, h, hr.MODE_PATHTRACING, 1, passes, max_batch=1)
    want = ref.read_accum()
    ref.close()
    for mb in (2, 5, 0):
        ctx = render_gpu(hr, dev, scene, w, h, hr.MODE_PATHTRACING, 1, passes, max_batch=mb)
        assert np.array_equal(bits(ctx.read_accum()), bits(want)), mb
        ctx.close()
    # passes split across calls
    ctx = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING)
    ctx.render_passes(1, 2)
    ctx.render_passes(3, 3)
    ctx.synchronize()
    assert np.array_equal(bits(ctx.read_accum()), bits(want))
    ctx.close()
    for nranks, tile in ((2, 8), (4, 4), (8, 4), (8, 2), (3, 16)):
        shards = []
        for rank in range(nranks):
            ctx = render_gpu(hr, dev, scene, w, h, hr.MODE_PATHTRACING, 1, passes, shard=(rank, nranks, tile))
            assert ctx.owned_rows == hd.padded_rows(h, nranks, tile)
            assert ctx.local_rows().tolist() == [hd.local_row_to_global(lr, rank, nranks, tile) for lr in range(ctx.owned_rows)]
            shards.append(ctx.read_accum())
            ctx.close()
        full = hd.deinterleave_numpy(np.stack(shards), h, nranks, tile)
        assert np.array_equal(bits(full), bits(want)), (nranks, tile)


def test_deinterleave_and_resolve_on_device(hr, core, oracle, get_scene, get_device_scene):
    """The multi-GPU resolve path on one device: shards -> (gathered buffer) -> k_deinterleave -> resolve == 1-GPU image."""
    import torch
    from hanamaru_renderer_b200 import dist as hd
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    w, h, passes, nranks, tile = 200, 117, 2, 4, 8
    one = render_gpu(hr, dev, scene, w, h, hr.MODE_PATHTRACING, 1, passes)
    want_img = one.resolve(passes)
    want = one.read_accum()
    one.close()
    ctxs = [render_gpu(hr, dev, scene, w, h, hr.MODE_PATHTRACING, 1, passes, shard=(r, nranks, tile)) for r in range(nranks)]
    gathered = torch.stack([hd.accum_as_tensor(c).clone() for c in ctxs])
    full = torch.zeros((h, w, 3), dtype=torch.float64, device="cuda:0")
    torch.cuda.synchronize()
    ctxs[0].deinterleave(gathered.data_ptr(), full.data_ptr())
    assert np.array_equal(bits(full.cpu().numpy()), bits(want))
    img = ctxs[0].resolve(passes, accum_full_device=full.data_ptr())
    assert np.array_equal(img, want_img)
    assert np.array_equal(img, oracle.resolve(scene.desc.contents.config, want, passes))
    for c in ctxs:
        c.close()


def test_resolve_in_two_halves(hr, core, get_scene, get_device_scene):
    """hnm_resolve_begin / hnm_resolve_end: the image enqueued behind step i is the image resolve() returns for step i, also when the
    passes of step i + 1 are enqueued before it is collected; misuse is an error."""
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    w, h = 160, 90
    ctx = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING)
    ref = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING)
    imgs = []
    for i in range(3):
        ctx.render_passes(1 + 2 * i, 2)
        if i > 0:
            imgs.append(ctx.resolve_end())
        ctx.resolve_begin(2 * (i + 1))
        with pytest.raises(hr.HanamaruError):
            ctx.resolve_begin(2 * (i + 1))      # one image may be pending
    imgs.append(ctx.resolve_end())
    with pytest.raises(hr.HanamaruError):
        ctx.resolve_end()                        # nothing pending
    for i in range(3):
        ref.render_passes(1 + 2 * i, 2)
        assert np.array_equal(ref.resolve(2 * (i + 1)), imgs[i]), i
    ctx.close(); ref.close()


def test_resolve_edge_cases(hr, core, oracle, get_scene, get_device_scene):
    """update_imgbuf on adversarial buffers: NaN, negatives, huge values, tiny images (u32 wrap of the bilateral taps)."""
    import torch
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    cfg = scene.desc.contents.config
    rng = np.random.default_rng(8)
    for (w, h) in ((3, 3), (1, 1), (17, 5), (150, 90)):
        acc = rng.gamma(0.5, 2.0, size=(h, w, 3))
        acc.flat[:: 7] = 0.0
        if w * h > 4:
            acc[0, 0] = np.nan
            acc[h - 1, w - 1] = [-3.0, 1e30, 1e-310]
        ctx = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING)
        t = torch.from_numpy(acc).cuda()
        for sampling in (1, 7):
            img = ctx.resolve(sampling, accum_full_device=t.data_ptr())
            assert np.array_equal(img, oracle.resolve(cfg, acc, sampling)), (w, h, sampling)
        ctx.close()


def test_renderer_surface(hr, core, oracle, get_scene, get_device_scene):
    """The reference-facing API: PathTracingRenderer / DebugRenderer .render(scene, camera, imgbuf) -> passes done."""
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    w, h = 128, 72
    img = np.zeros((h, w, 3), np.uint8)
    r = hr.PathTracingRenderer(3, time_limit_sec=1e9, report_interval_sec=1e9, batch=2)
    assert r.render(dev, scene.camera, img) == 3
    want, _ = oracle.render(scene, w, h, hr.MODE_PATHTRACING, 1, 3, counters=False)
    assert np.array_equal(img, oracle.resolve(scene.desc.contents.config, want, 3))
    # a zero time limit stops after the first call, like the reference's `used + offset > time_limit`
    r = hr.PathTracingRenderer(50, time_limit_sec=0.0, batch=1)
    assert r.render(dev, scene.camera, img) == 1
    d = hr.DebugRenderer(hr.MODE_DEBUG_FOCALPLANE)
    assert d.render(dev, scene.camera, img) == 1
    want, _ = oracle.render(scene, w, h, hr.MODE_DEBUG_FOCALPLANE, 1, 1, counters=False)
    assert np.array_equal(img, oracle.resolve(scene.desc.contents.config, want, 1))


def test_cpp_host_renderer_matches_oracle(hr, core, oracle, get_scene):
    """`renderer.render(&scene, &camera, &mut imgbuf)` through the C++ host mirror (csrc/host: PathTracingRenderer and
    DebugRenderer over the C ABI, dlopen of the core): same u8 image as the oracle, same pass count as the reference loop."""
    scene = get_scene("rtcamp6")
    w, h = 160, 90
    img, done = hr.host_render(scene, hr.MODE_PATHTRACING, w, h, sampling=3)
    want, _ = oracle.render(scene, w, h, hr.MODE_PATHTRACING, 1, 3, counters=False)
    assert done == 3 and np.array_equal(img, oracle.resolve(scene.desc.contents.config, want, 3))
    img, done = hr.host_render(scene, hr.MODE_PATHTRACING, w, h, sampling=5, passes_per_call=2)
    want, _ = oracle.render(scene, w, h, hr.MODE_PATHTRACING, 1, 5, counters=False)
    assert done == 5 and np.array_equal(img, oracle.resolve(scene.desc.contents.config, want, 5))
    img, done = hr.host_render(scene, hr.MODE_DEBUG_NORMAL, w, h)
    want, _ = oracle.render(scene, w, h, hr.MODE_DEBUG_NORMAL, 1, 1, counters=False)
    assert done == 1 and np.array_equal(img, oracle.resolve(scene.desc.contents.config, want, 1))
    # a time limit that is already over stops after the first call (src/renderer.rs:216-226)
    img, done = hr.host_render(scene, hr.MODE_PATHTRACING, w, h, sampling=50, time_limit_sec=0.0)
    assert done == 1


def test_full_size_properties(hr, core, oracle, get_scene, get_device_scene):
    """BASELINE config 2 size (1920x1080): one full pass is bit-identical to the oracle (8.3 M paths), and
    the multi-pass run keeps the size-independent invariants: determinism, additivity over passes, counters."""
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    w, h = 1920, 1080
    ctx = render_gpu(hr, dev, scene, w, h, hr.MODE_PATHTRACING, 1, 1)
    a1 = ctx.read_accum()
    want, cnt = oracle.render(scene, w, h, hr.MODE_PATHTRACING, 1, 1)
    assert np.array_equal(bits(a1), bits(want))
    c = ctx.counters()
    assert (c["paths"], c["segments"], c["shadow_rays"]) == (cnt["paths"], cnt["segments"], cnt["shadow_rays"])
    ctx.render_passes(2, 3)
    ctx.synchronize()
    a4 = ctx.read_accum()
    ctx.close()
    ctx = render_gpu(hr, dev, scene, w, h, hr.MODE_PATHTRACING, 2, 3)
    a234 = ctx.read_accum()
    ctx.close()
    # passes 2..4 alone differ from 1..4 exactly by pass 1 up to the order of f64 additions
    assert np.allclose(a4, a1 + a234, rtol=1e-12, atol=0)
    ctx = render_gpu(hr, dev, scene, w, h, hr.MODE_PATHTRACING, 1, 4, max_batch=2)
    assert np.array_equal(bits(ctx.read_accum()), bits(a4))   # deterministic, batch-independent
    assert ctx.counters()["paths"] == 4 * w * h * 4
    ctx.close()
    assert np.isfinite(a4).all() and (a4 >= 0).all()


def test_4k_bands_bit_exact(hr, core, oracle, get_scene, get_device_scene):
    """BASELINE config 5 resolution (3840x2160, 33.2 M paths per pass) on one GPU: bands of rows against the oracle
    (top / sky, the armadillo ring, floor), the whole-image path count, and the resolved image of those bands' interior."""
    scene, dev = get_scene("rtcamp6"), get_device_scene("rtcamp6")
    w, h = 3840, 2160
    ctx = render_gpu(hr, dev, scene, w, h, hr.MODE_PATHTRACING, 1, 1)
    got = ctx.read_accum()
    c = ctx.counters()
    assert c["paths"] == w * h * 4
    for r0 in (0, 1000, 1400, 2152):
        want, _ = oracle.render(scene, w, h, hr.MODE_PATHTRACING, 1, 1, rows=(r0, r0 + 8), counters=False)
        assert np.array_equal(bits(got[r0:r0 + 8]), bits(want[r0:r0 + 8])), r0
    assert np.isfinite(got).all()
    ctx.close()
