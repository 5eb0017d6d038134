"""Pins the oracle (CPU tests, no GPU): rand's ISAAC-64 known-answer vectors, hand-computed values for
the pure functions, the reference's golden image (statistically), committed regression vectors, and
the glibc-vs-deterministic-libm flavours against each other."""
import os

import numpy as np
import pytest
from PIL import Image

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# rand 0.4 src/prng/isaac64.rs, test_rng_64_true_values (quoted in SURVEY 8c)
KAT1_SEED = [1, 23, 456, 7890, 12345]
KAT1 = [547121783600835980, 14377643087320773276, 17351601304698403469, 1238879483818134882, 11952566807690396487,
        13970131091560099343, 4469761996653280935, 15552757044682284409, 6860251611068737823, 13722198873481261842]
KAT2_SEED = [12345, 67890, 54321, 9876]
KAT2 = [18143823860592706164, 8491801882678285927, 2699425367717515619, 17196852593171130876, 2606123525235546165,
        15790932315217671084, 596345674630742204, 9947027391921273664, 11788097613744130851, 10391409374914919106]


def test_isaac64_known_answers(oracle):
    assert oracle.isaac64(KAT1_SEED, 10).tolist() == KAT1
    assert oracle.isaac64(KAT2_SEED, 10, skip=10000).tolist() == KAT2


def test_isaac64_f64_mapping(oracle):
    # Rng::next_f64 of rand 0.4: 0x3FF0... | (u64 & (2^52-1)) reinterpreted, minus 1.0  -> [0, 1)
    u = oracle.isaac64(KAT1_SEED, 64)
    f = oracle.isaac64_f64(KAT1_SEED, 64)
    want = ((u & np.uint64((1 << 52) - 1)) | np.uint64(0x3FF0000000000000)).view(np.float64) - 1.0
    assert np.array_equal(f, want)
    assert (f >= 0).all() and (f < 1).all()


def test_isaac64_regression_vector(oracle):
    v = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    assert np.array_equal(oracle.isaac64([8700304, 1, 223146, 300912], 32), v["isaac_first_path"])


def test_scene_counts_match_survey(get_scene):
    # default scene = 11 elements, 12,294 triangles, 7,231 mesh nodes + 5 top nodes, 7 images, 1 light
    c = get_scene("rtcamp6").counts()
    assert c == {"elements": 11, "triangles": 12294, "mesh_nodes": 7231, "top_nodes": 5, "images": 7, "emissions": 1}


def test_workload_shape_matches_survey(oracle, get_scene, hr):
    # SURVEY section 6 measured 2.11 segments, 0.94 shadow rays, 1.27 lens iterations per path,
    # 106.8 node visits, 48.7 triangle tests per ray with an independent restatement
    _, c = oracle.render(get_scene("rtcamp6"), 240, 135, hr.MODE_PATHTRACING, 1, 1)
    paths = c["paths"]
    rays = c["segments"] + c["shadow_rays"]
    assert paths == 240 * 135 * 4
    assert abs(c["segments"] / paths - 2.11) < 0.03
    assert abs(c["shadow_rays"] / paths - 0.94) < 0.03
    assert abs(c["lens_iters"] / paths - 1.27) < 0.02
    assert abs(c["node_visits"] / rays - 106.8) < 3.0
    assert abs(c["tri_tests"] / rays - 48.7) < 2.0


def test_committed_vectors(oracle, oracle_glibc, get_scene, hr):
    v = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    scene = get_scene("rtcamp6")
    cfg = scene.desc.contents.config
    acc, cnt = oracle.render(scene, 64, 36, hr.MODE_PATHTRACING, 1, 2)
    assert np.array_equal(acc.view(np.uint64), v["pt_64x36_s2_accum"].view(np.uint64))
    assert [cnt[k] for k in ("paths", "segments", "shadow_rays", "lens_iters")] == v["pt_64x36_s2_counters"].tolist()
    assert np.array_equal(oracle.resolve(cfg, acc, 2), v["pt_64x36_s2_rgb8"])
    # the glibc flavour (what Rust's std calls) agrees to rounding noise: same discrete decisions, ulp-level radiance
    acc_g, cnt_g = oracle_glibc.render(scene, 64, 36, hr.MODE_PATHTRACING, 1, 2)
    assert cnt_g == cnt
    rel = np.linalg.norm(acc_g - acc, axis=2) / np.maximum(np.linalg.norm(acc, axis=2), 1e-300)
    assert rel.max() < 1e-11
    img_g = oracle_glibc.resolve(cfg, acc_g, 2)
    assert np.abs(img_g.astype(int) - v["pt_64x36_s2_rgb8"].astype(int)).max() <= 1
    for name, mode in (("normal", hr.MODE_DEBUG_NORMAL), ("depth", hr.MODE_DEBUG_DEPTH), ("focal", hr.MODE_DEBUG_FOCALPLANE),
                       ("shading", hr.MODE_DEBUG_SHADING)):
        for orc in (oracle, oracle_glibc):
            a, _ = orc.render(scene, 96, 54, mode, 1, 1)
            assert np.array_equal(orc.resolve(cfg, a, 1), v["debug_%s_96x54_rgb8" % name]), (name, orc.flavor)


@pytest.mark.parametrize("name,lights", [("tbf3_pl", 4), ("rtcamp6_v2_pl", 5), ("rtcamp6_v3", 2), ("rtcamp5_pl", 1)])
def test_multi_light_scenes_both_flavours(oracle, oracle_glibc, get_scene, hr, name, lights):
    """NEE loops over EVERY emissive sphere (src/renderer.rs:275): one shadow ray per light and NEE-able hit.  The glibc
    and the deterministic-libm flavours take the same discrete decisions on these scenes too."""
    scene = get_scene(name)
    assert scene.counts()["emissions"] == lights
    acc, cnt = oracle.render(scene, 48, 27, hr.MODE_PATHTRACING, 1, 1)
    acc_g, cnt_g = oracle_glibc.render(scene, 48, 27, hr.MODE_PATHTRACING, 1, 1)
    assert cnt == cnt_g and cnt["shadow_rays"] % lights == 0 and cnt["shadow_rays"] > 0
    rel = np.linalg.norm(acc_g - acc, axis=2) / np.maximum(np.linalg.norm(acc, axis=2), 1e-300)
    assert rel.max() < 1e-10
    assert np.isfinite(acc).all() and (acc >= 0).all() and acc.max() > 0


def test_golden_image_statistical(oracle, get_scene, hr):
    """The reference's only golden artefact: rtcamp6_1000x4spp.png.  16 passes at 480x270 against the
    4x4 box-downsampled golden (tools/make_golden.py).  Different sample count and a different JPEG decoder, so
    this is statistical: PSNR after a further 4x4 box filter, with a vertical flip as the control."""
    scene = get_scene("rtcamp6")
    acc, _ = oracle.render(scene, 480, 270, hr.MODE_PATHTRACING, 1, 8, counters=False)
    img = oracle.resolve(scene.desc.contents.config, acc, 8).astype(np.float64)
    gold = np.asarray(Image.open(os.path.join(GOLDEN, "rtcamp6_golden_480x270.png")).convert("RGB"), dtype=np.float64)

    def box(x, k=4):
        h, w, _ = x.shape
        return x[:h // k * k, :w // k * k].reshape(h // k, k, w // k, k, 3).mean(axis=(1, 3))

    def psnr(a, b):
        return 10 * np.log10(255.0 ** 2 / np.mean((a - b) ** 2))

    good = psnr(box(img), box(gold))
    control = psnr(box(img[::-1]), box(gold))
    assert good > 27.0, good
    assert control < 20.0, control
    assert abs(img.mean() - gold.mean()) < 6.0


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the LancellottiChapel cubemap of the reference checkout")
def test_rtcamp5_layout_matches_published_image(oracle, hr):
    """rtcamp5.png, the reference's other published render: 42 diamonds whose position, scale and two rotation angles
    come from StdRng::gen_range (src/main.rs:451-496) filtered by add_with_check_collisions.  The image pins the
    host's restatement of rand 0.4's u64 -> f64 mapping and draw order: any other mapping scatters the diamonds
    elsewhere.  (The published floor texture differs from the one the current source names, hence the modest PSNR.)"""
    scene = hr.build_scene("rtcamp5", hr.AssetStore.from_reference("/root/reference", ("rtcamp5",)))
    acc, _ = oracle.render(scene, 480, 270, hr.MODE_PATHTRACING, 1, 4, counters=False)
    img = oracle.resolve(scene.desc.contents.config, acc, 4).astype(np.float64)
    gold = np.asarray(Image.open(os.path.join(GOLDEN, "rtcamp5_golden_480x270.png")).convert("RGB"), dtype=np.float64)

    def box(x, k=4):
        h, w, _ = x.shape
        return x[:h // k * k, :w // k * k].reshape(h // k, k, w // k, k, 3).mean(axis=(1, 3))

    def psnr(a, b):
        return 10 * np.log10(255.0 ** 2 / np.mean((a - b) ** 2))

    good = psnr(box(img), box(gold))
    assert good > 21.0, good
    assert psnr(box(img[::-1]), box(gold)) < 15.0 and psnr(box(img[:, ::-1]), box(gold)) < 18.0
    # the upper half of the image is sky + floating diamonds only: their silhouettes line up
    top = slice(0, 100)
    assert psnr(box(img[top]), box(gold[top])) > psnr(box(img[top][:, ::-1]), box(gold[top])) + 4.0


# ---- pure functions, hand-computed --------------------------------------------------------------
def _ms(oracle, hr, surface, param, rough, r0, r1, pos, view, normal):
    from hanamaru_renderer_b200 import _ffi
    cfg = _ffi.Config(eps=1e-4, offset=1e-4, inf=1e100, gamma_factor=2.2, supersampling=2, bounce_limit=10)
    return oracle.material_sample(cfg, [[surface, param, rough, r0, r1] + list(pos) + list(view) + list(normal)])[0]


def test_material_specular_is_mirror(oracle, hr):
    v = np.array([1.0, 1.0, 0.0]) / np.sqrt(2.0)
    out = _ms(oracle, hr, hr.SURFACE_SPECULAR, 0, 0, 0.3, 0.7, (0, 0, 0), v, (0, 1, 0))
    assert out[0] == 1.0 and out[7] == 1.0
    assert np.allclose(out[1:4], [0, 1e-4, 0])
    assert np.allclose(out[4:7], [-v[0], v[1], 0], atol=1e-15)


def test_material_diffuse_hemisphere(oracle, hr):
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        out = _ms(oracle, hr, hr.SURFACE_DIFFUSE, 0, 0, rng.random(), rng.random(), (1, 2, 3), n, n)
        d = out[4:7]
        assert abs(np.linalg.norm(d) - 1) < 1e-12 and d @ n >= 0 and out[7] == 1.0
        assert np.allclose(out[1:4], np.array([1, 2, 3]) + 1e-4 * n)


def test_material_refraction_total_internal_reflection(oracle, hr):
    # leaving glass (view . normal < 0 means the RAY enters against the normal: is_incoming uses the ray = -view)
    # ray travelling +y inside a medium of index 1.5 hitting the surface at a grazing angle -> TIR -> mirror, reflectance 1
    ray = np.array([0.9, np.sqrt(1 - 0.81), 0.0])
    out = _ms(oracle, hr, hr.SURFACE_REFRACTION, 1.5, 0, 0.99, 0.5, (0, 0, 0), -ray, (0, 1, 0))
    assert out[0] == 1.0 and out[7] == 1.0
    assert np.allclose(out[4:7], [0.9, -np.sqrt(1 - 0.81), 0.0], atol=1e-15)
    assert np.allclose(out[1:4], [0, -1e-4, 0])   # oriented normal is -n: the origin stays inside


def test_material_refraction_fresnel_branches(oracle, hr):
    ray = np.array([0.0, -1.0, 0.0])  # straight down onto glass: fr = ((1.5-1)/(1.5+1))^2 = 0.04
    refl = _ms(oracle, hr, hr.SURFACE_REFRACTION, 1.5, 0, 0.039, 0.5, (0, 0, 0), -ray, (0, 1, 0))
    refr = _ms(oracle, hr, hr.SURFACE_REFRACTION, 1.5, 0, 0.041, 0.5, (0, 0, 0), -ray, (0, 1, 0))
    assert np.allclose(refl[4:7], [0, 1, 0]) and refl[7] == 1.0
    assert np.allclose(refr[4:7], [0, -1, 0]) and abs(refr[7] - (1 / 1.5) ** 2) < 1e-15
    assert np.allclose(refr[1:4], [0, -1e-4, 0])


def test_material_ggx_below_hemisphere_is_none(oracle, hr):
    # grazing view + rough surface: some half vectors reflect the ray below the surface -> None (src/material.rs:125-127)
    rng = np.random.default_rng(5)
    v = np.array([0.999, 0.0447101778, 0.0])
    v /= np.linalg.norm(v)
    outs = np.array([_ms(oracle, hr, hr.SURFACE_GGX, 0.8, 0.9, rng.random(), rng.random(), (0, 0, 0), v, (0, 1, 0)) for _ in range(200)])
    assert (outs[:, 0] == 0).any() and (outs[:, 0] == 1).any()
    some = outs[outs[:, 0] == 1]
    assert (some[:, 5] >= 0).all() and (some[:, 7] >= 0).all() and (some[:, 7] <= 1).all()


def test_bsdf_values(oracle, hr):
    n = [0, 1, 0]
    v = [0, 1, 0]
    assert oracle.material_bsdf([[hr.SURFACE_DIFFUSE, 0, 0.5] + v + n + [0, 1, 0]])[0] == 1.0 / np.pi
    # GGX at normal incidence, roughness 1: D = 1/pi, G = 1, F = f0 -> f0 / (4 pi)
    got = oracle.material_bsdf([[hr.SURFACE_GGX, 0.8, 1.0] + v + n + [0, 1, 0]])[0]
    assert abs(got - 0.8 / (4 * np.pi)) < 1e-15
    # light below the surface -> 0
    assert oracle.material_bsdf([[hr.SURFACE_GGX, 0.8, 0.3] + v + n + [0, -1, 0]])[0] == 0.0


def test_resolve_quirks(oracle, hr, get_scene):
    cfg = get_scene("rtcamp6").desc.contents.config
    # uniform image: bilateral filter is the identity; Reinhard + gamma by hand
    acc = np.full((5, 7, 3), 4.0)  # 1 pass x 4 sub-pixels of radiance 1.0
    img = oracle.resolve(cfg, acc, 1)
    c = 1.5
    lum = (0.22 + 0.707 + 0.071) * c
    want = int(255.0 * (c * (lum / 900.0 + 1.0) / (lum + 1.0)) ** (1 / 2.2))
    assert (img == want).all()
    # NaN / negative -> 0, huge -> 255 (saturate + truncating cast)
    acc = np.zeros((4, 4, 3))
    acc[0, 0] = np.nan
    assert (oracle.resolve(cfg, acc, 1) == 0).all()   # NaN -> saturate -> 0 everywhere (f64::max ignores NaN)
    acc = np.zeros((4, 4, 3))
    acc[2, 2] = 1e30
    img = oracle.resolve(cfg, acc, 1)
    assert (img[2, 2] > 20).all() and img[2, 2, 0] == img.max()   # saturates to 1.0, then the bilateral filter blends it
    # negative radiance: Reinhard's (L+1) denominator flips the sign back -> bright, as in the reference
    acc = np.zeros((4, 4, 3))
    acc[1, 1] = -20.0
    img = oracle.resolve(cfg, acc, 1)
    assert (img[1, 1] > 20).all() and img[1, 1, 0] == img.max()
    # u32 wrap of the left/top neighbour (src/filter.rs:43-44): on a tiny image the far edge leaks in
    acc = np.zeros((3, 3, 3))
    acc[:, 2] = 40.0
    img = oracle.resolve(cfg, acc, 1)
    assert img[1, 0, 0] > 0  # x = 0 sees x = width-1 through the wrap, weight exp(-2^2/512)


def test_texture_edge_semantics(oracle, get_scene, hr):
    """src/texture.rs:29-63: texel corners at integers, v flip, and the u32 wrap that maps the row above the top
    texel row to the BOTTOM row."""
    scene = get_scene("rtcamp6")
    d = scene.desc.contents
    img_id = d.materials[d.elements[4].material].albedo.image  # floor: magic-circle3.png
    assert img_id >= 0
    im = d.images[img_id]
    w, h = im.width, im.height
    import ctypes
    px = np.ctypeslib.as_array(ctypes.cast(im.rgba, ctypes.POINTER(ctypes.c_uint8)), shape=(h, w, 4))

    def lin(p):
        return (p[:3] / 255.0) ** 2.2

    # exactly on a texel corner: weight 1 on texel (x, h-1-y)
    got = oracle.texture_sample(scene, img_id, (1, 1, 1), [[10 / w, 20 / h]])[0]
    assert np.allclose(got, lin(px[h - 1 - 20, 10].astype(np.float64)), rtol=1e-12)
    # v = 1: y1 = h -> h - h - 1 wraps to u32::MAX -> clamps to the bottom row (index h-1)
    got = oracle.texture_sample(scene, img_id, (1, 1, 1), [[10 / w, 1.0]])[0]
    assert np.allclose(got, lin(px[h - 1, 10].astype(np.float64)), rtol=1e-12)
    # u = 1: x clamps to w-1
    got = oracle.texture_sample(scene, img_id, (1, 1, 1), [[1.0, 20 / h]])[0]
    assert np.allclose(got, lin(px[h - 1 - 20, w - 1].astype(np.float64)), rtol=1e-12)
    # tint multiplies after the gamma
    got = oracle.texture_sample(scene, img_id, (0.5, 2.0, 0.0), [[10 / w, 20 / h]])[0]
    assert np.allclose(got, lin(px[h - 1 - 20, 10].astype(np.float64)) * [0.5, 2.0, 0.0], rtol=1e-12)


def test_skybox_face_selection(oracle, get_scene):
    scene = get_scene("rtcamp6")
    # ties go to z (strict >, src/scene.rs:300-318): (1,1,1) samples pz
    a = oracle.skybox_sample(scene, [[1, 1, 1], [0.5, 0.5, 1.0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, -1]])
    assert np.isfinite(a).all() and (a >= 0).all() and (a <= 1.0).all()
    # -0.0 in the dominant axis selects the negative face (is_sign_positive is a sign-bit test)
    b = oracle.skybox_sample(scene, [[0.0, 0.0, 1e-300], [0.0, 0.0, -1e-300]])
    assert np.isfinite(b).all()


def test_camera_ray_dof(oracle, get_scene):
    scene = get_scene("rtcamp6")
    cam = scene.camera.contents
    r = oracle.camera_ray(scene, 480, 270, 240, 135, 0, 0, 1, dof=True)
    o, d, iters = r[:3], r[3:6], r[6]
    assert abs(np.linalg.norm(d) - 1) < 1e-14 and iters >= 1
    eye = np.array(cam.eye.tuple())
    assert np.linalg.norm(o - eye) <= cam.lens_radius + 1e-15
    p = oracle.camera_ray(scene, 480, 270, 240, 135, 0, 0, 1, dof=False)
    assert np.array_equal(p[:3], eye)
    # both pass through (nearly) the same point of the focal plane
    fo = np.array(cam.forward.tuple())
    t1 = cam.focus_distance / (d @ fo)
    t2 = cam.focus_distance / (p[3:6] @ fo)
    assert np.linalg.norm((o + d * t1) - (eye + p[3:6] * t2)) < 1e-9


def test_intersect_hand_checked(oracle, hr, assets):
    b = hr.SceneBuilder(assets)
    b.camera((0, 0, 5), (0, 0, 0))
    mat = hr.SceneBuilder.material(hr.SURFACE_DIFFUSE, albedo=(0.5, 0.5, 0.5))
    b.add_sphere((0, 0, 0), 1.0, mat)
    b.add_cuboid((-5, -2, -5), (5, -1, 5), mat)
    b.add_mesh([[-1, -1, -3], [1, -1, -3], [0, 1, -3]], [[0, 1, 2]], mat)
    b.skybox()
    scene = b.finish()
    hits = oracle.intersect(scene, [[0, 0, 5], [0, 0, 0], [0, 5, 0], [0, 0, 5], [3, 3, 5]],
                            [[0, 0, -1], [0, 0, -1], [0, -1, 0], [0, 1, 0], [0, 0, -1]])
    # 1: the sphere front at z=1 (near root only)
    assert hits[0]["hit"] == 1 and hits[0]["element"] == 0 and abs(hits[0]["distance"] - 4.0) < 1e-15
    assert np.allclose(hits[0]["normal"], [0, 0, 1])
    # 2: from INSIDE the sphere only the near root exists (t<0) -> the sphere is missed, the triangle behind it is hit
    assert hits[1]["hit"] == 1 and hits[1]["element"] == 2 and abs(hits[1]["distance"] - 3.0) < 1e-15
    assert np.allclose(hits[1]["normal"], [0, 0, 1])  # geometric normal e1 x e2, not flipped toward the ray
    assert hits[1]["face"] == 0
    # 3: down onto the sphere top
    assert hits[2]["element"] == 0 and abs(hits[2]["distance"] - 4.0) < 1e-15
    # 4: up: miss -> Intersection::empty() + sky emission
    assert hits[3]["hit"] == 0 and hits[3]["element"] == -1 and hits[3]["distance"] == 1e100 and hits[3]["roughness"] == 0.2
    assert (hits[3]["emission"] > 0).any()
    # 5: past everything except nothing (z travel above the floor top y=-1: y=3) -> miss
    assert hits[4]["hit"] == 0
    # cuboid top face: uv = (x, 1 - z) of the normalised position
    h = oracle.intersect(scene, [[2.5, 5, -2.5]], [[0, -1, 0]])[0]
    assert h["element"] == 1 and np.allclose(h["normal"], [0, 1, 0]) and abs(h["distance"] - 6.0) < 1e-15
    assert abs(h["u"] - 0.75) < 1e-15 and abs(h["v"] - 0.75) < 1e-15
