"""The C++ host's image I/O (csrc/host/hanamaru_image.cpp): `image::open` (src/texture.rs:18) and `DynamicImage::save`
(src/renderer.rs:97, src/main.rs:1217) of the reference.  PNG is lossless: bit-exact against PIL.  JPEG decoding is an
implementation choice (IDCT, chroma upsampling, colour conversion): bounded against libjpeg-turbo (PIL)."""
import io
import os

import numpy as np
import pytest
from PIL import Image


def synthetic(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(xx / 9.0) * np.cos(yy / 13.0), 128 + 90 * np.cos(xx / 5.0 + yy / 17.0), (xx * 3 + yy * 2) % 256], axis=2)
    return np.clip(base + rng.normal(0, 6, base.shape), 0, 255).astype(np.uint8)


def test_png_encode_is_read_back_exactly(hr, tmp_path):
    for (h, w) in ((1, 1), (37, 53), (270, 480)):
        img = synthetic(h, w, h)
        p = tmp_path / ("x_%d.png" % h)
        hr.save_png(p, img)
        back = Image.open(p)
        assert back.mode == "RGB" and np.array_equal(np.asarray(back), img)
        assert np.array_equal(hr.image_decode(p.read_bytes())[..., :3], img)


@pytest.mark.parametrize("mode,bits", [("RGB", 8), ("RGBA", 8), ("L", 8), ("LA", 8), ("P", 8), ("1", 1), ("I;16", 16)])
def test_png_decode_matches_pil(hr, mode, bits):
    rgb = synthetic(61, 47, 7)
    if mode == "RGB":
        im = Image.fromarray(rgb)
    elif mode == "RGBA":
        im = Image.fromarray(np.dstack([rgb, 255 - rgb[..., 0]]))
    elif mode == "L":
        im = Image.fromarray(rgb[..., 0])
    elif mode == "LA":
        im = Image.fromarray(np.dstack([rgb[..., 0], rgb[..., 1]]), "LA")
    elif mode == "P":
        im = Image.fromarray(rgb).quantize(37)
    elif mode == "1":
        im = Image.fromarray(rgb[..., 0] > 128)
    else:
        im = Image.fromarray((rgb[..., 0].astype(np.uint16) << 8) | rgb[..., 1], "I;16")
    for opt in (False, True):
        buf = io.BytesIO()
        im.save(buf, "PNG", optimize=opt)
        got = hr.image_decode(buf.getvalue())
        if mode == "I;16":
            want = np.dstack([rgb[..., 0]] * 3 + [np.full(rgb.shape[:2], 255, np.uint8)])  # the high byte, like image 0.19's to_rgba
        else:
            want = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGBA"))
        assert np.array_equal(got, want), (mode, opt)


def test_png_decode_rejects_garbage(hr):
    with pytest.raises(hr.HanamaruError):
        hr.image_decode(b"\x89PNG\r\n\x1a\n" + b"\0" * 40)
    with pytest.raises(hr.HanamaruError):
        hr.image_decode(b"GIF89a" + b"\0" * 40)
    buf = io.BytesIO()
    Image.fromarray(synthetic(16, 16, 1)).save(buf, "PNG")
    b = bytearray(buf.getvalue())
    b[60] ^= 0xFF   # corrupt the IDAT: CRC mismatch
    with pytest.raises(hr.HanamaruError):
        hr.image_decode(bytes(b))


@pytest.mark.parametrize("subsampling,quality,size", [(2, 90, (96, 128)), (2, 75, (101, 67)), (1, 85, (64, 80)), (0, 95, (33, 47)), (2, 50, (8, 8))])
def test_jpeg_decode_close_to_libjpeg(hr, subsampling, quality, size):
    """Baseline JPEGs written by PIL (4:2:0 like every JPEG under the reference's textures/, 4:2:2, 4:4:4; odd sizes that do
    not fill whole MCUs): the host decoder against libjpeg-turbo, per channel value."""
    img = synthetic(size[0], size[1], subsampling * 100 + quality)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, "JPEG", quality=quality, subsampling=subsampling)
    got = hr.image_decode(buf.getvalue())
    want = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGBA"))
    assert got.shape == want.shape
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 3 and (d <= 1).mean() >= 0.98, (int(d.max()), float((d <= 1).mean()))


def test_jpeg_greyscale_and_restart_markers(hr):
    img = synthetic(72, 90, 3)
    buf = io.BytesIO()
    Image.fromarray(img[..., 0]).save(buf, "JPEG", quality=90)
    got = hr.image_decode(buf.getvalue())
    want = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGBA"))
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 2
    buf = io.BytesIO()
    try:
        Image.fromarray(img).save(buf, "JPEG", quality=90, subsampling=2, restart_marker_blocks=3)
    except TypeError:
        pytest.skip("this PIL cannot write restart markers")
    got = hr.image_decode(buf.getvalue())
    want = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGBA"))
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 3


def test_progressive_jpeg_is_refused_not_misdecoded(hr):
    buf = io.BytesIO()
    Image.fromarray(synthetic(32, 32, 2)).save(buf, "JPEG", progressive=True)
    with pytest.raises(hr.HanamaruError, match="progressive"):
        hr.image_decode(buf.getvalue())


def test_cubemap_pack_decodes_and_scenes_build(hr, assets):
    """assets/hanamaru_cubemaps.hnmpack ships the reference's LancellottiChapel / Ryfjallet JPEG FILES; the host decodes them on
    first use, so every scene builder of src/main.rs runs under its OWN cube map (round 1 substituted Powerlines)."""
    if not os.path.exists(hr.CUBEMAP_PACK):
        pytest.skip("cube-map pack not present")
    for name, images in (("rtcamp6_v4", 6), ("rtcamp5", 9), ("simple", 8)):
        scene = hr.build_scene(name, assets)
        d = scene.desc.contents
        assert d.num_images == images
        face = d.images[d.skybox_images[0]]
        assert (face.width, face.height) == (2048, 2048)


@pytest.mark.skipif(not os.path.isdir("/root/reference/textures"), reason="needs the reference checkout")
def test_reference_textures_decode_like_pil(hr):
    for p, exact in (("textures/cube/Powerlines/posx.jpg", False), ("textures/cube/Ryfjallet/negy.jpg", False),
                     ("textures/2d/magic-circle3.png", True), ("textures/2d/checkered_diagonal_10_0.5_1.0_512.png", True),
                     ("textures/cube/pisa/px.png", True), ("textures/2d/earth_inverse_2048.jpg", False)):
        data = open(os.path.join("/root/reference", p), "rb").read()
        got = hr.image_decode(data)
        want = np.asarray(Image.open(io.BytesIO(data)).convert("RGBA"))
        d = np.abs(got.astype(int) - want.astype(int))
        if exact:
            assert d.max() == 0, p
        else:
            assert d.max() <= 3 and (d == 0).mean() > 0.99, (p, int(d.max()), float((d == 0).mean()))
