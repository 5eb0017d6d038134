#!/usr/bin/env python
"""Builds assets/hanamaru_assets.hnmpack from a checkout of the reference.

The pack holds the INPUTS of the host-side scene builders in decoded form, so
that the GPU box (which has no /root/reference) can build the benchmark scenes:
  * OBJ files parsed with the product's ObjLoader restatement (src/loader.rs
    semantics, libhanamaru_host.so) -> local-space f64 vertices + u32 faces;
  * textures decoded with PIL to RGBA8 (what `image::open` + `get_pixel`
    yields in the reference, src/texture.rs:16-20,59-63).  JPEG decoders may
    differ by +-1-2 levels from the `image 0.19` crate; decoded texels are ABI
    inputs shared by the oracle and the GPU path.

Usage: python tools/make_asset_pack.py [/root/reference] [out.hnmpack]
"""
import ctypes
import os
import struct
import sys
import zlib

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

MESHES = [
    "models/bunny/bunny_wired_300.obj",
    "models/box.obj",
    "models/picture_frame.obj",
    "models/armadilo_1000.obj",
    "models/fractal_icosahedron.obj",
    "models/fractal_dodecahedron.obj",
    "models/round_brilliant.obj",
    "models/bunny/bunny_face1000.obj",
    "models/bunny/bunny_face1000_flip.obj",
    "models/dia/dia.obj",
    "models/klab_logo/klab_logo_triangle.obj",
    "models/houdini_boss.obj",
]
IMAGES = [
    "textures/2d/magic-circle3.png",
    "textures/2d/checkered_diagonal_10_0.5_1.0_512.png",
    "textures/2d/checkered_diagonal_10_0.1_0.6_512.png",
    "textures/2d/earth_inverse_2048.jpg",
    "textures/2d/MarbleFloorTiles2/TexturesCom_MarbleFloorTiles2_1024_c_diffuse.tiff",
    "textures/2d/MarbleFloorTiles2/TexturesCom_MarbleFloorTiles2_1024_roughness.png",
] + ["textures/cube/Powerlines/%s.jpg" % f for f in ("posx", "negx", "posy", "negy", "posz", "negz")]


# Cube maps of the scenes that do not use Powerlines, shipped as the reference's own JPEG FILES (12 MB instead of 200 MB of
# texels) in a second pack, assets/hanamaru_cubemaps.hnmpack; the C++ host decodes them (hanamaru_image.cpp).
CUBEMAP_FILES = ["textures/cube/%s/%s.jpg" % (d, f) for d in ("LancellottiChapel", "Ryfjallet")
                 for f in ("posx", "negx", "posy", "negy", "posz", "negz")]


def write_pack(out, entries):
    with open(out, "wb") as f:
        f.write(b"HNMPACK1")
        f.write(struct.pack("<I", len(entries)))
        for name, kind, a, b, raw in entries:
            comp = zlib.compress(raw, 9 if kind != 3 else 1)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<IIIQQ", kind, a, b, len(raw), len(comp)))
            f.write(comp)
            print("%-55s kind=%d %8d x %-8d raw=%10d comp=%9d" % (name, kind, a, b, len(raw), len(comp)))
    print("wrote", out, os.path.getsize(out), "bytes")


def decode_rgba(path):
    """RGBA8, row 0 = top: DynamicImage::get_pixel semantics (Luma -> l,l,l,255; RGB -> r,g,b,255)."""
    im = Image.open(path)
    if im.mode not in ("RGBA", "RGB", "L", "LA", "P"):
        raise SystemExit("unsupported image mode %s in %s" % (im.mode, path))
    return np.ascontiguousarray(np.asarray(im.convert("RGBA"), dtype=np.uint8))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "assets", "hanamaru_assets.hnmpack")
    host = ctypes.CDLL(os.path.join(ROOT, "hanamaru_renderer_b200", "libhanamaru_host.so"))
    host.hnmh_assets_create.restype = ctypes.c_void_p
    host.hnmh_last_error.restype = ctypes.c_char_p
    store = ctypes.c_void_p(host.hnmh_assets_create())
    host.hnmh_assets_set_root(store, ref.encode())

    entries = []
    for name in MESHES:
        nv, nf = ctypes.c_uint32(), ctypes.c_uint32()
        if host.hnmh_assets_obj_counts(store, name.encode(), ctypes.byref(nv), ctypes.byref(nf)) != 0:
            raise SystemExit(host.hnmh_last_error().decode())
        verts = np.empty(nv.value * 3, dtype=np.float64)
        faces = np.empty(nf.value * 3, dtype=np.uint32)
        host.hnmh_assets_obj_copy(store, name.encode(), verts.ctypes.data_as(ctypes.c_void_p), faces.ctypes.data_as(ctypes.c_void_p))
        entries.append((name, 1, nv.value, nf.value, verts.tobytes() + faces.tobytes()))
    for name in IMAGES:
        px = decode_rgba(os.path.join(ref, name))
        entries.append((name, 2, px.shape[1], px.shape[0], px.tobytes()))

    write_pack(out, entries)
    cube = [(name, 3, 0, 0, open(os.path.join(ref, name), "rb").read()) for name in CUBEMAP_FILES]
    write_pack(os.path.join(os.path.dirname(out), "hanamaru_cubemaps.hnmpack"), cube)


if __name__ == "__main__":
    main()
