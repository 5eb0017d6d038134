#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference checkout and from the oracle.

  rtcamp6_golden_480x270.png   the reference's only golden output, rtcamp6_1000x4spp.png (1920x1080,
                               1000 passes x 4 spp), box-downsampled 4x4 so that it is small enough to commit
  rtcamp5_golden_480x270.png   rtcamp5.png (1920x1080), the same way
  oracle_vectors.npz           outputs of the oracle (both flavours agree on these) on fixed inputs:
                               regression pins for the oracle itself, and inputs the GPU tests replay
Usage: python tools/make_golden.py [/root/reference]
"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    gold = np.asarray(Image.open(os.path.join(ref, "rtcamp6_1000x4spp.png")).convert("RGB"), dtype=np.float64)
    small = gold.reshape(270, 4, 480, 4, 3).mean(axis=(1, 3))
    Image.fromarray(np.clip(np.rint(small), 0, 255).astype(np.uint8)).save(os.path.join(out, "rtcamp6_golden_480x270.png"))

    # the reference's second published image: init_scene_rtcamp5 (45 diamonds placed by StdRng::gen_range)
    gold5 = np.asarray(Image.open(os.path.join(ref, "rtcamp5.png")).convert("RGB"), dtype=np.float64)
    small5 = gold5.reshape(270, 4, 480, 4, 3).mean(axis=(1, 3))
    Image.fromarray(np.clip(np.rint(small5), 0, 255).astype(np.uint8)).save(os.path.join(out, "rtcamp5_golden_480x270.png"))

    import hanamaru_renderer_b200 as hr
    from oracle_ffi import Oracle
    oracle = Oracle("det")
    assets = hr.AssetStore.from_pack()
    scene = hr.build_scene("rtcamp6", assets)
    vec = {}
    # path tracing, 64x36, passes 1..2: the f64 accumulation buffer (det flavour)
    acc, cnt = oracle.render(scene, 64, 36, hr.MODE_PATHTRACING, 1, 2)
    vec["pt_64x36_s2_accum"] = acc
    vec["pt_64x36_s2_counters"] = np.array([cnt[k] for k in ("paths", "segments", "shadow_rays", "lens_iters")], np.uint64)
    vec["pt_64x36_s2_rgb8"] = oracle.resolve(scene.desc.contents.config, acc, 2)
    for mode, name in ((hr.MODE_DEBUG_NORMAL, "normal"), (hr.MODE_DEBUG_DEPTH, "depth"), (hr.MODE_DEBUG_FOCALPLANE, "focal"),
                       (hr.MODE_DEBUG_SHADING, "shading")):
        a, _ = oracle.render(scene, 96, 54, mode, 1, 1)
        vec["debug_%s_96x54_rgb8" % name] = oracle.resolve(scene.desc.contents.config, a, 1)
    # the random stream of the first camera path of the image (pixel 0,0 sub 0,0, pass 1)
    vec["isaac_first_path"] = oracle.isaac64([8700304, 1, 223146, 300912], 32)
    np.savez_compressed(os.path.join(out, "oracle_vectors.npz"), **vec)
    print("wrote", out, {k: v.shape for k, v in vec.items()})


if __name__ == "__main__":
    main()
