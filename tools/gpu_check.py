#!/usr/bin/env python
"""Exploratory GPU-vs-oracle comparison with timings (not a test; see tests/ for the parity suite)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hanamaru_renderer_b200 as hr  # noqa: E402
from oracle_ffi import Oracle  # noqa: E402


def main():
    scene_name = sys.argv[1] if len(sys.argv) > 1 else "rtcamp6"
    w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (480, 270)
    passes = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    oracle = Oracle("det")
    assets = hr.AssetStore.from_pack()
    scene = hr.build_scene(scene_name, assets)
    print("scene", scene_name, scene.counts())
    t = time.time()
    dev = hr.DeviceScene(scene, 0)
    print("scene upload %.3f s" % (time.time() - t))

    # unit: isaac
    seeds = np.array([[8700304, 1, 403480, 401216], [1, 23, 456, 7890]], np.uint64)
    g = hr.isaac64_batch(seeds, 32)
    o = np.stack([oracle.isaac64(s, 32) for s in seeds])
    print("isaac tail equal:", np.array_equal(g, o))
    g = hr.isaac64_batch(seeds, 600)
    o = np.stack([oracle.isaac64(s, 600) for s in seeds])
    print("isaac full (refill) equal:", np.array_equal(g, o))

    # unit: math
    rng = np.random.default_rng(0)
    x = rng.random(100000) * 2 * np.pi
    for fn, name in ((0, "sin"), (1, "cos")):
        print(name, "equal:", np.array_equal(hr.math_batch(fn, x), oracle.math(fn, x)))
    x = rng.random(100000)
    print("pow equal:", np.array_equal(hr.math_batch(3, x, np.full_like(x, 2.2)), oracle.math(3, x, np.full_like(x, 2.2))))
    print("exp equal:", np.array_equal(hr.math_batch(2, -x * 700), oracle.math(2, -x * 700)))
    print("acos equal:", np.array_equal(hr.math_batch(4, x * 2 - 1), oracle.math(4, x * 2 - 1)))

    # debug modes
    for mode, name in ((hr.MODE_DEBUG_NORMAL, "normal"), (hr.MODE_DEBUG_DEPTH, "depth"), (hr.MODE_DEBUG_FOCALPLANE, "focal"),
                       (hr.MODE_DEBUG_SHADING, "shading")):
        ctx = hr.RenderContext(dev, scene.camera, w, h, mode)
        ctx.render_passes(1, 1)
        ctx.synchronize()
        got = ctx.read_accum()
        img = ctx.resolve(1)
        want, _ = oracle.render(scene, w, h, mode, 1, 1)
        wimg = oracle.resolve(scene.desc.contents.config, want, 1)
        print("debug %-8s accum bit-equal: %s (diff px %d)  u8 equal: %s (diff %d)" % (
            name, np.array_equal(got.view(np.uint64), want.view(np.uint64)), int((got != want).any(axis=2).sum()),
            np.array_equal(img, wimg), int((img != wimg).sum())))
        ctx.close()

    # path tracing
    ctx = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING)
    ctx.set_profiling(True)
    t = time.time()
    ctx.render_passes(1, passes)
    ctx.synchronize()
    dt = time.time() - t
    got = ctx.read_accum()
    print("gpu render %dx%d x%d passes: %.3f s  %.2f Msamples/s" % (w, h, passes, dt, w * h * 4 * passes / dt / 1e6))
    print("counters", ctx.counters())
    for k, v in ctx.kernel_times().items():
        print("   %-14s %9.3f ms  %4d launches" % (k, v[0], v[1]))
    t = time.time()
    want, cnt = oracle.render(scene, w, h, hr.MODE_PATHTRACING, 1, passes)
    dt = time.time() - t
    print("oracle render: %.3f s  %.3f Msamples/s" % (dt, w * h * 4 * passes / dt / 1e6), cnt)
    eq = np.array_equal(got.view(np.uint64), want.view(np.uint64))
    bad = (got != want).any(axis=2)
    print("path accum bit-equal:", eq, "diff px", int(bad.sum()), "of", w * h)
    if bad.any():
        ys, xs = np.nonzero(bad)
        for y, x in list(zip(ys, xs))[:10]:
            print("   px", x, y, got[y, x], want[y, x])
    img = ctx.resolve(passes)
    wimg = oracle.resolve(scene.desc.contents.config, want, passes)
    print("resolved u8 equal:", np.array_equal(img, wimg), "diff", int((img != wimg).sum()))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    from PIL import Image
    Image.fromarray(img).save(os.path.join(ROOT, "gpurun_out", "gpu_%s_%dx%d_%d.png" % (scene_name, w, h, passes)))


if __name__ == "__main__":
    main()
