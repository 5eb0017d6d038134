#!/usr/bin/env python
"""FAST_MATH acceptance numbers (PSNR / bias vs oracle at equal spp and seed) -- tuning helper for the test of the same name."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hanamaru_renderer_b200 as hr
from oracle_ffi import Oracle
scene = hr.build_scene("rtcamp6", hr.AssetStore.from_pack())
dev = hr.DeviceScene(scene, 0)
w, h, passes = 480, 270, 16
want, cnt = Oracle("det").render(scene, w, h, hr.MODE_PATHTRACING, 1, passes)
want_img = Oracle("det").resolve(scene.desc.contents.config, want, passes).astype(float)
ctx = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING)
ctx.set_precision(1)
ctx.render_passes(1, passes); ctx.synchronize()
img = ctx.resolve(passes).astype(float); acc = ctx.read_accum(); c = ctx.counters()
mse = ((img - want_img) ** 2).mean()
rel = np.linalg.norm(acc - want, axis=2) / np.maximum(np.linalg.norm(want, axis=2), 1e-300)
print("lib=%s PSNR %.2f dB bias %s median rel %.3g  >1e-3: %d  segments %d vs %d  mean accum ratio %s" % (
    os.path.basename(os.environ.get("HNM_CORE_LIB", "default")), 10 * np.log10(255 ** 2 / mse), np.round((img - want_img).mean(axis=(0, 1)), 3).tolist(),
    float(np.median(rel)), int((rel > 1e-3).sum()), c["segments"], cnt["segments"], np.round(acc.mean(axis=(0, 1)) / want.mean(axis=(0, 1)), 5).tolist()))
