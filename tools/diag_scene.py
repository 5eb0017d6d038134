#!/usr/bin/env python
"""Whole-path parity of one scene against the oracle, with a breakdown of where the differences are (tuning helper)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hanamaru_renderer_b200 as hr
from oracle_ffi import Oracle
name, w, h, first, count = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
scene = hr.build_scene(name, hr.AssetStore.from_pack())
dev = hr.DeviceScene(scene, 0)
want, cnt = Oracle("det").render(scene, w, h, hr.MODE_PATHTRACING, first, count)
for rep in range(2):
    ctx = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING)
    ctx.render_passes(first, count); ctx.synchronize()
    got = ctx.read_accum(); c = ctx.counters()
    if os.environ.get("HNM_DIAG_QUEUES"):
        qc = ctx.queue_counters()
        print("   per bounce [rays miss delta nee events shadow]:", " | ".join("b%d %s" % (b, qc[b, :6].tolist()) for b in range(1, 11)))
    ctx.close()
    bad = (got.view(np.uint64) != want.view(np.uint64))
    print("%s lib=%s rep %d: %d bad pixels; per channel %s; segments %d vs %d, shadow %d vs %d; rows with bad pixels: %d; max rel %.3g" % (
        name, os.path.basename(os.environ.get("HNM_CORE_LIB", "default")), rep, int(bad.any(axis=2).sum()), bad.sum(axis=(0, 1)).tolist(),
        c["segments"], cnt["segments"], c["shadow_rays"], cnt["shadow_rays"], int(bad.any(axis=(1, 2)).sum()),
        float((np.abs(got - want) / np.maximum(np.abs(want), 1e-300)).max())))
