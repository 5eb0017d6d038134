#!/usr/bin/env python
"""Small renders through every kernel, for `compute-sanitizer --tool memcheck python tools/sanitize_run.py`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hanamaru_renderer_b200 as hr  # noqa: E402

a = hr.AssetStore.from_pack()
for name in ("rtcamp6", "tbf3_pl"):
    s = hr.build_scene(name, a)
    d = hr.DeviceScene(s, 0)
    c = hr.RenderContext(d, s.camera, 160, 90, hr.MODE_PATHTRACING)
    c.render_passes(1, 3)
    c.synchronize()
    print(name, c.resolve(3).mean(), c.counters())
    c.close()
    c = hr.RenderContext(d, s.camera, 97, 61, hr.MODE_PATHTRACING, shard=(1, 3, 4))
    c.render_passes(2, 2)
    c.synchronize()
    print(name, "shard", c.read_accum().sum())
    c.close()
    c = hr.RenderContext(d, s.camera, 64, 36, hr.MODE_DEBUG_SHADING)
    c.render_passes(1, 1)
    c.synchronize()
    print(name, "debug", c.resolve(1).mean())
    c.close()
    o = np.random.default_rng(0).normal(size=(5000, 3)) * 3 + [0, 2, 0]
    dd = -o / np.linalg.norm(o, axis=1, keepdims=True)
    print(name, "hit rate", d.intersect(o, dd)["hit"].mean())
