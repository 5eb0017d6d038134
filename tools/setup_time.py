#!/usr/bin/env python
"""Where the e2e set-up time goes: scene build (host), hnm_scene_create, hnm_renderer_create, first pass."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hanamaru_renderer_b200 as hr  # noqa: E402

t = time.perf_counter()
a = hr.AssetStore.from_pack()
t1 = time.perf_counter()
s = hr.build_scene("rtcamp6", a)
t2 = time.perf_counter()
hr.device_count()
t3 = time.perf_counter()
d = hr.DeviceScene(s, 0)
t4 = time.perf_counter()
c = hr.RenderContext(d, s.camera, 1920, 1080, hr.MODE_PATHTRACING)
t5 = time.perf_counter()
c.render_passes(1, 2)
c.synchronize()
t6 = time.perf_counter()
c.render_passes(3, 2)
c.synchronize()
t7 = time.perf_counter()
print("pack %.3f  host scene %.3f  cuda init %.3f  scene_create %.3f  renderer_create %.3f  first batch %.3f  second batch %.3f" % (
    t1 - t, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6))
d2 = hr.DeviceScene(s, 0)
t8 = time.perf_counter()
c2 = hr.RenderContext(d2, s.camera, 1920, 1080, hr.MODE_PATHTRACING)
t9 = time.perf_counter()
print("again: scene_create %.3f  renderer_create %.3f" % (t8 - t7, t9 - t8))
