#!/usr/bin/env python
"""One pass of a scene at a given size, kernels serialised (HNM_RNG_OVERLAP=0): the workload bench.py profiles under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` to MEASURE the dominant kernel's DRAM traffic in the run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hanamaru_renderer_b200 as hr  # noqa: E402

name, w, h = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
scene = hr.build_scene(name, hr.AssetStore.from_pack())
dev = hr.DeviceScene(scene, 0)
ctx = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING, max_batch=1)
ctx.render_passes(1, 1)
ctx.synchronize()
print(ctx.counters())
