#!/usr/bin/env python
"""Times N passes of the path tracer with per-kernel CUDA-event times (tuning helper)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hanamaru_renderer_b200 as hr

def main():
    scene_name = sys.argv[1] if len(sys.argv) > 1 else "rtcamp6"
    w, h, passes = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    batch = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    scene = hr.build_scene(scene_name, hr.AssetStore.from_pack())
    dev = hr.DeviceScene(scene, 0)
    ctx = hr.RenderContext(dev, scene.camera, w, h, hr.MODE_PATHTRACING, max_batch=batch)
    ctx.render_passes(1, passes); ctx.synchronize(); ctx.clear()
    ctx.mark(0); ctx.render_passes(1, passes); ctx.mark(1); ctx.synchronize()
    ms = ctx.elapsed_ms(0, 1)
    print("%s lib=%s %dx%d x%d batch=%d: %.2f ms/pass  %.1f Msamples/s" % (scene_name, os.path.basename(os.environ.get("HNM_CORE_LIB", "default")), w, h, passes, batch, ms / passes, w * h * 4 * passes / ms / 1e3))
    if os.environ.get("HNM_WID_STATS"):
        print("   warp slots, overlapped run (generation, trace, shade): " + "  ".join("%016x" % m for m in ctx.warp_slots()[:3]))
    ctx.clear(); ctx.set_profiling(True); ctx.render_passes(1, passes); ctx.synchronize()
    print("   " + "  ".join("%s=%.2f" % (k, v[0] / passes) for k, v in ctx.kernel_times().items()))
    if os.environ.get("HNM_TRACE_STATS"):
        c = ctx.counters()
        rays = c["segments"] + c["shadow_rays"]
        print("   per path: %.3f segments %.3f shadow rays; per ray: %.2f node visits, %.3f exact tests, %.5f list overflows" % (
            c["segments"] / c["paths"], c["shadow_rays"] / c["paths"], c["node_visits"] / rays, c["prim_tests"] / rays, c["cand_overflows"] / rays))
    if os.environ.get("HNM_WID_STATS"):
        print("   warp slots, serialised run (generation, trace, shade): " + "  ".join("%016x" % m for m in ctx.warp_slots()[:3]))

if __name__ == "__main__":
    main()
