#!/usr/bin/env python
"""How much would ray sorting buy?  Traces the same incoherent rays unsorted and sorted by (origin Morton cell,
direction octant / finer direction key); run under `ncu --metrics gpu__time_duration.sum -k regex:k_trace`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hanamaru_renderer_b200 as hr  # noqa: E402


def morton3(q, bits):
    out = np.zeros(len(q), np.uint64)
    for b in range(bits):
        for k in range(3):
            out |= ((q[:, k] >> b) & 1).astype(np.uint64) << np.uint64(3 * b + k)
    return out


def main():
    a = hr.AssetStore.from_pack()
    s = hr.build_scene("rtcamp6", a)
    d = hr.DeviceScene(s, 0)
    rng = np.random.default_rng(3)
    n = 4_000_000
    # secondary-ray-like: origins on surfaces hit by camera-ish rays, cosine-ish random directions
    cam = np.array(s.camera.contents.eye.tuple())
    tgt = rng.uniform(-3, 3, size=(n, 3)) * [1, 0.4, 1] + [0, 0.6, 0]
    d0 = tgt - cam
    d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
    h = d.intersect(np.tile(cam, (n, 1)), d0)
    ok = h["hit"] == 1
    o = (h["position"] + h["normal"] * 1e-4)[ok]
    nrm = h["normal"][ok]
    r = rng.normal(size=o.shape)
    r /= np.linalg.norm(r, axis=1, keepdims=True)
    dirs = nrm + r
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    print("rays", len(o))
    d.intersect(o, dirs)                      # launch 2: unsorted (order of the camera rays = random)
    lo, hi = o.min(0), o.max(0)
    for bits, dbits in ((4, 0), (6, 0), (6, 1), (8, 1)):
        q = np.minimum(((o - lo) / (hi - lo) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
        key = morton3(q, bits)
        if dbits:
            octant = (dirs[:, 0] < 0) * 1 + (dirs[:, 1] < 0) * 2 + (dirs[:, 2] < 0) * 4
            key = key * np.uint64(8) + octant.astype(np.uint64)
        idx = np.argsort(key, kind="stable")
        d.intersect(o[idx], dirs[idx])        # launches 3..: sorted
        print("sorted bits", bits, dbits)


if __name__ == "__main__":
    main()
