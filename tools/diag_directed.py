#!/usr/bin/env python
"""Diagnostics for tests/test_gpu_edge_cases.py::test_directed_rays_match_oracle: mismatches per ray category."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hanamaru_renderer_b200 as hr
from oracle_ffi import Oracle
import test_gpu_edge_cases as T

name = sys.argv[1] if len(sys.argv) > 1 else "rtcamp6"
scene = hr.build_scene(name, hr.AssetStore.from_pack())
dev = hr.DeviceScene(scene, 0)
oracle = Oracle("det")
rng = np.random.default_rng(21)
o, d = T.directed_rays(scene, rng)
got, want = dev.intersect(o, d), oracle.intersect(scene, o, d)
n_each = 4000
bad = np.zeros(len(o), bool)
for f in got.dtype.names:
    g, w = got[f], want[f]
    same = (g.view(np.uint64) == w.view(np.uint64)) if g.dtype == np.float64 else (g == w)
    bad |= ~same.reshape(len(g), -1).all(axis=1)
print(name, "rays", len(o), "bad", int(bad.sum()))
for k in range(0, len(o), n_each):
    b = bad[k:k + n_each]
    if b.any():
        print("  block %2d (rays %d..): %d bad" % (k // n_each, k, int(b.sum())))
idx = np.nonzero(bad)[0]
np.set_printoptions(precision=17, linewidth=200)
for i in idx[:12]:
    print("ray", i, "o", o[i], "d", d[i])
    print("   gpu : hit %d elem %d face %d dist %r uv (%r, %r) n %s" % (got["hit"][i], got["element"][i], got["face"][i], got["distance"][i], got["u"][i], got["v"][i], got["normal"][i]))
    print("   orcl: hit %d elem %d face %d dist %r uv (%r, %r) n %s" % (want["hit"][i], want["element"][i], want["face"][i], want["distance"][i], want["u"][i], want["v"][i], want["normal"][i]))
for i in idx[len(idx)//2: len(idx)//2 + 6]:
    print("ray", i, "o", o[i], "d", d[i])
    print("   gpu : hit %d elem %d face %d dist %r uv (%r, %r)" % (got["hit"][i], got["element"][i], got["face"][i], got["distance"][i], got["u"][i], got["v"][i]))
    print("   orcl: hit %d elem %d face %d dist %r uv (%r, %r)" % (want["hit"][i], want["element"][i], want["face"][i], want["distance"][i], want["u"][i], want["v"][i]))
