import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import hanamaru_renderer_b200 as hr
from oracle_ffi import Oracle
oracle = Oracle("det")
a = hr.AssetStore.from_pack()
s = hr.build_scene("rtcamp6", a)
d = hr.DeviceScene(s, 0)
rng = np.random.default_rng(9)
n = 20000
o = rng.uniform(-4, 4, size=(n, 3)) * [1, 0.5, 1] + [0, 1.0, 0]
dd = rng.normal(size=(n, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
g = d.intersect(o, dd); w = oracle.intersect(s, o, dd)
for f in g.dtype.names:
    bad = (g[f] != w[f]) if g[f].ndim == 1 else (g[f] != w[f]).any(axis=1)
    print(f, int(bad.sum()))
bad = (g["emission"] != w["emission"]).any(axis=1)
print("bad among hits", int((bad & (g["hit"] == 1)).sum()), "bad among misses", int((bad & (g["hit"] == 0)).sum()), "misses", int((g["hit"] == 0).sum()))
i = np.nonzero(bad)[0][:5]
print(g["emission"][i]); print(w["emission"][i]); print((g["emission"][i] != w["emission"][i]))
