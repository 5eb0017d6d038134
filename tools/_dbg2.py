import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import hanamaru_renderer_b200 as hr
from oracle_ffi import Oracle
o=Oracle('det'); a=hr.AssetStore.from_pack(); s=hr.build_scene('rtcamp6',a); dev=hr.DeviceScene(s,0)
rng=np.random.default_rng(9)
n=4096
d=rng.normal(size=(n,3)); d[:,1]=np.abs(d[:,1]); d/=np.linalg.norm(d,axis=1,keepdims=True)
oo=np.tile([0,3.0,0],(n,1))
got=dev.intersect(oo,d); want=o.intersect(s,oo,d)
for f in got.dtype.names:
    a1=np.ascontiguousarray(got[f]); b1=np.ascontiguousarray(want[f])
    print(f, a1.tobytes()==b1.tobytes())
bad=(got['emission']!=want['emission'])
print('bad per channel', bad.sum(axis=0), 'of', n)
i=np.nonzero(bad.any(axis=1))[0][:5]
print(got['emission'][i]); print(want['emission'][i]); print(d[i])
m=got['hit']==0
print('param(e.z copy)==oracle z', np.mean(got['param'][m]==want['emission'][m,2]), ' emission.z==oracle', np.mean(got['emission'][m,2]==want['emission'][m,2]))
print('sky_b as seen by kernel:', np.unique(got['u'][m])[:5], ' sky_g:', np.unique(got['v'][m])[:5])
