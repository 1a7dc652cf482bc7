import sys, os
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import povray_b200 as pv
import oracle_lib
from oracle_lib import RAY_DTYPE
W,H=96,54
name=sys.argv[1]
base='tools/dbg' if os.path.exists(f'tools/dbg/{name}.pvs') else 'tests/golden'
rays=np.fromfile(f'{base}/{name}.rays',dtype=RAY_DTYPE).reshape(H,W)
rgbt=np.fromfile(f'{base}/{name}.rgbt',dtype=np.float32).reshape(H,W,4)
s=pv.Scene.load(f'{base}/{name}.pvs').finalize(0)
img,st=s.render_image(W,H)
d=np.abs(img-rgbt).max(axis=2)
bad=d>1/255
print('bad',bad.sum())
for ob in np.unique(rays['obj']):
    m=rays['obj']==ob
    print(ob, m.sum(), (bad&m).sum(), d[m].max())
ys,xs=np.where(bad)
for y,x in list(zip(ys,xs))[:6]: print(y,x,rays['obj'][y,x],img[y,x],rgbt[y,x])
