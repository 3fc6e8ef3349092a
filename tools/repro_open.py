import sys, numpy as np
sys.path.insert(0, '.')
from shoulder_b200 import _lib, meshio
_lib.init(0)
v, f = meshio.icosphere(2, 10.0)
keep = np.ones(len(f), dtype=bool)
keep[np.argsort(v[f].mean(axis=1)[:, 0])[-40:]] = False
v2, f2 = meshio.icosphere(2, 7.0)
vv = np.vstack([v, v2 + np.array([-30.0, 1.0, 0.5])]); ff = np.vstack([f[keep], f2 + len(v)])
zs = np.linspace(8.0, -8.0, 9)
for mask in (_lib.OUT_PLANE | _lib.OUT_SEGMENTS | _lib.OUT_CONTOURS | _lib.OUT_ALL_PROFILES, _lib.OUT_PLANE | _lib.OUT_CONTOURS | _lib.OUT_IXY):
    r = _lib.sweep_batch([(vv, ff)], [(0, float(zs.mean()), zs - zs.mean(), 16)], mask)
    print(r.array(_lib.ARR_STATUS), r.array(_lib.ARR_N_ENT), r.array(_lib.ARR_N_SEG))
    r.close()
