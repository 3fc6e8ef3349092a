"""cfg5 front end, time inside each C-ABI call vs around it (Python glue): python tools/c5_probe2.py [bones]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bench
from shoulder_b200 import _lib, features
g = bench.Gpu(0, 1, 0)
bones = int(sys.argv[1]) if len(sys.argv) > 1 else 32
forest = features.Forest.from_arrays(np.load("tests/golden/forest_rfc_bg3.npz"))
lib = _lib.load()
CT = {}
class Timed:
    def __init__(self, name, fn): self.name, self.fn = name, fn
    def __call__(self, *a):
        t0 = time.perf_counter(); r = self.fn(*a); CT[self.name] = CT.get(self.name, 0) + time.perf_counter() - t0; return r
class LibProxy:
    def __getattr__(self, k):
        f = getattr(lib, k)
        return Timed(k, f) if k.startswith("shb_") else f
_lib.load = lambda: LibProxy()
features._lib.load = _lib.load
meshes, sweeps = bench.make_bones("cfg4", bones, 0, 0, 0)
req = bench.consumer_requests(sweeps)
packed = tuple(g.torch.from_numpy(a).pin_memory().numpy() for a in _lib._pack(meshes, sweeps))
full = [3 * b for b in range(bones)]; prox = [3 * b + 2 for b in range(bones)]
zs_of = lambda s: np.asarray(sweeps[s][2]) + sweeps[s][1]
zs_g = [zs_of(s)[150:480] for s in prox]
T = {}
def tick(name, t0):
    g.torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.perf_counter() - t0; return time.perf_counter()
for it in range(6):
    if it == 1: T.clear(); CT.clear()
    t = time.perf_counter()
    res = _lib.sweep_batch(None, None, _lib.OUT_PLANE, 0, packed=packed, lazy=True, requests=req); t = tick("sweep_batch", t)
    cz = np.stack([np.c_[res.array(_lib.ARR_CENTROID, s)[50:130], zs_of(s)[50:130]] for b, s in enumerate(full)])
    mid = cz.mean(axis=1, keepdims=True); dirn = np.linalg.svd(cz - mid)[2][:, 0, :]
    axes = np.stack([mid[:, 0] + dirn * 50, mid[:, 0] - dirn * 50], axis=1)
    t = tick("canal", t)
    ft = features.groove_features(res, prox, zs_g, axes); t = tick("groove_features", t)
    cuts = np.cumsum([0] + [len(f["X"]) for f in ft])
    proba = forest.predict_proba(np.vstack([f["X"] for f in ft])); t = tick("forest", t)
    bg = features.groove_theta_batch([f["peak_theta"] for f in ft], [proba[a:b, 1] for a, b in zip(cuts[:-1], cuts[1:])]); t = tick("theta", t)
    pts = features.groove_points(res, prox, zs_g, bg, 512); t = tick("points", t)
    imgs = features.neck_image(res, prox, bg); t = tick("neck_image", t)
    res.close(); t = tick("close", t)
print({k: round(1e3 * v / 5, 2) for k, v in T.items()}, "ms per step of", bones, "bones; total", round(1e3 * sum(T.values()) / 5, 2))
print("inside C calls:", {k: round(1e3 * v / 5, 2) for k, v in CT.items()})
