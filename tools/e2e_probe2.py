import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from shoulder_b200 import _lib
torch.cuda.set_device(0); _lib.init(0)
meshes, sweeps = bench.make_bones("cfg2", 32, 0, 2048, 360)
packed = tuple(torch.from_numpy(a).pin_memory().numpy() for a in _lib._pack(meshes, sweeps))
mask = _lib.OUT_PLANE | _lib.OUT_IXY | _lib.OUT_ITR_START | _lib.OUT_ITR_CENTERED_START | _lib.OUT_RADIAL
for K in (1, 4, 8):
    chunks, first = _lib.split_packed(packed, K)
    chunks = [tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in c) for c in chunks]
    for it in range(3):
        torch.cuda.synchronize(); T0 = time.perf_counter(); log = []
        parts, prev = [], None
        for c in chunks:
            t0 = time.perf_counter(); b = _lib.SweepBatch(None, None, packed=c); t1 = time.perf_counter()
            r = b.run(mask, 360); t2 = time.perf_counter(); b.close(); t3 = time.perf_counter()
            if prev is not None:
                prev.fetch(mask); parts.append(prev)
            t4 = time.perf_counter(); prev = r
            log.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
        t5 = time.perf_counter(); prev.fetch(mask); parts.append(prev); t6 = time.perf_counter()
        for p in parts: p.close()
        torch.cuda.synchronize(); T1 = time.perf_counter()
    a = np.array(log) * 1e3
    print(f"K={K}: total {1e3*(T1-T0):.2f} ms; per chunk create {a[:,0].mean():.2f} run {a[:,1].mean():.2f} bclose {a[:,2].mean():.2f} fetch(prev) {a[:,3].mean():.2f}; last fetch {1e3*(t6-t5):.2f}; close-all {1e3*(T1-t6):.2f}")
