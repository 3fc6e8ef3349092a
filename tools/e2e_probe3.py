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
    for it in range(4):
        torch.cuda.synchronize(); T0 = time.perf_counter(); log = []
        parts = []
        for c in chunks:
            t0 = time.perf_counter(); b = _lib.SweepBatch(None, None, packed=c); t1 = time.perf_counter()
            r = b.run(mask, 360); t2 = time.perf_counter(); b.close(); r.fetch_async(mask); t3 = time.perf_counter()
            parts.append(r); log.append((t1 - t0, t2 - t1, t3 - t2))
        t5 = time.perf_counter()
        waits = []
        for p in parts:
            w0 = time.perf_counter(); p.fetch(mask); waits.append(time.perf_counter() - w0)
        t6 = time.perf_counter()
        for p in parts: p.close()
        torch.cuda.synchronize(); T1 = time.perf_counter()
    a = np.array(log) * 1e3
    print(f"K={K}: total {1e3*(T1-T0):.2f} ms; enqueue phase {1e3*(t5-T0):.2f} (create {a[:,0].sum():.2f} run {a[:,1].sum():.2f} async {a[:,2].sum():.2f}); waits {[round(1e3*w,2) for w in waits]}; close {1e3*(T1-t6):.2f}")
