import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from shoulder_b200 import _lib
torch.cuda.set_device(0); _lib.init(0)
meshes, sweeps = bench.make_bones("cfg2", 32, 0, 2048, 360)
packed = tuple(torch.from_numpy(a).pin_memory().numpy() for a in _lib._pack(meshes, sweeps))
mask = _lib.OUT_PLANE | _lib.OUT_IXY | _lib.OUT_ITR_START | _lib.OUT_ITR_CENTERED_START | _lib.OUT_RADIAL
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b = _lib.SweepBatch(None, None, packed=packed); torch.cuda.synchronize(); t1 = time.perf_counter()
    r = b.run(mask, 360); torch.cuda.synchronize(); t2 = time.perf_counter()
    r.fetch(mask); t3 = time.perf_counter()
    r.close(); b.close(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} ms  run {1e3*(t2-t1):.2f} ms  fetch {1e3*(t3-t2):.2f} ms  free {1e3*(t4-t3):.2f} ms")
# raw pinned D2H bandwidth for reference
x = torch.empty(1326448768, dtype=torch.uint8, device='cuda'); h = torch.empty_like(x, device='cpu').pin_memory()
for _ in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()
    print(f"raw D2H 1.33 GB: {1e3*(t1-t0):.2f} ms = {1.326448768/(t1-t0):.1f} GB/s")
