"""cfg5 front end through the fused call: time inside each C-ABI call: python tools/c5_probe3.py [bones]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bench
from shoulder_b200 import _lib, features
g = bench.Gpu(0, 1, 0)
bones = int(sys.argv[1]) if len(sys.argv) > 1 else 32
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
orig = bench.landmark_record
t0 = time.perf_counter()
rec = bench.landmark_record(g, bones, 10)
print(rec["ms_per_step"], "ms per step,", rec["value"], "bones/s")
print("inside C calls, ms per step (12 steps incl. 2 warm-up):", {k: round(1e3 * v / 12, 2) for k, v in CT.items()})
