import sys, cProfile, pstats
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import single_bone as sb
from shoulder_b200 import _lib, meshio
from shoulder_b200.slice import GpuDistalSlices, GpuFullSlices, GpuProximalSlices, run_batch
_lib.init(0)
obb = meshio.PcaObb(ROOT / "tests" / "golden" / "bones" / "humerus_left.npz")
def once():
    full, dist, prox = GpuFullSlices(obb), GpuDistalSlices(obb), GpuProximalSlices(obb, sb.Neck())
    run_batch([full, dist, prox]); return sb.touch(full, dist, prox)
for _ in range(3): once()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): once()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
