"""BASELINE config 1 shape: ONE bone, the three default sweeps of bone.Humerus (200x100, 200x500, 600x512),
through the Python Slices mirror.  Prints GPU end-to-end latency and the CPU restatement's time."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle
from shoulder_b200 import _lib, meshio
from shoulder_b200.slice import GpuDistalSlices, GpuFullSlices, GpuProximalSlices, run_batch


class Neck:
    neck_z = 60.0


def touch(full, dist, prox):
    # what the landmark code reads: canal.py:40-46, surgical_neck.py:31-34, bicipital_groove.py:161, anatomic_neck.py:35
    return (full.centroids((0.35, 0.75)).sum() + full.areas1((0.70, 0.99)).sum() + prox.itr_centered_start((0.2, 0.75)).sum()
            + prox.itr_start((0.0, 0.852)).sum() + dist.centroids((0.8, 0.99)).sum())


def main():
    _lib.init(0)
    obb = meshio.PcaObb(ROOT / "tests" / "golden" / "bones" / "humerus_left.npz")
    for it in range(6):
        t0 = time.perf_counter()
        full, dist, prox = GpuFullSlices(obb), GpuDistalSlices(obb), GpuProximalSlices(obb, Neck())
        run_batch([full, dist, prox])
        v = touch(full, dist, prox)
        t1 = time.perf_counter()
        print(f"gpu e2e one bone, 1000 planes, 3 sweeps: {1e3 * (t1 - t0):.2f} ms (checksum {v:.6f})")
    t0 = time.perf_counter()
    m = obb.mesh
    o = [oracle.OracleSlices(m.vertices, m.faces, s._zs, s._interp_num) for s in (full, dist, prox)]
    c = (o[0].window(o[0].centroids, (0.35, 0.75)).sum() + o[0].window(o[0].areas1, (0.70, 0.99)).sum()
         + o[2].window(o[2].itr_centered_start, (0.2, 0.75)).sum() + o[2].window(o[2].itr_start, (0.0, 0.852)).sum()
         + o[1].window(o[1].centroids, (0.8, 0.99)).sum())
    t1 = time.perf_counter()
    print(f"cpu restatement (1 core): {1e3 * (t1 - t0):.1f} ms (checksum {c:.6f})")


if __name__ == "__main__":
    main()
