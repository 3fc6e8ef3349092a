"""Scope row f2 measured: bones/s from STL BYTES to a framed resident mesh (shb_mesh_from_stl) and on to the three default
sweeps of bone.Humerus, beside the host path (numpy parse + weld + PCA frame, then upload).  Run on the GPU box:
    python tools/stl_probe.py [n_bones]"""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from shoulder_b200 import _lib, meshio
from shoulder_b200.mesh import GpuMesh
from oracle import meshload        # only to ENCODE the test input (STL bytes of the synthetic bones); not timed, not compared

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
_lib.init(0)
names = ["humerus_left", "humerus_right", "humerus_left_trab", "humerus_left_flipped"]
bases = [meshio.load_mesh(ROOT / "tests" / "golden" / "bones" / f"{k}.npz") for k in names]
raws = []
for i in range(n):
    m = meshio.synthetic_bone(bases[i % 4], i)
    raws.append(meshload.encode_stl(m.vertices, m.faces))
mb = sum(len(r) for r in raws) / 1e6


def sweeps_of(z):
    full = np.linspace(0.99 * z.max(), 0.99 * z.min(), 200)
    dist = np.linspace(0.99 * z.min(), 0.0, 200)
    prox = np.linspace(0.99 * z.max(), 0.55 * z.max(), 600)
    return [(full, 100), (dist, 500), (prox, 512)]


def gpu_load_only():
    for r in raws:
        GpuMesh.from_stl(r, frame=True)


def host_load_only():
    out = []
    for r in raws:
        n_t = np.frombuffer(r, dtype="<u4", count=1, offset=80)[0]
        rec = np.frombuffer(r, dtype=meshload.STL_RECORD, count=n_t, offset=84)          # what meshio.read_stl does on a file
        v, f = meshio.weld(np.array(rec["vertices"], dtype=np.float32))
        out.append(meshio.PcaObb(meshio.Mesh(v, f)))
    return out


def timeit(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


t_gpu, t_host = timeit(gpu_load_only), timeit(host_load_only, 1)


def gpu_full():
    for r in raws:
        m, info = GpuMesh.from_stl(r, frame=True)
        zb = np.array(sorted(info["z_bounds"])) * (-1 if info["flipped"] else 1)
        z = np.array([zb.min(), zb.max()])
        for zs, N in sweeps_of(z):
            m.section_multiplane([0, 0, float(zs.mean())], [0, 0, 1], zs - zs.mean())


t_full = timeit(gpu_full, 2)
print(json.dumps({"bones": n, "stl_MB": round(mb, 2),
                  "gpu_stl_to_framed_resident_mesh": {"bones_per_s": round(n / t_gpu, 1), "ms_per_bone": round(1e3 * t_gpu / n, 3),
                                                      "includes": "H2D of the file bytes, parse, weld, PCA frame, end test, adjacency, D2H of the welded arrays for the host attributes"},
                  "host_numpy_parse_weld_frame": {"bones_per_s": round(n / t_host, 1), "ms_per_bone": round(1e3 * t_host / n, 3)},
                  "gpu_stl_to_three_default_sweeps_contours": {"bones_per_s": round(n / t_full, 1), "ms_per_bone": round(1e3 * t_full / n, 3),
                                                               "note": "one bone at a time through GpuMesh.section_multiplane (plane records + contours), not the batched call"}}))
