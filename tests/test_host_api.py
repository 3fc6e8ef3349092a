"""Host-side logic that needs no GPU: the C ABI library loads and exports what the header
declares, the Slices mirror windows exactly like slice.py:157-164, sharding covers the work."""
import re
from pathlib import Path

import numpy as np
import pytest

from shoulder_b200 import _lib, sharding
from shoulder_b200.slice import GpuDistalSlices, GpuFullSlices, GpuProximalSlices, GpuSlices

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "shoulder_b200.h").read_text()
    declared = set(re.findall(r"SHB_API [\w\s\*]+?\b(shb_\w+)\(", header))
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.shb_abi_version() == _lib.ABI_VERSION


def test_header_constants_match_python_mirror():
    header = (ROOT / "include" / "shoulder_b200.h").read_text()
    for name, val in re.findall(r"#define (SHB_(?:OUT|ST)_\w+)\s+(0x[0-9A-Fa-f]+)u", header):
        assert getattr(_lib, name[4:]) == int(val, 16), name


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.shb_init(0) == -2                     # SHB_E_CUDA, loudly
    assert b"no CPU fallback" in lib.shb_last_error()
    with pytest.raises(_lib.BackendError):
        _lib.sweep_batch([(np.zeros((3, 3)), np.array([[0, 1, 2]]))], [(0, 0.0, np.zeros(1), 8)], _lib.OUT_PLANE)


class _Obb:
    def __init__(self, mesh):
        self.mesh = mesh


class _Neck:
    neck_z = 40.0


def test_slices_mirror_heights_and_windows(bone_obbs):
    obb = bone_obbs("humerus_left")
    zmax, zmin = obb.mesh.bounds[1, 2], obb.mesh.bounds[0, 2]
    full, dist, prox = GpuFullSlices(obb), GpuDistalSlices(obb), GpuProximalSlices(obb, _Neck())
    assert np.array_equal(full._zs, np.linspace(0.99 * zmax, 0.99 * zmin, 200))      # slice.py:219-224
    assert np.array_equal(dist._zs, np.linspace(0.99 * zmin, 0, 200))                # slice.py:271-276
    assert np.array_equal(prox._zs, np.linspace(0.99 * zmax, 40.0, 600))             # slice.py:248-253
    assert full._z_orig == np.mean(full._zs) and np.array_equal(full._z_incrs, full._zs - full._z_orig)
    # slice.py:157-164 — int() truncation: (0.8, 0.99)@200 -> rows 2..38 (37 rows); (0, 0.852)@600 -> 88..599 (512 rows)
    assert np.array_equal(dist.zs((0.8, 0.99)), dist._zs[2:39]) and len(dist.zs((0.8, 0.99))) == 37
    assert len(prox.zs((0.0, 0.852))) == 512 and prox.zs((0.0, 0.852))[0] == prox._zs[88]
    odd = GpuFullSlices(obb, return_odd=True)
    assert len(odd.zs((0.35, 0.75))) % 2 == 1
    for name in ("slices", "centroids", "areas1", "ixy", "ixy_centered", "itr", "itr_start", "itr_start_even_theta",
                 "itr_centered", "itr_centered_start", "zs", "_cutoff"):
        assert callable(getattr(GpuSlices, name))


def test_bone_sharding_is_a_partition():
    for n, w in ((1024, 8), (10, 4), (3, 8)):
        got = np.concatenate([sharding.shard_bones(n, r, w) for r in range(w)])
        assert np.array_equal(np.sort(got), np.arange(n))
    cost = np.random.default_rng(0).uniform(1, 5, 64)
    parts = [sharding.shard_bones(64, r, 4, cost) for r in range(4)]
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(64))
    loads = [cost[p].sum() for p in parts]
    assert max(loads) / min(loads) < 1.15


def test_plane_sharding_tiles_the_sweep():
    for n, w in ((8192, 8), (600, 4), (5, 8)):
        r = [sharding.shard_planes(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def test_struct_layouts_match_the_header(tmp_path):
    """The structs that cross the C ABI by pointer (shb_landmark_args, shb_sweep_request) as the ctypes mirrors lay them
    out against what a C compiler makes of the header: sizes and every field offset."""
    import ctypes
    import shutil
    import subprocess
    from shoulder_b200 import features
    if shutil.which("gcc") is None:
        pytest.skip("gcc not found")
    fields = [n for n, _ in features._LandmarkArgs._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "shoulder_b200.h"\nint main(void) {\n'
                   '  printf("size %zu\\n", sizeof(shb_landmark_args));\n'
                   + "".join(f'  printf("{f} %zu\\n", offsetof(shb_landmark_args, {f}));\n' for f in fields)
                   + '  printf("request %zu\\n", sizeof(shb_sweep_request));\n  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", f"-I{ROOT / 'include'}", "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    assert int(out["size"]) == ctypes.sizeof(features._LandmarkArgs)
    for f in fields:
        assert int(out[f]) == getattr(features._LandmarkArgs, f).offset, f
    req = _lib.make_requests([{}])
    assert int(out["request"]) == ctypes.sizeof(req) // len(req)
