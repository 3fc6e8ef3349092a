"""Scope row f2: STL bytes -> welded mesh -> frame.  CPU: the oracle (oracle/meshload.py) against the reference's own STL
fixtures where they are readable (this container) and against the committed re-encodings; GPU: shb_mesh_from_stl /
shb_mesh_read through the C ABI against the oracle on the same bytes."""
from pathlib import Path

import numpy as np
import pytest

from oracle import meshload
from shoulder_b200 import meshio

BONES = Path(__file__).parent / "golden" / "bones"
NAMES = ["humerus_left", "humerus_right", "humerus_left_trab", "humerus_left_flipped"]
REF_STL = Path("/root/reference/tests/test_bones")


def _stl_bytes(name):
    m = meshio.load_mesh(BONES / f"{name}.npz")
    return meshload.encode_stl(m.vertices, m.faces), m


@pytest.mark.parametrize("name", NAMES)
def test_oracle_loader_round_trips_the_committed_bones(name):
    raw, m = _stl_bytes(name)
    v, f = meshload.load_mesh(raw)
    assert np.array_equal(f, m.faces) and np.array_equal(v, m.vertices.astype(np.float32).astype(np.float64))


@pytest.mark.skipif(not REF_STL.exists(), reason="the reference's STL files exist only in the build container")
@pytest.mark.parametrize("name", NAMES)
def test_oracle_loader_on_the_reference_stl_files_equals_the_committed_bones(name):
    v, f = meshload.load_mesh((REF_STL / f"{name}.stl").read_bytes())
    m = meshio.load_mesh(BONES / f"{name}.npz")
    assert np.array_equal(v, m.vertices) and np.array_equal(f, m.faces)


def test_merge_is_by_rounded_cell_and_keeps_first_occurrence():
    # two corners 4e-9 apart share a 1e-8 cell (merged, first coordinates kept); -0.0 and +0.0 are one vertex
    tri = np.array([[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]],
                    [[-0.0, 0.0, 0.0], [1.0, 0.0, 4e-9], [0.0, 0.0, 1.0]]], dtype=np.float32)
    raw = meshload.encode_stl(tri.reshape(-1, 3), np.arange(6).reshape(2, 3))
    v, f = meshload.load_mesh(raw)
    assert f.tolist() == [[0, 1, 2], [0, 1, 3]] and len(v) == 4 and v[1, 2] == 0.0


def test_frame_of_the_oracle_equals_the_host_stand_in():
    raw, m = _stl_bytes("humerus_right")
    v, f = meshload.load_mesh(raw)
    T, out, zb, zl, _ = meshload.pca_frame(v)
    o = meshio.PcaObb(meshio.Mesh(v, f))
    assert np.array_equal(T, o.transform) and np.array_equal(out, o.mesh.vertices) and zb == tuple(o.z_bounds)
    assert abs(out[:, :3].min(axis=0) + out[:, :3].max(axis=0)).max() < 1e-9        # AABB centred on the origin


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_weld_is_bit_exact_and_frame_matches(gpu_backend, name):
    from shoulder_b200.mesh import GpuMesh, GpuObb
    raw, m = _stl_bytes(name)
    v, f = meshload.load_mesh(raw)
    g = GpuMesh.from_stl(raw)
    assert np.array_equal(g.faces, f) and np.array_equal(g.vertices, v)              # numbering, order, coordinates: exact
    obb = GpuObb(raw, name=name)
    T, out, zb, zl, resid = meshload.pca_frame(v)
    assert np.array_equal(obb.mesh.faces, f)
    scale = np.abs(out).max()
    assert np.abs(obb.transform - T).max() < 1e-9 * max(1.0, np.abs(T).max())
    assert np.abs(obb.mesh.vertices - out).max() < 1e-9 * scale                      # float64 reductions in another order
    assert abs(obb.z_bounds[0] - zb[0]) < 1e-9 * scale and abs(obb.z_bounds[1] - zb[1]) < 1e-9 * scale
    assert abs(obb.z_length - zl) < 1e-9 * scale
    # the resident mesh slices like an uploaded copy of the same arrays
    zs = np.linspace(0.9 * zb[1], 0.9 * zb[0], 7)
    a = obb.mesh.section_multiplane([0, 0, 0.0], [0, 0, 1], zs)
    b = GpuMesh(obb.mesh.vertices, obb.mesh.faces).section_multiplane([0, 0, 0.0], [0, 0, 1], zs)
    for pa, pb in zip(a, b):
        assert len(pa.discrete) == len(pb.discrete) and all(np.array_equal(x, y) for x, y in zip(pa.discrete, pb.discrete))


@pytest.mark.gpu
def test_gpu_weld_edge_cases(gpu_backend):
    from shoulder_b200 import _lib
    from shoulder_b200.mesh import GpuMesh
    tri = np.array([[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]],
                    [[-0.0, 0.0, 0.0], [1.0, 0.0, 4e-9], [0.0, 0.0, 1.0]]], dtype=np.float32)
    raw = meshload.encode_stl(tri.reshape(-1, 3), np.arange(6).reshape(2, 3))
    g = GpuMesh.from_stl(raw)
    v, f = meshload.load_mesh(raw)
    assert np.array_equal(g.faces, f) and np.array_equal(g.vertices, v)
    # jittered, rotated copies (another vertex order of magnitude, another frame): exact weld, frame to rounding
    for bid in (3, 8):
        m = meshio.synthetic_bone(meshio.load_mesh(BONES / "humerus_left_trab.npz"), bid)
        raw = meshload.encode_stl(m.vertices, m.faces)
        v, f = meshload.load_mesh(raw)
        g, info = GpuMesh.from_stl(raw, frame=True)
        T, out, zb, zl, _ = meshload.pca_frame(v)
        assert np.array_equal(g.faces, f)
        assert np.abs(info["transform"] - T).max() < 1e-9 * np.abs(T).max() and np.abs(g.vertices - out).max() < 1e-9 * np.abs(out).max()
    with pytest.raises(_lib.BackendError):
        GpuMesh.from_stl(b"solid ascii\nfacet normal 0 0 0\n" + b" " * 100)
    with pytest.raises(_lib.BackendError):
        GpuMesh.from_stl(raw[:-7])                                                  # truncated file
