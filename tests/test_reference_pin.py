"""Pins of the post-trimesh half of the path to the REFERENCE'S OWN CODE.

The reference's ``slice.py`` (and ``canal.py``) are imported unchanged from /root/reference (tests/refload.py) and run
over ``obb.mesh.section_multiplane`` answers supplied by the oracle's restatement of trimesh.  Three layers:

  1. live (build container only, /root/reference present): the reference's classes at their DEFAULT sizes on the four
     test bones == ``oracle/slice_arrays.py``, ``np.array_equal`` on all eight cached arrays, quirks included;
  2. the committed vectors ``tests/golden/refvec_*.npz`` (made by tests/golden/make_reference_vectors.py from the
     reference's classes) still equal a live run, and equal the oracle on any host;
  3. ``-m gpu``: the CUDA path behind ``GpuFullSlices / GpuDistalSlices / GpuProximalSlices`` reproduces those
     reference-made vectors (the GPU box has no /root/reference).

What stays unpinned is the trimesh half (segments -> Path2D), see oracle/__init__.py.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle
from oracle.landmarks import axis_angle_deg, canal_axis_obb
from shoulder_b200 import meshio

import refload

HERE = Path(__file__).resolve().parent
GOLD = HERE / "golden"
sys.path.insert(0, str(GOLD))
import make_reference_vectors as mrv  # noqa: E402

NAMES = mrv.NAMES
PAIRS = (("_centroids", "centroids"), ("_areas1", "areas1"), ("_ixy", "ixy"), ("_ixy_centered", "ixy_centered"), ("_itr", "itr"),
         ("_itr_start", "itr_start"), ("_itr_start_even_theta", "itr_start"), ("_itr_centered", "itr_centered"),
         ("_itr_centered_start", "itr_centered_start"))
needs_reference = pytest.mark.skipif(not refload.available(), reason="/root/reference is not on this host")


def _bone(name):
    ct = meshio.load_mesh(GOLD / "bones" / f"{name}.npz")
    g = np.load(GOLD / f"refvec_{name}.npz")
    v, f = refload.exact_frame(ct.vertices, ct.faces, g["transform"])
    return g, v, f


@needs_reference
@pytest.mark.parametrize("name", NAMES)
def test_reference_slice_py_over_oracle_paths_equals_the_oracle_arrays(name, bone_obbs):
    """Layer 1, default sizes (bone.py:116-121): FullSlices 200x100, DistalSlices 200x500, ProximalSlices 600x512."""
    S = refload.reference_modules()["slice"]
    m = bone_obbs(name).mesh
    made = {}

    def factory(mesh, origin, normal, heights):          # answer the reference's one trimesh call from the oracle
        o = made["orc"]
        assert origin == [0, 0, o.z_orig] and normal == [0, 0, 1] and np.array_equal(heights, o.z_incrs)
        return o.paths

    obb = refload.OracleObb(m.vertices, m.faces, factory=factory)
    zmax = obb.mesh.bounds[1, 2]
    for ref in (S.FullSlices(obb), S.DistalSlices(obb), S.ProximalSlices(obb, refload.Neck(mrv.NECK_FRAC * zmax))):
        made["orc"] = orc = oracle.OracleSlices(m.vertices, m.faces, ref._zs, ref._interp_num)
        assert ref._z_orig == orc.z_orig and np.array_equal(ref._z_incrs, orc.z_incrs)                # slice.py:18-19
        for a, b in PAIRS:
            assert np.array_equal(getattr(ref, a), getattr(orc, b)), (type(ref).__name__, a)
        for cutoff in ((0.35, 0.75), (0.8, 0.99), (0.0, 0.852), (0.2, 0.75), (0.70, 0.99)):           # the consumers' windows
            lo, hi = oracle.cutoff_window(len(ref._zs), cutoff)
            assert np.array_equal(ref.zs(cutoff), ref._zs[lo:hi])
            assert np.array_equal(ref.itr(cutoff), orc.ixy[lo:hi])                                    # sic, slice.py:99-100
            assert np.array_equal(ref.itr_start_even_theta(cutoff), orc.itr_start[lo:hi])             # sic, slice.py:121-122


@needs_reference
def test_reference_cutoff_equals_the_mirror_on_random_windows(bone_obbs):
    """slice.py:157-164 against GpuSlices._cutoff (pure host code) incl. return_odd."""
    from shoulder_b200.slice import GpuFullSlices
    S = refload.reference_modules()["slice"]
    obb = bone_obbs("humerus_right")
    rng = np.random.default_rng(3)
    for odd in (False, True):
        ref = S.FullSlices(refload.OracleObb(obb.mesh.vertices, obb.mesh.faces), return_odd=odd)
        got = GpuFullSlices(obb, return_odd=odd)
        for _ in range(200):
            n = int(rng.integers(1, 700))
            a, b = np.sort(rng.uniform(0, 1, 2))
            ent = np.arange(n)
            assert np.array_equal(ref._cutoff(ent, (a, b)), got._cutoff(ent, (a, b)))


@needs_reference
def test_install_rebinds_the_reference_module():
    from shoulder_b200 import slice as gs
    S = refload.reference_modules()["slice"]
    saved = (S.FullSlices, S.ProximalSlices, S.DistalSlices)
    try:
        gs.install(S)
        assert (S.FullSlices, S.ProximalSlices, S.DistalSlices) == (gs.GpuFullSlices, gs.GpuProximalSlices, gs.GpuDistalSlices)
        import inspect
        for ours, ref in zip((gs.GpuFullSlices, gs.GpuProximalSlices, gs.GpuDistalSlices), saved):       # same constructor arguments
            assert list(inspect.signature(ours.__init__).parameters) == list(inspect.signature(ref.__init__).parameters)
        on_device = {"_resample_polygon", "_cart2pol", "_cart2pol_no_sort"}      # the CPU helpers K4 replaces
        names = [n for n in vars(saved[0].__mro__[1]) if not n.startswith("__") and not n.startswith("_abc") and n not in on_device]
        for n in names:                                   # every method / cached property of slice.Slices has a mirror
            assert hasattr(gs.GpuFullSlices, n), n
    finally:
        S.FullSlices, S.ProximalSlices, S.DistalSlices = saved


@needs_reference
@pytest.mark.parametrize("name", NAMES)
def test_committed_vectors_equal_a_live_run_of_the_reference(name):
    g, v, f = _bone(name)
    live = mrv.vectors(v, f)
    for key, val in live.items():
        assert np.array_equal(g[key], val), key


@pytest.mark.parametrize("name", NAMES)
def test_committed_vectors_equal_the_oracle(name):
    """Runs anywhere: the vectors are what oracle/slice_arrays.py produces on the same input."""
    g, v, f = _bone(name)
    for key, (P, N) in mrv.SIZES.items():
        zs = g[f"{key}__zs"]
        assert len(zs) == P
        orc = oracle.OracleSlices(v, f, zs, N)
        assert np.array_equal(g[f"{key}__n_entities"], orc.n_entities)
        assert np.array_equal(g[f"{key}__centroids"], orc.centroids)
        for a, b in PAIRS[1:]:
            assert np.allclose(g[f"{key}_{a}"], getattr(orc, b), rtol=1e-13, atol=1e-13), (key, a)
    zl = abs(v[:, 2].min()) + abs(v[:, 2].max())
    ax = canal_axis_obb(g["full__centroids"], g["full__zs"], zl)
    assert axis_angle_deg(ax, g["canal_axis"]) < 1e-4 and np.allclose(ax, g["canal_axis"], rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_slices_reproduce_the_reference_made_vectors(gpu_backend, name):
    """Layer 3: rows a1, a2, a7-a13 of the scope table against vectors computed by the reference's own slice.py."""
    from shoulder_b200.slice import GpuDistalSlices, GpuFullSlices, GpuProximalSlices, run_batch
    g, v, f = _bone(name)

    class Obb:
        mesh = meshio.Mesh(v, f)
    zmax = Obb.mesh.bounds[1, 2]
    objs = {"full": GpuFullSlices(Obb(), *mrv.SIZES["full"]), "distal": GpuDistalSlices(Obb(), *mrv.SIZES["distal"]),
            "proximal": GpuProximalSlices(Obb(), refload.Neck(mrv.NECK_FRAC * zmax), *mrv.SIZES["proximal"])}
    run_batch(list(objs.values()))
    for key, s in objs.items():
        assert np.array_equal(s._zs, g[f"{key}__zs"]) and s._z_orig == g[f"{key}__z_orig"]
        assert np.array_equal(s._z_incrs, g[f"{key}__z_incrs"])
        assert np.array_equal([len(p.entities) for p in s._slices], g[f"{key}__n_entities"])
        assert np.array_equal(s._centroids, g[f"{key}__centroids"])                       # bit exact
        assert np.allclose(s._areas1, g[f"{key}__areas1"], rtol=1e-12, atol=0)
        for a, _ in PAIRS[2:]:
            ref = g[f"{key}_{a}"]
            got = np.asarray(getattr(s, a))
            assert got.shape == ref.shape, (key, a)
            err = np.abs(got - ref).max() / np.abs(ref).max()
            assert err < 1e-9, (key, a, err)                                              # north_star budget: 1e-5
        assert np.array_equal(s.centroids((0.35, 0.75)), g[f"{key}__win_035_075"])
        assert np.allclose(s.itr((0.2, 0.8)), g[f"{key}__win_itr_bug"], rtol=1e-9, atol=1e-9)
    zl = abs(v[:, 2].min()) + abs(v[:, 2].max())
    ax = canal_axis_obb(objs["full"]._centroids, objs["full"]._zs, zl)
    assert axis_angle_deg(ax, g["canal_axis"]) < 0.01                                     # north_star: axes within 0.01 deg
    assert np.allclose(ax, g["canal_axis"], rtol=1e-9, atol=1e-9)
