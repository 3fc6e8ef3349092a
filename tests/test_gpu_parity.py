"""Parity of the CUDA path with the oracle, through the C ABI (needs a B200)."""
import numpy as np
import pytest

from shoulder_b200 import meshio

from helpers import compare_sweep

pytestmark = pytest.mark.gpu


def _zs(v, n, inset=0.99):
    return np.linspace(inset * v[:, 2].max(), inset * v[:, 2].min(), n)


def test_ellipsoid(gpu_backend):
    v, f = meshio.icosphere(3, 1.0, scale=(20.0, 30.0, 170.0))
    rep = compare_sweep(v, f, _zs(v, 64), 100, n_angles=90)
    assert rep["contours"] == 64 and not rep["h4_exceptions"]


def test_torus_two_loops(gpu_backend):
    v, f = meshio.torus(30.0, 8.0, 64, 32)
    rep = compare_sweep(v, f, _zs(v, 40, 0.97), 128, n_angles=90)      # centroid outside the outline: all-candidates rays
    assert rep["contours"] > 40            # planes through the hole cut two loops


@pytest.mark.parametrize("name", ["humerus_left", "humerus_right", "humerus_left_trab", "humerus_left_flipped"])
def test_bones_full_sweep(gpu_backend, bone_obbs, name):
    m = bone_obbs(name).mesh
    rep = compare_sweep(m.vertices, m.faces, _zs(m.vertices, 200), 100)
    assert rep["segments"] > 20000
    assert not rep["h4_exceptions"]


def test_bone_proximal_sweep_512(gpu_backend, bone_obbs):
    m = bone_obbs("humerus_left").mesh
    z = m.vertices[:, 2]
    zs = np.linspace(0.99 * z.max(), 0.55 * z.max(), 600)
    rep = compare_sweep(m.vertices, m.faces, zs, 512, n_angles=360)
    assert rep["max_rel"] < 1e-9


def test_ascending_and_unsorted_heights(gpu_backend, bone_obbs):
    m = bone_obbs("humerus_right").mesh
    z = m.vertices[:, 2]
    zs = np.linspace(0.99 * z.min(), 0.0, 50)                      # ascending, like DistalSlices
    compare_sweep(m.vertices, m.faces, zs, 500)
    rng = np.random.default_rng(7)
    compare_sweep(m.vertices, m.faces, rng.permutation(zs), 64)     # arbitrary order


def test_fast_mode_matches_full_mode(gpu_backend, bone_obbs):
    """Without SHB_OUT_SEGMENTS the stitcher skips the canonical sort and evaluates one copy per node;
    contours, plane records and profiles must not change by a single bit."""
    from shoulder_b200 import _lib
    from helpers import run_gpu
    m = bone_obbs("humerus_left_trab").mesh
    zs = np.linspace(0.99 * m.vertices[:, 2].max(), 0.99 * m.vertices[:, 2].min(), 300)
    fast = run_gpu(m.vertices, m.faces, zs, 128, _lib.OUT_PLANE | _lib.OUT_CONTOURS | _lib.OUT_ALL_PROFILES | _lib.OUT_RADIAL, 72)
    full = run_gpu(m.vertices, m.faces, zs, 128, _lib.OUT_PLANE | _lib.OUT_SEGMENTS | _lib.OUT_CONTOURS | _lib.OUT_ALL_PROFILES | _lib.OUT_RADIAL, 72)
    for w in (_lib.ARR_N_SEG, _lib.ARR_N_ENT, _lib.ARR_BOUNDS, _lib.ARR_CENTROID, _lib.ARR_SEL, _lib.ARR_CONTOUR_OFF,
              _lib.ARR_CONTOUR_PT_OFF, _lib.ARR_POINTS, _lib.ARR_IXY, _lib.ARR_ITR_START, _lib.ARR_ITR_CENTERED_START,
              _lib.ARR_ITR, _lib.ARR_ITR_CENTERED, _lib.ARR_RADIAL):
        assert np.array_equal(fast.array(w), full.array(w)), w
    assert np.allclose(fast.array(_lib.ARR_AREA1), full.array(_lib.ARR_AREA1), rtol=1e-13)
    with pytest.raises(_lib.BackendError):
        fast.array(_lib.ARR_SEGMENTS)


def test_radial_ray_parallel_path_equals_the_all_candidates_path(gpu_backend, bone_obbs):
    """Star-shaped outlines take a ray-parallel path (one owner edge per ray); it must deliver the bits of the
    edge-parallel all-candidates path that every other outline takes."""
    import os
    from shoulder_b200 import _lib
    from helpers import run_gpu
    m = bone_obbs("humerus_left").mesh
    zs = np.linspace(0.99 * m.vertices[:, 2].max(), 0.99 * m.vertices[:, 2].min(), 400)
    mask = _lib.OUT_PLANE | _lib.OUT_IXY | _lib.OUT_RADIAL
    fast = run_gpu(m.vertices, m.faces, zs, 64, mask, 360).array(_lib.ARR_RADIAL)
    os.environ["SHB_DEBUG_RADIAL_GENERAL"] = "1"
    try:
        slow = run_gpu(m.vertices, m.faces, zs, 64, mask, 360).array(_lib.ARR_RADIAL)
        slow7 = run_gpu(m.vertices, m.faces, zs, 64, mask, 7).array(_lib.ARR_RADIAL)
    finally:
        del os.environ["SHB_DEBUG_RADIAL_GENERAL"]
    assert np.array_equal(fast, slow) and np.isfinite(fast).all() and (fast > 0).mean() > 0.9     # 0 = ray misses the outline
    assert np.array_equal(run_gpu(m.vertices, m.faces, zs, 64, mask, 7).array(_lib.ARR_RADIAL), slow7)     # few, wide rays


def test_dense_sweep_with_several_contour_planes(gpu_backend, bone_obbs):
    """1,531 planes over the trabecular bone: the ends of the bone and its cavities give planes with two and three
    contours, whose order and start nodes follow CPython's set in the reference."""
    m = bone_obbs("humerus_left_trab").mesh
    z = m.vertices[:, 2]
    zs = np.linspace(0.995 * z.max(), 0.995 * z.min(), 1531)
    rep = compare_sweep(m.vertices, m.faces, zs, 128, n_angles=90, expect_all_closed=False)
    assert rep["contours"] > len(zs) + 20 and rep["pts_bitexact"] and not rep["h4_exceptions"]
    assert rep["max_rel"] < 1e-12


def test_polar_forms_against_numpy_bits(gpu_backend, bone_obbs):
    """The unroll's polar pass (``shb_polar``: one reciprocal square root per sample feeds both r and theta) against numpy
    on the samples the device itself delivered: r is ``np.sqrt(x**2 + y**2)`` (slice.py:100-101,138-139) BIT FOR BIT,
    theta is within 4 ulp of ``np.arctan2`` (slice.py:99,137), and the roll puts the smallest theta first."""
    from shoulder_b200 import _lib
    from helpers import run_gpu
    m = bone_obbs("humerus_right").mesh
    zs = np.linspace(0.99 * m.vertices[:, 2].max(), 0.99 * m.vertices[:, 2].min(), 700)
    res = run_gpu(m.vertices, m.faces, zs, 360, _lib.OUT_PLANE | _lib.OUT_ALL_PROFILES)
    for xy_name, tr_name in ((_lib.ARR_IXY, _lib.ARR_ITR_START), (_lib.ARR_IXY_CENTERED, _lib.ARR_ITR_CENTERED_START)):
        xy, tr = res.array(xy_name, 0), res.array(tr_name, 0)
        x, y = xy[:, 0, :], xy[:, 1, :]
        theta, r = np.arctan2(y, x), np.sqrt(x ** 2 + y ** 2)
        km = np.argmin(theta, axis=1)
        idx = (np.arange(x.shape[1])[None, :] + km[:, None]) % x.shape[1]
        theta, r = np.take_along_axis(theta, idx, 1), np.take_along_axis(r, idx, 1)
        assert np.array_equal(tr[:, 1, :], r)
        assert np.abs(tr[:, 0, :] - theta).max() <= 4 * np.spacing(np.pi)
        assert (tr[:, 0, 0] <= tr[:, 0, :].min(axis=1)).all()
