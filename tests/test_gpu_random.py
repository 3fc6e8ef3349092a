"""Seeded random solids against the oracle, through the C ABI (needs a B200).

Lumpy closed surfaces (an icosphere whose radius is modulated by a few random spherical waves) under a random
affine map: non-convex, so planes cut several loops, outlines that are not star-shaped about their centroid, and —
because some of the heights are vertex heights exactly — planes through vertices.  Every array is compared the way
tests/helpers.py compares the bones: triangle sets / order / connectivity exactly, coordinates at 1e-11, profiles
and the radius image at 1e-9 (north-star budget 1e-5)."""
import numpy as np
import pytest

from shoulder_b200 import meshio

from helpers import compare_sweep

pytestmark = pytest.mark.gpu


def lumpy(seed: int, levels: int):
    rng = np.random.default_rng(seed)
    v, f = meshio.icosphere(levels, 1.0)
    r = np.ones(len(v))
    for _ in range(4):
        k = rng.normal(size=3)
        k *= rng.uniform(2.0, 6.0) / np.linalg.norm(k)
        r += rng.uniform(0.20, 0.45) * np.sin(v @ k + rng.uniform(0, 2 * np.pi))
    v = v * np.maximum(r, 0.25)[:, None]
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    a = q * rng.uniform(8.0, 30.0, size=3)                     # rotation times anisotropic scale (mm-sized coordinates)
    v = v @ a.T + rng.uniform(-40.0, 40.0, size=3)
    if np.linalg.det(a) < 0:
        f = f[:, ::-1]                                         # keep the winding outward
    return np.ascontiguousarray(v), np.ascontiguousarray(f)


@pytest.mark.parametrize("seed", range(10))
def test_lumpy_solid(gpu_backend, seed):
    rng = np.random.default_rng(1000 + seed)
    v, f = lumpy(seed, 3 + seed % 2)
    z = v[:, 2]
    zs = np.linspace(z.max() - 1e-3 * np.ptp(z), z.min() + 1e-3 * np.ptp(z), int(rng.integers(24, 48)))
    if seed % 3 == 0:                                          # planes through vertices (exact vertex heights), one of them twice
        extra = rng.choice(z[(z > zs.min()) & (z < zs.max())], size=4, replace=False)
        zs = np.concatenate([zs, extra, extra[:1]])
    if seed % 2:
        zs = rng.permutation(zs)                               # arbitrary plane order
    rep = compare_sweep(v, f, zs, int(rng.choice([37, 64, 100, 360])), n_angles=int(rng.choice([7, 72, 360])),
                        expect_all_closed=False)
    assert rep["segments"] > 1000 and rep["contours"] >= len(zs) - 8


@pytest.mark.parametrize("seed", range(4))
def test_scene_of_blobs(gpu_backend, seed):
    """Several separate lumpy solids in one mesh: many contours per plane, so the order in which the reference starts
    them (CPython's set, rebuilt as components are removed) is exercised well beyond two or three contours."""
    rng = np.random.default_rng(500 + seed)
    vs, fs, off = [], [], 0
    for k in range(int(rng.integers(6, 13))):
        v, f = lumpy(100 * seed + k, 2 + k % 2)
        v = 0.35 * (v - v.mean(axis=0)) + np.array([rng.uniform(-60, 60), rng.uniform(-60, 60), rng.uniform(-3, 3)])
        vs.append(v); fs.append(f + off); off += len(v)
    v, f = np.vstack(vs), np.vstack(fs)
    z = v[:, 2]
    zs = np.linspace(np.percentile(z, 80), np.percentile(z, 20), 30)
    rep = compare_sweep(v, f, zs, 48, n_angles=36, expect_all_closed=False)
    assert rep["contours"] >= 4 * len(zs)


def test_bucketed_and_all_pairs_node_ranks_agree(gpu_backend, monkeypatch):
    """The node ids behind the contour order come from a bucketed ranking on large planes and from an all-pairs count on
    small ones (or when the shared scratch is short): one scene, both ways, identical results — and the oracle again."""
    from shoulder_b200 import _lib
    rng = np.random.default_rng(77)
    vs, fs, off = [], [], 0
    for k in range(9):
        v, f = lumpy(900 + k, 4)                                # 5,120 faces each: several hundred segments per plane
        v = 0.35 * (v - v.mean(axis=0)) + np.array([rng.uniform(-60, 60), rng.uniform(-60, 60), rng.uniform(-2, 2)])
        vs.append(v); fs.append(f + off); off += len(v)
    v, f = np.vstack(vs), np.vstack(fs)
    z = v[:, 2]
    zs = np.linspace(np.percentile(z, 70), np.percentile(z, 30), 12)
    mask = _lib.OUT_PLANE | _lib.OUT_CONTOURS
    a = _lib.sweep_batch([(v, f)], [(0, float(zs.mean()), zs - zs.mean(), 16)], mask)
    assert a.array(_lib.ARR_N_SEG).min() >= 256                 # large enough for the bucketed path
    monkeypatch.setenv("SHB_DEBUG_ALLPAIRS_RANK", "1")
    b = _lib.sweep_batch([(v, f)], [(0, float(zs.mean()), zs - zs.mean(), 16)], mask)
    monkeypatch.delenv("SHB_DEBUG_ALLPAIRS_RANK")
    for w in (_lib.ARR_CONTOUR_OFF, _lib.ARR_CONTOUR_PT_OFF, _lib.ARR_POINTS, _lib.ARR_CONTOUR_AREA, _lib.ARR_STATUS):
        assert np.array_equal(a.array(w), b.array(w))
    rep = compare_sweep(v, f, zs, 16, expect_all_closed=False)
    assert rep["contours"] >= 4 * len(zs)
