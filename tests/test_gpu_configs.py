"""Full-array oracle comparisons on the BASELINE configs AS BENCHMARKED (needs a B200).

  cfg2  bones 0..3 of bench.py's default batch: humerus_left under the config-4 jitter, 2,048 planes, N = 360, 360 rays,
        through the kernel instantiations the bench times (no SHB_OUT_SEGMENTS)
  cfg3  Loop-subdivided humerus_left, 8,192 planes: level 2 (519,040 triangles) on 256 sampled planes incl. ixy,
        itr_start, itr_centered_start and the radius image; level 3 (2,076,160 triangles) properties + 32 sampled planes
  cfg4  8 jittered bones x the three default sweeps at full 200 / 200 / 600 planes, with the consumers' windows
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle
from oracle.slice_arrays import radial_image, rows_for_paths
from shoulder_b200 import _lib, meshio

from helpers import rel_err

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
pytestmark = pytest.mark.gpu

BENCH_MASK = _lib.OUT_PLANE | _lib.OUT_IXY | _lib.OUT_ITR_START | _lib.OUT_ITR_CENTERED_START | _lib.OUT_RADIAL


def _check_rows(res, sweep, idx, paths, interp, angles, win_lo=0):
    """Plane records + profile rows of the planes ``idx`` against oracle paths of those planes."""
    A = lambda w: res.array(w, sweep)
    status = A(_lib.ARR_STATUS)
    keep = [k for k, p in enumerate(paths) if p is not None and p.info["agree"] and
            all(p.entity_closed(e) for e in range(len(p.entities)))]
    assert len(keep) >= 0.98 * len(idx)
    ii = np.asarray(idx)[keep]
    pp = [paths[k] for k in keep]
    assert not (status[ii] & (_lib.ST_EMPTY | _lib.ST_OPEN | _lib.ST_NONMANIFOLD)).any()
    assert np.array_equal(A(_lib.ARR_N_SEG)[ii], [len(p.metadata["face_index"]) for p in pp])
    assert np.array_equal(A(_lib.ARR_N_ENT)[ii], [len(p.entities) for p in pp])
    assert np.array_equal(A(_lib.ARR_CENTROID)[ii], np.array([p.centroid for p in pp]))          # bit exact
    assert np.array_equal(A(_lib.ARR_BOUNDS)[ii].reshape(-1, 2, 2), np.array([p.bounds for p in pp]))
    rows = rows_for_paths(pp, interp)
    assert rel_err(A(_lib.ARR_AREA1)[ii], rows["areas1"]) < 1e-12
    merged = [k for k, p in zip(ii, pp) if p.info.get("n_merged")]
    assert all(status[k] & _lib.ST_MERGED for k in merged)
    worst = 0.0
    for name, which in (("ixy", _lib.ARR_IXY), ("itr_start", _lib.ARR_ITR_START), ("itr_centered_start", _lib.ARR_ITR_CENTERED_START)):
        got = A(which)
        e = rel_err(got[ii - win_lo], rows[name])
        worst = max(worst, e)
        assert e < 1e-9, (name, e)                                                                 # north_star budget: 1e-5
    if angles:
        e = rel_err(A(_lib.ARR_RADIAL)[ii - win_lo], radial_image(pp, angles))
        assert e < 1e-9, ("radial", e)
    return {"planes": len(ii), "merged": merged, "max_rel": worst}


def test_cfg2_as_benchmarked_full_arrays():
    import bench
    meshes, sweeps = bench.make_bones("cfg2", 4, 0, 2048, 360)
    res = _lib.sweep_batch(meshes, sweeps, BENCH_MASK | _lib.OUT_CONTOURS, 360)
    total_merged = 0
    for s, (k, zo, h, n) in enumerate(sweeps):
        v, f = meshes[k]
        zs = np.asarray(h) + zo
        assert n == 360 and len(zs) == 2048
        paths = oracle.section_multiplane(v, f, [0, 0, zo], [0, 0, 1], h, merge="topo")
        rep = _check_rows(res, s, np.arange(len(zs)), paths, 360, 360)
        assert rep["planes"] >= 2040
        total_merged += len(rep["merged"])
        # contours of every plane: count, order, start vertex, orientation, coordinates — bit exact
        ct_off, ctpt, pts = res.array(_lib.ARR_CONTOUR_OFF, s), res.array(_lib.ARR_CONTOUR_PT_OFF, s), res.array(_lib.ARR_POINTS, s)
        for i, p in enumerate(paths):
            if p is None or not p.info["agree"]:
                continue
            c0 = int(ct_off[i])
            for c, dsc in enumerate(p.discrete):
                assert np.array_equal(pts[int(ctpt[c0 + c]):int(ctpt[c0 + c + 1])], dsc), (s, i, c)
    print("cfg2: planes where Path.merge_vertices fused nodes:", total_merged)


def _subdivided(level):
    base = meshio.PcaObb(ROOT / "tests" / "golden" / "bones" / "humerus_left.npz").mesh
    v, f = meshio.loop_subdivide(base.vertices, base.faces, level)
    z = v[:, 2]
    return v, f, np.linspace(0.99 * z.max(), 0.99 * z.min(), 8192)


def test_cfg3_level2_sampled_planes_all_arrays():
    v, f, zs = _subdivided(2)
    assert len(f) == 519040
    zo = zs.mean()
    res = _lib.sweep_batch([(v, f)], [(0, float(zo), zs - zo, 360)], BENCH_MASK | _lib.OUT_CONTOURS, 360)
    idx = np.unique(np.r_[np.linspace(0, 8191, 248).astype(int), np.arange(0, 8), np.arange(8184, 8192)])      # incl. both ends
    assert len(idx) >= 256
    paths = oracle.section_multiplane(v, f, [0, 0, zo], [0, 0, 1], (zs - zo)[idx], merge="topo")
    rep = _check_rows(res, 0, idx, paths, 360, 360)
    ct_off, ctpt, pts = res.array(_lib.ARR_CONTOUR_OFF), res.array(_lib.ARR_CONTOUR_PT_OFF), res.array(_lib.ARR_POINTS)
    for i, p in zip(idx, paths):
        c0 = int(ct_off[i])
        for c, dsc in enumerate(p.discrete):
            assert np.array_equal(pts[int(ctpt[c0 + c]):int(ctpt[c0 + c + 1])], dsc), (i, c)
    print("cfg3 L2:", rep)


def test_cfg3_level3_two_million_triangles():
    v, f, zs = _subdivided(3)
    assert len(f) == 2076160
    zo = zs.mean()
    res = _lib.sweep_batch([(v, f)], [(0, float(zo), zs - zo, 360)], BENCH_MASK | _lib.OUT_CONTOURS, 360)
    status, n_ent, n_seg = res.array(_lib.ARR_STATUS), res.array(_lib.ARR_N_ENT), res.array(_lib.ARR_N_SEG)
    assert not (status & (_lib.ST_EMPTY | _lib.ST_OPEN | _lib.ST_NONMANIFOLD | _lib.ST_GENERAL)).any()
    assert n_seg.sum() > 9_000_000
    ctpt, pts = res.array(_lib.ARR_CONTOUR_PT_OFF), res.array(_lib.ARR_POINTS)
    assert np.array_equal(pts[ctpt[:-1]], pts[ctpt[1:] - 1])                                   # every contour closed
    merged = int(((status & _lib.ST_MERGED) != 0).sum())
    assert np.diff(ctpt).sum() <= n_seg.sum() + n_ent.sum() and np.diff(ctpt).sum() >= n_seg.sum() + n_ent.sum() - 8 * max(merged, 1)
    ixy, itr = res.array(_lib.ARR_IXY), res.array(_lib.ARR_ITR_START)
    assert np.array_equal(ixy[:, :, 0], ixy[:, :, -1])
    assert np.array_equal(itr[:, 0, 0], itr[:, 0, :].min(axis=1))
    idx = np.linspace(3, 8188, 32).astype(int)
    paths = oracle.section_multiplane(v, f, [0, 0, zo], [0, 0, 1], (zs - zo)[idx], merge="topo")
    rep = _check_rows(res, 0, idx, paths, 360, 360)
    print("cfg3 L3:", rep, "planes with merged vertices:", merged)


def test_cfg4_default_sweeps_with_consumer_windows():
    """8 jittered bones x (Full 200x100, Distal 200x500, Proximal 600x512) in ONE call, each sweep asking for what its
    consumers read (plane records; itr_start rows 88..599, itr_centered_start rows 150..479 of the proximal sweep)."""
    import bench
    meshes, sweeps = bench.make_bones("cfg4", 8, 0, 0, 0)
    requests = bench.consumer_requests(sweeps)
    res = _lib.sweep_batch(meshes, sweeps, _lib.OUT_PLANE, 0, requests=requests)
    delivered = 0
    for s, (k, zo, h, n) in enumerate(sweeps):
        v, f = meshes[k]
        P = len(h)
        paths = oracle.section_multiplane(v, f, [0, 0, zo], [0, 0, 1], h, merge="topo")
        ok = [i for i, p in enumerate(paths) if p is not None and p.info["agree"]]
        assert len(ok) == P
        assert np.array_equal(res.array(_lib.ARR_CENTROID, s), np.array([p.centroid for p in paths]))
        rows = rows_for_paths(paths, n)
        assert rel_err(res.array(_lib.ARR_AREA1, s), rows["areas1"]) < 1e-12
        delivered += 76 * P
        if P == 600:
            lo, hi = res.window(_lib.ARR_ITR_START, s)
            assert (lo, hi) == (88, 600)
            a = res.array(_lib.ARR_ITR_START, s)
            assert a.shape == (512, 2, 512) and rel_err(a, rows["itr_start"][88:600]) < 1e-9
            lo, hi = res.window(_lib.ARR_ITR_CENTERED_START, s)
            assert (lo, hi) == (150, 480)
            b = res.array(_lib.ARR_ITR_CENTERED_START, s)
            assert b.shape == (330, 2, 512) and rel_err(b, rows["itr_centered_start"][150:480]) < 1e-9
            delivered += a.nbytes + b.nbytes
            with pytest.raises(_lib.BackendError):
                res.array(_lib.ARR_IXY, s)                                   # not requested for this sweep
        else:
            with pytest.raises(_lib.BackendError):
                res.array(_lib.ARR_ITR_START, s)
    assert delivered / 8 <= 7.0e6                                            # bytes per bone (was ~20 MB with every row of every array)


def test_plane_range_shards_concatenate_to_the_unsharded_run():
    """BASELINE config 3's sharding (one large mesh, contiguous plane ranges, z_orig of the FULL list): the per-rank
    outputs of the CUDA path, concatenated, equal the single run bit for bit."""
    from shoulder_b200 import sharding
    base = meshio.PcaObb(ROOT / "tests" / "golden" / "bones" / "humerus_left_trab.npz").mesh
    v, f = meshio.loop_subdivide(base.vertices, base.faces, 1)
    z = v[:, 2]
    zs = np.linspace(0.99 * z.max(), 0.99 * z.min(), 1000)
    mask = BENCH_MASK | _lib.OUT_CONTOURS
    zo = float(np.mean(zs))
    one = _lib.sweep_batch([(v, f)], [(0, zo, zs - zo, 128)], mask, 90)
    for world in (2, 3, 8):
        parts = []
        for rank in range(world):
            zo_r, h_r, (lo, hi) = sharding.plane_shard_heights(zs, rank, world)
            assert zo_r == zo
            parts.append(_lib.sweep_batch([(v, f)], [(0, zo_r, h_r, 128)], mask, 90))
        arrays = (_lib.ARR_N_SEG, _lib.ARR_N_ENT, _lib.ARR_STATUS, _lib.ARR_CENTROID, _lib.ARR_BOUNDS, _lib.ARR_AREA1, _lib.ARR_IXY,
                  _lib.ARR_ITR_START, _lib.ARR_ITR_CENTERED_START, _lib.ARR_RADIAL)
        for w in arrays + (_lib.ARR_POINTS, _lib.ARR_CONTOUR_AREA):
            cat = np.concatenate([p.array(w) for p in parts])
            assert np.array_equal(cat, one.array(w), equal_nan=True), (world, w)
        # the block-cyclic deal bench.py uses: per-plane arrays scattered back by index
        cyc = []
        for rank in range(world):
            zo_r, h_r, idx = sharding.plane_shard_heights_cyclic(zs, rank, world, block=32)
            cyc.append((idx, _lib.sweep_batch([(v, f)], [(0, zo_r, h_r, 128)], mask, 90)))
        for w in arrays:
            ref = one.array(w)
            got = np.zeros_like(ref)
            for idx, p in cyc:
                got[idx] = p.array(w)
            assert np.array_equal(got, ref, equal_nan=True), (world, w, "cyclic")
