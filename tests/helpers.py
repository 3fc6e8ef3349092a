"""Comparison of the CUDA path (through the C ABI) with the oracle on identical inputs."""
from __future__ import annotations

import numpy as np

import oracle
from oracle.slice_arrays import radial_image
from shoulder_b200 import _lib

# north_star tolerance: coordinates / profiles within 1e-5 relative (fp32 budget).  The device
# computes in fp64 with numpy's operation order, so the tests hold it to a far tighter bound and
# to exact equality wherever the arithmetic is elementwise.
REL_TOL = 1e-5
TIGHT = 1e-11


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max() if b.size else 0.0, 1e-300)
    return float(np.abs(a - b).max() / scale) if a.size else 0.0


def run_gpu(vertices, faces, zs, interp_num, mask=_lib.OUT_PLANE | _lib.OUT_SEGMENTS | _lib.OUT_CONTOURS | _lib.OUT_ALL_PROFILES,
            n_angles=0, **kw):
    zs = np.asarray(zs, dtype=np.float64)
    z_orig = np.mean(zs)
    return _lib.sweep_batch([(vertices, faces)], [(0, float(z_orig), zs - z_orig, interp_num)], mask, n_angles)


FULL_MASK = _lib.OUT_PLANE | _lib.OUT_SEGMENTS | _lib.OUT_CONTOURS | _lib.OUT_ALL_PROFILES
FAST_MASK = FULL_MASK & ~_lib.OUT_SEGMENTS

# arrays that must not depend on whether mesh_multiplane's own outputs were requested (the FULL / FAST instantiations
# of the stitch kernels) — compared bit for bit
_MODE_INDEPENDENT = ("ARR_N_SEG", "ARR_SEG_OFF", "ARR_N_ENT", "ARR_STATUS", "ARR_BOUNDS", "ARR_CENTROID", "ARR_AREA1", "ARR_SEL",
                     "ARR_CONTOUR_OFF", "ARR_CONTOUR_PT_OFF", "ARR_CONTOUR_AREA", "ARR_POINTS", "ARR_IXY", "ARR_IXY_CENTERED",
                     "ARR_ITR", "ARR_ITR_START", "ARR_ITR_CENTERED", "ARR_ITR_CENTERED_START")


def compare_sweep(vertices, faces, zs, interp_num, res=None, sweep=0, n_angles=0, expect_all_closed=True, modes=("full", "fast")):
    """Asserts parity of every array with the oracle; returns a small report dict.

    Without ``res`` the sweep is run TWICE — with ``SHB_OUT_SEGMENTS`` (the FULL instantiation of the stitch
    kernels: canonical (class, face) sort, both endpoint copies, face_index / lines_2D delivered) and without it (the
    FAST instantiation and ``shb_stitch_warp``, which is what ``bench.py`` times) — and each run is held against the
    oracle on its own; the two must also agree bit for bit."""
    orc = oracle.OracleSlices(vertices, faces, zs, interp_num, merge="topo")
    if res is not None:
        return _compare(orc, res, sweep, zs, n_angles, expect_all_closed, has_segments=True)
    rep, runs = None, {}
    for mode in modes:
        mask = FULL_MASK if mode == "full" else FAST_MASK
        if n_angles:
            mask |= _lib.OUT_RADIAL
        runs[mode] = run_gpu(vertices, faces, zs, interp_num, mask, n_angles)
        r = _compare(orc, runs[mode], 0, zs, n_angles, expect_all_closed, has_segments=(mode == "full"))
        r["mode"] = mode
        rep = r if rep is None else rep
        if mode != rep["mode"]:
            assert r["contours"] == rep["contours"] and r["pts_bitexact"] == rep["pts_bitexact"]
            rep["max_rel"] = max(rep["max_rel"], r["max_rel"])
    if len(runs) == 2:
        for name in _MODE_INDEPENDENT + (("ARR_RADIAL",) if n_angles else ()):
            a, b = runs["full"].array(getattr(_lib, name)), runs["fast"].array(getattr(_lib, name))
            if name in ("ARR_AREA1", "ARR_CONTOUR_AREA"):       # fixed-order sums in both, but the two stitchers add in different orders
                assert np.allclose(a, b, rtol=1e-13, atol=0), f"{name}: FULL and FAST runs differ"
            else:
                assert np.array_equal(a, b, equal_nan=True), f"{name}: FULL and FAST runs differ"
    for r in runs.values():
        r.close()
    return rep


def _compare(orc, res, sweep, zs, n_angles, expect_all_closed, has_segments):
    A = lambda w: res.array(w, sweep)
    P = len(zs)
    n_seg, seg_off, n_ent, status = A(_lib.ARR_N_SEG), A(_lib.ARR_SEG_OFF), A(_lib.ARR_N_ENT), A(_lib.ARR_STATUS)
    if has_segments:
        fidx, segs = A(_lib.ARR_FACE_INDEX), A(_lib.ARR_SEGMENTS)
    ct_off, ctpt, ctarea, pts = A(_lib.ARR_CONTOUR_OFF), A(_lib.ARR_CONTOUR_PT_OFF), A(_lib.ARR_CONTOUR_AREA), A(_lib.ARR_POINTS)
    cent, bounds, area1 = A(_lib.ARR_CENTROID), A(_lib.ARR_BOUNDS), A(_lib.ARR_AREA1)
    rep = {"planes": P, "segments": 0, "contours": 0, "seg_bitexact": True, "pts_bitexact": True, "max_rel": 0.0,
           "h4_exceptions": [], "merged_planes": [], "invalid_ring_planes": []}
    good = []
    for i, p in enumerate(orc.paths):
        if p is None:
            assert status[i] & _lib.ST_EMPTY and n_seg[i] == 0, f"plane {i}: oracle empty, gpu n_seg={n_seg[i]}"
            continue
        assert not (status[i] & _lib.ST_EMPTY), f"plane {i}: gpu says empty"
        # --- which triangles intersect the plane, in mesh_plane order: bit exact
        s0, s1 = int(seg_off[i]), int(seg_off[i + 1])
        assert s1 - s0 == len(p.metadata["face_index"]) == n_seg[i], f"plane {i}: segment count"
        if has_segments:
            assert np.array_equal(fidx[s0:s1], p.metadata["face_index"]), f"plane {i}: face_index differs"
            og = p.metadata["segments"]
            if not np.array_equal(segs[s0:s1], og):
                rep["seg_bitexact"] = False
                assert rel_err(segs[s0:s1], og) < TIGHT, f"plane {i}: segment coordinates {rel_err(segs[s0:s1], og)}"
        rep["segments"] += s1 - s0
        if not p.info["agree"]:
            rep["h4_exceptions"].append(i)      # coordinate-hash merge != topological merge
            continue
        if p.info.get("n_merged"):              # Path.__init__'s merge_vertices fused nodes closer than ~1e-6 mm
            rep["merged_planes"].append(i)
            assert status[i] & _lib.ST_MERGED, f"plane {i}: merged vertices not flagged"
        closed = all(p.entity_closed(k) for k in range(len(p.entities)))
        if not closed:
            assert status[i] & (_lib.ST_OPEN | _lib.ST_NONMANIFOLD), f"plane {i}: open/non-manifold not flagged"
            if status[i] & _lib.ST_NONMANIFOLD:
                continue
        else:
            assert not (status[i] & (_lib.ST_OPEN | _lib.ST_NONMANIFOLD)), f"plane {i}: status {status[i]}"
        # --- connectivity, contour order, start vertex, orientation: exact (closed contours; open chains carry none)
        disc = p.discrete
        c0, c1 = int(ct_off[i]), int(ct_off[i + 1])
        assert c1 - c0 == len(disc), f"plane {i}: {c1 - c0} contours vs {len(disc)}"
        if closed:
            assert len(p.entities) == n_ent[i], f"plane {i}: {n_ent[i]} entities vs {len(p.entities)}"
        polys = p.polygons_closed
        for k, d in enumerate(disc):
            g = pts[int(ctpt[c0 + k]):int(ctpt[c0 + k + 1])]
            assert g.shape == d.shape, f"plane {i} contour {k}: {g.shape} vs {d.shape}"
            if not np.array_equal(g, d):
                rep["pts_bitexact"] = False
                assert rel_err(g, d) < TIGHT, f"plane {i} contour {k}: order/start/orientation differ"
            a_ref = polys[k].area
            assert abs(ctarea[c0 + k] - a_ref) <= 1e-12 * max(a_ref, 1.0), f"plane {i} contour {k}: area"
        if p.info.get("invalid_rings"):         # shapely would call repair_invalid here (GEOS buffer, not restated)
            rep["invalid_ring_planes"].append(i)
        rep["contours"] += c1 - c0
        assert np.array_equal(bounds[i].reshape(2, 2), p.bounds), f"plane {i}: bounds"
        assert np.array_equal(cent[i], p.centroid), f"plane {i}: centroid"
        if closed:
            good.append(i)
        else:
            assert n_ent[i] > c1 - c0, f"plane {i}: open chains not counted as entities"
    good = np.array(good, dtype=np.int64)
    names = [("ixy", _lib.ARR_IXY), ("ixy_centered", _lib.ARR_IXY_CENTERED), ("itr", _lib.ARR_ITR),
             ("itr_start", _lib.ARR_ITR_START), ("itr_centered", _lib.ARR_ITR_CENTERED),
             ("itr_centered_start", _lib.ARR_ITR_CENTERED_START)]
    if len(good) == P:          # the reference's array properties need every plane to have a section
        assert rel_err(area1, orc.areas1) < 1e-12
        for name, which in names:
            got, ref = A(which), getattr(orc, name)
            assert got.shape == ref.shape, name
            e = rel_err(got, ref)
            rep["max_rel"] = max(rep["max_rel"], e)
            assert e < 1e-9, f"{name}: rel err {e} (north-star budget {REL_TOL})"
    elif expect_all_closed:
        raise AssertionError(f"{P - len(good)} planes without a closed section")
    elif len(good):
        # some planes carry no closed section (the reference's array properties would raise): the planes that do are
        # still compared row by row, so the several-contour / open-chain cases are not waved through
        sub = oracle.slice_arrays.rows_for_paths([orc.paths[i] for i in good], orc.interp_num)
        for name, which in names:
            e = rel_err(A(which)[good], sub[name])
            rep["max_rel"] = max(rep["max_rel"], e)
            assert e < 1e-9, f"{name} (closed planes only): rel err {e}"
        assert rel_err(area1[good], sub["areas1"]) < 1e-12
    if n_angles:
        got, ref = A(_lib.ARR_RADIAL), radial_image(orc.paths, n_angles)
        e = rel_err(got[good], ref[good])
        rep["radial_rel"] = e
        assert e < 1e-9, f"radial image rel err {e}"
    rep["oracle"] = orc
    return rep
