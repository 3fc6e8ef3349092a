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
            n_angles=0):
    zs = np.asarray(zs, dtype=np.float64)
    z_orig = np.mean(zs)
    return _lib.sweep_batch([(vertices, faces)], [(0, float(z_orig), zs - z_orig, interp_num)], mask, n_angles)


def compare_sweep(vertices, faces, zs, interp_num, res=None, sweep=0, n_angles=0, expect_all_closed=True):
    """Asserts parity of every array; returns a small report dict."""
    orc = oracle.OracleSlices(vertices, faces, zs, interp_num, merge="topo")
    own = res is None
    if own:
        mask = _lib.OUT_PLANE | _lib.OUT_SEGMENTS | _lib.OUT_CONTOURS | _lib.OUT_ALL_PROFILES
        if n_angles:
            mask |= _lib.OUT_RADIAL
        res = run_gpu(vertices, faces, zs, interp_num, mask, n_angles)
    A = lambda w: res.array(w, sweep)
    P = len(zs)
    n_seg, seg_off, n_ent, status = A(_lib.ARR_N_SEG), A(_lib.ARR_SEG_OFF), A(_lib.ARR_N_ENT), A(_lib.ARR_STATUS)
    fidx, segs = A(_lib.ARR_FACE_INDEX), A(_lib.ARR_SEGMENTS)
    ct_off, ctpt, ctarea, pts = A(_lib.ARR_CONTOUR_OFF), A(_lib.ARR_CONTOUR_PT_OFF), A(_lib.ARR_CONTOUR_AREA), A(_lib.ARR_POINTS)
    cent, bounds, area1 = A(_lib.ARR_CENTROID), A(_lib.ARR_BOUNDS), A(_lib.ARR_AREA1)
    rep = {"planes": P, "segments": 0, "contours": 0, "seg_bitexact": True, "pts_bitexact": True, "max_rel": 0.0,
           "h4_exceptions": []}
    good = []
    for i, p in enumerate(orc.paths):
        if p is None:
            assert status[i] & _lib.ST_EMPTY and n_seg[i] == 0, f"plane {i}: oracle empty, gpu n_seg={n_seg[i]}"
            continue
        assert not (status[i] & _lib.ST_EMPTY), f"plane {i}: gpu says empty"
        # --- which triangles intersect the plane, in mesh_plane order: bit exact
        s0, s1 = int(seg_off[i]), int(seg_off[i + 1])
        assert s1 - s0 == len(p.metadata["face_index"]) == n_seg[i], f"plane {i}: segment count"
        assert np.array_equal(fidx[s0:s1], p.metadata["face_index"]), f"plane {i}: face_index differs"
        og = p.metadata["segments"]
        if not np.array_equal(segs[s0:s1], og):
            rep["seg_bitexact"] = False
            assert rel_err(segs[s0:s1], og) < TIGHT, f"plane {i}: segment coordinates {rel_err(segs[s0:s1], og)}"
        rep["segments"] += s1 - s0
        if not p.info["agree"]:
            rep["h4_exceptions"].append(i)      # coordinate-hash merge != topological merge
            continue
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
        for k, d in enumerate(disc):
            g = pts[int(ctpt[c0 + k]):int(ctpt[c0 + k + 1])]
            assert g.shape == d.shape, f"plane {i} contour {k}: {g.shape} vs {d.shape}"
            if not np.array_equal(g, d):
                rep["pts_bitexact"] = False
                assert rel_err(g, d) < TIGHT, f"plane {i} contour {k}: order/start/orientation differ"
            a_ref = p.polygons_closed[k].area
            assert abs(ctarea[c0 + k] - a_ref) <= 1e-12 * max(a_ref, 1.0), f"plane {i} contour {k}: area"
        rep["contours"] += c1 - c0
        assert np.array_equal(bounds[i].reshape(2, 2), p.bounds), f"plane {i}: bounds"
        assert np.array_equal(cent[i], p.centroid), f"plane {i}: centroid"
        if closed:
            good.append(i)
        else:
            assert n_ent[i] > c1 - c0, f"plane {i}: open chains not counted as entities"
    good = np.array(good, dtype=np.int64)
    if len(good) == P:          # the reference's array properties need every plane to have a section
        assert rel_err(area1, orc.areas1) < 1e-12
        names = [("ixy", _lib.ARR_IXY), ("ixy_centered", _lib.ARR_IXY_CENTERED), ("itr", _lib.ARR_ITR),
                 ("itr_start", _lib.ARR_ITR_START), ("itr_centered", _lib.ARR_ITR_CENTERED),
                 ("itr_centered_start", _lib.ARR_ITR_CENTERED_START)]
        for name, which in names:
            got, ref = A(which), getattr(orc, name)
            assert got.shape == ref.shape, name
            e = rel_err(got, ref)
            rep["max_rel"] = max(rep["max_rel"], e)
            assert e < 1e-9, f"{name}: rel err {e} (north-star budget {REL_TOL})"
    elif expect_all_closed:
        raise AssertionError(f"{P - len(good)} planes without a closed section")
    if n_angles:
        got, ref = A(_lib.ARR_RADIAL), radial_image(orc.paths, n_angles)
        e = rel_err(got[good], ref[good])
        rep["radial_rel"] = e
        assert e < 1e-9, f"radial image rel err {e}"
    if own:
        res.close()
    rep["oracle"] = orc
    return rep
