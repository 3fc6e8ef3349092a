"""Known answers for the oracle from analytic solids (the reference pins none)."""
import numpy as np
import pytest

import oracle
from oracle import trimesh_path as tp
from shoulder_b200 import meshio


def test_ellipsoid_sections_match_analytic_area_and_centroid():
    a, b, c = 20.0, 30.0, 170.0
    v, f = meshio.icosphere(4, 1.0, scale=(a, b, c))
    zs = np.linspace(0.9 * c, -0.9 * c, 41)
    s = oracle.OracleSlices(v, f, zs, 100)
    assert (s.n_entities == 1).all()
    exact = np.pi * a * b * (1 - (zs / c) ** 2)
    # inscribed polyhedron: area slightly below the analytic ellipse, converging with refinement
    ratio = s.areas1 / exact
    assert (ratio < 1.0).all() and (ratio > 0.985).all()
    assert np.abs(s.centroids).max() < 0.35          # AABB midpoint of a centred ellipse
    # every contour closed, CCW, starting at its minimum-rank vertex (vertex 0 of the path)
    for p in s.paths:
        d = p.discrete[0]
        assert np.array_equal(d[0], d[-1]) and tp.ring_area_signed(d) > 0
        k1, k2 = tp.rank_key(d[:-1], p.info["packed"])          # vertex 0 of lines_to_path = minimum np.unique rank
        assert np.lexsort((k2, k1))[0] == 0
        assert p.info["n_merged"] == 0                          # Path.__init__'s merge_vertices merges nothing here
    # arc-length resample: first == last sample, equal spacing along the polyline
    ixy = s.ixy
    assert np.allclose(ixy[:, :, 0], ixy[:, :, -1])


def test_torus_planes_cut_two_loops_and_area_nests():
    v, f = meshio.torus(30.0, 8.0, 64, 32)
    z = v[:, 2]
    zs = np.linspace(0.5 * z.max(), 0.5 * z.min(), 9)
    s = oracle.OracleSlices(v, f, zs, 64)
    assert (s.n_entities == 2).all()
    for p, a1 in zip(s.paths, s.areas1):
        areas = [q.area for q in p.polygons_closed]
        assert a1 == max(areas)
        assert p.area == pytest.approx(sum(areas))     # side-by-side loops: no nesting


def _cube():
    v = np.array([[x, y, z] for x in (0.0, 1.0) for y in (0.0, 1.0) for z in (0.0, 1.0)])
    f = np.array([[0, 2, 3], [0, 3, 1], [4, 5, 7], [4, 7, 6], [0, 1, 5], [0, 5, 4], [2, 6, 7], [2, 7, 3],
                  [0, 4, 6], [0, 6, 2], [1, 3, 7], [1, 7, 5]])
    return v, f


def test_cube_degenerate_planes_hit_every_case():
    v, f = _cube()
    # z = 0.5: all basic.  z = 0 (bottom face coplanar): only the +side faces keep their on-plane edge (code 16).
    # z = 1 (top face coplanar): on-plane edges belong to -side faces (code 6) -> no section at all.
    segs, _, fidx, keys, klass = tp.mesh_multiplane(v, f, [0, 0, 0.5], [0, 0, 1], np.array([0.0, -0.5, 0.5]))
    assert (klass[0] == tp.CLASS_BASIC).all() and len(fidx[0]) == 8
    assert (klass[1] == tp.CLASS_EDGE).all() and len(fidx[1]) == 4
    assert len(fidx[2]) == 0
    paths = tp.section_multiplane(v, f, [0, 0, 0.5], [0, 0, 1], np.array([0.0, -0.5, 0.5]))
    assert paths[2] is None
    for p in paths[:2]:
        assert len(p.entities) == 1 and p.polygons_closed[0].area == pytest.approx(1.0)
        assert p.centroid == pytest.approx([0.5, 0.5])


def test_tilted_cube_plane_through_vertices_uses_vertex_case():
    v, f = _cube()
    # rotate so that a body diagonal is +z: planes through the 3+3 mid vertices hit the on-vertex case (code 8)
    d = np.array([1.0, 1.0, 1.0]) / np.sqrt(3)
    x = np.cross(d, [0, 0, 1.0]); x /= np.linalg.norm(x)
    r = np.stack([x, np.cross(d, x), d])
    vr = (v - 0.5) @ r.T
    zmid = np.sort(np.unique(np.round(vr[:, 2], 12)))[1]
    segs, _, fidx, keys, klass = tp.mesh_multiplane(vr, f, [0, 0, 0], [0, 0, 1], np.array([zmid]))
    assert (klass[0] == tp.CLASS_VERTEX).sum() > 0
    p = tp.section_multiplane(vr, f, [0, 0, 0], [0, 0, 1], np.array([zmid]))[0]
    assert len(p.entities) == 1 and p.entity_closed(0)
    assert p.info["agree"]


def test_rank_key_reproduces_np_unique_order():
    rng = np.random.default_rng(3)
    for scale, version in ((10.0, "4"), (40.0, "4"), (10.0, "3"), (40.0, "3")):
        pts = rng.uniform(-scale, scale, size=(500, 2))
        pts[::7, 0] = pts[3, 0]                       # shared x: forces ties on the first word
        h, packed = tp.hashable_rows(pts, version)
        assert packed == (scale < 21.0)
        order_ref = np.argsort(h, kind="stable") if packed else np.argsort(h, kind="stable")
        k1, k2 = tp.rank_key(pts, packed, version)
        order = np.lexsort((k2, k1))
        assert np.array_equal(h[order_ref], h[order])


def test_hash_merge_equals_topological_merge_on_a_bone(bone_obbs):
    m = bone_obbs("humerus_left_trab").mesh
    zs = np.linspace(0.99 * m.vertices[:, 2].max(), 0.99 * m.vertices[:, 2].min(), 60)
    a = oracle.OracleSlices(m.vertices, m.faces, zs, 50, merge="hash")
    b = oracle.OracleSlices(m.vertices, m.faces, zs, 50, merge="topo")
    assert all(p.info["agree"] for p in a.paths)
    for p, q in zip(a.paths, b.paths):
        assert np.array_equal(p.vertices, q.vertices)
        assert len(p.entities) == len(q.entities)
        assert all(np.array_equal(x, y) for x, y in zip(p.entities, q.entities))
    assert (a.n_entities > 1).any()                    # the trabecular bone has multi-contour planes
