"""Row f1: single-plane / multi-plane sections with an arbitrary normal (reference mesh.py:95-99,158-161,
surgical_neck.py:37-50, anatomic_neck.py:160-165, arthroplasty.py:71) against the oracle's general-normal path."""
import numpy as np
import pytest

import oracle
from shoulder_b200 import section as sec

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _same_cycle(a, b, tol=1e-9):
    """closed polylines equal as cyclic sequences (same orientation), within tol relative"""
    if a.shape != b.shape:
        return False
    a0, b0 = a[:-1], b[:-1]
    scale = max(np.abs(b0).max(), 1e-300)
    k = int(np.argmin(np.abs(b0 - a0[0]).sum(axis=1)))
    return bool(np.abs(np.roll(b0, -k, axis=0) - a0).max() / scale < tol)


def test_z_sections_match_oracle(gpu_backend, bone_obbs):
    """``mesh.section(plane_normal, plane_origin)`` through the drop-in proxy against the oracle's restatement of
    ``Trimesh.section`` (3-D route: row hashes over three columns, no CCW normalisation): same polylines, same start
    vertex, same direction — asserted, not waived."""
    from shoulder_b200.mesh import GpuMesh
    base = bone_obbs("humerus_left").mesh
    m = GpuMesh(base.vertices, base.faces)
    z = m.vertices[:, 2]
    for zz in (0.95 * z.max(), 0.95 * z.min(), 0.0, 41.7, -60.0):
        p3 = m.section([0, 0, 1], [0, 0, zz])                                  # trimesh's positional order: normal first
        o = oracle.section(m.vertices, m.faces, [0, 0, 1], [0, 0, zz])
        assert len(p3.entities) == len(o.entities)
        if len(o.entities) == 1:                                               # the group stitcher's planes: exact
            a, b = p3.discrete[0], o.discrete[0]
            assert a.shape == b.shape
            assert np.array_equal(a[:, :2], b[:, :2]) and np.abs(a[:, 2] - zz).max() < 1e-12 and np.abs(b[:, 2] - zz).max() < 1e-12
        else:                                                                  # several contours: same cycles
            for a in p3.discrete:
                assert any(_same_cycle(a[:, :2], b[:, :2]) or _same_cycle(a[::-1, :2], b[:, :2]) for b in o.discrete)
        o2 = oracle.section_multiplane(m.vertices, m.faces, [0, 0, zz], [0, 0, 1], np.array([0.0]))[0]
        p2, to_3d = p3.to_planar()
        assert abs(p2.area - o2.area) <= 1e-12 * o2.area
        assert np.array_equal(to_3d, o2.metadata["to_3D"])
    assert m.section([0, 0, 1], [0, 0, 2 * z.max()]) is None                   # the plane misses the mesh
    assert m._handle is not None and m.section(plane_origin=[0, 0, 1.0], plane_normal=[0, 0, 1]) is not None      # keywords, as mesh.py:95-99
    h = m._handle
    m.section([0, 0, 1], [0, 0, 3.0])
    assert m._handle is h                                                      # uploaded once
    # the free-function form of round 1 (origin first) still answers
    assert np.array_equal(sec.section(m, [0, 0, 41.7], [0, 0, 1]).discrete[0], m.section([0, 0, 1], [0, 0, 41.7]).discrete[0])


def test_hundred_z_sections_in_one_sweep(gpu_backend, bone_obbs):
    """ProxObb's loop (mesh.py:158-161): 100 sections along z, area of each."""
    m = bone_obbs("humerus_right").mesh
    zb = m.bounds[:, 2]
    zs = np.linspace(zb[0] * 0.99, zb[1] * 0.99, 100)
    from shoulder_b200.mesh import GpuMesh
    paths = GpuMesh(m.vertices, m.faces).section_multiplane([0, 0, 0], [0, 0, 1], zs)      # the drop-in method, one upload
    assert np.array_equal(np.array([p.area for p in paths]), np.array([p.area for p in sec.section_multiplane(m, [0, 0, 0], [0, 0, 1], zs).paths()]))
    sweep = sec.section_multiplane(m, [0, 0, 0], [0, 0, 1], zs)
    got = np.array([p.area for p in sweep.paths()])
    ref = np.array([p.area for p in oracle.section_multiplane(m.vertices, m.faces, [0, 0, 0], [0, 0, 1], zs)])
    assert rel_err(got, ref) < 1e-12


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_tilted_plane_sections(gpu_backend, bone_obbs, seed):
    """anatomic_neck.py:160-165 / arthroplasty.py:71: a plane with an arbitrary normal through the head."""
    m = bone_obbs("humerus_left").mesh
    rng = np.random.default_rng(seed)
    n = rng.normal(size=3); n /= np.linalg.norm(n)
    origin = np.array([rng.uniform(-5, 5), rng.uniform(-5, 5), 0.6 * m.vertices[:, 2].max()])
    heights = np.array([-6.0, 0.0, 7.5])
    sweep = sec.section_multiplane(m, origin, n, heights)
    ref = oracle.section_multiplane(m.vertices, m.faces, origin, n, heights)
    for i, (p, o) in enumerate(zip(sweep.paths(), ref)):
        assert (p is None) == (o is None)
        if o is None:
            continue
        assert len(p.entities) == len(o.entities)
        gd, od = p.discrete, o.discrete
        for a in gd:                                  # same contours (cyclic sequences), any entity order
            assert any(_same_cycle(a, b) for b in od)
        assert abs(p.area - o.area) <= 1e-9 * o.area
        assert np.allclose(sweep.to_3D(i), o.metadata["to_3D"], atol=1e-12)
    from shoulder_b200.mesh import GpuMesh
    gm = GpuMesh(m.vertices, m.faces)
    for i, (p, o) in enumerate(zip(gm.section_multiplane(origin, n, heights), ref)):           # tilt applied on the device
        assert (p is None) == (o is None)
        if o is not None:
            assert len(p.entities) == len(o.entities) and abs(p.area - o.area) <= 1e-9 * o.area
            assert np.allclose(p.metadata["to_3D"], o.metadata["to_3D"], atol=1e-12)
    p3 = gm.section(n, origin)                                                                  # arthroplasty.py:71: (normal, point)
    o3 = oracle.section(m.vertices, m.faces, n, origin)
    assert len(p3.entities) == len(o3.entities)
    d = np.abs(p3.vertices[:, None, :] - o3.vertices[None, :, :]).sum(axis=2).min(axis=1)
    assert p3.vertices.shape == o3.vertices.shape and d.max() < 1e-8
    for a in p3.discrete:                                        # same closed polylines (start vertex / direction of a TILTED
        assert any(_same_cycle(a, b) or _same_cycle(a[::-1], b) for b in o3.discrete)      # section follow the plane frame, see mesh.py)
    o = ref[1]
    lifted = np.c_[o.vertices, np.zeros(len(o.vertices)), np.ones(len(o.vertices))].dot(o.metadata["to_3D"].T)[:, :3]
    assert p3.vertices.shape == lifted.shape
    # every section point lies on the plane and the two point sets coincide
    assert np.abs((p3.vertices - origin).dot(n)).max() < 1e-9
    d = np.abs(p3.vertices[:, None, :] - lifted[None, :, :]).sum(axis=2).min(axis=1)
    assert d.max() < 1e-8


def test_ray_casts_match_the_restated_trimesh_test(gpu_backend, bone_obbs):
    """anatomic_neck.py:184-191,217-224: rays from a point inside the head along +-normal -> where they leave the bone."""
    from oracle import ray as oray
    from shoulder_b200.mesh import GpuMesh
    base = bone_obbs("humerus_left").mesh
    m = GpuMesh(base.vertices, base.faces)
    z = m.vertices[:, 2]
    head = np.array([0.0, 0.0, 0.8 * z.max()])
    rng = np.random.default_rng(4)
    dirs = rng.normal(size=(12, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs = np.vstack([dirs, -dirs, [[0, 0, 1.0]], [[0, 0, -1.0]]])
    orgs = np.repeat(head[None, :], len(dirs), axis=0) + rng.uniform(-1, 1, (len(dirs), 3))
    loc, ray, tri = m.ray.intersects_location(orgs, dirs)
    oloc, oray_i, otri = oray.intersects_location(m.vertices, m.faces, orgs, dirs)
    k = np.lexsort((otri, oray_i))
    assert np.array_equal(ray, oray_i[k]) and np.array_equal(tri, otri[k])        # which triangles every ray meets: exact
    assert np.array_equal(loc, oloc[k])                                            # same arithmetic, same bits
    assert len(np.unique(ray)) == len(dirs)                                        # from inside a closed bone every ray leaves it
    one, r1, t1 = m.ray.intersects_location(orgs[:2], dirs[:2], multiple_hits=False)
    assert len(one) == 2
