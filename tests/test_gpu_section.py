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
    m = bone_obbs("humerus_left").mesh
    z = m.vertices[:, 2]
    for zz in (0.95 * z.max(), 0.95 * z.min(), 0.0, 41.7):
        p3 = sec.section(m, [0, 0, zz], [0, 0, 1])
        o = oracle.section_multiplane(m.vertices, m.faces, [0, 0, zz], [0, 0, 1], np.array([0.0]))[0]
        assert len(p3.entities) == len(o.entities)
        for a, b in zip(p3.discrete, o.discrete):
            assert a.shape == (len(b), 3)
            assert np.array_equal(a[:, :2], b) and np.array_equal(a[:, 2], np.full(len(b), zz))
        p2, to_3d = p3.to_planar()
        assert abs(p2.area - o.area) <= 1e-12 * o.area
        assert np.array_equal(to_3d, o.metadata["to_3D"])
    assert sec.section(m, [0, 0, 2 * z.max()], [0, 0, 1]) is None            # the plane misses the mesh


def test_hundred_z_sections_in_one_sweep(gpu_backend, bone_obbs):
    """ProxObb's loop (mesh.py:158-161): 100 sections along z, area of each."""
    m = bone_obbs("humerus_right").mesh
    zb = m.bounds[:, 2]
    zs = np.linspace(zb[0] * 0.99, zb[1] * 0.99, 100)
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
    p3 = sec.section(m, origin, n)
    o = ref[1]
    lifted = np.c_[o.vertices, np.zeros(len(o.vertices)), np.ones(len(o.vertices))].dot(o.metadata["to_3D"].T)[:, :3]
    assert p3.vertices.shape == lifted.shape
    # every section point lies on the plane and the two point sets coincide
    assert np.abs((p3.vertices - origin).dot(n)).max() < 1e-9
    d = np.abs(p3.vertices[:, None, :] - lifted[None, :, :]).sum(axis=2).min(axis=1)
    assert d.max() < 1e-8
