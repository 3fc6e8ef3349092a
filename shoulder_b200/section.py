"""Single-plane and multi-plane sections with an arbitrary normal, on the same kernels.

Replaces the reference's remaining ``trimesh.Trimesh.section`` calls:
``mesh.py:95-99`` (two z-normal sections + ``to_planar``), ``mesh.py:158-161`` (100 z-normal sections — one sweep
here), ``surgical_neck.py:37-50`` (``entities`` / ``discrete``), ``anatomic_neck.py:160-165`` (``vertices``, tilted
plane) and ``arthroplasty.py:71`` (tilted plane, ``entities`` / ``discrete``).

A tilted plane is handled by moving the vertices into trimesh's own plane frame on the host
(``geometry.plane_transform``: ``align_vectors(normal, +z)`` then the origin shift — one (V,3)x(3,3) product) and
running the z-normal sweep there; ``to_3D`` is the inverse, exactly the matrix ``section_multiplane`` attaches.
Differences from trimesh that are visible: a ``Path3D``'s polylines come out counter-clockwise about the plane
normal from the minimum-rank vertex (trimesh leaves 3-D paths in DFS direction), and ``vertices`` lists the nodes
in contour order.  Point sets, connectivity and areas are the same.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .path2d import GpuPath2D, _Entity


def _align_to_z(normal: np.ndarray) -> np.ndarray:
    """trimesh ``geometry.align_vectors(normal, [0,0,1])`` (SVD bases, right-handed), as a 3x3 rotation."""
    a = np.asarray(normal, dtype=np.float64).reshape(3)
    b = np.array([0.0, 0.0, 1.0])
    au = np.linalg.svd(a.reshape(-1, 1))[0]
    bu = np.linalg.svd(b.reshape(-1, 1))[0]
    if np.linalg.det(au) < 0:
        au[:, -1] *= -1.0
    if np.linalg.det(bu) < 0:
        bu[:, -1] *= -1.0
    return bu.dot(au.T)


def plane_transform(origin, normal) -> np.ndarray:
    """4x4 ``to_2D``: moves the plane (origin, normal) onto z = 0 (trimesh ``geometry.plane_transform``)."""
    n = np.asarray(normal, dtype=np.float64).reshape(3)
    n = n / np.linalg.norm(n)
    m = np.eye(4)
    m[:3, :3] = _align_to_z(n)
    m[:3, 3] = -m[:3, :3].dot(np.asarray(origin, dtype=np.float64).reshape(3))
    return m


def _is_plus_z(normal) -> bool:
    n = np.asarray(normal, dtype=np.float64).reshape(3)
    return n[0] == 0.0 and n[1] == 0.0 and n[2] > 0.0


class SectionSweep:
    """Result of :func:`section_multiplane`: per-plane 2-D views plus the 3-D frames."""

    def __init__(self, result, heights, to_2d, origin_z):
        self.result = result
        self.heights = np.asarray(heights, dtype=np.float64)
        self._base = np.linalg.inv(to_2d) if to_2d is not None else None
        self._origin_z = origin_z

    def to_3D(self, i: int) -> np.ndarray:
        t = np.eye(4)
        if self._base is None:                        # +z normal: pure translation, as mesh_multiplane builds it
            t[2, 3] = self._origin_z + self.heights[i]
            return t
        t[2, 3] = self.heights[i]
        return self._base.dot(t)

    def paths(self):
        status = self.result.array(_lib.ARR_STATUS, 0)
        out = []
        for i in range(len(self.heights)):
            if status[i] & _lib.ST_EMPTY:
                out.append(None)
                continue
            p = GpuPath2D(self.result, 0, i, 0.0)
            p._to_3d = self.to_3D(i)
            out.append(p)
        return out


def section_multiplane(mesh, plane_origin, plane_normal, heights, interp_num: int = 2, extra_mask: int = 0) -> SectionSweep:
    """``Trimesh.section_multiplane`` for any normal.  ``mesh`` needs ``vertices`` and ``faces``."""
    v = np.asarray(mesh.vertices, dtype=np.float64)
    f = np.asarray(mesh.faces, dtype=np.int64)
    origin = np.asarray(plane_origin, dtype=np.float64).reshape(3)
    heights = np.asarray(heights, dtype=np.float64).reshape(-1)
    mask = _lib.OUT_PLANE | _lib.OUT_CONTOURS | extra_mask
    if _is_plus_z(plane_normal):
        res = _lib.sweep_batch([(v, f)], [(0, float(origin[2]), heights, interp_num)], mask, lazy=True)
        return SectionSweep(res, heights, None, float(origin[2]))
    to_2d = plane_transform(origin, plane_normal)
    vr = np.ascontiguousarray(v.dot(to_2d[:3, :3].T) + to_2d[:3, 3])
    res = _lib.sweep_batch([(vr, f)], [(0, 0.0, heights, interp_num)], mask, lazy=True)
    return SectionSweep(res, heights, to_2d, 0.0)


from .mesh import GpuMesh, GpuPath3D  # noqa: E402  (the drop-in proxy; this module keeps the free-function forms)


def section(mesh, plane_origin, plane_normal):
    """Free-function form, ORIGIN FIRST (kept for callers of round 1); the drop-in with trimesh's own signature is
    :meth:`shoulder_b200.mesh.GpuMesh.section` ``(plane_normal, plane_origin)``."""
    if isinstance(mesh, GpuMesh):
        return mesh.section(plane_normal, plane_origin)
    return GpuMesh(mesh.vertices, mesh.faces).section(plane_normal, plane_origin)


__all__ = ["section", "section_multiplane", "plane_transform", "GpuPath3D", "SectionSweep", "_Entity"]
