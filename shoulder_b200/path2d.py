"""Per-plane view with the ``trimesh.path.Path2D`` attributes the reference reads.

Consumers in the reference: ``slice.py:38,53-59,70-76`` (centroid, entities, polygons_closed,
area, discrete), ``canal.py:46`` (centroid), ``epicondyle.py:36,43`` (polygons_closed[0]),
``mesh.py:102`` (vertices), plus ``metadata['face_index']`` / ``['to_3D']`` which
``section_multiplane`` attaches.  Everything is a lazy view into one ``SweepResult``.
"""
from __future__ import annotations

import numpy as np

from . import _lib

try:  # real shapely polygons when the host has shapely (epicondyle.py needs GEOS operations)
    from shapely.geometry import Polygon as _ShapelyPolygon
except Exception:  # pragma: no cover - not installed in the build image
    _ShapelyPolygon = None


class _Polygon:
    """Minimal stand-in when shapely is absent: the ring and the area the device computed."""

    def __init__(self, ring: np.ndarray, area: float):
        self.ring = ring
        self.area = float(area)

    @property
    def exterior_coords(self) -> np.ndarray:
        return self.ring


class _Entity:
    """Closed polyline entity (``trimesh.path.entities.Line``): only ``len(entities)`` and
    ``closed`` / ``points`` are meaningful."""

    closed = True

    def __init__(self, points: np.ndarray):
        self.points = points


def _point_in_ring(p, ring) -> bool:
    x, y = p
    x0, y0, x1, y1 = ring[:-1, 0], ring[:-1, 1], ring[1:, 0], ring[1:, 1]
    cond = (y0 > y) != (y1 > y)
    with np.errstate(divide="ignore", invalid="ignore"):
        xint = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
    return bool((cond & (x < xint)).sum() % 2 == 1)


class GpuPath2D:
    def __init__(self, result: "_lib.SweepResult", sweep: int, plane: int, z: float, full=None):
        self._res, self._k, self._i, self._z = result, sweep, plane, z
        self._full = full        # callable -> (SweepResult, sweep) of a run that kept segments / face_index

    def _a(self, which):
        return self._res.array(which, self._k)

    # ---- per-plane scalars ------------------------------------------------------------------
    @property
    def status(self) -> int:
        return int(self._a(_lib.ARR_STATUS)[self._i])

    @property
    def bounds(self) -> np.ndarray:
        return np.array(self._a(_lib.ARR_BOUNDS)[self._i])

    @property
    def centroid(self) -> np.ndarray:
        return np.array(self._a(_lib.ARR_CENTROID)[self._i])

    # ---- contours ---------------------------------------------------------------------------
    def _contour_range(self):
        off = self._a(_lib.ARR_CONTOUR_OFF)
        return int(off[self._i]), int(off[self._i + 1])

    @property
    def discrete(self):
        c0, c1 = self._contour_range()
        ptoff = self._a(_lib.ARR_CONTOUR_PT_OFF)
        pts = self._a(_lib.ARR_POINTS)
        return [np.array(pts[int(ptoff[c]):int(ptoff[c + 1])]) for c in range(c0, c1)]

    @property
    def entities(self):
        n = int(self._a(_lib.ARR_N_ENT)[self._i])
        lens = [len(d) for d in self.discrete] if n else []
        out, first = [], 0
        for m in lens:
            idx = np.r_[np.arange(first, first + m - 1), first]
            out.append(_Entity(idx))
            first += m - 1
        for _ in range(n - len(out)):                 # open chains (non-watertight input): counted, no geometry
            e = _Entity(np.zeros(0, dtype=np.int64))
            e.closed = False
            out.append(e)
        return out

    @property
    def vertices(self) -> np.ndarray:
        d = self.discrete
        return np.vstack([c[:-1] for c in d]) if d else np.zeros((0, 2))

    @property
    def polygons_closed(self):
        c0, c1 = self._contour_range()
        areas = self._a(_lib.ARR_CONTOUR_AREA)
        out = []
        for c, ring in zip(range(c0, c1), self.discrete):
            if len(ring) < 4:
                out.append(None)
            elif _ShapelyPolygon is not None:
                out.append(_ShapelyPolygon(ring))
            else:
                out.append(_Polygon(ring, areas[c]))
        return out

    @property
    def area(self) -> float:
        """``Path2D.area``: shells minus the holes directly inside them (``polygons_full``)."""
        polys = [p for p in self.polygons_closed if p is not None]
        if len(polys) == 1:
            return float(polys[0].area)
        rings = [np.asarray(p.exterior.coords) if _ShapelyPolygon is not None and isinstance(p, _ShapelyPolygon)
                 else p.ring for p in polys]
        depth = [sum(_point_in_ring(rings[i][0], rings[j]) for j in range(len(rings)) if j != i)
                 for i in range(len(rings))]
        total = 0.0
        for i, p in enumerate(polys):
            if depth[i] % 2:
                continue
            total += p.area
            for j, q in enumerate(polys):
                if depth[j] == depth[i] + 1 and _point_in_ring(rings[j][0], rings[i]):
                    total -= q.area
        return float(total)

    # ---- what section_multiplane attaches ---------------------------------------------------
    @property
    def metadata(self):
        res, k = self._full() if self._full is not None else (self._res, self._k)
        off = res.array(_lib.ARR_SEG_OFF, k)
        s0, s1 = int(off[self._i]), int(off[self._i + 1])
        to_3d = getattr(self, "_to_3d", None)
        if to_3d is None:
            to_3d = np.eye(4)
            to_3d[2, 3] = self._z
        meta = {"to_3D": to_3d}
        try:
            meta["face_index"] = np.array(res.array(_lib.ARR_FACE_INDEX, k)[s0:s1], dtype=np.int64)
            meta["segments"] = np.array(res.array(_lib.ARR_SEGMENTS, k)[s0:s1])
        except _lib.BackendError:
            pass                                   # the run did not keep mesh_multiplane's own outputs
        return meta
