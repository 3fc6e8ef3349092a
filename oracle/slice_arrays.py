"""Restatement of the reference's own per-slice post-processing — TEST INFRASTRUCTURE.

Follows ``/root/reference/src/shoulder/humerus/slice.py`` (this part of the path *is* in the
reference repo, so it is followed from source, not from recollection):
  heights            slice.py:10-19      z_orig = mean(zs), z_incrs = zs - z_orig
  zs                 slice.py:219-224 (Full), 248-253 (Proximal), 271-276 (Distal)
  centroids          slice.py:34-39      Path2D.centroid per plane
  areas1             slice.py:49-60      largest closed polygon if >1 entity else Path2D.area
  ixy                slice.py:65-80      outline choice + arc-length resample (slice.py:166-189)
  ixy_centered       slice.py:85-87
  itr / itr_centered slice.py:92-97, 124-134   polar, argsort by theta (slice.py:191-198)
  itr_start / itr_centered_start  slice.py:102-108, 136-144   polar, rolled to argmin theta
  cutoff window      slice.py:157-164    int() truncation, optional odd length
The ``radial_image`` at the end is NOT in the reference: it is the definition of the extra
"radius image at fixed angular steps" product that BASELINE config 2 / north_star stage 4 name.
"""
from __future__ import annotations

import numpy as np

from . import trimesh_path as tp


def full_zs(bounds, n):
    return np.linspace(0.99 * np.max(bounds[:, -1]), 0.99 * np.min(bounds[:, -1]), n)


def proximal_zs(bounds, neck_z, n):
    return np.linspace(0.99 * np.max(bounds[:, -1]), neck_z, n)


def distal_zs(bounds, n):
    return np.linspace(0.99 * np.min(bounds[:, -1]), 0, n)


def cutoff_window(length: int, cutoff, return_odd: bool = False):
    lo = int((1 - cutoff[1]) * length)
    hi = int((1 - cutoff[0]) * length)
    if return_odd and len(range(length)[lo:hi]) % 2 == 0:
        hi -= 1
    return lo, hi


def resample_closed_polyline(xy: np.ndarray, n: int) -> np.ndarray:
    step = np.sqrt((np.diff(xy, axis=0) ** 2).sum(axis=1))
    d = np.cumsum(np.r_[0, step])
    s = np.linspace(0, d.max(), n)
    return np.c_[np.interp(s, d, xy[:, 0]), np.interp(s, d, xy[:, 1])]


def polar_rows(x, y, sort: bool):
    r = np.sqrt(x ** 2 + y ** 2)
    th = np.arctan2(y, x)
    if sort:
        k = np.argsort(th)
        return np.vstack((th[k], r[k]))
    return np.vstack((th, r))


def roll_to_theta_min(pol: np.ndarray) -> np.ndarray:
    k = np.argmin(pol[0])
    return np.c_[pol[:, k:], pol[:, :k]]


def chosen_outline(path) -> np.ndarray:
    """``slice.py:70-76``: the largest closed polygon's polyline when a plane has several entities."""
    if len(path.entities) > 1:
        return path.discrete[int(np.argmax([p.area for p in path.polygons_closed]))]
    return path.discrete[0]


class OracleSlices:
    """All cached arrays of ``slice.Slices`` for one sweep, from oracle paths."""

    def __init__(self, vertices, faces, zs, interp_num, merge="hash", version="4", return_odd=False, **section_kw):
        self.zs_all = np.asarray(zs, dtype=np.float64)
        self.interp_num = int(interp_num)
        self.return_odd = return_odd
        self.z_orig = np.mean(self.zs_all)
        self.z_incrs = self.zs_all - self.z_orig
        self.paths = tp.section_multiplane(vertices, faces, [0, 0, self.z_orig], [0, 0, 1], self.z_incrs,
                                           merge=merge, version=version, **section_kw)
        self._cache = {}

    def _get(self, name, fn):
        if name not in self._cache:
            self._cache[name] = fn()
        return self._cache[name]

    @property
    def centroids(self):
        return self._get("centroids", lambda: np.array([p.centroid for p in self.paths]))

    @property
    def n_entities(self):
        return np.array([0 if p is None else len(p.entities) for p in self.paths])

    @property
    def areas1(self):
        def fn():
            out = np.zeros(len(self.paths))
            for i, p in enumerate(self.paths):
                if len(p.entities) > 1:
                    out[i] = max(q.area for q in p.polygons_closed)
                else:
                    out[i] = p.area
            return out
        return self._get("areas1", fn)

    @property
    def ixy(self):
        def fn():
            out = np.zeros((len(self.paths), 2, self.interp_num))
            for i, p in enumerate(self.paths):
                out[i] = resample_closed_polyline(chosen_outline(p), self.interp_num).T
            return out
        return self._get("ixy", fn)

    @property
    def ixy_centered(self):
        return self._get("ixy_centered", lambda: self.ixy - self.centroids[:, :, None])

    def _polar(self, src, sort, roll):
        out = np.zeros(src.shape)
        for i in range(len(src)):
            pol = polar_rows(src[i][0], src[i][1], sort)
            out[i] = roll_to_theta_min(pol) if roll else pol
        return out

    @property
    def itr(self):
        return self._get("itr", lambda: self._polar(self.ixy, True, False))

    @property
    def itr_start(self):
        return self._get("itr_start", lambda: self._polar(self.ixy, False, True))

    @property
    def itr_centered(self):
        return self._get("itr_centered", lambda: self._polar(self.ixy_centered, True, False))

    @property
    def itr_centered_start(self):
        return self._get("itr_centered_start", lambda: self._polar(self.ixy_centered, False, True))

    def window(self, arr, cutoff):
        lo, hi = cutoff_window(len(arr), cutoff, self.return_odd)
        return arr[lo:hi]


def rows_for_paths(paths, interp_num: int) -> dict:
    """The per-plane rows of every cached array for an arbitrary list of (non-None, closed) paths — what
    ``Slices`` would hold if its sweep consisted of just these planes (slice.py:34-147 is row-wise)."""
    class _Sub(OracleSlices):
        def __init__(self, paths, n):
            self.paths, self.interp_num, self.return_odd, self._cache = paths, int(n), False, {}
    s = _Sub(list(paths), interp_num)
    areas = np.array([max(q.area for q in p.polygons_closed) if len(p.entities) > 1 else p.area for p in s.paths])
    return {"areas1": areas, "centroids": s.centroids, "ixy": s.ixy, "ixy_centered": s.ixy_centered, "itr": s.itr,
            "itr_start": s.itr_start, "itr_centered": s.itr_centered, "itr_centered_start": s.itr_centered_start}


def radial_image(paths, n_angles: int) -> np.ndarray:
    """Radius image: for each plane, from the AABB centroid (the same centre ``ixy_centered``
    uses) cast ``n_angles`` rays at theta_k = -pi + 2*pi*k/n_angles and record the distance to
    the OUTERMOST crossing of the chosen outline (same outline rule as ``ixy``); 0 where a ray
    misses.  A crossing of edge p->q is accepted when den != 0, t >= 0 and 0 <= u <= 1 with
        e = q - p ; w = p - c ; den = dx*e_y - dy*e_x ; t = (w_x*e_y - w_y*e_x)/den ; u = (w_x*dy - w_y*dx)/den
    """
    k = np.arange(n_angles)
    theta = -np.pi + (2.0 * np.pi / n_angles) * k
    dx, dy = np.cos(theta), np.sin(theta)
    out = np.zeros((len(paths), n_angles))
    for i, path in enumerate(paths):
        if path is None:
            continue
        xy = chosen_outline(path)
        c = path.centroid
        p, q = xy[:-1], xy[1:]
        e = q - p
        w = p - c
        den = dx[:, None] * e[None, :, 1] - dy[:, None] * e[None, :, 0]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (w[None, :, 0] * e[None, :, 1] - w[None, :, 1] * e[None, :, 0]) / den
            u = (w[None, :, 0] * dy[:, None] - w[None, :, 1] * dx[:, None]) / den
        ok = (den != 0) & (t >= 0) & (u >= 0) & (u <= 1)
        out[i] = np.where(ok, t, 0.0).max(axis=1)
    return out
