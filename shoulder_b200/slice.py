"""GPU slice provider — host-side mirror of reference ``src/shoulder/humerus/slice.py``.

``GpuFullSlices`` / ``GpuProximalSlices`` / ``GpuDistalSlices`` take the constructor arguments
of the reference's ``FullSlices`` (slice.py:209-224), ``ProximalSlices`` (:227-253) and
``DistalSlices`` (:256-276) and expose the accessor set of ``Slices`` (:31-155) with the same
names, argument meaning, windowing (:157-164) and quirks:

  * ``itr(cutoff)`` hands back the *cartesian* ``_ixy`` window (slice.py:99-100),
  * ``itr_start_even_theta(cutoff)`` hands back ``_itr_start`` (slice.py:121-122),
  * ``_cutoff`` truncates with ``int()`` and can drop a row to force odd length.

Everything numeric comes from one call into ``libshoulder_b200.so`` per object (or per batch of
objects, see :func:`run_batch`); arrays are fetched from the device lazily, on first access, the
way the reference memoises with ``cached_property``.  There is no CPU path.

What a provider asks the device for is what the reference's consumers read (``consumer_windows``): the Full sweep
feeds ``canal.py:40-46`` (centroids) and ``surgical_neck.py:31-34`` (areas) — plane records only; the Distal sweep
feeds ``epicondyle.py:33-43`` (37 polygons) — plane records + contours on demand; the Proximal sweep feeds
``anatomic_neck.py:34-36`` (``itr_start((0, 0.852))``: rows 88..599 of 600) and ``bicipital_groove.py:161-162``
(``itr_centered_start((0.2, 0.75))``: rows 150..479).  Only those rows are computed and cross PCIe.  Any other array or
window is still served, by a second run on first access, so the accessor set behaves like the reference's.
"""
from __future__ import annotations

from functools import cached_property

import numpy as np

from . import _lib
from .path2d import GpuPath2D

_PROFILE_ARRAYS = {
    "_ixy": _lib.ARR_IXY,
    "_ixy_centered": _lib.ARR_IXY_CENTERED,
    "_itr": _lib.ARR_ITR,
    "_itr_start": _lib.ARR_ITR_START,
    "_itr_centered": _lib.ARR_ITR_CENTERED,
    "_itr_centered_start": _lib.ARR_ITR_CENTERED_START,
}
_RUN_MASK = _lib.OUT_PLANE
_OUT_BIT = {"_ixy": _lib.OUT_IXY, "_ixy_centered": _lib.OUT_IXY_CENTERED, "_itr": _lib.OUT_ITR, "_itr_start": _lib.OUT_ITR_START,
            "_itr_centered": _lib.OUT_ITR_CENTERED, "_itr_centered_start": _lib.OUT_ITR_CENTERED_START}


class GpuSlices:
    """Common part (reference ``Slices``, slice.py:9-207)."""

    #: dtype of the profile arrays handed back.  float64 mirrors the reference; float32 (computed in fp64 on
    #: the device, rounded on store) halves the device->host bytes and stays inside north_star's 1e-5 budget.
    profile_dtype = np.float64

    #: cached array -> fractional window its consumer reads (slice.py:157-164 cutoffs); what the first run delivers.
    #: ``None`` asks for every profile array over every plane (the reference's behaviour, 3x the bytes).
    consumer_windows = None

    @classmethod
    def _run_mask(cls) -> int:
        return _RUN_MASK | (_lib.OUT_F32 if np.dtype(cls.profile_dtype) == np.float32 else 0)

    def _request_spec(self):
        """Per-sweep request (see ``_lib.make_requests``) of the first run."""
        if self.consumer_windows is None:
            return _lib.OUT_ALL_PROFILES
        n = len(self._z_incrs)
        spec = {}
        for name, cutoff in self.consumer_windows.items():
            lo, hi = int((1 - cutoff[1]) * n), int((1 - cutoff[0]) * n)
            spec[_OUT_BIT[name]] = (lo, hi)
        return spec

    def __init__(self, obb, zslice_num: int, interp_num: int, return_odd: bool = False):
        self._mesh_oriented_uobb = obb.mesh
        self.obb = obb
        self.return_odd = return_odd
        self._zslice_num = zslice_num
        self._interp_num = interp_num
        self._z_orig = np.mean(self._zs)
        self._z_incrs = self._zs - self._z_orig
        self._attached = None            # (SweepResult, sweep index) when filled by run_batch

    # ---- backend call -------------------------------------------------------------------
    def _sweep_spec(self, mesh_index: int = 0):
        return (mesh_index, float(self._z_orig), np.asarray(self._z_incrs, dtype=np.float64), int(self._interp_num))

    def _mesh_arrays(self):
        m = self.obb.mesh
        return np.asarray(m.vertices, dtype=np.float64), np.asarray(m.faces, dtype=np.int64)

    @cached_property
    def _result(self):
        if self._attached is not None:
            return self._attached
        res = _lib.sweep_batch([self._mesh_arrays()], [self._sweep_spec(0)], self._run_mask(), lazy=True,
                               requests=[self._request_spec()])
        return res, 0

    def _result_for(self, name: str):
        """A run that holds EVERY row of profile array ``name`` (made on first need, one per array)."""
        cache = self.__dict__.setdefault("_full_runs", {})
        if name not in cache:
            cache[name] = (_lib.sweep_batch([self._mesh_arrays()], [self._sweep_spec(0)], self._run_mask(), lazy=True,
                                            requests=[{_OUT_BIT[name]: (0, -1)}]), 0)
        return cache[name]

    def _profile_rows(self, name: str, lo: int, hi: int) -> np.ndarray:
        """Rows [lo, hi) of cached array ``name``: from the first run when its window holds them, else from a full run."""
        self._require_sections()
        which = _PROFILE_ARRAYS[name]
        res, k = self._result
        wlo, whi = res.window(which, k)
        if whi > wlo and wlo <= lo and hi <= whi:
            return res.array(which, k)[lo - wlo:hi - wlo]
        res, k = self._result_for(name)
        return res.array(which, k)[lo:hi]

    @cached_property
    def _result_full(self):
        """Second, on-demand run that also keeps mesh_multiplane's own outputs (lines_2D, face_index)."""
        res = _lib.sweep_batch([self._mesh_arrays()], [self._sweep_spec(0)],
                               _lib.OUT_PLANE | _lib.OUT_SEGMENTS | _lib.OUT_CONTOURS)
        return res, 0

    def _arr(self, which: int) -> np.ndarray:
        res, k = self._result
        return res.array(which, k)

    # ---- what slice.py caches -----------------------------------------------------------
    @cached_property
    def _slices(self):
        res, k = self._result
        status = self._arr(_lib.ARR_STATUS)
        return [None if status[i] & _lib.ST_EMPTY else GpuPath2D(res, k, i, float(self._zs[i]), full=lambda: self._result_full)
                for i in range(len(self._z_incrs))]

    @cached_property
    def _centroids(self):
        self._require_sections()
        return self._arr(_lib.ARR_CENTROID)

    @cached_property
    def _centroids_repeated(self):
        return np.repeat(self._centroids.reshape(-1, 2, 1), self._interp_num, axis=2)

    @cached_property
    def _areas1(self):
        self._require_sections()
        area = self._arr(_lib.ARR_AREA1)
        n_ent = self._arr(_lib.ARR_N_ENT)
        # one entity: slice.py:59 asks Path2D.area; identical to the polygon's own area.
        # several entities: slice.py:55-57 takes the largest closed polygon — what the device stored.
        assert (n_ent >= 1).all()
        return area

    @cached_property
    def _itr_start_even_theta(self):
        return np.array(self._itr_start)   # slice.py:113-119 recomputes the very same array

    def _require_sections(self):
        if self.__dict__.get("_sections_ok"):
            return
        status = self._arr(_lib.ARR_STATUS)
        no_outline = self._arr(_lib.ARR_SEL)[:, 1] == 0          # no closed contour to resample
        bad = np.nonzero((status & (_lib.ST_EMPTY | _lib.ST_NONMANIFOLD)).astype(bool) | no_outline)[0]
        if len(bad):
            # the reference fails here too (``None.centroid`` / ``p.area`` on None, slice.py:38,56)
            raise ValueError(f"planes {bad[:8].tolist()} have no closed section (status {status[bad[:8]].tolist()})")
        self._sections_ok = True

    # ---- windowing ------------------------------------------------------------------------
    def _cutoff(self, entity, cutoff: tuple):
        n = len(entity)
        lo, hi = int((1 - cutoff[1]) * n), int((1 - cutoff[0]) * n)
        if self.return_odd and len(entity[lo:hi]) % 2 == 0:
            hi -= 1
        return entity[lo:hi]

    def slices(self, cutoff: tuple):
        return self._cutoff(self._slices, cutoff)

    def centroids(self, cutoff: tuple):
        return self._cutoff(self._centroids, cutoff)

    def areas1(self, cutoff: tuple):
        return self._cutoff(self._areas1, cutoff)

    def _cut_rows(self, name: str, cutoff: tuple) -> np.ndarray:
        """``_cutoff(self.<name>, cutoff)`` without materialising rows nobody asked for."""
        n = len(self._z_incrs)
        idx = self._cutoff(range(n), cutoff)
        if len(idx) == 0:
            return self._profile_rows(name, 0, 0)
        return self._profile_rows(name, idx[0], idx[-1] + 1)

    def ixy(self, cutoff: tuple):
        return self._cut_rows("_ixy", cutoff)

    def ixy_centered(self, cutoff: tuple):
        return self._cut_rows("_ixy_centered", cutoff)

    def itr(self, cutoff: tuple) -> np.ndarray:
        return self._cut_rows("_ixy", cutoff)             # sic: slice.py:99-100

    def itr_start(self, cutoff: tuple):
        return self._cut_rows("_itr_start", cutoff)

    def itr_start_even_theta(self, cutoff: tuple):
        return self._cut_rows("_itr_start", cutoff)       # sic: slice.py:121-122

    def itr_centered(self, cutoff: tuple):
        return self._cut_rows("_itr_centered", cutoff)

    def itr_centered_start(self, cutoff: tuple):
        return self._cut_rows("_itr_centered_start", cutoff)

    def zs(self, cutoff) -> np.ndarray:
        return self._cutoff(self._zs, cutoff)

    # extra product (not in the reference): ray-cast radius image, see DESIGN.md
    def radial(self, n_angles: int = 360) -> np.ndarray:
        res = _lib.sweep_batch([self._mesh_arrays()], [self._sweep_spec(0)], _lib.OUT_PLANE | _lib.OUT_RADIAL, n_angles)
        return np.array(res.array(_lib.ARR_RADIAL, 0))


def _profile_property(name: str, which: int):
    def get(self):
        return self._profile_rows(name, 0, len(self._z_incrs))      # view of a result's pinned buffer (no host copy)
    get.__name__ = name
    prop = cached_property(get)
    prop.__set_name__(GpuSlices, name)
    return prop


for _name, _which in _PROFILE_ARRAYS.items():
    setattr(GpuSlices, _name, _profile_property(_name, _which))


class GpuFullSlices(GpuSlices):
    consumer_windows = {}                                   # canal.py:40-46, surgical_neck.py:31-34: plane records only

    def __init__(self, obb, zslice_num=200, interp_num=100, return_odd=False):
        super().__init__(obb, zslice_num, interp_num, return_odd)

    @cached_property
    def _zs(self) -> np.ndarray:
        z = self.obb.mesh.bounds[:, -1]
        return np.linspace(0.99 * np.max(z), 0.99 * np.min(z), self._zslice_num)


class GpuProximalSlices(GpuSlices):
    consumer_windows = {"_itr_start": (0.0, 0.852),         # anatomic_neck.py:34-36
                        "_itr_centered_start": (0.2, 0.75)}  # bicipital_groove.py:161-162

    def __init__(self, obb, surgical_neck, zslice_num=600, interp_num=512, return_odd=False):
        # 600 x 512 "must not change": the anatomic-neck CNN input (slice.py:236-237)
        self.surgical_neck = surgical_neck
        super().__init__(obb, zslice_num, interp_num, return_odd)

    @cached_property
    def _zs(self) -> np.ndarray:
        z = self.obb.mesh.bounds[:, -1]
        return np.linspace(0.99 * np.max(z), self.surgical_neck.neck_z, self._zslice_num)


class GpuDistalSlices(GpuSlices):
    consumer_windows = {}                                   # epicondyle.py:33-43: 37 polygons_closed[0] (contours, on demand)

    def __init__(self, obb, zslice_num=200, interp_num=500, return_odd=False):
        super().__init__(obb, zslice_num, interp_num, return_odd)

    @cached_property
    def _zs(self) -> np.ndarray:
        z = self.obb.mesh.bounds[:, -1]
        return np.linspace(0.99 * np.min(z), 0, self._zslice_num)


def run_batch(slices_objects, extra_mask: int = 0):
    """Fill many slice providers (any mix of bones / sweeps) with ONE backend call.  Objects that
    share an ``obb`` share the uploaded mesh.  Returns the shared ``SweepResult``."""
    meshes, mesh_id, sweeps = [], {}, []
    for s in slices_objects:
        key = id(s.obb.mesh)
        if key not in mesh_id:
            mesh_id[key] = len(meshes)
            meshes.append(s._mesh_arrays())
        sweeps.append(s._sweep_spec(mesh_id[key]))
    res = _lib.sweep_batch(meshes, sweeps, type(slices_objects[0])._run_mask() | extra_mask, lazy=True,
                           requests=[s._request_spec() for s in slices_objects])
    for k, s in enumerate(slices_objects):
        s._attached = (res, k)
        s.__dict__.pop("_result", None)
    return res


def install(reference_slice_module) -> None:
    """Drop-in switch: make ``shoulder.humerus.slice`` resolve to the GPU providers, so that
    ``bone.Humerus`` (bone.py:116-121) builds them instead of the trimesh-backed ones."""
    reference_slice_module.FullSlices = GpuFullSlices
    reference_slice_module.ProximalSlices = GpuProximalSlices
    reference_slice_module.DistalSlices = GpuDistalSlices
