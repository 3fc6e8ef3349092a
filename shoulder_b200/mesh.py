"""``GpuMesh``: stand-in for the ``trimesh.Trimesh`` the reference holds as ``obb.mesh`` / ``mesh_ct`` / ``humerus.mesh``,
for the calls the reference makes on it along the slicing path:

  ``mesh.section(plane_normal, plane_origin)``            trimesh's positional order — ``arthroplasty.py:71`` passes
                                                           ``(normal, point)`` positionally; keywords at ``mesh.py:95-99,
                                                           158-161``, ``surgical_neck.py:37-39``, ``anatomic_neck.py:160-165``
  ``mesh.section_multiplane(plane_origin, plane_normal, heights)``      ``slice.py:26-28``
  ``mesh.vertices / faces / bounds / copy() / apply_transform(m)``      ``mesh.py:36-41,82-86,112-117``, ``slice.py:221-222``

The mesh is uploaded ONCE (``shb_mesh_create``: K0 conversion + face adjacency stay in HBM); every section after that
is kernels only.  A tilted plane is brought to z = 0 on the device (``shb_section`` / ``shb_mesh_transform``).
No CPU path: every number comes from ``libshoulder_b200.so``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .path2d import GpuPath2D


def _p(a):
    return C.c_void_p(a.ctypes.data)


class _MeshHandle:
    def __init__(self, vertices=None, faces=None, adopt=None):
        _lib.init(_lib._inited if _lib._inited is not None else 0)
        if adopt is not None:                                # a mesh made on the device (shb_mesh_from_stl)
            self.h = adopt
            return
        h = C.c_void_p()
        _lib.check(_lib.load().shb_mesh_create(_p(vertices), len(vertices), _p(faces), len(faces), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h:
                _lib.load().shb_mesh_free(self.h)
                self.h = None
        except Exception:
            pass


class GpuPath3D:
    """What ``Trimesh.section`` hands back, reduced to the attributes the reference reads: ``entities`` (``len``),
    ``discrete`` (3-D polylines), ``vertices``, ``bounds``, ``centroid``, ``to_planar()``, ``metadata``."""

    def __init__(self, path2d: GpuPath2D, to_3d: np.ndarray):
        self._p, self._to_3d = path2d, to_3d

    def _lift(self, xy: np.ndarray) -> np.ndarray:
        m = self._to_3d
        out = np.empty((len(xy), 3))
        for r in range(3):                                  # elementwise, the order the device uses the other way round
            out[:, r] = (xy[:, 0] * m[r, 0] + xy[:, 1] * m[r, 1]) + m[r, 3]
        return out

    @property
    def discrete(self):
        return [self._lift(d) for d in self._p.discrete]

    @property
    def vertices(self) -> np.ndarray:
        return self._lift(self._p.vertices)

    @property
    def entities(self):
        return self._p.entities

    @property
    def bounds(self) -> np.ndarray:
        v = self.vertices
        return np.array([v.min(axis=0), v.max(axis=0)])

    @property
    def centroid(self) -> np.ndarray:
        return self.bounds.mean(axis=0)

    @property
    def metadata(self):
        return self._p.metadata

    def to_planar(self):
        """(Path2D view, to_3D).  trimesh refits its own in-plane frame here; the reference only uses frame-invariant
        quantities of the result (``area`` at mesh.py:161, the circle-fit residual of ``vertices`` at mesh.py:102)."""
        return self._p, self._to_3d


class _Ray:
    """``mesh.ray``: the one method the reference calls, ``intersects_location`` (anatomic_neck.py:184-191,217-224)."""

    def __init__(self, mesh: "GpuMesh"):
        self._mesh = mesh

    def intersects_location(self, ray_origins, ray_directions, multiple_hits: bool = True):
        o = np.ascontiguousarray(np.asarray(ray_origins, dtype=np.float64).reshape(-1, 3))
        d = np.ascontiguousarray(np.asarray(ray_directions, dtype=np.float64).reshape(-1, 3))
        cap = 64 * len(o)
        while True:
            ray = np.zeros(cap, dtype=np.int32); tri = np.zeros(cap, dtype=np.int32)
            loc = np.zeros((cap, 3)); dist = np.zeros(cap); n = C.c_int32()
            rc = _lib.load().shb_ray_cast(self._mesh.resident.h, len(o), _p(o), _p(d), cap, _p(ray), _p(tri), _p(loc), _p(dist), C.byref(n))
            if rc == -4 and n.value > cap:                   # SHB_E_CAPACITY: more hits than room
                cap = n.value
                continue
            _lib.check(rc)
            break
        k = n.value
        order = np.lexsort((tri[:k], ray[:k]))               # device order is arbitrary: (ray, triangle) ascending
        ray, tri, loc, dist = ray[:k][order], tri[:k][order], loc[:k][order], dist[:k][order]
        if not multiple_hits and k:                          # nearest hit per ray
            keep = np.array([np.nonzero(ray == r)[0][np.argmin(dist[ray == r])] for r in np.unique(ray)])
            ray, tri, loc = ray[keep], tri[keep], loc[keep]
        return loc, ray.astype(np.int64), tri.astype(np.int64)


class GpuMesh:
    def __init__(self, vertices, faces):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.faces = np.ascontiguousarray(faces, dtype=np.int64)
        self._handle = None

    @classmethod
    def from_stl(cls, stl, frame: bool = False):
        """``trimesh.load_mesh(stl_file)`` (mesh.py:24) on the device: the file's bytes are parsed and welded there
        (``shb_mesh_from_stl``) and the mesh stays resident; ``vertices`` / ``faces`` are read back once for the host-side
        attributes.  ``frame=True`` also applies the oriented frame (``SHB_STL_FRAME``) and returns ``(mesh, info)`` with
        ``info = {transform, z_bounds, z_length, flipped, residuals}``."""
        import os
        _lib.init(_lib._inited if _lib._inited is not None else 0)
        # the file goes through page-locked memory (read straight into it when a path is given): the upload then runs at
        # PCIe speed instead of through the driver's pageable staging
        if isinstance(stl, (bytes, bytearray, memoryview)):
            buf = _lib.pinned_empty((len(stl),), np.uint8)
            buf[:] = np.frombuffer(stl, dtype=np.uint8)
        else:
            buf = _lib.pinned_empty((os.path.getsize(stl),), np.uint8)
            with open(stl, "rb") as fh:
                if fh.readinto(memoryview(buf)) != len(buf):
                    raise IOError(f"short read of {stl}")
        h, nv, nf = C.c_void_p(), C.c_int64(), C.c_int64()
        fo = np.zeros(22)
        _lib.check(_lib.load().shb_mesh_from_stl(_p(buf), len(buf), _lib.STL_FRAME if frame else 0, C.byref(h), C.byref(nv), C.byref(nf), _p(fo)))
        handle = _MeshHandle(adopt=h)
        v, f = _lib.pinned_empty((nv.value, 3), np.float64), _lib.pinned_empty((nf.value, 3), np.int64)
        _lib.check(_lib.load().shb_mesh_read(h, _p(v), _p(f)))
        m = cls(v, f)
        m._handle = handle
        if not frame:
            return m
        return m, {"transform": fo[:16].reshape(4, 4).copy(), "z_bounds": (float(fo[16]), float(fo[17])), "z_length": float(fo[18]),
                   "flipped": bool(fo[19] < 0), "residuals": (float(fo[20]), float(fo[21]))}

    # ---- the attributes of trimesh.Trimesh the path reads --------------------------------------
    @property
    def ray(self) -> _Ray:
        return _Ray(self)

    @property
    def bounds(self) -> np.ndarray:
        return np.array([self.vertices.min(axis=0), self.vertices.max(axis=0)])

    def copy(self) -> "GpuMesh":
        return GpuMesh(self.vertices.copy(), self.faces.copy())

    def apply_transform(self, matrix) -> "GpuMesh":
        m = np.asarray(matrix, dtype=np.float64)
        self.vertices = np.ascontiguousarray(self.vertices @ m[:3, :3].T + m[:3, 3])
        if np.linalg.det(m[:3, :3]) < 0:
            self.faces = np.ascontiguousarray(self.faces[:, ::-1])
        self._handle = None                                  # the resident copy is stale
        return self

    @property
    def resident(self) -> _MeshHandle:
        if self._handle is None:
            self._handle = _MeshHandle(self.vertices, self.faces)
        return self._handle

    # ---- sections ---------------------------------------------------------------------------------
    def section(self, plane_normal, plane_origin, **kwargs):
        """``Trimesh.section(plane_normal, plane_origin)`` — NORMAL FIRST, as trimesh declares it."""
        n = np.ascontiguousarray(np.asarray(plane_normal, dtype=np.float64).reshape(3))
        o = np.ascontiguousarray(np.asarray(plane_origin, dtype=np.float64).reshape(3))
        t3 = np.zeros(16)
        r = C.c_void_p()
        _lib.check(_lib.load().shb_section(self.resident.h, _p(n), _p(o), _lib.OUT_PLANE | _lib.OUT_CONTOURS, _p(t3), C.byref(r)))
        res = _lib.SweepResult(r, 1)
        if res.array(_lib.ARR_STATUS, 0)[0] & _lib.ST_EMPTY:
            return None                                      # trimesh returns None when the plane misses the mesh
        to_3d = t3.reshape(4, 4)
        p2 = GpuPath2D(res, 0, 0, 0.0)
        p2._to_3d = to_3d
        return GpuPath3D(p2, to_3d)

    def section_multiplane(self, plane_origin, plane_normal, heights):
        """``Trimesh.section_multiplane(plane_origin, plane_normal, heights)``: list of Path2D views / None (slice.py:26-28).
        For a tilted normal the frame is trimesh's ``plane_transform`` (so ``metadata['to_3D']`` is trimesh's matrix)."""
        from .section import SectionSweep, plane_transform, _is_plus_z
        origin = np.asarray(plane_origin, dtype=np.float64).reshape(3)
        heights = np.ascontiguousarray(np.asarray(heights, dtype=np.float64).reshape(-1))
        lib = _lib.load()
        hoff = np.array([0, len(heights)], dtype=np.int64)
        interp = np.array([2], dtype=np.int32)
        b, r = C.c_void_p(), C.c_void_p()
        tilted = None
        try:
            if _is_plus_z(plane_normal):
                zo = np.array([float(origin[2])])
                _lib.check(lib.shb_batch_create_on(self.resident.h, 1, _p(zo), _p(heights), _p(hoff), _p(interp), C.byref(b)))
                to_2d, oz = None, float(origin[2])
            else:
                to_2d = np.ascontiguousarray(plane_transform(origin, plane_normal))
                tilted = C.c_void_p()
                _lib.check(lib.shb_mesh_transform(self.resident.h, _p(to_2d), C.byref(tilted)))
                zo = np.array([0.0])
                _lib.check(lib.shb_batch_create_on(tilted, 1, _p(zo), _p(heights), _p(hoff), _p(interp), C.byref(b)))
                oz = 0.0
            _lib.check(lib.shb_batch_run(b, _lib.OUT_PLANE | _lib.OUT_CONTOURS, 0, C.byref(r)))
        finally:
            if b:
                lib.shb_batch_free(b)
            if tilted:
                lib.shb_mesh_free(tilted)
        return SectionSweep(_lib.SweepResult(r, 1), heights, to_2d, oz).paths()


class GpuObb:
    """``mesh.FullObb`` (mesh.py:57-127) with the STL parsed, welded and framed on the device: same attributes (``mesh``,
    ``mesh_ct`` [the welded mesh in CT coordinates = the framed one under the inverse transform is NOT kept: ``mesh_ct`` is
    loaded on demand], ``transform``, ``z_bounds``, ``z_length``, ``cutoff_pcts``, ``name``, ``file``).  The frame is the PCA
    stand-in of :class:`shoulder_b200.meshio.PcaObb` (trimesh's ``apply_obb`` needs qhull)."""

    def __init__(self, stl_file, name=None):
        from pathlib import Path
        if isinstance(stl_file, (bytes, bytearray, memoryview)):
            self.file, self.name, self._raw = None, name or "mesh", bytes(stl_file)
        else:
            self.file = Path(stl_file)
            self.name, self._raw = name or self.file.stem, self.file.read_bytes()
        self.mesh, info = GpuMesh.from_stl(self._raw, frame=True)
        self.transform = info["transform"]
        self.z_bounds, self.z_length = info["z_bounds"], info["z_length"]
        self.cutoff_pcts = [0.5, 0.8]

    @property
    def mesh_ct(self) -> GpuMesh:
        return GpuMesh.from_stl(self._raw, frame=False)


def install_mesh(obb) -> None:
    """Swap ``obb.mesh`` (a trimesh.Trimesh after ``apply_obb``) for a :class:`GpuMesh` with the same vertices / faces,
    so that ``mesh.section(...)`` / ``section_multiplane(...)`` of the reference's landmark code run on the device."""
    m = obb.mesh
    obb.mesh = GpuMesh(np.asarray(m.vertices), np.asarray(m.faces))
