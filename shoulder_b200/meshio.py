"""Mesh input for the slicing backend: STL parse + weld, a lightweight mesh record, the
OBB-frame stand-in and the synthetic-bone generators the BASELINE configs name.

The reference hands its hot path a ``trimesh.Trimesh`` already rotated into the oriented
bounding-box frame (reference ``src/shoulder/humerus/mesh.py:82-86``: ``apply_obb`` then an
optional 180 degree flip).  The backend reads exactly three attributes of that object —
``vertices`` (V,3) f64, ``faces`` (T,3) i64 and ``bounds`` (2,3) — (reference
``src/shoulder/humerus/slice.py:26-28,221-222,250,273``), so :class:`Mesh` carries those
and nothing else.  trimesh is not installable in this environment, therefore the frame is
produced by :class:`PcaObb` (PCA axes, AABB centred on the origin, long axis = +z), which
has the one property the slice provider relies on (``slice.py:219-224``).
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

__all__ = [
    "Mesh",
    "read_stl",
    "weld",
    "load_mesh",
    "PcaObb",
    "loop_subdivide",
    "jitter_transform",
    "synthetic_bone",
    "icosphere",
    "torus",
]


class Mesh:
    """The three attributes of ``trimesh.Trimesh`` the hot path touches."""

    def __init__(self, vertices, faces):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.faces = np.ascontiguousarray(faces, dtype=np.int64)
        if self.vertices.ndim != 2 or self.vertices.shape[1] != 3:
            raise ValueError("vertices must be (V,3)")
        if self.faces.ndim != 2 or self.faces.shape[1] != 3:
            raise ValueError("faces must be (T,3)")

    @property
    def bounds(self) -> np.ndarray:
        b = getattr(self, "_bounds", None)          # cached like trimesh's; reset by apply_transform
        if b is None or self._bounds_of is not self.vertices:
            b = np.array([self.vertices.min(axis=0), self.vertices.max(axis=0)])
            self._bounds, self._bounds_of = b, self.vertices
        return b

    def copy(self) -> "Mesh":
        return Mesh(self.vertices.copy(), self.faces.copy())

    def apply_transform(self, matrix) -> "Mesh":
        m = np.asarray(matrix, dtype=np.float64)
        self.vertices = np.ascontiguousarray(self.vertices @ m[:3, :3].T + m[:3, 3])
        if np.linalg.det(m[:3, :3]) < 0:  # keep outward winding under reflections
            self.faces = np.ascontiguousarray(self.faces[:, ::-1])
        return self

    @property
    def is_watertight(self) -> bool:
        e = np.sort(self.faces[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2), axis=1)
        _, counts = np.unique(e, axis=0, return_counts=True)
        return bool((counts == 2).all())


def read_stl(path) -> np.ndarray:
    """Binary (or ASCII) STL -> (T,3,3) float32 triangle soup, file order preserved."""
    raw = Path(path).read_bytes()
    if len(raw) >= 84:
        (n,) = struct.unpack_from("<I", raw, 80)
        if 84 + 50 * n == len(raw):
            rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                                count=n, offset=84)
            return np.array(rec["v"], dtype=np.float32)
    tris = []
    for line in raw.decode("ascii", errors="ignore").splitlines():
        tok = line.split()
        if len(tok) == 4 and tok[0] == "vertex":
            tris.append([float(t) for t in tok[1:]])
    return np.asarray(tris, dtype=np.float32).reshape(-1, 3, 3)


def weld(tris: np.ndarray):
    """Merge bit-identical corners.  Vertex ids follow first occurrence, face order is kept
    (what ``trimesh.load_mesh`` + ``merge_vertices`` gives the reference at ``mesh.py:24``;
    vertex numbering itself never reaches the slicing results)."""
    pts = np.ascontiguousarray(tris.reshape(-1, 3)) + np.float32(0.0)  # -0.0 -> +0.0
    key = pts.view(np.dtype((np.void, pts.dtype.itemsize * 3))).reshape(-1)
    _, first, inverse = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")  # first-occurrence numbering
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    vertices = pts[first[order]].astype(np.float64)
    faces = rank[inverse.reshape(-1)].reshape(-1, 3).astype(np.int64)
    return vertices, faces


def load_mesh(path) -> Mesh:
    """STL file or a ``.npz`` fixture holding ``vertices`` (f32/f64) and ``faces``."""
    path = Path(path)
    if path.suffix == ".npz":
        z = np.load(path)
        return Mesh(z["vertices"].astype(np.float64), z["faces"].astype(np.int64))
    return Mesh(*weld(read_stl(path)))


def _kasa_residual(xy: np.ndarray) -> float:
    """Algebraic least-squares circle fit residual (stand-in for ``circle_fit`` at mesh.py:102)."""
    a = np.c_[2 * xy, np.ones(len(xy))]
    b = (xy ** 2).sum(axis=1)
    sol, *_ = np.linalg.lstsq(a, b, rcond=None)
    r = np.sqrt(sol[2] + sol[0] ** 2 + sol[1] ** 2)
    return float(((np.hypot(xy[:, 0] - sol[0], xy[:, 1] - sol[1]) - r) ** 2).sum() / max(len(xy), 1))


class PcaObb:
    """Stand-in for ``mesh.FullObb`` (reference ``mesh.py:57-127``): same attributes
    (``mesh``, ``mesh_ct``, ``transform``, ``z_bounds``, ``z_length``, ``cutoff_pcts``), frame
    from PCA instead of qhull's minimum-volume box.  The humeral head is put at +z by the
    same rule the reference uses — the rounder of the two ends, judged at 0.95 of each z
    limit (``mesh.py:91-117``) — evaluated on a thin vertex band rather than a section."""

    def __init__(self, mesh_or_path, name: str | None = None):
        if isinstance(mesh_or_path, Mesh):
            self._mesh_ct = mesh_or_path
            self.name = name or "mesh"
            self.file = None
        else:
            self.file = Path(mesh_or_path)
            self.name = self.file.stem
            self._mesh_ct = load_mesh(self.file)
        self.mesh = self._mesh_ct.copy()
        self.transform = self._obb()
        self.cutoff_pcts = [0.5, 0.8]

    @property
    def mesh_ct(self) -> Mesh:
        return self._mesh_ct.copy()

    def _obb(self) -> np.ndarray:
        v = self.mesh.vertices
        c = v.mean(axis=0)
        w, vec = np.linalg.eigh(np.cov((v - c).T))
        axes = vec[:, np.argsort(w)].T  # rows: smallest .. largest variance  -> x, y, z
        for a in axes:  # deterministic sign
            if a[np.argmax(np.abs(a))] < 0:
                a *= -1
        if np.linalg.det(axes) < 0:
            axes[0] *= -1
        rot = np.eye(4)
        rot[:3, :3] = axes
        p = v @ axes.T
        mid = 0.5 * (p.min(axis=0) + p.max(axis=0))
        rot[:3, 3] = -mid
        self.mesh.apply_transform(rot)

        self.z_bounds = (self.mesh.bounds[0][-1], self.mesh.bounds[1][-1])
        self.z_length = abs(self.z_bounds[0]) + abs(self.z_bounds[1])
        humeral_end, best = 0.0, np.inf
        band = 0.01 * self.z_length
        for z_limit in self.z_bounds:
            z_slice = 0.95 * z_limit
            sel = np.abs(self.mesh.vertices[:, 2] - z_slice) < band
            res = _kasa_residual(self.mesh.vertices[sel, :2]) if sel.sum() >= 8 else np.inf
            if res < best:
                best, humeral_end = res, z_limit
        flip = np.eye(4)
        if humeral_end < 0:
            flip = np.diag([-1.0, 1.0, -1.0, 1.0])
            self.mesh.apply_transform(flip)
        return flip @ rot


def loop_subdivide(vertices: np.ndarray, faces: np.ndarray, levels: int = 1):
    """Loop subdivision of a closed manifold triangle mesh (4x faces per level), used to
    build the "1M-triangle" synthetic humerus of BASELINE config 3.  Child faces of parent
    f are emitted at 4f..4f+3 so the face order stays deterministic."""
    v = np.asarray(vertices, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    for _ in range(levels):
        nv = len(v)
        he = f[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2)           # half-edges, 3 per face
        opp = f[:, [2, 0, 1]].reshape(-1)                       # vertex opposite each half-edge
        key = np.sort(he, axis=1)
        ukey, inv = np.unique(key[:, 0] * nv + key[:, 1], return_inverse=True)
        ne = len(ukey)
        ea, eb = ukey // nv, ukey % nv
        oppsum = np.zeros((ne, 3))
        np.add.at(oppsum, inv, v[opp])
        cnt = np.bincount(inv, minlength=ne)
        if not (cnt == 2).all():
            raise ValueError("loop_subdivide needs a closed manifold mesh")
        epts = 0.375 * (v[ea] + v[eb]) + 0.125 * oppsum
        # old vertices
        nsum = np.zeros((nv, 3))
        np.add.at(nsum, ea, v[eb])
        np.add.at(nsum, eb, v[ea])
        val = np.bincount(np.r_[ea, eb], minlength=nv).astype(np.float64)
        val[val == 0] = 1.0
        beta = (0.625 - (0.375 + 0.25 * np.cos(2 * np.pi / val)) ** 2) / val
        vnew = v * (1 - val * beta)[:, None] + nsum * beta[:, None]
        m = (inv + nv).reshape(-1, 3)                           # edge-point ids per face: (01,12,20)
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        mab, mbc, mca = m[:, 0], m[:, 1], m[:, 2]
        f = np.stack([np.c_[a, mab, mca], np.c_[mab, b, mbc], np.c_[mca, mbc, c], np.c_[mab, mbc, mca]],
                     axis=1).reshape(-1, 3)
        v = np.vstack([vnew, epts])
    return v, f


def jitter_transform(bone_id: int, seed: int = 20261018) -> np.ndarray:
    """Random similarity + per-axis stretch of BASELINE config 4: rotation uniform on SO(3),
    isotropic scale U[0.85,1.15], per-axis stretch U[0.95,1.05], translation U[-50,50] mm."""
    rng = np.random.default_rng(seed + int(bone_id))
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    r = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])
    s = rng.uniform(0.85, 1.15)
    stretch = rng.uniform(0.95, 1.05, size=3)
    t = rng.uniform(-50, 50, size=3)
    m = np.eye(4)
    m[:3, :3] = r @ np.diag(s * stretch)
    m[:3, 3] = t
    return m


def synthetic_bone(base: Mesh, bone_id: int, seed: int = 20261018) -> Mesh:
    """Bone ``bone_id`` of the config-4 batch: ``base`` under :func:`jitter_transform`."""
    return base.copy().apply_transform(jitter_transform(bone_id, seed))


def icosphere(levels: int = 2, radius: float = 1.0, scale=(1.0, 1.0, 1.0)):
    """Analytic test solid (ellipsoid when ``scale`` is anisotropic)."""
    t = (1 + 5 ** 0.5) / 2
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(levels):
        nv = len(v)
        he = np.sort(f[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2), axis=1)
        ukey, inv = np.unique(he[:, 0] * nv + he[:, 1], return_inverse=True)
        mid = 0.5 * (v[ukey // nv] + v[ukey % nv])
        m = (inv + nv).reshape(-1, 3)
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        f = np.stack([np.c_[a, m[:, 0], m[:, 2]], np.c_[m[:, 0], b, m[:, 1]], np.c_[m[:, 2], m[:, 1], c],
                      np.c_[m[:, 0], m[:, 1], m[:, 2]]], axis=1).reshape(-1, 3)
        v = np.vstack([v, mid])
    v = v / np.linalg.norm(v, axis=1, keepdims=True) * radius
    return v * np.asarray(scale, dtype=np.float64), f


def torus(major: float = 3.0, minor: float = 1.0, nu: int = 48, nv: int = 24, wobble: float = 0.3):
    """Genus-1 test solid with its axis along +x, so z planes cut two separate loops.  The tube
    radius varies around the ring (``wobble``) so the two loops of a plane differ in area —
    an exact tie would leave ``argmax(area)`` (slice.py:55-57) to rounding noise."""
    u = np.linspace(0, 2 * np.pi, nu, endpoint=False)
    w = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    uu, ww = np.meshgrid(u, w, indexing="ij")
    rad = minor * (1.0 + wobble * np.cos(uu + 0.4))
    y = (major + rad * np.cos(ww)) * np.cos(uu)
    z = (major + rad * np.cos(ww)) * np.sin(uu)
    x = rad * np.sin(ww)
    v = np.c_[x.reshape(-1), y.reshape(-1), z.reshape(-1)]
    idx = np.arange(nu * nv).reshape(nu, nv)
    a = idx
    b = np.roll(idx, -1, axis=0)
    c = np.roll(np.roll(idx, -1, axis=0), -1, axis=1)
    d = np.roll(idx, -1, axis=1)
    f = np.vstack([np.c_[a.reshape(-1), b.reshape(-1), c.reshape(-1)],
                   np.c_[a.reshape(-1), c.reshape(-1), d.reshape(-1)]])
    return v, f.astype(np.int64)
