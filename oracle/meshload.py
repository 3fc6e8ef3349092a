"""TEST INFRASTRUCTURE (CPU oracle, never on the product path): numpy restatement of what the reference does to an STL
file before the slicing path starts (scope row f2).

* ``load_stl`` / ``merge_vertices`` follow trimesh (third-party, absent here; pinned 3.23.5 in ``poetry.lock:4145-4146``,
  ``>=4.0.0`` in ``pyproject.toml:12``) as ``trimesh.load_mesh`` runs them at reference ``mesh.py:24`` [RECALLED, parity
  unpinned like the rest of the trimesh half]: ``exchange/stl.py load_stl_binary`` (80-byte header, uint32 triangle count,
  records of normal / 3 float32 corners / uint16 attribute; ``faces = arange(3 T).reshape(-1, 3)``) and
  ``Trimesh(process=True)`` -> ``grouping.merge_vertices``: ``(vertices * 10**8).round().astype(int64)`` rows hashed,
  ``unique_rows(keep_order=True)`` -> vertices in order of first occurrence holding the first occurrence's coordinates.
* ``pca_frame`` is the stand-in for ``mesh.apply_obb()`` (reference ``mesh.py:82``; qhull's minimum-volume box cannot be
  restated) that every config of this repo uses, plus the reference's own end test (``mesh.py:91-117``): the definition the
  device kernels (``csrc/shb_meshio.cu``) are held to.
"""
from __future__ import annotations

import numpy as np

STL_RECORD = np.dtype([("normals", "<f4", 3), ("vertices", "<f4", (3, 3)), ("attributes", "<u2")])


def load_stl(raw: bytes):
    """bytes of a binary STL -> (vertices (3T,3) float64, faces (T,3) int64), unmerged."""
    if len(raw) < 84:
        raise ValueError("not a binary STL")
    n = int(np.frombuffer(raw, dtype="<u4", count=1, offset=80)[0])
    if 84 + 50 * n != len(raw):
        raise ValueError("binary STL length does not match its header")
    blob = np.frombuffer(raw, dtype=STL_RECORD, count=n, offset=84)
    vertices = blob["vertices"].reshape(-1, 3).astype(np.float64)
    faces = np.arange(3 * n, dtype=np.int64).reshape(-1, 3)
    return vertices, faces


def merge_vertices(vertices: np.ndarray, faces: np.ndarray, digits: int = 8):
    """trimesh ``merge_vertices``: one vertex per 10**-digits cell, first-occurrence order and coordinates."""
    referenced = np.zeros(len(vertices), dtype=bool)
    referenced[faces] = True
    stacked = (vertices * 10 ** digits).round().astype(np.int64)
    rows = np.ascontiguousarray(stacked[referenced])
    keys = rows.view(np.dtype((np.void, rows.dtype.itemsize * 3))).reshape(-1)
    _, unique, inverse = np.unique(keys, return_index=True, return_inverse=True)
    order = unique.argsort()                               # keep_order=True: unique rows in order of first occurrence
    lookup = np.empty(len(order), dtype=np.int64)
    lookup[order] = np.arange(len(order))
    unique, inverse = unique[order], lookup[inverse.reshape(-1)]
    full_inverse = np.zeros(len(vertices), dtype=np.int64)
    full_inverse[referenced] = inverse
    mask = np.zeros(len(vertices), dtype=bool)
    mask[np.nonzero(referenced)[0][unique]] = True
    return vertices[mask], full_inverse[faces]


def load_mesh(raw: bytes):
    return merge_vertices(*load_stl(raw))


def encode_stl(vertices: np.ndarray, faces: np.ndarray) -> bytes:
    """(V,3), (T,3) -> bytes of a binary STL (float32 corners, zero normals): test input generator."""
    rec = np.zeros(len(faces), dtype=STL_RECORD)
    rec["vertices"] = np.asarray(vertices, dtype=np.float32)[np.asarray(faces)]
    return b"\0" * 80 + np.uint32(len(faces)).tobytes() + rec.tobytes()


def kasa_residual(xy: np.ndarray) -> float:
    a = np.c_[2 * xy, np.ones(len(xy))]
    b = (xy ** 2).sum(axis=1)
    sol, *_ = np.linalg.lstsq(a, b, rcond=None)
    r = np.sqrt(sol[2] + sol[0] ** 2 + sol[1] ** 2)
    return float(((np.hypot(xy[:, 0] - sol[0], xy[:, 1] - sol[1]) - r) ** 2).sum() / max(len(xy), 1))


def pca_frame(vertices: np.ndarray):
    """-> (transform (4,4), framed vertices, z_bounds, z_length, residuals of the two ends)."""
    v = np.asarray(vertices, dtype=np.float64)
    c = v.mean(axis=0)
    w, vec = np.linalg.eigh(np.cov((v - c).T))
    axes = vec[:, np.argsort(w)].T.copy()
    for a in axes:
        if a[np.argmax(np.abs(a))] < 0:
            a *= -1
    if np.linalg.det(axes) < 0:
        axes[0] *= -1
    rot = np.eye(4)
    rot[:3, :3] = axes
    p = v @ axes.T
    rot[:3, 3] = -0.5 * (p.min(axis=0) + p.max(axis=0))
    out = v @ rot[:3, :3].T + rot[:3, 3]
    z_bounds = (out[:, 2].min(), out[:, 2].max())
    z_length = abs(z_bounds[0]) + abs(z_bounds[1])
    humeral_end, best, resid = 0.0, np.inf, []
    for z_limit in z_bounds:
        sel = np.abs(out[:, 2] - 0.95 * z_limit) < 0.01 * z_length
        res = kasa_residual(out[sel, :2]) if sel.sum() >= 8 else np.inf
        resid.append(res)
        if res < best:
            best, humeral_end = res, z_limit
    flip = np.eye(4)
    if humeral_end < 0:
        flip = np.diag([-1.0, 1.0, -1.0, 1.0])
        out = out @ flip[:3, :3].T
    return flip @ rot, out, z_bounds, z_length, resid
