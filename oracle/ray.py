"""Ray - mesh queries as trimesh's numpy backend answers them — TEST INFRASTRUCTURE ([RECALLED], parity unpinned).

``mesh.ray.intersects_location(ray_origins, ray_directions)`` (reference: anatomic_neck.py:184-191,217-224) ->
``ray.ray_triangle.ray_triangle_id``: the ray meets the plane of every candidate triangle (``planes_lines``: skipped when
``|direction . normal| <= 1e-5``), the point counts when its barycentric coordinates (``points_to_barycentric``, Cramer)
are inside ``(-tol.zero, 1 + tol.zero)`` and its distance along the ray is ``> -1e-6``.  trimesh prefilters candidate
triangles with an R-tree (conservative), so brute force over all triangles finds the same hits; their ORDER in trimesh is
the R-tree's and is not restated — both sides are compared sorted by (ray, triangle)."""
import numpy as np

TOL_ZERO = 1e-13


def _dd(a, b):
    return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]


def intersects_location(vertices, faces, ray_origins, ray_directions):
    v = np.asarray(vertices, dtype=np.float64)
    tri = v[np.asarray(faces, dtype=np.int64)]
    e0, e1 = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    cr = np.cross(e0, e1)
    nn = np.sqrt(_dd(cr, cr))
    ok = nn > 1e-13
    normal = np.zeros_like(cr)
    normal[ok] = cr[ok] / nn[ok, None]
    locs, rays, tris = [], [], []
    for r, (o, d) in enumerate(zip(np.asarray(ray_origins, dtype=np.float64).reshape(-1, 3), np.asarray(ray_directions, dtype=np.float64).reshape(-1, 3))):
        p_ori = _dd(tri[:, 0] - o, normal)
        p_dir = _dd(np.broadcast_to(d, normal.shape), normal)
        valid = ok & (np.abs(p_dir) > 1e-5)
        idx = np.nonzero(valid)[0]
        dist = p_ori[idx] / p_dir[idx]
        on = d[None, :] * dist[:, None] + o
        w = on - tri[idx, 0]
        a, b = e0[idx], e1[idx]
        d00, d01, d02, d11, d12 = _dd(a, a), _dd(a, b), _dd(a, w), _dd(b, b), _dd(b, w)
        inv = 1.0 / (d00 * d11 - d01 * d01)
        b2 = (d00 * d12 - d01 * d02) * inv
        b1 = (d11 * d02 - d01 * d12) * inv
        b0 = 1 - b1 - b2
        bary = np.c_[b0, b1, b2]
        hit = (bary > -TOL_ZERO).all(axis=1) & (bary < 1 + TOL_ZERO).all(axis=1)
        fwd = _dd(on - o, np.broadcast_to(d, on.shape)) > -1e-6
        keep = hit & fwd
        locs.append(on[keep]); tris.append(idx[keep]); rays.append(np.full(int(keep.sum()), r))
    return np.vstack(locs), np.concatenate(rays).astype(np.int64), np.concatenate(tris).astype(np.int64)
