"""numpy float64 restatement of the trimesh calls the reference's slice provider makes.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED: trimesh is not available
here, so every function below is a restatement of trimesh's published source, written from
recollection and checked only against analytic solids and its own
internal cross-checks.  Call site in the reference: ``src/shoulder/humerus/slice.py:24-28``::

    self.obb.mesh.section_multiplane(plane_origin=[0,0,z_orig], plane_normal=[0,0,1], heights=z_incrs)

Cost structure is kept on purpose (cached vertex dots, then one full O(T) classification per
plane inside a Python loop, then per-plane unique / DFS): this module doubles as the CPU
baseline and must cost what the reference path costs, not what an optimised slicer would.

Arithmetic conventions fixed here where numpy/BLAS leave them open (all float64):
  * ``unitize``:  n2 = (vx*vx + vy*vy) + vz*vz ; inv = 1/sqrt(n2) ; d = v * inv
  * ``plane_lines``:  t = n.(o - p0), b = n.d, X = p0 + (t/b) * d, each product / sum rounded
    separately (no FMA); for the +z normal  t = (o_z - p0_z)  and  b = d_z  exactly.
The CUDA kernels follow the same operation order with contraction disabled, so coordinates
are expected to agree to the last bit on the z-normal path.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import depth_first_order

TOL_MERGE = 1e-8   # trimesh.constants.tol.merge
TOL_ZERO = 1e-13   # trimesh.constants.tol.zero  (finfo(float64).resolution * 100)

CLASS_BASIC, CLASS_VERTEX, CLASS_EDGE = 0, 1, 2


# --------------------------------------------------------------------------------------
# intersections.mesh_plane / plane_lines
# --------------------------------------------------------------------------------------
def _unitize_rows(v: np.ndarray):
    n2 = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]
    norm = np.sqrt(n2)
    valid = norm > TOL_ZERO
    inv = np.zeros_like(norm)
    inv[valid] = 1.0 / norm[valid]
    return v * inv[:, None]


def _dot3(n, a):
    """n . a  for rows of ``a``; exact pick of the z component when n == +z."""
    if n[0] == 0.0 and n[1] == 0.0 and n[2] == 1.0:
        return a[:, 2].copy()
    return (n[0] * a[:, 0] + n[1] * a[:, 1]) + n[2] * a[:, 2]


def plane_lines(plane_origin, plane_normal, p0, p1):
    """trimesh ``intersections.plane_lines(..., line_segments=False)``: where the infinite line
    p0->p1 meets the plane.  Returns (points, valid)."""
    d = _unitize_rows(p1 - p0)
    t = _dot3(plane_normal, plane_origin[None, :] - p0)
    b = _dot3(plane_normal, d)
    valid = np.abs(b) > TOL_ZERO
    dist = t[valid] / b[valid]
    return p0[valid] + dist[:, None] * d[valid], valid


def _case_codes(signs_f: np.ndarray) -> np.ndarray:
    s = np.sort(signs_f, axis=1).astype(np.int64)
    return 14 + (s[:, 0] << 3) + (s[:, 1] << 2) + (s[:, 2] << 1)


def mesh_plane(vertices, faces, plane_origin, plane_normal, cached_dots=None):
    """trimesh ``intersections.mesh_plane(return_faces=True)``.

    Returns ``lines`` (n,2,3) f64, ``face_index`` (n,) i64, ``keys`` (n,2,2) i64 and ``klass``
    (n,) i8.  ``keys[i,j]`` names the mesh feature endpoint j lies on: the sorted vertex pair
    of the crossed mesh edge, or (v,v) for an on-plane mesh vertex — the topological identity
    the GPU stitcher merges on; trimesh itself only has the coordinates.
    """
    vertices = np.asarray(vertices, dtype=np.float64)
    faces = np.asarray(faces, dtype=np.int64)
    plane_origin = np.asarray(plane_origin, dtype=np.float64).reshape(3)
    plane_normal = np.asarray(plane_normal, dtype=np.float64).reshape(3)
    dots = cached_dots if cached_dots is not None else _dot3(plane_normal, vertices - plane_origin)

    signs_v = np.zeros(len(vertices), dtype=np.int8)
    signs_v[dots < -TOL_MERGE] = -1
    signs_v[dots > TOL_MERGE] = 1
    signs = signs_v[faces]                              # the O(T) gather, once per plane
    coded = _case_codes(signs)
    basic = (coded == 4) | (coded == 12)
    on_vertex = coded == 8
    on_edge = coded == 16                               # not 6: only the +side face keeps an on-plane edge

    out_l, out_k = [], []
    # --- basic: lone-sign vertex u; segment = [X(u->next), X(u->next2)] in cyclic face order
    fb, sb = faces[basic], signs[basic]
    if len(fb):
        lone = np.zeros_like(sb, dtype=bool)
        for value in (-1, 1):
            test = sb == value
            ok = test.sum(axis=1) == 1
            lone[ok] = test[ok]
        k = np.argmax(lone, axis=1)
        rows = np.arange(len(fb))
        u, n1, n2 = fb[rows, k], fb[rows, (k + 1) % 3], fb[rows, (k + 2) % 3]
        e0 = np.column_stack((u, n1, u, n2)).reshape(-1, 2)
        pts, valid = plane_lines(plane_origin, plane_normal, vertices[e0[:, 0]], vertices[e0[:, 1]])
        assert valid.all(), "culling broken: basic face without a valid crossing"
        out_l.append(pts.reshape(-1, 2, 3))
        out_k.append(np.sort(e0, axis=1).reshape(-1, 2, 2))
    # --- one vertex on the plane, the other two on different sides
    fv, sv = faces[on_vertex], signs[on_vertex]
    keep_v = np.ones(len(fv), dtype=bool)
    if len(fv):
        von = fv[sv == 0]
        thru = fv[sv != 0].reshape(-1, 2)
        pts, valid = plane_lines(plane_origin, plane_normal, vertices[thru[:, 0]], vertices[thru[:, 1]])
        keep_v = valid
        out_l.append(np.column_stack((vertices[von[valid]], pts)).reshape(-1, 2, 3))
        kk = np.stack((np.column_stack((von, von)), np.sort(thru, axis=1)), axis=1)
        out_k.append(kk[valid])
    # --- two vertices on the plane, third on the + side
    fe, se = faces[on_edge], signs[on_edge]
    if len(fe):
        ed = fe[se == 0].reshape(-1, 2)
        out_l.append(vertices[ed])
        out_k.append(np.stack((np.column_stack((ed[:, 0], ed[:, 0])), np.column_stack((ed[:, 1], ed[:, 1]))), axis=1))

    idx_b, idx_v, idx_e = np.nonzero(basic)[0], np.nonzero(on_vertex)[0][keep_v], np.nonzero(on_edge)[0]
    face_index = np.hstack((idx_b, idx_v, idx_e)).astype(np.int64)
    klass = np.hstack((np.full(len(idx_b), CLASS_BASIC), np.full(len(idx_v), CLASS_VERTEX),
                       np.full(len(idx_e), CLASS_EDGE))).astype(np.int8)
    if out_l:
        lines = np.vstack(out_l)
        keys = np.vstack(out_k).astype(np.int64)
    else:
        lines = np.zeros((0, 2, 3))
        keys = np.zeros((0, 2, 2), dtype=np.int64)
    return lines, face_index, keys, klass


# --------------------------------------------------------------------------------------
# geometry.align_vectors / plane_transform  (only the +z normal is on the hot path)
# --------------------------------------------------------------------------------------
def align_vectors(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(3)
    b = np.asarray(b, dtype=np.float64).reshape(3)
    au = np.linalg.svd(a.reshape(-1, 1))[0]
    bu = np.linalg.svd(b.reshape(-1, 1))[0]
    if np.linalg.det(au) < 0:
        au[:, -1] *= -1.0
    if np.linalg.det(bu) < 0:
        bu[:, -1] *= -1.0
    m = np.eye(4)
    m[:3, :3] = bu.dot(au.T)
    return m


def plane_transform(origin, normal):
    m = align_vectors(normal, [0.0, 0.0, 1.0])
    m[:3, 3] = -np.dot(m, np.append(origin, 1.0))[:3]
    return m


def mesh_multiplane(vertices, faces, plane_origin, plane_normal, heights):
    """trimesh ``intersections.mesh_multiplane``.  Per plane: 2-D segments (n,2,2), the 4x4
    ``to_3D``, face_index (n,), plus this oracle's extras (keys, klass)."""
    vertices = np.asarray(vertices, dtype=np.float64)
    faces = np.asarray(faces, dtype=np.int64)
    plane_normal = np.asarray(plane_normal, dtype=np.float64).reshape(3)
    plane_normal = plane_normal * (1.0 / np.sqrt(float((plane_normal * plane_normal).sum())))
    plane_origin = np.asarray(plane_origin, dtype=np.float64).reshape(3)
    heights = np.asarray(heights, dtype=np.float64)

    vertex_dots = _dot3(plane_normal, vertices - plane_origin)          # once per sweep
    z_normal = plane_normal[0] == 0.0 and plane_normal[1] == 0.0 and plane_normal[2] == 1.0
    base = np.linalg.inv(plane_transform(plane_origin, plane_normal))
    translation = np.eye(4)
    segments, transforms, face_index, keys, klass = [], [], [], [], []
    for height in heights:                                                # the reference's Python loop
        new_origin = plane_origin + plane_normal * height
        new_dots = vertex_dots - height                                   # two-step subtraction (H2)
        lines, idx, kk, cl = mesh_plane(vertices, faces, new_origin, plane_normal, cached_dots=new_dots)
        translation[2, 3] = height
        to_3d = np.dot(base, translation)
        transforms.append(to_3d)
        if z_normal:                                                      # rotation part is exactly I
            lines_2d = np.ascontiguousarray(lines[:, :, :2])
        else:
            to_2d = np.linalg.inv(to_3d)
            flat = lines.reshape(-1, 3)
            lines_2d = (np.dot(to_2d, np.column_stack((flat, np.ones(len(flat)))).T).T[:, :2]).reshape(-1, 2, 2)
        segments.append(lines_2d)
        face_index.append(idx)
        keys.append(kk)
        klass.append(cl)
    return segments, np.array(transforms, dtype=np.float64), face_index, keys, klass


# --------------------------------------------------------------------------------------
# grouping.float_to_int / hashable_rows / unique_rows
# --------------------------------------------------------------------------------------
def float_to_int(data, digits: int = 8):
    return np.round(np.asarray(data, dtype=np.float64) * 10 ** digits - 1e-6).astype(np.int64)


def hashable_rows(data, version: str = "4"):
    """Row hash of (n,2) float data.  Returns (hash array, packed flag).  ``version`` selects
    trimesh 4.x (declared requirement, default) or 3.23.5 (stale lock) semantics."""
    q = float_to_int(data)
    if len(q) == 0:
        return np.zeros(0, dtype=np.uint64), True
    return _hash_int_rows(q, version)


def rank_key(points, packed: bool, version: str = "4"):
    """The device-side rank rule, as two uint64 words whose lexicographic order equals the
    ``np.unique`` order of :func:`hashable_rows` (checked in tests/test_oracle_rank.py).
    packed 4.x: the uint64 hash itself; packed 3.x: int64 order -> flip the sign bit;
    void: memcmp over little-endian int64 bytes -> byte-swapped words, x then y."""
    q = float_to_int(points)
    if packed:
        if version == "4":
            u = (q + (2 ** 31 + 1)).astype(np.uint64)
            return u[:, 0] ^ (u[:, 1] << np.uint64(32)), np.zeros(len(q), dtype=np.uint64)
        h = q[:, 0] ^ (q[:, 1] << 32)
        return h.view(np.uint64) ^ np.uint64(1 << 63), np.zeros(len(q), dtype=np.uint64)
    u = np.ascontiguousarray(q).view(np.uint64)
    return u[:, 0].byteswap(), u[:, 1].byteswap()


# --------------------------------------------------------------------------------------
# path.exchange.misc.lines_to_path / edges_to_path, graph.traversals / fill / split
# --------------------------------------------------------------------------------------
def _edge_hash(e: np.ndarray) -> np.ndarray:
    return e[:, 0].astype(np.int64) * np.int64(1 << 32) + e[:, 1].astype(np.int64)


def _traversals_dfs(edges: np.ndarray):
    edges = np.array(edges, dtype=np.int64)
    edges.sort(axis=1)
    nodes = set(edges.reshape(-1))
    count = int(edges.max()) + 1
    graph = coo_matrix((np.ones(len(edges), dtype=bool), (edges[:, 0], edges[:, 1])), dtype=bool, shape=(count, count))
    out = []
    while len(nodes) > 0:
        start = nodes.pop()
        ordered = depth_first_order(graph, i_start=start, return_predecessors=False, directed=False).astype(np.int64)
        out.append(ordered)
        nodes.difference_update(ordered)
    return out


def _split_traversal(traversal, edges_hash):
    trav_edge = np.column_stack((traversal[:-1], traversal[1:]))
    contained = np.isin(_edge_hash(np.sort(trav_edge, axis=1)), edges_hash)
    if contained.all():
        split = [traversal]
    else:
        split, i, n = [], 0, len(contained)
        while i < n:                                    # contiguous runs of existing edges
            if not contained[i]:
                i += 1
                continue
            j = i
            while j < n and contained[j]:
                j += 1
            split.append(np.append(trav_edge[i:j, 0], trav_edge[j - 1, 1]))
            i = j
    for i, t in enumerate(split):
        t = np.asarray(t, dtype=np.int64)
        split[i] = t
        if len(t) <= 2:
            continue
        lo, hi = min(t[0], t[-1]), max(t[0], t[-1])
        if lo == hi:
            continue
        if np.isin(lo * np.int64(1 << 32) + hi, edges_hash):
            split[i] = np.append(t, t[0]).astype(np.int64)
    return split


def edges_to_path(edges: np.ndarray, vertices: np.ndarray):
    """Returns the entity list (each an int64 vertex-index sequence, = ``Line.points``)."""
    if len(edges) == 0:
        return []
    edges = np.asarray(edges, dtype=np.int64).copy()
    edges.sort(axis=1)
    edges_hash = _edge_hash(edges)
    splits = []
    for nodes in _traversals_dfs(edges):
        if len(nodes) < 2:
            continue
        splits.extend(_split_traversal(nodes, edges_hash))
    if splits:
        inc = np.vstack([np.column_stack((s[:-1], s[1:])) for s in splits])
        inc.sort(axis=1)
        missing = ~np.isin(edges_hash, _edge_hash(inc))
        splits.extend([e for e in edges[missing]])
    else:
        splits = [e for e in edges]
    return [np.asarray(s, dtype=np.int64) for s in splits]


def lines_to_path(segments_2d: np.ndarray, keys=None, merge: str = "hash", version: str = "4"):
    """(n,2,2) segments -> (vertices, entities, info).

    merge="hash": trimesh's own rule — endpoints merge when their 1e-8-rounded coordinates hash
    equal; vertex id = rank of the hash in ``np.unique`` order; kept coordinate = first
    occurrence in ``lines`` order.
    merge="topo": the GPU's rule — endpoints merge when they lie on the same mesh edge /
    vertex (``keys``); vertex id = rank of the kept coordinate's hash.  ``info['agree']`` says
    whether the two rules give the same graph on this plane."""
    segments_2d = np.asarray(segments_2d, dtype=np.float64)
    pts = segments_2d.reshape(-1, segments_2d.shape[-1])           # 2 columns (section_multiplane) or 3 (Trimesh.section)
    h, packed = hashable_rows(pts, version)
    _, uidx, inv = np.unique(h, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    info = {"packed": packed, "agree": True, "n_hash_nodes": len(uidx)}
    if merge == "hash" and keys is None:
        return pts[uidx], edges_to_path(inv.reshape(-1, 2), pts[uidx]), info
    kflat = np.asarray(keys, dtype=np.int64).reshape(-1, 2)
    _, tfirst, tinv = np.unique(_edge_hash(kflat), return_index=True, return_inverse=True)
    tinv = tinv.reshape(-1)
    # the two partitions of the endpoints agree iff they induce the same equivalence classes
    info["agree"] = bool(len(tfirst) == len(uidx) and (uidx[inv] == tfirst[tinv]).all())
    info["n_topo_nodes"] = len(tfirst)
    if merge == "hash":
        return pts[uidx], edges_to_path(inv.reshape(-1, 2), pts[uidx]), info
    kept = pts[tfirst]                                   # first occurrence in lines order
    k1, k2 = rank_key(kept, packed, version)
    order = np.lexsort((k2, k1))
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(len(order))
    if len(order) > 1:
        s1, s2 = k1[order], k2[order]
        info["rank_ties"] = int(((s1[1:] == s1[:-1]) & (s2[1:] == s2[:-1])).sum())
    verts = kept[order]
    return verts, edges_to_path(rank[tinv].reshape(-1, 2), verts), info


# --------------------------------------------------------------------------------------
# path.Path.__init__(process=True): merge_vertices / remove_duplicate_entities / remove_unreferenced_vertices
# --------------------------------------------------------------------------------------
def decimal_to_digits(decimal: float, min_digits=None) -> int:
    """trimesh ``util.decimal_to_digits``: ``abs(int(log10(decimal)))`` (int() truncates towards zero)."""
    digits = abs(int(np.log10(decimal)))
    if min_digits is not None:
        digits = int(np.clip(digits, min_digits, 20))
    return int(digits)


def path_scale(vertices: np.ndarray) -> float:
    """trimesh ``Path.scale``: length of the diagonal of the vertices' bounding box."""
    return float((np.ptp(vertices, axis=0) ** 2).sum() ** 0.5)


def merge_runs(idx: np.ndarray) -> np.ndarray:
    """trimesh ``grouping.merge_runs`` on an index sequence: consecutive repeats collapse to one."""
    idx = np.asarray(idx, dtype=np.int64)
    if len(idx) == 0:
        return idx
    keep = np.ones(len(idx), dtype=bool)
    keep[1:] = idx[1:] != idx[:-1]
    return idx[keep]


def process_path(vertices: np.ndarray, entities, version: str = "4"):
    """What ``load_path`` does to the output of ``lines_to_path`` by constructing ``Path2D(..., process=True)``:

    ``merge_vertices``: vertices whose coordinates round equal at ``digits = decimal_to_digits(tol.merge * scale,
    min_digits=1)`` (6 digits for a section 10..100 mm across, i.e. closer than ~1e-6 mm) become one vertex — the
    first in vertex order is kept, the vertex array is re-ordered by the hash of the rounded rows — and runs of
    repeated indices inside an entity collapse; a 3-point loop a-b-a becomes the line a-b, entities with fewer
    than two points are dropped.  ``remove_duplicate_entities`` and ``remove_unreferenced_vertices`` follow.
    Returns (vertices, entities, n_merged).  On the reference's test bones n_merged == 0 on every plane
    (tests/test_oracle_process.py), i.e. the step renumbers vertices and changes nothing the consumers read.
    """
    vertices = np.asarray(vertices, dtype=np.float64)
    if len(vertices) == 0:
        return vertices, list(entities), 0
    digits = decimal_to_digits(TOL_MERGE * path_scale(vertices), min_digits=1)
    q = np.round(vertices * 10 ** digits - 1e-6).astype(np.int64)             # grouping.float_to_int(data, digits)
    h, _ = _hash_int_rows(q, version)
    _, unique, inverse = np.unique(h, return_index=True, return_inverse=True)
    inverse = inverse.reshape(-1)
    n_merged = len(vertices) - len(unique)
    out = []
    for e in entities:
        pts = merge_runs(inverse[np.asarray(e, dtype=np.int64)])
        if len(pts) == 3 and pts[0] == pts[-1]:
            pts = pts[:2]
        elif len(pts) < 2:
            continue
        out.append(pts)
    new_vertices = vertices[unique]
    # remove_duplicate_entities: entities with identical point lists (either direction for closed loops hash
    # differently in trimesh, so only exact repeats are dropped)
    seen, dedup = set(), []
    for pts in out:
        key = pts.tobytes()
        if key in seen:
            continue
        seen.add(key)
        dedup.append(pts)
    # remove_unreferenced_vertices
    if dedup:
        ref = np.unique(np.concatenate(dedup))
        remap = -np.ones(len(new_vertices), dtype=np.int64)
        remap[ref] = np.arange(len(ref))
        dedup = [remap[pts] for pts in dedup]
        new_vertices = new_vertices[ref]
    return new_vertices, dedup, n_merged


def _hash_int_rows(q: np.ndarray, version: str = "4"):
    """``hashable_rows`` on already-integer rows (see :func:`hashable_rows`).  Rows of D columns are packed into one
    64-bit word when every value fits ``64 // D`` bits (32 for the 2-D paths of ``section_multiplane``; 21 for the 3-D
    paths of ``Trimesh.section``, i.e. |coordinate| < 0.0105 mm — never on a bone, so 3-D rows are always void rows)."""
    precision = 64 // q.shape[1]
    threshold = 2 ** (precision - 1)
    if version == "4":
        if q.max() < threshold and q.min() > -threshold:
            bang = (q.T + (threshold + 1)).astype(np.uint64)
            h = np.zeros(len(q), dtype=np.uint64)
            for offset, col in enumerate(bang):
                np.bitwise_xor(h, col << np.uint64(offset * precision), out=h)
            return h, True
    else:
        if np.abs(q).max() < threshold:
            h = np.zeros(len(q), dtype=np.int64)
            for offset, col in enumerate(q.T):
                np.bitwise_xor(h, col << (offset * precision), out=h)
            return h, True
    void = np.ascontiguousarray(q).view(np.dtype((np.void, q.dtype.itemsize * q.shape[1]))).reshape(-1)
    return void, False


# --------------------------------------------------------------------------------------
# shapely / GEOS validity of a ring (path.polygons.paths_to_polygons keeps a polygon only if ``is_valid``)
# --------------------------------------------------------------------------------------
def ring_is_valid(xy: np.ndarray) -> bool:
    """GEOS ``IsValidOp`` for a polygon that has only a shell: the closed ring must not cross or touch itself.
    Every pair of non-adjacent edges is tested with orientation signs (O(m^2), vectorised); adjacent edges may
    only share their common vertex (a collinear fold-back is a self-intersection)."""
    xy = np.asarray(xy, dtype=np.float64)
    if len(xy) < 4 or not np.array_equal(xy[0], xy[-1]) or not np.isfinite(xy).all():
        return False
    p, q = xy[:-1], xy[1:]
    m = len(p)
    if (np.abs(q - p).sum(axis=1) == 0).all():
        return False

    def orient(a, b, c):
        return np.sign((b[..., 0] - a[..., 0]) * (c[..., 1] - a[..., 1]) - (b[..., 1] - a[..., 1]) * (c[..., 0] - a[..., 0]))

    i, j = np.triu_indices(m, k=2)
    keep = ~((i == 0) & (j == m - 1))                      # first and last edge are adjacent through the closing vertex
    i, j = i[keep], j[keep]
    a, b, c, d = p[i], q[i], p[j], q[j]
    o1, o2, o3, o4 = orient(a, b, c), orient(a, b, d), orient(c, d, a), orient(c, d, b)
    proper = (o1 * o2 < 0) & (o3 * o4 < 0)

    def on_seg(a, b, c, o):                                 # c collinear with a-b and inside its box
        return (o == 0) & (np.minimum(a[:, 0], b[:, 0]) <= c[:, 0]) & (c[:, 0] <= np.maximum(a[:, 0], b[:, 0])) & \
               (np.minimum(a[:, 1], b[:, 1]) <= c[:, 1]) & (c[:, 1] <= np.maximum(a[:, 1], b[:, 1]))
    touch = on_seg(a, b, c, o1) | on_seg(a, b, d, o2) | on_seg(c, d, a, o3) | on_seg(c, d, b, o4)
    if (proper | touch).any():
        return False
    # adjacent edges: fold-back along the same line
    nxt = np.r_[np.arange(1, m), 0]
    u, w = q - p, q[nxt] - p[nxt]
    cross = u[:, 0] * w[:, 1] - u[:, 1] * w[:, 0]
    dot = u[:, 0] * w[:, 0] + u[:, 1] * w[:, 1]
    return not bool(((cross == 0) & (dot < 0)).any())


# --------------------------------------------------------------------------------------
# path.Path2D — the attributes the reference consumes
# --------------------------------------------------------------------------------------
def ring_area_signed(xy: np.ndarray) -> float:
    """GEOS ``algorithm::Area::ofRingSigned`` (what ``shapely.Polygon.area`` evaluates): shoelace
    about the first x, summed in ring order.  GEOS' sign is positive for clockwise; negated
    here so that CCW is positive."""
    n = len(xy)
    if n < 3:
        return 0.0
    x = xy[1:-1, 0] - xy[0, 0]
    return float(-np.sum(x * (xy[:-2, 1] - xy[2:, 1])) / 2.0)


def is_ccw(xy: np.ndarray) -> bool:
    """trimesh ``path.util.is_ccw``: sum(y_i*x_{i+1} - x_i*y_{i+1})/2 < 0."""
    prod = xy[:-1, 1] * xy[1:, 0] - xy[:-1, 0] * xy[1:, 1]
    return bool(prod.sum() / 2.0 < 0.0)


def _point_in_ring(p, ring) -> bool:
    x, y = p
    x0, y0 = ring[:-1, 0], ring[:-1, 1]
    x1, y1 = ring[1:, 0], ring[1:, 1]
    cond = (y0 > y) != (y1 > y)
    with np.errstate(divide="ignore", invalid="ignore"):
        xint = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
    return bool((cond & (x < xint)).sum() % 2 == 1)


class OraclePolygon:
    """The two shapely attributes ``slice.py:55-57`` reads (``area``) plus the ring itself."""

    def __init__(self, ring: np.ndarray):
        self.ring = ring
        self.area = abs(ring_area_signed(ring))
        self.valid = True

    @property
    def exterior_coords(self):
        return self.ring


class OraclePath2D:
    """What ``load_path(segments)`` hands back, reduced to the attributes the reference reads:
    ``entities`` (``slice.py:53,70``), ``discrete`` (``:72,76``), ``polygons_closed`` (``:55-57,73``;
    ``epicondyle.py:36,43``), ``area`` (``:59``), ``centroid`` (``:38``; ``canal.py:46``), ``bounds``,
    ``vertices`` (``mesh.py:102``) and ``metadata['face_index']``."""

    def __init__(self, vertices, entities, metadata=None, info=None, validate=True):
        self.vertices = vertices
        self.entities = entities
        self.metadata = metadata or {}
        self.info = info or {}
        self.validate = validate

    def entity_closed(self, i) -> bool:
        e = self.entities[i]
        return len(e) > 2 and e[0] == e[-1]

    @property
    def paths(self):
        # closed entities, in entity order, each a single-entity path.  Open entities would go
        # through networkx cycle_basis in trimesh; watertight input never produces them and the
        # oracle does not reconstruct that branch (they still count in ``len(entities)``).
        return [[i] for i in range(len(self.entities)) if self.entity_closed(i)]

    @property
    def discrete(self):
        out = []
        for (i,) in self.paths:
            d = self.vertices[self.entities[i]]
            if not is_ccw(d):
                d = np.ascontiguousarray(d[::-1])
            out.append(d)
        return out

    @property
    def polygons_closed(self):
        """trimesh ``paths_to_polygons``: a polygon per closed path with >= 4 points.  trimesh keeps it as is when
        shapely calls it valid and otherwise hands it to ``repair_invalid`` (GEOS ``buffer`` tricks), which cannot be
        restated without GEOS: such a ring keeps its un-repaired polygon here (``valid = False``; shapely's ``area``
        of an invalid polygon is still the shoelace sum) and is listed in ``info['invalid_rings']``.  It happens on
        the reference's own test bones — CT surfaces self-intersect a little (e.g. humerus_left, DistalSlices
        plane 15 at 30 planes: two edges three apart cross) — and moves the area by the size of the tiny loop."""
        cached = getattr(self, "_polys", None)
        if cached is not None:
            return cached
        out = []
        for k, d in enumerate(self.discrete):
            if len(d) < 4:
                out.append(None)
                continue
            poly = OraclePolygon(d)
            if self.validate and not ring_is_valid(d):
                poly.valid = False
                self.info.setdefault("invalid_rings", []).append(k)
            out.append(poly)
        self._polys = out
        return out

    @property
    def bounds(self):
        pts = np.array([[self.vertices[e].min(axis=0), self.vertices[e].max(axis=0)] for e in self.entities])
        pts = pts.reshape(-1, 2)
        return np.array([pts.min(axis=0), pts.max(axis=0)])

    @property
    def centroid(self):
        return self.bounds.mean(axis=0)

    @property
    def area(self) -> float:
        polys = [p for p in self.polygons_closed if p is not None]
        depth = [sum(_point_in_ring(b.ring[0], a.ring) for j, a in enumerate(polys) if j != i)
                 for i, b in enumerate(polys)]
        total = 0.0
        for i, p in enumerate(polys):
            if depth[i] % 2:
                continue
            total += p.area
            for j, c in enumerate(polys):
                if depth[j] == depth[i] + 1 and _point_in_ring(c.ring[0], p.ring):
                    total -= c.area
        return total


class OraclePath3D:
    """What ``Trimesh.section`` hands back (reference call sites mesh.py:95-99,158-161, surgical_neck.py:37-50,
    anatomic_neck.py:160-165, arthroplasty.py:71): the 3-D segments of ``mesh_plane`` through ``load_path`` — row hashes
    over THREE columns (always void rows: memcmp order), the same traversal, and ``discrete`` WITHOUT the counter-clockwise
    normalisation of 2-D paths: a closed polyline runs from its start node towards the neighbour with the lower id."""

    def __init__(self, vertices, entities, metadata=None, info=None):
        self.vertices, self.entities, self.metadata, self.info = vertices, entities, metadata or {}, info or {}

    @property
    def discrete(self):
        return [self.vertices[e] for e in self.entities]

    @property
    def bounds(self):
        return np.array([self.vertices.min(axis=0), self.vertices.max(axis=0)])


def section(vertices, faces, plane_normal, plane_origin, version="4", process=True):
    """trimesh ``Trimesh.section(plane_normal, plane_origin)``; ``None`` when the plane misses the mesh."""
    vertices = np.asarray(vertices, dtype=np.float64)
    n = np.asarray(plane_normal, dtype=np.float64).reshape(3)
    lines, fidx, keys, klass = mesh_plane(vertices, faces, np.asarray(plane_origin, dtype=np.float64).reshape(3), n)
    if len(lines) == 0:
        return None
    verts, ents, info = lines_to_path(lines, None, merge="hash", version=version)
    if process:
        verts, ents, info["n_merged"] = process_path(verts, ents, version)
    return OraclePath3D(verts, ents, metadata={"face_index": fidx}, info=info)


def section_multiplane(vertices, faces, plane_origin, plane_normal, heights, merge="hash", version="4",
                       check_merge=True, process=True, validate=True):
    """trimesh ``Trimesh.section_multiplane``: list of P paths, ``None`` where a plane misses.
    ``process``: apply ``Path.__init__``'s processing (:func:`process_path`); ``validate``: shapely's ``is_valid``
    gate of ``polygons_closed``.  ``check_merge=False`` skips the oracle-only comparison of the two merge rules
    (what the timed CPU arm uses: the reference does no such work)."""
    segs, transforms, fidx, keys, klass = mesh_multiplane(vertices, faces, plane_origin, plane_normal, heights)
    paths = [None] * len(segs)
    for i in range(len(segs)):
        if len(segs[i]) == 0:
            continue
        kk = keys[i] if (check_merge or merge == "topo") else None
        verts, ents, info = lines_to_path(segs[i], kk, merge=merge, version=version)
        if process:
            verts, ents, info["n_merged"] = process_path(verts, ents, version)
        paths[i] = OraclePath2D(verts, ents, info=info, validate=validate, metadata={
            "to_3D": transforms[i], "face_index": fidx[i], "segments": segs[i], "keys": keys[i], "klass": klass[i]})
    return paths
