"""Edge cases, batches, the Python Slices mirror and the full-size property checks (needs a B200)."""
import os

import numpy as np
import pytest

import oracle
from oracle.landmarks import axis_angle_deg, canal_axis_obb
from shoulder_b200 import _lib, meshio
from shoulder_b200.slice import GpuDistalSlices, GpuFullSlices, GpuProximalSlices, run_batch

from helpers import compare_sweep, rel_err, run_gpu

pytestmark = pytest.mark.gpu


def _cube():
    v = np.array([[x, y, z] for x in (0.0, 1.0) for y in (0.0, 1.0) for z in (0.0, 1.0)])
    f = np.array([[0, 2, 3], [0, 3, 1], [4, 5, 7], [4, 7, 6], [0, 1, 5], [0, 5, 4], [2, 6, 7], [2, 7, 3],
                  [0, 4, 6], [0, 6, 2], [1, 3, 7], [1, 7, 5]])
    return v, f


def test_cube_planes_through_faces_and_vertices(gpu_backend):
    """sign == 0 cases: coplanar bottom face (code 16 kept, code 6 dropped), coplanar top face (no section),
    planes that miss the mesh (None), plus the tilted cube whose planes pass exactly through vertices (code 8)."""
    v, f = _cube()
    zs = np.array([0.5, 0.0, 1.0, 2.0, -1.0, 0.25])
    rep = compare_sweep(v, f, zs, 16, expect_all_closed=False)
    assert rep["segments"] == 8 + 4 + 8
    d = np.array([1.0, 1.0, 1.0]) / np.sqrt(3)
    x = np.cross(d, [0, 0, 1.0]); x /= np.linalg.norm(x)
    r = np.stack([x, np.cross(d, x), d])
    vr = (v - 0.5) @ r.T
    lv = np.sort(np.unique(np.round(vr[:, 2], 12)))
    zs = np.array([lv[1], lv[2], 0.0, 0.5 * (lv[0] + lv[1])])
    rep = compare_sweep(vr, f, zs, 24, expect_all_closed=False)
    klass = np.concatenate([p.metadata["klass"] for p in rep["oracle"].paths if p is not None])
    assert (klass == oracle.trimesh_path.CLASS_VERTEX).any()


def test_open_mesh_is_flagged_and_closed_contours_survive(gpu_backend):
    """A sphere with a hole plus an intact sphere beside it: planes through the hole carry an open chain (flagged,
    counted as an entity, no contour) AND the closed contour of the intact sphere, which must equal the oracle's."""
    v, f = meshio.icosphere(2, 10.0)
    keep = np.ones(len(f), dtype=bool)
    keep[np.argsort(v[f].mean(axis=1)[:, 0])[-40:]] = False          # cut a hole at +x
    v2, f2 = meshio.icosphere(2, 7.0)
    vv = np.vstack([v, v2 + np.array([-30.0, 1.0, 0.5])])
    ff = np.vstack([f[keep], f2 + len(v)])
    zs = np.linspace(8.0, -8.0, 9)
    rep = compare_sweep(vv, ff, zs, 16, expect_all_closed=False)
    orc = rep["oracle"]
    res = run_gpu(vv, ff, zs, 16)
    status, n_ent = res.array(_lib.ARR_STATUS), res.array(_lib.ARR_N_ENT)
    n_open_planes = 0
    for i, p in enumerate(orc.paths):
        closed = all(p.entity_closed(k) for k in range(len(p.entities)))
        assert bool(status[i] & _lib.ST_OPEN) == (not closed), i
        n_open_planes += not closed
    assert 0 < n_open_planes < len(zs)
    # every plane still has the closed contour of the intact sphere -> the array API works on all planes
    from shoulder_b200.slice import GpuFullSlices

    class Obb:
        mesh = meshio.Mesh(vv, ff)
    g = GpuFullSlices(Obb(), zslice_num=9, interp_num=16)
    o = oracle.OracleSlices(vv, ff, g._zs, 16)
    assert rel_err(g._ixy, o.ixy) < 1e-9
    assert rel_err(g._areas1, np.array([max(q.area for q in p.polygons_closed) for p in o.paths])) < 1e-12


def test_batch_of_bones_and_mixed_sweeps(gpu_backend, bone_obbs):
    """config-4 shape: several bones, the three default sweeps each (200x100, 200x500, 600x512 scaled down), one call."""
    names = ["humerus_left", "humerus_left_trab", "humerus_right"]
    meshes, sweeps, specs = [], [], []
    for k, name in enumerate(names):
        m = meshio.synthetic_bone(bone_obbs(name).mesh, 100 + k)
        m = meshio.PcaObb(m).mesh
        z = m.vertices[:, 2]
        meshes.append((m.vertices, m.faces))
        for zs, n in ((np.linspace(0.99 * z.max(), 0.99 * z.min(), 50), 100), (np.linspace(0.99 * z.min(), 0.0, 40), 500),
                      (np.linspace(0.99 * z.max(), 0.5 * z.max(), 75), 512)):
            sweeps.append((k, float(zs.mean()), zs - zs.mean(), n))
            specs.append((m, zs, n))
    mask = _lib.OUT_PLANE | _lib.OUT_SEGMENTS | _lib.OUT_CONTOURS | _lib.OUT_ALL_PROFILES
    res = _lib.sweep_batch(meshes, sweeps, mask)
    for s, (m, zs, n) in enumerate(specs):
        rep = compare_sweep(m.vertices, m.faces, zs, n, res=res, sweep=s)
        assert rep["planes"] == len(zs)


def test_python_slices_mirror_and_canal_axis(gpu_backend, bone_obbs):
    class Neck:
        neck_z = 60.0
    obb = bone_obbs("humerus_left")
    full, dist, prox = GpuFullSlices(obb), GpuDistalSlices(obb), GpuProximalSlices(obb, Neck(), zslice_num=120, interp_num=512)
    run_batch([full, dist, prox])
    m = obb.mesh
    for g, zs in ((full, full._zs), (dist, dist._zs), (prox, prox._zs)):
        o = oracle.OracleSlices(m.vertices, m.faces, zs, g._interp_num)
        assert np.array_equal(g._centroids, o.centroids)
        assert rel_err(g._areas1, o.areas1) < 1e-12
        for name in ("ixy", "ixy_centered", "itr", "itr_start", "itr_centered", "itr_centered_start"):
            assert rel_err(getattr(g, "_" + name), getattr(o, name)) < 1e-9, name
        assert np.array_equal(g.itr((0.2, 0.8)), g.ixy((0.2, 0.8)))                      # slice.py:99-100
        assert np.array_equal(g.itr_start_even_theta((0.2, 0.8)), g.itr_start((0.2, 0.8)))  # slice.py:121-122
        p = g.slices((0.35, 0.75))[3]
        q = o.window(o.paths, (0.35, 0.75))[3]
        assert np.array_equal(p.centroid, q.centroid) and len(p.entities) == len(q.entities)
        assert np.array_equal(p.discrete[0], q.discrete[0])
        assert abs(p.area - q.area) <= 1e-12 * q.area
        assert np.array_equal(p.metadata["face_index"], q.metadata["face_index"])
    # final landmark parity that is reachable without the missing models: canal axis (canal.py:40-85), angle budget 0.01 deg
    o = oracle.OracleSlices(m.vertices, m.faces, full._zs, 100)
    ax_g = canal_axis_obb(full._centroids, full._zs, obb.z_length)
    ax_o = canal_axis_obb(o.centroids, full._zs, obb.z_length)
    assert axis_angle_deg(ax_g, ax_o) < 0.01 and rel_err(ax_g, ax_o) < 1e-9


def test_oversized_planes_take_the_global_workspace_path(gpu_backend, bone_obbs):
    m = bone_obbs("humerus_right").mesh
    zs = np.linspace(0.99 * m.vertices[:, 2].max(), 0.99 * m.vertices[:, 2].min(), 60)
    os.environ["SHB_DEBUG_SMEM_CAP"] = "96"            # most planes have > 96 segments
    try:
        rep = compare_sweep(m.vertices, m.faces, zs, 100, n_angles=36)
    finally:
        del os.environ["SHB_DEBUG_SMEM_CAP"]
    assert rep["contours"] >= 60


def test_half_million_triangles_properties(gpu_backend, bone_obbs):
    """BASELINE config 3 scale (Loop-subdivided humerus, 519,040 triangles, 8,192 planes): size-independent
    properties at full size + the oracle on a sample of planes."""
    base = bone_obbs("humerus_left").mesh
    v, f = meshio.loop_subdivide(base.vertices, base.faces, 2)
    assert len(f) == 519040
    z = v[:, 2]
    zs = np.linspace(0.99 * z.max(), 0.99 * z.min(), 8192)
    res = run_gpu(v, f, zs, 360, _lib.OUT_PLANE | _lib.OUT_SEGMENTS | _lib.OUT_CONTOURS | _lib.OUT_IXY | _lib.OUT_ITR_START)
    status, n_ent, n_seg = res.array(_lib.ARR_STATUS), res.array(_lib.ARR_N_ENT), res.array(_lib.ARR_N_SEG)
    assert not (status & (_lib.ST_EMPTY | _lib.ST_OPEN | _lib.ST_NONMANIFOLD)).any()       # watertight in, closed out
    ct_off, ctpt, pts = res.array(_lib.ARR_CONTOUR_OFF), res.array(_lib.ARR_CONTOUR_PT_OFF), res.array(_lib.ARR_POINTS)
    # every contour closed; points per plane == segments + contours (each node once + closing duplicates)
    first, last = pts[ctpt[:-1]], pts[ctpt[1:] - 1]
    assert np.array_equal(first, last)
    # ... except where Path.merge_vertices fused nodes closer than ~1e-6 mm (flagged SHB_ST_MERGED; a few planes in a thousand
    # at this mesh density): those deliver fewer points
    per_plane = np.add.reduceat(np.diff(ctpt).reshape(-1), ct_off[:-1].astype(np.int64))
    merged = (status & _lib.ST_MERGED) != 0
    assert np.array_equal(per_plane[~merged], (n_seg + n_ent)[~merged])
    assert (per_plane[merged] < (n_seg + n_ent)[merged]).all() and merged.sum() < 0.01 * len(zs)
    # face_index ascending within class blocks => strictly increasing except at <= 2 class boundaries per plane
    off, fi = res.array(_lib.ARR_SEG_OFF), res.array(_lib.ARR_FACE_INDEX)
    drops = np.add.reduceat((np.diff(fi) <= 0).astype(np.int64), off[:-1][:-0 or None])[: len(zs)]
    interior = np.ones(len(fi) - 1, dtype=bool)
    interior[off[1:-1] - 1] = False                         # ignore plane boundaries
    assert ((np.diff(fi) <= 0) & interior).sum() == 0       # real bones never touch a vertex: one class per plane
    # ixy rows are closed polylines; itr_start rows start at their minimum theta
    ixy, itr = res.array(_lib.ARR_IXY), res.array(_lib.ARR_ITR_START)
    assert np.array_equal(ixy[:, :, 0], ixy[:, :, -1])
    assert np.array_equal(itr[:, 0, 0], itr[:, 0, :].min(axis=1))
    # area1 bounded by the bounding box, areas vary smoothly along the shaft
    b = res.array(_lib.ARR_BOUNDS)
    box = (b[:, 1, 0] - b[:, 0, 0]) * (b[:, 1, 1] - b[:, 0, 1])
    a1 = res.array(_lib.ARR_AREA1)
    assert (a1 > 0).all() and (a1 <= box * (1 + 1e-12)).all()
    # oracle on 24 sampled planes of the same sweep (same z_orig => identical planes)
    idx = np.linspace(5, 8186, 24).astype(int)
    z_orig = zs.mean()
    paths = oracle.section_multiplane(v, f, [0, 0, z_orig], [0, 0, 1], (zs - z_orig)[idx], merge="topo")
    for i, p in zip(idx, paths):
        assert np.array_equal(fi[off[i]:off[i + 1]], p.metadata["face_index"])
        c0 = int(ct_off[i])
        for k, dsc in enumerate(p.discrete):
            assert np.array_equal(pts[int(ctpt[c0 + k]):int(ctpt[c0 + k + 1])], dsc)
        assert np.array_equal(res.array(_lib.ARR_CENTROID)[i], p.centroid)


def test_float32_profile_outputs_stay_inside_the_north_star_budget(gpu_backend, bone_obbs):
    """SHB_OUT_F32: float32 stores; the cartesian arrays are the float64 values rounded once, the polar forms and the radius
    image are COMPUTED in float32 (atan2f / sqrtf / float quotient) with the theta-min roll still decided in float64.
    Tolerance = north_star's 1e-5 relative; the roll (a discrete decision) must be the float64 one on every row."""
    for name in ("humerus_left", "humerus_left_trab"):
        m = bone_obbs(name).mesh
        zs = np.linspace(0.99 * m.vertices[:, 2].max(), 0.99 * m.vertices[:, 2].min(), 300)
        mask = _lib.OUT_PLANE | _lib.OUT_ALL_PROFILES | _lib.OUT_RADIAL
        r64 = run_gpu(m.vertices, m.faces, zs, 360, mask, 360)
        r32 = run_gpu(m.vertices, m.faces, zs, 360, mask | _lib.OUT_F32, 360)
        for w in (_lib.ARR_IXY, _lib.ARR_IXY_CENTERED):
            a, b = r32.array(w), r64.array(w)
            assert a.dtype == np.float32 and b.dtype == np.float64 and a.shape == b.shape
            assert np.array_equal(a, b.astype(np.float32)), w          # identical values, rounded once
        for w in (_lib.ARR_ITR, _lib.ARR_ITR_START, _lib.ARR_ITR_CENTERED, _lib.ARR_ITR_CENTERED_START, _lib.ARR_RADIAL):
            a, b = r32.array(w), r64.array(w)
            assert a.dtype == np.float32 and a.shape == b.shape
            assert rel_err(a, b) < 1e-5, (w, rel_err(a, b))
        for w in (_lib.ARR_ITR_START, _lib.ARR_ITR_CENTERED_START):     # same roll: the radius rows line up sample by sample
            a, b = r32.array(w), r64.array(w)
            assert np.abs(a[:, 1, :] - b[:, 1, :]).max() < 1e-5 * np.abs(b[:, 1, :]).max()
            assert np.abs(a[:, 0, :] - b[:, 0, :]).max() < 1e-5
        assert np.array_equal(r32.array(_lib.ARR_CENTROID), r64.array(_lib.ARR_CENTROID))     # plane records stay f64


def test_many_and_nested_contours(gpu_backend):
    """A row of spheres (dozens of contours per plane, ordered by the rank of their start vertex) and a hollow
    ball (outer shell + inward-facing inner shell: nested loops, Path2D.area = shell minus hole)."""
    vs, fs, off = [], [], 0
    rng = np.random.default_rng(11)
    for k in range(30):
        v, f = meshio.icosphere(2, 1.0 + 0.02 * k)
        vs.append(v + np.array([3.1 * (k % 6) + rng.uniform(-0.2, 0.2), 3.3 * (k // 6) + rng.uniform(-0.2, 0.2), rng.uniform(-0.3, 0.3)]))
        fs.append(f + off); off += len(v)
    v, f = np.vstack(vs), np.vstack(fs)
    rep = compare_sweep(v, f, np.linspace(0.8, -0.8, 17), 64, expect_all_closed=True)
    assert rep["contours"] >= 17 * 25
    vo, fo = meshio.icosphere(3, 10.0)
    vi, fi = meshio.icosphere(3, 6.0)
    v = np.vstack([vo, vi + np.array([0.7, -0.4, 0.2])])
    f = np.vstack([fo, fi[:, ::-1] + len(vo)])                  # inner surface faces inward
    zs = np.linspace(9.0, -9.0, 25)
    rep = compare_sweep(v, f, zs, 90, n_angles=36)
    orc = rep["oracle"]
    from shoulder_b200.slice import GpuFullSlices

    class Obb:
        mesh = meshio.Mesh(v, f)
    g = GpuFullSlices(Obb(), zslice_num=25, interp_num=90)
    nested = 0
    for p, q in zip(g._slices, oracle.OracleSlices(v, f, g._zs, 90).paths):
        assert abs(p.area - q.area) <= 1e-12 * max(q.area, 1.0)
        nested += len(q.entities) == 2
    assert nested >= 10


def test_inconsistent_winding_uses_the_two_cycle_path(gpu_backend, bone_obbs):
    """Flipping the winding of random faces breaks the direction rule of the fast path; results must still
    equal the oracle's on the same (flipped) input."""
    m = bone_obbs("humerus_right").mesh
    f = m.faces.copy()
    flip = np.random.default_rng(5).random(len(f)) < 0.3
    f[flip] = f[flip][:, [0, 2, 1]]
    zs = np.linspace(0.99 * m.vertices[:, 2].max(), 0.99 * m.vertices[:, 2].min(), 80)
    compare_sweep(m.vertices, f, zs, 100)
    from shoulder_b200 import _lib as L
    fast = run_gpu(m.vertices, f, zs, 100, L.OUT_PLANE | L.OUT_CONTOURS | L.OUT_IXY)
    full = run_gpu(m.vertices, f, zs, 100, L.OUT_PLANE | L.OUT_SEGMENTS | L.OUT_CONTOURS | L.OUT_IXY)
    for w in (L.ARR_POINTS, L.ARR_CONTOUR_PT_OFF, L.ARR_IXY, L.ARR_CENTROID):
        assert np.array_equal(fast.array(w), full.array(w))


def test_duplicate_and_unsorted_heights(gpu_backend):
    v, f = meshio.icosphere(3, 1.0, scale=(20.0, 30.0, 170.0))
    zs = np.array([10.0, -50.0, 10.0, 120.0, 0.0, 0.0, 169.9, -169.9, 500.0])
    rep = compare_sweep(v, f, zs, 50, expect_all_closed=False)
    assert rep["contours"] == 8


def test_pipelined_host_call_equals_the_single_call(gpu_backend, bone_obbs):
    meshes, sweeps = [], []
    for k, name in enumerate(["humerus_left", "humerus_right", "humerus_left_trab", "humerus_left_flipped", "humerus_right"]):
        m = meshio.PcaObb(meshio.synthetic_bone(bone_obbs(name).mesh, 40 + k)).mesh
        z = m.vertices[:, 2]
        meshes.append((m.vertices, m.faces))
        for zs, n in ((np.linspace(0.99 * z.max(), 0.99 * z.min(), 64), 90), (np.linspace(0.99 * z.min(), 0.0, 33), 120)):
            sweeps.append((k, float(zs.mean()), zs - zs.mean(), n))
    mask = _lib.OUT_PLANE | _lib.OUT_IXY | _lib.OUT_ITR_START | _lib.OUT_ITR_CENTERED_START | _lib.OUT_RADIAL | _lib.OUT_CONTOURS
    packed = _lib._pack(meshes, sweeps)
    one = _lib.sweep_batch(None, None, mask, 45, packed=packed)
    chunks, first = _lib.split_packed(packed, 3)
    piped = _lib.sweep_batch_pipelined(chunks, first, mask, 45)
    assert piped.n_sweep == len(sweeps)
    for s in range(len(sweeps)):
        for w in (_lib.ARR_N_SEG, _lib.ARR_CENTROID, _lib.ARR_IXY, _lib.ARR_ITR_START, _lib.ARR_ITR_CENTERED_START,
                  _lib.ARR_RADIAL, _lib.ARR_POINTS, _lib.ARR_CONTOUR_PT_OFF):
            assert np.array_equal(one.array(w, s), piped.array(w, s)), (s, w)
        # multi-contour planes sum their areas with shared-memory atomics: reproducible to rounding, not to the bit
        assert np.allclose(one.array(_lib.ARR_AREA1, s), piped.array(_lib.ARR_AREA1, s), rtol=1e-13)


def test_invalid_inputs_fail_loudly(gpu_backend):
    v, f = meshio.icosphere(1, 1.0)
    bad = f.copy(); bad[3, 1] = len(v) + 5
    with pytest.raises(_lib.BackendError, match="face index out of range"):
        _lib.sweep_batch([(v, bad)], [(0, 0.0, np.linspace(-0.5, 0.5, 4), 8)], _lib.OUT_PLANE)
    with pytest.raises(_lib.BackendError):
        _lib.sweep_batch([(v, f)], [(0, 0.0, np.linspace(-0.5, 0.5, 4), 1)], _lib.OUT_PLANE)        # interp_num < 2
    with pytest.raises(_lib.BackendError):
        _lib.sweep_batch([(v, f)], [(2, 0.0, np.linspace(-0.5, 0.5, 4), 8)], _lib.OUT_PLANE)        # sweep names a missing mesh
    # the library is still usable afterwards
    res = _lib.sweep_batch([(v, f)], [(0, 0.0, np.linspace(-0.5, 0.5, 4), 8)], _lib.OUT_PLANE | _lib.OUT_IXY)
    assert (res.array(_lib.ARR_N_ENT) == 1).all()


def test_trim_releases_caches_and_the_library_keeps_working(gpu_backend):
    v, f = meshio.icosphere(3, 5.0)
    zs = np.linspace(4.0, -4.0, 32)
    a = run_gpu(v, f, zs, 64).array(_lib.ARR_IXY).copy()
    _lib.trim()
    b = run_gpu(v, f, zs, 64).array(_lib.ARR_IXY)
    assert np.array_equal(a, b)


def test_degenerate_batches(gpu_backend):
    """planes that all miss, a mesh without faces, a sweep without planes beside a normal one"""
    v, f = meshio.icosphere(2, 1.0)
    res = _lib.sweep_batch([(v, f)], [(0, 0.0, np.array([5.0, 6.0, -7.0]), 8)], _lib.OUT_PLANE | _lib.OUT_IXY | _lib.OUT_CONTOURS)
    assert (res.array(_lib.ARR_STATUS) == _lib.ST_EMPTY).all() and res.totals()["segments"] == 0
    assert np.isnan(res.array(_lib.ARR_IXY)).all() and res.array(_lib.ARR_POINTS).shape == (0, 2)
    empty = (np.zeros((0, 3)), np.zeros((0, 3), dtype=np.int64))
    res = _lib.sweep_batch([empty, (v, f)], [(0, 0.0, np.array([0.0, 0.1]), 8), (1, 0.0, np.zeros(0), 8), (1, 0.0, np.array([0.0, 0.3]), 16)],
                           _lib.OUT_PLANE | _lib.OUT_IXY)
    assert (res.array(_lib.ARR_STATUS, 0) == _lib.ST_EMPTY).all()
    assert res.array(_lib.ARR_N_ENT, 1).shape == (0,)
    assert (res.array(_lib.ARR_N_ENT, 2) == 1).all() and res.array(_lib.ARR_IXY, 2).shape == (2, 2, 16)
    res = _lib.sweep_batch([empty], [(0, 0.0, np.array([0.0]), 4)], _lib.OUT_PLANE)      # nothing to slice at all
    assert res.array(_lib.ARR_N_SEG)[0] == 0


def test_h4_coordinate_hash_vs_topological_merge_is_reported_not_hidden(gpu_backend):
    """two distinct crossing points closer than 1e-8 hash equal in trimesh (they fuse into a degree-4
    node), while the GPU merges on the mesh edge and keeps two contours.  The oracle reports such planes
    (info['agree'] == False); the comparison lists them instead of asserting connectivity on them."""
    v0 = np.array([[x, y, z] for x in (0.0, 1.0) for y in (0.0, 1.0) for z in (0.0, 1.0)])
    f0 = np.array([[0, 2, 3], [0, 3, 1], [4, 5, 7], [4, 7, 6], [0, 1, 5], [0, 5, 4], [2, 6, 7], [2, 7, 3],
                   [0, 4, 6], [0, 6, 2], [1, 3, 7], [1, 7, 5]])
    v = np.vstack([v0, v0 + np.array([1.0 + 2e-9, 0.0, 0.0])])        # second cube 2e-9 to the right of the first
    f = np.vstack([f0, f0 + len(v0)])
    zs = np.array([0.25, 0.5, 0.75])
    rep = compare_sweep(v, f, zs, 16, expect_all_closed=False)
    assert rep["h4_exceptions"] == [0, 1, 2]
    res = run_gpu(v, f, zs, 16)
    assert (res.array(_lib.ARR_N_ENT) == 2).all()                      # two unit squares, not one fused figure
    areas = res.array(_lib.ARR_CONTOUR_AREA)
    assert np.allclose(areas, 1.0, rtol=1e-12)
    assert (res.array(_lib.ARR_STATUS) & (_lib.ST_OPEN | _lib.ST_NONMANIFOLD) == 0).all()


def test_path_merge_vertices_behind_both_stitchers(gpu_backend):
    """Planes a hair above a mesh vertex: the edges leaving the vertex cross the plane within 1e-6 mm of each other, so
    trimesh's Path.__init__ -> merge_vertices fuses those contour nodes (the oracle restates it).  One solid = one contour
    per plane = the group stitcher; two solids = a second contour = declined to the CTA stitcher; with SHB_OUT_SEGMENTS
    every plane goes through the CTA stitcher.  All of them must deliver the merged contour and flag SHB_ST_MERGED."""
    v, f = meshio.icosphere(3, 1.0, scale=(20.0, 25.0, 30.0))
    rng = np.random.default_rng(5)
    zv = rng.choice(v[np.abs(v[:, 2]) < 24.0, 2], size=10, replace=False)
    zs = np.concatenate([zv + 3e-8, np.linspace(-20.0, 20.0, 7)])
    rep = compare_sweep(v, f, zs, 64, n_angles=36)
    assert len(rep["merged_planes"]) >= 3, rep["merged_planes"]
    v2 = np.vstack([v, v * 0.6 + np.array([70.0, 0.0, 1.0])])
    f2 = np.vstack([f, f + len(v)])
    rep2 = compare_sweep(v2, f2, zs, 64, n_angles=36, expect_all_closed=False)
    assert len(rep2["merged_planes"]) >= 3 and rep2["contours"] > len(zs)


def test_calls_from_another_host_thread(gpu_backend):
    """CUDA's current device is per host thread: an entry point called from a thread that never saw shb_init must still
    run on the library's device and streams (device guard at every entry point)."""
    import threading
    v, f = meshio.icosphere(3, 1.0, scale=(20.0, 30.0, 40.0))
    zs = np.linspace(-35.0, 35.0, 40)
    mask = _lib.OUT_PLANE | _lib.OUT_IXY | _lib.OUT_CONTOURS
    ref = run_gpu(v, f, zs, 64, mask)
    out = {}

    def work():
        try:
            r = run_gpu(v, f, zs, 64, mask)
            out["ixy"], out["pts"] = r.array(_lib.ARR_IXY).copy(), r.array(_lib.ARR_POINTS).copy()
            r.close()
        except Exception as e:          # surfaced in the main thread
            out["err"] = e

    ts = [threading.Thread(target=work) for _ in range(2)]
    ts[0].start(); ts[0].join()
    assert "err" not in out, out.get("err")
    assert np.array_equal(out["ixy"], ref.array(_lib.ARR_IXY)) and np.array_equal(out["pts"], ref.array(_lib.ARR_POINTS))
    ts[1].start()                        # a second thread while the main thread keeps calling
    again = run_gpu(v, f, zs, 64, mask)
    ts[1].join()
    assert "err" not in out and np.array_equal(again.array(_lib.ARR_IXY), ref.array(_lib.ARR_IXY))
