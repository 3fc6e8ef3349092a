"""Feature stage between slice.py's polar stacks and the landmark models (scope row f3), pinned to the reference's own
code: ``tests/golden/refgroove_*.npz`` were made by running ``DeepGroove.points()`` and ``AnatomicNeck.points()`` of the
reference unchanged (tests/golden/make_groove_vectors.py; trimesh answered by the oracle, onnxruntime by the ONNX reader).

  CPU   oracle/groove.py (restatement with scipy / numpy) == those vectors; the vectors == a live run (build container)
  GPU   shb_groove_features / shb_forest / shb_groove_points / shb_neck_image == those vectors (1e-5 budget of north_star;
        discrete results — which samples are peaks, bg_theta, which sample is the local minimum — exact)
"""
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle
from oracle import groove
from shoulder_b200 import meshio

import refload

HERE = Path(__file__).resolve().parent
GOLD = HERE / "golden"
sys.path.insert(0, str(GOLD))
import make_groove_vectors as mgv  # noqa: E402

NAMES = mgv.NAMES
needs_reference = pytest.mark.skipif(not refload.available(), reason="/root/reference is not on this host")


def _bone(name):
    g = np.load(GOLD / f"refgroove_{name}.npz")
    ct = meshio.load_mesh(GOLD / "bones" / f"{name}.npz")
    v, f = refload.exact_frame(ct.vertices, ct.faces, g["transform"])
    zs = np.linspace(0.99 * v[:, 2].max(), float(g["neck_z"]), 600)            # ProximalSlices._zs, slice.py:248-253
    return g, v, f, zs


def _sorted_rows(X, theta, row):
    """peaks of one stack row may come in any order (np.argpartition): sort by (row, theta)"""
    k = np.lexsort((theta, row))
    return X[k], theta[k], row[k]


@needs_reference
@pytest.mark.parametrize("name", NAMES[:1])
def test_committed_groove_vectors_equal_a_live_run_of_the_reference(name):
    g, v, f, zs = _bone(name)
    live, _ = mgv.run_reference(v, f, g["transform"])
    for key in ("X", "peak_theta", "proba", "groove_points_obb", "canal_axis"):
        assert np.array_equal(live[key], g[key]), key
    assert float(live["bg_theta"]) == float(g["bg_theta"])
    assert np.array_equal(live["image"][::mgv.IMAGE_STRIDE], g["image_rows"])


@pytest.mark.parametrize("name", NAMES)
def test_groove_oracle_equals_the_reference_made_vectors(name):
    g, v, f, zs = _bone(name)
    orc = oracle.OracleSlices(v, f, zs, 512)
    lo, hi = oracle.cutoff_window(600, (0.2, 0.75))
    pol = orc.itr_centered_start[lo:hi]
    ft = groove.groove_features(pol, zs[lo:hi], g["canal_axis"], 512)
    assert ft["X"].shape == g["X"].shape and np.abs(ft["X"] - g["X"]).max() < 1e-12
    assert np.array_equal(ft["peak_theta"], g["peak_theta"])
    bg, _ = groove.groove_theta(ft["peak_theta"], g["proba"][:, 1])
    assert bg == float(g["bg_theta"])
    pts, _ = groove.groove_points(pol, ft["polar_0"], zs[lo:hi], orc.centroids[lo:hi], bg, 512)
    assert np.array_equal(pts, g["groove_points_obb"])
    lo2, hi2 = oracle.cutoff_window(600, (0.0, 0.852))
    img, _, _ = groove.neck_image(orc.itr_start[lo2:hi2], bg)
    assert np.array_equal(img.astype(np.float32)[::mgv.IMAGE_STRIDE], g["image_rows"])
    assert abs(img.astype(np.float32).astype(np.float64).sum() - float(g["image_sum"])) < 1e-6


@needs_reference
def test_forest_reader_reproduces_the_reference_blob_semantics():
    """oracle/onnx_forest.py on the reference's rfc_bg3.onnx: 40 trees, binary, scores in [0, 1], leaves sum to the class-1
    fraction; the stored probabilities of the vectors come from it (no onnxruntime in the image to compare with)."""
    from oracle.onnx_forest import Forest
    fo = Forest(refload.REF / "humerus" / "models" / "rfc_bg3.onnx")
    assert fo.n_trees == 40 and fo.n_classes == 2 and fo.binary_single
    g = np.load(GOLD / "refgroove_humerus_left.npz")
    p = fo.predict_proba(g["X"])
    assert np.array_equal(p, g["proba"]) and (p >= 0).all() and np.allclose(p.sum(axis=1), 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_feature_stage_reproduces_the_reference_made_vectors(gpu_backend, name):
    from shoulder_b200 import _lib, features
    g, v, f, zs = _bone(name)
    lo, hi = oracle.cutoff_window(600, (0.2, 0.75))
    lo2, hi2 = oracle.cutoff_window(600, (0.0, 0.852))
    zo = float(np.mean(zs))
    req = [{_lib.OUT_ITR_START: (lo2, hi2), _lib.OUT_ITR_CENTERED_START: (lo, hi)}]
    res = _lib.sweep_batch([(v, f)], [(0, zo, zs - zo, 512)], _lib.OUT_PLANE, 0, requests=req, lazy=True)
    ft = features.groove_features(res, [0], [zs[lo:hi]], [g["canal_axis"]])[0]
    # which samples are peaks, per row: exact; features within north_star's 1e-5
    orc = oracle.OracleSlices(v, f, zs, 512)
    ref = groove.groove_features(orc.itr_centered_start[lo:hi], zs[lo:hi], g["canal_axis"], 512)
    assert np.array_equal(ft["n_peaks"], np.bincount(ref["peak_row"], minlength=hi - lo))
    Xg, tg, rg = _sorted_rows(ft["X"], ft["peak_theta"], ft["peak_row"])
    Xr, tr, rr = _sorted_rows(g["X"], g["peak_theta"], ref["peak_row"])
    assert np.array_equal(rg, rr) and np.allclose(tg, tr, rtol=0, atol=1e-12)       # same samples; theta is the device's atan2 (<= 1 ulp from libm)
    assert np.abs(Xg - Xr).max() < 1e-5 * max(1.0, np.abs(Xr).max())
    # forest on the device == the ONNX semantics on the host; density arg-max == the reference's bg_theta
    if (GOLD / "forest_rfc_bg3.npz").exists():
        forest = features.Forest.from_arrays(np.load(GOLD / "forest_rfc_bg3.npz"))
        k = np.lexsort((ft["peak_theta"], ft["peak_row"]))
        proba = forest.predict_proba(ft["X"][k])
        kr = np.lexsort((g["peak_theta"], ref["peak_row"]))
        assert np.abs(proba - g["proba"][kr]).max() < 1e-6
        bg = features.groove_theta(ft["peak_theta"][k], proba[:, 1])
        assert features.groove_theta_batch([ft["peak_theta"][k]], [proba[:, 1]])[0] == bg          # the device arg-max == the host one
    else:
        bg = features.groove_theta(g["peak_theta"], g["proba"][:, 1])
    assert abs(bg - float(g["bg_theta"])) < 1e-12
    bg = float(g["bg_theta"])
    pts, _ = features.groove_points(res, [0], [zs[lo:hi]], [bg], 512)[0]
    assert np.abs(pts - g["groove_points_obb"]).max() < 1e-9 * np.abs(g["groove_points_obb"]).max()
    img, (mn, mx), shft = features.neck_image(res, [0], [bg], want_shifted=True)[0]
    assert img.shape == tuple(g["image_shape"]) and img.dtype == np.float32
    assert np.abs(img[::mgv.IMAGE_STRIDE].astype(np.float64) - g["image_rows"]).max() < 1e-5
    assert abs(img.astype(np.float64).sum() - float(g["image_sum"])) < 1e-5 * float(g["image_sum"])
    _, shft_ref, _ = groove.neck_image(orc.itr_start[lo2:hi2], bg)
    assert np.abs(shft - shft_ref).max() < 1e-9 * np.abs(shft_ref).max()


@pytest.mark.gpu
def test_device_groove_angle_equals_the_host_density_arg_max(gpu_backend):
    """shb_groove_theta against features.groove_theta (numpy, what the reference's KernelDensity arg-max reduces to) on
    random peak sets of bone-like sizes, several bones per call, incl. an empty set and a set without accepted peaks."""
    from shoulder_b200 import features
    rng = np.random.default_rng(3)
    ths, prs = [], []
    for n in (2310, 1500, 0, 40, 900):
        c = rng.uniform(-np.pi, np.pi, 3)                                    # three clusters + background
        th = np.concatenate([rng.normal(c[k % 3], 0.15, n // 4) for k in range(3)] + [rng.uniform(-np.pi, np.pi, n - 3 * (n // 4))])
        ths.append(np.clip(th, -np.pi, np.pi)); prs.append(rng.uniform(0, 1, len(th)).astype(np.float32))
    ths.append(rng.uniform(-3, 3, 50)); prs.append(np.full(50, 0.1, dtype=np.float32))  # nothing accepted: density 0 everywhere -> first grid angle
    got = features.groove_theta_batch(ths, prs)
    ref = np.array([features.groove_theta(t, p) for t, p in zip(ths, prs)])
    # the device sums the density in peak order, numpy pairwise: where two neighbouring grid angles tie within rounding the
    # arg-max may land on the other one — then the host density at the device's angle must equal the maximum
    tlin = np.linspace(-np.pi, np.pi, 1024)
    for g_, r_, th, pr in zip(got, ref, ths, prs):
        if g_ == r_:
            continue
        pts = th[pr > 0.4]
        dens = np.maximum(0.0, 1.0 - np.abs(tlin[:, None] - pts[None, :])).sum(axis=1)
        k = int(np.argmin(np.abs(tlin - g_)))
        assert tlin[k] == g_ and dens[k] >= dens.max() * (1 - 1e-12), (g_, r_)
    assert (got == ref).sum() >= len(ref) - 1


@pytest.mark.gpu
def test_fused_front_end_equals_the_step_by_step_chain(gpu_backend):
    """shb_landmark_front (canal axis, features, StandardScaler, forest, groove angle, groove points, neck image in one
    enqueue, everything between the stages staying on the device) against the same chain driven call by call: the canal
    axis agrees with the host fit (numpy SVD, what scikit-spatial's Line.best_fit runs) to 1e-9 (Jacobi eigenvector against
    LAPACK), and with that axis handed to the step-by-step chain every later result is EQUAL bit for bit — including the
    StandardScaler, which the device sums in numpy's order."""
    from shoulder_b200 import _lib, features
    if not (GOLD / "forest_rfc_bg3.npz").exists():
        pytest.skip("forest fixture missing")
    forest = features.Forest.from_arrays(np.load(GOLD / "forest_rfc_bg3.npz"))
    meshes, sweeps, zs_full, zs_prox = [], [], [], []
    names = NAMES[:3]
    nb = len(names)
    for b, name in enumerate(names):
        g, v, f, zs = _bone(name)
        zf = np.linspace(0.99 * v[:, 2].max(), 0.99 * v[:, 2].min(), 200)        # FullSlices._zs
        meshes.append((v, f))
        sweeps += [(b, float(np.mean(zf)), zf - np.mean(zf), 100), (b, float(np.mean(zs)), zs - np.mean(zs), 512)]
        zs_full.append(zf); zs_prox.append(zs)
    lo, hi = oracle.cutoff_window(600, (0.2, 0.75))
    lo2, hi2 = oracle.cutoff_window(600, (0.0, 0.852))
    c_lo, c_hi = oracle.cutoff_window(200, (0.35, 0.75))
    req = []
    for b in range(nb):
        req += [{}, {_lib.OUT_ITR_START: (lo2, hi2), _lib.OUT_ITR_CENTERED_START: (lo, hi)}]
    res = _lib.sweep_batch(meshes, sweeps, _lib.OUT_PLANE, 0, requests=req, lazy=True)
    full, prox = [2 * b for b in range(nb)], [2 * b + 1 for b in range(nb)]
    half = np.array([0.55 * (abs(z[0]) + abs(z[-1])) / 2 for z in zs_full])
    zs_g = [z[lo:hi] for z in zs_prox]
    # ---- one call
    fe = features.LandmarkFrontEnd(forest, 512)
    out = fe(res, full, prox, (c_lo, c_hi), np.stack([z[c_lo:c_hi] for z in zs_full]), half, zs_g)
    out = {k: np.array(v) for k, v in out.items()}
    # ---- step by step
    cz = np.stack([np.c_[res.array(_lib.ARR_CENTROID, s)[c_lo:c_hi], zs_full[b][c_lo:c_hi]] for b, s in enumerate(full)])
    mid = cz.mean(axis=1, keepdims=True)
    dirn = np.linalg.svd(cz - mid)[2][:, 0, :]
    dirn = np.where(dirn[:, -1:] < 0, -dirn, dirn)
    axes = np.stack([mid[:, 0] + dirn * half[:, None], mid[:, 0] - dirn * half[:, None]], axis=1)
    assert np.abs(out["canal_axes"] - axes).max() < 1e-9 * np.abs(axes).max()
    ft = features.groove_features(res, prox, zs_g, out["canal_axes"])
    cuts = np.cumsum([0] + [len(f["X"]) for f in ft])
    proba = forest.predict_proba(np.vstack([f["X"] for f in ft]))
    bg = features.groove_theta_batch([f["peak_theta"] for f in ft], [proba[a:b, 1] for a, b in zip(cuts[:-1], cuts[1:])])
    pts = features.groove_points(res, prox, zs_g, bg, 512)
    imgs = features.neck_image(res, prox, bg, interp_num=512)
    for b in range(nb):
        r0, r1 = out["row_cuts"][b], out["row_cuts"][b + 1]
        cnt = out["n_peaks"][r0:r1]
        assert np.array_equal(cnt, ft[b]["n_peaks"])
        sel = np.arange(features.N_TOP)[None, :] < cnt[:, None]
        assert np.array_equal(out["peak_index"][r0:r1][sel], ft[b]["peak_index"])
        assert np.array_equal(out["peak_theta"][r0:r1][sel], ft[b]["peak_theta"])
        raw = out["feat"][r0:r1][sel]
        assert np.array_equal(raw, ft[b]["raw"])
        # StandardScaler on the device == numpy on the device's own raw rows, bit for bit
        mean, std = raw.mean(axis=0), raw.std(axis=0)
        std = np.where(std == 0.0, 1.0, std)
        assert np.array_equal(out["scaler"][b, 0], mean) and np.array_equal(out["scaler"][b, 1], std)
        assert np.array_equal(out["X"][r0:r1][sel], ((raw - mean) / std).astype(np.float32))
        assert np.abs(out["proba1"][r0:r1][sel] - proba[cuts[b]:cuts[b + 1], 1]).max() < 1e-6      # float atomics: the trees' order
        assert out["bg_theta"][b] == bg[b]
        assert np.array_equal(out["points"][r0:r1], pts[b][0]) and np.array_equal(out["local_theta"][r0:r1], pts[b][1])
        i0, i1 = out["image_cuts"][b], out["image_cuts"][b + 1]
        assert np.array_equal(out["image"][i0:i1], imgs[b][0]) and tuple(out["minmax"][b]) == imgs[b][1]
    # the no-wait form (copies on the library's copy stream) delivers the same bytes
    fe2 = features.LandmarkFrontEnd(forest, 512)
    out2 = fe2(res, full, prox, (c_lo, c_hi), np.stack([z[c_lo:c_hi] for z in zs_full]), half, zs_g, wait=False)
    fe2.wait(res)
    for k in ("canal_axes", "feat", "peak_theta", "peak_index", "n_peaks", "X", "scaler", "bg_theta", "points", "local_theta", "image", "minmax"):
        assert np.array_equal(out2[k], out[k]), k
    res.close()
    forest.close()
