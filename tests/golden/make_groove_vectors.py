"""Vectors of the groove / neck-image feature stage made by the REFERENCE'S OWN CODE (build container only):

    python tests/golden/make_groove_vectors.py        -> tests/golden/refgroove_<bone>.npz

``bicipital_groove.DeepGroove.points()`` and ``anatomic_neck.AnatomicNeck.points()`` are imported unchanged
(tests/refload.py) and run over the reference's ``ProximalSlices`` (600 x 512, "must not change", slice.py:236) and
``Canal``, with trimesh answered by the oracle and onnxruntime by ``oracle/onnx_forest.py`` reading the reference's own
``rfc_bg3.onnx``.  The UNet blob of the anatomic neck is missing from the checkout, so that method is run up to the model
call: the stand-in session captures the 1 x 1 x 512 x 512 float32 input the reference built and stops there.
Stored: scaled feature matrix X, peak thetas, bg_theta, the groove points (OBB frame), every 8th row of the neck
image, the forest's probabilities, and the inputs (frame transform, neck_z) to rebuild everything on another host.
"""
from __future__ import annotations

import importlib.resources
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

NAMES = ["humerus_left", "humerus_right"]
NECK_FRAC = 0.55
IMAGE_STRIDE = 8


class _Captured(Exception):
    pass


def run_reference(vertices, faces, transform):
    """Returns dict of what the reference computed for one bone (frame-transformed vertices in)."""
    import refload
    from oracle import onnx_forest
    m = refload.reference_modules()
    S, C, BG = m["slice"], m["canal"], m["bicipital_groove"]
    AN = refload.load_extra("anatomic_neck")
    obb = refload.OracleObb(vertices, faces, transform=transform)

    class T:
        matrix = np.eye(4)
    full = S.FullSlices(obb)
    canal = C.Canal(full, T())
    zmax = obb.mesh.bounds[1, 2]
    prox = S.ProximalSlices(obb, refload.Neck(NECK_FRAC * zmax))
    captured = {}

    class UnetStub:
        def __init__(self, *a, **k):
            pass

        def get_inputs(self):
            return [type("I", (), {"name": "input"})()]

        def run(self, _, feed):
            captured["image"] = next(iter(feed.values())).copy()
            raise _Captured()

    forest_cls = onnx_forest.InferenceSession
    probs = {}

    class ForestSpy(forest_cls):
        def run(self, _, feed):
            out = super().run(_, feed)
            probs["proba"] = out[1].copy()
            return out

    orig_files = importlib.resources.files
    tmp = Path("/tmp/shb_ref_models")
    (tmp / "humerus" / "models").mkdir(parents=True, exist_ok=True)
    (tmp / "humerus" / "models" / "unetcrf_anp.onnx").write_bytes(b"missing from the reference checkout")
    (tmp / "humerus" / "models" / "rfc_bg3.onnx").write_bytes((refload.REF / "humerus" / "models" / "rfc_bg3.onnx").read_bytes())
    importlib.resources.files = lambda name: tmp
    try:
        BG.rt.InferenceSession = ForestSpy
        dg = BG.DeepGroove(prox, canal, T())
        dg.points()
        AN.rt.InferenceSession = UnetStub
        an = AN.AnatomicNeck(prox, dg, T())
        try:
            an.points()
        except _Captured:
            pass
    finally:
        importlib.resources.files = orig_files
    cutoff = (0.2, 0.75)
    return {
        "X": dg._X, "peak_theta": dg._peak_theta, "proba": probs["proba"], "bg_theta": np.float64(dg.bg_theta),
        "groove_points_obb": dg._points_obb, "canal_axis": np.asarray(canal.axis()),
        "image": captured["image"][0, 0], "groove_zs": np.asarray(prox.zs(cutoff)),
        "neck_z": np.float64(NECK_FRAC * zmax),
    }, prox


def main():
    import refload
    from shoulder_b200.meshio import PcaObb, load_mesh
    for name in NAMES:
        ct = load_mesh(HERE / "bones" / f"{name}.npz")
        transform = PcaObb(ct).transform
        vv, ff = refload.exact_frame(ct.vertices, ct.faces, transform)
        out, _ = run_reference(vv, ff, transform)
        img = out.pop("image")
        out["image_rows"] = img[::IMAGE_STRIDE].copy()
        out["image_shape"] = np.array(img.shape)
        out["image_sum"] = np.float64(img.astype(np.float64).sum())
        out["transform"] = transform
        np.savez_compressed(HERE / f"refgroove_{name}.npz", **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items()}, sum(np.asarray(v).nbytes for v in out.values()) // 1024, "KiB")


if __name__ == "__main__":
    main()
