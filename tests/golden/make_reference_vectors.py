"""Vectors produced by the REFERENCE'S OWN CODE, run in the build container where /root/reference exists:

    python tests/golden/make_reference_vectors.py        -> tests/golden/refvec_<bone>.npz

``/root/reference/src/shoulder/humerus/slice.py`` is imported unchanged (tests/refload.py) and its
``FullSlices`` / ``DistalSlices`` / ``ProximalSlices`` are run over an ``obb.mesh`` whose ``section_multiplane`` is
answered by the oracle's restatement of trimesh (trimesh itself is not installable).  Everything from
``Slices.__init__`` to the eight cached arrays (slice.py:10-147) is therefore computed by reference code; the files
pin rows a1, a2, a7-a13 of the scope table to it.  ``canal.py`` (imported unchanged, scikit-spatial's
``Line.best_fit`` stubbed by its definition) adds the canal axis.  The GPU box has no /root/reference: the ``-m gpu``
tests compare the CUDA path with these files; a CPU test checks that the files still equal a live run.
Sizes are reduced (constructor arguments of the reference classes) to keep the fixtures small.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

NAMES = ["humerus_left", "humerus_right", "humerus_left_trab", "humerus_left_flipped"]
SIZES = {"full": (36, 32), "distal": (30, 40), "proximal": (40, 64)}       # (zslice_num, interp_num)
ARRAYS = ["_centroids", "_areas1", "_ixy", "_ixy_centered", "_itr", "_itr_start", "_itr_start_even_theta", "_itr_centered",
          "_itr_centered_start"]
NECK_FRAC = 0.55          # neck_z stand-in (ruptures, which finds it in surgical_neck.py:31-34, is not installable)


def reference_objects(vertices, faces, sizes=SIZES, merge="hash"):
    import refload
    S = refload.reference_modules()["slice"]
    obb = refload.OracleObb(vertices, faces, merge=merge)
    zmax = obb.mesh.bounds[1, 2]
    return obb, {
        "full": S.FullSlices(obb, zslice_num=sizes["full"][0], interp_num=sizes["full"][1]),
        "distal": S.DistalSlices(obb, zslice_num=sizes["distal"][0], interp_num=sizes["distal"][1]),
        "proximal": S.ProximalSlices(obb, refload.Neck(NECK_FRAC * zmax), zslice_num=sizes["proximal"][0],
                                     interp_num=sizes["proximal"][1]),
    }


def vectors(vertices, faces):
    import refload
    obb, objs = reference_objects(vertices, faces)
    out = {}
    for key, s in objs.items():
        out[f"{key}__zs"] = np.asarray(s._zs)
        out[f"{key}__z_orig"] = np.asarray(s._z_orig)
        out[f"{key}__z_incrs"] = np.asarray(s._z_incrs)
        for a in ARRAYS:
            out[f"{key}_{a}"] = np.asarray(getattr(s, a))
        out[f"{key}__n_entities"] = np.array([len(p.entities) for p in s._slices])
        # windows as the consumers ask for them (slice.py:157-164 through the public accessors)
        out[f"{key}__win_035_075"] = np.asarray(s.centroids((0.35, 0.75)))
        out[f"{key}__win_itr_bug"] = np.asarray(s.itr((0.2, 0.8)))                 # sic: returns _ixy (slice.py:99-100)
    C = refload.reference_modules()["canal"]

    class T:
        matrix = np.eye(4)
    out["canal_axis"] = np.asarray(C.Canal(objs["full"], T()).axis())              # canal.py:40-85 on the Full sweep
    return out


def main():
    import refload
    from shoulder_b200.meshio import PcaObb, load_mesh
    for name in NAMES:
        ct = load_mesh(HERE / "bones" / f"{name}.npz")
        transform = PcaObb(ct).transform                 # stored: the test rebuilds the frame bit for bit (refload.exact_frame)
        vv, ff = refload.exact_frame(ct.vertices, ct.faces, transform)
        v = vectors(vv, ff)
        v["transform"] = transform
        np.savez_compressed(HERE / f"refvec_{name}.npz", **v)
        print(name, sum(a.nbytes for a in v.values()) // 1024, "KiB")


if __name__ == "__main__":
    main()
