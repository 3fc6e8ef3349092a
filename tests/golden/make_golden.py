"""Freezes ORACLE outputs for the reference's test bones as small regression fixtures.

    python tests/golden/make_golden.py

These pin the oracle (and, through the GPU tests, the CUDA path) against drift.  They are NOT
reference outputs: the reference pins none and trimesh cannot run here (oracle/__init__.py).
Per bone: the 40-plane Full sweep (N=64): per-plane segment count, a checksum of face_index,
entity count, centroid, area1, and the ixy / itr_start / itr_centered_start arrays."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from shoulder_b200.meshio import PcaObb  # noqa: E402

HERE = Path(__file__).resolve().parent
P, N = 40, 64


def golden_for(name):
    obb = PcaObb(HERE / "bones" / f"{name}.npz")
    m = obb.mesh
    z = m.vertices[:, 2]
    zs = np.linspace(0.99 * z.max(), 0.99 * z.min(), P)
    s = oracle.OracleSlices(m.vertices, m.faces, zs, N)
    fsum = np.array([int(p.metadata["face_index"].astype(np.int64).sum()) for p in s.paths])
    fxor = np.array([int(np.bitwise_xor.reduce(p.metadata["face_index"].astype(np.int64) * 2654435761 % (1 << 31))) for p in s.paths])
    return dict(transform=obb.transform, zs=zs, n_seg=np.array([len(p.metadata["face_index"]) for p in s.paths]), face_sum=fsum, face_xor=fxor,
                n_ent=s.n_entities, centroids=s.centroids, areas1=s.areas1, ixy=s.ixy, itr_start=s.itr_start,
                itr_centered_start=s.itr_centered_start)


if __name__ == "__main__":
    for name in ("humerus_left", "humerus_right", "humerus_left_trab", "humerus_left_flipped"):
        g = golden_for(name)
        np.savez_compressed(HERE / f"golden_{name}.npz", **g)
        print(name, int(g["n_seg"].sum()), "segments")
