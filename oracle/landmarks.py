"""Canal axis from slice centroids — TEST INFRASTRUCTURE (restates reference
``src/shoulder/humerus/canal.py:40-85`` with numpy only; skspatial's ``Line.best_fit`` is the
centroid plus the first right-singular vector of the centred points)."""
import numpy as np

from .slice_arrays import cutoff_window


def canal_axis_obb(centroids_xy, zs, z_length, cutoff_pcts=(0.35, 0.75)):
    """(2,3) proximal/distal end points of the canal centre line in the OBB frame."""
    lo, hi = cutoff_window(len(zs), cutoff_pcts)
    pts = np.c_[np.asarray(centroids_xy)[lo:hi], np.asarray(zs)[lo:hi]]
    mid = pts.mean(axis=0)
    _, _, vh = np.linalg.svd(pts - mid)
    direction = vh[0]
    if direction[-1] < 0:
        direction = -direction
    half = z_length * np.mean(cutoff_pcts) / 2
    return np.array([mid + direction * half, mid - direction * half])


def axis_angle_deg(a, b):
    u, v = a[0] - a[1], b[0] - b[1]
    c = abs(np.dot(u, v)) / (np.linalg.norm(u) * np.linalg.norm(v))
    return float(np.degrees(np.arccos(min(1.0, c))))
