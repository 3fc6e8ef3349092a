"""Restatement of the reference's per-row feature extraction on the polar stacks — TEST INFRASTRUCTURE.

Follows the reference source (scipy / scikit-learn are installed, so the library calls are the reference's own):
  groove_features   bicipital_groove.py:28-156   ``_X_process``: Savitzky-Golay(10, 1) of the negated zero-mean radius,
                    roll to the arg-min, ``find_peaks(height=-10, prominence=0.6, width=0.1)``, the 7 most prominent
                    peaks, 9 features per peak, ``StandardScaler``
  groove_theta      bicipital_groove.py:184-188   linear-kernel density of the accepted peaks' theta over 1,024 angles
  groove_points     bicipital_groove.py:190-238   local minimum of the zero-mean radius within +-deg_window of bg_theta
  neck_image        anatomic_neck.py:34-58        even-theta re-interpolation, roll to the groove, MinMax image
Checked against the reference's own ``DeepGroove.points()`` run in the build container (tests/test_groove_pin.py).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.signal

N_TOP = 7          # bicipital_groove.py:122


def _closest_angles(array, v):
    return np.abs([math.atan2(math.sin(v - a), math.cos(v - a)) for a in array])


def _peak_nearest(theta, which):
    """``peak_nearest`` (which = 0) / ``peak_next_nearest`` (which = 1), bicipital_groove.py:40-65."""
    n = len(theta)
    if n <= which + 1:
        return np.zeros(n)
    out = []
    for p in theta:
        angs = _closest_angles(theta, p)
        angs = angs[np.round(angs, 2) != 0]
        angs.sort()
        out.append(angs[which])
    return np.array(out)


def row_peaks(rpol0_radius, interp_num):
    """Peaks of one row: (indices into the row, properties dict) after the top-7 selection."""
    radius = scipy.signal.savgol_filter(-1 * rpol0_radius, 10, 1)
    rmin = -1 * np.argmin(radius)
    radius_roll = np.roll(radius, rmin)
    peaks, prop = scipy.signal.find_peaks(radius_roll, height=-10, prominence=0.6, width=0.1)
    peaks = (peaks - rmin) % interp_num
    if len(peaks) > N_TOP:
        part = np.argpartition(prop["prominences"], -N_TOP)[-N_TOP:]
        peaks = peaks[part]
        prop = {k: np.asarray([v[i] for i in part]) for k, v in prop.items()}
    return peaks, prop, radius


def groove_features(polar, zs, canal_axis, interp_num):
    """polar: (R, 2, N) ``itr_centered_start`` window; zs: (R,); canal_axis: (2, 3) as ``Canal.axis()`` returns it.
    Returns dict with the raw feature matrix (one row per peak), its StandardScaler form ``X``, ``peak_theta``,
    ``peak_row`` (which stack row each peak belongs to) and the smoothed rows."""
    polar = np.asarray(polar, dtype=np.float64)
    zs = np.asarray(zs, dtype=np.float64)
    polar_0 = polar.copy()
    polar_0[:, 1, :] = polar[:, 1, :] - polar[:, 1, :].mean(axis=1, keepdims=True)      # apply_along_axis(x - mean(x)), :166-168
    z_scale = (zs - zs.min()) / (zs.max() - zs.min())                                    # MinMaxScaler, :92
    u = canal_axis[0] - canal_axis[1]
    canal_u = u / np.linalg.norm(u)                                                      # utils.unit_vector
    rows, theta_all, prow, smooth = [], [], [], []
    for i, (rpol, rpol0) in enumerate(zip(polar, polar_0)):
        theta, radius_og = rpol0[0], rpol[1]
        peaks, prop, sm = row_peaks(rpol0[1], interp_num)
        smooth.append(sm)
        th = theta[peaks]
        canal_xy = (canal_u.reshape(-1, 1) @ np.repeat(zs[i], len(peaks)).reshape(1, -1))[:2, :]
        pk_xy = np.c_[radius_og[peaks] * np.cos(th), radius_og[peaks] * np.sin(th)].T
        dist = np.sqrt(np.sum((pk_xy - canal_xy) ** 2, axis=0))
        for k in range(len(peaks)):
            rows.append([radius_og[peaks][k], _peak_nearest(th, 0)[k], _peak_nearest(th, 1)[k], z_scale[i], prop["prominences"][k],
                         prop["widths"][k], prop["width_heights"][k], dist[k], len(peaks) / N_TOP])
        theta_all.extend(th)
        prow.extend([i] * len(peaks))
    raw = np.asarray(rows, dtype=np.float64).reshape(-1, 9)
    mean, std = raw.mean(axis=0), raw.std(axis=0)
    std = np.where(std == 0.0, 1.0, std)                                                 # sklearn's _handle_zeros_in_scale
    return {"raw": raw, "X": (raw - mean) / std, "peak_theta": np.asarray(theta_all), "peak_row": np.asarray(prow, dtype=np.int64),
            "smooth": np.asarray(smooth), "polar_0": polar_0}


def groove_theta(peak_theta, proba1, threshold=0.4):
    """KernelDensity(kernel='linear', bandwidth=1.0) over the accepted peaks, evaluated on 1,024 angles; its arg-max."""
    pts = np.asarray(peak_theta)[np.asarray(proba1) > threshold]
    tlin = np.linspace(-np.pi, np.pi, 1024)
    dens = np.maximum(0.0, 1.0 - np.abs(tlin[:, None] - pts[None, :])).sum(axis=1)       # linear kernel, h = 1 (normalisation is monotone)
    return float(tlin[np.argmax(dens)]), dens


def groove_points(polar, polar_0, zs, centroids, bg_theta, interp_num, deg_window=7):
    ivar = max(1, int(round(deg_window / (360 / interp_num))))
    out = np.zeros((len(zs), 3))
    local_theta = np.zeros(len(zs))
    for i, z in enumerate(zs):
        row = polar_0[i, 0, :]
        est = int(np.searchsorted(row, bg_theta, side="left"))
        if est == len(row):
            est -= 1
        if ivar > est:
            rng = np.concatenate((polar_0[i, :, (est - ivar):], polar_0[i, :, :(est + ivar)]), axis=1)
        else:
            rng = polar_0[i, :, (est - ivar):(est + ivar)]
        loc = int(np.argmin(rng[1, :])) + (est - ivar)
        local_theta[i] = polar[i, 0, loc]
        r, t = polar[i, 1, loc], polar[i, 0, loc]
        out[i] = [r * np.cos(t) + centroids[i, 0], r * np.sin(t) + centroids[i, 1], z]
    return out, local_theta


def neck_image(itr, bg_theta):
    """anatomic_neck.py:38-58: rows re-interpolated on even theta, rolled to the groove, MinMax over the whole image.
    Returns (image float64 in [0, 1], itr_shft, (min, max))."""
    itr = np.asarray(itr, dtype=np.float64)
    image = np.zeros((itr.shape[0], itr.shape[2]))
    shft = np.zeros(itr.shape)
    for i, tr in enumerate(itr):
        t = np.linspace(tr[0][0], tr[0][-2], tr.shape[1])
        tr2 = np.c_[t, np.interp(t, tr[0, :-1], tr[1, :-1])].T
        k = int(np.argmin(np.abs(tr2[0] - bg_theta)))
        tr2 = np.c_[tr2[:, k:], tr2[:, :k]]
        image[i] = tr2[1]
        shft[i] = tr2
    lo, hi = image.min(), image.max()
    return (image - lo) / (hi - lo), shft, (lo, hi)
