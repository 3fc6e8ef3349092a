"""Host side of the device feature extraction (scope row f3) and of the groove detector's random forest (part of f4).

Mirrors what ``bicipital_groove.DeepGroove.points()`` (bicipital_groove.py:94-238) and ``anatomic_neck.AnatomicNeck.points()``
(anatomic_neck.py:34-58) do between ``slice.py``'s polar stacks and their models, with the per-row loops on the device
(``shb_groove_features`` / ``shb_groove_points`` / ``shb_neck_image``) and the forest on the device (``shb_forest_*``).
What stays on the host is O(peaks): the StandardScaler (two numbers per column) and the linear-kernel density over the
accepted peaks' theta (1,024 x a few hundred).  No CPU fallback: every array comes from the library.
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from . import _lib

N_TOP, N_FEAT = 7, 9


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _windows(result, sweeps, which):
    return [result.window(which, s) for s in sweeps]


def groove_features(result, sweeps, zs_list, canal_axes):
    """Per listed sweep: dict(raw (n_peaks, 9), X (StandardScaler), peak_theta, peak_row, peak_index)."""
    lib = _lib.load()
    sweeps = np.asarray(sweeps, dtype=np.int32)
    zs = np.ascontiguousarray(np.concatenate([np.asarray(z, dtype=np.float64) for z in zs_list]))
    axes = np.ascontiguousarray(np.asarray(canal_axes, dtype=np.float64).reshape(len(sweeps), 2, 3))
    rows = len(zs)
    feat = _lib.pinned_empty((rows, N_TOP, N_FEAT), np.float64); theta = _lib.pinned_empty((rows, N_TOP), np.float64)
    idx = _lib.pinned_empty((rows, N_TOP), np.int32); cnt = _lib.pinned_empty((rows,), np.int32)
    _lib.check(lib.shb_groove_features(result._h, len(sweeps), _p(sweeps), _p(zs), _p(axes), _p(feat), _p(theta), _p(idx), _p(cnt)))
    # one gather for the whole batch, then per-bone slices (the StandardScaler is per bone)
    sel = np.arange(N_TOP)[None, :] < cnt[:, None]
    raw_all, theta_all, idx_all = feat[sel], theta[sel], idx[sel]
    out, r0, p0 = [], 0, 0
    for z in zs_list:
        n = len(z)
        c = cnt[r0:r0 + n].copy()
        p1 = p0 + int(c.sum())
        raw = raw_all[p0:p1]
        mean, std = raw.mean(axis=0), raw.std(axis=0)
        std = np.where(std == 0.0, 1.0, std)                         # sklearn's _handle_zeros_in_scale
        out.append({"raw": raw, "X": (raw - mean) / std, "peak_theta": theta_all[p0:p1],
                    "peak_row": np.repeat(np.arange(n), c), "peak_index": idx_all[p0:p1], "n_peaks": c})
        r0, p0 = r0 + n, p1
    return out


def groove_theta(peak_theta, proba1, threshold: float = 0.4) -> float:
    """bicipital_groove.py:184-188: arg-max over 1,024 angles of the linear-kernel density (bandwidth 1) of the accepted peaks."""
    pts = np.asarray(peak_theta)[np.asarray(proba1) > threshold]
    tlin = np.linspace(-np.pi, np.pi, 1024)
    dens = np.maximum(0.0, 1.0 - np.abs(tlin[:, None] - pts[None, :])).sum(axis=1)
    return float(tlin[np.argmax(dens)])


def groove_theta_batch(peak_thetas, probas1, threshold: float = 0.4) -> np.ndarray:
    """:func:`groove_theta` for many bones in one device call (``shb_groove_theta``): lists of per-bone peak angles and
    class-1 probabilities -> (n_bones,) groove angles.  Same grid, same first-maximum rule; the density is summed in peak
    order (numpy sums pairwise), so the two can only differ where two grid angles tie within rounding."""
    off = np.cumsum([0] + [len(t) for t in peak_thetas]).astype(np.int64)
    th = np.ascontiguousarray(np.concatenate([np.asarray(t, dtype=np.float64) for t in peak_thetas])) if off[-1] else np.zeros(1)
    pr = np.ascontiguousarray(np.concatenate([np.asarray(q, dtype=np.float32) for q in probas1])) if off[-1] else np.zeros(1, np.float32)
    bg = np.zeros(len(peak_thetas))
    _lib.init(_lib._inited if _lib._inited is not None else 0)
    _lib.check(_lib.load().shb_groove_theta(len(peak_thetas), _p(off), _p(th), _p(pr), C.c_float(threshold), _p(bg), C.c_void_p(0)))
    return bg


def groove_points(result, sweeps, zs_list, bg_thetas, interp_num: int, deg_window: float = 7):
    lib = _lib.load()
    sweeps = np.asarray(sweeps, dtype=np.int32)
    zs = np.ascontiguousarray(np.concatenate([np.asarray(z, dtype=np.float64) for z in zs_list]))
    bg = np.ascontiguousarray(np.asarray(bg_thetas, dtype=np.float64))
    ivar = max(1, int(round(deg_window / (360 / interp_num))))
    pts = np.zeros((len(zs), 3)); lt = np.zeros(len(zs))
    _lib.check(lib.shb_groove_points(result._h, len(sweeps), _p(sweeps), _p(bg), ivar, _p(zs), _p(pts), _p(lt)))
    cuts = np.cumsum([0] + [len(z) for z in zs_list])
    return [(pts[a:b], lt[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]


def neck_image(result, sweeps, bg_thetas, want_shifted: bool = False, interp_num: int | None = None):
    """Per listed sweep: (image float32 (rows, N) in [0, 1], (min, max), itr_shft or None).  Pass ``interp_num`` (the N of
    the listed sweeps): without it N is read off the itr_start array itself, which copies that whole stack to the host."""
    lib = _lib.load()
    sweeps = np.asarray(sweeps, dtype=np.int32)
    bg = np.ascontiguousarray(np.asarray(bg_thetas, dtype=np.float64))
    wins = _windows(result, sweeps, _lib.ARR_ITR_START)
    rows = [hi - lo for lo, hi in wins]
    N = int(interp_num) if interp_num is not None else result.array_shape(_lib.ARR_ITR_START, int(sweeps[0]))[2]
    image = _lib.pinned_empty((sum(rows), N), np.float32)
    shft = _lib.pinned_empty((sum(rows), 2, N), np.float64) if want_shifted else None
    mm = np.zeros((len(sweeps), 2))
    _lib.check(lib.shb_neck_image(result._h, len(sweeps), _p(sweeps), _p(bg), _p(image), _p(shft) if want_shifted else C.c_void_p(0), _p(mm)))
    cuts = np.cumsum([0] + rows)
    return [(image[a:b], tuple(mm[k]), shft[a:b] if want_shifted else None) for k, (a, b) in enumerate(zip(cuts[:-1], cuts[1:]))]


# ------------------------------------------------------------------------------------------
# the groove detector's random forest: rfc_bg3.onnx -> flat node arrays -> device
# ------------------------------------------------------------------------------------------
def _varint(buf, pos):
    val, shift = 0, 0
    while True:
        b = buf[pos]; pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _fields(buf):
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"wire type {wt}")
        yield num, wt, val


def read_onnx_tree_ensemble(path) -> dict:
    """Attributes of the TreeEnsembleClassifier node of an onnx file (protobuf wire format read directly: the image has
    no ``onnx`` package).  ModelProto.graph = 7, GraphProto.node = 1, NodeProto.op_type = 4 / .attribute = 5,
    AttributeProto.name = 1 / .f = 2 / .i = 3 / .s = 4 / .floats = 7 / .ints = 8 / .strings = 9."""
    model = memoryview(open(path, "rb").read())
    graph = next(v for num, wt, v in _fields(model) if num == 7 and wt == 2)
    for num, wt, node in _fields(graph):
        if num != 1 or wt != 2:
            continue
        op, attrs = None, {}
        for fn, fwt, fv in _fields(node):
            if fn == 4 and fwt == 2:
                op = bytes(fv).decode()
            elif fn == 5 and fwt == 2:
                name, floats, ints, strings, scalar = None, [], [], [], None
                for an, awt, av in _fields(fv):
                    if an == 1:
                        name = bytes(av).decode()
                    elif an == 7:
                        floats.append(np.frombuffer(av, dtype="<f4") if awt == 2 else np.frombuffer(av, dtype="<f4", count=1))
                    elif an == 8:
                        if awt == 2:
                            pos, b = 0, bytes(av)
                            while pos < len(b):
                                v, pos = _varint(b, pos)
                                ints.append(v if v < (1 << 63) else v - (1 << 64))
                        else:
                            ints.append(av)
                    elif an == 9:
                        strings.append(bytes(av).decode())
                    elif an == 4:
                        scalar = bytes(av).decode()
                    elif an == 3:
                        scalar = av
                    elif an == 2:
                        scalar = struct.unpack("<f", av)[0]
                if floats:
                    attrs[name] = np.concatenate(floats)
                elif ints:
                    attrs[name] = np.asarray(ints, dtype=np.int64)
                elif strings:
                    attrs[name] = strings
                else:
                    attrs[name] = scalar
        if op == "TreeEnsembleClassifier":
            return attrs
    raise ValueError(f"no TreeEnsembleClassifier node in {path}")


def flatten_tree_ensemble(a: dict) -> dict:
    """onnx TreeEnsembleClassifier attributes -> flat node arrays (tree after tree, children behind their parent)."""
    modes = np.asarray(a["nodes_modes"])
    if not set(modes) <= {"BRANCH_LEQ", "LEAF"} or (a.get("post_transform") or "NONE") != "NONE":
        raise _lib.BackendError("only BRANCH_LEQ forests without post_transform are supported")
    if len(a["classlabels_int64s"]) != 2 or len(set(a["class_ids"].tolist())) != 1 or a["class_weights"].min() < 0:
        raise _lib.BackendError("expected a binary forest with one non-negative weight per leaf (what skl2onnx writes)")
    tid, nid = a["nodes_treeids"], a["nodes_nodeids"]
    n_trees, n = int(tid.max()) + 1, len(tid)
    count = np.bincount(tid, minlength=n_trees)
    base = np.concatenate([[0], np.cumsum(count)[:-1]])
    flat = base[tid] + nid
    order = np.argsort(flat)
    if not np.array_equal(flat[order], np.arange(n)):
        raise _lib.BackendError("node ids of a tree are not 0 .. n-1")
    leaf = (modes == "LEAF")[order]
    weight = np.zeros(n, dtype=np.float32)
    np.add.at(weight, base[a["class_treeids"]] + a["class_nodeids"], a["class_weights"].astype(np.float32))
    return {"root": base.astype(np.uint32), "feature": np.where(leaf, -1, a["nodes_featureids"][order]).astype(np.int32),
            "value": a["nodes_values"][order].astype(np.float32),
            "true_child": (base[tid[order]] + a["nodes_truenodeids"][order]).astype(np.uint32),
            "false_child": (base[tid[order]] + a["nodes_falsenodeids"][order]).astype(np.uint32), "weight": weight,
            "labels": np.asarray(a["classlabels_int64s"], dtype=np.int64)}


class Forest:
    """Device copy of a binary TreeEnsembleClassifier (``shb_forest_create``); ``predict_proba`` as onnxruntime returns it."""

    def __init__(self, onnx_path=None, arrays=None):
        fa = arrays if arrays is not None else flatten_tree_ensemble(read_onnx_tree_ensemble(onnx_path))
        self.arrays = {k: np.ascontiguousarray(fa[k]) for k in ("root", "feature", "value", "true_child", "false_child", "weight", "labels")}
        self.n_trees, self.n_nodes = len(self.arrays["root"]), len(self.arrays["feature"])
        self.n_features = int(self.arrays["feature"].max()) + 1
        self.labels = self.arrays["labels"]
        _lib.init(_lib._inited if _lib._inited is not None else 0)
        h = C.c_void_p()
        order = ("root", "feature", "value", "true_child", "false_child", "weight")
        _lib.check(_lib.load().shb_forest_create(self.n_nodes, self.n_trees, self.n_features, *[_p(self.arrays[k]) for k in order], C.byref(h)))
        self._h = h

    @classmethod
    def from_arrays(cls, arrays) -> "Forest":
        return cls(arrays=arrays)

    def predict_proba(self, X) -> np.ndarray:
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float32))
        if X.ndim != 2 or X.shape[1] != self.n_features:
            raise ValueError(f"X must be (n, {self.n_features})")
        s = np.zeros(len(X), dtype=np.float32)
        _lib.check(_lib.load().shb_forest_predict(self._h, _p(X), len(X), _p(s)))
        return np.stack([1.0 - s, s], axis=1).astype(np.float32)      # onnxruntime's binary case with all-positive weights

    def close(self):
        if self._h:
            _lib.load().shb_forest_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def detect_groove(result, sweep, zs, canal_axis, forest: Forest, interp_num: int, deg_window: float = 7):
    """The whole pre-UNet groove chain for one proximal sweep, stacks resident on the device:
    features -> StandardScaler -> forest -> density arg-max -> local minima.  Returns dict(bg_theta, points_obb, X, ...)."""
    ft = groove_features(result, [sweep], [zs], [canal_axis])[0]
    proba = forest.predict_proba(ft["X"])
    bg = groove_theta(ft["peak_theta"], proba[:, 1])
    pts, local_theta = groove_points(result, [sweep], [zs], [bg], interp_num, deg_window)[0]
    return {"bg_theta": bg, "points_obb": pts, "local_theta": local_theta, "proba": proba, **ft}


# ------------------------------------------------------------------------------------------
# the whole front end of a batch in one enqueue (shb_landmark_front)
# ------------------------------------------------------------------------------------------
class _LandmarkArgs(C.Structure):
    _fields_ = [("n_bones", C.c_int32), ("canal_lo", C.c_int32), ("canal_hi", C.c_int32),
                ("full_sweeps", C.c_void_p), ("prox_sweeps", C.c_void_p), ("canal_z", C.c_void_p), ("canal_half", C.c_void_p),
                ("groove_zs", C.c_void_p), ("forest", C.c_void_p), ("threshold", C.c_float), ("ivar", C.c_int32),
                ("canal_axes", C.c_void_p), ("feat", C.c_void_p), ("peak_theta", C.c_void_p), ("peak_index", C.c_void_p),
                ("n_peaks", C.c_void_p), ("X", C.c_void_p), ("proba1", C.c_void_p), ("scaler", C.c_void_p), ("bg_theta", C.c_void_p),
                ("points", C.c_void_p), ("local_theta", C.c_void_p), ("image", C.c_void_p), ("minmax", C.c_void_p), ("flags", C.c_uint32)]


class LandmarkFrontEnd:
    """Canal axis -> groove features -> StandardScaler -> forest -> groove angle -> groove points -> neck image for a batch of
    bones whose sweeps are resident in a ``SweepResult``, in ONE device enqueue and one host wait (``shb_landmark_front``):
    what ``canal.Canal.axis`` (canal.py:40-85), ``DeepGroove.points`` (bicipital_groove.py:94-238) and the image stage of
    ``AnatomicNeck.points`` (anatomic_neck.py:38-58) do between ``slice.py``'s arrays and their models.  The object owns the
    page-locked output buffers and reuses them call after call (same batch shape), so a step allocates nothing.

    ``canal_rows`` = ``Slices._cutoff((0.35, 0.75))`` of the Full sweep, ``canal_z`` (n_bones, rows) the z of those planes,
    ``canal_half`` (n_bones,) = ``obb.z_length * mean(cutoff_pcts) / 2``; ``groove_zs`` one z array per bone over the
    ``itr_centered_start`` window of its proximal sweep."""

    def __init__(self, forest: Forest, interp_num: int = 512, deg_window: float = 7, threshold: float = 0.4):
        self.forest, self.threshold = forest, float(threshold)
        self.ivar = max(1, int(round(deg_window / (360 / interp_num))))
        self.interp_num = int(interp_num)
        self._shape, self._buf = None, {}

    def _buffers(self, n_bones, rows, img_rows, N):
        shape = (n_bones, rows, img_rows, N)
        if shape != self._shape:
            pe = _lib.pinned_empty
            self._buf = {"canal_axes": pe((n_bones, 2, 3), np.float64), "feat": pe((rows, N_TOP, N_FEAT), np.float64),
                         "peak_theta": pe((rows, N_TOP), np.float64), "peak_index": pe((rows, N_TOP), np.int32), "n_peaks": pe((rows,), np.int32),
                         "X": pe((rows, N_TOP, N_FEAT), np.float32), "proba1": pe((rows, N_TOP), np.float32), "scaler": pe((n_bones, 2, N_FEAT), np.float64),
                         "bg_theta": pe((n_bones,), np.float64), "points": pe((rows, 3), np.float64), "local_theta": pe((rows,), np.float64),
                         "image": pe((img_rows, N), np.float32), "minmax": pe((n_bones, 2), np.float64)}
            self._shape = shape
        return self._buf

    def __call__(self, result, full_sweeps, prox_sweeps, canal_rows, canal_z, canal_half, groove_zs, wait: bool = True):
        """``wait=False`` returns once the work is enqueued (``SHB_LF_NO_WAIT``): the arrays are filled by the time
        :meth:`wait` returns; meanwhile the next batch can be uploaded and computed (use one object per batch in flight).

        Returns a dict of batch arrays (views of the object's buffers, valid until the next call): ``canal_axes`` (B,2,3),
        ``feat`` / ``X`` (rows,7,9), ``peak_theta`` / ``peak_index`` / ``proba1`` (rows,7), ``n_peaks`` (rows,), ``scaler`` (B,2,9),
        ``bg_theta`` (B,), ``points`` (rows,3), ``local_theta`` (rows,), ``image`` (image rows, N) float32, ``minmax`` (B,2), and
        ``row_cuts`` / ``image_cuts``: the row range of bone b is ``cuts[b]:cuts[b+1]``."""
        lib = _lib.load()
        full = np.ascontiguousarray(full_sweeps, dtype=np.int32); prox = np.ascontiguousarray(prox_sweeps, dtype=np.int32)
        nb = len(full)
        cz = np.ascontiguousarray(canal_z, dtype=np.float64).reshape(nb, canal_rows[1] - canal_rows[0])
        half = np.ascontiguousarray(canal_half, dtype=np.float64).reshape(nb)
        zs = np.ascontiguousarray(np.concatenate([np.asarray(z, dtype=np.float64) for z in groove_zs]))
        wins = [result.window(_lib.ARR_ITR_START, int(s)) for s in prox]
        img_rows = [hi - lo for lo, hi in wins]
        N = self.interp_num
        buf = self._buffers(nb, len(zs), sum(img_rows), N)
        a = _LandmarkArgs(nb, int(canal_rows[0]), int(canal_rows[1]), full.ctypes.data, prox.ctypes.data, cz.ctypes.data, half.ctypes.data,
                          zs.ctypes.data, self.forest._h, self.threshold, self.ivar, *[buf[k].ctypes.data for k in
                          ("canal_axes", "feat", "peak_theta", "peak_index", "n_peaks", "X", "proba1", "scaler", "bg_theta", "points",
                           "local_theta", "image", "minmax")], 0 if wait else 1)
        self._keep = (full, prox, cz, half, zs)                       # read by the enqueued copies until they are staged (the call stages them)
        _lib.check(lib.shb_landmark_front(result._h, C.byref(a)))
        out = dict(buf)
        out["row_cuts"] = np.cumsum([0] + [len(z) for z in groove_zs])
        out["image_cuts"] = np.cumsum([0] + img_rows)
        return out

    @staticmethod
    def wait(result) -> None:
        """Blocks until the outputs of the ``wait=False`` calls on ``result`` have arrived."""
        _lib.check(_lib.load().shb_landmark_wait(result._h))
