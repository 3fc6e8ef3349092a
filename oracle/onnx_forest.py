"""Reader + evaluator for the reference's random-forest blob — TEST INFRASTRUCTURE.

``src/shoulder/humerus/models/rfc_bg3.onnx`` (skl2onnx export of a scikit-learn RandomForestClassifier: one
``TreeEnsembleClassifier`` node, input ``X`` float32 (n, 9), outputs ``label`` and ``probabilities``) is opened by
``bicipital_groove.py:174-181`` through onnxruntime, which is not installable here.  This module reads the protobuf wire
format directly (no ``onnx`` package) and evaluates the ensemble as the ONNX-ML operator specifies:
every tree is walked from its root — ``BRANCH_LEQ``: go to the true child when ``x[feature] <= value`` — the leaf's
class weights are summed over the trees, ``post_transform = NONE``.  The blob is a BINARY classifier whose leaves carry
one weight each, all under class id 0 (how skl2onnx writes binary forests: the weight is the leaf's class-1 fraction
divided by the number of trees); onnxruntime's binary case turns the summed score s into probabilities [1 - s, s]
(TreeAggregatorClassifier::FinalizeScores with all-positive weights), which is what ``pred_proba[:, 1] > 0.4``
(bicipital_groove.py:184-186) reads.
Used (a) as the ``onnxruntime`` stand-in when the reference's own ``DeepGroove.points()`` is run to make vectors, and
(b) as the checker of the device forest (``shoulder_b200`` ships its own reader: product code never imports oracle/).
"""
from __future__ import annotations

import struct

import numpy as np


def _varint(buf, pos):
    val, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _fields(buf):
    """Yields (field number, wire type, value) of one protobuf message (bytes)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield num, wt, val


def _packed_varints(b):
    out, pos = [], 0
    while pos < len(b):
        v, pos = _varint(b, pos)
        out.append(v if v < (1 << 63) else v - (1 << 64))
    return out


def read_tree_ensemble(path) -> dict:
    """Attributes of the first TreeEnsembleClassifier node of an ONNX file, as numpy arrays / lists."""
    model = open(path, "rb").read()
    graph = next(v for num, wt, v in _fields(model) if num == 7 and wt == 2)              # ModelProto.graph
    for num, wt, node in _fields(graph):
        if num != 1 or wt != 2:                                                             # GraphProto.node
            continue
        op, attrs = None, {}
        for fn, fwt, fv in _fields(node):
            if fn == 4 and fwt == 2:                                                        # NodeProto.op_type
                op = bytes(fv).decode()
            elif fn == 5 and fwt == 2:                                                      # NodeProto.attribute
                name, floats, ints, strings, scalar = None, [], [], [], None
                for an, awt, av in _fields(fv):
                    if an == 1:
                        name = bytes(av).decode()
                    elif an == 7:                                                           # floats (packed or not)
                        floats += list(struct.unpack(f"<{len(av) // 4}f", av)) if awt == 2 else [struct.unpack("<f", av)[0]]
                    elif an == 8:                                                           # ints
                        ints += _packed_varints(av) if awt == 2 else [av]
                    elif an == 9:
                        strings.append(bytes(av).decode())
                    elif an == 4:
                        scalar = bytes(av).decode()
                    elif an == 3:
                        scalar = av
                    elif an == 2:
                        scalar = struct.unpack("<f", av)[0]
                attrs[name] = floats or ints or strings or scalar
        if op == "TreeEnsembleClassifier":
            return attrs
    raise ValueError("no TreeEnsembleClassifier node in " + str(path))


class Forest:
    def __init__(self, path):
        a = read_tree_ensemble(path)
        self.attrs = a
        tid = np.asarray(a["nodes_treeids"], dtype=np.int64)
        nid = np.asarray(a["nodes_nodeids"], dtype=np.int64)
        self.n_trees = int(tid.max()) + 1
        self.n_classes = len(a["classlabels_int64s"])
        base = np.zeros(self.n_trees + 1, dtype=np.int64)
        np.add.at(base, tid + 1, 1)
        self.base = np.cumsum(base)[:-1]                                 # first flat node of every tree (nodes are stored tree by tree)
        flat = self.base[tid] + nid
        n = len(tid)
        assert np.array_equal(np.sort(flat), np.arange(n))
        self.feature = np.zeros(n, dtype=np.int64); self.feature[flat] = a["nodes_featureids"]
        self.value = np.zeros(n, dtype=np.float32); self.value[flat] = np.asarray(a["nodes_values"], dtype=np.float32)
        modes = np.array(a["nodes_modes"])
        assert set(modes) <= {"BRANCH_LEQ", "LEAF"}, set(modes)
        self.leaf = np.zeros(n, dtype=bool); self.leaf[flat] = modes == "LEAF"
        self.true_id = np.zeros(n, dtype=np.int64); self.true_id[flat] = self.base[tid] + np.asarray(a["nodes_truenodeids"])
        self.false_id = np.zeros(n, dtype=np.int64); self.false_id[flat] = self.base[tid] + np.asarray(a["nodes_falsenodeids"])
        self.weights = np.zeros((n, self.n_classes), dtype=np.float32)
        cflat = self.base[np.asarray(a["class_treeids"])] + np.asarray(a["class_nodeids"])
        np.add.at(self.weights, (cflat, np.asarray(a["class_ids"])), np.asarray(a["class_weights"], dtype=np.float32))
        assert a.get("post_transform", "NONE") == "NONE"
        self.binary_single = self.n_classes == 2 and len(set(a["class_ids"])) == 1
        assert not self.binary_single or float(np.min(a["class_weights"])) >= 0.0

    def predict_proba(self, X) -> np.ndarray:
        X = np.asarray(X, dtype=np.float32)
        out = np.zeros((len(X), self.n_classes), dtype=np.float32)
        for t in range(self.n_trees):
            cur = np.full(len(X), self.base[t], dtype=np.int64)
            live = ~self.leaf[cur]
            while live.any():
                c = cur[live]
                go_true = X[live, self.feature[c]] <= self.value[c]
                cur[live] = np.where(go_true, self.true_id[c], self.false_id[c])
                live = ~self.leaf[cur]
            out += self.weights[cur]
        if self.binary_single:                                   # onnxruntime's binary case: [1 - s, s]
            sc = out.sum(axis=1)
            out = np.stack([1.0 - sc, sc], axis=1).astype(np.float32)
        return out


class InferenceSession:
    """The part of ``onnxruntime.InferenceSession`` that bicipital_groove.py:178-181 uses."""

    def __init__(self, model_bytes, providers=None):
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".onnx") as f:
            f.write(model_bytes)
            f.flush()
            self.forest = Forest(f.name)

    def run(self, _, feed):
        p = self.forest.predict_proba(feed["X"])
        labels = np.asarray(self.forest.attrs["classlabels_int64s"], dtype=np.int64)[p.argmax(axis=1)]
        return [labels, p]
