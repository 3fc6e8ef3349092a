"""Frozen oracle outputs for the reference's four test bones (tests/golden/make_golden.py).
CPU: the oracle still reproduces them.  GPU: the CUDA path reproduces them through the C ABI.
(The reference itself pins no outputs; these are regression pins of the restatement.)"""
from pathlib import Path

import numpy as np
import pytest

import oracle
from shoulder_b200 import meshio

GOLD = Path(__file__).resolve().parent / "golden"
NAMES = ["humerus_left", "humerus_right", "humerus_left_trab", "humerus_left_flipped"]
ARR_TOL = 1e-8      # frame transform goes through BLAS; last-bit differences between hosts are allowed


def _load(name):
    g = np.load(GOLD / f"golden_{name}.npz")
    m = meshio.load_mesh(GOLD / "bones" / f"{name}.npz").apply_transform(g["transform"])
    return g, m


def _face_xor(fi):
    return int(np.bitwise_xor.reduce(np.asarray(fi, dtype=np.int64) * 2654435761 % (1 << 31)))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_its_own_frozen_outputs(name):          # regression pin of the restatement, not a reference vector
    g, m = _load(name)
    s = oracle.OracleSlices(m.vertices, m.faces, g["zs"], g["ixy"].shape[2])
    assert np.array_equal([len(p.metadata["face_index"]) for p in s.paths], g["n_seg"])
    assert np.array_equal([_face_xor(p.metadata["face_index"]) for p in s.paths], g["face_xor"])
    assert np.array_equal(s.n_entities, g["n_ent"])
    for key in ("centroids", "areas1", "ixy", "itr_start", "itr_centered_start"):
        assert np.allclose(getattr(s, key), g[key], rtol=ARR_TOL, atol=ARR_TOL), key


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_the_frozen_oracle_outputs(gpu_backend, name):   # regression pin, not a reference vector
    from shoulder_b200 import _lib
    from helpers import run_gpu
    g, m = _load(name)
    res = run_gpu(m.vertices, m.faces, g["zs"], g["ixy"].shape[2])
    off = res.array(_lib.ARR_SEG_OFF)
    fi = res.array(_lib.ARR_FACE_INDEX)
    assert np.array_equal(res.array(_lib.ARR_N_SEG), g["n_seg"])
    assert np.array_equal([_face_xor(fi[off[i]:off[i + 1]]) for i in range(len(g["zs"]))], g["face_xor"])
    assert np.array_equal([int(fi[off[i]:off[i + 1]].sum()) for i in range(len(g["zs"]))], g["face_sum"])
    assert np.array_equal(res.array(_lib.ARR_N_ENT), g["n_ent"])
    for key, which in (("centroids", _lib.ARR_CENTROID), ("areas1", _lib.ARR_AREA1), ("ixy", _lib.ARR_IXY),
                       ("itr_start", _lib.ARR_ITR_START), ("itr_centered_start", _lib.ARR_ITR_CENTERED_START)):
        assert np.allclose(res.array(which), g[key], rtol=ARR_TOL, atol=ARR_TOL), key
