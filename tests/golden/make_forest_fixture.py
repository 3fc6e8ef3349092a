"""tests/golden/forest_rfc_bg3.npz: the node arrays of the reference's groove forest (``src/shoulder/humerus/models/
rfc_bg3.onnx``, a data asset like the test bones) as ``shoulder_b200.features.flatten_tree_ensemble`` lays them out, so that
the ``-m gpu`` tests can run the device forest on the GPU box, which has no /root/reference.

    python tests/golden/make_forest_fixture.py
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from shoulder_b200 import features  # noqa: E402

if __name__ == "__main__":
    a = features.flatten_tree_ensemble(features.read_onnx_tree_ensemble("/root/reference/src/shoulder/humerus/models/rfc_bg3.onnx"))
    np.savez_compressed(HERE / "forest_rfc_bg3.npz", **a)
    print({k: v.shape for k, v in a.items()})
