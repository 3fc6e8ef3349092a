"""Regenerates tests/golden/bones/*.npz from the reference's STL fixtures.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_bones.py

Each ``.npz`` holds the welded mesh of one of the reference's four test bones
(``/root/reference/tests/test_bones/*.stl``): ``vertices`` float32 exactly as stored in the
STL, ``faces`` int32 in file order.  They are *inputs* (the reference pins no outputs).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from shoulder_b200.meshio import read_stl, weld  # noqa: E402

SRC = Path("/root/reference/tests/test_bones")
DST = Path(__file__).resolve().parent / "bones"


def main():
    DST.mkdir(exist_ok=True)
    for stl in sorted(SRC.glob("*.stl")):
        v, f = weld(read_stl(stl))
        assert (v.astype(np.float32).astype(np.float64) == v).all()
        np.savez_compressed(DST / (stl.stem + ".npz"), vertices=v.astype(np.float32), faces=f.astype(np.int32))
        print(stl.name, "T=%d V=%d" % (len(f), len(v)))


if __name__ == "__main__":
    main()
