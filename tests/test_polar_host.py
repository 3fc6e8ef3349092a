"""shb_polar (the unroll's theta / r of a sample, csrc/shb_kernels.cu) restated on the host with the device's own table:
the algorithm's error against atan2l and the exactness of its square root, without a GPU.  The device function itself is
held to numpy on the box (tests/test_gpu_parity.py::test_polar_forms_against_numpy_bits)."""
import ctypes
import re
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
CU = ROOT / "shoulder_b200" / "csrc" / "shb_kernels.cu"


def _table_rows():
    src = CU.read_text()
    body = src[src.index("g_polar_tab[23][4] = {"):]
    body = body[:body.index("};")]
    rows = re.findall(r"\{([^{}]+)\}", body)
    return [[float(v) for v in r.split(",")] for r in rows]


def test_table_is_what_the_generator_makes():
    rows = np.array(_table_rows())
    assert rows.shape == (23, 4)
    ld = np.longdouble
    assert tuple(rows[0]) == (0.0, 1.0, 0.0, 0.0)
    for i in range(1, 23):
        phi = np.float64(np.arcsin(ld(i + 0.5) / ld(32)))
        assert rows[i][2] == phi and rows[i][0] == np.float64(np.sin(ld(phi))) and rows[i][1] == np.float64(np.cos(ld(phi)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not found")
def test_polar_algorithm_error_and_exact_sqrt(tmp_path):
    (tmp_path / "tab.inc").write_text("".join("    {%.20e, %.20e, %.20e, 0.0},\n" % tuple(r[:3]) for r in _table_rows()))
    so = tmp_path / "polar_host.so"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", f"-I{tmp_path}", "-o", str(so), str(ROOT / "tests" / "polar_host.c"), "-lm"],
                   check=True)
    lib = ctypes.CDLL(str(so))
    lib.polar_check.restype = ctypes.c_double
    lib.polar_check.argtypes = [ctypes.c_long, ctypes.c_long, ctypes.POINTER(ctypes.c_long)]
    bad = ctypes.c_long()
    ulp = lib.polar_check(4_000_000, 1, ctypes.byref(bad))
    assert ulp <= 3.0, ulp                    # measured 2.5: rounding of the unit vector, of b sin(phi) and of the last two sums
    assert bad.value == 0                     # r is the correctly rounded root of fl(fl(x^2) + fl(y^2)): numpy's np.sqrt(x**2 + y**2)
