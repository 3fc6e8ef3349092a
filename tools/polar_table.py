"""Generates g_polar_tab of csrc/shb_kernels.cu (shb_polar): entry i holds sin, cos and the value of phi_i with sin phi_i =
(i + 1/2) / 32 (entry 0 is the identity), computed in 80-bit long double and rounded to double.
    python tools/polar_table.py"""
import numpy as np
ld = np.longdouble
K = 32
rows = []
for i in range(23):
    if i == 0:
        rows.append((0.0, 1.0, 0.0)); continue
    phi = np.float64(np.arcsin(ld(i + 0.5) / ld(K)))
    S = np.float64(np.sin(ld(phi))); C = np.float64(np.cos(ld(phi)))
    rows.append((float(S), float(C), float(phi)))
for r in rows:
    print("    {%.20e, %.20e, %.20e, 0.0}," % r)
