#!/usr/bin/env python
"""Emit the Runge-Kutta tableaus as C headers (hex-float literals, bit-exact).

Sources of truth (nothing is read from /root/reference):
  * DOP853 (A 16x16, B, C, E3, E5, D 4x16) and DOPRI5/RK45 (A, B, C, E, P):
    SciPy's scipy.integrate._ivp tables.  SURVEY.md section 8c records that the
    reference's vendored hiten/algorithms/integrators/coefficients/dop853.py is
    array_equal to them; tests/golden/make_coeff_check.py re-checks every table
    emitted here against the reference arrays.
  * RK4 (classical), "RK6" (= the 7-stage DOPRI5 tableau)
    and RK8 (Prince-Dormand 8(7)13M, 8th-order weights): the published rationals,
    evaluated as double/double exactly like a Python literal expression would.

Writes  hiten_b200/csrc/hb_coeffs.h  (product, prefix HB_)  and
        oracle/ho_coeffs.h           (test oracle, prefix HO_).
"""
import os
from fractions import Fraction  # noqa: F401  (kept for ad-hoc checks)

import numpy as np
from scipy.integrate._ivp import dop853_coefficients as dop
from scipy.integrate._ivp.rk import RK45

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def q(a, b=1.0):
    return float(a) / float(b)


def rk4():
    A = np.zeros((4, 4))
    A[1, 0] = 0.5
    A[2, 1] = 0.5
    A[3, 2] = 1.0
    B = np.array([q(1, 6), q(1, 3), q(1, 3), q(1, 6)])
    C = np.array([0.0, 0.5, 0.5, 1.0])
    return A, B, C


def rk6():
    # DOPRI5 with the FSAL row kept as a 7th stage (what the reference ships as "RK6").
    A = np.zeros((7, 7))
    A[1, 0] = q(1, 5)
    A[2, :2] = [q(3, 40), q(9, 40)]
    A[3, :3] = [q(44, 45), q(-56, 15), q(32, 9)]
    A[4, :4] = [q(19372, 6561), q(-25360, 2187), q(64448, 6561), q(-212, 729)]
    A[5, :5] = [q(9017, 3168), q(-355, 33), q(46732, 5247), q(49, 176), q(-5103, 18656)]
    A[6, :6] = [q(35, 384), 0.0, q(500, 1113), q(125, 192), q(-2187, 6784), q(11, 84)]
    B = np.array([q(35, 384), 0.0, q(500, 1113), q(125, 192), q(-2187, 6784), q(11, 84), 0.0])
    C = np.array([0.0, q(1, 5), q(3, 10), q(4, 5), q(8, 9), 1.0, 1.0])
    return A, B, C


def rk8():
    # Prince & Dormand (1981) RK8(7)13M, 8th-order solution weights.
    A = np.zeros((13, 13))
    A[1, 0] = q(1, 18)
    A[2, :2] = [q(1, 48), q(1, 16)]
    A[3, :3] = [q(1, 32), 0.0, q(3, 32)]
    A[4, :4] = [q(5, 16), 0.0, q(-75, 64), q(75, 64)]
    A[5, :5] = [q(3, 80), 0.0, 0.0, q(3, 16), q(3, 20)]
    A[6, :6] = [q(29443841, 614563906), 0.0, 0.0, q(77736538, 692538347), q(-28693883, 1125000000),
                q(23124283, 1800000000)]
    A[7, :7] = [q(16016141, 946692911), 0.0, 0.0, q(61564180, 158732637), q(22789713, 633445777),
                q(545815736, 2771057229), q(-180193667, 1043307555)]
    A[8, :8] = [q(39632708, 573591083), 0.0, 0.0, q(-433636366, 683701615), q(-421739975, 2616292301),
                q(100302831, 723423059), q(790204164, 839813087), q(800635310, 3783071287)]
    A[9, :9] = [q(246121993, 1340847787), 0.0, 0.0, q(-37695042795, 15268766246), q(-309121744, 1061227803),
                q(-12992083, 490766935), q(6005943493, 2108947869), q(393006217, 1396673457),
                q(123872331, 1001029789)]
    A[10, :10] = [q(-1028468189, 846180014), 0.0, 0.0, q(8478235783, 508512852), q(1311729495, 1432422823),
                  q(-10304129995, 1701304382), q(-48777925059, 3047939560), q(15336726248, 1032824649),
                  q(-45442868181, 3398467696), q(3065993473, 597172653)]
    A[11, :11] = [q(185892177, 718116043), 0.0, 0.0, q(-3185094517, 667107341), q(-477755414, 1098053517),
                  q(-703635378, 230739211), q(5731566787, 1027545527), q(5232866602, 850066563),
                  q(-4093664535, 808688257), q(3962137247, 1805957418), q(65686358, 487910083)]
    A[12, :12] = [q(403863854, 491063109), 0.0, 0.0, q(-5068492393, 434740067), q(-411421997, 543043805),
                  q(652783627, 914296604), q(11173962825, 925320556), q(-13158990841, 6184727034),
                  q(3936647629, 1978049680), q(-160528059, 685178525), q(248638103, 1413531060), 0.0]
    B = np.array([q(14005451, 335480064), 0.0, 0.0, 0.0, 0.0, q(-59238493, 1068277825),
                  q(181606767, 758867731), q(561292985, 797845732), q(-1041891430, 1371343529),
                  q(760417239, 1151165299), q(118820643, 751138087), q(-528747749, 2220607170), q(1, 4)])
    C = np.array([0.0, q(1, 18), q(1, 12), q(1, 8), q(5, 16), q(3, 8), q(59, 400), q(93, 200),
                  q(5490023248, 9719169821), q(13, 20), q(1201146811, 1299019798), 1.0, 1.0])
    return A, B, C


def tables():
    t = {}
    t["DOP853_A"] = np.array(dop.A, dtype=np.float64)          # 16x16 (rows 12..15 = B row + dense stages)
    t["DOP853_B"] = np.array(dop.B, dtype=np.float64)          # 12
    t["DOP853_C"] = np.array(dop.C, dtype=np.float64)          # 16
    t["DOP853_E3"] = np.array(dop.E3, dtype=np.float64)        # 13
    t["DOP853_E5"] = np.array(dop.E5, dtype=np.float64)        # 13
    t["DOP853_D"] = np.array(dop.D, dtype=np.float64)          # 4x16
    t["RK45_A"] = np.array(RK45.A, dtype=np.float64)           # 6x5
    t["RK45_B"] = np.array(RK45.B, dtype=np.float64)           # 6
    t["RK45_C"] = np.array(RK45.C, dtype=np.float64)           # 6
    t["RK45_E"] = np.array(RK45.E, dtype=np.float64)           # 7
    t["RK45_P"] = np.array(RK45.P, dtype=np.float64)           # 7x4
    for name, fn in (("RK4", rk4), ("RK6", rk6), ("RK8", rk8)):
        A, B, C = fn()
        t[name + "_A"], t[name + "_B"], t[name + "_C"] = A, B, C
    return t


def emit(prefix, path, qualifier):
    t = tables()
    out = [
        "/* GENERATED by tools/gen_coeffs.py -- do not edit.",
        " * Runge-Kutta tableaus as exact hex-float doubles (C99 / C++17).",
        " * DOP853 + RK45 from SciPy's tables (bit-identical to the reference's vendored",
        " * hiten/algorithms/integrators/coefficients/{dop853,rk45}.py); RK4, RK6 (= DOPRI5,",
        " * 7 stages) and RK8 (Prince-Dormand 8(7)13M) from their published rationals",
        " * (reference: coefficients/{rk4,rk6,rk8}.py). */",
        "#pragma once",
        "",
    ]
    for name, arr in t.items():
        if arr.ndim == 1:
            out.append(f"{qualifier} double {prefix}{name}[{arr.shape[0]}] = {{")
            out.append("  " + ", ".join(float(v).hex() for v in arr))
            out.append("};")
        else:
            out.append(f"{qualifier} double {prefix}{name}[{arr.shape[0]}][{arr.shape[1]}] = {{")
            for row in arr:
                out.append("  {" + ", ".join(float(v).hex() for v in row) + "},")
            out.append("};")
        out.append("")
    with open(path, "w") as fh:
        fh.write("\n".join(out))
    print("wrote", path)


if __name__ == "__main__":
    emit("HB_", os.path.join(REPO, "hiten_b200", "csrc", "hb_coeffs.h"), "static constexpr")
    emit("HO_", os.path.join(REPO, "oracle", "ho_coeffs.h"), "static const")
