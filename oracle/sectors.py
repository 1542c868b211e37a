"""Sector algebra of the symmetry groups HubbardTN uses (oracle; test infrastructure).

Reference: src/HubbardFunctions.jl:245-255 (`SymSpace(P,Q,spin)`) and :341-346
(`SymSpace()`):
    spin=false : I = fZ2 x SU2Irrep x U1Irrep   -> kind SU2U1, label (p, 2j, n)
    spin=true  : I = fZ2 x U1Irrep  x U1Irrep   -> kind U1U1,  label (p, m,  n)  (m = 2 S_z)
    mu models  : I = fZ2 x SU2Irrep             -> kind SU2U1 with n == 0 everywhere
The group theory itself lives in TensorKitSectors 0.1.4 / WignerSymbols 2.0.0
(Manifest.toml:1176,1302; not vendored): restated here from the textbook
definitions (Clebsch-Gordan coefficients by Racah's formula, Condon-Shortley phase).

Fermionic signs are NOT carried by the sectors in this restatement: the MPO carries
explicit Jordan-Wigner parity strings (oracle/hubbard.py), so every tensor is an
ordinary (bosonic) invariant tensor and `p` is just an additive Z2 label.
"""
from __future__ import annotations

from fractions import Fraction
from functools import lru_cache
from math import factorial, sqrt

import numpy as np

SU2U1 = 0
U1U1 = 1
KIND_NAMES = {SU2U1: "fZ2xSU2xU1", U1U1: "fZ2xU1xU1"}


def dim(kind: int, s) -> int:
    """Quantum dimension of sector s (2j+1 for SU(2), 1 for abelian)."""
    return s[1] + 1 if kind == SU2U1 else 1


def dual(kind: int, s):
    if kind == SU2U1:
        return (s[0], s[1], -s[2])
    return (s[0], -s[1], -s[2])


def trivial(kind: int):
    return (0, 0, 0)


def fuse(kind: int, a, b):
    """All sectors c in a (x) b (each appears once: SU(2) is multiplicity free)."""
    p = (a[0] + b[0]) & 1
    n = a[2] + b[2]
    if kind == SU2U1:
        return [(p, tj, n) for tj in range(abs(a[1] - b[1]), a[1] + b[1] + 1, 2)]
    return [(p, a[1] + b[1], n)]


def allowed(kind: int, a, b, c) -> bool:
    """True when c is contained in a (x) b."""
    if ((a[0] + b[0]) & 1) != c[0] or a[2] + b[2] != c[2]:
        return False
    if kind == SU2U1:
        return abs(a[1] - b[1]) <= c[1] <= a[1] + b[1] and ((a[1] + b[1] + c[1]) & 1) == 0
    return a[1] + b[1] == c[1]


def _u1key(n: int):
    # TensorKit orders U(1) charges 0, +1, -1, +2, -2, ... (SURVEY.md App. A [EXT])
    return (abs(n), 0 if n >= 0 else 1)


def sort_key(kind: int, s):
    """Canonical ordering: product sectors compare by reversed tuple (last factor most
    significant); U(1): 0,+1,-1,...; SU(2): by j; Z2: 0<1 (SURVEY.md App. A [EXT])."""
    if kind == SU2U1:
        return (_u1key(s[2]), s[1], s[0])
    return (_u1key(s[2]), _u1key(s[1]), s[0])


@lru_cache(maxsize=None)
def _cg_su2(tj1: int, tj2: int, tj3: int) -> np.ndarray:
    """<j1 m1; j2 m2 | j3 m3> as array [2j1+1, 2j2+1, 2j3+1]; index i <-> m = -j + i."""
    d1, d2, d3 = tj1 + 1, tj2 + 1, tj3 + 1
    out = np.zeros((d1, d2, d3))
    if not (abs(tj1 - tj2) <= tj3 <= tj1 + tj2 and (tj1 + tj2 + tj3) % 2 == 0):
        return out
    f = factorial

    def h(x):  # x is a doubled integer that must be even
        assert x % 2 == 0
        return x // 2

    pref = Fraction(
        d3 * f(h(tj3 + tj1 - tj2)) * f(h(tj3 - tj1 + tj2)) * f(h(tj1 + tj2 - tj3)),
        f(h(tj1 + tj2 + tj3) + 1),
    )
    for i1 in range(d1):
        tm1 = -tj1 + 2 * i1
        for i2 in range(d2):
            tm2 = -tj2 + 2 * i2
            tm3 = tm1 + tm2
            if abs(tm3) > tj3:
                continue
            i3 = (tm3 + tj3) // 2
            rad = pref * (
                f(h(tj3 + tm3)) * f(h(tj3 - tm3)) * f(h(tj1 - tm1)) * f(h(tj1 + tm1))
                * f(h(tj2 - tm2)) * f(h(tj2 + tm2))
            )
            ssum = Fraction(0)
            for k in range(0, tj1 + tj2 + 2):
                a = h(tj1 + tj2 - tj3) - k
                b = h(tj1 - tm1) - k
                c = h(tj2 + tm2) - k
                d = h(tj3 - tj2 + tm1) + k
                e = h(tj3 - tj1 - tm2) + k
                if min(a, b, c, d, e) < 0:
                    continue
                ssum += Fraction((-1) ** k, f(k) * f(a) * f(b) * f(c) * f(d) * f(e))
            out[i1, i2, i3] = float(ssum) * sqrt(float(rad))
    return out


_ONE = np.ones((1, 1, 1))
_ZERO = np.zeros((1, 1, 1))


def cg(kind: int, a, b, c) -> np.ndarray:
    """Invariant coupling tensor a (x) b -> c, shape [dim a, dim b, dim c], normalised
    as an isometry: sum_{ma,mb} cg[ma,mb,mc] cg[ma,mb,mc'] = delta(mc,mc')."""
    if not allowed(kind, a, b, c):
        return np.zeros((dim(kind, a), dim(kind, b), dim(kind, c)))
    if kind == SU2U1:
        return _cg_su2(a[1], b[1], c[1])
    return _ONE
