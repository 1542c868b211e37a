"""Spaces of the hot path: physical space, synthetic bond spaces, initial MPS spaces (oracle).

* `physical_space(kind,P,Q)`            <- SymSpace, src/HubbardFunctions.jl:245-255, 341-346
* `synthetic_bond_space(kind,D,btype)`  <- SURVEY.md 8(d) "Synthetic inputs" (bench shapes)
* `initial_bond_spaces(...)`            <- initialize_mps, src/HubbardFunctions.jl:917-959
"""
from __future__ import annotations

import math

from . import sectors as S
from .tensors import Legs, Space


def physical_space(kind: int, P: int = 1, Q: int = 1, mu_model: bool = False) -> Legs:
    """Physical multiplets.  U(1) charge = Q*occupation - P (HF:248,251); the mu models
    (HF:343) drop the U(1) factor and hold empty+double as (0,0)=>2."""
    if mu_model:
        assert kind == S.SU2U1
        return Legs(kind, [(0, 0, 0), (0, 0, 0), (1, 1, 0)])            # empty, double, single
    if kind == S.SU2U1:
        return Legs(kind, [(0, 0, -P), (0, 0, 2 * Q - P), (1, 1, Q - P)])  # empty, double, single
    return Legs(kind, [(0, 0, -P), (0, 0, 2 * Q - P), (1, 1, Q - P), (1, -1, Q - P)])  # 0,2,up,dn


def fuse_spaces(kind: int, a: dict, b: dict) -> dict:
    out = {}
    for sa, na in a.items():
        for sb, nb in b.items():
            for c in S.fuse(kind, sa, sb):
                out[c] = out.get(c, 0) + na * nb
    return out


def synthetic_bond_space(kind: int, D: int, btype: int = 0) -> Space:
    """Bond space of SURVEY.md 8(d): multiplicities n_c = round(D w_c / sum w), remainder
    added to the largest sector.  btype 0 ('A'): (p=0, integer spin, n even) and (p=1,
    half-integer spin, n odd); btype 1 ('B'): parities of n swapped."""
    w = {}
    if kind == S.SU2U1:
        for tj in range(0, 7):                      # j <= 3
            p = tj & 1
            for n in range(-4, 5):
                if ((n & 1) == p) != (btype == 0):
                    continue
                j = tj / 2.0
                w[(p, tj, n)] = (tj + 1) * math.exp(-(j + 0.5) ** 2 / 2.0) * math.exp(-n * n / (2 * 1.2 ** 2))
    else:
        for m in range(-5, 6):
            p = m & 1
            for n in range(-4, 5):
                if ((n & 1) == p) != (btype == 0):
                    continue
                w[(p, m, n)] = math.exp(-m * m / (2 * 1.3 ** 2)) * math.exp(-n * n / (2 * 1.2 ** 2))
    tot = sum(w.values())
    mult = {s: int(round(D * v / tot)) for s, v in w.items()}
    mult = {s: n for s, n in mult.items() if n > 0}
    if not mult:                       # D too small for any rounded multiplicity
        mult = {max(w, key=lambda s: w[s]): 0}
    big = max(mult, key=lambda s: (mult[s], -abs(s[2]), -abs(s[1])))
    mult[big] += D - sum(mult.values())
    return Space(kind, mult)


def infimum(a: dict, b: dict) -> dict:
    return {s: min(n, b[s]) for s, n in a.items() if s in b and min(n, b[s]) > 0}


def initial_bond_spaces(kind: int, phys: list, P: int, bond_dim: int) -> list:
    """Virtual spaces of the random initial state, HF:917-959.  phys[i] is the Legs of site
    i (i = 0..L-1).  Returns V[i] = right bond of site i (V[L-1] is also the left bond of
    site 0).  V_right = cumulative fuse from the left, V_left = cumulative fuse of duals from
    the right (circularly shifted), V = infimum, capped per sector at bond_dim inside the
    window p in {0,1}, j <= 3 (|m| <= L), |n| <= L*P (HF:931-947)."""
    L = len(phys)

    def as_dict(legs):
        d = {}
        for s in legs.sectors:
            d[s] = d.get(s, 0) + 1
        return d

    pd = [as_dict(p) for p in phys]
    v_right = []
    acc = None
    for i in range(L):
        acc = dict(pd[i]) if acc is None else fuse_spaces(kind, acc, pd[i])
        v_right.append(acc)
    # accumulate(fuse, dual.(Ps); init=one) -> L entries: dual(P1), dual(P1)(x)dual(P2), ...
    v_l = []
    acc = {S.trivial(kind): 1}
    for i in range(L):
        acc = fuse_spaces(kind, acc, {S.dual(kind, s): n for s, n in pd[i].items()})
        v_l.append(acc)
    v_left = list(reversed(v_l))
    v_left = v_left[1:] + v_left[:1]            # HF:926 (= circshift(V_left, 1) as written there)
    out = []
    for i in range(L):
        v = infimum(v_left[i], v_right[i])
        capped = {}
        for s, n in v.items():
            if kind == S.SU2U1:
                inside = s[1] <= 6 and abs(s[2]) <= L * P
            else:
                inside = abs(s[1]) <= L and abs(s[2]) <= L * P
            if inside or s == S.trivial(kind):
                capped[s] = min(n, bond_dim + (1 if s == (0, 0, 0) else 0)) if inside else 1   # HF:935/944
        out.append(Space(kind, capped))
    return out
