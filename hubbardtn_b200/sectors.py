"""Host-side sector rules used when ASSEMBLING inputs for the C ABI (pure integers).

Mirrors `SymSpace` of the reference (src/HubbardFunctions.jl:245-255, 341-346).  The
recoupling arithmetic itself is done inside libhtn (csrc/htn_sectors.cpp); this module only
knows labels and fusion rules so that Python callers can build spaces and MPO entry lists.
"""
from __future__ import annotations

SU2U1 = 0   # (parity, 2j, n)      fZ2 x SU2 x U1   (HF:250)
U1U1 = 1    # (parity, 2Sz, n)     fZ2 x U1 x U1    (HF:247)


def dim(sym, s):
    return s[1] + 1 if sym == SU2U1 else 1


def fuse(sym, a, b):
    p, n = (a[0] + b[0]) & 1, a[2] + b[2]
    if sym == SU2U1:
        return [(p, q, n) for q in range(abs(a[1] - b[1]), a[1] + b[1] + 1, 2)]
    return [(p, a[1] + b[1], n)]


def allowed(sym, a, b, c):
    if ((a[0] + b[0]) & 1) != c[0] or a[2] + b[2] != c[2]:
        return False
    if sym == SU2U1:
        return abs(a[1] - b[1]) <= c[1] <= a[1] + b[1] and ((a[1] + b[1] + c[1]) & 1) == 0
    return a[1] + b[1] == c[1]


def physical_space(sym, P=1, Q=1):
    """Physical multiplets at filling P/Q (HF:248, 251): empty, double, single(s)."""
    if sym == SU2U1:
        return [(0, 0, -P), (0, 0, 2 * Q - P), (1, 1, Q - P)]
    return [(0, 0, -P), (0, 0, 2 * Q - P), (1, 1, Q - P), (1, -1, Q - P)]


def mpo_entry_keys(sym, Ml, P, Mr, pairs=None):
    """All symmetry-allowed reduced entries (a, s', s, b, c) of an MPO tensor, optionally
    restricted to level pairs (a,b) in `pairs`."""
    out = []
    for a, ca in enumerate(Ml):
        for sp, csp in enumerate(P):
            for c in fuse(sym, ca, csp):
                for s, cs in enumerate(P):
                    for b, cb in enumerate(Mr):
                        if pairs is not None and (a, b) not in pairs:
                            continue
                        if allowed(sym, cs, cb, c):
                            out.append((a, sp, s, b, c))
    return out
