#!/usr/bin/env python
"""Writes tests/golden/reference_energies.json.

Two kinds of pins for the ground-state path:
  * `reference`: the energies hard-coded in the reference's own tests (copied VALUES with
    file:line, read from /root/reference at generation time to make sure they are still there);
  * `lieb_wu`: exact Bethe-ansatz energies per site of the half-filled one-band Hubbard model,
    E/N = -4 t int_0^inf J0(w) J1(w) / (w (1 + exp(w U / 2t))) dw, evaluated with scipy.
Run in the build container (needs /root/reference); the GPU box only reads the JSON.
"""
import json
import os
import re

import numpy as np
from scipy import integrate, special

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/test"


def lieb_wu(U, t=1.0):
    if U == 0:
        return -4.0 * t / np.pi
    f = lambda w: special.j0(w) * special.j1(w) / (w * (1.0 + np.exp(0.5 * w * U / t)))  # noqa: E731
    val, _ = integrate.quad(f, 0, np.inf, limit=2000)
    return -4.0 * t * val


def grab(path, line, expect):
    txt = open(os.path.join(REF, path)).read().splitlines()[line - 1]
    assert re.search(re.escape(expect), txt), (path, line, txt)
    return float(expect)


ref = [
    dict(cite="test/OB.jl:21 (U=0)", model="OB", t=[1.0], u=[0.0], P=1, Q=1, spin=False, E=grab("OB.jl", 21, "-1.2696767"), atol=1e-2),
    dict(cite="test/OB.jl:21 (U=1)", model="OB", t=[1.0], u=[1.0], P=1, Q=1, spin=False, E=grab("OB.jl", 21, "-1.037173"), atol=1e-2),
    dict(cite="test/OB.jl:21 (U=2)", model="OB", t=[1.0], u=[2.0], P=1, Q=1, spin=False, E=grab("OB.jl", 21, "-0.84163698"), atol=1e-2),
    dict(cite="test/OB.jl:44 (P/Q=1)", model="OB", t=[1.0], u=[5.0], P=1, Q=1, spin=False, E=grab("OB.jl", 44, "-0.48460447"), atol=1e-2),
    dict(cite="test/OB.jl:44 (P/Q=1/2)", model="OB", t=[1.0], u=[5.0], P=1, Q=2, spin=False, E=grab("OB.jl", 44, "-0.73920032"), atol=1e-2),
    dict(cite="test/OB.jl:44 (P/Q=3/2)", model="OB", t=[1.0], u=[5.0], P=3, Q=2, spin=False, E=grab("OB.jl", 44, "1.76073968"), atol=1e-2),
    dict(cite="test/Spin.jl:42", model="OB", t=[1.0], u=[8.0], P=1, Q=1, spin=True, E=grab("Spin.jl", 42, "-0.32637"), atol=1e-1),
]
ref_mb = [
    # test/MB.jl:24-35: two uncoupled bands, t_IS = 1, U = 3, bond_dim 20, svalue 2.0; compared with atol = tol = 1e-1 (MB.jl:12,65)
    dict(cite="test/MB.jl:59", model="MB", t=[[0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]],
         u=[[3.0, 0.0, 0.0, 0.0], [0.0, 3.0, 0.0, 0.0]], P=1, Q=1, spin=False, bond_dim=20,
         E=grab("MB.jl", 59, "-0.630375296"), atol=1e-1),
    # test/Spin.jl:20-30,49: the same two-band model with spin=true (U(1)xU(1)); atol = tol = 1e-1 (Spin.jl:12)
    dict(cite="test/Spin.jl:49", model="MB", t=[[0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]],
         u=[[3.0, 0.0, 0.0, 0.0], [0.0, 3.0, 0.0, 0.0]], P=1, Q=1, spin=True, bond_dim=20,
         E=grab("Spin.jl", 49, "-0.63093"), atol=1e-1),
]
out = dict(
    note="reference: values hard-coded in DaanVrancken/HubbardTN tests (truncation-limited, loose atol); "
         "lieb_wu: exact half-filling energies per site (scipy quad of the Lieb-Wu integral)",
    reference=ref,
    reference_mb=ref_mb,
    lieb_wu={str(U): lieb_wu(U) for U in (0, 1, 2, 3, 5, 6, 8)},
)
with open(os.path.join(ROOT, "tests", "golden", "reference_energies.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out["lieb_wu"], indent=1))
