"""CPU-side checks of the drop-in boundary: libhtn.so loads, exports every symbol that
include/htn.h declares, and its host-side recoupling coefficients equal the oracle's."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import HAVE_GPU, ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "htn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(htn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hubbardtn_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True)
    exported = set(re.findall(r" T (htn_[a-z0-9_]+)", out.stdout))
    declared = _declared()
    assert len(declared) >= 25
    missing = [s for s in declared if s not in exported]
    assert not missing, "declared in include/htn.h but not exported: %s" % missing
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and header disagree"


def test_version_and_error_string():
    from hubbardtn_b200 import _lib
    assert _lib.lib.htn_version() >= 100
    assert isinstance(_lib.last_error(None), str)


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-device error path")
def test_no_cpu_fallback():
    from hubbardtn_b200 import _lib, device
    with pytest.raises(_lib.HtnError) as ei:
        device.Context(0)
    assert ei.value.code == _lib.HTN_ERR_NO_DEVICE


def test_network_coefficients_match_oracle():
    """C++ (Racah loops) vs numpy (exact-rational CG + einsum): independent implementations."""
    from hubbardtn_b200 import device
    from oracle import heff, sectors as S
    rng = np.random.default_rng(0)
    phys = [(0, 0, -1), (0, 0, 1), (1, 1, 0)]
    levels = [(0, 0, 0), (1, 1, 1), (1, 1, -1), (0, 2, 0), (0, 0, 2), (0, 2, -2), (0, 4, 0), (1, 3, 1)]
    checked = nonzero = 0
    for _ in range(600):
        l = (int(rng.integers(0, 2)), int(rng.integers(0, 7)), int(rng.integers(-3, 4)))
        s, sp = phys[rng.integers(0, 3)], phys[rng.integers(0, 3)]
        a, b = levels[rng.integers(0, len(levels))], levels[rng.integers(0, len(levels))]
        for lp in S.fuse(S.SU2U1, a, l):
            for r in S.fuse(S.SU2U1, l, s):
                for c in S.fuse(S.SU2U1, a, sp):
                    for rp in S.fuse(S.SU2U1, b, r):
                        ref = heff.network(S.SU2U1, lp, sp, rp, l, s, r, a, b, c)
                        got = device.network_coefficient(0, [lp, sp, rp, l, s, r, a, b, c])
                        assert abs(got - ref) < 1e-12
                        checked += 1
                        nonzero += ref != 0.0
    assert checked > 300 and nonzero > 30
    # abelian: 1 when every vertex is allowed
    assert device.network_coefficient(1, [(1, 1, 1), (1, 1, 0), (0, 2, 1), (0, 0, 0), (1, 1, 0), (1, 1, 0),
                                          (1, 1, 1), (1, 1, 1), (0, 2, 1)]) == 1.0


def test_small_hessenberg_eigensolver_matches_numpy():
    """Host part of the Arnoldi fixed-point solver (gauge fixing): complex single-shift QR + inverse
    iteration against numpy.linalg.eig, including clustered and complex sub-dominant spectra."""
    import ctypes as C
    from hubbardtn_b200 import _lib
    rng = np.random.default_rng(3)
    pd = C.POINTER(C.c_double)
    for m in (1, 2, 5, 17, 30):
        for trial in range(6):
            # spectrum: dominant real eigenvalue, a close real neighbour, complex pairs inside the disc
            lam = np.zeros(m, dtype=complex)
            lam[0] = 0.93 + 0.07 * rng.random()
            for i in range(1, m):
                lam[i] = (0.999 * lam[0] if i == 1 and trial % 2 else 0.9 * rng.random() * np.exp(2j * np.pi * rng.random()))
            if m > 2 and m % 2 == 0:
                lam[-1] = 0.5 * rng.random()
            # real matrix with that spectrum (complex pairs as 2x2 rotations), random similarity, then Hessenberg
            B = np.zeros((m, m))
            i = 0
            while i < m:
                if abs(lam[i].imag) > 0 and i + 1 < m and i >= 2:
                    a, b = lam[i].real, lam[i].imag
                    B[i:i + 2, i:i + 2] = [[a, b], [-b, a]]
                    lam[i + 1] = np.conj(lam[i])
                    i += 2
                else:
                    lam[i] = lam[i].real if i else lam[i]
                    B[i, i] = lam[i].real
                    i += 1
            S_ = rng.standard_normal((m, m)) + 2 * np.eye(m)
            A = S_ @ B @ np.linalg.inv(S_)
            from scipy.linalg import hessenberg
            H = np.ascontiguousarray(hessenberg(A))
            ev, evec = np.linalg.eig(H)
            k = int(np.argmax(np.abs(ev)))
            theta, y = C.c_double(), np.zeros(m)
            rc = _lib.lib.htn_test_hessenberg_dominant(m, H.ctypes.data_as(pd), C.byref(theta), y.ctypes.data_as(pd))
            assert rc == 0
            assert abs(theta.value - ev[k].real) < 1e-9 * abs(ev[k])
            ref = np.real(evec[:, k])
            ref /= np.linalg.norm(ref)
            if ref @ y < 0:
                ref = -ref
            assert np.abs(ref - y).max() < 1e-6
