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
