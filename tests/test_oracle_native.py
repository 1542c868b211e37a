"""The native-thread CPU baseline (oracle/native: single-threaded OpenBLAS dgemm per block, one pthread per core) computes
the same H_AC apply as the numpy plan and as the defining triple product (tests the checker / baseline, not the product)."""
import os

import numpy as np

from oracle import sectors as S
from oracle.heff import HeffACPlan, heff_ac_apply_naive
from oracle.native import NativeHeffAC
from oracle.spaces import physical_space, synthetic_bond_space
from oracle.tensors import EnvTensor, Legs, MPOTensor, MPSTensor


def _case(kind, D, seed):
    rng = np.random.default_rng(seed)
    P = physical_space(kind, 1, 1)
    Va, Vb = synthetic_bond_space(kind, D, 0), synthetic_bond_space(kind, D, 1)
    levels = [(0, 0, 0), (1, 1, 1), (1, 1, -1), (0, 2, 0), (0, 0, 0)] if kind == S.SU2U1 else \
        [(0, 0, 0), (1, 1, 1), (1, -1, 1), (1, 1, -1), (0, 0, 0)]
    M = Legs(kind, levels)
    GL = EnvTensor("L", Va, M, identity_levels=[0]).randomize(rng)
    GR = EnvTensor("R", Vb, M, identity_levels=[len(levels) - 1]).randomize(rng)
    W = MPOTensor(M, P, M).randomize(rng)
    x = MPSTensor(Va, P, Vb).randomize(rng)
    return GL, W, GR, x


def test_native_apply_matches_numpy_plan_and_naive():
    for kind, D in ((S.SU2U1, 40), (S.U1U1, 30)):
        GL, W, GR, x = _case(kind, D, 5)
        plan = HeffACPlan(GL, W, GR, x)
        ref = plan.apply(x)
        naive = heff_ac_apply_naive(GL, W, GR, x)
        nat = NativeHeffAC(plan)
        xf = nat.pack_x(x)
        for threads in (1, min(4, os.cpu_count() or 1)):
            yf = nat.apply_flat(xf, np.empty_like(xf), threads)
            y = nat.unpack_y(yf, x)
            for k in x.keys:
                scale = max(np.abs(ref.blocks[k]).max(), 1e-300)
                assert np.abs(y.blocks[k] - ref.blocks[k]).max() / scale < 1e-12
                assert np.abs(y.blocks[k] - naive.blocks[k]).max() / scale < 1e-11
