"""Oracle pins that need no GPU: every block-sparse contraction of oracle/heff.py equals its
dense (symmetry-free) expansion, and the Clebsch-Gordan tensors match sympy's."""
import numpy as np
import pytest

from oracle import sectors as S
from oracle.heff import (HeffACPlan, heff_ac_apply_dense, heff_ac_apply_naive, heff_c_apply,
                         heff_c_apply_dense, transfer_left, transfer_left_dense, transfer_right,
                         transfer_right_dense)
from oracle.spaces import physical_space, synthetic_bond_space
from oracle.tensors import BondTensor, EnvTensor, Legs, MPOTensor, MPSTensor, inner

LEVELS = {
    S.SU2U1: [(0, 0, 0), (1, 1, 1), (1, 1, -1), (0, 2, 0), (0, 0, 2), (0, 2, -2), (0, 0, 0)],
    S.U1U1: [(0, 0, 0), (1, 1, 1), (1, -1, -1), (0, 2, 0), (0, 0, 2), (0, -2, -2), (0, 0, 0)],
}


def _case(kind, D=12, seed=1):
    rng = np.random.default_rng(seed)
    P = physical_space(kind, 1, 1)
    Va, Vb = synthetic_bond_space(kind, D, 0), synthetic_bond_space(kind, D, 1)
    M = Legs(kind, LEVELS[kind])
    GL = EnvTensor("L", Va, M, identity_levels=[0]).randomize(rng)
    GR = EnvTensor("R", Vb, M, identity_levels=[6]).randomize(rng)
    W = MPOTensor(M, P, M).randomize(rng)
    x = MPSTensor(Va, P, Vb).randomize(rng)
    return rng, P, Va, Vb, M, GL, GR, W, x


def test_cg_matches_sympy():
    from sympy import Rational
    from sympy.physics.quantum.cg import CG
    for tj1, tj2 in [(1, 1), (2, 1), (2, 2), (3, 2), (4, 1), (6, 1), (4, 2)]:
        for tj3 in range(abs(tj1 - tj2), tj1 + tj2 + 1, 2):
            g = S._cg_su2(tj1, tj2, tj3)
            for i1 in range(tj1 + 1):
                for i2 in range(tj2 + 1):
                    tm1, tm2 = -tj1 + 2 * i1, -tj2 + 2 * i2
                    tm3 = tm1 + tm2
                    if abs(tm3) > tj3:
                        continue
                    ref = float(CG(Rational(tj1, 2), Rational(tm1, 2), Rational(tj2, 2), Rational(tm2, 2),
                                   Rational(tj3, 2), Rational(tm3, 2)).doit())
                    assert abs(g[i1, i2, (tm3 + tj3) // 2] - ref) < 1e-13


def test_cg_is_isometry():
    for tj1, tj2 in [(1, 1), (2, 1), (3, 1), (4, 2), (6, 1)]:
        for tj3 in range(abs(tj1 - tj2), tj1 + tj2 + 1, 2):
            g = S._cg_su2(tj1, tj2, tj3)
            assert np.allclose(np.einsum("abc,abd->cd", g, g), np.eye(tj3 + 1), atol=1e-13)


@pytest.mark.parametrize("kind", [S.SU2U1, S.U1U1])
def test_heff_ac_equals_dense(kind):
    rng, P, Va, Vb, M, GL, GR, W, x = _case(kind)
    y = heff_ac_apply_naive(GL, W, GR, x)
    yd = heff_ac_apply_dense(GL, W, GR, x)
    assert np.abs(y.to_dense() - yd).max() < 1e-11 * max(1.0, np.abs(yd).max())
    # staged plan (the algorithm of the CUDA path) == defining triple products
    plan = HeffACPlan(GL, W, GR, x)
    y2 = plan.apply(x)
    for k in y.blocks:
        assert np.abs(y.blocks[k] - y2.blocks[k]).max() < 1e-11
    assert plan.flops == plan.flops_L + plan.flops_R > 0
    # weighted inner product == dense inner product
    assert abs(inner(x, y) - np.vdot(x.to_dense(), yd)) < 1e-9 * abs(np.vdot(x.to_dense(), yd))


@pytest.mark.parametrize("kind", [S.SU2U1, S.U1U1])
def test_transfers_equal_dense(kind):
    rng, P, Va, Vb, M, GL, GR, W, x = _case(kind, seed=3)
    A = MPSTensor(Va, P, Vb).randomize(rng)
    B = MPSTensor(Va, P, Vb).randomize(rng)
    gl = transfer_left(GL, W, A, B)
    assert np.abs(gl.to_dense() - transfer_left_dense(GL, W, A, B)).max() < 1e-11
    gr = transfer_right(GR, W, A, B)
    assert np.abs(gr.to_dense() - transfer_right_dense(GR, W, A, B)).max() < 1e-11


@pytest.mark.parametrize("kind", [S.SU2U1, S.U1U1])
def test_heff_c_equals_dense(kind):
    rng, P, Va, Vb, M, GL, GR, W, x = _case(kind, seed=5)
    GL2 = EnvTensor("L", Vb, M).randomize(rng)
    C = BondTensor(Vb)
    for c in C.blocks:
        C.blocks[c] = rng.standard_normal(C.blocks[c].shape)
    y = heff_c_apply(GL2, GR, C)
    assert np.abs(y.to_dense() - heff_c_apply_dense(GL2, GR, C)).max() < 1e-11


def test_synthetic_space_matches_survey_appendix_c():
    """D=1024 SU2xU1 -> 28 / 25 sectors, largest block 167, D_full 2624."""
    Va, Vb = synthetic_bond_space(S.SU2U1, 1024, 0), synthetic_bond_space(S.SU2U1, 1024, 1)
    assert (len(Va), len(Vb)) == (28, 25)
    assert max(Va.mult) == 167 and Va.red_dim == 1024 and Vb.red_dim == 1024
    assert Va.full_dim == 2624
    assert Va.as_dict()[(0, 2, 0)] == 167 and Va.as_dict()[(0, 0, 0)] == 154
