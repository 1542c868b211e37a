"""Oracle pins for the ground-state path (no GPU): MPO == exact diagonalisation, Krylov solvers,
gauge fixing, and end-to-end VUMPS energies against the reference's golden values
(tests/golden/reference_energies.json <- /root/reference/test/OB.jl:21,44, test/Spin.jl:42)
and the exact Lieb-Wu energies."""
import json
import os

import numpy as np
import pytest

from oracle import mps as M
from oracle import sectors as S
from oracle.hubbard import OB_Sim, dense_mpo, ed_hamiltonian, mpo, mpo_to_hamiltonian
from oracle.krylov import gmres, lanczos_lowest
from oracle.spaces import physical_space, synthetic_bond_space
from oracle.tensors import BondTensor, MPSTensor, inner

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_energies.json")))


@pytest.mark.parametrize("kind", [S.SU2U1, S.U1U1])
def test_mpo_equals_exact_hamiltonian(kind):
    t, u, mu = [1.0, 0.3], [4.0, 0.7], 0.2
    Wd, Mlev = dense_mpo(kind, t, u, mu)
    for N in (2, 3, 4):
        H = mpo_to_hamiltonian(Wd, Mlev, N)
        assert np.abs(H - ed_hamiltonian(kind, N, t, u, mu)).max() < 1e-12
    # the reduced form re-expands to the dense tensor (from_dense raises otherwise)
    Ws, P, Mr = mpo(OB_Sim(t=t, u=u, mu=mu, kwargs={"spin": kind == S.U1U1}))
    assert np.abs(Ws[0].to_dense() - Wd).max() < 1e-12


class _Vec:
    def __init__(self, v):
        self.blocks = {0: v}

    def weight(self, k):
        return 1

    def copy(self):
        return _Vec(self.blocks[0].copy())


def test_krylov_solvers_on_dense_matrices():
    rng = np.random.default_rng(0)
    n = 60
    A = rng.standard_normal((n, n))
    A = A + A.T
    ev, x, info = lanczos_lowest(lambda v: _Vec(A @ v.blocks[0]), _Vec(rng.standard_normal(n)), tol=1e-11, maxiter=50)
    assert info["converged"] and abs(ev - np.linalg.eigvalsh(A)[0]) < 1e-9
    B = np.eye(n) + 0.3 * rng.standard_normal((n, n)) / np.sqrt(n)
    b = rng.standard_normal(n)
    sol, info = gmres(lambda v: _Vec(B @ v.blocks[0]), _Vec(b), tol=1e-12)
    assert info["converged"] and np.abs(B @ sol.blocks[0] - b).max() < 1e-9


@pytest.mark.parametrize("kind", [S.SU2U1, S.U1U1])
def test_gauge_fixing_against_dense(kind):
    rng = np.random.default_rng(2)
    P = physical_space(kind)
    spaces = M.trim_spaces(kind, [synthetic_bond_space(kind, 14, 1), synthetic_bond_space(kind, 14, 0)], [P, P])
    A = MPSTensor(spaces[1], P, spaces[0]).randomize(rng)
    Q, R = M.left_orth(A)
    Qd = Q.to_dense()
    n = Qd.shape[2]
    assert np.abs(np.einsum("lsr,lsq->rq", Qd, Qd) - np.eye(n)).max() < 1e-12       # left isometry (dense)
    assert np.abs(M.mul_right(Q, R).to_dense() - A.to_dense()).max() < 1e-12
    assert all((np.diag(b) > 0).all() for b in R.blocks.values())
    Lm, Q2 = M.right_orth(A)
    Q2d = Q2.to_dense()
    assert np.abs(np.einsum("lsr,msr->lm", Q2d, Q2d) - np.eye(Q2d.shape[0])).max() < 1e-12  # right isometry
    assert np.abs(M.mul_left(Lm, Q2).to_dense() - A.to_dense()).max() < 1e-12
    st = M.random_state(kind, spaces, [P, P], rng)
    for i in range(2):
        lhs = M.mul_right(st["AL"][i], st["C"][i]).to_dense()
        rhs = M.mul_left(st["C"][i - 1], st["AR"][i]).to_dense()
        assert np.abs(lhs - rhs).max() < 1e-9


def _run(kind, u, D, tol, maxiter):
    sim = OB_Sim(t=[1.0], u=[u], kwargs={"spin": kind == S.U1U1})
    Ws, P, _ = mpo(sim)
    spaces = M.trim_spaces(kind, [synthetic_bond_space(kind, D, 1), synthetic_bond_space(kind, D, 0)], [P, P])
    st = M.random_state(kind, spaces, [P, P], np.random.default_rng(1))
    st, envs, eps, log = M.vumps(st, Ws, tol=tol, maxiter=maxiter)
    return st, envs, eps


HALF = [i for i, r in enumerate(GOLD["reference"]) if r["P"] == r["Q"]]      # fixed synthetic spaces: half filling only


@pytest.mark.parametrize("idx", HALF)
def test_vumps_energy_matches_reference_golden(idx):
    """Same comparison the reference's tests make (E/site vs hard-coded value, their atol);
    the variational energy must also sit above the exact Lieb-Wu value and within 1e-2 of it."""
    g = GOLD["reference"][idx]
    kind = S.U1U1 if g["spin"] else S.SU2U1
    st, envs, eps = _run(kind, g["u"][0], 16, 1e-5, 40)
    E = envs.energy_per_site
    assert abs(E - g["E"]) < g["atol"], (g["cite"], E, g["E"])
    exact = GOLD["lieb_wu"][str(int(g["u"][0]))]
    assert exact - 1e-9 < E < exact + 1e-2
    assert abs(envs.energy_cell_left - envs.energy_cell_right) < 1e-6
    # filling is conserved exactly by the symmetry (test/OB.jl:99: sum(density)/2 ~ P/Q)
    vals = [0, 2, 1] if kind == S.SU2U1 else [0, 2, 1, 1]
    n = [M.expval_diag(st["AC"][i], vals) for i in range(2)]
    assert abs(sum(n) / 2 - 1.0) < 1e-8


def test_vumps_converges_tightly_gapped():
    st, envs, eps = _run(S.SU2U1, 8.0, 10, 1e-10, 80)
    assert eps < 1e-10
    exact = GOLD["lieb_wu"]["8"]
    assert exact < envs.energy_per_site < exact + 2e-3
    # H_AC eigen-equation holds at the fixed point: AC is an eigenvector of H_AC
    from oracle.heff import HeffACPlan
    sim = OB_Sim(t=[1.0], u=[8.0])
    Ws, P, _ = mpo(sim)
    plan = HeffACPlan(envs.GL[0], Ws[0], envs.GR[0], st["AC"][0])
    y = plan.apply(st["AC"][0])
    lam = inner(st["AC"][0], y)
    res = np.sqrt(sum(st["AC"][0].weight(k) * np.sum((y.blocks[k] - lam * st["AC"][0].blocks[k]) ** 2) for k in y.blocks))
    assert res < 1e-8
    spec = M.entanglement_spectrum(st["C"][0])
    tot = sum((k[1] + 1) * np.sum(v ** 2) for k, v in spec.items())
    assert abs(tot - 1.0) < 1e-12
