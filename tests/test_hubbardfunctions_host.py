"""Host logic of the product-side mirror of HubbardFunctions.jl (no GPU): the finite-state-machine MPO
+ libhtn's Wigner-Eckart projection give exactly the oracle's reduced MPO (itself pinned on exact
diagonalisation), and the initial bond spaces equal the oracle's restatement of HF:917-959."""
import numpy as np
import pytest

from hubbardtn_b200 import device as dev, hubbardfunctions as hf, sectors as PS
from oracle import mps as M
from oracle import sectors as S
from oracle.hubbard import OB_Sim as OracleSim, mpo as oracle_mpo
from oracle.spaces import initial_bond_spaces


@pytest.mark.parametrize("spin", [False, True])
def test_hamiltonian_matches_oracle_mpo(spin):
    t, u, mu = [1.0, 0.3], [4.0, 0.7], 0.2
    H = hf.hamiltonian(hf.OB_Sim(t=t, u=u, mu=mu, kwargs={"spin": spin}), ctx=None)
    Ws, P, Mlev = oracle_mpo(OracleSim(t=t, u=u, mu=mu, kwargs={"spin": spin}))
    assert H.levels == Mlev.sectors and H.phys == P.sectors and len(H) == len(Ws) == 2
    got, ref = H.W[0].entries(), Ws[0].entries
    assert set(got) == set(ref)
    assert max(abs(got[k] - ref[k]) for k in ref) < 1e-13


def test_dense_projection_rejects_non_invariant_tensor():
    from hubbardtn_b200 import _lib
    sim = hf.OB_Sim(t=[1.0], u=[2.0])
    Wd, levels = hf.hamiltonian_dense(sim)
    Wd = Wd.copy()
    Wd[0, 0, 2, -1] += 0.5          # couples |0> to one member of the doublet only: breaks SU(2) and parity
    P = dev.Legs(None, PS.SU2U1, PS.physical_space(PS.SU2U1, 1, 1))
    Mleg = dev.Legs(None, PS.SU2U1, levels)
    with pytest.raises(_lib.HtnError):
        dev.Mpo.from_dense(None, Mleg, P, Mleg, Wd)


@pytest.mark.parametrize("spin,P,Q", [(False, 1, 1), (True, 1, 1), (False, 1, 2), (False, 3, 2)])
def test_initial_spaces_match_oracle(spin, P, Q):
    sym = PS.U1U1 if spin else PS.SU2U1
    sim = hf.OB_Sim(t=[1.0], u=[5.0], P=P, Q=Q, kwargs={"spin": spin})
    L = sim.unit_cell
    phys = PS.physical_space(sym, P, Q)
    got = hf.initial_spaces(sym, phys, L, P, 50)
    from oracle.spaces import physical_space
    Po = physical_space(sym, P, Q)
    ref = M.trim_spaces(sym, initial_bond_spaces(sym, [Po] * L, P, 50), [Po] * L)
    assert [dict(sorted(g.items())) for g in got] == [dict(sorted(r.as_dict().items())) for r in ref]
