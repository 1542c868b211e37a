"""Host logic of the multi-band MPO builder (hubbardfunctions.MB_Sim / fsm_mpo_dense / mb_terms, mirroring
HF:477-561, 645-659, 811-910): the finite-state-machine MPO contracted over an open chain must equal the
Hamiltonian written down directly with Jordan-Wigner matrices in the 4^N Fock space.  CPU only."""
import numpy as np
import pytest

from hubbardtn_b200 import hubbardfunctions as hf, sectors as PS


def chain_from_mpo(Ws, N):
    """Sum of all MPO terms that open and close inside an N-site open chain (site p uses Ws[p % L])."""
    chi = Ws[0].shape[0]
    cur = {0: np.eye(1)}                       # level -> operator on the sites so far
    for p in range(N):
        W = Ws[p % len(Ws)]
        nxt = {}
        for a, op in cur.items():
            for b in range(chi):
                w = W[a, :, :, b]
                if not np.any(w):
                    continue
                term = np.kron(op, w)
                nxt[b] = nxt[b] + term if b in nxt else term
        cur = nxt
    return cur[chi - 1]


def jw_ops(sym, N):
    """c_{p,up}, c_{p,dn} on the N-site chain in the local basis / ordering conventions of _fermion_ops."""
    cu, cd, par, num, dbl = hf._fermion_ops(sym)
    ops = []
    for p in range(N):
        def emb(o):
            m = np.eye(1)
            for q in range(N):
                m = np.kron(m, par if q < p else (o if q == p else np.eye(4)))
            return m
        ops.append((emb(cu), emb(cd)))
    return ops


def direct_hamiltonian(sim, N):
    B = sim.bands
    c = jw_ops(sim.sym, N)
    dim = 4 ** N
    n = [c[p][0].T @ c[p][0] + c[p][1].T @ c[p][1] for p in range(N)]
    d = [c[p][0].T @ c[p][0] @ c[p][1].T @ c[p][1] for p in range(N)]
    H = np.zeros((dim, dim))
    t, u = sim.t, sim.u
    J = sim.J

    def exch(p, q, j):
        """j [ sum_ss' c+_ps c+_qs' c_ps' c_qs  +  (c+_p,up c+_p,dn c_q,dn c_q,up + h.c.) ]  written with the Jordan-Wigner
        matrices themselves (the operators HF:580/675 and HF:604/690 are named after; docstring of hf.mb_terms)"""
        if j == 0.0 or not (0 <= p < N and 0 <= q < N):
            return
        for s1 in (0, 1):
            for s2 in (0, 1):
                H[...] += j * (c[p][s1].T @ c[q][s2].T @ c[p][s2] @ c[q][s1])
        ph = c[p][0].T @ c[p][1].T @ c[q][1] @ c[q][0]
        H[...] += j * (ph + ph.T)

    def hop(p, q, amp):
        if 0 <= p < N and 0 <= q < N:
            for s in (0, 1):
                H[...] += amp * (c[q][s].T @ c[p][s])

    for p in range(N):
        b = p % B
        H += u[b, b] * d[p] - t[b, b] * n[p]                                     # HF:531-551
    for cell in range(-(N // B) - 4, N // B + 4):
        for bi in range(B):
            for bf in range(B):
                pi, pf = cell * B + bi, cell * B + bf
                if bi != bf and 0 <= pi < N and 0 <= pf < N:
                    hop(pi, pf, -t[bi, bf])                                       # HF:499 -t cdc{(bf,site),(bi,site)}
                    if bi > bf:
                        H += 0.5 * (u[bi, bf] + u[bf, bi]) * n[pi] @ n[pf]        # HF:548-560
                        exch(pi, pf, 0.5 * (J[bi, bf] + J[bf, bi]))               # HF:563-616
                for k in range(1, t.shape[1] // B):
                    pf2 = (cell + k) * B + bf
                    if 0 <= pi < N and 0 <= pf2 < N:
                        hop(pi, pf2, -t[bi, k * B + bf])                          # HF:518 twosite = cdc + cdc'
                        hop(pf2, pi, -t[bi, k * B + bf])
                for k in range(1, u.shape[1] // B):
                    pf2 = (cell + k) * B + bf
                    if 0 <= pi < N and 0 <= pf2 < N:
                        H += u[bi, k * B + bf] * n[pi] @ n[pf2]                    # HF:658
                for k in range(1, J.shape[1] // B):
                    exch(pi, (cell + k) * B + bf, J[bi, k * B + bf])              # HF:668-700
    return H


@pytest.mark.parametrize("spin", [False, True])
def test_fsm_mpo_equals_direct_sum(spin):
    t = np.array([[0.3, 0.1, 1.0, 0.5], [0.1, -0.2, 0.25, 0.8]])
    u = np.array([[3.0, 0.7, 0.25, 0.1], [0.5, 2.0, 0.0, 0.4]])
    sim = hf.MB_Sim(t, u, P=1, Q=1, kwargs={"spin": spin})
    assert sim.unit_cell == 4 and sim.bands == 2
    Ws, levels = hf.fsm_mpo_dense(sim.sym, sim.Q, *hf.mb_terms(sim))
    N = 5
    Hm, Hd = chain_from_mpo(Ws, N), direct_hamiltonian(sim, N)
    assert np.abs(Hd - Hd.T).max() < 1e-13
    assert np.abs(Hm - Hd).max() < 1e-12
    assert levels[0] == levels[-1] == (0, 0, 0)


def test_decoupled_bands_are_the_one_band_chain():
    """test/MB.jl:24-35: two uncoupled bands with t_IS = 1, U = 3 are two interleaved one-band chains, i.e.
    OB_Sim(t=[0,1], u=[3]) on the same sites."""
    t = np.array([[0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]])
    u = np.array([[3.0, 0.0, 0.0, 0.0], [0.0, 3.0, 0.0, 0.0]])
    mb = hf.MB_Sim(t, u, np.zeros((2, 2)))
    Ws, _ = hf.fsm_mpo_dense(mb.sym, mb.Q, *hf.mb_terms(mb))
    W1, _ = hf.hamiltonian_dense(hf.OB_Sim(t=[0.0, 1.0], u=[3.0]))
    N = 5
    assert np.abs(chain_from_mpo(Ws, N) - chain_from_mpo([W1], N)).max() < 1e-12


def test_unmirrored_terms_raise():
    t = np.array([[0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]])
    u = np.array([[3.0, 0.0], [0.0, 3.0]])
    with pytest.raises(NotImplementedError):
        hf.mb_terms(hf.MB_Sim(t, u, None, np.array([[0.0, 0.5], [0.5, 0.0]])))          # U13 / U_ijjj
    with pytest.raises(NotImplementedError):
        hf.mb_terms(hf.MB_Sim(t, u, kwargs={"U112": {(1, 1, 2, 3): 0.1}}))
    with pytest.raises(ValueError):
        hf.mb_terms(hf.MB_Sim(t, u, np.array([[0.3, 0.5], [0.5, 0.0]])))                # on-band J (HF:575)
    with pytest.raises(ValueError):
        hf.MB_Sim(t, np.eye(3))


@pytest.mark.parametrize("spin", [False, True])
def test_exchange_terms_equal_the_second_quantised_operators(spin):
    """HF:563-616 / 668-700: on-site and inter-site exchange (Hund's coupling -2J S.S - J/2 nn) and pair hopping through the
    triplet and charge-+-2 levels of the finite-state machine == the four-fermion operators written with Jordan-Wigner
    matrices; the dense MPO tensors are invariant (the library's Wigner-Eckart projection accepts them) and the
    polyacetylene example of the reference (examples/polyacetylene.jl:29-31, BASELINE config C4) builds."""
    from hubbardtn_b200 import device as dev
    t = np.array([[0.3, 0.1, 1.0, 0.5], [0.1, -0.2, 0.25, 0.8]])
    u = np.array([[3.0, 0.7, 0.25, 0.1], [0.5, 2.0, 0.0, 0.4]])
    J = np.array([[0.0, 0.37, 0.21, 0.05], [0.11, 0.0, 0.0, 0.3]])
    sim = hf.MB_Sim(t, u, J, P=1, Q=1, kwargs={"spin": spin})
    Ws, levels = hf.fsm_mpo_dense(sim.sym, sim.Q, *hf.mb_terms(sim))
    N = 5
    Hm, Hd = chain_from_mpo(Ws, N), direct_hamiltonian(sim, N)
    assert np.abs(Hd - Hd.T).max() < 1e-13
    assert np.abs(Hm - Hd).max() < 1e-12
    # every site tensor is an invariant tensor of the symmetry group: the projection onto reduced entries succeeds
    P = dev.Legs(None, sim.sym, PS.physical_space(sim.sym, 1, 1))
    Mleg = dev.Legs(None, sim.sym, levels)
    for W in Ws:
        assert dev.Mpo.from_dense(None, Mleg, P, Mleg, W).nnz > 0
    # the exchange part alone is -2J S.S - J/2 nn + J (pair hopping): two sites, one pair
    c = jw_ops(sim.sym, 2)
    Sv = [[0.5 * (c[p][0].T @ c[p][1] + c[p][1].T @ c[p][0]), None, 0.5 * (c[p][0].T @ c[p][0] - c[p][1].T @ c[p][1])] for p in (0, 1)]
    SS = Sv[0][2] @ Sv[1][2] + 0.5 * (c[0][0].T @ c[0][1] @ c[1][1].T @ c[1][0] + c[0][1].T @ c[0][0] @ c[1][0].T @ c[1][1])
    n = [c[p][0].T @ c[p][0] + c[p][1].T @ c[p][1] for p in (0, 1)]
    X = sum(c[0][s1].T @ c[1][s2].T @ c[0][s2] @ c[1][s1] for s1 in (0, 1) for s2 in (0, 1))
    assert np.abs(X + 2.0 * SS + 0.5 * n[0] @ n[1]).max() < 1e-13


def test_polyacetylene_model_of_the_reference_builds():
    """examples/polyacetylene.jl:29-31 (BASELINE config C4): t, U, J of the two-band model; the MPO has 1 + 2x3 hopping
    doublets... levels and every site tensor projects onto reduced entries; the one-band exchange of HF:445-450 goes the
    same way."""
    from hubbardtn_b200 import device as dev
    t = np.array([[0.000, 3.803, -0.548, 0.000], [3.803, 0.000, 2.977, -0.501]])
    U = np.array([[10.317, 6.264, 0.000, 0.000], [6.264, 10.317, 6.162, 0.000]])
    J = np.array([[0.000, 0.123, 0.000, 0.000], [0.123, 0.000, 0.113, 0.000]])
    sim = hf.MB_Sim(t, U, J, P=1, Q=1, svalue=2.5, bond_dim=20)
    Ws, levels = hf.fsm_mpo_dense(sim.sym, sim.Q, *hf.mb_terms(sim))
    assert len(Ws) == 4 and levels[0] == levels[-1] == (0, 0, 0)
    assert (0, 2, 0) in levels and (0, 0, 2) in levels and (0, 0, -2) in levels
    N = 4
    assert np.abs(chain_from_mpo(Ws, N) - direct_hamiltonian(sim, N)).max() < 1e-11
    H = hf.hamiltonian(sim, ctx=None)
    assert len(H) == 4 and H.chi == len(levels)
    ob = hf.OB_Sim([1.0], [4.0], 0.0, [0.4, 0.1], 1, 1, 2.0)
    Wo, lev = hf.fsm_mpo_dense(ob.sym, ob.Q, *hf.ob_extended_terms(ob))
    c = jw_ops(ob.sym, 4)
    Hd = np.zeros((4 ** 4, 4 ** 4))
    for p in range(4):
        nu, nd = c[p][0].T @ c[p][0], c[p][1].T @ c[p][1]
        Hd += 4.0 * nu @ nd
        if p + 1 < 4:
            for s in (0, 1):
                Hd += -1.0 * (c[p][s].T @ c[p + 1][s] + c[p + 1][s].T @ c[p][s])
        for r, jr in ((1, 0.4), (2, 0.1)):
            if p + r < 4:
                q = p + r
                for s1 in (0, 1):
                    for s2 in (0, 1):
                        Hd += jr * (c[p][s1].T @ c[q][s2].T @ c[p][s2] @ c[q][s1])
                ph = c[p][0].T @ c[p][1].T @ c[q][1] @ c[q][0]
                Hd += jr * (ph + ph.T)
    assert np.abs(chain_from_mpo(Wo, 4) - Hd).max() < 1e-12
    assert len(hf.hamiltonian(ob, ctx=None)) == 2


def _ob_direct(sim, N, hop_dists, field=None):
    c = jw_ops(sim.sym, N)
    H = np.zeros((4 ** N, 4 ** N))
    for p in range(N):
        nu, nd = c[p][0].T @ c[p][0], c[p][1].T @ c[p][1]
        H += sim.u[0] * nu @ nd - sim.mu * (nu + nd)
        if field is not None:
            H += field * (-1.0) ** ((p % sim.unit_cell) + 1) * 0.5 * (nu - nd)
        for d in hop_dists:
            if p + d < N:
                for s in (0, 1):
                    H += -sim.t[0] * (c[p][s].T @ c[p + d][s] + c[p + d][s].T @ c[p][s])
    return H


def test_helix_and_staggered_field_variants():
    """HF:463-465 (helix of circumference `period`: hops at distance 1 and period) and HF:458-462 (staggered field
    J_inter Ms (-1)^i S^z_i, spin-resolved symmetry only) against the directly written Hamiltonians."""
    N = 5
    helix = hf.OB_Sim([1.3], [4.0], 0.2, [0.0], 1, 1, 2.0, 50, 3)
    Ws, _ = hf.fsm_mpo_dense(helix.sym, helix.Q, *hf.ob_extended_terms(helix))
    assert np.abs(chain_from_mpo(Ws, N) - _ob_direct(helix, N, (1, 3))).max() < 1e-12
    with pytest.raises(ValueError):
        hf.ob_extended_terms(hf.OB_Sim([1.0, 0.5], [4.0], 0.0, [0.0], 1, 1, 2.0, 50, 3))     # HF:467
    stag = hf.OB_Sim([1.0], [6.0], 0.0, [0.0], 1, 1, 2.0, 50, 0, kwargs={"spin": True, "JMs": (0.7, 0.4)})
    Ws, _ = hf.fsm_mpo_dense(stag.sym, stag.Q, *hf.ob_extended_terms(stag))
    assert len(Ws) == 2 and np.abs(Ws[0] - Ws[1]).max() > 0.1                                 # site-dependent
    assert np.abs(chain_from_mpo(Ws, N) - _ob_direct(stag, N, (1,), field=0.7 * 0.4)).max() < 1e-12
    # without spin resolution the field is ignored, as in the reference (HF:458 `&& spin`)
    plain = hf.OB_Sim([1.0], [6.0], 0.0, [0.0], 1, 1, 2.0, 50, 0, kwargs={"JMs": (0.7, 0.4)})
    W1, _ = hf.hamiltonian_dense(plain)
    W0, _ = hf.hamiltonian_dense(hf.OB_Sim([1.0], [6.0], 0.0, [0.0], 1, 1, 2.0))
    assert np.array_equal(W1, W0)
