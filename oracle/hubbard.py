"""One-band Hubbard MPO in reduced (symmetric) form (oracle; test infrastructure).

Restates `hamiltonian(::OB_Sim)` of the reference (src/HubbardFunctions.jl:386-472):

    H = sum_i [ u[1] n_up n_dn - mu n ]_i  -  sum_r t[r] sum_i sum_s (c+_{i,s} c_{i+r,s} + h.c.)
        + sum_{r>=2} u[r] sum_i n_i n_{i+r-1}

with the physical space, U(1) charge convention (charge = Q*occ - P) and symmetry sectors of
`SymSpace` (HF:245-255) and the local operator entries of `Hopping`, `OSInteraction`,
`Number` (HF:257-327).  The reference builds fermionic (graded) TensorMaps and lets
MPSKitModels' `@mpoham` assemble the MPO (not vendored).  Here fermionic signs are made
explicit with Jordan-Wigner parity strings, the MPO is written as a finite-state machine in
the plain 4-dimensional occupation basis, checked against exact diagonalisation, and then
projected onto its reduced entries (Wigner-Eckart; `MPOTensor.from_dense` verifies the
symmetry).  The MPO level layout is therefore defined by this repo;
energies and observables are convention independent.

Exchange (J), U13, staggered-field and helix terms (HF:445-469) are not restated yet
(SURVEY.md 8(f) "next" #1).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import sectors as S
from .spaces import physical_space
from .tensors import Legs, MPOTensor


@dataclass
class OB_Sim:
    """Mirror of `OB_Sim(t, u, mu, J, P, Q, svalue, bond_dim, period; kwargs...)` (HF:76-93)."""
    t: list
    u: list
    mu: float = 0.0
    P: int = 1
    Q: int = 1
    svalue: float = 2.0
    bond_dim: int = 50
    period: int = 0
    kwargs: dict = field(default_factory=dict)

    @property
    def spin(self) -> bool:
        return bool(self.kwargs.get("spin", False))

    @property
    def unit_cell(self) -> int:
        return self.Q if self.P % 2 == 0 else 2 * self.Q      # HF:408-412


# ---- local operators in the occupation basis |0>, |up>, |dn>, |up dn> = c+_up c+_dn |0> --------
def _local_ops():
    cu = np.zeros((4, 4))
    cd = np.zeros((4, 4))
    cu[0, 1] = 1.0      # c_up |up> = |0>
    cu[2, 3] = 1.0      # c_up |up dn> = |dn>
    cd[0, 2] = 1.0      # c_dn |dn> = |0>
    cd[1, 3] = -1.0     # c_dn c+_up c+_dn |0> = -|up>
    par = np.diag([1.0, -1.0, -1.0, 1.0])
    nup, ndn = cu.T @ cu, cd.T @ cd
    return cu, cd, par, nup + ndn, nup @ ndn


def _basis_perm(kind: int):
    """occupation-basis index of every state of the multiplet basis used by `physical_space`:
    SU2U1: empty, double, single(m=-1/2 -> dn, m=+1/2 -> up); U1U1: empty, double, up, dn."""
    return [0, 3, 2, 1] if kind == S.SU2U1 else [0, 3, 1, 2]


def local_operators(kind: int):
    """(c_up, c_dn, parity, n, double occupancy) as 4x4 matrices in the multiplet basis."""
    perm = _basis_perm(kind)
    return tuple(op[np.ix_(perm, perm)] for op in _local_ops())


def dense_mpo(kind: int, t, u, mu: float, Q: int = 1):
    """Finite-state-machine MPO W[a, s', s, b] (dense) and the sector label of every level.

    levels: 0 = start, then for k = 1..R:  A_k (c+ ... c chain, one level per spin projection),
    B_k (c ... c+ chain), then N_k (n ... n chain, k = 1..len(u)-1), last = end."""
    cu, cd, par, num, dbl = local_operators(kind)
    R, Rn = len(t), max(len(u) - 1, 0)
    labels = [(0, 0, 0)]
    idx = {}
    # spin projection order inside a doublet: m = -1/2 first (oracle CG convention)
    for k in range(1, R + 1):
        if kind == S.SU2U1:
            labels.append((1, 1, Q)); idx[("A", k)] = len(labels) - 1       # doublet (two dense levels)
        else:
            labels.append((1, -1, Q)); idx[("A", k, -1)] = len(labels) - 1
            labels.append((1, 1, Q)); idx[("A", k, 1)] = len(labels) - 1
    for k in range(1, R + 1):
        if kind == S.SU2U1:
            labels.append((1, 1, -Q)); idx[("B", k)] = len(labels) - 1
        else:
            labels.append((1, -1, -Q)); idx[("B", k, -1)] = len(labels) - 1
            labels.append((1, 1, -Q)); idx[("B", k, 1)] = len(labels) - 1
    for k in range(1, Rn + 1):
        labels.append((0, 0, 0)); idx[("N", k)] = len(labels) - 1
    labels.append((0, 0, 0))
    M = Legs(kind, labels)
    end = len(labels) - 1
    W = np.zeros((M.full_dim, 4, 4, M.full_dim))

    def dense_level(key, m):
        """dense index of chain level `key` with spin projection 2m in {-1,+1}."""
        if kind == S.SU2U1:
            return M.full_offset[idx[key]] + (0 if m < 0 else 1)
        return M.full_offset[idx[key + (m,)]]

    o0, oe = M.full_offset[0], M.full_offset[end]
    W[o0, :, :, o0] = np.eye(4)
    W[oe, :, :, oe] = np.eye(4)
    W[o0, :, :, oe] = (u[0] if len(u) else 0.0) * dbl - mu * num
    cdag = {1: cu.T, -1: cd.T}
    cann = {1: cu, -1: cd}
    for k in range(1, R + 1):
        for m in (-1, 1):
            # A chain: c+_{i,s} c_{j,s} = (c+_s P)_i (x) P ... (x) (c_s)_j ; level spin = spin added = s
            a = dense_level(("A", k), m)
            if k == 1:
                W[o0, :, :, a] = cdag[m] @ par
            else:
                W[dense_level(("A", k - 1), m), :, :, a] = par
            W[a, :, :, oe] = -t[k - 1] * cann[m]
            # B chain: h.c. = (P c_s)_i (x) P ... (x) (c+_s)_j ; level spin added = -s.  (c_dn, -c_up)
            # is the standard spinor conjugate to (c+_up, c+_dn): level m carries sign(m) * c_{-m}
            b = dense_level(("B", k), m)
            sgn = 1.0 if m > 0 else -1.0
            if k == 1:
                W[o0, :, :, b] = sgn * (par @ cann[-m])
            else:
                W[dense_level(("B", k - 1), m), :, :, b] = par
            W[b, :, :, oe] = -t[k - 1] * sgn * cdag[-m]
    for k in range(1, Rn + 1):
        n_ = M.full_offset[idx[("N", k)]]
        if k == 1:
            W[o0, :, :, n_] = num
        else:
            W[M.full_offset[idx[("N", k - 1)]], :, :, n_] = np.eye(4)
        W[n_, :, :, oe] = u[k] * num
    return W, M


def mpo(sim: OB_Sim):
    """Per-site reduced MPO tensors of the unit cell: ([MPOTensor]*L, physical Legs, level Legs)."""
    kind = S.U1U1 if sim.spin else S.SU2U1
    if sim.period != 0 or any(k in sim.kwargs for k in ("U13", "JMs")):
        raise NotImplementedError("helix / U13 / staggered-field terms are not restated yet (HF:452-469)")
    Wd, M = dense_mpo(kind, list(sim.t), list(sim.u), sim.mu, sim.Q)
    P = physical_space(kind, sim.P, sim.Q)
    W = MPOTensor.from_dense(Wd, M, P, M)
    return [W] * sim.unit_cell, P, M


# ---- exact diagonalisation reference (open chain), for the MPO check -----------------------------
def ed_hamiltonian(kind: int, N: int, t, u, mu: float):
    """Second-quantised H on an open N-site chain, built with explicit Jordan-Wigner strings,
    in the product of the multiplet bases (same local basis as `dense_mpo`)."""
    cu, cd, par, num, dbl = local_operators(kind)
    I4 = np.eye(4)

    def site_op(ops):
        out = np.ones((1, 1))
        for i in range(N):
            out = np.kron(out, ops.get(i, I4))
        return out

    def c(i, s):   # annihilator of spin s on site i with the string on sites < i
        ops = {k: par for k in range(i)}
        ops[i] = cu if s > 0 else cd
        return site_op(ops)

    H = np.zeros((4 ** N, 4 ** N))
    for i in range(N):
        H += site_op({i: (u[0] if len(u) else 0.0) * dbl - mu * num})
        for r, tr in enumerate(t, start=1):
            if i + r < N:
                for s in (1, -1):
                    hop = c(i, s).T @ c(i + r, s)
                    H += -tr * (hop + hop.T)
        for r in range(1, len(u)):
            if i + r < N:
                H += u[r] * site_op({i: num, i + r: num})
    return H


def mpo_to_hamiltonian(Wd: np.ndarray, M: Legs, N: int):
    """Contract N copies of the dense MPO between the start and end level (open chain)."""
    o0, oe = M.full_offset[0], M.full_offset[len(M) - 1]
    left = np.zeros((M.full_dim, 1, 1))
    left[o0, 0, 0] = 1.0
    for _ in range(N):
        d = left.shape[1]
        left = np.einsum("apq,asrb->bpsqr", left, Wd).reshape(M.full_dim, d * 4, d * 4)
    return left[oe]
