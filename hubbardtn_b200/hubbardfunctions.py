"""Host-side mirror of HubbardTN's public interface for the ground-state path, running on libhtn.

Names, arguments and return conventions follow `src/HubbardFunctions.jl` of the reference so that the
parity tests read like the reference's own tests (test/OB.jl, test/Spin.jl):

    OB_Sim(t, u, mu, J, P, Q, svalue, bond_dim, period; kwargs...)      HF:76-93
    MB_Sim(t, u, J, U13, P, Q, svalue, bond_dim; kwargs...)             HF:117-134
    hamiltonian(simul)                                                  HF:386-472 (t, u, mu terms), HF:811-910 (hopping,
                                                                        band energies, on-band U, direct interactions)
    initialize_mps(H, P, max_dimension, spin)                           HF:917-959
    compute_groundstate(simul; tol, verbosity, maxiter)                 HF:993-1030
    produce_groundstate(simul; force)                                   HF:1145-1166 (in-memory cache only)
    TruncState(simul, trunc_dim; trunc_scheme=0)                        HF:1351-1387 (VUMPSSvdCut / SvdCut)
    dim_state(psi)                                                      HF:1399-1405
    density_state(psi)                                                  HF:1475-1542

The reference delegates the numerics to MPSKit; here `compute_groundstate` drives the same schedule
(IDMRG2 with truncbelow(10^-svalue), then VUMPS) through the C ABI on the GPU.  This module contains
host bookkeeping only (spaces, the MPO as a finite-state machine, random initial blocks); it never
touches `oracle/`.  The helix (`period`) and staggered-field (`JMs`) variants of the one-band model (HF:458-465)
go through the same site-dependent finite-state-machine builder as `MB_Sim`.  Not mirrored yet: exchange / U13 /
U_ijkk / U_ijkl terms (HF:445-457, 563-643, 662-809) and one-site unit cells (the VUMPS + SvdCut bond-growing
loop, HF:1011-1022).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import device as dev
from . import sectors as S


@dataclass
class OB_Sim:
    """One-band Hubbard simulation parameters (HF:76-93)."""
    t: list
    u: list
    mu: float = 0.0
    J: list = field(default_factory=lambda: [0.0])
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
    def sym(self) -> int:
        return S.U1U1 if self.spin else S.SU2U1

    @property
    def unit_cell(self) -> int:
        return self.Q if self.P % 2 == 0 else 2 * self.Q          # HF:408-412


# ---------------------------------------------------------------------------------------------------
# Hamiltonian as a finite-state-machine MPO (dense, multiplet basis) -> reduced form inside libhtn
# ---------------------------------------------------------------------------------------------------
def _fermion_ops(sym: int):
    """(c_up, c_dn, parity, n, n_up n_dn) as 4x4 matrices in the basis of `S.physical_space`:
    SU2U1: empty, double, single(m=-1/2 = dn, m=+1/2 = up);  U1U1: empty, double, up, dn.
    |double> = c+_up c+_dn |0>."""
    cu = np.zeros((4, 4))
    cd = np.zeros((4, 4))
    cu[0, 1] = 1.0
    cu[2, 3] = 1.0
    cd[0, 2] = 1.0
    cd[1, 3] = -1.0
    par = np.diag([1.0, -1.0, -1.0, 1.0])
    perm = [0, 3, 2, 1] if sym == S.SU2U1 else [0, 3, 1, 2]
    ix = np.ix_(perm, perm)
    nup, ndn = cu.T @ cu, cd.T @ cd
    return cu[ix], cd[ix], par[ix], (nup + ndn)[ix], (nup @ ndn)[ix]


def hamiltonian_dense(simul: OB_Sim):
    """H = sum_i [u1 n_up n_dn - mu n]_i - sum_r t_r sum_i,s (c+_{i,s} c_{i+r,s} + h.c.) + sum_{r>=2} u_r n_i n_{i+r-1}
    (HF:424-444) as an upper-triangular MPO W[a, s', s, b] with explicit Jordan-Wigner strings.
    Returns (W dense, level sectors)."""
    if simul.period != 0 or "U13" in simul.kwargs or any(abs(j) > 0 for j in simul.J):
        raise NotImplementedError("U13 terms (HF:452-457) are not mirrored yet; helix, staggered field and exchange go "
                                  "through ob_extended_terms")
    sym, Q = simul.sym, simul.Q
    cu, cd, par, num, dbl = _fermion_ops(sym)
    t, u = list(simul.t), list(simul.u)
    R, Rn = len(t), max(len(u) - 1, 0)
    levels = [(0, 0, 0)]
    first = {}

    def add(name, sectors):
        first[name] = len(levels)
        levels.extend(sectors)

    for k in range(1, R + 1):
        add(("A", k), [(1, 1, Q)] if sym == S.SU2U1 else [(1, -1, Q), (1, 1, Q)])
    for k in range(1, R + 1):
        add(("B", k), [(1, 1, -Q)] if sym == S.SU2U1 else [(1, -1, -Q), (1, 1, -Q)])
    for k in range(1, Rn + 1):
        add(("N", k), [(0, 0, 0)])
    levels.append((0, 0, 0))
    off, acc = [], 0
    for s in levels:
        off.append(acc)
        acc += S.dim(sym, s)
    Dm, end = acc, off[-1]
    W = np.zeros((Dm, 4, 4, Dm))

    def lvl(name, m):       # dense index of the doublet member with 2 m = -1 / +1
        return off[first[name]] + (0 if m < 0 else 1)

    W[0, :, :, 0] = np.eye(4)
    W[end, :, :, end] = np.eye(4)
    W[0, :, :, end] = (u[0] if u else 0.0) * dbl - simul.mu * num
    cdag, cann = {1: cu.T, -1: cd.T}, {1: cu, -1: cd}
    for k in range(1, R + 1):
        for m in (-1, 1):
            a = lvl(("A", k), m)                    # c+_{i,s} P ... P c_{j,s}
            if k == 1:
                W[0, :, :, a] = cdag[m] @ par
            else:
                W[lvl(("A", k - 1), m), :, :, a] = par
            W[a, :, :, end] = -t[k - 1] * cann[m]
            b = lvl(("B", k), m)                    # h.c.: the conjugate spinor (c_dn, -c_up)
            sgn = 1.0 if m > 0 else -1.0
            if k == 1:
                W[0, :, :, b] = sgn * (par @ cann[-m])
            else:
                W[lvl(("B", k - 1), m), :, :, b] = par
            W[b, :, :, end] = -t[k - 1] * sgn * cdag[-m]
    for k in range(1, Rn + 1):
        n_ = off[first[("N", k)]]
        if k == 1:
            W[0, :, :, n_] = num
        else:
            W[off[first[("N", k - 1)]], :, :, n_] = np.eye(4)
        W[n_, :, :, end] = u[k] * num
    return W, levels


@dataclass
class MB_Sim:
    """Multi-band Hubbard simulation parameters (HF:117-134).  `t`, `u`, `J` are B x (B + B*range) matrices:
    the leading B x B block is the on-site part (band energies / on-band U on its diagonal), block k >= 1
    couples band i of a cell to band j of the k-th next cell."""
    t: np.ndarray
    u: np.ndarray
    J: np.ndarray = None
    U13: np.ndarray = None
    P: int = 1
    Q: int = 1
    svalue: float = 2.0
    bond_dim: int = 50
    kwargs: dict = field(default_factory=dict)

    def __post_init__(self):
        self.t = np.atleast_2d(np.asarray(self.t, float))
        self.u = np.atleast_2d(np.asarray(self.u, float))
        B = self.t.shape[0]
        self.J = np.zeros((B, B)) if self.J is None else np.atleast_2d(np.asarray(self.J, float))
        self.U13 = np.zeros((B, B)) if self.U13 is None else np.atleast_2d(np.asarray(self.U13, float))
        if not (self.u.shape[0] == self.J.shape[0] == self.U13.shape[0] == B):
            raise ValueError("Number of bands is incosistent.")                      # HF:826
        for m in (self.t, self.u, self.J):
            if m.shape[1] % B or m.shape[1] < B:
                raise ValueError("parameter matrices must be B x (B + B*range)")

    @property
    def bands(self) -> int:
        return self.t.shape[0]

    @property
    def spin(self) -> bool:
        return bool(self.kwargs.get("spin", False))

    @property
    def sym(self) -> int:
        return S.U1U1 if self.spin else S.SU2U1

    @property
    def unit_cell(self) -> int:
        return (self.Q if self.P % 2 == 0 else 2 * self.Q) * self.bands          # HF:829-834, InfiniteStrip(B, T*B)


def _spin_ops(sym: int):
    """(S^+, S^-, S^z, Delta^+) as 4x4 matrices in the local basis of `_fermion_ops`; Delta^+ = c+_up c+_dn creates the
    doubly occupied state (|double> = c+_up c+_dn |0>)."""
    cu, cd, par, num, dbl = _fermion_ops(sym)
    sp = cu.T @ cd                      # S^+ = c+_up c_dn
    sz = 0.5 * (cu.T @ cu - cd.T @ cd)
    dp = cu.T @ cd.T                    # Delta^+
    return sp, sp.T, sz, dp


def fsm_mpo_dense(sym: int, Q: int, onsite: list, hops: dict, dens: dict, spins: dict = None, pairs: dict = None):
    """Finite-state-machine MPO of  sum_p onsite[p] + sum c (c+_{p,s} c_{p+d,s} + h.c.) + sum v n_p n_{p+d}
    + sum g S_p . S_{p+d} + sum h (Delta+_p Delta_{p+d} + h.c.)
    on a chain with an L-site unit cell (L = len(onsite)).  hops / dens / spins / pairs: {(p mod L, d >= 1): coefficient}.
    The spin-spin term travels on a triplet level (0, j=1, 0) whose three dense members carry the spherical components
    S^(q), q = -1, 0, +1, on the opening site and their adjoints on the closing site (U(1)xU(1): three abelian levels
    S^z, S^+, S^-); the pair term on the two scalar levels of charge +-2Q.  Both are bosonic: identity fillers.
    A level (kind, d) means "d more sites until the term closes"; the coefficient sits on the opening
    site, so every site shares the level list and only the dense entries differ.  Fermionic pairs carry
    an explicit Jordan-Wigner parity string.  Returns ([W_p dense (chi,4,4,chi)], level sectors)."""
    L = len(onsite)
    cu, cd, par, num, dbl = _fermion_ops(sym)
    spins, pairs = spins or {}, pairs or {}
    dh = max([d for (_, d) in hops] + [0])
    dn = max([d for (_, d) in dens] + [0])
    ds = max([d for (_, d) in spins] + [0])
    dp_ = max([d for (_, d) in pairs] + [0])
    levels = [(0, 0, 0)]
    first = {}

    def add(name, sectors):
        first[name] = len(levels)
        levels.extend(sectors)

    for d in range(1, dh + 1):
        add(("A", d), [(1, 1, Q)] if sym == S.SU2U1 else [(1, -1, Q), (1, 1, Q)])
    for d in range(1, dh + 1):
        add(("B", d), [(1, 1, -Q)] if sym == S.SU2U1 else [(1, -1, -Q), (1, 1, -Q)])
    for d in range(1, dn + 1):
        add(("N", d), [(0, 0, 0)])
    for d in range(1, ds + 1):
        add(("S", d), [(0, 2, 0)] if sym == S.SU2U1 else [(0, -2, 0), (0, 0, 0), (0, 2, 0)])
    for d in range(1, dp_ + 1):
        add(("Dp", d), [(0, 0, 2 * Q)])
        add(("Dm", d), [(0, 0, -2 * Q)])
    levels.append((0, 0, 0))
    off, acc = [], 0
    for s in levels:
        off.append(acc)
        acc += S.dim(sym, s)
    Dm, end = acc, off[-1]

    def lvl(name, m):
        return off[first[name]] + (0 if m < 0 else 1)

    cdag, cann = {1: cu.T, -1: cd.T}, {1: cu, -1: cd}
    Ws = []
    for p in range(L):
        W = np.zeros((Dm, 4, 4, Dm))
        W[0, :, :, 0] = np.eye(4)
        W[end, :, :, end] = np.eye(4)
        W[0, :, :, end] = onsite[p]
        for d in range(1, dh + 1):
            for m in (-1, 1):
                sgn = 1.0 if m > 0 else -1.0
                a, b = lvl(("A", d), m), lvl(("B", d), m)
                c = hops.get((p, d), 0.0)
                if c != 0.0:
                    W[0, :, :, a] = c * (cdag[m] @ par)                # c+_{p,s} P ... P c_{p+d,s}
                    W[0, :, :, b] = c * sgn * (par @ cann[-m])         # h.c. through the conjugate spinor (c_dn, -c_up)
                if d > 1:
                    W[a, :, :, lvl(("A", d - 1), m)] = par
                    W[b, :, :, lvl(("B", d - 1), m)] = par
                else:
                    W[a, :, :, end] = cann[m]
                    W[b, :, :, end] = sgn * cdag[-m]
        for d in range(1, dn + 1):
            n_ = off[first[("N", d)]]
            v = dens.get((p, d), 0.0)
            if v != 0.0:
                W[0, :, :, n_] = v * num
            if d > 1:
                W[n_, :, :, off[first[("N", d - 1)]]] = np.eye(4)
            else:
                W[n_, :, :, end] = num
        if ds:
            sp, sm, sz, _ = _spin_ops(sym)
            # spherical components q = -1, 0, +1 (dense members of the triplet in this order) and their adjoints:
            # sum_q S^(q)_i (S^(q))^+_j = S_i . S_j
            sph = [sm / np.sqrt(2.0), sz, -sp / np.sqrt(2.0)]
            for d in range(1, ds + 1):
                s0 = off[first[("S", d)]]
                g = spins.get((p, d), 0.0)
                for q in range(3):
                    if g != 0.0:
                        W[0, :, :, s0 + q] = g * sph[q]
                    if d > 1:
                        W[s0 + q, :, :, off[first[("S", d - 1)]] + q] = np.eye(4)
                    else:
                        W[s0 + q, :, :, end] = sph[q].T
        if dp_:
            _, _, _, dpl = _spin_ops(sym)
            for d in range(1, dp_ + 1):
                a, b = off[first[("Dp", d)]], off[first[("Dm", d)]]
                h = pairs.get((p, d), 0.0)
                if h != 0.0:
                    W[0, :, :, a] = h * dpl          # Delta+_p ... Delta_{p+d}
                    W[0, :, :, b] = h * dpl.T        # h.c.
                if d > 1:
                    W[a, :, :, off[first[("Dp", d - 1)]]] = np.eye(4)
                    W[b, :, :, off[first[("Dm", d - 1)]]] = np.eye(4)
                else:
                    W[a, :, :, end] = dpl.T
                    W[b, :, :, end] = dpl
        Ws.append(W)
    return Ws, levels


def mb_terms(simul: MB_Sim):
    """Term lists of the multi-band Hamiltonian (HF:811-910) on the chain p = band + cell * B:
    on-band U and band energies (HF:531-551, 852-870), on-site and inter-site hopping (HF:477-519), direct
    on-site and inter-site interactions (HF:542-561, 645-659), on-site and inter-site exchange (HF:563-616, 668-700).
    U_ijjj and the three-/four-band dictionaries (HF:617-643, 702-809) are not mirrored.

    Exchange: the reference contracts two hopping tensors into `C4_1 = sum_ss' c+_is c+_js' c_is' c_js` (HF:580, 675) and
    `C4_2 = sum_ss' c+_is c+_is' c_js' c_js` (HF:604, 690).  In operators C4_1 = -(2 S_i.S_j + n_i n_j / 2) and
    C4_2{i,j} + C4_2{j,i} = 2 (Delta+_i Delta_j + h.c.), so a pair with exchange integral J carries
    -2J S_i.S_j - (J/2) n_i n_j + J (Delta+_i Delta_j + h.c.)  (Hund's coupling and pair hopping of the Kanamori form).
    The signs are those of the second-quantised operators the terms are named after; TensorKit's fermionic braiding signs
    of the `@tensor` lines cannot be checked without Julia (no reference golden has J != 0).
    Returns (onsite, hops, dens, spins, pairs)."""
    B, L = simul.bands, simul.unit_cell
    if np.any(simul.U13 != 0) or any(k in simul.kwargs for k in ("U112", "U1111", "U13_IS")):
        raise NotImplementedError("U_ijjj / U_ijkk / U_ijkl terms (HF:617-643, 702-809) are not mirrored yet")
    if np.any(np.diag(simul.J[:, :B]) != 0):
        raise ValueError("On-band interaction is not taken into account in Exchange_OS.")    # HF:575 (a warning there)
    _, _, _, num, dbl = _fermion_ops(simul.sym)
    t, u = simul.t, simul.u
    onsite = [u[p % B, p % B] * dbl - t[p % B, p % B] * num for p in range(L)]      # OB_interaction + Chem_pot
    hops, dens, spins, pairs = {}, {}, {}, {}
    J = simul.J

    def acc(dct, p, d, c):
        if c != 0.0:
            dct[(p % L, d)] = dct.get((p % L, d), 0.0) + c

    def exchange(p, d, j):
        acc(spins, p, d, -2.0 * j)
        acc(dens, p, d, -0.5 * j)
        acc(pairs, p, d, j)

    for cell in range(L // B):
        for bi in range(B):
            for bf in range(bi + 1, B):
                p = cell * B + bi
                acc(hops, p, bf - bi, -0.5 * (t[bi, bf] + t[bf, bi]))                  # OS_Hopping (both orders of (bi,bf))
                acc(dens, p, bf - bi, 0.5 * (u[bi, bf] + u[bf, bi]))                   # Direct_OS: U_av on the lower triangle
                exchange(p, bf - bi, 0.5 * (J[bi, bf] + J[bf, bi]))                    # Exchange_OS: 0.5 J over both orders
        for k in range(1, t.shape[1] // B):
            for bi in range(B):
                for bf in range(B):
                    acc(hops, cell * B + bi, k * B + bf - bi, -t[bi, k * B + bf])      # IS_Hopping: twosite = cdc + cdc'
        for k in range(1, u.shape[1] // B):
            for bi in range(B):
                for bf in range(B):
                    acc(dens, cell * B + bi, k * B + bf - bi, u[bi, k * B + bf])       # Direct_IS
        for k in range(1, J.shape[1] // B):
            for bi in range(B):
                for bf in range(B):
                    exchange(cell * B + bi, k * B + bf - bi, J[bi, k * B + bf])        # Exchange_IS
    return onsite, hops, dens, spins, pairs


def ob_extended_terms(simul: OB_Sim):
    """Term lists of the one-band variants that go beyond `hamiltonian_dense`: the helix of circumference `period`
    (HF:463-465: -t (c+_i c_{i+1} + c+_i c_{i+period} + h.c.), on-site U only) and the staggered field
    J_inter Ms (-1)^i S^z_i of `JMs` (HF:458-462, U(1)xU(1) only; i = 1..T as `enumerate` counts)."""
    L = simul.unit_cell
    cu, cd, par, num, dbl = _fermion_ops(simul.sym)
    t, u = list(simul.t), list(simul.u)
    JMs = simul.kwargs.get("JMs", (0.0, 0.0))
    if "U13" in simul.kwargs:
        raise NotImplementedError("U13 terms (HF:452-457) are not mirrored yet")
    onsite = [(u[0] if u else 0.0) * dbl - simul.mu * num for _ in range(L)]
    hops, dens, spins, pairs = {}, {}, {}, {}
    if simul.period != 0:
        if len(t) != 1 or len(u) != 1:
            raise ValueError("Extended models in 2D not implemented.")            # HF:467
        for p in range(L):
            for d in (1, simul.period):
                hops[(p, d)] = hops.get((p, d), 0.0) - t[0]
    else:
        for p in range(L):
            for r, tr in enumerate(t, start=1):
                if tr != 0.0:
                    hops[(p, r)] = -tr
            for r, ur in enumerate(u[1:], start=1):
                if ur != 0.0:
                    dens[(p, r)] = ur
    if simul.period == 0:
        for p in range(L):
            for r, jr in enumerate(simul.J, start=1):                  # HF:445-450: J1 + J2 at range r (see mb_terms)
                if jr != 0.0:
                    spins[(p, r)] = spins.get((p, r), 0.0) - 2.0 * jr
                    dens[(p, r)] = dens.get((p, r), 0.0) - 0.5 * jr
                    pairs[(p, r)] = pairs.get((p, r), 0.0) + jr
    if JMs[1] != 0.0 and simul.spin:
        sz = 0.5 * (cu.T @ cu - cd.T @ cd)
        for p in range(L):
            onsite[p] = onsite[p] + JMs[0] * JMs[1] * (-1.0) ** (p + 1) * sz
    return onsite, hops, dens, spins, pairs


class Hamiltonian:
    """`InfiniteMPOHamiltonian` stand-in: per-site reduced MPO tensors held by libhtn."""

    def __init__(self, ctx, simul):
        self.simul, self.sym = simul, simul.sym
        if isinstance(simul, MB_Sim):
            Wd, self.levels = fsm_mpo_dense(self.sym, simul.Q, *mb_terms(simul))
        elif (simul.period != 0 or (simul.kwargs.get("JMs", (0.0, 0.0))[1] != 0.0 and simul.spin)
              or any(abs(j) > 0 for j in simul.J)):
            Wd, self.levels = fsm_mpo_dense(self.sym, simul.Q, *ob_extended_terms(simul))
        else:
            W1, self.levels = hamiltonian_dense(simul)
            Wd = [W1]
        self.phys = S.physical_space(self.sym, simul.P, simul.Q)
        self.P = dev.Legs(ctx, self.sym, self.phys)
        self.M = dev.Legs(ctx, self.sym, self.levels)
        mpos = [dev.Mpo.from_dense(ctx, self.M, self.P, self.M, w) for w in Wd]
        self.W = [mpos[i % len(mpos)] for i in range(simul.unit_cell)]
        self.chi = len(self.levels)

    def __len__(self):
        return len(self.W)


def hamiltonian(simul, ctx=None) -> Hamiltonian:
    return Hamiltonian(ctx, simul)


# ---------------------------------------------------------------------------------------------------
# initial state (HF:917-959)
# ---------------------------------------------------------------------------------------------------
def _fuse_spaces(sym, a: dict, b: dict) -> dict:
    out = {}
    for sa, na in a.items():
        for sb, nb in b.items():
            for c in S.fuse(sym, sa, sb):
                out[c] = out.get(c, 0) + na * nb
    return out


def _dual(sym, s):
    return (s[0], s[1], -s[2]) if sym == S.SU2U1 else (s[0], -s[1], -s[2])


def initial_spaces(sym: int, phys: list, L: int, P: int, max_dimension: int) -> list:
    """V[i] = right bond of site i: infimum of the spaces fused from the left and (dual) from the right,
    capped per sector at `max_dimension` inside the window of HF:931-947."""
    pd = {}
    for s in phys:
        pd[s] = pd.get(s, 0) + 1
    v_right, acc = [], None
    for _ in range(L):
        acc = dict(pd) if acc is None else _fuse_spaces(sym, acc, pd)
        v_right.append(acc)
    v_l, acc = [], {(0, 0, 0): 1}
    dual_pd = {_dual(sym, s): n for s, n in pd.items()}
    for _ in range(L):
        acc = _fuse_spaces(sym, acc, dual_pd)
        v_l.append(acc)
    v_left = list(reversed(v_l))
    v_left = v_left[1:] + v_left[:1]
    out = []
    for i in range(L):
        v = {s: min(n, v_right[i][s]) for s, n in v_left[i].items() if s in v_right[i]}
        capped = {}
        for s, n in v.items():
            inside = (s[1] <= 6 if sym == S.SU2U1 else abs(s[1]) <= L) and abs(s[2]) <= L * P
            if inside or s == (0, 0, 0):
                # HF:935/944: Vmax holds (0,0,0)=>1 next to (0,0,0)=>max_dimension
                capped[s] = min(n, max_dimension + (1 if s == (0, 0, 0) else 0)) if inside else 1
        out.append({s: n for s, n in capped.items() if n > 0})
    return _full_rank(sym, out, phys)


def _full_rank(sym, spaces, phys):
    """Shrink multiplicities until every site tensor can be left- and right-isometric."""
    L = len(spaces)
    mult = [dict(m) for m in spaces]
    changed = True
    while changed:
        changed = False
        for i in range(L):
            vl, vr = mult[i - 1], mult[i]
            cap_r = {r: 0 for r in vr}
            cap_l = {l: 0 for l in vl}
            for l, nl in vl.items():
                for s in phys:
                    for r in S.fuse(sym, l, s):
                        if r in vr:
                            cap_r[r] += nl
                            cap_l[l] += vr[r]
            for r in list(vr):
                if vr[r] > cap_r[r]:
                    vr[r], changed = cap_r[r], True
                if vr[r] == 0:
                    del vr[r]
                    changed = True
            for l in list(vl):
                if vl[l] > cap_l[l]:
                    vl[l], changed = cap_l[l], True
                if vl[l] == 0:
                    del vl[l]
                    changed = True
    return mult


class InfiniteMPS:
    """Uniform MPS in mixed gauge, resident on the device: AL, AR, AC, C (C[i] right of site i)."""

    def __init__(self, ctx, sym, AL, AR, C, AC, phys=None):
        self.ctx, self.sym, self.AL, self.AR, self.C, self.AC = ctx, sym, AL, AR, C, AC
        self.phys = list(phys) if phys is not None else None      # physical multiplets (sector labels), needed by save_state

    def __len__(self):
        return len(self.AL)

    def bond_space(self, i):
        sp = self.C[i].space(0, self.sym)
        return dict(zip(sp.sectors, sp.mult))


def initialize_mps(H: Hamiltonian, P: int, max_dimension: int, spin: bool = False, ctx=None, seed: int = 20261018):
    """Random uniform MPS on the spaces of HF:917-959, brought to mixed gauge on the device."""
    sym, L = H.sym, len(H)
    spaces = initial_spaces(sym, H.phys, L, P, max_dimension)
    V = [dev.Space(ctx, sym, m) for m in spaces]
    rng = np.random.Generator(np.random.Philox(key=seed))
    AL, AR, AC, C = [], [], [], []
    for i in range(L):
        a = dev.Tensor.mps(ctx, V[i - 1], H.P, V[i])
        a.upload(rng.standard_normal(a.nelem))
        AL.append(a)
        AR.append(a.like())
        AC.append(a.like())
        C.append(dev.Tensor.bond(ctx, V[i]))
    guess = dev.Tensor.bond(ctx, V[L - 1])
    g = np.zeros(guess.nelem)
    for lab, blk in guess.block_views(g).items():
        blk[...] = np.eye(blk.shape[0]) + 0.1 * rng.standard_normal(blk.shape)
    guess.upload(g)
    dev.mixed_gauge(ctx, AL, guess, AR, C, AC, tol=1e-12)
    return InfiniteMPS(ctx, sym, AL, AR, C, AC, H.phys)


# ---------------------------------------------------------------------------------------------------
# ground state (HF:993-1030)
# ---------------------------------------------------------------------------------------------------
def _make_envs(ctx, psi: InfiniteMPS, H: Hamiltonian):
    L = len(psi)
    V = [psi.C[i].space(0, psi.sym) for i in range(L)]
    GL = [dev.Tensor.env(ctx, 0, V[i - 1], H.M, identity_level=0) for i in range(L)]
    GR = [dev.Tensor.env(ctx, 1, V[i], H.M, identity_level=H.chi - 1) for i in range(L)]
    return GL, GR


def compute_groundstate(simul: OB_Sim, ctx=None, tol: float = 1e-6, verbosity: int = 0, maxiter: int = 1000,
                        init_state=None, seed: int = 20261018):
    """HF:993-1030: IDMRG2(trscheme=truncbelow(10^-svalue), tol) followed by VUMPS(tol, maxiter).
    Returns the dictionary of HF:1029 with an extra "energy" (per site) and iteration logs."""
    own_ctx = ctx is None
    if own_ctx:
        ctx = dev.Context(0)
    H = hamiltonian(simul, ctx)
    if len(H) < 2:
        raise NotImplementedError("one-site unit cells (VUMPS + SvdCut bond growing, HF:1011-1022) are not mirrored yet")
    psi = init_state if init_state is not None else initialize_mps(H, simul.P, simul.bond_dim, simul.spin, ctx, seed)
    schmidtcut = 10.0 ** (-simul.svalue)                                   # HF:1007
    # htn_idmrg2 replaces the tensors it is given: work on copies so that the caller's init_state survives (HF:997-1005)
    start = [[t.like_copy() for t in lst] for lst in (psi.AL, psi.AR, psi.C, psi.AC)]
    AL, AR, C, AC, info1 = dev.idmrg2(ctx, start[0], start[1], start[2], start[3], H.W, cut=schmidtcut, tol=tol,
                                      maxiter=min(maxiter, 200))           # HF:1010 (MPSKit default maxiter 200)
    AL, AR, C, AC = dev.uniform_from_right(ctx, AR, C[-1], simul.sym)      # MPSKit: InfiniteMPS(psi.AR) at the end of IDMRG2
    psi = InfiniteMPS(ctx, simul.sym, AL, AR, C, AC, H.phys)
    GL, GR = _make_envs(ctx, psi, H)
    info2 = dev.vumps(ctx, psi.AL, psi.AR, psi.C, psi.AC, H.W, GL, GR, tol=tol, maxiter=min(maxiter, 1000))  # HF:1025-1027
    # ... & GradientGrassmann(; maxiter, tol): MPSKit runs both stages; the second returns at once when tol is already met
    info3 = None
    if True:
        info3 = dev.gradient_grassmann(ctx, psi.AL, psi.AR, psi.C, psi.AC, H.W, GL, GR, tol=tol, maxiter=min(maxiter, 1000))
        if info3["iterations"] > 0 or not info2["converged"]:
            info2 = dict(info2, delta=info3["delta"], energy_per_site=info3["energy_per_site"], converged=info3["converged"])
    if verbosity > 0:
        print("IDMRG2: %d iterations, delta %.3e; VUMPS: %d iterations, galerkin %.3e, E/site %.10f"
              % (info1["iterations"], info1["delta"], info2["iterations"], info2["delta"], info2["energy_per_site"]))
    return {"groundstate": psi, "environments": (GL, GR), "ham": H, "delta": info2["delta"], "config": simul,
            "energy": info2["energy_per_site"], "idmrg2": info1, "vumps": info2, "gradient_grassmann": info3, "ctx": ctx}


_CACHE = {}


def produce_groundstate(simul: OB_Sim, force: bool = False, **kw):
    """HF:1145-1166 without the DrWatson/JLD2 disk cache (out of scope): memoised per parameter set."""
    if isinstance(simul, MB_Sim):
        key = ("MB", simul.t.tobytes(), simul.u.tobytes(), simul.t.shape, simul.u.shape, simul.P, simul.Q, simul.svalue,
               simul.bond_dim, simul.spin)
    else:
        key = (tuple(simul.t), tuple(simul.u), simul.mu, simul.P, simul.Q, simul.svalue, simul.bond_dim, simul.spin,
               simul.period, tuple(simul.kwargs.get("JMs", (0.0, 0.0))))
    key = key + (kw.get("tol", 1e-6), kw.get("maxiter", 1000), kw.get("seed", 20261018), id(kw.get("ctx")))
    if kw.get("init_state") is not None:                 # a caller-supplied start is never served from the cache
        return compute_groundstate(simul, **kw)
    if force or key not in _CACHE:
        _CACHE[key] = compute_groundstate(simul, **kw)
    return _CACHE[key]


def TruncState(simul, trunc_dim: int, trunc_scheme: int = 0, **kw):
    """HF:1351-1366: ground state truncated to (full) bond dimension `trunc_dim` with
    `changebonds(psi, H, VUMPSSvdCut(trscheme = truncdim(trunc_dim)))` (trunc_scheme 0, the reference's default) or
    `changebonds(psi, SvdCut(trscheme = truncdim(trunc_dim)))` (trunc_scheme 1).  The cap is applied to the number of
    kept multiplets such that the full dimension sum_c dim(c) n_c stays <= trunc_dim (TensorKit's truncdim counts the
    full dimension)."""
    if trunc_dim <= 0:
        raise ValueError("trunc_dim should be a positive integer.")           # HF:1353
    if trunc_scheme not in (0, 1):
        raise ValueError("trunc_scheme should be either 0 (VUMPSSvdCut) or 1 (SvdCut).")   # HF:1356
    d = produce_groundstate(simul, **kw)
    psi, H, ctx = d["groundstate"], d["ham"], d["ctx"]

    # htn_changebonds with maxdim < 0 = truncdim(trunc_dim) on the FULL dimension: the kept set of every bond is decided
    # in one pass from its singular values (TensorKit's truncdim), no search over multiplet caps
    AL, AR, C, AC = dev.changebonds(ctx, 1 if trunc_scheme == 0 else 0, psi.AL, psi.AR, psi.C, psi.AC, H.W, maxdim=-int(trunc_dim))
    return InfiniteMPS(ctx, psi.sym, AL, AR, C, AC, psi.phys)


def produce_TruncState(simul, trunc_dim: int, trunc_scheme: int = 0, force: bool = False, **kw):
    """HF:1378-1387 without the disk cache: {"ψ_trunc": state, "envs_trunc": (GL, GR)} as `TruncState` returns in
    the reference (HF:1366); the environments are those of the truncated state."""
    if force:
        produce_groundstate(simul, force=True, **kw)
    psi = TruncState(simul, trunc_dim, trunc_scheme, **kw)
    d = produce_groundstate(simul, **kw)
    GL, GR = _make_envs(d["ctx"], psi, d["ham"])
    dev.environments(d["ctx"], psi.AL, psi.AR, psi.C, d["ham"].W, GL, GR, tol=1e-10)
    return {"ψ_trunc": psi, "psi_trunc": psi, "envs_trunc": (GL, GR)}


# ---------------------------------------------------------------------------------------------------
# state I/O (HF:1669-1691)
# ---------------------------------------------------------------------------------------------------
STATE_FORMAT = "hubbardtn-b200/state-v1"


def save_state(psi: InfiniteMPS, path: str, name: str):
    """HF:1669-1677: one file per site under `path/name/` holding the left isometry AL[i] as a plain dictionary --
    the reference stores `convert(Dict, psi.AL[i])` in `state$i.jld2`; here `state{i}.npz` (1-based like the
    reference) with the symmetry, the three graded spaces (sector labels (p, q, n) and multiplicities), the block
    table (labels, rows, cols, offsets of the packed row-major blocks in canonical order) and the data.  JLD2/HDF5
    cannot be written from this image; INTEGRATION.md shows the NPZ.jl reader that turns a file back into a
    TensorMap."""
    import os
    d = os.path.join(path, name)
    os.mkdir(d)                                                   # like the reference: fails if the state exists
    if psi.phys is None:
        raise ValueError("save_state: the state does not carry its physical space (InfiniteMPS(..., phys=...))")
    phys = psi.phys
    for i in range(len(psi)):
        A = psi.AL[i]
        vl, vr = A.space(0, psi.sym), A.space(1, psi.sym)
        np.savez(os.path.join(d, "state%d.npz" % (i + 1)), format=STATE_FORMAT, sym=psi.sym,
                 left_sectors=np.array(vl.sectors, dtype=np.int32).reshape(-1, 3), left_mult=np.array(vl.mult, dtype=np.int32),
                 right_sectors=np.array(vr.sectors, dtype=np.int32).reshape(-1, 3), right_mult=np.array(vr.mult, dtype=np.int32),
                 phys_sectors=np.array(phys, dtype=np.int32).reshape(-1, 3),
                 labels=A.labels, rows=A.rows, cols=A.cols, offsets=A.offsets, data=A.download())
        print("State %d saved." % (i + 1))


def load_state(path: str, ctx=None, tol: float = 1e-12):
    """HF:1679-1691: reads `state1 .. stateN` and rebuilds the uniform MPS from the left isometries
    (`InfiniteMPS(PeriodicArray(A))`: gauge fixing on the device)."""
    import os
    files = sorted((f for f in os.listdir(path) if os.path.isfile(os.path.join(path, f)) and f.startswith("state")),
                   key=lambda f: int("".join(ch for ch in f if ch.isdigit())))
    if not files:
        raise FileNotFoundError("no state files under %s" % path)
    if ctx is None:
        ctx = dev.Context(0)
    AL, sym = [], None
    for f in files:
        z = np.load(os.path.join(path, f))
        if str(z["format"]) != STATE_FORMAT:
            raise ValueError("%s: unknown state format %r" % (f, str(z["format"])))
        sym = int(z["sym"])
        Vl = dev.Space(ctx, sym, {tuple(int(v) for v in s): int(m) for s, m in zip(z["left_sectors"], z["left_mult"])})
        Vr = dev.Space(ctx, sym, {tuple(int(v) for v in s): int(m) for s, m in zip(z["right_sectors"], z["right_mult"])})
        P = dev.Legs(ctx, sym, [tuple(int(v) for v in s) for s in z["phys_sectors"]])
        A = dev.Tensor.mps(ctx, Vl, P, Vr)
        if not (np.array_equal(A.labels, z["labels"]) and np.array_equal(A.rows, z["rows"]) and np.array_equal(A.cols, z["cols"])
                and np.array_equal(A.offsets, z["offsets"])):
            raise ValueError("%s: block table differs from the canonical one of its spaces" % f)
        A.upload(np.ascontiguousarray(z["data"], dtype=np.float64))
        AL.append(A)
    n = len(AL)
    AR = [a.like() for a in AL]
    AC = [a.like() for a in AL]
    Cs = [dev.Tensor.bond(ctx, AL[i].space(1, sym)) for i in range(n)]
    dev.mixed_gauge(ctx, AL, dev._identity_bond(ctx, AL[n - 1].space(1, sym)), AR, Cs, AC, tol=tol)
    return InfiniteMPS(ctx, sym, AL, AR, Cs, AC, [tuple(int(v) for v in s) for s in z["phys_sectors"]])


def dim_state(psi: InfiniteMPS):
    """Full (quantum-dimension weighted) bond dimension of every site's left bond (HF:1399-1405)."""
    out = []
    for i in range(len(psi)):
        sp = psi.AL[i].space(0, psi.sym)
        out.append(sum(S.dim(psi.sym, s) * n for s, n in zip(sp.sectors, sp.mult)))
    return out


def entanglement_spectrum(psi: InfiniteMPS, site: int = 0):
    """MPSKit `entanglement_spectrum(psi, site)`: {sector label: Schmidt values} of the bond right of `site`."""
    sp = psi.C[site].space(0, psi.sym)
    spec = dev.entanglement_spectrum(psi.C[site])
    return {sp.sectors[c]: v for c, v in spec.items()}


def density_state(psi: InfiniteMPS):
    """<n_i> per site (HF:1495-1542; `expectation_value(psi, i => n)` HF:1507)."""
    vals = [0.0, 2.0, 1.0] if psi.sym == S.SU2U1 else [0.0, 2.0, 1.0, 1.0]
    return [dev.expval_diag(psi.AC[i], vals) for i in range(len(psi))]


def density_spin(psi: InfiniteMPS):
    """(<n_up>_i, <n_down>_i) per site for a U(1)xU(1) state (HF:1413-1454): n_up is 1 on the doubly occupied and
    the up multiplet, n_down on the doubly occupied and the down multiplet."""
    if psi.sym != S.U1U1:
        raise ValueError("This system is spin independent.")                  # HF:1424
    up = [dev.expval_diag(psi.AC[i], [0.0, 1.0, 1.0, 0.0]) for i in range(len(psi))]
    down = [dev.expval_diag(psi.AC[i], [0.0, 1.0, 0.0, 1.0]) for i in range(len(psi))]
    return up, down


def calc_ms(psi: InfiniteMPS):
    """Staggered magnetisation |<n_up> - <n_down>| of the first site (HF:1461-1468)."""
    up, down = density_spin(psi)
    mag = [u - d for u, d in zip(up, down)]
    return abs(mag[0])
