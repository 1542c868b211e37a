"""Uniform MPS algebra in reduced form: gauge fixing, environments, VUMPS (oracle).

TEST INFRASTRUCTURE.  Restates, from their published form (Zauner-Stauber et al., PRB 97,
045145 (2018)), the MPSKit 0.13.1 routines HubbardTN reaches through
`find_groundstate(psi, H, VUMPS(...))` (src/HubbardFunctions.jl:1012,1017,1025-1027) and
`InfiniteMPS(...)` (HF:958,990); MPSKit is not vendored (Manifest.toml:722):

  left_orth / right_orth        <- TensorKit `leftorth!(.., QRpos())` / `rightorth!(.., LQpos())`
  uniform_rightorth             <- MPSKit `uniform_rightorth!` (iterated LQ through the unit cell)
  regauge                       <- MPSKit `regauge!`: AL = Q(AC) Q(C)^T
  environments                  <- MPSKit `environments(psi, H)` for Jordan-form MPOs: level by
                                   level transfer, identity-diagonal level by GMRES
  vumps                         <- MPSKit `find_groundstate(.., VUMPS)`: eigsolve H_AC, H_C per
                                   site, regauge, gauge fix, environments, galerkin error

Conventions.  Site tensors A[l,s,r] (n_l x n_r blocks), bond matrices C[c]; C[i] sits on the
bond to the RIGHT of site i; GL[i] on the bond to the LEFT of site i, GR[i] on the bond to its
RIGHT.  With the isometric Clebsch-Gordan normalisation of oracle/tensors.py:
  left-orthonormal :  sum_{l,s} A[lsr]^T A[lsr]                 = 1   for every r
  right-orthonormal:  sum_{s,r} (d_r/d_l) A[lsr] A[lsr]^T       = 1   for every l
MPO levels: level 0 and level chi-1 carry the identity on the diagonal, all other diagonal
entries vanish (the Jordan form oracle/hubbard.py produces).
"""
from __future__ import annotations

import numpy as np

from . import sectors as S
from .heff import HeffACPlan, heff_c_apply, network
from .krylov import gmres, lanczos_lowest, vaxpy, vcopy, vdot, vnorm, vscale
from .tensors import BondTensor, EnvTensor, Legs, MPOTensor, MPSTensor, Space


# ----------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------
def _env_weight(self, key):
    """Inner-product weight of an environment block = dim of the coupled (bra) sector."""
    return self.V.dims[key[1]] if self.side == "L" else self.V.dims[key[2]]


def _env_copy(self):
    return EnvTensor(self.side, self.V, self.M, {k: v.copy() for k, v in self.blocks.items()},
                     identity_levels=self.identity_levels)


EnvTensor.weight = _env_weight
EnvTensor.copy = _env_copy


def mul_right(A: MPSTensor, C: BondTensor) -> MPSTensor:
    """A . C  (C on the right bond)."""
    return MPSTensor(A.Vl, A.P, A.Vr, {k: v @ C.blocks[k[2]] for k, v in A.blocks.items()})


def mul_left(C: BondTensor, A: MPSTensor) -> MPSTensor:
    """C . A  (C on the left bond)."""
    return MPSTensor(A.Vl, A.P, A.Vr, {k: C.blocks[k[0]] @ v for k, v in A.blocks.items()})


def bond_norm(C: BondTensor) -> float:
    return float(np.sqrt(sum(C.V.dims[c] * np.vdot(b, b) for c, b in C.blocks.items())))


def trim_spaces(kind, spaces, phys):
    """Shrink multiplicities until every site tensor can be both left- and right-isometric:
    n_r <= sum_{(l,s)->r} n_l and n_l <= sum_{(s,r)<-l} n_r (what TensorKit's infimum/fuse
    logic guarantees for the spaces of HF:917-959).  spaces[i] = right bond of site i."""
    L = len(spaces)
    mult = [dict(sp.as_dict()) for sp in spaces]
    changed = True
    while changed:
        changed = False
        for i in range(L):
            vl, vr, P = mult[i - 1], mult[i], phys[i]
            cap_r = {r: 0 for r in vr}
            cap_l = {l: 0 for l in vl}
            for l, nl in vl.items():
                for s in P.sectors:
                    for r in S.fuse(kind, l, s):
                        if r in vr:
                            cap_r[r] += nl
                            cap_l[l] += vr[r]
            for r in list(vr):
                if vr[r] > cap_r[r]:
                    vr[r] = cap_r[r]
                    changed = True
                if vr[r] == 0:
                    del vr[r]
                    changed = True
            for l in list(vl):
                if vl[l] > cap_l[l]:
                    vl[l] = cap_l[l]
                    changed = True
                if vl[l] == 0:
                    del vl[l]
                    changed = True
    return [Space(kind, m) for m in mult]


# ----------------------------------------------------------------------------------------
# positive QR / LQ per coupled sector
# ----------------------------------------------------------------------------------------
def _qrpos(M):
    q, r = np.linalg.qr(M)
    sgn = np.sign(np.diag(r))
    sgn[sgn == 0] = 1.0
    return q * sgn[None, :], r * sgn[:, None]


def left_orth(A: MPSTensor):
    """A = Q . R with Q left-orthonormal and diag(R) > 0, per right sector r (rows of the
    coupled block run over the fusion trees (l,s) -> r in canonical key order)."""
    by_r = {}
    for key in A.keys:
        by_r.setdefault(key[2], []).append(key)
    Q = A.zeros_like()
    R = BondTensor(A.Vr)
    for r, keys in by_r.items():
        M = np.vstack([A.blocks[k] for k in keys])
        assert M.shape[0] >= M.shape[1], "left_orth: block of sector %d is wide (%s)" % (r, M.shape)
        q, rr = _qrpos(M)
        o = 0
        for k in keys:
            n = A.blocks[k].shape[0]
            Q.blocks[k] = q[o:o + n]
            o += n
        R.blocks[r] = rr
    return Q, R


def right_orth(A: MPSTensor):
    """A = L . Q with Q right-orthonormal and diag(L) > 0, per left sector l (columns run over
    (s,r); blocks enter with weight sqrt(d_r/d_l))."""
    by_l = {}
    for key in A.keys:
        by_l.setdefault(key[0], []).append(key)
    Q = A.zeros_like()
    Lm = BondTensor(A.Vl)
    for l, keys in by_l.items():
        wts = [np.sqrt(A.Vr.dims[k[2]] / A.Vl.dims[l]) for k in keys]
        M = np.hstack([w * A.blocks[k] for w, k in zip(wts, keys)])
        assert M.shape[1] >= M.shape[0], "right_orth: block of sector %d is tall (%s)" % (l, M.shape)
        q, rr = _qrpos(M.T)
        Lm.blocks[l] = rr.T
        qt = q.T
        o = 0
        for w, k in zip(wts, keys):
            n = A.blocks[k].shape[1]
            Q.blocks[k] = qt[:, o:o + n] / w
            o += n
    return Lm, Q


def regauge(AC: MPSTensor, C: BondTensor) -> MPSTensor:
    """AL = Q(AC) . Q(C)^T   (MPSKit `regauge!`, QRpos on both)."""
    Qac, _ = left_orth(AC)
    out = AC.zeros_like()
    qc = {c: _qrpos(b)[0] for c, b in C.blocks.items()}
    for k, v in Qac.blocks.items():
        out.blocks[k] = v @ qc[k[2]].T
    return out


EIG_MINITER = 10     # plain QR/LQ sweeps before the Arnoldi-accelerated ones from a cold start (MPSKit); VUMPS passes 2


def arnoldi_dominant(apply, x0, tol=1e-12, krylovdim=30, maxiter=10):
    """Dominant eigenvector of a (non-symmetric) linear map by Arnoldi with CGS2 and explicit restarts
    (KrylovKit `eigsolve(f, x0, 1, :LM, Arnoldi(...))` as used by MPSKit's uniform_left/rightorth!)."""
    x = vscale(vcopy(x0), 1.0 / vnorm(x0))
    res = np.inf
    for _ in range(maxiter):
        V = [x]
        H = np.zeros((krylovdim + 1, krylovdim))
        m = 0
        for j in range(krylovdim):
            w = apply(V[j])
            for _pass in range(2):
                for i, q in enumerate(V):
                    c = vdot(q, w)
                    H[i, j] += c
                    vaxpy(-c, q, w)
            hn = vnorm(w)
            H[j + 1, j] = hn
            m = j + 1
            ev, evec = np.linalg.eig(H[:m, :m])
            k = int(np.argmax(np.abs(ev)))
            y = np.real(evec[:, k])
            y /= np.linalg.norm(y)
            res = abs(hn * y[-1])
            if res < tol or hn < 1e-14 or j == krylovdim - 1:
                break
            V.append(vscale(w, 1.0 / hn))
        xn = vscale(vcopy(V[0]), y[0])
        for q, c in zip(V[1:m], y[1:m]):
            vaxpy(c, q, xn)
        x = vscale(xn, 1.0 / vnorm(xn))
        if res < tol:
            break
    return x, res


def _bond_env(C: BondTensor, side: str, triv) -> "EnvTensor":
    return EnvTensor(side, C.V, triv, {(0, c, c): b.copy() for c, b in C.blocks.items()})


def _env_bond(E: "EnvTensor") -> BondTensor:
    return BondTensor(E.V, {k[1]: b.copy() for k, b in E.blocks.items()})


def _tri_factor(C: BondTensor, lower: bool) -> BondTensor:
    out = BondTensor(C.V)
    for c, b in C.blocks.items():
        out.blocks[c] = _qrpos(b.T)[1].T if lower else _qrpos(b)[1]
    nrm = bond_norm(out)
    for b in out.blocks.values():
        b /= nrm
    return out


def uniform_rightorth(AL, C_last: BondTensor, tol=1e-13, maxiter=10000, eig_miniter=EIG_MINITER):
    """From left-orthonormal AL[0..L-1] and a guess for C[L-1]: AR[i], C[i] with
    AL[i] C[i] = C[i-1] AR[i]  (iterated LQ through the unit cell until C[L-1] is stationary)."""
    L = len(AL)
    C = [None] * L
    C[L - 1] = C_last.copy()
    nrm = bond_norm(C[L - 1])
    for b in C[L - 1].blocks.values():
        b /= nrm
    AR = [None] * L
    delta = np.inf
    its = 0
    idR = None
    for its in range(1, maxiter + 1):
        if its > eig_miniter:
            if idR is None:
                triv = Legs(AL[0].kind, [S.trivial(AL[0].kind)])
                idR = [TransferPlan("R", identity_mpo(AL[i].P), AL[i].Vl, AL[i].P, AL[i].Vr) for i in range(L)]

            def op(X):
                T = _bond_env(X, "R", triv)
                for k in range(L - 1, -1, -1):
                    T = idR[k].apply(T, AL[k], AR[k])
                return _env_bond(T)

            x, _ = arnoldi_dominant(op, C[L - 1], tol=min(1e-3, max(delta * delta, 1e-13)))
            C[L - 1] = _tri_factor(x, lower=True)
        Cold = C[L - 1]
        for i in range(L - 1, -1, -1):
            Lm, Q = right_orth(mul_right(AL[i], C[i]))
            nrm = bond_norm(Lm)
            for b in Lm.blocks.values():
                b /= nrm
            C[i - 1 if i > 0 else L - 1] = Lm
            AR[i] = Q
        delta = np.sqrt(sum(C[L - 1].V.dims[c] * np.sum((C[L - 1].blocks[c] - Cold.blocks[c]) ** 2)
                            for c in Cold.blocks))
        if delta < tol:
            break
    return AR, C, dict(iterations=its, delta=float(delta))


def uniform_leftorth(AR, C_last: BondTensor, tol=1e-13, maxiter=10000, eig_miniter=EIG_MINITER):
    """Mirror of `uniform_rightorth`: from (approximately) right-orthonormal AR[0..L-1] and a guess for
    C[L-1]: AL[i], C[i] with AL[i] C[i] = C[i-1] AR[i]  (iterated positive QR, MPSKit `uniform_leftorth!`)."""
    L = len(AR)
    C = [None] * L
    C[L - 1] = C_last.copy()
    nrm = bond_norm(C[L - 1])
    for b in C[L - 1].blocks.values():
        b /= nrm
    AL = [None] * L
    delta, its = np.inf, 0
    idL = None
    for its in range(1, maxiter + 1):
        if its > eig_miniter:
            if idL is None:
                triv = Legs(AR[0].kind, [S.trivial(AR[0].kind)])
                idL = [TransferPlan("L", identity_mpo(AR[i].P), AR[i].Vl, AR[i].P, AR[i].Vr) for i in range(L)]

            def op(X):
                T = _bond_env(X, "L", triv)
                for k in range(L):
                    T = idL[k].apply(T, AR[k], AL[k])
                return _env_bond(T)

            x, _ = arnoldi_dominant(op, C[L - 1], tol=min(1e-3, max(delta * delta, 1e-13)))
            C[L - 1] = _tri_factor(x, lower=False)
        Cold = C[L - 1]
        for i in range(L):
            Q, R = left_orth(mul_left(C[i - 1 if i > 0 else L - 1], AR[i]))
            nrm = bond_norm(R)
            for b in R.blocks.values():
                b /= nrm
            C[i] = R
            AL[i] = Q
        delta = np.sqrt(sum(C[L - 1].V.dims[c] * np.sum((C[L - 1].blocks[c] - Cold.blocks[c]) ** 2)
                            for c in Cold.blocks))
        if delta < tol:
            break
    return AL, C, dict(iterations=its, delta=float(delta))


def random_state(kind, spaces, phys, rng):
    """Random uniform MPS in mixed gauge on the given bond spaces (right bond of site i)."""
    L = len(spaces)
    AL = []
    for i in range(L):
        A = MPSTensor(spaces[i - 1], phys[i], spaces[i]).randomize(rng)
        AL.append(left_orth(A)[0])
    C0 = BondTensor(spaces[L - 1])
    for c in C0.blocks:
        n = C0.blocks[c].shape[0]
        C0.blocks[c] = np.eye(n) + 0.1 * rng.standard_normal((n, n))
    # iterate the left gauge once more so AL is a consistent uniform state (any left-orthonormal
    # AL is a valid state; only the right gauge needs the fixed point)
    AR, C, _ = uniform_rightorth(AL, C0, tol=1e-12)
    return dict(AL=AL, AR=AR, C=C, AC=[mul_right(AL[i], C[i]) for i in range(L)])


# ----------------------------------------------------------------------------------------
# transfers with cached term lists
# ----------------------------------------------------------------------------------------
class TransferPlan:
    """Term list of the MPO transfer through one site for fixed spaces (side 'L': GL on the left
    bond -> GL on the right bond; side 'R': GR on the right bond -> GR on the left bond)."""

    def __init__(self, side, W: MPOTensor, Vl: Space, P: Legs, Vr: Space):
        self.side, self.W, self.Vl, self.P, self.Vr = side, W, Vl, P, Vr
        k = Vl.kind
        akeys = MPSTensor(Vl, P, Vr).keys
        by_s = {}
        for (l, s, r) in akeys:
            by_s.setdefault(s, []).append((l, r))
        aset = set(akeys)
        acc = {}
        for (a, sp, s, b, c), w in W.entries.items():
            ca, cb = W.Ml.sectors[a], W.Mr.sectors[b]
            for (l, r) in by_s.get(s, ()):
                for clp in S.fuse(k, ca, Vl.sectors[l]):
                    lp = Vl.index.get(clp)
                    if lp is None:
                        continue
                    for crp in S.fuse(k, cb, Vr.sectors[r]):
                        rp = Vr.index.get(crp)
                        if rp is None or (lp, sp, rp) not in aset:
                            continue
                        n = network(k, clp, P.sectors[sp], crp, Vl.sectors[l], P.sectors[s],
                                    Vr.sectors[r], ca, cb, c)
                        if n == 0.0:
                            continue
                        if side == "L":
                            key = ((b, rp, r), (a, lp, l), (l, s, r), (lp, sp, rp))
                            cf = w * n / Vr.dims[rp]
                        else:
                            key = ((a, l, lp), (b, r, rp), (l, s, r), (lp, sp, rp))
                            cf = w * n / Vl.dims[lp]
                        acc[key] = acc.get(key, 0.0) + cf
        self.terms = [(k0, k1, k2, k3, cf) for (k0, k1, k2, k3), cf in acc.items() if cf != 0.0]

    def apply(self, G: EnvTensor, A: MPSTensor, Abar: MPSTensor = None) -> EnvTensor:
        Abar = A if Abar is None else Abar
        if self.side == "L":
            out = EnvTensor("L", self.Vr, self.W.Mr)
            for (ko, kg, ka, kb, cf) in self.terms:
                g = G.blocks.get(kg)
                if g is None:
                    continue
                out.blocks[ko] += cf * (Abar.blocks[kb].T @ g @ A.blocks[ka])
        else:
            out = EnvTensor("R", self.Vl, self.W.Ml)
            for (ko, kg, ka, kb, cf) in self.terms:
                g = G.blocks.get(kg)
                if g is None:
                    continue
                out.blocks[ko] += cf * (A.blocks[ka] @ g @ Abar.blocks[kb].T)
        return out


def identity_mpo(P: Legs) -> MPOTensor:
    M = Legs(P.kind, [S.trivial(P.kind)])
    return MPOTensor(M, P, M, {(0, s, s, 0, cs): 1.0 for s, cs in enumerate(P.sectors)})


def mpo_depth(W_list) -> int:
    """Longest path (in sites) through the strictly upper-triangular part of the level graph."""
    chi = len(W_list[0].Ml)
    depth = [0] * chi
    for _ in range(chi):
        changed = False
        for W in W_list:
            for (a, sp, s, b, c) in W.entries:
                if a != b and depth[b] < depth[a] + 1:
                    depth[b] = depth[a] + 1
                    changed = True
        if not changed:
            break
    return max(depth)


# ----------------------------------------------------------------------------------------
# environments of a Jordan-form MPO Hamiltonian
# ----------------------------------------------------------------------------------------
class Environments:
    """GL[i], GR[i] for every site and the energy per unit cell.

    Left side: GL[i][0] = 1.  Levels 1..chi-2 follow from repeated site transfers (the level
    graph is strictly upper triangular there).  Last level: X = GL[0][chi-1] solves
        X - T_cell(X) + (X|rho) 1 = Y - (Y|rho) 1 ,   rho = C C^T on the bond left of site 0,
    where Y is what one trip round the unit cell deposits on the last level and
    (Y|rho) = energy per unit cell.  The right side is the mirror image (first level)."""

    def __init__(self, state, W_list, tol=1e-12, krylovdim=30, maxiter=200):
        self.W = W_list
        L = self.L = len(W_list)
        AL, AR, C = state["AL"], state["AR"], state["C"]
        chi = len(W_list[0].Ml)
        self.chi = chi
        depth = mpo_depth(W_list)
        kind = AL[0].kind
        self.planL = [TransferPlan("L", W_list[i], AL[i].Vl, AL[i].P, AL[i].Vr) for i in range(L)]
        self.planR = [TransferPlan("R", W_list[i], AR[i].Vl, AR[i].P, AR[i].Vr) for i in range(L)]
        idL = [TransferPlan("L", identity_mpo(AL[i].P), AL[i].Vl, AL[i].P, AL[i].Vr) for i in range(L)]
        idR = [TransferPlan("R", identity_mpo(AR[i].P), AR[i].Vl, AR[i].P, AR[i].Vr) for i in range(L)]
        M = W_list[0].Ml
        triv = Legs(kind, [S.trivial(kind)])

        def unit_env(side, V, M, level):
            e = EnvTensor(side, V, M, identity_levels=[level])
            e.fix_identity_levels()
            return e

        def set_level(e, level, src, src_level=0):
            for key in e.keys:
                if key[0] == level:
                    sk = (src_level, key[1], key[2])
                    e.blocks[key] = src.blocks[sk].copy() if src is not None else np.zeros_like(e.blocks[key])

        def bond_dot(X, rho, side):
            # (X|rho): X an env restricted to one trivial level, rho a BondTensor
            return float(sum(X.V.dims[k[1]] * np.vdot(X.blocks[k], rho.blocks[k[1]]) for k in X.blocks))

        # ---------------- left ----------------
        # GL[i] lives on bond i-1 (left of site i); bond index -1 == L-1
        GL = [unit_env("L", AL[i].Vl, W_list[i].Ml, 0) for i in range(L)]
        steps = depth + L - 1
        i = 0
        for _ in range(max(steps, L)):
            nxt = (i + 1) % L
            out = self.planL[i].apply(GL[i], AL[i])
            set_level(out, chi - 1, None)
            for key in out.keys:                       # level 0 stays the unit tensor
                if key[0] == 0:
                    out.blocks[key] = np.eye(out.blocks[key].shape[0])
            out.identity_levels = {0}
            GL[nxt] = out
            i = nxt
        # make sure the sweep ended so that every site has been refreshed with complete inputs:
        # run one more full ring starting at site 0 (levels < chi-1 are now exact everywhere)
        for i in range(L):
            nxt = (i + 1) % L
            out = self.planL[i].apply(GL[i], AL[i])
            if nxt != 0:
                GL[nxt] = out
                GL[nxt].identity_levels = {0}
            else:
                Ylast = out
        rhoL = BondTensor(C[L - 1].V, {c: b @ b.T for c, b in C[L - 1].blocks.items()})
        Y = EnvTensor("L", AL[0].Vl, triv)
        set_level(Y, 0, Ylast, chi - 1)
        eL = bond_dot(Y, rhoL, "L")
        one = unit_env("L", AL[0].Vl, triv, 0)

        def opL(X):
            T = X
            for i in range(L):
                T = idL[i].apply(T, AL[i])
            out = vcopy(X)
            vaxpy(-1.0, T, out)
            vaxpy(bond_dot(X, rhoL, "L"), one, out)
            return out

        rhs = vcopy(Y)
        vaxpy(-eL, one, rhs)
        X, infoL = gmres(opL, rhs, tol=tol, krylovdim=krylovdim, maxiter=maxiter)
        set_level(GL[0], chi - 1, X, 0)
        for i in range(L - 1):
            GL[i + 1] = self.planL[i].apply(GL[i], AL[i])
            GL[i + 1].identity_levels = {0}
        # ---------------- right ----------------
        GR = [unit_env("R", AR[i].Vr, W_list[i].Mr, chi - 1) for i in range(L)]
        i = L - 1
        for _ in range(max(steps, L)):
            prv = (i - 1) % L
            out = self.planR[i].apply(GR[i], AR[i])
            set_level(out, 0, None)
            for key in out.keys:
                if key[0] == chi - 1:
                    out.blocks[key] = np.eye(out.blocks[key].shape[0])
            out.identity_levels = {chi - 1}
            GR[prv] = out
            i = prv
        for i in range(L - 1, -1, -1):
            prv = (i - 1) % L
            out = self.planR[i].apply(GR[i], AR[i])
            if prv != L - 1:
                GR[prv] = out
                GR[prv].identity_levels = {chi - 1}
            else:
                Yfirst = out
        # GR[L-1] sits on bond L-1 (right of site L-1 == left of site 0): rho = C^T C there
        rhoR = BondTensor(C[L - 1].V, {c: b.T @ b for c, b in C[L - 1].blocks.items()})
        Yr = EnvTensor("R", AR[L - 1].Vr, triv)
        set_level(Yr, 0, Yfirst, 0)
        eR = bond_dot(Yr, rhoR, "R")
        oneR = unit_env("R", AR[L - 1].Vr, triv, 0)

        def opR(X):
            T = X
            for i in range(L - 1, -1, -1):
                T = idR[i].apply(T, AR[i])
            out = vcopy(X)
            vaxpy(-1.0, T, out)
            vaxpy(bond_dot(X, rhoR, "R"), oneR, out)
            return out

        rhs = vcopy(Yr)
        vaxpy(-eR, oneR, rhs)
        Xr, infoR = gmres(opR, rhs, tol=tol, krylovdim=krylovdim, maxiter=maxiter)
        set_level(GR[L - 1], 0, Xr, 0)
        for i in range(L - 1, 0, -1):
            GR[i - 1] = self.planR[i].apply(GR[i], AR[i])
            GR[i - 1].identity_levels = {chi - 1}
        self.GL, self.GR = GL, GR
        self.energy_cell_left, self.energy_cell_right = eL, eR
        self.energy_per_site = 0.5 * (eL + eR) / L
        self.info = dict(left=infoL, right=infoR, depth=depth)


# ----------------------------------------------------------------------------------------
# VUMPS
# ----------------------------------------------------------------------------------------
def _bond_as_vec(C):
    return C


class _HC:
    def __init__(self, GL, GR):
        self.GL, self.GR = GL, GR

    def __call__(self, C):
        return heff_c_apply(self.GL, self.GR, C)


def galerkin(state, envs, plans):
    """max_i || H_AC AC_i - AL_i (AL_i^T H_AC AC_i) ||   (MPSKit `calc_galerkin`)."""
    eps = 0.0
    for i in range(len(state["AL"])):
        AL, AC = state["AL"][i], state["AC"][i]
        y = plans[i].apply(AC)
        proj = BondTensor(AL.Vr)
        for k, v in AL.blocks.items():
            proj.blocks[k[2]] += v.T @ y.blocks[k]
        for k, v in AL.blocks.items():
            y.blocks[k] = y.blocks[k] - v @ proj.blocks[k[2]]
        eps = max(eps, vnorm(y))
    return eps


def vumps(state, W_list, tol=1e-10, maxiter=100, krylovdim=30, verbose=False, env_tol=None):
    """VUMPS ground-state search on fixed bond spaces.  Returns (state, envs, delta, log)."""
    L = len(W_list)
    state = dict(state)
    eps = 1.0
    log = []
    envs = Environments(state, W_list, tol=env_tol or 1e-10)
    for it in range(1, maxiter + 1):
        tol_eig = min(1e-4, max(eps * 1e-3, 1e-14))
        tol_env = env_tol or min(1e-6, max(eps * 1e-4, 1e-14))
        newAC, newC = [], []
        for i in range(L):
            plan = HeffACPlan(envs.GL[i], W_list[i], envs.GR[i], state["AC"][i])
            _, ac, _ = lanczos_lowest(plan.apply, state["AC"][i], tol=tol_eig, krylovdim=krylovdim, maxiter=20)
            if vdot(ac, state["AC"][i]) < 0:
                vscale(ac, -1.0)
            hc = _HC(envs.GL[(i + 1) % L], envs.GR[i])
            _, c, _ = lanczos_lowest(hc, state["C"][i], tol=tol_eig, krylovdim=krylovdim, maxiter=20)
            if vdot(c, state["C"][i]) < 0:
                vscale(c, -1.0)
            newAC.append(ac)
            newC.append(c)
        AL = [regauge(newAC[i], newC[i]) for i in range(L)]
        AR, C, ginfo = uniform_rightorth(AL, newC[L - 1], tol=min(1e-8, max(eps * 1e-6, 1e-14)), eig_miniter=2)
        state = dict(AL=AL, AR=AR, C=C, AC=[mul_right(AL[i], C[i]) for i in range(L)])
        envs = Environments(state, W_list, tol=tol_env)
        plans = [HeffACPlan(envs.GL[i], W_list[i], envs.GR[i], state["AC"][i]) for i in range(L)]
        eps = galerkin(state, envs, plans)
        log.append(dict(iter=it, galerkin=eps, energy=envs.energy_per_site, gauge_iters=ginfo["iterations"]))
        if verbose:
            print("vumps %3d  eps %.3e  E/site %.12f  gauge its %d" % (it, eps, envs.energy_per_site, ginfo["iterations"]))
        if eps < tol:
            break
    return state, envs, eps, log


# ----------------------------------------------------------------------------------------
# observables
# ----------------------------------------------------------------------------------------
def expval_diag(AC: MPSTensor, values) -> float:
    """<op> for a one-site operator that is a scalar values[s] on physical multiplet s (number
    operator: HF:316-323; the reference evaluates `expectation_value(psi, i => n)` HF:1507)."""
    num = sum(AC.Vr.dims[k[2]] * values[k[1]] * np.vdot(v, v) for k, v in AC.blocks.items())
    den = sum(AC.Vr.dims[k[2]] * np.vdot(v, v) for k, v in AC.blocks.items())
    return float(num / den)


def entanglement_spectrum(C: BondTensor):
    """Schmidt values per sector (each appears dim(c) times in the full spectrum)."""
    return {C.V.sectors[c]: np.linalg.svd(b, compute_uv=False) for c, b in C.blocks.items()}
