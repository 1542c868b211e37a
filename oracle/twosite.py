"""Two-site effective Hamiltonian, per-sector truncated SVD and IDMRG2 (oracle; test infrastructure).

Restates, from their published form (McCulloch, arXiv:0804.2509), the MPSKit
0.13.1 pieces HubbardTN reaches through `find_groundstate(psi, H, IDMRG2(trscheme=truncbelow(cut),
tol))` (src/HubbardFunctions.jl:1010) -- `∂∂AC2`, `tsvd!(..; trunc)`, the IDMRG2 sweep -- MPSKit and
TensorKit are not vendored (Manifest.toml:722,1156).

Two-site tensors live in the fusion-tree basis (l,s1 -> m), (m,s2 -> r):
    x2[l, s1, m, s2, r]  (n_l x n_r block; m is a sector LABEL: every m in l (x) s1 with r in m (x) s2)
    x2_full = sum_m x2[...] (x) CG(l,s1|m) CG(m,s2|r)
and the effective Hamiltonian is two nested copies of the H_AC recoupling network:
    y2[l',s1',m',s2',r'] = sum  w1 N(l',s1',m'; l,s1,m; a,b,c1)/dim(m')
                               * w2 N(m',s2',r'; m,s2,r; b,c,c2)/dim(r')
                               * GL[a,l',l] . x2[l,s1,m,s2,r] . GR[c,r,r']
"""
from __future__ import annotations

import numpy as np

from . import sectors as S
from .heff import network
from .krylov import lanczos_lowest, vdot
from .mps import Environments, TransferPlan, bond_norm, mul_left, mul_right, uniform_rightorth
from .tensors import BondTensor, EnvTensor, Legs, MPOTensor, MPSTensor, Space


def two_site_keys(Vl: Space, P1: Legs, P2: Legs, Vr: Space):
    """Canonical block order: coupled sector r, then s2, then m (canonical sector order), s1, l."""
    k = Vl.kind
    keys = []
    for r, cr in enumerate(Vr.sectors):
        for s2, cs2 in enumerate(P2.sectors):
            ms = set()
            for l, cl in enumerate(Vl.sectors):
                for s1, cs1 in enumerate(P1.sectors):
                    for m in S.fuse(k, cl, cs1):
                        if S.allowed(k, m, cs2, cr):
                            ms.add(m)
            for m in sorted(ms, key=lambda s: S.sort_key(k, s)):
                for s1, cs1 in enumerate(P1.sectors):
                    for l, cl in enumerate(Vl.sectors):
                        if S.allowed(k, cl, cs1, m):
                            keys.append((l, s1, m, s2, r))
    return keys


class TwoSiteTensor:
    def __init__(self, Vl: Space, P1: Legs, P2: Legs, Vr: Space, blocks=None):
        self.Vl, self.P1, self.P2, self.Vr = Vl, P1, P2, Vr
        self.kind = Vl.kind
        self.keys = two_site_keys(Vl, P1, P2, Vr)
        self.blocks = blocks if blocks is not None else {
            k: np.zeros((Vl.mult[k[0]], Vr.mult[k[4]])) for k in self.keys}

    def zeros_like(self):
        return TwoSiteTensor(self.Vl, self.P1, self.P2, self.Vr)

    def copy(self):
        return TwoSiteTensor(self.Vl, self.P1, self.P2, self.Vr, {k: v.copy() for k, v in self.blocks.items()})

    def weight(self, key):
        return self.Vr.dims[key[4]]

    def randomize(self, rng):
        for k in self.keys:
            self.blocks[k] = rng.standard_normal(self.blocks[k].shape)
        return self

    def nelem(self):
        return sum(v.size for v in self.blocks.values())

    def to_dense(self):
        k = self.kind
        out = np.zeros((self.Vl.full_dim, self.P1.full_dim, self.P2.full_dim, self.Vr.full_dim))
        for (l, s1, m, s2, r), blk in self.blocks.items():
            g1 = S.cg(k, self.Vl.sectors[l], self.P1.sectors[s1], m)          # [ml, m1, mm]
            g2 = S.cg(k, m, self.P2.sectors[s2], self.Vr.sectors[r])          # [mm, m2, mr]
            g = np.einsum("xyz,zuv->xyuv", g1, g2)
            nl, nr = blk.shape
            dl, d1, d2, dr = g.shape
            t = np.einsum("ab,xyuv->axyubv", blk, g).reshape(nl * dl, d1, d2, nr * dr)
            ol, o1, o2, orr = (self.Vl.full_offset[l], self.P1.full_offset[s1], self.P2.full_offset[s2],
                               self.Vr.full_offset[r])
            out[ol:ol + nl * dl, o1:o1 + d1, o2:o2 + d2, orr:orr + nr * dr] += t
        return out


def contract_two_site(A1: MPSTensor, A2: MPSTensor) -> TwoSiteTensor:
    """x2[l,s1,m,s2,r] = A1[l,s1,m] . A2[m,s2,r]  (m runs over the shared bond space)."""
    assert A1.Vr == A2.Vl
    x = TwoSiteTensor(A1.Vl, A1.P, A2.P, A2.Vr)
    Vm = A1.Vr
    for (l, s1, m, s2, r) in x.keys:
        mi = Vm.index.get(m)
        if mi is None:
            continue
        x.blocks[(l, s1, m, s2, r)] = A1.blocks[(l, s1, mi)] @ A2.blocks[(mi, s2, r)]
    return x


class HeffAC2Plan:
    """Term list of H_AC2 for fixed spaces; apply() evaluates the defining triple products."""

    def __init__(self, GL: EnvTensor, W1: MPOTensor, W2: MPOTensor, GR: EnvTensor, x: TwoSiteTensor):
        self.GL, self.GR = GL, GR
        k = x.kind
        Vl, Vr, P1, P2 = x.Vl, x.Vr, x.P1, x.P2
        pl, pr = {}, {}
        for (a, lp, l) in GL.blocks:
            pl.setdefault((a, l), []).append(lp)
        for (c, r, rp) in GR.blocks:
            pr.setdefault((c, r), []).append(rp)
        w1_by_s, w2_by_s = {}, {}
        for key, w in W1.entries.items():
            w1_by_s.setdefault(key[2], []).append((key, w))
        for key, w in W2.entries.items():
            w2_by_s.setdefault((key[0], key[2]), []).append((key, w))
        yset = set(x.keys)
        acc = {}
        for (l, s1, m, s2, r) in x.keys:
            cl, cr = Vl.sectors[l], Vr.sectors[r]
            for (a, s1p, _, b, c1), w1 in w1_by_s.get(s1, ()):
                for lp in pl.get((a, l), ()):
                    clp = Vl.sectors[lp]
                    for mp in S.fuse(k, clp, P1.sectors[s1p]):
                        n1 = network(k, clp, P1.sectors[s1p], mp, cl, P1.sectors[s1], m,
                                     W1.Ml.sectors[a], W1.Mr.sectors[b], c1)
                        if n1 == 0.0:
                            continue
                        for (_, s2p, _, c, c2), w2 in w2_by_s.get((b, s2), ()):
                            for rp in pr.get((c, r), ()):
                                crp = Vr.sectors[rp]
                                ky = (lp, s1p, mp, s2p, rp)
                                if ky not in yset:
                                    continue
                                n2 = network(k, mp, P2.sectors[s2p], crp, m, P2.sectors[s2], cr,
                                             W2.Ml.sectors[b], W2.Mr.sectors[c], c2)
                                if n2 == 0.0:
                                    continue
                                cf = (w1 * n1 / S.dim(k, mp)) * (w2 * n2 / S.dim(k, crp))
                                key = (ky, (a, lp, l), (l, s1, m, s2, r), (c, r, rp))
                                acc[key] = acc.get(key, 0.0) + cf
        self.terms = [(k0, k1, k2, k3, cf) for (k0, k1, k2, k3), cf in acc.items() if cf != 0.0]
        self.flops = sum(2 * Vl.mult[k1[1]] * Vl.mult[k1[2]] * Vr.mult[k3[1]] for (_, k1, _, k3, _) in self.terms)

    def apply(self, x: TwoSiteTensor) -> TwoSiteTensor:
        y = x.zeros_like()
        cache = {}
        for (ky, kgl, kx, kgr, cf) in self.terms:
            t = cache.get((kgl, kx))
            if t is None:
                t = cache[(kgl, kx)] = self.GL.blocks[kgl] @ x.blocks[kx]
            y.blocks[ky] += cf * (t @ self.GR.blocks[kgr])
        return y


def heff_ac2_apply_dense(GL, W1, W2, GR, x: TwoSiteTensor) -> np.ndarray:
    t = np.einsum("pal,lstr->pastr", GL.to_dense(), x.to_dense())
    t = np.einsum("pastr,aysb->pybtr", t, W1.to_dense())
    t = np.einsum("pybtr,bztc->pyzcr", t, W2.to_dense())
    return np.einsum("pyzcr,rcq->pyzq", t, GR.to_dense())


# ----------------------------------------------------------------------------------------
# truncated SVD per middle sector
# ----------------------------------------------------------------------------------------
def tsvd(x: TwoSiteTensor, cut: float = 0.0, maxdim: int = None):
    """x2 = AL . C . AR with C diagonal (Schmidt values), per middle sector m:
        M_m[(l,s1),(s2,r)] = sqrt(d_r/d_m) x2[l,s1,m,s2,r] = U S V^T
    `truncbelow(cut)` semantics (HF:1010): keep singular values >= cut; optional cap `maxdim` on the
    number of kept multiplets (largest first, across sectors).  Returns (AL, C, AR, info)."""
    k = x.kind
    Vl, Vr, P1, P2 = x.Vl, x.Vr, x.P1, x.P2
    by_m = {}
    for key in x.keys:
        by_m.setdefault(key[2], []).append(key)
    fac = {}
    for m, keys in by_m.items():
        rows = sorted({(kk[1], kk[0]) for kk in keys})           # (s1, l)
        cols = sorted({(kk[4], kk[3]) for kk in keys})           # (r, s2)
        ro, co = {}, {}
        o = 0
        for (s1, l) in rows:
            ro[(s1, l)] = o
            o += Vl.mult[l]
        nrow = o
        o = 0
        for (r, s2) in cols:
            co[(r, s2)] = o
            o += Vr.mult[r]
        ncol = o
        Mm = np.zeros((nrow, ncol))
        for (l, s1, _, s2, r) in keys:
            w = np.sqrt(Vr.dims[r] / S.dim(k, m))
            Mm[ro[(s1, l)]:ro[(s1, l)] + Vl.mult[l], co[(r, s2)]:co[(r, s2)] + Vr.mult[r]] = w * x.blocks[(l, s1, m, s2, r)]
        if min(Mm.shape) == 0:
            continue
        U, sv, Vt = np.linalg.svd(Mm, full_matrices=False)
        # deterministic sign: largest-magnitude entry of every left vector positive
        for j in range(U.shape[1]):
            i = np.argmax(np.abs(U[:, j]))
            if U[i, j] < 0:
                U[:, j] *= -1
                Vt[j, :] *= -1
        fac[m] = (U, sv, Vt, ro, co)
    # global truncation
    allsv = sorted(((sv_j, m) for m, f in fac.items() for sv_j in f[1]), key=lambda t: -t[0])
    nrm = np.sqrt(sum(S.dim(k, m) * sv_j ** 2 for sv_j, m in allsv))
    keep = {m: 0 for m in fac}
    kept = 0
    for sv_j, m in allsv:
        if sv_j < cut * nrm or sv_j <= 1e-14 * nrm or (maxdim is not None and kept >= maxdim):
            break
        keep[m] += 1
        kept += 1
    Vm = Space(k, {m: n for m, n in keep.items() if n > 0})
    AL = MPSTensor(Vl, P1, Vm)
    AR = MPSTensor(Vm, P2, Vr)
    Cb = BondTensor(Vm)
    disc = 0.0
    for m, (U, sv, Vt, ro, co) in fac.items():
        n = keep[m]
        disc += S.dim(k, m) * float(np.sum(sv[n:] ** 2))
        if n == 0:
            continue
        mi = Vm.index[m]
        Cb.blocks[mi] = np.diag(sv[:n])
        for (s1, l), o in ro.items():
            AL.blocks[(l, s1, mi)] = U[o:o + Vl.mult[l], :n]
        for (r, s2), o in co.items():
            w = np.sqrt(Vr.dims[r] / S.dim(k, m))
            AR.blocks[(mi, s2, r)] = Vt[:n, o:o + Vr.mult[r]] / w
    return AL, Cb, AR, dict(kept=kept, discarded_weight=disc / max(nrm ** 2, 1e-300), space=Vm)


# ----------------------------------------------------------------------------------------
# IDMRG2
# ----------------------------------------------------------------------------------------
def _inv_diag(C: BondTensor) -> BondTensor:
    """inv(C) block by block (MPSKit idmrg2: `inv(psi.C[end])`): the SVD-made bonds are diagonal, the caller's initial
    C[L-1] is a general matrix (triangular from the QR gauge, dense after VUMPS)."""
    out = BondTensor(C.V)
    for c, b in C.blocks.items():
        if np.count_nonzero(b - np.diag(np.diag(b))) == 0:
            out.blocks[c] = np.diag(1.0 / np.diag(b))
        else:
            out.blocks[c] = np.linalg.inv(b)
    return out


def _normalize_bond(C: BondTensor):
    n = bond_norm(C)
    for b in C.blocks.values():
        b /= n
    return C


def _unit_env(side, V, M, level):
    e = EnvTensor(side, V, M, identity_levels=[level])
    e.fix_identity_levels()
    return e


def idmrg2(state, W_list, cut=1e-2, tol=1e-6, maxiter=100, krylovdim=30, eig_tol=1e-8, maxdim=None, verbose=False,
           init_env="infinite"):
    """Two-site infinite DMRG (MPSKit `IDMRG2`): sweeps L->R, edge, R->L, edge over the unit cell
    (L >= 2), growing the environments by site transfers and the bond spaces by the truncated SVD.
    Convergence: || C_new - C_old || on the common subspace of the edge bond.  Returns
    (AL list, C list, AR list, eps, log)."""
    L = len(W_list)
    assert L >= 2
    AL, AR, AC, C = list(state["AL"]), list(state["AR"]), list(state["AC"]), list(state["C"])
    chi = len(W_list[0].Ml)
    if init_env == "infinite":
        # MPSKit: `find_groundstate(psi, H, alg::IDMRG2, envs = environments(psi, H))` -- the sweeps start
        # from the infinite environments of the initial uniform state, not from empty (unit) ones
        env0 = Environments(dict(AL=AL, AR=AR, C=C), W_list, tol=1e-10)
        GL, GR = list(env0.GL), list(env0.GR)
        for i in range(L):
            GL[i].identity_levels = {0}
            GR[i].identity_levels = {chi - 1}
    else:
        GL = [_unit_env("L", AL[i].Vl, W_list[i].Ml, 0) for i in range(L)]
        GR = [_unit_env("R", AR[i].Vr, W_list[i].Mr, chi - 1) for i in range(L)]
        # a first pass of transfers so that the environments contain one unit cell
        for i in range(L - 1):
            GL[i + 1] = TransferPlan("L", W_list[i], AL[i].Vl, AL[i].P, AL[i].Vr).apply(GL[i], AL[i])
            GL[i + 1].identity_levels = {0}
        for i in range(L - 1, 0, -1):
            GR[i - 1] = TransferPlan("R", W_list[i], AR[i].Vl, AR[i].P, AR[i].Vr).apply(GR[i], AR[i])
            GR[i - 1].identity_levels = {chi - 1}
    log = []
    eps = np.inf
    applies = 0

    def solve(i, j, x2):
        nonlocal applies
        plan = HeffAC2Plan(GL[i], W_list[i], W_list[j], GR[j], x2)
        ev, y, info = lanczos_lowest(plan.apply, x2, tol=eig_tol, krylovdim=krylovdim, maxiter=3)
        applies += info["applies"]
        if vdot(y, x2) < 0:
            for b in y.blocks.values():
                b *= -1
        return ev, y

    def grow_left(i):      # GL of site i+1 from site i
        j = (i + 1) % L
        GL[j] = TransferPlan("L", W_list[i], AL[i].Vl, AL[i].P, AL[i].Vr).apply(GL[i], AL[i])
        GL[j].identity_levels = {0}

    def grow_right(j):     # GR of site j-1 from site j
        i = (j - 1) % L
        GR[i] = TransferPlan("R", W_list[j], AR[j].Vl, AR[j].P, AR[j].Vr).apply(GR[j], AR[j])
        GR[i].identity_levels = {chi - 1}

    ev = 0.0
    for it in range(1, maxiter + 1):
        C_old = C[L - 1]
        # ---- left -> right ----
        for i in range(L - 1):
            ev, x2 = solve(i, i + 1, contract_two_site(AC[i], AR[i + 1]))
            al, c, ar, info = tsvd(x2, cut, maxdim)
            _normalize_bond(c)
            AL[i], C[i], AR[i + 1] = al, c, ar
            AC[i], AC[i + 1] = mul_right(al, c), mul_left(c, ar)
            grow_left(i)
            grow_right(i + 1)
        # ---- edge (sites L-1, 0) ----
        left = mul_right(AC[L - 1], _inv_diag(C[L - 1]))
        ev, x2 = solve(L - 1, 0, contract_two_site(left, mul_right(AL[0], C[0])))
        al, c, ar, info = tsvd(x2, cut, maxdim)
        _normalize_bond(c)
        AL[L - 1], C[L - 1], AR[0] = al, c, ar
        AC[L - 1], AC[0] = mul_right(al, c), mul_left(c, ar)
        AL[0] = mul_right(AC[0], _inv_diag(C[0]))
        grow_left(L - 1)
        grow_right(0)
        # ---- right -> left ----
        for i in range(L - 2, -1, -1):
            ev, x2 = solve(i, i + 1, contract_two_site(AL[i], AC[i + 1]))
            al, c, ar, info = tsvd(x2, cut, maxdim)
            _normalize_bond(c)
            AL[i], C[i], AR[i + 1] = al, c, ar
            AC[i], AC[i + 1] = mul_right(al, c), mul_left(c, ar)
            grow_left(i)
            grow_right(i + 1)
        # ---- edge again ----
        right = mul_left(_inv_diag(C[L - 1]), AC[0])
        ev, x2 = solve(L - 1, 0, contract_two_site(mul_left(C[L - 2], AR[L - 1]), right))
        al, c, ar, info = tsvd(x2, cut, maxdim)
        _normalize_bond(c)
        AL[L - 1], C[L - 1], AR[0] = al, c, ar
        AC[L - 1], AC[0] = mul_right(al, c), mul_left(c, ar)
        AR[L - 1] = mul_left(_inv_diag(C[L - 2]), AC[L - 1])
        grow_left(L - 1)
        grow_right(0)
        # ---- error: Schmidt values of the edge bond on the common subspace ----
        d2 = 0.0
        old = {C_old.V.sectors[k]: np.sort(np.abs(np.diag(b)))[::-1] if b.ndim == 2 else b for k, b in C_old.blocks.items()}
        old = {s: (np.linalg.svd(C_old.blocks[k], compute_uv=False)) for k, s in enumerate(C_old.V.sectors)}
        new = {s: np.diag(c.blocks[k]) for k, s in enumerate(c.V.sectors)}
        for s in set(old) | set(new):
            a, b = old.get(s, np.zeros(0)), new.get(s, np.zeros(0))
            n = min(len(a), len(b))
            d2 += S.dim(c.kind, s) * float(np.sum((a[:n] - b[:n]) ** 2))
        eps = float(np.sqrt(d2))
        log.append(dict(iter=it, eps=eps, D=[cc.V.red_dim for cc in C], applies=applies, eigenvalue=ev))
        if verbose:
            print("idmrg2 %3d eps %.3e  D_red %s  applies %d" % (it, eps, [cc.V.red_dim for cc in C], applies))
        if eps < tol:
            break
    return AL, C, AR, eps, log


def idmrg2_to_uniform(AR, C, tol=1e-12):
    """Gauge-fix the IDMRG2 result into a consistent uniform MPS, as MPSKit does with
    `InfiniteMPS(psi.AR)` at the end of IDMRG2 (the AR list is the one whose bond spaces chain after a
    full iteration): AL, C by the iterated QR from AR, then AR, C by the iterated LQ from AL."""
    from .mps import uniform_leftorth
    ALq, Cl, _ = uniform_leftorth(AR, C[-1], tol=tol)
    ARq, Cn, _ = uniform_rightorth(ALq, Cl[-1], tol=tol)
    return dict(AL=ALq, AR=ARq, C=Cn, AC=[mul_right(ALq[i], Cn[i]) for i in range(len(AR))])
