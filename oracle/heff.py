"""Effective-Hamiltonian applications and environment transfers in reduced form (oracle).

Restates what MPSKit 0.13.1's derivative operators do for HubbardTN
(src/HubbardFunctions.jl:1010,1012,1017,1027 -> `find_groundstate` -> `∂AC`, `∂C`,
`∂AC2`, `TransferMatrix`; MPSKit/TensorKit are not vendored):

    H_AC x  = sum_{a,b} GL[a] . x . W[a,b] . GR[b]
    H_C  x  = sum_a     GL[a] . x . GR[a]
    T_L(GL) = sum_a  conj(A) . GL[a] . W[a,b] . A        (left environment growth)
    T_R(GR) = sum_b  A . W[a,b] . GR[b] . conj(A)        (right environment growth)

In reduced form every term is  coef * GL[a,l',l] @ x[l,s,r] @ GR[b,r,r']  with a scalar
recoupling coefficient.  ALL coefficients derive from one number, the full contraction
of the six Clebsch-Gordan tensors of the network (computed numerically here, no 6j
tables):

    N = sum_m CG(l',s'|r') CG(a,l|l') CG(l,s|r) CG(a,s'|c) CG(s,b|c) CG(b,r|r')
    H_AC, T_L : coef = w * N / dim(r')         T_R : coef = w * N / dim(l')

`heff_ac_apply_naive` is the defining triple product per term; `HeffACPlan` is the
staged algorithm the CUDA path implements (stage L: T = GL.x once per (a,l',l,s,r);
stage W: U = sum coef T; stage R: y += U.GR) and is what the CPU baseline times.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

from . import sectors as S
from .tensors import BondTensor, EnvTensor, MPOTensor, MPSTensor


@lru_cache(maxsize=None)
def _network_su2(jlp, jsp, jrp, jl, js, jr, ja, jb, jc) -> float:
    g = S._cg_su2
    v = float(np.einsum(
        "ptq,alp,lsr,atc,sbc,brq->",
        g(jlp, jsp, jrp), g(ja, jl, jlp), g(jl, js, jr), g(ja, jsp, jc), g(js, jb, jc),
        g(jb, jr, jrp), optimize=True))
    return 0.0 if abs(v) < 1e-14 else v   # same zero threshold as csrc/htn_sectors.cpp


def network(kind, clp, csp, crp, cl, cs, cr, ca, cb, cc) -> float:
    """Full contraction N of the six coupling tensors of the H_AC network (0 if any
    vertex is not allowed)."""
    ok = (S.allowed(kind, clp, csp, crp) and S.allowed(kind, ca, cl, clp)
          and S.allowed(kind, cl, cs, cr) and S.allowed(kind, ca, csp, cc)
          and S.allowed(kind, cs, cb, cc) and S.allowed(kind, cb, cr, crp))
    if not ok:
        return 0.0
    if kind == S.SU2U1:
        return _network_su2(clp[1], csp[1], crp[1], cl[1], cs[1], cr[1], ca[1], cb[1], cc[1])
    return 1.0


# ----------------------------------------------------------------------------------------
# H_AC
# ----------------------------------------------------------------------------------------

def _env_partners(env: EnvTensor):
    """(level, in) -> list of out positions: GL: (a,l)->[l'], GR: (b,r)->[r']."""
    out = {}
    if env.side == "L":
        for (a, lp, l) in env.blocks:
            out.setdefault((a, l), []).append(lp)
    else:
        for (b, r, rp) in env.blocks:
            out.setdefault((b, r), []).append(rp)
    return out


def heff_ac_terms(GL: EnvTensor, W: MPOTensor, GR: EnvTensor, x: MPSTensor):
    """Enumerate ((l',s',r'), (a,l',l), (l,s,r), (b,r,r'), coef) with the sum over the
    MPO coupled sector c already carried out."""
    k = x.kind
    Vl, P, Vr = x.Vl, x.P, x.Vr
    pl, pr = _env_partners(GL), _env_partners(GR)
    xs = {}
    for (l, s, r) in x.blocks:
        xs.setdefault(s, []).append((l, r))
    acc = {}
    for (a, sp, s, b, c), w in W.entries.items():
        for (l, r) in xs.get(s, ()):
            for lp in pl.get((a, l), ()):
                for rp in pr.get((b, r), ()):
                    if (lp, sp, rp) not in x.blocks:
                        continue
                    n = network(k, Vl.sectors[lp], P.sectors[sp], Vr.sectors[rp],
                                Vl.sectors[l], P.sectors[s], Vr.sectors[r],
                                W.Ml.sectors[a], W.Mr.sectors[b], c)
                    if n == 0.0:
                        continue
                    key = ((lp, sp, rp), (a, lp, l), (l, s, r), (b, r, rp))
                    acc[key] = acc.get(key, 0.0) + w * n / Vr.dims[rp]
    return [(k0, k1, k2, k3, cf) for (k0, k1, k2, k3), cf in acc.items() if cf != 0.0]


def heff_ac_apply_naive(GL, W, GR, x: MPSTensor) -> MPSTensor:
    y = x.zeros_like()
    for (ky, kgl, kx, kgr, cf) in heff_ac_terms(GL, W, GR, x):
        y.blocks[ky] += cf * (GL.blocks[kgl] @ x.blocks[kx] @ GR.blocks[kgr])
    return y


def heff_ac_apply_dense(GL, W, GR, x: MPSTensor) -> np.ndarray:
    """Symmetry-free evaluation: y[p,t,q] = GL[p,a,l] x[l,s,r] W[a,t,s,b] GR[r,b,q]."""
    return np.einsum("pal,lsr,atsb,rbq->ptq", GL.to_dense(), x.to_dense(), W.to_dense(),
                     GR.to_dense(), optimize=True)


class HeffACPlan:
    """Staged H_AC apply (the algorithm of the CUDA path; CPU baseline).

    stage L : T[a,l',l,s,r]    = GL[a,l',l] @ x[l,s,r]       (skipped for identity levels a)
    stage W : U[b,l',s',r',r]  = sum coef * T[...]            (or straight into y when b is
                                                                an identity level)
    stage R : y[l',s',r']     += U[b,l',s',r',r] @ GR[b,r,r']

    flops = sum 2 m n k over the stage-L and stage-R GEMM lists.
    """

    def __init__(self, GL: EnvTensor, W: MPOTensor, GR: EnvTensor, x: MPSTensor):
        self.GL, self.GR, self.x0 = GL, GR, x
        terms = heff_ac_terms(GL, W, GR, x)
        idL, idR = GL.identity_levels, GR.identity_levels
        self.t_list = []      # (a,lp,l,s,r) needing a GEMM
        t_index = {}
        self.u_list = []      # (b,lp,sp,rp,r) needing a GEMM
        u_index = {}
        self.mix = {}         # target ('U',i) / ('Y',key) -> list of (('T',i)/('X',key), coef)
        for (ky, kgl, kx, kgr, cf) in terms:
            a, lp, l = kgl
            _, s, r = kx
            b, _, rp = kgr
            sp = ky[1]
            if a in idL:
                src = ("X", kx)
            else:
                tk = (a, lp, l, s, r)
                if tk not in t_index:
                    t_index[tk] = len(self.t_list)
                    self.t_list.append(tk)
                src = ("T", t_index[tk])
            if b in idR:
                dst = ("Y", ky)
            else:
                uk = (b, lp, sp, rp, r)
                if uk not in u_index:
                    u_index[uk] = len(self.u_list)
                    self.u_list.append(uk)
                dst = ("U", u_index[uk])
            self.mix.setdefault(dst, []).append((src, cf))
        Vl, Vr = x.Vl, x.Vr
        self.flops_L = sum(2 * Vl.mult[lp] * Vl.mult[l] * Vr.mult[r]
                           for (a, lp, l, s, r) in self.t_list)
        self.flops_R = sum(2 * Vl.mult[lp] * Vr.mult[r] * Vr.mult[rp]
                           for (b, lp, sp, rp, r) in self.u_list)
        self.flops = self.flops_L + self.flops_R

    def apply(self, x: MPSTensor, threads: int = 1) -> MPSTensor:
        """threads > 1: worker threads over sector blocks with single-threaded BLAS — the
        reference's CPU policy (HubbardFunctions.jl:29 BLAS threads = 1, :37 scheduler)."""
        GL, GR = self.GL, self.GR
        if threads > 1:
            from concurrent.futures import ThreadPoolExecutor
            pool = ThreadPoolExecutor(threads)
            pmap = lambda f, it: list(pool.map(f, it, chunksize=8))  # noqa: E731
        else:
            pool = None
            pmap = lambda f, it: [f(v) for v in it]  # noqa: E731
        T = pmap(lambda k: GL.blocks[(k[0], k[1], k[2])] @ x.blocks[(k[2], k[3], k[4])], self.t_list)

        def mix_one(item):
            dst, srcs = item
            acc = None
            for (kind, i), cf in srcs:
                blk = T[i] if kind == "T" else x.blocks[i]
                acc = cf * blk if acc is None else acc + cf * blk
            return dst, acc

        y = x.zeros_like()
        U = [None] * len(self.u_list)
        for dst, acc in pmap(mix_one, list(self.mix.items())):
            if dst[0] == "Y":
                y.blocks[dst[1]] += acc
            else:
                U[dst[1]] = acc
        if not hasattr(self, "_by_y"):
            self._by_y = {}
            for i, (b, lp, sp, rp, r) in enumerate(self.u_list):
                self._by_y.setdefault((lp, sp, rp), []).append((i, (b, r, rp)))

        def stage_r(item):
            ky, lst = item
            acc = y.blocks[ky]
            for i, kgr in lst:
                acc += U[i] @ GR.blocks[kgr]

        pmap(stage_r, list(self._by_y.items()))
        if pool is not None:
            pool.shutdown()
        return y


# ----------------------------------------------------------------------------------------
# H_C
# ----------------------------------------------------------------------------------------

def heff_c_apply(GL: EnvTensor, GR: EnvTensor, C: BondTensor) -> BondTensor:
    """y[c'] = sum_{a,c} GL[a,c',c] @ C[c] @ GR[a,c,c'] (all recoupling coefficients are 1:
    sum_m CG(a,c|c')CG(a,c|c') = delta)."""
    y = BondTensor(C.V)
    for (a, lp, l), g in GL.blocks.items():
        kr = (a, l, lp)
        if kr in GR.blocks:
            y.blocks[lp] += g @ C.blocks[l] @ GR.blocks[kr]
    return y


def heff_c_apply_dense(GL, GR, C: BondTensor) -> np.ndarray:
    return np.einsum("pal,lr,raq->pq", GL.to_dense(), C.to_dense(), GR.to_dense(), optimize=True)


# ----------------------------------------------------------------------------------------
# environment transfers
# ----------------------------------------------------------------------------------------

def transfer_left(GL: EnvTensor, W: MPOTensor, A: MPSTensor, Abar: MPSTensor = None) -> EnvTensor:
    """GL'[b,r',r] = sum coef * Abar[l',s',r']^T @ GL[a,l',l] @ A[l,s,r], coef = w N / dim(r')."""
    Abar = A if Abar is None else Abar
    out = EnvTensor("L", A.Vr, W.Mr)
    k = A.kind
    Vl, P, Vr = A.Vl, A.P, A.Vr
    pl = _env_partners(GL)
    by_lps = {}
    for (lp, sp, rp) in Abar.blocks:
        by_lps.setdefault((lp, sp), []).append(rp)
    for (a, sp, s, b, c), w in W.entries.items():
        for (l, s2, r), ablk in A.blocks.items():
            if s2 != s:
                continue
            for lp in pl.get((a, l), ()):
                for rp in by_lps.get((lp, sp), ()):
                    if (b, rp, r) not in out.blocks:
                        continue
                    n = network(k, Vl.sectors[lp], P.sectors[sp], Vr.sectors[rp],
                                Vl.sectors[l], P.sectors[s], Vr.sectors[r],
                                W.Ml.sectors[a], W.Mr.sectors[b], c)
                    if n == 0.0:
                        continue
                    out.blocks[(b, rp, r)] += (w * n / Vr.dims[rp]) * (
                        Abar.blocks[(lp, sp, rp)].T @ GL.blocks[(a, lp, l)] @ ablk)
    return out


def transfer_right(GR: EnvTensor, W: MPOTensor, A: MPSTensor, Abar: MPSTensor = None) -> EnvTensor:
    """GR'[a,l,l'] = sum coef * A[l,s,r] @ GR[b,r,r'] @ Abar[l',s',r']^T, coef = w N / dim(l')."""
    Abar = A if Abar is None else Abar
    out = EnvTensor("R", A.Vl, W.Ml)
    k = A.kind
    Vl, P, Vr = A.Vl, A.P, A.Vr
    pr = _env_partners(GR)
    by_sprp = {}
    for (lp, sp, rp) in Abar.blocks:
        by_sprp.setdefault((sp, rp), []).append(lp)
    for (a, sp, s, b, c), w in W.entries.items():
        for (l, s2, r), ablk in A.blocks.items():
            if s2 != s:
                continue
            for rp in pr.get((b, r), ()):
                for lp in by_sprp.get((sp, rp), ()):
                    if (a, l, lp) not in out.blocks:
                        continue
                    n = network(k, Vl.sectors[lp], P.sectors[sp], Vr.sectors[rp],
                                Vl.sectors[l], P.sectors[s], Vr.sectors[r],
                                W.Ml.sectors[a], W.Mr.sectors[b], c)
                    if n == 0.0:
                        continue
                    out.blocks[(a, l, lp)] += (w * n / Vl.dims[lp]) * (
                        ablk @ GR.blocks[(b, r, rp)] @ Abar.blocks[(lp, sp, rp)].T)
    return out


def transfer_left_dense(GL, W, A, Abar=None) -> np.ndarray:
    Ad = A.to_dense()
    Bd = Ad if Abar is None else Abar.to_dense()
    return np.einsum("ptq,pal,atsb,lsr->qbr", Bd, GL.to_dense(), W.to_dense(), Ad, optimize=True)


def transfer_right_dense(GR, W, A, Abar=None) -> np.ndarray:
    Ad = A.to_dense()
    Bd = Ad if Abar is None else Abar.to_dense()
    return np.einsum("lsr,atsb,rbq,ptq->lap", Ad, W.to_dense(), GR.to_dense(), Bd, optimize=True)
