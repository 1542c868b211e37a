"""Block-sparse symmetric tensors in reduced (Wigner-Eckart) form + dense expansion.

Oracle restatement (test infrastructure) of the storage TensorKit 0.14.6 gives the
objects of the hot path (TensorKit is not vendored):

  bond space  V   : sorted sectors c with multiplicity n_c          (GradedSpace)
  leg list    P/M : ordered list of sectors, one multiplet each     (physical space
                    HubbardFunctions.jl:248,251,343; MPO virtual levels = SumSpace)
  MPS tensor  A   : blocks (l,s,r) -> [n_l, n_r],  r in l(x)s       ((V_l (x) P) <- V_r)
  bond matrix C   : blocks c -> [n_c, n_c]
  left env    GL  : blocks (a,l',l) -> [n_l', n_l], l' in a(x)l     (bra, MPO level, ket)
  right env   GR  : blocks (b,r,r') -> [n_r, n_r'], r' in b(x)r     (ket, MPO level, bra)
  MPO tensor  W   : entries (a,s',s,b,c) -> scalar, c in a(x)s' and c in s(x)b
                    ((M_l (x) P) <- (P (x) M_r) with coupled sector c)

Index letters are POSITIONS in the corresponding sector list (not labels).  The full
(symmetry-free) tensors are obtained by attaching Clebsch-Gordan tensors:

  A_full [l am, s m_s, r bm]  = A[l,s,r][a,b]     * CG(l,s|r)
  GL_full[l' a'm', a m_a, l am] = GL[a,l',l][a',a] * CG(a,l|l')
  GR_full[r bm, b m_b, r' b'm'] = GR[b,r,r'][b,b'] * CG(b,r|r')
  W_full [a m_a, s' m', s m, b m_b] = sum_c w * sum_mc CG(a,s'|c) CG(s,b|c)

with inner product <x,y> = sum_blocks dim(r) tr(x^T y)  (SURVEY.md App. A).
"""
from __future__ import annotations

import numpy as np

from . import sectors as S


class Space:
    """Graded bond space: sorted sectors with multiplicities."""

    def __init__(self, kind: int, mults: dict):
        self.kind = kind
        items = [(s, int(n)) for s, n in mults.items() if n > 0]
        items.sort(key=lambda t: S.sort_key(kind, t[0]))
        self.sectors = [s for s, _ in items]
        self.mult = [n for _, n in items]
        self.index = {s: i for i, s in enumerate(self.sectors)}
        self.dims = [S.dim(kind, s) for s in self.sectors]
        off, acc = [], 0
        for n, d in zip(self.mult, self.dims):
            off.append(acc)
            acc += n * d
        self.full_offset = off
        self.full_dim = acc          # D_full = sum dim(c) n_c  (hf.dim_state, HF:1402)
        self.red_dim = sum(self.mult)  # D_red  = sum n_c

    def __len__(self):
        return len(self.sectors)

    def as_dict(self):
        return dict(zip(self.sectors, self.mult))

    def __eq__(self, other):
        return (isinstance(other, Space) and self.kind == other.kind
                and self.sectors == other.sectors and self.mult == other.mult)

    def __repr__(self):
        return "Space(%s)" % ", ".join("%s=>%d" % (s, n) for s, n in zip(self.sectors, self.mult))


class Legs:
    """Ordered list of single multiplets (physical space or MPO virtual levels)."""

    def __init__(self, kind: int, sectors):
        self.kind = kind
        self.sectors = [tuple(s) for s in sectors]
        self.dims = [S.dim(kind, s) for s in self.sectors]
        off, acc = [], 0
        for d in self.dims:
            off.append(acc)
            acc += d
        self.full_offset = off
        self.full_dim = acc

    def __len__(self):
        return len(self.sectors)


# ----------------------------------------------------------------------------------------
# block key enumeration (canonical orders; these DEFINE the block tables of this repo)
# ----------------------------------------------------------------------------------------

def mps_keys(Vl: Space, P: Legs, Vr: Space):
    """Allowed (l,s,r) blocks, ordered by coupled sector r, then s, then l (mirrors the
    TensorKit block for coupled sector c=r whose rows run over fusion trees (l,s)->c)."""
    k = Vl.kind
    keys = []
    for r, cr in enumerate(Vr.sectors):
        for s, cs in enumerate(P.sectors):
            for l, cl in enumerate(Vl.sectors):
                if S.allowed(k, cl, cs, cr):
                    keys.append((l, s, r))
    return keys


def envl_keys(V: Space, M: Legs):
    """Allowed (a,l',l) blocks of a left environment: l' in a (x) l."""
    k = V.kind
    return [(a, lp, l) for a, ca in enumerate(M.sectors) for lp, clp in enumerate(V.sectors)
            for l, cl in enumerate(V.sectors) if S.allowed(k, ca, cl, clp)]


def envr_keys(V: Space, M: Legs):
    """Allowed (b,r,r') blocks of a right environment: r' in b (x) r."""
    k = V.kind
    return [(b, r, rp) for b, cb in enumerate(M.sectors) for r, cr in enumerate(V.sectors)
            for rp, crp in enumerate(V.sectors) if S.allowed(k, cb, cr, crp)]


def mpo_keys(Ml: Legs, P: Legs, Mr: Legs):
    """All symmetry-allowed reduced entries (a,s',s,b,c) of an MPO tensor."""
    k = P.kind
    keys = []
    for a, ca in enumerate(Ml.sectors):
        for sp, csp in enumerate(P.sectors):
            for c in S.fuse(k, ca, csp):
                for s, cs in enumerate(P.sectors):
                    for b, cb in enumerate(Mr.sectors):
                        if S.allowed(k, cs, cb, c):
                            keys.append((a, sp, s, b, c))
    return keys


# ----------------------------------------------------------------------------------------
# containers
# ----------------------------------------------------------------------------------------

class MPSTensor:
    def __init__(self, Vl: Space, P: Legs, Vr: Space, blocks=None):
        self.Vl, self.P, self.Vr = Vl, P, Vr
        self.kind = Vl.kind
        self.keys = mps_keys(Vl, P, Vr)
        self.blocks = {} if blocks is None else blocks
        if blocks is None:
            for (l, s, r) in self.keys:
                self.blocks[(l, s, r)] = np.zeros((Vl.mult[l], Vr.mult[r]))

    def zeros_like(self):
        return MPSTensor(self.Vl, self.P, self.Vr)

    def copy(self):
        return MPSTensor(self.Vl, self.P, self.Vr, {k: v.copy() for k, v in self.blocks.items()})

    def weight(self, key):
        return self.Vr.dims[key[2]]

    def nelem(self):
        return sum(v.size for v in self.blocks.values())

    def randomize(self, rng):
        for k in self.keys:
            self.blocks[k] = rng.standard_normal(self.blocks[k].shape)
        return self

    def to_dense(self):
        out = np.zeros((self.Vl.full_dim, self.P.full_dim, self.Vr.full_dim))
        k = self.kind
        for (l, s, r), blk in self.blocks.items():
            cl, cs, cr = self.Vl.sectors[l], self.P.sectors[s], self.Vr.sectors[r]
            g = S.cg(k, cl, cs, cr)
            dl, ds, dr = g.shape
            nl, nr = blk.shape
            t = np.einsum("ab,xyz->axybz", blk, g).reshape(nl * dl, ds, nr * dr)
            ol, os_, orr = self.Vl.full_offset[l], self.P.full_offset[s], self.Vr.full_offset[r]
            out[ol:ol + nl * dl, os_:os_ + ds, orr:orr + nr * dr] += t
        return out


class BondTensor:
    def __init__(self, V: Space, blocks=None):
        self.V = V
        self.kind = V.kind
        self.blocks = blocks if blocks is not None else {
            c: np.zeros((n, n)) for c, n in enumerate(V.mult)}

    def copy(self):
        return BondTensor(self.V, {k: v.copy() for k, v in self.blocks.items()})

    def weight(self, key):
        return self.V.dims[key]

    def to_dense(self):
        out = np.zeros((self.V.full_dim, self.V.full_dim))
        for c, blk in self.blocks.items():
            d, n, o = self.V.dims[c], self.V.mult[c], self.V.full_offset[c]
            out[o:o + n * d, o:o + n * d] = np.kron(blk, np.eye(d))
        return out


class EnvTensor:
    """side 'L': blocks (a,l',l) [n_l', n_l];  side 'R': blocks (b,r,r') [n_r, n_r']."""

    def __init__(self, side: str, V: Space, M: Legs, blocks=None, identity_levels=()):
        assert side in ("L", "R")
        self.side, self.V, self.M = side, V, M
        self.kind = V.kind
        self.keys = envl_keys(V, M) if side == "L" else envr_keys(V, M)
        self.identity_levels = set(identity_levels)
        self.blocks = {} if blocks is None else blocks
        if blocks is None:
            for key in self.keys:
                self.blocks[key] = np.zeros(self.shape(key))

    def shape(self, key):
        return (self.V.mult[key[1]], self.V.mult[key[2]])

    def randomize(self, rng):
        for key in self.keys:
            self.blocks[key] = rng.standard_normal(self.shape(key))
        self.fix_identity_levels()
        return self

    def fix_identity_levels(self):
        """Levels flagged identity hold the unit tensor (GL[1] = 1 in left-canonical
        gauge, GR[chi] = 1 in right-canonical gauge)."""
        for key in self.keys:
            if key[0] in self.identity_levels:
                assert key[1] == key[2]
                self.blocks[key] = np.eye(self.V.mult[key[1]])

    def to_dense(self):
        V, M, k = self.V, self.M, self.kind
        if self.side == "L":
            out = np.zeros((V.full_dim, M.full_dim, V.full_dim))      # [l' , a, l]
            for (a, lp, l), blk in self.blocks.items():
                g = S.cg(k, M.sectors[a], V.sectors[l], V.sectors[lp])  # [ma, ml, mlp]
                da, dl, dlp = g.shape
                nlp, nl = blk.shape
                t = np.einsum("pq,xyz->pzxqy", blk, g).reshape(nlp * dlp, da, nl * dl)
                o1, o2, o3 = V.full_offset[lp], M.full_offset[a], V.full_offset[l]
                out[o1:o1 + nlp * dlp, o2:o2 + da, o3:o3 + nl * dl] += t
        else:
            out = np.zeros((V.full_dim, M.full_dim, V.full_dim))      # [r, b, r']
            for (b, r, rp), blk in self.blocks.items():
                g = S.cg(k, M.sectors[b], V.sectors[r], V.sectors[rp])  # [mb, mr, mrp]
                db, dr, drp = g.shape
                nr, nrp = blk.shape
                t = np.einsum("pq,xyz->pyxqz", blk, g).reshape(nr * dr, db, nrp * drp)
                o1, o2, o3 = V.full_offset[r], M.full_offset[b], V.full_offset[rp]
                out[o1:o1 + nr * dr, o2:o2 + db, o3:o3 + nrp * drp] += t
        return out


class MPOTensor:
    """Reduced MPO tensor: entries[(a, s', s, b, c)] = w with c a sector LABEL."""

    def __init__(self, Ml: Legs, P: Legs, Mr: Legs, entries=None):
        self.Ml, self.P, self.Mr = Ml, P, Mr
        self.kind = P.kind
        self.entries = {} if entries is None else entries

    def randomize(self, rng, pattern=None):
        for key in mpo_keys(self.Ml, self.P, self.Mr):
            if pattern is None or (key[0], key[3]) in pattern:
                self.entries[key] = float(rng.standard_normal())
        return self

    def to_dense(self):
        Ml, P, Mr, k = self.Ml, self.P, self.Mr, self.kind
        out = np.zeros((Ml.full_dim, P.full_dim, P.full_dim, Mr.full_dim))  # [a, s', s, b]
        for (a, sp, s, b, c), w in self.entries.items():
            g1 = S.cg(k, Ml.sectors[a], P.sectors[sp], c)   # [ma, msp, mc]
            g2 = S.cg(k, P.sectors[s], Mr.sectors[b], c)    # [ms, mb, mc]
            t = w * np.einsum("xyc,zwc->xyzw", g1, g2)
            oa, osp, os_, ob = (Ml.full_offset[a], P.full_offset[sp], P.full_offset[s],
                                Mr.full_offset[b])
            da, dsp, ds, db = t.shape
            out[oa:oa + da, osp:osp + dsp, os_:os_ + ds, ob:ob + db] += t
        return out

    @staticmethod
    def from_dense(dense, Ml: Legs, P: Legs, Mr: Legs, tol=1e-12):
        """Wigner-Eckart projection of an invariant dense MPO tensor [a,s',s,b] onto its
        reduced entries; raises if `dense` is not invariant (re-expansion mismatch)."""
        k = P.kind
        W = MPOTensor(Ml, P, Mr)
        for key in mpo_keys(Ml, P, Mr):
            a, sp, s, b, c = key
            g1 = S.cg(k, Ml.sectors[a], P.sectors[sp], c)
            g2 = S.cg(k, P.sectors[s], Mr.sectors[b], c)
            oa, osp, os_, ob = (Ml.full_offset[a], P.full_offset[sp], P.full_offset[s],
                                Mr.full_offset[b])
            sub = dense[oa:oa + g1.shape[0], osp:osp + g1.shape[1],
                        os_:os_ + g2.shape[0], ob:ob + g2.shape[1]]
            w = np.einsum("xyzw,xyc,zwc->", sub, g1, g2) / S.dim(k, c)
            if abs(w) > tol:
                W.entries[key] = float(w)
        err = np.abs(W.to_dense() - dense).max() if dense.size else 0.0
        if err > 1e-10:
            raise ValueError("dense MPO tensor is not symmetric (re-expansion error %.3e)" % err)
        return W


# ----------------------------------------------------------------------------------------
# inner products / linear algebra on MPSTensor-shaped vectors
# ----------------------------------------------------------------------------------------

def inner(x, y) -> float:
    """<x,y> = sum_blocks dim(coupled sector) tr(x^T y)  (SURVEY.md App. A)."""
    return float(sum(x.weight(k) * np.vdot(x.blocks[k], y.blocks[k]) for k in x.blocks))


def norm(x) -> float:
    return float(np.sqrt(inner(x, x)))


def axpy(alpha: float, x, y):
    """y += alpha x (in place)."""
    for k in x.blocks:
        y.blocks[k] += alpha * x.blocks[k]
    return y


def scale(x, alpha: float):
    for k in x.blocks:
        x.blocks[k] *= alpha
    return x
