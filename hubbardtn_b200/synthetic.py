"""Synthetic models of the benchmark shapes.

BASELINE.json quotes its metric on "two-band MB_Sim with full U_ijkl, U(1)xSU(2), D=1024"
(config C4).  Until the multi-band term builders (HubbardFunctions.jl:476-910) are restated,
C4 is the synthetic instance SURVEY.md 8(d) defines: physical space of HF:251, two
alternating bond types with Gaussian-weighted multiplicities summing to D, chi = 96 MPO
levels (1 + 48 fermionic (1,1/2,+-1) + 46 bosonic + 1) and random reduced W entries in an
upper-triangular pattern with 4 non-zero level pairs per row.  Throughput does not depend on
the values.
"""
from __future__ import annotations

import math

import numpy as np

from . import sectors as S

SEED = 20261018


def bond_space(sym: int, D: int, btype: int) -> dict:
    """{sector: multiplicity}: n_c = round(D w_c / sum w), remainder to the largest sector."""
    w = {}
    if sym == S.SU2U1:
        for tj in range(0, 7):
            p = tj & 1
            for n in range(-4, 5):
                if ((n & 1) == p) != (btype == 0):
                    continue
                j = tj / 2.0
                w[(p, tj, n)] = (tj + 1) * math.exp(-(j + 0.5) ** 2 / 2.0) * math.exp(-n * n / (2 * 1.2 ** 2))
    else:
        for m in range(-5, 6):
            p = m & 1
            for n in range(-4, 5):
                if ((n & 1) == p) != (btype == 0):
                    continue
                w[(p, m, n)] = math.exp(-m * m / (2 * 1.3 ** 2)) * math.exp(-n * n / (2 * 1.2 ** 2))
    tot = sum(w.values())
    mult = {s: int(round(D * v / tot)) for s, v in w.items()}
    mult = {s: n for s, n in mult.items() if n > 0}
    if not mult:                       # D too small for any rounded multiplicity
        mult = {max(w, key=lambda s: w[s]): 0}
    big = max(mult, key=lambda s: (mult[s], -abs(s[2]), -abs(s[1])))
    mult[big] += D - sum(mult.values())
    return mult


def mpo_levels(sym: int, chi: int) -> list:
    """[trivial] + (chi-2) intermediate levels + [trivial]; ~half fermionic hop channels."""
    if chi < 2:
        raise ValueError("chi must be >= 2")
    inner = chi - 2
    nferm = (inner + 1) // 2
    if sym == S.SU2U1:
        ferm = [(1, 1, 1), (1, 1, -1)]
        bos = [(0, 0, 0), (0, 2, 0), (0, 0, 2), (0, 0, -2), (0, 2, 2), (0, 2, -2)]
    else:
        ferm = [(1, 1, 1), (1, -1, 1), (1, 1, -1), (1, -1, -1)]
        bos = [(0, 0, 0), (0, 2, 0), (0, -2, 0), (0, 0, 2), (0, 0, -2), (0, 2, 2), (0, -2, -2)]
    levels = [(0, 0, 0)]
    levels += [ferm[i % len(ferm)] for i in range(nferm)]
    levels += [bos[i % len(bos)] for i in range(inner - nferm)]
    levels.append((0, 0, 0))
    return levels


def mpo_entries(sym: int, levels: list, phys: list, nnz_per_row: int = 4, seed: int = SEED) -> dict:
    """Random reduced W entries {(a,s',s,b,c): w}: W[0,0] = W[chi-1,chi-1] = 1 (Jordan form),
    every row a < chi-1 connects to `nnz_per_row` levels b > a (the closing level chi-1
    first, when the charges allow it)."""
    rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, 7]))
    chi = len(levels)

    def connects(a, b):
        for csp in phys:
            for c in S.fuse(sym, levels[a], csp):
                for cs in phys:
                    if S.allowed(sym, cs, levels[b], c):
                        return True
        return False

    pairs = set()
    for a in range(chi - 1):
        cand = [b for b in range(a + 1, chi) if connects(a, b)]
        chosen = []
        if chi - 1 in cand:
            chosen.append(chi - 1)
            cand.remove(chi - 1)
        k = min(len(cand), nnz_per_row - len(chosen))
        if k > 0:
            chosen += [int(v) for v in rng.choice(cand, size=k, replace=False)]
        pairs.update((a, b) for b in chosen)
    entries = {}
    for key in S.mpo_entry_keys(sym, levels, phys, levels, pairs):
        entries[key] = float(rng.standard_normal())
    for s, cs in enumerate(phys):          # identity on the first and last level
        entries[(0, s, s, 0, cs)] = 1.0
        entries[(chi - 1, s, s, chi - 1, cs)] = 1.0
    return entries


def random_packed(nelem: int, stream: int, seed: int = SEED) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, 100 + stream]))
    return rng.standard_normal(nelem)


class HeffCase:
    """Device-resident synthetic H_AC problem: spaces, GL, W, GR, x, y and the plan."""

    def __init__(self, ctx, sym: int = S.SU2U1, D: int = 1024, chi: int = 96, nnz_per_row: int = 4,
                 seed: int = SEED, site: int = 0, spaces=None):
        """`spaces` = (vl_mult, vr_mult, phys, levels) overrides the SURVEY 8(d) recipe (used by
        tools/ and tests to build special shapes, e.g. one dense sector)."""
        from . import device as dev
        self.ctx, self.sym, self.D, self.chi, self.site = ctx, sym, D, chi, site
        if spaces is not None:
            self.vl_mult, self.vr_mult, self.phys, self.levels = spaces
            chi = self.chi = len(self.levels)
        else:
            self.phys = S.physical_space(sym, 1, 1)
            # site parity picks the (left,right) bond types: A|B on even sites, B|A on odd ones
            self.vl_mult = bond_space(sym, D, site & 1)
            self.vr_mult = bond_space(sym, D, 1 - (site & 1))
            self.levels = mpo_levels(sym, chi)
        self.w_entries = mpo_entries(sym, self.levels, self.phys, nnz_per_row, seed + site)
        self.Vl = dev.Space(ctx, sym, self.vl_mult)
        self.Vr = dev.Space(ctx, sym, self.vr_mult)
        self.P = dev.Legs(ctx, sym, self.phys)
        self.M = dev.Legs(ctx, sym, self.levels)
        self.GL = dev.Tensor.env(ctx, 0, self.Vl, self.M, identity_level=0)
        self.GR = dev.Tensor.env(ctx, 1, self.Vr, self.M, identity_level=chi - 1)
        self.x = dev.Tensor.mps(ctx, self.Vl, self.P, self.Vr)
        self.y = self.x.like()
        self.W = dev.Mpo(ctx, self.M, self.P, self.M, self.w_entries)
        # data: iid N(0,1), scaled so that y stays O(1); identity levels hold the unit tensor
        self.gl_host = random_packed(self.GL.nelem, 1 + 10 * site, seed) / math.sqrt(max(D, 1))
        self.gr_host = random_packed(self.GR.nelem, 2 + 10 * site, seed) / math.sqrt(max(D, 1))
        self.x_host = random_packed(self.x.nelem, 3 + 10 * site, seed)
        for (a, i, j), blk in self.GL.block_views(self.gl_host).items():
            if a == 0:
                blk[...] = np.eye(blk.shape[0])
        for (b, i, j), blk in self.GR.block_views(self.gr_host).items():
            if b == chi - 1:
                blk[...] = np.eye(blk.shape[0])
        self.GL.upload(self.gl_host)
        self.GR.upload(self.gr_host)
        self.x.upload(self.x_host)
        self.plan = dev.HeffAC(ctx, self.GL, self.W, self.GR, self.x)
