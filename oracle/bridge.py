"""Packed-array <-> oracle-tensor conversion (oracle side; test infrastructure).

The C ABI moves tensors as packed host arrays described by a block table
(`htn_tensor_blocktable`: label positions, rows, cols, offset).  These helpers rebuild the
oracle containers from such arrays so the CUDA path and the oracle see IDENTICAL inputs,
and check that the library's canonical block order equals the oracle's (the "block indexing
bit-exact" part of parity).
"""
from __future__ import annotations

import numpy as np

from .tensors import BondTensor, EnvTensor, Legs, MPOTensor, MPSTensor, Space


def _views(labels, rows, cols, offsets, packed):
    out = {}
    for i in range(len(rows)):
        o, r, c = int(offsets[i]), int(rows[i]), int(cols[i])
        out[tuple(int(v) for v in labels[i])] = packed[o:o + r * c].reshape(r, c)
    return out


def check_table(keys, shapes, labels, rows, cols, offsets):
    """Library block table == oracle canonical enumeration (order, shapes, packed offsets)."""
    assert len(keys) == len(rows), "block count differs: oracle %d, library %d" % (len(keys), len(rows))
    off = 0
    for i, key in enumerate(keys):
        assert tuple(int(v) for v in labels[i]) == tuple(key), "block %d label differs" % i
        assert (int(rows[i]), int(cols[i])) == tuple(shapes[i]), "block %d shape differs" % i
        assert int(offsets[i]) == off, "block %d offset differs" % i
        off += int(rows[i]) * int(cols[i])
    return off


def mps_from_packed(Vl: Space, P: Legs, Vr: Space, table, packed) -> MPSTensor:
    t = MPSTensor(Vl, P, Vr)
    check_table(t.keys, [t.blocks[k].shape for k in t.keys], *table)
    for k, v in _views(*table, packed).items():
        t.blocks[k] = np.array(v)
    return t


def mps_to_packed(t: MPSTensor, table) -> np.ndarray:
    labels, rows, cols, offsets = table
    n = int(sum(int(r) * int(c) for r, c in zip(rows, cols)))
    out = np.zeros(n)
    for k, v in _views(labels, rows, cols, offsets, out).items():
        v[...] = t.blocks[k]
    return out


def env_from_packed(side: str, V: Space, M: Legs, table, packed, identity_levels=()) -> EnvTensor:
    t = EnvTensor(side, V, M, identity_levels=identity_levels)
    check_table(t.keys, [t.shape(k) for k in t.keys], *table)
    for k, v in _views(*table, packed).items():
        t.blocks[k] = np.array(v)
    return t


def bond_from_packed(V: Space, table, packed) -> BondTensor:
    t = BondTensor(V)
    for (c, _, _), v in _views(*table, packed).items():
        t.blocks[c] = np.array(v)
    return t


def mpo_from_entries(Ml: Legs, P: Legs, Mr: Legs, entries: dict) -> MPOTensor:
    return MPOTensor(Ml, P, Mr, {(a, sp, s, b, tuple(c)): float(w) for (a, sp, s, b, c), w in entries.items()})
