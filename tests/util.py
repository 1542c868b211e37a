"""Shared helpers of the parity tests: build the oracle's view of a device-side case."""
import numpy as np

from oracle import bridge
from oracle import sectors as OS
from oracle.tensors import Legs, Space


def table(t):
    return (t.labels, t.rows, t.cols, t.offsets)


def oracle_view(case):
    """Oracle tensors holding exactly the data uploaded for a hubbardtn_b200.synthetic.HeffCase."""
    kind = case.sym
    Vl, Vr = Space(kind, case.vl_mult), Space(kind, case.vr_mult)
    assert Vl.sectors == case.Vl.sectors and Vl.mult == case.Vl.mult, "canonical sector order differs"
    assert Vr.sectors == case.Vr.sectors and Vr.mult == case.Vr.mult, "canonical sector order differs"
    P, M = Legs(kind, case.phys), Legs(kind, case.levels)
    GL = bridge.env_from_packed("L", Vl, M, table(case.GL), case.gl_host, identity_levels=[0])
    GR = bridge.env_from_packed("R", Vr, M, table(case.GR), case.gr_host, identity_levels=[len(case.levels) - 1])
    W = bridge.mpo_from_entries(M, P, M, case.w_entries)
    x = bridge.mps_from_packed(Vl, P, Vr, table(case.x), case.x_host)
    return dict(Vl=Vl, Vr=Vr, P=P, M=M, GL=GL, GR=GR, W=W, x=x)


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
