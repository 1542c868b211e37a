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


# ---- oracle <-> device conversion for the uniform-MPS tests ----------------------------------
def pack_blocks(dev_tensor, blocks, key=lambda lab: lab):
    """Packed host array of `dev_tensor`'s layout from an oracle block dict."""
    out = np.zeros(dev_tensor.nelem)
    for lab, view in dev_tensor.block_views(out).items():
        view[...] = blocks[key(lab)]
    return out


def unpack_blocks(dev_tensor, key=lambda lab: lab):
    return {key(lab): np.array(v) for lab, v in dev_tensor.block_views(dev_tensor.download()).items()}


class DevUniform:
    """Device mirror of an oracle uniform MPS (oracle/mps.py state dict) + MPO + environments."""

    def __init__(self, ctx, kind, state, W_list=None):
        from hubbardtn_b200 import device as dev
        self.ctx, self.kind = ctx, kind
        L = self.L = len(state["AL"])
        self.P = [dev.Legs(ctx, kind, state["AL"][i].P.sectors) for i in range(L)]
        self.V = [dev.Space(ctx, kind, state["AL"][i].Vr.as_dict()) for i in range(L)]   # right bond of site i
        for i in range(L):
            assert self.V[i].sectors == state["AL"][i].Vr.sectors, "canonical sector order differs"
        self.AL, self.AR, self.AC, self.C = [], [], [], []
        for i in range(L):
            for name, lst in (("AL", self.AL), ("AR", self.AR), ("AC", self.AC)):
                t = dev.Tensor.mps(ctx, self.V[i - 1], self.P[i], self.V[i])
                t.upload(pack_blocks(t, state[name][i].blocks))
                lst.append(t)
            c = dev.Tensor.bond(ctx, self.V[i])
            c.upload(pack_blocks(c, state["C"][i].blocks, key=lambda lab: lab[0]))
            self.C.append(c)
        self.W = self.GL = self.GR = None
        if W_list is not None:
            self.M = dev.Legs(ctx, kind, W_list[0].Ml.sectors)
            chi = len(W_list[0].Ml)
            self.W = [dev.Mpo(ctx, self.M, self.P[i], self.M, W_list[i].entries) for i in range(L)]
            self.GL = [dev.Tensor.env(ctx, 0, self.V[i - 1], self.M, identity_level=0) for i in range(L)]
            self.GR = [dev.Tensor.env(ctx, 1, self.V[i], self.M, identity_level=chi - 1) for i in range(L)]

    def mps_blocks(self, t):
        return unpack_blocks(t)

    def bond_blocks(self, t):
        return unpack_blocks(t, key=lambda lab: lab[0])


def max_block_err(a: dict, b: dict):
    scale = max(max(np.abs(v).max() for v in b.values() if v.size), 1e-300)
    return max(np.abs(a[k] - b[k]).max() for k in b if b[k].size) / scale
