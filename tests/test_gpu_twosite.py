"""Parity of the device two-site path (H_AC2 apply, two-site contraction, truncated SVD) against
oracle/twosite.py on identical inputs, through the C ABI."""
import numpy as np
import pytest

from hubbardtn_b200 import device as dev
from oracle import sectors as S
from oracle import twosite as T2
from oracle.mps import mul_right
from oracle.spaces import physical_space, synthetic_bond_space
from oracle.tensors import EnvTensor, Legs, MPOTensor, MPSTensor
from util import max_block_err, pack_blocks, unpack_blocks

pytestmark = pytest.mark.gpu

LEVELS = {
    S.SU2U1: [(0, 0, 0), (1, 1, 1), (1, 1, -1), (0, 2, 0), (0, 0, 0)],
    S.U1U1: [(0, 0, 0), (1, 1, 1), (1, -1, -1), (0, 2, 0), (0, 0, 0)],
}


def _key5(t):
    return lambda lab: (lab[0], lab[1], t.mid[lab[2]], lab[3], lab[4])


def _setup(ctx, kind, D, seed=4):
    rng = np.random.default_rng(seed)
    P = physical_space(kind, 1, 1)
    Va, Vb = synthetic_bond_space(kind, D, 0), synthetic_bond_space(kind, D, 0)
    Mleg = Legs(kind, LEVELS[kind])
    dVa, dVb = dev.Space(ctx, kind, Va.as_dict()), dev.Space(ctx, kind, Vb.as_dict())
    dP, dM = dev.Legs(ctx, kind, P.sectors), dev.Legs(ctx, kind, LEVELS[kind])
    return rng, P, Va, Vb, Mleg, dVa, dVb, dP, dM


@pytest.mark.parametrize("kind,D", [(S.SU2U1, 12), (S.U1U1, 12), (S.SU2U1, 70)])
def test_heff_ac2_matches_oracle(ctx, kind, D):
    rng, P, Va, Vb, Mleg, dVa, dVb, dP, dM = _setup(ctx, kind, D)
    GL = EnvTensor("L", Va, Mleg, identity_levels=[0]).randomize(rng)
    GR = EnvTensor("R", Vb, Mleg, identity_levels=[4]).randomize(rng)
    W1 = MPOTensor(Mleg, P, Mleg).randomize(rng)
    W2 = MPOTensor(Mleg, P, Mleg).randomize(rng)
    x = T2.TwoSiteTensor(Va, P, P, Vb).randomize(rng)
    ref = T2.HeffAC2Plan(GL, W1, W2, GR, x).apply(x)
    dGL = dev.Tensor.env(ctx, 0, dVa, dM, identity_level=0)
    dGR = dev.Tensor.env(ctx, 1, dVb, dM, identity_level=4)
    dGL.upload(pack_blocks(dGL, GL.blocks))
    dGR.upload(pack_blocks(dGR, GR.blocks))
    dW1, dW2 = dev.Mpo(ctx, dM, dP, dM, W1.entries), dev.Mpo(ctx, dM, dP, dM, W2.entries)
    dx = dev.Tensor.mps2(ctx, dVa, dP, dP, dVb)
    # block tables agree (order, labels) with the oracle's canonical enumeration
    assert [_key5(dx)(tuple(int(v) for v in lab)) for lab in dx.labels] == x.keys
    dx.upload(pack_blocks(dx, x.blocks, key=_key5(dx)))
    dy = dx.like()
    plan = dev.HeffAC2(ctx, dGL, dW1, dW2, dGR, dx)
    plan.apply(dx, dy)
    got = unpack_blocks(dy, key=_key5(dy))
    assert max_block_err(got, ref.blocks) < 1e-12


@pytest.mark.parametrize("kind", [S.SU2U1, S.U1U1])
def test_contract_and_tsvd_match_oracle(ctx, kind):
    rng, P, Va, Vb, Mleg, dVa, dVb, dP, dM = _setup(ctx, kind, 24, seed=8)
    Vm = synthetic_bond_space(kind, 20, 1)
    A1 = MPSTensor(Va, P, Vm).randomize(rng)
    A2 = MPSTensor(Vm, P, Vb).randomize(rng)
    x = T2.contract_two_site(A1, A2)
    dVm = dev.Space(ctx, kind, Vm.as_dict())
    dA1, dA2 = dev.Tensor.mps(ctx, dVa, dP, dVm), dev.Tensor.mps(ctx, dVm, dP, dVb)
    dA1.upload(pack_blocks(dA1, A1.blocks))
    dA2.upload(pack_blocks(dA2, A2.blocks))
    dx = dev.Tensor.mps2(ctx, dVa, dP, dP, dVb)
    dev.contract_two_site(dA1, dA2, dx)
    assert max_block_err(unpack_blocks(dx, key=_key5(dx)), x.blocks) < 1e-12
    # generic (full-rank) two-site tensor: SVD without and with truncation
    y = T2.TwoSiteTensor(Va, P, P, Vb).randomize(rng)
    dx.upload(pack_blocks(dx, y.blocks, key=_key5(dx)))
    for cut, maxdim in ((0.0, 0), (0.05, 0), (0.0, 17)):
        ALo, Co, ARo, info = T2.tsvd(y, cut, maxdim if maxdim else None)
        V, AL, Cb, AR, dinfo = dev.tsvd(dx, cut, maxdim, sym=kind)
        assert V.sectors == info["space"].sectors and V.mult == info["space"].mult
        assert dinfo["kept"] == info["kept"]
        assert abs(dinfo["discarded_weight"] - info["discarded_weight"]) < 1e-12
        assert max_block_err(unpack_blocks(Cb, key=lambda lab: lab[0]), Co.blocks) < 1e-12
        assert max_block_err(unpack_blocks(AL), ALo.blocks) < 1e-9
        assert max_block_err(unpack_blocks(AR), ARo.blocks) < 1e-9
    # exact reconstruction when nothing is cut
    ALo, Co, ARo, info = T2.tsvd(y, 0.0)
    V, AL, Cb, AR, dinfo = dev.tsvd(dx, 0.0, 0, sym=kind)
    ALh = MPSTensor(Va, P, info["space"], unpack_blocks(AL))
    ARh = MPSTensor(info["space"], P, Vb, unpack_blocks(AR))
    from oracle.tensors import BondTensor
    Ch = BondTensor(info["space"], unpack_blocks(Cb, key=lambda lab: lab[0]))
    rec = T2.contract_two_site(mul_right(ALh, Ch), ARh)
    assert max_block_err(rec.blocks, y.blocks) < 1e-12


def test_idmrg2_then_vumps_reproduces_reference_golden(ctx):
    """The reference's schedule (HF:1010 IDMRG2 with truncbelow(1e-2), then HF:1025-1027 VUMPS) on the
    device, from the same random initial state as the oracle: bond spaces identical, Schmidt values
    and energy per site equal to the oracle's, energy equal to the reference's hard-coded golden
    (test/OB.jl:44, U=5, P/Q=1: -0.48460447) to its printed digits."""
    import json
    import os
    from oracle import mps as M
    from oracle.hubbard import OB_Sim, mpo
    from oracle.spaces import initial_bond_spaces
    from util import DevUniform
    kind = S.SU2U1
    Ws, P, _ = mpo(OB_Sim(t=[1.0], u=[5.0]))
    sp = M.trim_spaces(kind, initial_bond_spaces(kind, [P, P], 1, 50), [P, P])
    st0 = M.random_state(kind, sp, [P, P], np.random.default_rng(3))   # lands in the reference's truncated space
    ALo, Co, ARo, eps_o, log_o = T2.idmrg2(st0, Ws, cut=1e-2, tol=1e-6, maxiter=60)
    du = DevUniform(ctx, kind, st0, Ws)
    AL, AR, Cs, AC, info = dev.idmrg2(ctx, du.AL, du.AR, du.C, du.AC, du.W, cut=1e-2, tol=1e-6, maxiter=60)
    assert info["converged"] and abs(info["iterations"] - len(log_o)) <= 1
    for i in range(2):
        V = Cs[i].space(0, kind)
        assert V.sectors == Co[i].V.sectors and V.mult == Co[i].V.mult        # same truncated bond spaces
        assert max_block_err(unpack_blocks(Cs[i], key=lambda lab: lab[0]), Co[i].blocks) < 1e-6
    # gauge-fix and polish with VUMPS on the grown spaces, both sides
    sto = T2.idmrg2_to_uniform(ARo, Co)
    sto, envs, eps, _ = M.vumps(sto, Ws, tol=1e-8, maxiter=60)
    AL, AR, Cs, AC = dev.uniform_from_right(ctx, AR, Cs[1], kind)
    V = [Cs[i].space(0, kind) for i in range(2)]
    chi = len(Ws[0].Ml)
    GL = [dev.Tensor.env(ctx, 0, V[i - 1], du.M, identity_level=0) for i in range(2)]
    GR = [dev.Tensor.env(ctx, 1, V[i], du.M, identity_level=chi - 1) for i in range(2)]
    res = dev.vumps(ctx, AL, AR, Cs, AC, du.W, GL, GR, tol=1e-8, maxiter=60)
    assert res["converged"]
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_energies.json")))
    g = [r for r in gold["reference"] if r["u"] == [5.0] and r["P"] == r["Q"]][0]
    assert abs(res["energy_per_site"] - envs.energy_per_site) < 1e-9
    assert abs(res["energy_per_site"] - g["E"]) < 5e-8, (res["energy_per_site"], g["E"])


def test_multiband_idmrg2_vumps_reproduces_reference_golden(ctx):
    """test/MB.jl:24-35,59: two uncoupled bands on a 4-site unit cell (site-dependent MPO from MB_Sim), the
    reference's schedule on the device from the oracle's initial state: same truncated bond spaces as the
    oracle and E/site = -0.630375296, every digit the reference prints (its own atol is 1e-1)."""
    from test_oracle_twosite import mb_golden_setup
    from oracle import mps as M
    from util import DevUniform
    g, kind, Ws, P, st0 = mb_golden_setup()
    ALo, Co, ARo, eps_o, log_o = T2.idmrg2(st0, Ws, cut=1e-2, tol=1e-6, maxiter=30)
    du = DevUniform(ctx, kind, st0, Ws)
    AL, AR, Cs, AC, info = dev.idmrg2(ctx, du.AL, du.AR, du.C, du.AC, du.W, cut=1e-2, tol=1e-6, maxiter=30)
    assert info["converged"] and abs(info["iterations"] - len(log_o)) <= 1
    for i in range(4):
        V = Cs[i].space(0, kind)
        assert V.sectors == Co[i].V.sectors and V.mult == Co[i].V.mult
    AL, AR, Cs, AC = dev.uniform_from_right(ctx, AR, Cs[3], kind)
    V = [Cs[i].space(0, kind) for i in range(4)]
    chi = len(Ws[0].Ml)
    GL = [dev.Tensor.env(ctx, 0, V[i - 1], du.M, identity_level=0) for i in range(4)]
    GR = [dev.Tensor.env(ctx, 1, V[i], du.M, identity_level=chi - 1) for i in range(4)]
    res = dev.vumps(ctx, AL, AR, Cs, AC, du.W, GL, GR, tol=1e-6, maxiter=60)
    assert res["converged"]
    assert abs(res["energy_per_site"] - g["E"]) < 2e-9, (res["energy_per_site"], g["E"])
