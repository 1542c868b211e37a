"""Parity of the device gauge / environment / Krylov / VUMPS path (through the C ABI) against the
oracle (oracle/mps.py, oracle/heff.py) on identical inputs.

Tolerances: single kernels 1e-12 relative on O(1) data; converged ground-state observables
(energy per site, densities, entanglement spectrum) 1e-10 relative -- the north-star bound."""
import numpy as np
import pytest

from hubbardtn_b200 import device as dev, sectors as PS
from oracle import heff as oheff
from oracle import mps as M
from oracle import sectors as S
from oracle.hubbard import OB_Sim, mpo
from oracle.krylov import lanczos_lowest
from oracle.spaces import physical_space, synthetic_bond_space
from oracle.tensors import BondTensor, EnvTensor, Legs, MPOTensor, MPSTensor
from util import DevUniform, max_block_err, pack_blocks, unpack_blocks

pytestmark = pytest.mark.gpu

LEVELS = {
    S.SU2U1: [(0, 0, 0), (1, 1, 1), (1, 1, -1), (0, 2, 0), (0, 0, 2), (0, 2, -2), (0, 0, 0)],
    S.U1U1: [(0, 0, 0), (1, 1, 1), (1, -1, -1), (0, 2, 0), (0, 0, 2), (0, -2, -2), (0, 0, 0)],
}


def _hubbard_state(kind, u=4.0, D=14, seed=1, t=(1.0,)):
    sim = OB_Sim(t=list(t), u=[u], kwargs={"spin": kind == S.U1U1})
    Ws, P, _ = mpo(sim)
    spaces = M.trim_spaces(kind, [synthetic_bond_space(kind, D, 1), synthetic_bond_space(kind, D, 0)], [P, P])
    st = M.random_state(kind, spaces, [P, P], np.random.default_rng(seed))
    return Ws, P, spaces, st


@pytest.mark.parametrize("kind", [S.SU2U1, S.U1U1])
def test_heff_c_matches_oracle(ctx, kind):
    rng = np.random.default_rng(5)
    V = synthetic_bond_space(kind, 40, 1)
    Mleg = Legs(kind, LEVELS[kind])
    GL = EnvTensor("L", V, Mleg, identity_levels=[0]).randomize(rng)
    GR = EnvTensor("R", V, Mleg, identity_levels=[6]).randomize(rng)
    Cb = BondTensor(V)
    for c in Cb.blocks:
        Cb.blocks[c] = rng.standard_normal(Cb.blocks[c].shape)
    ref = oheff.heff_c_apply(GL, GR, Cb)
    dV = dev.Space(ctx, kind, V.as_dict())
    dM = dev.Legs(ctx, kind, LEVELS[kind])
    dGL = dev.Tensor.env(ctx, 0, dV, dM, identity_level=0)
    dGR = dev.Tensor.env(ctx, 1, dV, dM, identity_level=6)
    dGL.upload(pack_blocks(dGL, GL.blocks))
    dGR.upload(pack_blocks(dGR, GR.blocks))
    x = dev.Tensor.bond(ctx, dV)
    x.upload(pack_blocks(x, Cb.blocks, key=lambda lab: lab[0]))
    y = x.like()
    plan = dev.HeffC(ctx, dGL, dGR, x)
    plan.apply(x, y)
    got = unpack_blocks(y, key=lambda lab: lab[0])
    assert max_block_err(got, ref.blocks) < 1e-12


@pytest.mark.parametrize("kind", [S.SU2U1, S.U1U1])
def test_transfers_match_oracle(ctx, kind):
    rng = np.random.default_rng(3)
    P = physical_space(kind, 1, 1)
    Va, Vb = synthetic_bond_space(kind, 36, 0), synthetic_bond_space(kind, 36, 1)
    Mleg = Legs(kind, LEVELS[kind])
    W = MPOTensor(Mleg, P, Mleg).randomize(rng)
    A = MPSTensor(Va, P, Vb).randomize(rng)
    GL = EnvTensor("L", Va, Mleg, identity_levels=[0]).randomize(rng)
    GR = EnvTensor("R", Vb, Mleg, identity_levels=[6]).randomize(rng)
    refL = oheff.transfer_left(GL, W, A)
    refR = oheff.transfer_right(GR, W, A)
    dVa, dVb = dev.Space(ctx, kind, Va.as_dict()), dev.Space(ctx, kind, Vb.as_dict())
    dP, dM = dev.Legs(ctx, kind, P.sectors), dev.Legs(ctx, kind, LEVELS[kind])
    dW = dev.Mpo(ctx, dM, dP, dM, W.entries)
    dA = dev.Tensor.mps(ctx, dVa, dP, dVb)
    dA.upload(pack_blocks(dA, A.blocks))
    dAt = dA.transposed()
    dA.transpose_into(dAt)
    gl_in = dev.Tensor.env(ctx, 0, dVa, dM, identity_level=0)
    gl_out = dev.Tensor.env(ctx, 0, dVb, dM, identity_level=0)
    gr_in = dev.Tensor.env(ctx, 1, dVb, dM, identity_level=6)
    gr_out = dev.Tensor.env(ctx, 1, dVa, dM, identity_level=6)
    gl_in.upload(pack_blocks(gl_in, GL.blocks))
    gr_in.upload(pack_blocks(gr_in, GR.blocks))
    dev.Transfer(ctx, 0, dW, dA, dAt, gl_in, gl_out).apply(dA, dAt, gl_in, gl_out)
    dev.Transfer(ctx, 1, dW, dA, dAt, gr_in, gr_out).apply(dA, dAt, gr_in, gr_out)
    assert max_block_err(unpack_blocks(gl_out), refL.blocks) < 1e-12
    assert max_block_err(unpack_blocks(gr_out), refR.blocks) < 1e-12
    # MPO-free transfer matrix on bond tensors (gauge / GMRES operator)
    X = BondTensor(Va)
    for c in X.blocks:
        X.blocks[c] = rng.standard_normal(X.blocks[c].shape)
    idm = M.identity_mpo(P)
    Xe = EnvTensor("L", Va, idm.Ml, {(0, c, c): b for c, b in X.blocks.items()})
    ref1 = M.TransferPlan("L", idm, Va, P, Vb).apply(Xe, A)
    dId = dev.Mpo(ctx, dev.Legs(ctx, kind, [(0, 0, 0)]), dP, dev.Legs(ctx, kind, [(0, 0, 0)]), idm.entries)
    bx, by = dev.Tensor.bond(ctx, dVa), dev.Tensor.bond(ctx, dVb)
    bx.upload(pack_blocks(bx, X.blocks, key=lambda lab: lab[0]))
    dev.Transfer(ctx, 0, dId, dA, dAt, bx, by).apply(dA, dAt, bx, by)
    got = unpack_blocks(by, key=lambda lab: lab[0])
    assert max_block_err(got, {k[1]: v for k, v in ref1.blocks.items()}) < 1e-12


@pytest.mark.parametrize("kind", [S.SU2U1, S.U1U1])
def test_qr_lq_regauge_match_oracle(ctx, kind):
    rng = np.random.default_rng(2)
    P = physical_space(kind)
    spaces = M.trim_spaces(kind, [synthetic_bond_space(kind, 60, 1), synthetic_bond_space(kind, 60, 0)], [P, P])
    A = MPSTensor(spaces[1], P, spaces[0]).randomize(rng)
    dVl, dVr = dev.Space(ctx, kind, spaces[1].as_dict()), dev.Space(ctx, kind, spaces[0].as_dict())
    dP = dev.Legs(ctx, kind, P.sectors)
    dA = dev.Tensor.mps(ctx, dVl, dP, dVr)
    dA.upload(pack_blocks(dA, A.blocks))
    Q, R = dA.like(), dev.Tensor.bond(ctx, dVr)
    dev.qrpos(dA, Q, R)
    Qo, Ro = M.left_orth(A)
    assert max_block_err(unpack_blocks(Q), Qo.blocks) < 1e-11
    assert max_block_err(unpack_blocks(R, key=lambda lab: lab[0]), Ro.blocks) < 1e-11
    Lm, Q2 = dev.Tensor.bond(ctx, dVl), dA.like()
    dev.lqpos(dA, Lm, Q2)
    Lo, Q2o = M.right_orth(A)
    assert max_block_err(unpack_blocks(Q2), Q2o.blocks) < 1e-11
    assert max_block_err(unpack_blocks(Lm, key=lambda lab: lab[0]), Lo.blocks) < 1e-11
    Cb = BondTensor(spaces[0])
    for c in Cb.blocks:
        Cb.blocks[c] = rng.standard_normal(Cb.blocks[c].shape)
    dC = dev.Tensor.bond(ctx, dVr)
    dC.upload(pack_blocks(dC, Cb.blocks, key=lambda lab: lab[0]))
    AL = dA.like()
    dev.regauge(dA, dC, AL)
    assert max_block_err(unpack_blocks(AL), M.regauge(A, Cb).blocks) < 1e-11


@pytest.mark.parametrize("kind", [S.SU2U1, S.U1U1])
def test_gauge_right_matches_oracle(ctx, kind):
    Ws, P, spaces, st = _hubbard_state(kind, D=20)
    rng = np.random.default_rng(9)
    C0 = BondTensor(spaces[1])
    for c in C0.blocks:
        n = C0.blocks[c].shape[0]
        C0.blocks[c] = np.eye(n) + 0.2 * rng.standard_normal((n, n))
    ARo, Co, info = M.uniform_rightorth(st["AL"], C0, tol=1e-13)
    du = DevUniform(ctx, kind, st)
    g = dev.Tensor.bond(ctx, du.V[1])
    g.upload(pack_blocks(g, C0.blocks, key=lambda lab: lab[0]))
    res = dev.gauge_right(ctx, du.AL, g, du.AR, du.C, tol=1e-13)
    assert res["converged"] and abs(res["iterations"] - info["iterations"]) <= 2
    for i in range(2):
        assert max_block_err(du.mps_blocks(du.AR[i]), ARo[i].blocks) < 1e-9
        assert max_block_err(du.bond_blocks(du.C[i]), Co[i].blocks) < 1e-9


@pytest.mark.parametrize("kind", [S.SU2U1, S.U1U1])
def test_environments_and_eigsolve_match_oracle(ctx, kind):
    Ws, P, spaces, st = _hubbard_state(kind, u=3.0, D=18, t=(1.0, 0.25))
    envs = M.Environments(st, Ws, tol=1e-13)
    du = DevUniform(ctx, kind, st, Ws)
    res = dev.environments(ctx, du.AL, du.AR, du.C, du.W, du.GL, du.GR, tol=1e-13)
    assert res["converged"]
    assert abs(res["energy_cell_left"] - envs.energy_cell_left) < 1e-11 * max(1.0, abs(envs.energy_cell_left))
    assert abs(res["energy_cell_right"] - envs.energy_cell_right) < 1e-11 * max(1.0, abs(envs.energy_cell_right))
    for i in range(2):
        assert max_block_err(unpack_blocks(du.GL[i]), envs.GL[i].blocks) < 1e-10
        assert max_block_err(unpack_blocks(du.GR[i]), envs.GR[i].blocks) < 1e-10
    # Lanczos on H_AC and H_C with these (Hermitian) environments
    hac = dev.HeffAC(ctx, du.GL[0], du.W[0], du.GR[0], du.AC[0])
    x = du.AC[0].like()
    ev, info = hac.eigsolve(du.AC[0], x, krylovdim=30, tol=1e-11, maxiter=20)
    oplan = oheff.HeffACPlan(envs.GL[0], Ws[0], envs.GR[0], st["AC"][0])
    ev_o, x_o, info_o = lanczos_lowest(oplan.apply, st["AC"][0], tol=1e-11, krylovdim=30, maxiter=20)
    assert info["converged"] and info_o["converged"]
    assert abs(ev - ev_o) < 1e-10 * max(1.0, abs(ev_o))
    if M.vdot(x_o, st["AC"][0]) < 0:
        M.vscale(x_o, -1.0)
    assert max_block_err(du.mps_blocks(x), x_o.blocks) < 1e-7     # eigenvector error ~ sqrt(residual)
    hc = dev.HeffC(ctx, du.GL[1], du.GR[0], du.C[0])
    c = du.C[0].like()
    evc, infoc = hc.eigsolve(du.C[0], c, krylovdim=30, tol=1e-11, maxiter=20)
    evc_o, _, _ = lanczos_lowest(lambda v: oheff.heff_c_apply(envs.GL[1], envs.GR[0], v), st["C"][0], tol=1e-11,
                                 krylovdim=30, maxiter=20)
    assert abs(evc - evc_o) < 1e-10 * max(1.0, abs(evc_o))


@pytest.mark.parametrize("kind,u,D", [(S.SU2U1, 8.0, 10), (S.U1U1, 8.0, 12)])
def test_vumps_ground_state_matches_oracle(ctx, kind, u, D):
    """End-to-end: same initial state, same MPO, same tolerances -> energy per site, densities and
    entanglement spectrum within 1e-10 relative of the oracle (north-star parity bound)."""
    Ws, P, spaces, st0 = _hubbard_state(kind, u=u, D=D)
    st, envs, eps, log = M.vumps(st0, Ws, tol=1e-10, maxiter=300)
    assert eps < 1e-10
    du = DevUniform(ctx, kind, st0, Ws)
    res = dev.vumps(ctx, du.AL, du.AR, du.C, du.AC, du.W, du.GL, du.GR, tol=1e-10, maxiter=300)
    assert res["converged"], res
    E, Eo = res["energy_per_site"], envs.energy_per_site
    assert abs(E - Eo) < 1e-10 * abs(Eo), (E, Eo)
    vals = [0, 2, 1] if kind == S.SU2U1 else [0, 2, 1, 1]
    for i in range(2):
        n_dev = dev.expval_diag(du.AC[i], vals)
        n_or = M.expval_diag(st["AC"][i], vals)
        assert abs(n_dev - n_or) < 1e-9 * abs(n_or)
        spec_o = M.entanglement_spectrum(st["C"][i])
        spec_d = dev.entanglement_spectrum(du.C[i])                  # device Jacobi SVD of C
        cb = du.bond_blocks(du.C[i])
        for c, blk in cb.items():
            ref = spec_o[du.V[i].sectors[c]]
            assert np.abs(spec_d[c] - ref).max() < 1e-9 * max(ref.max(), 1e-300) + 1e-12
            assert np.abs(np.linalg.svd(blk, compute_uv=False) - spec_d[c]).max() < 1e-12
    # same algorithm, but the inexact inner solves stop at slightly different points (the device checks
    # the Lanczos residual every 5 steps, the oracle every step): iteration counts are close, not equal
    assert abs(res["iterations"] - len(log)) <= 0.5 * len(log) + 2


def test_qrpos_is_orthogonal_on_ill_conditioned_panels(ctx):
    """Gauge fixing meets AL.C with cond(C) ~ 1/smallest Schmidt value (1e8 and beyond): the blocked QR kernel must
    keep Q^T Q = 1 at machine precision there (one block-Gram-Schmidt round would lose eps * cond)."""
    kind = S.SU2U1
    rng = np.random.default_rng(11)
    P = physical_space(kind)
    spaces = M.trim_spaces(kind, [synthetic_bond_space(kind, 150, 1), synthetic_bond_space(kind, 150, 0)], [P, P])
    A = MPSTensor(spaces[1], P, spaces[0]).randomize(rng)
    for k, v in A.blocks.items():                       # graded columns: singular values from 1 down to 1e-9
        n = v.shape[1]
        A.blocks[k] = v * np.logspace(0, -9, n)[None, :]
    dVl, dVr = dev.Space(ctx, kind, spaces[1].as_dict()), dev.Space(ctx, kind, spaces[0].as_dict())
    dA = dev.Tensor.mps(ctx, dVl, dev.Legs(ctx, kind, P.sectors), dVr)
    dA.upload(pack_blocks(dA, A.blocks))
    Q, R = dA.like(), dev.Tensor.bond(ctx, dVr)
    dev.qrpos(dA, Q, R)
    Qb = unpack_blocks(Q)
    Rb = unpack_blocks(R, key=lambda lab: lab[0])
    by_r = {}
    for (l, s, r), blk in Qb.items():
        by_r.setdefault(r, []).append(((l, s, r), blk))
    for r, lst in by_r.items():
        Qr = np.vstack([b for _, b in sorted(lst, key=lambda t: (t[0][1], t[0][0]))])
        Ar = np.vstack([A.blocks[k] for k, _ in sorted(lst, key=lambda t: (t[0][1], t[0][0]))])
        n = Qr.shape[1]
        assert np.abs(Qr.T @ Qr - np.eye(n)).max() < 5e-13
        assert np.abs(Qr @ Rb[r] - Ar).max() < 1e-13 * np.abs(Ar).max() * 10
        assert (np.diag(Rb[r]) > 0).all()


@pytest.mark.parametrize("kind,u,D", [(S.SU2U1, 8.0, 10), (S.U1U1, 4.0, 12)])
def test_gradient_grassmann_reaches_the_vumps_fixed_point(ctx, kind, u, D):
    """HF:1025-1027 `VUMPS(maxiter) & GradientGrassmann(maxiter, tol)`: VUMPS stopped early, the Riemannian
    CG polish takes the state to the same variational minimum the oracle's converged VUMPS finds on these
    bond spaces: energy per site within 1e-10 relative, densities within 1e-7, gradient norm below tol,
    energy non-increasing along the accepted steps."""
    Ws, P, spaces, st0 = _hubbard_state(kind, u=u, D=D)
    st, envs, eps, log = M.vumps(st0, Ws, tol=1e-10, maxiter=300)
    assert eps < 1e-10
    du = DevUniform(ctx, kind, st0, Ws)
    res = dev.vumps(ctx, du.AL, du.AR, du.C, du.AC, du.W, du.GL, du.GR, tol=1e-10, maxiter=3)
    assert not res["converged"] and res["delta"] > 1e-6
    gg = dev.gradient_grassmann(ctx, du.AL, du.AR, du.C, du.AC, du.W, du.GL, du.GR, tol=1e-8, maxiter=1500)
    assert gg["converged"] and gg["delta"] < 1e-8, (gg["delta"], gg["iterations"])
    E, Eo = gg["energy_per_site"], envs.energy_per_site
    assert E < res["energy_per_site"] + 1e-12
    assert abs(E - Eo) < 1e-10 * abs(Eo), (E, Eo)
    assert np.all(np.diff(gg["log"][:, 1]) < 1e-11)
    vals = [0, 2, 1] if kind == S.SU2U1 else [0, 2, 1, 1]
    for i in range(2):
        assert abs(dev.expval_diag(du.AC[i], vals) - M.expval_diag(st["AC"][i], vals)) < 1e-7
