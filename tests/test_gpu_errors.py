"""Error behaviour of the C ABI on the new entry points (status codes, no exceptions across the
boundary, messages through htn_last_error_string) and small / degenerate inputs."""
import numpy as np
import pytest

from hubbardtn_b200 import _lib, device as dev, hubbardfunctions as hf, sectors as PS

pytestmark = pytest.mark.gpu


def _small(ctx, sym=PS.SU2U1):
    phys = PS.physical_space(sym, 1, 1)
    P = dev.Legs(ctx, sym, phys)
    Va = dev.Space(ctx, sym, {(0, 0, 0): 2, (1, 1, 1): 1, (1, 1, -1): 1})
    Vb = dev.Space(ctx, sym, {(1, 1, 0): 2, (0, 0, 1): 1, (0, 0, -1): 1})
    return P, Va, Vb


def test_shape_and_kind_errors(ctx):
    P, Va, Vb = _small(ctx)
    A = dev.Tensor.mps(ctx, Va, P, Vb)
    Cb = dev.Tensor.bond(ctx, Vb)
    Ca = dev.Tensor.bond(ctx, Va)
    M = dev.Legs(ctx, PS.SU2U1, [(0, 0, 0), (0, 0, 0)])
    GLa = dev.Tensor.env(ctx, 0, Va, M, identity_level=0)
    GRb = dev.Tensor.env(ctx, 1, Vb, M, identity_level=1)
    GRa = dev.Tensor.env(ctx, 1, Va, M, identity_level=1)
    # H_C needs both environments on the bond of C
    with pytest.raises(_lib.HtnError) as ei:
        dev.HeffC(ctx, GLa, GRb, Ca)
    assert ei.value.code == _lib.HTN_ERR_SHAPE
    dev.HeffC(ctx, GLa, GRa, Ca)                       # consistent: fine
    # wrong tensor kinds
    with pytest.raises(_lib.HtnError) as ei:
        dev.HeffC(ctx, GRa, GLa, Ca)
    assert ei.value.code == _lib.HTN_ERR_INVALID
    with pytest.raises(_lib.HtnError):
        dev.tsvd(A)                                    # not a two-site tensor
    with pytest.raises(_lib.HtnError):
        dev.qrpos(A, A.like(), Ca)                     # R lives on the wrong space
    with pytest.raises(_lib.HtnError):
        dev.expval_diag(A, [0.0, 1.0])                 # one value per physical multiplet
    with pytest.raises(_lib.HtnError):
        dev.regauge(A, Ca, A.like())                   # C not on the right bond of AC
    # an eigensolve needs an effective-Hamiltonian plan and matching vectors
    plan = dev.HeffC(ctx, GLa, GRa, Ca)
    with pytest.raises(_lib.HtnError) as ei:
        plan.eigsolve(Cb, Cb.like())
    assert ei.value.code == _lib.HTN_ERR_SHAPE
    assert "structure" in _lib.last_error(ctx.h)


def test_rank_deficient_panel_is_reported(ctx):
    """QRpos of a panel with more columns than rows cannot be an isometry: HTN_ERR_SHAPE, not garbage."""
    sym = PS.SU2U1
    P = dev.Legs(ctx, sym, PS.physical_space(sym, 1, 1))
    Vl = dev.Space(ctx, sym, {(0, 0, 0): 1})
    Vr = dev.Space(ctx, sym, {(1, 1, 0): 5})
    A = dev.Tensor.mps(ctx, Vl, P, Vr)
    A.upload(np.ones(A.nelem))
    with pytest.raises(_lib.HtnError) as ei:
        dev.qrpos(A, A.like(), dev.Tensor.bond(ctx, Vr))
    assert ei.value.code == _lib.HTN_ERR_SHAPE


def test_idmrg2_needs_two_sites_and_rejects_unmirrored_models(ctx):
    with pytest.raises(NotImplementedError):
        hf.compute_groundstate(hf.OB_Sim([1.0], [4.0], 0.0, [0.0], 2, 1, 2.0), ctx=ctx)   # one-site unit cell
    with pytest.raises(NotImplementedError):
        hf.hamiltonian(hf.OB_Sim([1.0], [4.0], 0.0, [0.0], 1, 1, 2.0, kwargs={"U13": [0.3]}), ctx)   # U_ijjj term
    assert hf.hamiltonian(hf.OB_Sim([1.0], [4.0], 0.0, [0.5], 1, 1, 2.0), ctx).chi == 8       # exchange: mirrored (triplet + pair levels)
    H = hf.hamiltonian(hf.OB_Sim([1.0], [4.0], 0.0, [0.0], 1, 1, 2.0), ctx)
    psi = hf.initialize_mps(H, 1, 50, False, ctx)
    rc = _lib.lib.htn_idmrg2(ctx.h, 1, None, None, None, None, None, 1e-2, 1e-6, 1, 30, 1e-8, 0, None, None, None, 0)
    assert rc == _lib.HTN_ERR_INVALID
    rc = _lib.lib.htn_gradient_grassmann(ctx.h, 2, None, None, None, None, None, None, None, 1e-8, 10, 30, None, None, None, None, 0)
    assert rc == _lib.HTN_ERR_INVALID
    assert len(psi) == 2


def test_minimal_bond_dimension_state(ctx):
    """D = 1 per sector (the smallest non-trivial uniform MPS): gauge fixing, environments and VUMPS run and
    give a finite variational energy above the exact one."""
    model = hf.OB_Sim([1.0], [8.0], 0.0, [0.0], 1, 1, 2.0, 1)          # bond_dim cap 1 per sector
    H = hf.hamiltonian(model, ctx)
    psi = hf.initialize_mps(H, 1, 1, False, ctx)
    GL, GR = hf._make_envs(ctx, psi, H)
    info = dev.vumps(ctx, psi.AL, psi.AR, psi.C, psi.AC, H.W, GL, GR, tol=1e-8, maxiter=100)
    assert np.isfinite(info["energy_per_site"]) and info["energy_per_site"] > -0.3275305343795398
    n = hf.density_state(psi)
    assert abs(sum(n) / 2 - 1.0) < 1e-8
