"""ctypes binding of libhtn.so (the C ABI declared in include/htn.h).

This is the stand-in for the Julia `ccall` shim of INTEGRATION.md (Julia is not in the
image).  There is NO fallback: if the shared library is missing or cannot be loaded the
import fails loudly, and without a CUDA device `htn_ctx_create` returns HTN_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhtn.so")

HTN_OK = 0
HTN_NOT_CONVERGED = 1
HTN_ERR_INVALID = -1
HTN_ERR_NO_DEVICE = -2
HTN_ERR_CUDA = -3
HTN_ERR_OOM = -4
HTN_ERR_SHAPE = -5

SYM_SU2U1 = 0
SYM_U1U1 = 1
SIDE_LEFT = 0
SIDE_RIGHT = 1
T_MPS, T_BOND, T_ENVL, T_ENVR, T_MPST, T_MPS2 = 0, 1, 2, 3, 4, 5


class HtnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libhtn error %d: %s" % (code, msg))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "hubbardtn_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C hubbardtn_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    return C.CDLL(LIB_PATH)


lib = _load()

_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)
_pd = C.POINTER(C.c_double)
_pf = C.POINTER(C.c_float)
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/htn.h declares is listed here and
# tests/test_abi.py checks the two against each other.
SIGNATURES = {
    "htn_ctx_create": (_i32, [_i32, _pp]),
    "htn_ctx_destroy": (_i32, [_p]),
    "htn_last_error_string": (C.c_char_p, [_p]),
    "htn_version": (_i32, []),
    "htn_ctx_synchronize": (_i32, [_p]),
    "htn_ctx_stream": (_i32, [_p, _pp]),
    "htn_space_create": (_i32, [_p, _i32, _i32, _pi32, _pi32, _pp]),
    "htn_space_destroy": (_i32, [_p]),
    "htn_space_info": (_i32, [_p, _pi32, _pi32, _pi32]),
    "htn_legs_create": (_i32, [_p, _i32, _i32, _pi32, _pp]),
    "htn_legs_destroy": (_i32, [_p]),
    "htn_tensor_create_mps": (_i32, [_p, _p, _p, _p, _pp]),
    "htn_tensor_create_bond": (_i32, [_p, _p, _pp]),
    "htn_tensor_create_env": (_i32, [_p, _i32, _p, _p, _i32, _pp]),
    "htn_tensor_create_like": (_i32, [_p, _pp]),
    "htn_tensor_create_transposed": (_i32, [_p, _pp]),
    "htn_tensor_create_mps2": (_i32, [_p, _p, _p, _p, _p, _pp]),
    "htn_tensor_mid_sectors": (_i32, [_p, _pi32, _pi32]),
    "htn_tensor_blocktable5": (_i32, [_p, _pi32, _pi64, _pi32, _pi32, _pi32, _pi64]),
    "htn_tensor_kind": (_i32, [_p]),
    "htn_tensor_space": (_i32, [_p, _i32, _pp]),
    "htn_tensor_device_ptr": (_i32, [_p, _pp, _pi64]),
    "htn_plan_heff_ac2": (_i32, [_p, _p, _p, _p, _p, _p, _pp]),
    "htn_contract_two_site": (_i32, [_p, _p, _p]),
    "htn_tsvd": (_i32, [_p, C.c_double, _i32, _pp, _pp, _pp, _pp, _pd, _pi32]),
    "htn_tensor_transpose": (_i32, [_p, _p, _i32]),
    "htn_tensor_destroy": (_i32, [_p]),
    "htn_tensor_blocktable": (_i32, [_p, _pi32, _pi64, _pi32, _pi32, _pi32, _pi64]),
    "htn_tensor_upload": (_i32, [_p, _p, _i64]),
    "htn_tensor_download": (_i32, [_p, _p, _i64]),
    "htn_mpo_create": (_i32, [_p, _p, _p, _p, _i32, _pi32, _pi32, _pd, _pp]),
    "htn_mpo_create_dense": (_i32, [_p, _p, _p, _p, _pd, C.c_double, _pp]),
    "htn_mpo_entries": (_i32, [_p, _pi32, _pi32, _pi32, _pd]),
    "htn_mpo_destroy": (_i32, [_p]),
    "htn_plan_heff_ac": (_i32, [_p, _p, _p, _p, _p, _pp]),
    "htn_plan_heff_ac_sharded": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _pp]),
    "htn_plan_heff_c": (_i32, [_p, _p, _p, _p, _pp]),
    "htn_plan_destroy": (_i32, [_p]),
    "htn_plan_transfer": (_i32, [_p, _i32, _p, _p, _p, _p, _p, _pp]),
    "htn_transfer_apply": (_i32, [_p, _p, _p, _p, _p]),
    "htn_eigsolve": (_i32, [_p, _p, _p, _i32, C.c_double, _i32, _pd, _pd, _pi32]),
    "htn_qrpos": (_i32, [_p, _p, _p]),
    "htn_lqpos": (_i32, [_p, _p, _p]),
    "htn_regauge": (_i32, [_p, _p, _p]),
    "htn_gauge_right": (_i32, [_p, _i32, _pp, _p, _pp, _pp, C.c_double, _i32, _pi32, _pd]),
    "htn_environments": (_i32, [_p, _i32, _pp, _pp, _pp, _pp, _pp, _pp, C.c_double, _i32, _i32, _pd, _pd]),
    "htn_vumps": (_i32, [_p, _i32, _pp, _pp, _pp, _pp, _pp, _pp, _pp, C.c_double, _i32, _i32, _pd, _pd, _pi32,
                         _pd, _i32]),
    "htn_gradient_grassmann": (_i32, [_p, _i32, _pp, _pp, _pp, _pp, _pp, _pp, _pp, C.c_double, _i32, _i32, _pd, _pd,
                                      _pi32, _pd, _i32]),
    "htn_mul_bond": (_i32, [_p, _p, _i32, C.POINTER(C.c_void_p)]),
    "htn_expval_diag": (_i32, [_p, _pd, _i32, _pd]),
    "htn_entanglement_spectrum": (_i32, [_p, _pd, _i64]),
    "htn_idmrg2": (_i32, [_p, _i32, _pp, _pp, _pp, _pp, _pp, C.c_double, C.c_double, _i32, _i32, C.c_double, _i32, _pd,
                          _pi32, _pd, _i32]),
    "htn_mixed_gauge": (_i32, [_p, _i32, _pp, _p, _pp, _pp, _pp, _i32, C.c_double, _i32, _pi32]),
    "htn_changebonds": (_i32, [_p, _i32, _i32, _pp, _pp, _pp, _pp, _pp, C.c_double, _i32, _i32, C.c_double, C.c_double]),
    "htn_gauge_left": (_i32, [_p, _i32, _pp, _p, _pp, _pp, C.c_double, _i32, _pi32, _pd]),
    "htn_heff_apply": (_i32, [_p, _p, _p]),
    "htn_heff_apply_host": (_i32, [_p, _p, _p, _i64]),
    "htn_plan_stats": (_i32, [_p, _pd, _i32]),
    "htn_plan_profile": (_i32, [_p, _p, _p, _i32, _pf]),
    "htn_heff_time": (_i32, [_p, _p, _p, _i32, _pf]),
    "htn_tensor_dot": (_i32, [_p, _p, _pd]),
    "htn_tensor_axpby": (_i32, [C.c_double, _p, C.c_double, _p]),
    "htn_network_coefficient": (_i32, [_i32, _pi32, _pd]),
    "htn_test_hessenberg_dominant": (_i32, [_i32, _pd, _pd, _pd]),
    "htn_probe_krylov": (_i32, [_p, _i32, _i32, _pf, _pd]),
    "htn_probe_fp64_peak": (_i32, [_p, _i32, _pd]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)          # AttributeError here = ABI mismatch: fail loudly
    _f.restype = _res
    _f.argtypes = _args


def last_error(ctx=None) -> str:
    s = lib.htn_last_error_string(ctx)
    return s.decode() if s else ""


def check(rc: int, ctx=None) -> int:
    """Raise on hard errors (<0); pass through 0 and >0 ('not converged but usable')."""
    if rc < 0:
        raise HtnError(rc, last_error(ctx))
    return rc
