"""Host-side object layer over the C ABI (numpy in / numpy out, no arithmetic here).

Mirrors the objects HubbardTN hands to MPSKit at the drop-in boundary
(src/HubbardFunctions.jl:1010-1027): graded spaces (`Vect[I](...)`, HF:248-251), site
tensors of an `InfiniteMPS`, the per-site MPO tensor of an `InfiniteMPOHamiltonian`, the
environments, and the effective-Hamiltonian operator.  All numbers live on the GPU in the
library's arena; this layer only moves packed host arrays across the boundary.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import lib


def _i32arr(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


class Context:
    """One CUDA device + stream + staging buffers (htn_ctx)."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        rc = lib.htn_ctx_create(device, C.byref(h))
        if rc != 0:
            raise L.HtnError(rc, L.last_error(None))
        self.h = h
        self.device = device

    def synchronize(self):
        L.check(lib.htn_ctx_synchronize(self.h), self.h)

    @property
    def stream_ptr(self) -> int:
        """cudaStream_t of the library (for stream-ordered collectives of the caller)."""
        p = C.c_void_p()
        L.check(lib.htn_ctx_stream(self.h, C.byref(p)), self.h)
        return int(p.value or 0)

    def probe_fp64_peak(self, which: int = 0) -> float:
        out = C.c_double()
        L.check(lib.htn_probe_fp64_peak(self.h, which, C.byref(out)), self.h)
        return out.value

    def close(self):
        """Destroy the context.  Handles created from it become inert (their destructors
        check `ctx.h`); the C ABI itself requires children to be destroyed first."""
        if self.h:
            h, self.h = self.h, None
            lib.htn_ctx_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Space:
    """Graded bond space: {sector label (p,q,n): multiplicity}; stored in canonical order."""

    def __init__(self, ctx: Context, sym: int, mults: dict):
        self.ctx, self.sym = ctx, sym
        labels = [tuple(k) for k in mults.keys()]
        lab, plab = _i32arr(np.array(labels, dtype=np.int32).reshape(-1, 3))
        mul, pmul = _i32arr(list(mults.values()))
        h = C.c_void_p()
        L.check(lib.htn_space_create(ctx.h, sym, len(labels), plab, pmul, C.byref(h)), ctx.h)
        self.h = h
        n = C.c_int32()
        L.check(lib.htn_space_info(h, C.byref(n), None, None), ctx.h)
        lab2 = np.zeros((n.value, 3), dtype=np.int32)
        mul2 = np.zeros(n.value, dtype=np.int32)
        L.check(lib.htn_space_info(h, C.byref(n), lab2.ctypes.data_as(C.POINTER(C.c_int32)),
                                   mul2.ctypes.data_as(C.POINTER(C.c_int32))), ctx.h)
        self.sectors = [tuple(int(v) for v in row) for row in lab2]
        self.mult = [int(v) for v in mul2]

    def __del__(self):
        try:
            if self.h and self.ctx.h:      # a closed context has already released the device
                lib.htn_space_destroy(self.h)
            self.h = None
        except Exception:
            pass


class Legs:
    """Ordered single multiplets: the physical space or the MPO virtual levels."""

    def __init__(self, ctx: Context, sym: int, sectors):
        self.ctx, self.sym = ctx, sym
        self.sectors = [tuple(int(v) for v in s) for s in sectors]
        lab, plab = _i32arr(np.array(self.sectors, dtype=np.int32).reshape(-1, 3))
        h = C.c_void_p()
        # ctx may be None: legs are host-only objects (used by the device-free MPO projection tests)
        L.check(lib.htn_legs_create(ctx.h if ctx else None, sym, len(self.sectors), plab, C.byref(h)),
                ctx.h if ctx else None)
        self.h = h

    def __len__(self):
        return len(self.sectors)

    def __del__(self):
        try:
            if self.h and (self.ctx is None or self.ctx.h):
                lib.htn_legs_destroy(self.h)
            self.h = None
        except Exception:
            pass


class Tensor:
    """Block-sparse tensor resident in HBM.  `table` = list of (labels, rows, cols, offset)
    describing the packed host layout used by upload()/download()."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle
        nb, ne = C.c_int32(), C.c_int64()
        L.check(lib.htn_tensor_blocktable(handle, C.byref(nb), C.byref(ne), None, None, None, None), ctx.h)
        self.nblocks, self.nelem = nb.value, ne.value
        lab = np.zeros((self.nblocks, 3), dtype=np.int32)
        rows = np.zeros(self.nblocks, dtype=np.int32)
        cols = np.zeros(self.nblocks, dtype=np.int32)
        offs = np.zeros(self.nblocks, dtype=np.int64)
        pi32 = C.POINTER(C.c_int32)
        L.check(lib.htn_tensor_blocktable(handle, C.byref(nb), C.byref(ne), lab.ctypes.data_as(pi32),
                                          rows.ctypes.data_as(pi32), cols.ctypes.data_as(pi32),
                                          offs.ctypes.data_as(C.POINTER(C.c_int64))), ctx.h)
        self.labels, self.rows, self.cols, self.offsets = lab, rows, cols, offs
        self.kind = lib.htn_tensor_kind(handle)
        self.mid = None
        if self.kind == L.T_MPS2:
            lab5 = np.zeros((self.nblocks, 5), dtype=np.int32)
            L.check(lib.htn_tensor_blocktable5(handle, C.byref(nb), C.byref(ne), lab5.ctypes.data_as(pi32), None, None,
                                               None), ctx.h)
            self.labels = lab5
            nm = C.c_int32()
            L.check(lib.htn_tensor_mid_sectors(handle, C.byref(nm), None), ctx.h)
            mid = np.zeros((nm.value, 3), dtype=np.int32)
            L.check(lib.htn_tensor_mid_sectors(handle, C.byref(nm), mid.ctypes.data_as(pi32)), ctx.h)
            self.mid = [tuple(int(v) for v in row) for row in mid]

    # -- constructors --------------------------------------------------------------------
    @staticmethod
    def mps(ctx, Vl: Space, P: Legs, Vr: Space) -> "Tensor":
        h = C.c_void_p()
        L.check(lib.htn_tensor_create_mps(ctx.h, Vl.h, P.h, Vr.h, C.byref(h)), ctx.h)
        return Tensor(ctx, h)

    @staticmethod
    def mps2(ctx, Vl: Space, P1: Legs, P2: Legs, Vr: Space) -> "Tensor":
        """Two-site tensor (MPSKit AC2) in the fusion-tree basis; labels (l, s1, m, s2, r), m -> self.mid."""
        h = C.c_void_p()
        L.check(lib.htn_tensor_create_mps2(ctx.h, Vl.h, P1.h, P2.h, Vr.h, C.byref(h)), ctx.h)
        return Tensor(ctx, h)

    @staticmethod
    def bond(ctx, V: Space) -> "Tensor":
        h = C.c_void_p()
        L.check(lib.htn_tensor_create_bond(ctx.h, V.h, C.byref(h)), ctx.h)
        return Tensor(ctx, h)

    @staticmethod
    def env(ctx, side: int, V: Space, M: Legs, identity_level: int = -1) -> "Tensor":
        h = C.c_void_p()
        L.check(lib.htn_tensor_create_env(ctx.h, side, V.h, M.h, identity_level, C.byref(h)), ctx.h)
        return Tensor(ctx, h)

    def like(self) -> "Tensor":
        h = C.c_void_p()
        L.check(lib.htn_tensor_create_like(self.h, C.byref(h)), self.ctx.h)
        return Tensor(self.ctx, h)

    def space(self, which: int, sym: int = 0) -> "Space":
        """Copy of the left (0) / right (1) bond space this tensor was built on."""
        h = C.c_void_p()
        L.check(lib.htn_tensor_space(self.h, which, C.byref(h)), self.ctx.h)
        return _OwnedSpace(self.ctx, sym, h)

    def device_array(self):
        """Object exposing the padded device arena through `__cuda_array_interface__` (1-D float64), e.g. for
        `torch.as_tensor(t.device_array(), device="cuda")` -> NCCL allreduce in place."""
        ptr, n = C.c_void_p(), C.c_int64()
        L.check(lib.htn_tensor_device_ptr(self.h, C.byref(ptr), C.byref(n)), self.ctx.h)

        class _Arr:
            __cuda_array_interface__ = {"shape": (n.value,), "typestr": "<f8", "data": (ptr.value, False), "version": 3,
                                        "strides": None}
        return _Arr()

    def transposed(self) -> "Tensor":
        """Blockwise-transposed companion (kind MPST) of an MPS tensor (structure only)."""
        h = C.c_void_p()
        L.check(lib.htn_tensor_create_transposed(self.h, C.byref(h)), self.ctx.h)
        return Tensor(self.ctx, h)

    def transpose_into(self, dst: "Tensor", weighted: bool = False) -> "Tensor":
        L.check(lib.htn_tensor_transpose(self.h, dst.h, 1 if weighted else 0), self.ctx.h)
        return dst

    def like_copy(self) -> "Tensor":
        """New tensor with the same structure AND data (device-to-device)."""
        t = self.like()
        t.axpby(1.0, self, 0.0)
        return t

    # -- data movement -------------------------------------------------------------------
    def upload(self, packed: np.ndarray):
        packed = np.ascontiguousarray(packed, dtype=np.float64)
        L.check(lib.htn_tensor_upload(self.h, packed.ctypes.data, packed.size), self.ctx.h)
        return self

    def upload_ptr(self, ptr: int, nelem: int):
        """Upload from a raw host pointer (e.g. pinned memory)."""
        L.check(lib.htn_tensor_upload(self.h, ptr, nelem), self.ctx.h)
        return self

    def download_ptr(self, ptr: int, nelem: int):
        L.check(lib.htn_tensor_download(self.h, ptr, nelem), self.ctx.h)

    def download(self) -> np.ndarray:
        out = np.empty(self.nelem, dtype=np.float64)
        L.check(lib.htn_tensor_download(self.h, out.ctypes.data, out.size), self.ctx.h)
        return out

    def block_views(self, packed: np.ndarray) -> dict:
        """{(label0,label1,label2): 2-D view into `packed`}."""
        out = {}
        for i in range(self.nblocks):
            o, r, c = int(self.offsets[i]), int(self.rows[i]), int(self.cols[i])
            out[tuple(int(v) for v in self.labels[i])] = packed[o:o + r * c].reshape(r, c)
        return out

    # -- vector algebra ------------------------------------------------------------------
    def dot(self, other: "Tensor") -> float:
        out = C.c_double()
        L.check(lib.htn_tensor_dot(self.h, other.h, C.byref(out)), self.ctx.h)
        return out.value

    def axpby(self, alpha: float, x: "Tensor", beta: float):
        """self = alpha * x + beta * self"""
        L.check(lib.htn_tensor_axpby(alpha, x.h, beta, self.h), self.ctx.h)
        return self

    def __del__(self):
        try:
            if self.h and self.ctx.h:      # a closed context has already released the device
                lib.htn_tensor_destroy(self.h)
            self.h = None
        except Exception:
            pass


class Mpo:
    """One site of the MPO Hamiltonian in reduced form: entries {(a,s',s,b,c): w}."""

    @staticmethod
    def from_dense(ctx: Context, Ml: Legs, P: Legs, Mr: Legs, dense: np.ndarray, tol: float = 1e-12) -> "Mpo":
        """From the dense invariant tensor dense[a,s',s,b] (multiplets expanded, m = -j..+j); the library
        does the Wigner-Eckart projection and rejects non-invariant input."""
        dense = np.ascontiguousarray(dense, dtype=np.float64)
        h = C.c_void_p()
        L.check(lib.htn_mpo_create_dense(ctx.h if ctx else None, Ml.h, P.h, Mr.h,
                                         dense.ctypes.data_as(C.POINTER(C.c_double)), tol, C.byref(h)),
                ctx.h if ctx else None)
        self = Mpo.__new__(Mpo)
        self.ctx, self.Ml, self.P, self.Mr, self.h = ctx, Ml, P, Mr, h
        n = C.c_int32()
        L.check(lib.htn_mpo_entries(h, C.byref(n), None, None, None), ctx.h if ctx else None)
        self.nnz = n.value
        return self

    def entries(self) -> dict:
        ch = self.ctx.h if self.ctx else None
        n = C.c_int32()
        L.check(lib.htn_mpo_entries(self.h, C.byref(n), None, None, None), ch)
        idx = np.zeros((n.value, 4), dtype=np.int32)
        cl = np.zeros((n.value, 3), dtype=np.int32)
        val = np.zeros(n.value)
        pi32 = C.POINTER(C.c_int32)
        L.check(lib.htn_mpo_entries(self.h, C.byref(n), idx.ctypes.data_as(pi32), cl.ctypes.data_as(pi32),
                                    val.ctypes.data_as(C.POINTER(C.c_double))), ch)
        return {(int(i[0]), int(i[1]), int(i[2]), int(i[3]), tuple(int(v) for v in c)): float(w)
                for i, c, w in zip(idx, cl, val)}

    def __init__(self, ctx: Context, Ml: Legs, P: Legs, Mr: Legs, entries: dict):
        self.ctx, self.Ml, self.P, self.Mr = ctx, Ml, P, Mr
        keys = list(entries.keys())
        idx = np.array([[k[0], k[1], k[2], k[3]] for k in keys], dtype=np.int32).reshape(-1, 4)
        cl = np.array([list(k[4]) for k in keys], dtype=np.int32).reshape(-1, 3)
        val = np.array([entries[k] for k in keys], dtype=np.float64)
        h = C.c_void_p()
        L.check(lib.htn_mpo_create(ctx.h, Ml.h, P.h, Mr.h, len(keys),
                                   idx.ctypes.data_as(C.POINTER(C.c_int32)),
                                   cl.ctypes.data_as(C.POINTER(C.c_int32)),
                                   val.ctypes.data_as(C.POINTER(C.c_double)), C.byref(h)), ctx.h)
        self.h = h
        self.nnz = len(keys)

    def __del__(self):
        try:
            if self.h and (self.ctx is None or self.ctx.h):
                lib.htn_mpo_destroy(self.h)
            self.h = None
        except Exception:
            pass


PLAN_STAT_NAMES = ["flops", "flops_L", "flops_R", "n_gemm_L", "n_gemm_R", "n_mix_targets",
                   "n_mix_sources", "workspace_bytes", "n_tiles_L", "n_tiles_R", "padded_flops",
                   "launches_per_apply"]


class _Heff:
    """Common part of the effective-Hamiltonian plans (apply, eigsolve, timing)."""

    def _finish(self, like):
        st = (C.c_double * 12)()
        L.check(lib.htn_plan_stats(self.h, st, 12), self.ctx.h)
        self.stats = dict(zip(PLAN_STAT_NAMES, [float(v) for v in st]))
        self.nelem = like.nelem

    def eigsolve(self, x0: Tensor, x: Tensor, krylovdim: int = 30, tol: float = 1e-10, maxiter: int = 100):
        """Lowest eigenpair (KrylovKit `eigsolve(.., :SR, Lanczos)`): returns (eigenvalue, info)."""
        ev, res, napp = C.c_double(), C.c_double(), C.c_int32()
        rc = L.check(lib.htn_eigsolve(self.h, x0.h, x.h, krylovdim, tol, maxiter, C.byref(ev), C.byref(res),
                                      C.byref(napp)), self.ctx.h)
        return ev.value, dict(converged=rc == 0, residual=res.value, applies=napp.value)


class HeffAC(_Heff):
    """y = H_AC x  (MPSKit `AC_hamiltonian`): plan over fixed GL, W, GR."""

    def __init__(self, ctx: Context, GL: Tensor, W: Mpo, GR: Tensor, like: Tensor, nshards: int = 1, shard: int = 0):
        """`nshards` > 1: this plan computes shard `shard`'s PARTIAL y (sum over the shards = H_AC x)."""
        self.ctx, self.GL, self.W, self.GR = ctx, GL, W, GR   # keep GL/GR alive
        h = C.c_void_p()
        L.check(lib.htn_plan_heff_ac_sharded(ctx.h, GL.h, W.h, GR.h, like.h, nshards, shard, C.byref(h)), ctx.h)
        self.h = h
        self._finish(like)

    def apply(self, x: Tensor, y: Tensor):
        L.check(lib.htn_heff_apply(self.h, x.h, y.h), self.ctx.h)
        return y

    def apply_host(self, x_host: np.ndarray, y_host: np.ndarray):
        """Reference-facing call with HOST buffers (upload, apply, download)."""
        assert x_host.dtype == np.float64 and y_host.dtype == np.float64
        L.check(lib.htn_heff_apply_host(self.h, x_host.ctypes.data, y_host.ctypes.data, x_host.size),
                self.ctx.h)
        return y_host

    def apply_host_ptr(self, x_ptr: int, y_ptr: int, nelem: int):
        L.check(lib.htn_heff_apply_host(self.h, x_ptr, y_ptr, nelem), self.ctx.h)

    def time(self, x: Tensor, y: Tensor, reps: int) -> float:
        """Device time (ms, CUDA events on the library stream) of `reps` back-to-back applies."""
        ms = C.c_float()
        L.check(lib.htn_heff_time(self.h, x.h, y.h, reps, C.byref(ms)), self.ctx.h)
        return ms.value

    def profile(self, x: Tensor, y: Tensor, reps: int = 10):
        ms = (C.c_float * 4)()
        L.check(lib.htn_plan_profile(self.h, x.h, y.h, reps, ms), self.ctx.h)
        return {"total_ms": ms[0], "stage_L_ms": ms[1], "stage_W_ms": ms[2], "stage_R_ms": ms[3]}

    def __del__(self):
        try:
            if self.h and self.ctx.h:      # a closed context has already released the device
                lib.htn_plan_destroy(self.h)
            self.h = None
        except Exception:
            pass


class HeffAC2(_Heff):
    """y2 = H_AC2 x2  (MPSKit `AC2_hamiltonian`): GL of the first site, GR of the second site."""

    def __init__(self, ctx: Context, GL: Tensor, W1: Mpo, W2: Mpo, GR: Tensor, like: Tensor):
        self.ctx, self.GL, self.GR, self.W1, self.W2 = ctx, GL, GR, W1, W2
        h = C.c_void_p()
        L.check(lib.htn_plan_heff_ac2(ctx.h, GL.h, W1.h, W2.h, GR.h, like.h, C.byref(h)), ctx.h)
        self.h = h
        self._finish(like)

    def apply(self, x: Tensor, y: Tensor):
        L.check(lib.htn_heff_apply(self.h, x.h, y.h), self.ctx.h)
        return y

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                lib.htn_plan_destroy(self.h)
            self.h = None
        except Exception:
            pass


class HeffC(_Heff):
    """y = H_C x  (MPSKit `C_hamiltonian`): GL on the bond of C (left env of the next site), GR of this site."""

    def __init__(self, ctx: Context, GL: Tensor, GR: Tensor, like: Tensor):
        self.ctx, self.GL, self.GR = ctx, GL, GR
        h = C.c_void_p()
        L.check(lib.htn_plan_heff_c(ctx.h, GL.h, GR.h, like.h, C.byref(h)), ctx.h)
        self.h = h
        self._finish(like)

    def apply(self, x: Tensor, y: Tensor):
        L.check(lib.htn_heff_apply(self.h, x.h, y.h), self.ctx.h)
        return y

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                lib.htn_plan_destroy(self.h)
            self.h = None
        except Exception:
            pass


class Transfer:
    """Environment transfer through one site (MPSKit `TransferMatrix`): side 0 = left, 1 = right."""

    def __init__(self, ctx: Context, side: int, W: Mpo, A: Tensor, At: Tensor, env_in: Tensor, env_out: Tensor):
        self.ctx, self.side, self.W = ctx, side, W
        h = C.c_void_p()
        L.check(lib.htn_plan_transfer(ctx.h, side, W.h, A.h, At.h, env_in.h, env_out.h, C.byref(h)), ctx.h)
        self.h = h
        st = (C.c_double * 12)()
        L.check(lib.htn_plan_stats(h, st, 12), ctx.h)
        self.stats = dict(zip(PLAN_STAT_NAMES, [float(v) for v in st]))

    def apply(self, A: Tensor, At: Tensor, env_in: Tensor, env_out: Tensor):
        L.check(lib.htn_transfer_apply(self.h, A.h, At.h, env_in.h, env_out.h), self.ctx.h)
        return env_out

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                lib.htn_plan_destroy(self.h)
            self.h = None
        except Exception:
            pass


def _harr(tensors):
    arr = (C.c_void_p * len(tensors))(*[t.h for t in tensors])
    return arr


class _OwnedSpace(Space):
    """Space created by the library (htn_tsvd); wraps an existing handle."""

    def __init__(self, ctx, sym, handle):
        self.ctx, self.sym, self.h = ctx, sym, handle
        n = C.c_int32()
        L.check(lib.htn_space_info(handle, C.byref(n), None, None), ctx.h)
        lab2 = np.zeros((n.value, 3), dtype=np.int32)
        mul2 = np.zeros(n.value, dtype=np.int32)
        L.check(lib.htn_space_info(handle, C.byref(n), lab2.ctypes.data_as(C.POINTER(C.c_int32)),
                                   mul2.ctypes.data_as(C.POINTER(C.c_int32))), ctx.h)
        self.sectors = [tuple(int(v) for v in row) for row in lab2]
        self.mult = [int(v) for v in mul2]


def contract_two_site(A1: Tensor, A2: Tensor, x2: Tensor) -> Tensor:
    L.check(lib.htn_contract_two_site(A1.h, A2.h, x2.h), A1.ctx.h)
    return x2


def tsvd(x2: Tensor, cut: float = 0.0, maxdim: int = 0, sym: int = 0):
    """x2 = AL . C . AR per middle sector with global truncation (TensorKit `tsvd!` + `truncbelow`)."""
    ctx = x2.ctx
    hv, hal, hc, har = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    disc, kept = C.c_double(), C.c_int32()
    L.check(lib.htn_tsvd(x2.h, cut, maxdim, C.byref(hv), C.byref(hal), C.byref(hc), C.byref(har), C.byref(disc),
                         C.byref(kept)), ctx.h)
    return (_OwnedSpace(ctx, sym, hv), Tensor(ctx, hal), Tensor(ctx, hc), Tensor(ctx, har),
            dict(discarded_weight=disc.value, kept=kept.value))


def qrpos(A: Tensor, Q: Tensor, R: Tensor):
    L.check(lib.htn_qrpos(A.h, Q.h, R.h), A.ctx.h)


def lqpos(A: Tensor, Lm: Tensor, Q: Tensor):
    L.check(lib.htn_lqpos(A.h, Lm.h, Q.h), A.ctx.h)


def regauge(AC: Tensor, Cb: Tensor, AL: Tensor):
    L.check(lib.htn_regauge(AC.h, Cb.h, AL.h), AC.ctx.h)


def gauge_right(ctx: Context, AL, C_guess: Tensor, AR, Cs, tol=1e-13, maxiter=10000):
    it, delta = C.c_int32(), C.c_double()
    rc = L.check(lib.htn_gauge_right(ctx.h, len(AL), _harr(AL), C_guess.h, _harr(AR), _harr(Cs), tol, maxiter,
                                     C.byref(it), C.byref(delta)), ctx.h)
    return dict(converged=rc == 0, iterations=it.value, delta=delta.value)


def environments(ctx: Context, AL, AR, Cs, Ws, GL, GR, tol=1e-12, krylovdim=30, maxiter=200):
    el, er = C.c_double(), C.c_double()
    rc = L.check(lib.htn_environments(ctx.h, len(AL), _harr(AL), _harr(AR), _harr(Cs), _harr(Ws), _harr(GL), _harr(GR),
                                      tol, krylovdim, maxiter, C.byref(el), C.byref(er)), ctx.h)
    return dict(converged=rc == 0, energy_cell_left=el.value, energy_cell_right=er.value)


def vumps(ctx: Context, AL, AR, Cs, AC, Ws, GL, GR, tol=1e-10, maxiter=100, krylovdim=30):
    """`find_groundstate(psi, H, VUMPS(; tol, maxiter))` on fixed bond spaces (in/out tensors)."""
    delta, e, it = C.c_double(), C.c_double(), C.c_int32()
    log = np.zeros((maxiter, 8))
    rc = L.check(lib.htn_vumps(ctx.h, len(AL), _harr(AL), _harr(AR), _harr(Cs), _harr(AC), _harr(Ws), _harr(GL),
                               _harr(GR), tol, maxiter, krylovdim, C.byref(delta), C.byref(e), C.byref(it),
                               log.ctypes.data_as(C.POINTER(C.c_double)), maxiter), ctx.h)
    return dict(converged=rc == 0, delta=delta.value, energy_per_site=e.value, iterations=it.value,
                log=log[:it.value])


def gradient_grassmann(ctx: Context, AL, AR, Cs, AC, Ws, GL, GR, tol=1e-10, maxiter=100, krylovdim=30):
    """`find_groundstate(psi, H, GradientGrassmann(; tol, maxiter))` on fixed bond spaces (in/out tensors):
    the polish stage of HF:1025-1027."""
    delta, e, it = C.c_double(), C.c_double(), C.c_int32()
    log = np.zeros((maxiter, 8))
    rc = L.check(lib.htn_gradient_grassmann(ctx.h, len(AL), _harr(AL), _harr(AR), _harr(Cs), _harr(AC), _harr(Ws),
                                            _harr(GL), _harr(GR), tol, maxiter, krylovdim, C.byref(delta), C.byref(e),
                                            C.byref(it), log.ctypes.data_as(C.POINTER(C.c_double)), maxiter), ctx.h)
    return dict(converged=rc == 0, delta=delta.value, energy_per_site=e.value, iterations=it.value,
                log=log[:it.value])


def idmrg2(ctx: Context, AL, AR, Cs, AC, Ws, cut=1e-2, tol=1e-6, maxiter=100, krylovdim=30, eig_tol=1e-8, maxdim=0):
    """`find_groundstate(psi, H, IDMRG2(trscheme=truncbelow(cut), tol))` (HF:1010).  The bond spaces
    change, so the tensors are REPLACED: returns new (AL, AR, C, AC) lists and an info dict."""
    lists = [list(AL), list(AR), list(Cs), list(AC)]
    arrs = [_harr(x) for x in lists]
    delta, it = C.c_double(), C.c_int32()
    log = np.zeros((maxiter, 8))
    rc = lib.htn_idmrg2(ctx.h, len(AL), arrs[0], arrs[1], arrs[2], arrs[3], _harr(Ws), cut, tol, maxiter, krylovdim,
                        eig_tol, maxdim, C.byref(delta), C.byref(it), log.ctypes.data_as(C.POINTER(C.c_double)), maxiter)
    out = []
    for x, arr in zip(lists, arrs):
        new = []
        for i, t in enumerate(x):
            # the old wrapper is always retired: its tensor was either destroyed by the library or is
            # re-wrapped here (a recycled address may carry a different block structure)
            t.h = None
            new.append(Tensor(ctx, C.c_void_p(arr[i])))
        out.append(new)
    L.check(rc, ctx.h)
    return out[0], out[1], out[2], out[3], dict(converged=rc == 0, delta=delta.value, iterations=it.value,
                                                 log=log[:it.value])


def changebonds_svdcut(ctx: Context, AL, AR, Cs, AC, Ws, cut=0.0, maxdim=0, sym=0):
    """`changebonds(psi, SvdCut(trscheme))` (HF:1013,1018,1365): one truncation-only two-site sweep (every
    bond is re-split by the truncated SVD, no eigensolve), then the gauge fixing of `InfiniteMPS(psi.AR)`.
    Returns new (AL, AR, C, AC)."""
    AL, AR, Cs, AC, _ = idmrg2(ctx, AL, AR, Cs, AC, Ws, cut=cut, tol=0.0, maxiter=1, krylovdim=0, maxdim=maxdim)
    return uniform_from_right(ctx, AR, Cs[-1], sym)


def changebonds(ctx: Context, kind: int, AL, AR, Cs, AC, Ws, cut: float = 0.0, maxdim: int = 0, krylovdim: int = 30,
                eig_tol: float = 1e-10, tol_gauge: float = 1e-12):
    """`changebonds(psi, SvdCut(trscheme))` (kind 0) / `changebonds(psi, H, VUMPSSvdCut(trscheme))` (kind 1) in ONE library
    call (htn_changebonds; HF:1013-1018, 1363-1365).  maxdim > 0 caps the multiplets per bond, maxdim < 0 the full
    dimension (`truncdim(|maxdim|)`).  The inputs are left untouched (the library works on copies made here); returns new
    (AL, AR, C, AC) lists."""
    lists = [[t.like_copy() for t in lst] for lst in (AL, AR, Cs, AC)]
    arrs = [_harr(x) for x in lists]
    rc = lib.htn_changebonds(ctx.h, kind, len(AL), arrs[0], arrs[1], arrs[2], arrs[3], _harr(Ws), cut, maxdim, krylovdim,
                             eig_tol, tol_gauge)
    out = []
    for x, arr in zip(lists, arrs):
        new = []
        for i, t in enumerate(x):
            t.h = None                                   # retired: destroyed or re-wrapped below
            new.append(Tensor(ctx, C.c_void_p(arr[i])))
        out.append(new)
    L.check(rc, ctx.h)
    return out[0], out[1], out[2], out[3]


def mul_bond(A: Tensor, Cb: Tensor, right: bool = True) -> Tensor:
    """A . C (right) or C . A (left) as a new MPS tensor (MPSKit `_mul_tail` / `_mul_front`)."""
    h = C.c_void_p()
    L.check(lib.htn_mul_bond(A.h, Cb.h, 1 if right else 0, C.byref(h)), A.ctx.h)
    return Tensor(A.ctx, h)


def _identity_bond(ctx: Context, V: Space) -> Tensor:
    t = Tensor.bond(ctx, V)
    g = np.zeros(t.nelem)
    for blk in t.block_views(g).values():
        blk[...] = np.eye(blk.shape[0])
    return t.upload(g)


def changebonds_vumpssvdcut(ctx: Context, AL, AR, Cs, AC, Ws, P: Legs, sym: int = 0, cut: float = 0.0, maxdim: int = 0,
                            krylovdim: int = 30, eig_tol: float = 1e-10, tol_gauge: float = 1e-12):
    """`changebonds(psi, H, VUMPSSvdCut(; trscheme))` for unit cells of two or more sites (MPSKit `changebonds_n`;
    HF:1016, 1363): for every site loc the two-site tensor AC[loc] AR[loc+1] is replaced by the lowest eigenvector of
    H_AC2, the bond matrix C[loc+1] by that of H_C, the two-site tensor is split by a truncated SVD (AL1, S V), the
    second site is regauged as in VUMPS (AL2 = Q(S V) Q(C)^T), and the state and its environments are rebuilt from the
    new left isometries before the next site.  Difference to MPSKit: the bonds are first brought to the cap by SvdCut
    (see below).  Inputs are left untouched; returns new (AL, AR, C, AC, GL, GR)."""
    n = len(AL)
    if n < 2:
        raise NotImplementedError("VUMPSSvdCut on a one-site unit cell (MPSKit changebonds_1) is not mirrored")
    AL, AR, Cs, AC = [[t.like_copy() for t in lst] for lst in (AL, AR, Cs, AC)]
    if maxdim > 0 or cut > 0.0:
        # MPSKit rebuilds the state with InfiniteMPS(...) after every site, which shrinks the neighbouring bonds to
        # full rank; the positive QR used here needs that up front, so all bonds are first cut to the same cap by
        # SvdCut and the sweep below re-optimises and re-cuts them at that size.
        AL, AR, Cs, AC = changebonds_svdcut(ctx, AL, AR, Cs, AC, Ws, cut=cut, maxdim=maxdim, sym=sym)
    chi = len(Ws[0].Ml)

    def envs():
        V = [Cs[i].space(0, sym) for i in range(n)]
        GL = [Tensor.env(ctx, 0, V[i - 1], Ws[i].Ml, identity_level=0) for i in range(n)]
        GR = [Tensor.env(ctx, 1, V[i], Ws[i].Mr, identity_level=chi - 1) for i in range(n)]
        environments(ctx, AL, AR, Cs, Ws, GL, GR, tol=1e-10)
        return GL, GR

    GL, GR = envs()
    for loc in range(n):
        nxt, nn = (loc + 1) % n, (loc + 2) % n
        x2 = Tensor.mps2(ctx, AC[loc].space(0, sym), P, P, AR[nxt].space(1, sym))
        contract_two_site(AC[loc], AR[nxt], x2)
        y2 = x2.like()
        HeffAC2(ctx, GL[loc], Ws[loc], Ws[nxt], GR[nxt], x2).eigsolve(x2, y2, krylovdim, eig_tol)
        nC = Cs[nxt].like()
        HeffC(ctx, GL[nn], GR[nxt], Cs[nxt]).eigsolve(Cs[nxt], nC, krylovdim, eig_tol)
        _, AL1, S, V, _ = tsvd(y2, cut, maxdim, sym)
        ACn = mul_bond(V, S, right=False)
        AL2 = ACn.like()
        regauge(ACn, nC, AL2)
        AL[loc], AL[nxt] = AL1, AL2
        AR = [a.like() for a in AL]
        AC = [a.like() for a in AL]
        Cs = [Tensor.bond(ctx, AL[i].space(1, sym)) for i in range(n)]
        mixed_gauge(ctx, AL, _identity_bond(ctx, AL[n - 1].space(1, sym)), AR, Cs, AC, tol=tol_gauge)
        GL, GR = envs()
    return AL, AR, Cs, AC, GL, GR


def mixed_gauge(ctx: Context, AL, C_guess: Tensor, AR, Cs, AC, tol=1e-12, maxiter=10000, from_right=False):
    """`InfiniteMPS(A...)` gauge fixing.  from_right=False: AL holds left isometries (in/out), AR, C, AC are
    outputs.  from_right=True: AR holds right isometries (MPSKit `InfiniteMPS(psi.AR)` after IDMRG2); AL, C, AC
    must have AR's block structure / right bond spaces and are overwritten."""
    it = C.c_int32()
    rc = L.check(lib.htn_mixed_gauge(ctx.h, len(AL), _harr(AL), C_guess.h, _harr(AR), _harr(Cs), _harr(AC),
                                     1 if from_right else 0, tol, maxiter, C.byref(it)), ctx.h)
    return dict(converged=rc == 0, iterations=it.value)


def uniform_from_right(ctx: Context, AR, C_last: Tensor, sym: int = 0, tol=1e-12):
    """Consistent mixed-gauge uniform MPS from the right isometries an IDMRG2 run ends with: creates AL, AC
    (structure of AR[i]) and C (bond tensors on the right spaces) and gauge-fixes.  Returns (AL, AR, C, AC)."""
    n = len(AR)
    AL = [a.like() for a in AR]
    AC = [a.like() for a in AR]
    Cs = [Tensor.bond(ctx, AR[i].space(1, sym)) for i in range(n)]
    mixed_gauge(ctx, AL, C_last, AR, Cs, AC, tol=tol, from_right=True)
    return AL, AR, Cs, AC


def entanglement_spectrum(Cb: Tensor) -> dict:
    """{sector position c: Schmidt values (descending)} of a bond matrix, computed on the device."""
    n = int(sum(int(r) for r in Cb.rows))
    out = np.zeros(n)
    L.check(lib.htn_entanglement_spectrum(Cb.h, out.ctypes.data_as(C.POINTER(C.c_double)), n), Cb.ctx.h)
    res, o = {}, 0
    for i in range(Cb.nblocks):
        r = int(Cb.rows[i])
        res[int(Cb.labels[i][0])] = out[o:o + r].copy()
        o += r
    return res


def expval_diag(AC: Tensor, values) -> float:
    v = np.ascontiguousarray(values, dtype=np.float64)
    out = C.c_double()
    L.check(lib.htn_expval_diag(AC.h, v.ctypes.data_as(C.POINTER(C.c_double)), v.size, C.byref(out)), AC.ctx.h)
    return out.value


def probe_krylov(like: Tensor, nvec: int = 30, reps: int = 20) -> dict:
    """Device-timed Gram-Schmidt pass (multidot + multiaxpy) on vectors shaped like `like`."""
    ms = (C.c_float * 2)()
    by = (C.c_double * 2)()
    L.check(lib.htn_probe_krylov(like.h, nvec, reps, ms, by), like.ctx.h)
    return {"multidot_ms": ms[0], "multiaxpy_ms": ms[1], "multidot_GBs": by[0] / (ms[0] * 1e-3) / 1e9,
            "multiaxpy_GBs": by[1] / (ms[1] * 1e-3) / 1e9, "nvec": nvec, "vector_bytes": by[0] / (nvec + 1)}


def network_coefficient(sym: int, nine_labels) -> float:
    lab, plab = _i32arr(np.array(nine_labels, dtype=np.int32).reshape(9, 3))
    out = C.c_double()
    L.check(lib.htn_network_coefficient(sym, plab, C.byref(out)))
    return out.value
