"""TEST / BASELINE INFRASTRUCTURE (not the product): native-thread CPU execution of oracle.heff.HeffACPlan.

`NativeHeffAC(plan)` flattens the plan's tensors and its GEMM / mix lists into arrays and runs one apply through
oracle/native/libheff_cpu.so: single-threaded OpenBLAS dgemm per block, one pthread per core over the blocks -- the
reference's CPU policy (BLAS threads = 1 and all threads over sector blocks, /root/reference/src/HubbardFunctions.jl:29,37)
without the interpreter lock that bounds the pure-numpy `HeffACPlan.apply(threads=N)`.  Checked against that numpy path in
tests/test_oracle_native.py; timed by bench.py (cpu_baseline and --impl reference)."""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libheff_cpu.so")


def build():
    subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


def _openblas_dgemm():
    """Address of numpy's bundled ILP64 OpenBLAS `cblas_dgemm` and its set-num-threads call."""
    cands = glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so"))
    if not cands:
        raise OSError("numpy's bundled OpenBLAS was not found")
    blas = C.CDLL(os.path.realpath(cands[0]), mode=C.RTLD_GLOBAL)
    fn = getattr(blas, "scipy_cblas_dgemm64_")
    setn = getattr(blas, "scipy_openblas_set_num_threads64_")
    setn.argtypes = [C.c_int]
    return blas, C.cast(fn, C.c_void_p), setn


class _Gemm(C.Structure):
    _fields_ = [("a_off", C.c_int64), ("b_off", C.c_int64), ("c_off", C.c_int64), ("m", C.c_int32), ("n", C.c_int32),
                ("k", C.c_int32), ("pad", C.c_int32)]


class _Mix(C.Structure):
    _fields_ = [("dst_off", C.c_int64), ("dst_space", C.c_int32), ("nelem", C.c_int32), ("src_begin", C.c_int32),
                ("src_end", C.c_int32)]


class _Src(C.Structure):
    _fields_ = [("off", C.c_int64), ("space", C.c_int32), ("pad", C.c_int32), ("coef", C.c_double)]


class NativeHeffAC:
    def __init__(self, plan):
        if not os.path.exists(LIB):
            build()
        self.lib = C.CDLL(LIB)
        self.blas, self.dgemm, self.set_threads = _openblas_dgemm()
        self.plan = plan
        GL, GR, x = plan.GL, plan.GR, plan.x0

        def flatten(blocks):
            off, o = {}, 0
            for k, b in blocks.items():
                off[k] = o
                o += b.size
            flat = np.empty(max(o, 1))
            for k, b in blocks.items():
                flat[off[k]:off[k] + b.size] = np.ascontiguousarray(b).ravel()
            return flat, off

        self.gl, gl_off = flatten(GL.blocks)
        self.gr, gr_off = flatten(GR.blocks)
        self.x_off, o = {}, 0
        for k in x.keys:
            self.x_off[k] = o
            o += x.blocks[k].size
        self.nx = o
        Vl, Vr = x.Vl, x.Vr
        # T and U workspaces
        t_off, o = [], 0
        for (a, lp, l, s, r) in plan.t_list:
            t_off.append(o)
            o += Vl.mult[lp] * Vr.mult[r]
        self.T = np.zeros(max(o, 1))
        u_off, o = [], 0
        for (b, lp, sp, rp, r) in plan.u_list:
            u_off.append(o)
            o += Vl.mult[lp] * Vr.mult[r]
        self.U = np.zeros(max(o, 1))
        # stage L
        gl = (_Gemm * max(len(plan.t_list), 1))()
        for i, (a, lp, l, s, r) in enumerate(plan.t_list):
            gl[i] = _Gemm(gl_off[(a, lp, l)], self.x_off[(l, s, r)], t_off[i], Vl.mult[lp], Vr.mult[r], Vl.mult[l], 0)
        self.c_gl, self.ngl = gl, len(plan.t_list)
        # stage W
        mixes, srcs = [], []
        for dst, lst in plan.mix.items():
            b0 = len(srcs)
            for (kind, i), cf in lst:
                srcs.append(_Src(t_off[i] if kind == "T" else self.x_off[i], 1 if kind == "T" else 0, 0, float(cf)))
            if dst[0] == "U":
                b, lp, sp, rp, r = plan.u_list[dst[1]]
                mixes.append(_Mix(u_off[dst[1]], 2, Vl.mult[lp] * Vr.mult[r], b0, len(srcs)))
            else:
                lp, sp, rp = dst[1]
                mixes.append(_Mix(self.x_off[dst[1]], 3, Vl.mult[lp] * Vr.mult[rp], b0, len(srcs)))
        self.c_mix = (_Mix * max(len(mixes), 1))(*mixes)
        self.c_src = (_Src * max(len(srcs), 1))(*srcs)
        self.nmix = len(mixes)
        # stage R grouped by y block, heavy groups first
        by_y = {}
        for i, (b, lp, sp, rp, r) in enumerate(plan.u_list):
            by_y.setdefault((lp, sp, rp), []).append(
                _Gemm(u_off[i], gr_off[(b, r, rp)], self.x_off[(lp, sp, rp)], Vl.mult[lp], Vr.mult[rp], Vr.mult[r], 0))
        groups = sorted(by_y.values(), key=lambda g: -sum(t.m * t.n * t.k for t in g))
        flat = [t for g in groups for t in g]
        self.c_gr = (_Gemm * max(len(flat), 1))(*flat)
        offs = np.zeros(len(groups) + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(g) for g in groups])
        self.gr_group, self.ngroups = offs, len(groups)
        self.lib.htn_cpu_heff_apply.restype = C.c_int
        self.lib.htn_cpu_heff_apply.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int64] + [C.c_void_p] * 2 + [
            C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]

    def pack_x(self, x) -> np.ndarray:
        out = np.empty(self.nx)
        for k, o in self.x_off.items():
            out[o:o + x.blocks[k].size] = x.blocks[k].ravel()
        return out

    def unpack_y(self, flat, like):
        y = like.zeros_like()
        for k, o in self.x_off.items():
            y.blocks[k] = flat[o:o + y.blocks[k].size].reshape(y.blocks[k].shape).copy()
        return y

    def apply_flat(self, xf: np.ndarray, yf: np.ndarray, threads: int):
        """One apply on flat arrays (canonical block order of x); BLAS threads = 1, `threads` workers."""
        self.set_threads(1)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = self.lib.htn_cpu_heff_apply(self.dgemm, p(self.gl), p(self.gr), p(xf), p(yf), yf.size, p(self.T), p(self.U),
                                         C.cast(self.c_gl, C.c_void_p), self.ngl, C.cast(self.c_mix, C.c_void_p),
                                         C.cast(self.c_src, C.c_void_p), self.nmix, C.cast(self.c_gr, C.c_void_p),
                                         p(self.gr_group), self.ngroups, threads)
        if rc != 0:
            raise RuntimeError("htn_cpu_heff_apply failed")
        return yf
