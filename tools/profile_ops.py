#!/usr/bin/env python
"""Runs ONE of the non-headline device operations of the path a few times at D_red = 1024 (U1xSU2, synthetic spaces of
SURVEY.md 8(d)) so that `ncu --set full` can capture its kernels in steady state, and prints device-event timings.

usage: python tools/profile_ops.py {ac2|transfer|krylov|qr|svd} [chi]
  ac2      H_AC2 apply (two-site effective Hamiltonian, MPSKit `∂AC2`)           kernels: stack_gemm / grouped_gemm / mix
  transfer left environment transfer T_L (MPSKit `TransferMatrix`)               kernels: grouped_gemm / mix
  krylov   one Gram-Schmidt pass on 30 basis vectors of len(AC2) = 10 MB         kernels: multidot_partial / multiaxpy
  qr       positive QR of an MPS tensor per coupled sector (`leftorth!(QRpos)`)  kernel:  qr_bcgs2
  svd      truncated SVD of a two-site tensor (`tsvd!`)                          kernels: svd_round (one per Jacobi round)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hubbardtn_b200 import device as dev, sectors as S, synthetic as syn

op = sys.argv[1] if len(sys.argv) > 1 else "ac2"
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 24
D = int(os.environ.get("D_RED", "1024"))
ctx = dev.Context(0)
sym = S.SU2U1
phys = S.physical_space(sym, 1, 1)
levels = syn.mpo_levels(sym, chi)
Va = dev.Space(ctx, sym, syn.bond_space(sym, D, 0))
Vb = dev.Space(ctx, sym, syn.bond_space(sym, D, 1))
P, M = dev.Legs(ctx, sym, phys), dev.Legs(ctx, sym, levels)


def rnd(t, stream, scale=1.0):
    return t.upload(syn.random_packed(t.nelem, stream) * scale)


def env(side, V, ident, stream):
    t = dev.Tensor.env(ctx, side, V, M, identity_level=ident)
    h = syn.random_packed(t.nelem, stream) / np.sqrt(D)
    for (a, i, j), blk in t.block_views(h).items():
        if a == ident:
            blk[...] = np.eye(blk.shape[0])
    return t.upload(h)


def timed(fn, reps=3):
    fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


if op == "ac2":
    W1 = dev.Mpo(ctx, M, P, M, syn.mpo_entries(sym, levels, phys, 4, syn.SEED))
    W2 = dev.Mpo(ctx, M, P, M, syn.mpo_entries(sym, levels, phys, 4, syn.SEED + 1))
    GL, GR = env(0, Va, 0, 1), env(1, Va, chi - 1, 2)          # sites (0, 1): left bond A, middle B, right A
    x = rnd(dev.Tensor.mps2(ctx, Va, P, P, Va), 3)
    y = x.like()
    plan = dev.HeffAC2(ctx, GL, W1, W2, GR, x)
    ms = timed(lambda: plan.apply(x, y))
    print("H_AC2 apply D_red=%d chi=%d: %.3f ms, %.2f GF algorithmic -> %.2f TFLOP/s; len(x2) = %.1f MB"
          % (D, chi, ms, plan.stats["flops"] / 1e9, plan.stats["flops"] / ms / 1e9, x.nelem * 8 / 1e6))
elif op == "transfer":
    W = dev.Mpo(ctx, M, P, M, syn.mpo_entries(sym, levels, phys, 4, syn.SEED))
    A = rnd(dev.Tensor.mps(ctx, Va, P, Vb), 3, 1.0 / np.sqrt(D))
    At = A.transposed()
    A.transpose_into(At)
    gin, gout = env(0, Va, 0, 1), dev.Tensor.env(ctx, 0, Vb, M, identity_level=0)
    tr = dev.Transfer(ctx, 0, W, A, At, gin, gout)
    ms = timed(lambda: tr.apply(A, At, gin, gout))
    print("left transfer D_red=%d chi=%d: %.3f ms; env bytes in %.1f MB out %.1f MB" % (D, chi, ms, gin.nelem * 8 / 1e6, gout.nelem * 8 / 1e6))
elif op == "krylov":
    x = rnd(dev.Tensor.mps2(ctx, Va, P, P, Va), 3)
    r = dev.probe_krylov(x, nvec=30, reps=10)
    print("Gram-Schmidt pass on 30 vectors of %.1f MB (basis %.0f MB > 126 MB L2): multidot %.3f ms = %.0f GB/s, multiaxpy %.3f ms = %.0f GB/s"
          % (r["vector_bytes"] / 1e6, 31 * r["vector_bytes"] / 1e6, r["multidot_ms"], r["multidot_GBs"], r["multiaxpy_ms"], r["multiaxpy_GBs"]))
elif op == "qr":
    A = dev.Tensor.mps(ctx, Va, P, Vb)
    Q, R = A.like(), dev.Tensor.bond(ctx, Vb)
    h = syn.random_packed(A.nelem, 3)

    def run():
        A.upload(h)
        dev.qrpos(A, Q, R)
    ms = timed(run)
    print("QRpos of an MPS tensor D_red=%d (%d coupled sectors, panels up to %d x %d): %.3f ms incl. upload"
          % (D, len(Vb.mult), 3 * max(Va.mult), max(Vb.mult), ms))
elif op == "svd":
    x = rnd(dev.Tensor.mps2(ctx, Va, P, P, Va), 3)
    t0 = time.perf_counter()
    out = dev.tsvd(x, 0.0, D, sym)
    ctx.synchronize()
    print("tsvd of a random two-site tensor D_red=%d (kept %d multiplets): %.1f ms" % (D, D, (time.perf_counter() - t0) * 1e3))
else:
    raise SystemExit(__doc__)
