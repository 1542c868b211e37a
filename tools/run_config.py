#!/usr/bin/env python
"""Run the reference schedule (IDMRG2 -> VUMPS) for one of the BASELINE configs on the GPU with timing
per phase.  usage: python tools/run_config.py C2 [maxdim] [cut]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hubbardtn_b200 import device as dev, hubbardfunctions as hf

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
maxdim = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cut = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-5
models = {
    "C1": hf.OB_Sim([1.0], [8.0], 0.0, [0.0], 1, 1, 2.0),
    "C2": hf.OB_Sim([1.0, 0.2], [6.0], 0.0, [0.0], 1, 1, 5.0),
    "C3": hf.OB_Sim([1.0], [8.0], 0.0, [0.0], 1, 1, 5.0, kwargs={"spin": True}),
}
model = models[cfg]
ctx = dev.Context(0)
H = hf.hamiltonian(model, ctx)
psi = hf.initialize_mps(H, model.P, model.bond_dim, model.spin, ctx)
t0 = time.perf_counter()
AL, AR, C, AC, info1 = dev.idmrg2(ctx, psi.AL, psi.AR, psi.C, psi.AC, H.W, cut=cut, tol=1e-6, maxiter=int(os.environ.get("IDMRG_ITERS", "30")),
                                  maxdim=maxdim)
t1 = time.perf_counter()
print("IDMRG2: %.2f s, %d iterations, delta %.2e, D_red per bond %s, applies %d"
      % (t1 - t0, info1["iterations"], info1["delta"], [sum(c.space(0, model.sym).mult) for c in C], int(info1["log"][-1][2])))
print("   cumulative seconds: planning %.2f  lanczos %.2f  svd %.2f  env growth %.2f" % tuple(info1["log"][-1][3:7]))
if os.environ.get("STOP_AFTER_IDMRG"):
    sys.exit(0)
n = len(AR)
AL = [a.like() for a in AR]
AC = [a.like() for a in AR]
Cn = [dev.Tensor.bond(ctx, AR[i].space(1, model.sym)) for i in range(n)]
ginfo = dev.mixed_gauge(ctx, AL, C[-1], AR, Cn, AC, tol=1e-12, maxiter=int(os.environ.get("GAUGE_MAXITER", "10000")),
                        from_right=True)
C = Cn
print("mixed gauge:", ginfo)
t2 = time.perf_counter()
print("mixed gauge: %.2f s" % (t2 - t1))
psi = hf.InfiniteMPS(ctx, model.sym, AL, AR, C, AC)
GL, GR = hf._make_envs(ctx, psi, H)
info2 = dev.vumps(ctx, psi.AL, psi.AR, psi.C, psi.AC, H.W, GL, GR, tol=1e-8, maxiter=int(os.environ.get("VUMPS_ITERS", "40")))
t3 = time.perf_counter()
print("VUMPS: %.2f s, %d iterations, galerkin %.2e, E/site %.10f" % (t3 - t2, info2["iterations"], info2["delta"], info2["energy_per_site"]))
for row in info2["log"][:12]:
    print("   eps %.3e  E %.10f  gauge its %d  applies %d   t_eig %.3f t_gauge %.3f t_env %.3f (gmres applies %d)" % tuple(row))
print("D_full", hf.dim_state(psi), "density", hf.density_state(psi))
