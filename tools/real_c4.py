#!/usr/bin/env python
"""BASELINE config C4 with its REAL Hamiltonian: the reference's polyacetylene example (examples/polyacetylene.jl:29-31;
two bands, hopping + direct + exchange terms, U(1)xSU(2), 4-site unit cell, chi = 10 MPO levels) grown by IDMRG2 to
D_red = 1024 multiplets per bond, then one H_AC apply of every site is timed on the environments of that state.
usage: python tools/real_c4.py [D_cap] [sweeps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hubbardtn_b200 import device as dev, hubbardfunctions as hf

D_cap = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
t = np.array([[0.000, 3.803, -0.548, 0.000], [3.803, 0.000, 2.977, -0.501]])
U = np.array([[10.317, 6.264, 0.000, 0.000], [6.264, 10.317, 6.162, 0.000]])
J = np.array([[0.000, 0.123, 0.000, 0.000], [0.123, 0.000, 0.113, 0.000]])
model = hf.MB_Sim(t, U, J, None, 1, 1, 2.5, 20)
ctx = dev.Context(0)
H = hf.hamiltonian(model, ctx)
psi = hf.initialize_mps(H, model.P, model.bond_dim, False, ctx)
t0 = time.perf_counter()
AL, AR, C, AC, info = dev.idmrg2(ctx, psi.AL, psi.AR, psi.C, psi.AC, H.W, cut=1e-12, tol=1e-8, maxiter=sweeps, maxdim=D_cap)
AL, AR, C, AC = dev.uniform_from_right(ctx, AR, C[-1], model.sym)
ctx.synchronize()
print("IDMRG2 %d sweeps + gauge: %.1f s; D_red per bond %s, chi = %d levels" % (info["iterations"], time.perf_counter() - t0,
      [int(sum(c.space(0, model.sym).mult)) for c in C], H.chi))
psi = hf.InfiniteMPS(ctx, model.sym, AL, AR, C, AC, H.phys)
GL, GR = hf._make_envs(ctx, psi, H)
e = dev.environments(ctx, psi.AL, psi.AR, psi.C, H.W, GL, GR, tol=1e-10)
print("E/site %.8f, D_full %s" % (0.5 * (e["energy_cell_left"] + e["energy_cell_right"]) / len(psi), hf.dim_state(psi)))
peak = ctx.probe_fp64_peak(0)
for i in range(int(os.environ.get('NSITES', len(psi)))):
    plan = dev.HeffAC(ctx, GL[i], H.W[i], GR[i], psi.AC[i])
    y = psi.AC[i].like()
    plan.apply(psi.AC[i], y)
    ms = plan.time(psi.AC[i], y, 200) / 200
    f = plan.stats["flops"]
    if i == 0:
        print("   stages:", {k: round(v, 4) for k, v in plan.profile(psi.AC[i], y, 50).items()}, {k: int(v) for k, v in plan.stats.items() if k.startswith("n_")})
    print("site %d: H_AC apply %.4f ms, %.3f GF algorithmic -> %.2f TFLOP/s = %.2f of the DMMA peak (%.1f); %.0f applies/s"
          % (i, ms, f / 1e9, f / ms / 1e9, f / ms / 1e9 / peak, peak, 1e3 / ms))
