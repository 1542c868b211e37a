#!/usr/bin/env python
"""All shards of the library-level sharding of one C4 apply on ONE GPU: each shard's plan is applied, the partial results
are summed and compared with the unsharded apply.  usage: python tools/shard_check.py [nshards]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hubbardtn_b200 import device, sectors, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = device.Context(0)
case = synthetic.HeffCase(ctx, sectors.SU2U1, D=1024, chi=96)
case.plan.apply(case.x, case.y)
full = case.y.download()
acc = np.zeros_like(full)
for r in range(n):
    t0 = time.time()
    p = device.HeffAC(ctx, case.GL, case.W, case.GR, case.x, nshards=n, shard=r)
    print("shard %d/%d planned in %.1f s, flops %.2f GF" % (r, n, time.time() - t0, p.stats["flops"] / 1e9), flush=True)
    p.apply(case.x, case.y)
    ctx.synchronize()
    ms = p.time(case.x, case.y, 20) / 20
    acc += case.y.download()
    print("   apply %.3f ms" % ms, flush=True)
print("rel err of the shard sum:", float(np.abs(acc - full).max() / np.abs(full).max()))
