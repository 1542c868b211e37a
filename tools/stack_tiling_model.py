#!/usr/bin/env python
"""Host model of the stacked stage-L tiling (csrc/htn_plan.cpp:build_front): algorithmic vs executed flops by cause.
usage: python tools/stack_tiling_model.py [D] [chi]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

class A: pass
args = A(); args.sym = 0; args.D = int(sys.argv[1]) if len(sys.argv) > 1 else 1024; args.chi = int(sys.argv[2]) if len(sys.argv) > 2 else 96
plan, x = bench.host_case(args)
GL, GR = plan.GL, plan.GR
nl = {i: m for i, m in enumerate(x.Vl.mult)}
nr = {i: m for i, m in enumerate(x.Vr.mult)}
idl = list(GL.identity_levels)[0]
# panel order: (l; identity last; l'; a)
panel = collections.defaultdict(list)
for (a, lp, l) in sorted(GL.blocks, key=lambda k: (k[2], k[0] == idl, k[1], k[0])):
    panel[l].append((a, lp))
prow = {}
for l, lst in panel.items():
    r = 0
    for (a, lp) in lst:
        prow[(a, lp, l)] = r
        r += nl[lp]
by_x = collections.defaultdict(list)
for (a, lp, l, s, r) in plan.t_list:
    by_x[(l, s, r)].append((a, lp))
alg = 0; ex_rows_gap = 0; tot_exec = 0; ntiles = 0; nruns = 0
alg_by_l = collections.Counter(); exec_by_l = collections.Counter()
def r4(k): return (k + 3) // 4 * 4
for (l, s, r), need in by_x.items():
    K, N = nl[l], nr[r]
    rows = sorted((prow[(a, lp, l)], nl[lp]) for (a, lp) in need)
    a_flops = sum(2 * m * K * N for _, m in rows)
    alg += a_flops; alg_by_l[l] += a_flops
    # runs with gaps < GAP merged (ignoring waves)
    GAP = int(os.environ.get("GAP", "32"))
    runs = []
    for p, m in rows:
        if runs and p - runs[-1][1] < GAP:
            runs[-1][1] = p + m
        else:
            runs.append([p, p + m])
    atoms = (N + 7) // 8
    for r0, r1 in runs:
        nruns += 1
        t = (r1 - r0 + 63) // 64
        ntiles += t
        e = 2 * t * 64 * atoms * 8 * r4(K)
        tot_exec += e; exec_by_l[l] += e
print("x blocks %d, T blocks %d, runs %d, tiles(64 rows x all cols) %d" % (len(by_x), len(plan.t_list), nruns, ntiles))
print("algorithmic stage L %.3f GF, executed (model, waves ignored) %.3f GF  (x%.3f)" % (alg / 1e9, tot_exec / 1e9, tot_exec / alg))
full = sum(2 * sum(nl[lp] for (a, lp) in panel[l] if a != idl) * nl[l] * nr[r] for (l, s, r) in by_x)
print("all non-identity panel rows for every x block: %.3f GF" % (full / 1e9))
for l in sorted(alg_by_l, key=lambda l: -alg_by_l[l])[:12]:
    print("  l=%d n_l=%d  alg %.3f GF exec %.3f (x%.2f)  panel rows %d" % (l, nl[l], alg_by_l[l] / 1e9, exec_by_l[l] / 1e9, exec_by_l[l] / alg_by_l[l], sum(nl[lp] for a, lp in panel[l])))
# needed-level pattern per physical sector s
pat = collections.defaultdict(set)
for (a, lp, l, s, r) in plan.t_list:
    pat[s].add(a)
for s in pat: print("s=%d needs %d of %d levels" % (s, len(pat[s]), args.chi))
print("levels needed by all s:", len(set.intersection(*pat.values())), " by exactly one:", sum(1 for a in range(args.chi) if sum(a in p for p in pat.values()) == 1))
# ---- mix statistics: sources per target, consumers per T block, weighted by elements ----
import numpy as np
nsrc_hist = collections.Counter(); cons = collections.Counter(); el_by_nsrc = collections.Counter()
tot_u_el = 0; tot_t_el = 0; read_el = 0
for dst, srcs in plan.mix.items():
    if dst[0] == "U":
        b, lp, sp, rp, r = plan.u_list[dst[1]]
    else:
        lp, sp, rp = dst[1]; r = None
    nT = sum(1 for (k, i), cf in srcs if k == "T")
    for (k, i), cf in srcs:
        if k == "T":
            cons[i] += 1
    if dst[0] == "U":
        el = nl[lp] * nr[r]
        nsrc_hist[len(srcs)] += 1; el_by_nsrc[len(srcs)] += el; tot_u_el += el; read_el += el * len(srcs)
print("U targets %d (%.1f MB), sources/target histogram (count, MB):" % (len(plan.u_list), tot_u_el * 8 / 1e6),
      {k: (nsrc_hist[k], round(el_by_nsrc[k] * 8 / 1e6, 1)) for k in sorted(nsrc_hist)})
print("source reads for U targets %.1f MB" % (read_el * 8 / 1e6))
ch = collections.Counter(); el_c = collections.Counter()
for i, (a, lp, l, s, r) in enumerate(plan.t_list):
    ch[cons[i]] += 1; el_c[cons[i]] += nl[lp] * nr[r]; tot_t_el += nl[lp] * nr[r]
print("T blocks %d (%.1f MB), consumers/T histogram (count, MB):" % (len(plan.t_list), tot_t_el * 8 / 1e6), {k: (ch[k], round(el_c[k] * 8 / 1e6, 1)) for k in sorted(ch)})
ydir = [d for d in plan.mix if d[0] == "Y"]
print("direct y targets:", len(ydir), "sources:", sum(len(plan.mix[d]) for d in ydir))
print("flops L %.3f R %.3f" % (plan.flops_L / 1e9, plan.flops_R / 1e9))
# ---- targets sharing one source list (e.g. the two r' = l' +- 1/2 of a doublet s') ----
groups = collections.defaultdict(list)
for dst, srcs in plan.mix.items():
    key = tuple(sorted((k, i if k == "T" else tuple(i)) for (k, i), cf in srcs))
    groups[key].append(dst)
rd_now = 0; rd_grp = 0; wr = 0
gh = collections.Counter()
for key, dsts in groups.items():
    d0 = dsts[0]
    if d0[0] == "U":
        b, lp, sp, rp, r = plan.u_list[d0[1]]; el = nl[lp] * nr[r]
    else:
        lp, sp, rp = d0[1]; el = nl[lp] * nr[rp]
    gh[len(dsts)] += 1
    rd_now += el * len(key) * len(dsts); rd_grp += el * len(key); wr += el * len(dsts)
print("target groups with identical source lists:", dict(gh))
print("mix reads now %.1f MB -> grouped %.1f MB; writes %.1f MB" % (rd_now * 8 / 1e6, rd_grp * 8 / 1e6, wr * 8 / 1e6))
