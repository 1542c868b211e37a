#!/usr/bin/env python
"""Slot efficiency of the grouped-GEMM tile cover on the bench workload (C4 block mix): useful flops /
(DMMA issue slots the four consumer warps of every tile occupy).  CPU only.  usage: python tools/tile_analysis.py"""
import sys, math, collections
sys.path.insert(0,'/root/repo')
import numpy as np
from hubbardtn_b200 import sectors as PS
from oracle.spaces import synthetic_bond_space
from oracle import sectors as S
from oracle.heff import heff_ac_terms
from oracle.tensors import EnvTensor, Legs, MPOTensor, MPSTensor, Space
import importlib.util
spec=importlib.util.spec_from_file_location("syn","/root/repo/hubbardtn_b200/synthetic.py")
# avoid importing package (needs lib): copy functions
src=open('/root/repo/hubbardtn_b200/synthetic.py').read().replace("from . import sectors as S","import hubbardtn_b200_sectors as S")
import types
m=types.ModuleType("hubbardtn_b200_sectors"); exec(open('/root/repo/hubbardtn_b200/sectors.py').read(), m.__dict__); sys.modules["hubbardtn_b200_sectors"]=m
syn=types.ModuleType("syn"); exec(src, syn.__dict__)
sym=0; D=1024; chi=96
phys=m.physical_space(sym,1,1)
levels=syn.mpo_levels(sym,chi)
entries=syn.mpo_entries(sym,levels,phys,4,syn.SEED)
Vl,Vr=Space(sym,syn.bond_space(sym,D,0)),Space(sym,syn.bond_space(sym,D,1))
P,M=Legs(sym,phys),Legs(sym,levels)
GL=EnvTensor("L",Vl,M,identity_levels=[0]); GR=EnvTensor("R",Vr,M,identity_levels=[chi-1])
W=MPOTensor(M,P,M,dict(entries)); x=MPSTensor(Vl,P,Vr)
from oracle.heff import HeffACPlan
plan=HeffACPlan(GL,W,GR,x)
def split_flex(n):
    atoms=(n+7)//8; nt=(atoms+7)//8; base=atoms//nt; rem=atoms%nt; out=[]; o=0
    for i in range(nt):
        ln=min(8*(base+(1 if i<rem else 0)), n-o); out.append((o,ln)); o+=ln
    return out
def tile_cost(fe,fx):
    flex=(fe+7)//8; strips=(fx+15)//16
    return flex*2.0*strips+0.25*flex*2.0*(4-strips)
def tile_block(Mm,N):
    best=None;bc=1e300
    for variant in (0,1):
        v=[];cost=0
        if variant==0:
            nfull=N//64; rn=N%64
            for j in range(nfull):
                for (o,l) in split_flex(Mm): v.append((l,64,0)); cost+=tile_cost(l,64)
            if rn:
                for mo in range(0,Mm,64):
                    mn=min(64,Mm-mo); v.append((mn,rn,1)); cost+=tile_cost(rn,mn)
        else:
            mfull=Mm//64; rm=Mm%64
            for i in range(mfull):
                for (o,l) in split_flex(N): v.append((64,l,1)); cost+=tile_cost(l,64)
            if rm:
                for no in range(0,N,64):
                    nn=min(64,N-no); v.append((rm,nn,0)); cost+=tile_cost(rm,nn)
        if cost<bc: bc=cost;best=v
    return best
def analyze(gemms,name):
    # gemms: list of (M,N,K)
    alg=0; slot=0; used=0; ntiles=0; hist=collections.Counter()
    for (Mm,N,K) in gemms:
        alg+=2*Mm*N*K
        nk4=(K+3)//4
        for (mt,nt,lay) in tile_block(Mm,N):
            ntiles+=1
            flex=( (mt if lay==0 else nt)+7)//8
            fixed= nt if lay==0 else mt
            strips=(fixed+15)//16
            # CTA-time in DMMA slots per warp: flex*2*nk4 (each active warp), 4 warp slots
            slot+=4*flex*2*nk4
            used+= ((mt+7)//8)*((nt+7)//8)*nk4   # useful atoms (padded to 8)
            hist[(flex,strips)]+=flex*2*nk4
    print(name,"gemms",len(gemms),"tiles",ntiles,"alg GF %.2f"%(alg/1e9),"atom-padded GF %.2f"%(used*512/1e9),"warp-slot GF %.2f"%(slot*512/1e9),"slot efficiency %.1f%%"%(100*alg/(slot*512)))
    tot=sum(hist.values())
    print("   time share by (flex atoms, active strips):",sorted(((k,round(100*v/tot,1)) for k,v in hist.items()),key=lambda t:-t[1])[:12])
gL=[(Vl.mult[lp],Vr.mult[r],Vl.mult[l]) for (a,lp,l,s,r) in plan.t_list]
analyze(gL,"stage L")
# stage R: per y block M=n_lp,N=n_rp with K segments; tile by y block; K total
byy=collections.defaultdict(list)
for (b,lp,sp,rp,r) in plan.u_list: byy[(lp,sp,rp)].append(Vr.mult[r])
alg=0;slot=0;hist=collections.Counter();nt_=0
for (lp,sp,rp),Ks in byy.items():
    Mm,N=Vl.mult[lp],Vr.mult[rp]
    for (mt,nt,lay) in tile_block(Mm,N):
        flex=((mt if lay==0 else nt)+7)//8; fixed=nt if lay==0 else mt; strips=(fixed+15)//16
        for K in Ks:
            nk4=(K+3)//4; slot+=4*flex*2*nk4; hist[(flex,strips)]+=flex*2*nk4
    for K in Ks: alg+=2*Mm*N*K
print("stage R alg GF %.2f warp-slot GF %.2f slot efficiency %.1f%%"%(alg/1e9,slot*512/1e9,100*alg/(slot*512)))
tot=sum(hist.values()); print("   ",sorted(((k,round(100*v/tot,1)) for k,v in hist.items()),key=lambda t:-t[1])[:12])

# ---- mix statistics: how many sources feed each U block, how often a T block is used ----
import collections as _c
nsrc = _c.Counter()
t_use = _c.Counter()
elems_by_nsrc = _c.Counter()
for dst, srcs in plan.mix.items():
    if dst[0] != "U":
        continue
    nsrc[len(srcs)] += 1
    b, lp, sp, rp, r = plan.u_list[dst[1]]
    elems_by_nsrc[len(srcs)] += Vl.mult[lp] * Vr.mult[r]
    for (kind, i), cf in srcs:
        if kind == "T":
            t_use[i] += 1
print("U blocks by number of sources:", sorted(nsrc.items()))
tot = sum(elems_by_nsrc.values())
print("U elements by number of sources (%):", sorted((k, round(100 * v / tot, 1)) for k, v in elems_by_nsrc.items()))
print("T blocks by number of U targets:", sorted(_c.Counter(t_use.values()).items()))
