#!/usr/bin/env python
"""bench.py — H_eff applies/s and FP64 TFLOP/s at D=1024 (U(1)xSU(2)), BASELINE.json's metric.

A "step" is ONE application of the one-site effective Hamiltonian y = GL.x.W.GR (MPSKit's
`∂AC`, the inner operation of every Lanczos iteration of VUMPS) on the
synthetic instance of BASELINE config C4 (two-band-like MPO with chi=96 levels, D_red=1024,
28/25 symmetry sectors).  Environments and the MPO are plan-resident in HBM, as they are
across the Krylov iterations of the reference.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N>1 (launched by torchrun, one rank per GPU): the unit-cell sites of the VUMPS sweep are
independent, so rank r applies H_AC of site (r mod 4) with no data-path collective; value is
the aggregate applies/s ("weak" scaling, replicas of the per-site problem).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

UNIT = "applies/s"
METRIC = "H_eff (H_AC) applies/s at D=1024, U(1)xSU(2), chi=96"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--D", type=int, default=1024)
    ap.add_argument("--chi", type=int, default=96)
    ap.add_argument("--sym", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-groundstate", action="store_true", help="skip the time-to-converge legs")
    ap.add_argument("--no-groundstate-scale", action="store_true", help="skip the D_red=256 / 1024 time-to-converge legs")
    ap.add_argument("--shard", default="sector", choices=["sector", "site"],
                    help="N>1: 'sector' (default) = ONE apply sharded over the left symmetry sectors / MPO levels inside "
                         "libhtn (htn_plan_heff_ac_sharded) with an NCCL allreduce of y per apply, strong scaling; "
                         "'site' = independent per-site replicas, no collective, weak scaling")
    return ap.parse_args()


def workload_config(args, extra=None):
    cfg = {
        "workload": "C4-synthetic: one H_AC apply, U1xSU2, D_red=%d, chi=%d MPO levels (4 nnz level pairs/row), "
                    "d=3 multiplets, unit cell 4 (SURVEY.md 8(d))" % (args.D, args.chi),
        "D_red": args.D, "chi": args.chi, "symmetry": "fZ2xSU2xU1" if args.sym == 0 else "fZ2xU1xU1",
        "cache": "inputs larger than L2 (GL+GR+workspaces ~1.2 GB per apply vs 126 MB L2), no flush",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------
# CPU side: the oracle's H_AC on EXACTLY the inputs the GPU arm uploads -- cpu_baseline leg and --impl reference.
# Nothing here touches libhtn.so: the two pure-python input builders of the package are imported without running
# the package's __init__ (which dlopens the library).
# ----------------------------------------------------------------------------------------
def _pure_modules():
    import importlib
    import types
    if "hubbardtn_b200" not in sys.modules:
        pkg = types.ModuleType("hubbardtn_b200")
        pkg.__path__ = [os.path.join(ROOT, "hubbardtn_b200")]
        sys.modules["hubbardtn_b200"] = pkg          # stub: submodules resolve, __init__.py does not run
    return importlib.import_module("hubbardtn_b200.sectors"), importlib.import_module("hubbardtn_b200.synthetic")


def host_case(args, site=0):
    """Oracle tensors holding the data of hubbardtn_b200.synthetic.HeffCase(sym, D, chi, site) (same Philox streams,
    same canonical block order = packed host layout; tests/util.py:oracle_view checks that equality on the GPU)."""
    import math
    import numpy as np
    PS, synthetic = _pure_modules()
    from oracle.heff import HeffACPlan
    from oracle.tensors import EnvTensor, Legs, MPOTensor, MPSTensor, Space

    sym, D, chi = args.sym, args.D, args.chi
    phys = PS.physical_space(sym, 1, 1)
    levels = synthetic.mpo_levels(sym, chi)
    entries = synthetic.mpo_entries(sym, levels, phys, 4, synthetic.SEED + site)
    Vl = Space(sym, synthetic.bond_space(sym, D, site & 1))
    Vr = Space(sym, synthetic.bond_space(sym, D, 1 - (site & 1)))
    P, M = Legs(sym, phys), Legs(sym, levels)

    def fill(t, shapes, stream, scale, identity=None):
        n = int(sum(r * c for r, c in shapes))
        data = synthetic.random_packed(n, stream + 10 * site, synthetic.SEED) * scale
        off = 0
        for k, (r, c) in zip(t.keys, shapes):
            blk = data[off:off + r * c].reshape(r, c)
            t.blocks[k] = np.eye(r) if (identity is not None and k[0] == identity) else np.array(blk)
            off += r * c
        return t

    GL = EnvTensor("L", Vl, M, identity_levels=[0])
    GR = EnvTensor("R", Vr, M, identity_levels=[chi - 1])
    x = MPSTensor(Vl, P, Vr)
    fill(GL, [GL.shape(k) for k in GL.keys], 1, 1.0 / math.sqrt(max(D, 1)), identity=0)
    fill(GR, [GR.shape(k) for k in GR.keys], 2, 1.0 / math.sqrt(max(D, 1)), identity=chi - 1)
    fill(x, [x.blocks[k].shape for k in x.keys], 3, 1.0)
    W = MPOTensor(M, P, M, {k: v for k, v in entries.items()})
    return HeffACPlan(GL, W, GR, x), x


def cpu_apply_timing(args, threads, budget_s, max_reps=40):
    """Times FULL applies of the oracle's staged H_AC on the host cores through oracle/native: the plan's GEMM list as
    single-threaded OpenBLAS dgemm calls, `threads` native worker threads over the sector blocks (the reference's policy:
    BLAS threads = 1, all threads over blocks, HubbardFunctions.jl:29,37).  Returns a dict with flops, seconds per apply,
    repetitions, the checksum of y and -- for the record -- the seconds of ONE apply of the pure-numpy plan, whose block
    loop runs under Python's interpreter lock."""
    import numpy as np
    from threadpoolctl import threadpool_limits
    from oracle.native import NativeHeffAC
    plan, x = host_case(args)
    nat = NativeHeffAC(plan)
    xf = nat.pack_x(x)
    yf = np.empty_like(xf)
    t0 = time.perf_counter()
    nat.apply_flat(xf, yf, threads)                  # warm-up + duration estimate
    one = time.perf_counter() - t0
    n = max(1, min(max_reps, int(budget_s / max(one, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        nat.apply_flat(xf, yf, threads)
    dt = (time.perf_counter() - t0) / n
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        y = plan.apply(x, threads=threads)
        dt_numpy = time.perf_counter() - t0
    chk_numpy = float(sum(np.sum(y.blocks[k]) for k in x.keys))
    chk = float(yf.sum())
    assert abs(chk - chk_numpy) <= 1e-9 * max(abs(chk_numpy), 1e-300), "native and numpy CPU applies disagree"
    return {"flops": plan.flops, "seconds": dt, "reps": n, "checksum": chk, "numpy_plan_seconds": dt_numpy}


def reference_main(args):
    """--impl reference: the reference's CPU path for this metric.  Julia / MPSKit are not in the image, so
    this is the oracle port: the SAME workload (all chi MPO levels, same seeded inputs, same config dict) on all host
    cores through oracle/native (single-threaded OpenBLAS dgemm per block, one native thread per core); each step = one
    full apply."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle.native import NativeHeffAC
    threads = os.cpu_count() or 1
    plan, x = host_case(args)
    nat = NativeHeffAC(plan)
    xf = nat.pack_x(x)
    yf = np.empty_like(xf)
    W = max(1, min(args.warmup, 3))
    K = max(1, min(args.steps, 60))                 # bounded sample: each step is a full apply (~0.15 s)
    for _ in range(W):
        nat.apply_flat(xf, yf, threads)
    t0 = time.perf_counter()
    for _ in range(K):
        nat.apply_flat(xf, yf, threads)
    dt = (time.perf_counter() - t0) / K
    value = 1.0 / dt
    sample = ("oracle HeffACPlan through oracle/native (OpenBLAS dgemm per block with BLAS threads=1, %d native worker threads "
              "over blocks): %d full applies of the same workload (all %d MPO levels, the GPU arm's seeded inputs); %.1f GFLOP/s "
              "= %.2f GFLOP/s per core" % (threads, K, args.chi, plan.flops / dt / 1e9, plan.flops / dt / 1e9 / threads))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * dt, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, {"algorithmic_gflop_per_apply": plan.flops / 1e9}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "checksum": float(yf.sum())},
        "gflops_per_core": plan.flops / dt / 1e9 / threads,
        "note": "reference = CPU restatement (oracle port with native threads + OpenBLAS), NOT MPSKit: Julia is not in the "
                "image (DESIGN.md); baseline/run_reference.jl times the unmodified reference where Julia exists",
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------
# ground state time-to-converge (config C1: one-band Hubbard t=1, U=8, half filling, U1xSU2)
# ----------------------------------------------------------------------------------------
LIEB_WU_U8 = -0.3275305343795398      # exact E/site (tests/golden/reference_energies.json)


def groundstate_leg(ctx, args):
    """The reference's schedule (HubbardFunctions.jl:993-1030: IDMRG2 with truncbelow(10^-svalue) then
    VUMPS) for BASELINE config C1 on the GPU, wall-clock, and -- unless --no-cpu -- the oracle port of
    the same schedule on the host CPU beside it (numpy, single process; NOT MPSKit)."""
    from hubbardtn_b200 import hubbardfunctions as hf
    model = hf.OB_Sim([1.0], [8.0], 0.0, [0.0], 1, 1, 2.0)
    H0 = hf.hamiltonian(model, ctx)
    psi0 = hf.initialize_mps(H0, model.P, model.bond_dim, False, ctx)      # one seeded random start for both arms

    def fresh():
        return hf.InfiniteMPS(ctx, psi0.sym, *[[t.like_copy() for t in lst] for lst in (psi0.AL, psi0.AR, psi0.C, psi0.AC)])

    hf.compute_groundstate(model, ctx=ctx, tol=1e-8, init_state=fresh())   # warm-up (plans, allocations)
    start = fresh()
    ctx.synchronize()
    t0 = time.perf_counter()
    d = hf.compute_groundstate(model, ctx=ctx, tol=1e-8, init_state=start)
    ctx.synchronize()
    gpu_s = time.perf_counter() - t0
    out = {
        "workload": "C1: OB_Sim t=[1] u=[8] P=Q=1 svalue=2.0, U1xSU2; IDMRG2(truncbelow 1e-2, tol 1e-8) -> VUMPS(tol 1e-8)",
        "gpu_seconds": gpu_s, "energy_per_site": d["energy"], "galerkin": d["delta"],
        "idmrg2_iterations": d["idmrg2"]["iterations"], "vumps_iterations": d["vumps"]["iterations"],
        "D_full": hf.dim_state(d["groundstate"]), "energy_above_lieb_wu": d["energy"] - LIEB_WU_U8,
    }
    if not args.no_cpu:
        import numpy as np
        from oracle import mps as M, sectors as OS, twosite as T2
        from oracle.hubbard import OB_Sim as OSim, mpo
        from oracle import bridge
        from oracle.tensors import Space as OSpace
        Ws, P, _ = mpo(OSim(t=[1.0], u=[8.0]))
        # the same initial state as the GPU arm: its left isometries, brought to mixed gauge by the oracle
        L = len(psi0)
        Vd = [psi0.C[i].space(0, psi0.sym) for i in range(L)]
        Vo = [OSpace(OS.SU2U1, dict(zip(v.sectors, v.mult))) for v in Vd]
        tab = lambda t: (t.labels, t.rows, t.cols, t.offsets)  # noqa: E731
        ALo = [bridge.mps_from_packed(Vo[i - 1], P, Vo[i], tab(psi0.AL[i]), psi0.AL[i].download()) for i in range(L)]
        C0 = bridge.bond_from_packed(Vo[L - 1], tab(psi0.C[L - 1]), psi0.C[L - 1].download())
        t0 = time.perf_counter()
        ARo, Co, _ = M.uniform_rightorth(ALo, C0, tol=1e-12)
        st = dict(AL=ALo, AR=ARo, C=Co, AC=[M.mul_right(ALo[i], Co[i]) for i in range(L)])
        AL, C, AR, eps, log = T2.idmrg2(st, Ws, cut=1e-2, tol=1e-8, maxiter=200)
        st2, envs, eps2, log2 = M.vumps(T2.idmrg2_to_uniform(AR, C), Ws, tol=1e-8, maxiter=200)
        out["cpu_port_seconds"] = time.perf_counter() - t0
        out["cpu_port_energy_per_site"] = envs.energy_per_site
        out["cpu_port_note"] = ("oracle restatement (numpy, 1 process, python block loops), not MPSKit/Julia; at this "
                                "size (D_full ~ 30) both sides are latency-bound, not flop-bound")
    return out


def groundstate_scale_leg(ctx, args, D_cap, cut, vumps_tol, vumps_maxiter, idmrg_iters, cpu_iters):
    """BASELINE metric part 2 at a size where the GPU is not launch-bound: config C2 (one-band Hubbard with
    next-nearest-neighbour hopping, t=[1,0.2], U=6, half filling, U1xSU2) grown by IDMRG2 to D_red <= D_cap multiplets per
    bond, then VUMPS to `vumps_tol` -- the reference's schedule HF:993-1030 with a `truncdim`-style cap.  The CPU port
    (oracle VUMPS, numpy, BLAS threads = all cores) runs `cpu_iters` VUMPS iterations from the SAME post-IDMRG2 state
    (downloaded from the device): a bounded sample; its seconds per iteration sit beside the GPU's."""
    import numpy as np
    from hubbardtn_b200 import device as dev, hubbardfunctions as hf
    model = hf.OB_Sim([1.0, 0.2], [6.0], 0.0, [0.0], 1, 1, 5.0)
    H = hf.hamiltonian(model, ctx)
    psi = hf.initialize_mps(H, model.P, model.bond_dim, model.spin, ctx)
    ctx.synchronize()
    t0 = time.perf_counter()
    AL, AR, C, AC, info1 = dev.idmrg2(ctx, psi.AL, psi.AR, psi.C, psi.AC, H.W, cut=cut, tol=1e-6, maxiter=idmrg_iters, maxdim=D_cap)
    ctx.synchronize()
    t1 = time.perf_counter()
    AL, AR, C, AC = dev.uniform_from_right(ctx, AR, C[-1], model.sym)
    ctx.synchronize()
    t2 = time.perf_counter()
    psi = hf.InfiniteMPS(ctx, model.sym, AL, AR, C, AC)
    start_AL = [t.download() for t in psi.AL]
    start_C = psi.C[-1].download()
    GL, GR = hf._make_envs(ctx, psi, H)
    ctx.synchronize()
    t3 = time.perf_counter()
    info2 = dev.vumps(ctx, psi.AL, psi.AR, psi.C, psi.AC, H.W, GL, GR, tol=vumps_tol, maxiter=vumps_maxiter)
    ctx.synchronize()
    t4 = time.perf_counter()
    log = info2["log"]
    out = {
        "workload": "C2: OB_Sim t=[1,0.2] u=[6] P=Q=1, U1xSU2; IDMRG2(truncbelow %.0e, cap %d multiplets, %d sweeps) -> "
                    "InfiniteMPS gauge -> VUMPS(tol %.0e, maxiter %d)" % (cut, D_cap, idmrg_iters, vumps_tol, vumps_maxiter),
        "D_red": [int(sum(c.space(0, model.sym).mult)) for c in psi.C], "D_full": hf.dim_state(psi),
        "mpo_levels": H.chi,
        "gpu_seconds": {"idmrg2": t1 - t0, "gauge": t2 - t1, "vumps": t4 - t3, "total": (t1 - t0) + (t2 - t1) + (t4 - t3)},
        "idmrg2_phase_seconds": dict(zip(["planning", "lanczos", "svd", "env_growth"], [float(v) for v in info1["log"][-1][3:7]])),
        "idmrg2_iterations": info1["iterations"], "idmrg2_applies": int(info1["log"][-1][2]),
        "vumps_iterations": info2["iterations"], "vumps_converged": bool(info2["converged"]), "galerkin": info2["delta"],
        "energy_per_site": info2["energy_per_site"],
        "vumps_phase_seconds": {"eigensolves": float(log[:, 4].sum()), "gauge": float(log[:, 5].sum()), "environments": float(log[:, 6].sum())},
        "vumps_applies": int(log[:, 3].sum()),
        "gpu_seconds_per_vumps_iteration": (t4 - t3) / max(info2["iterations"], 1),
    }
    if not args.no_cpu and cpu_iters > 0:
        from oracle import bridge, mps as M, sectors as OS
        from oracle.hubbard import OB_Sim as OSim, mpo
        from oracle.tensors import Space as OSpace
        Ws, P, _ = mpo(OSim(t=[1.0, 0.2], u=[6.0]))
        L = len(psi)
        Vd = [psi.C[i].space(0, psi.sym) for i in range(L)]
        Vo = [OSpace(OS.SU2U1, dict(zip(v.sectors, v.mult))) for v in Vd]
        tab = lambda t: (t.labels, t.rows, t.cols, t.offsets)  # noqa: E731
        ALo = [bridge.mps_from_packed(Vo[i - 1], P, Vo[i], tab(psi.AL[i]), start_AL[i]) for i in range(L)]
        C0 = bridge.bond_from_packed(Vo[L - 1], tab(psi.C[L - 1]), start_C)
        c0 = time.perf_counter()
        ARo, Co, _ = M.uniform_rightorth(ALo, C0, tol=1e-12)
        st = dict(AL=ALo, AR=ARo, C=Co, AC=[M.mul_right(ALo[i], Co[i]) for i in range(L)])
        c1 = time.perf_counter()
        _, envs, eps_c, log_c = M.vumps(st, Ws, tol=vumps_tol, maxiter=cpu_iters)
        c2 = time.perf_counter()
        n_c = max(len(log_c), 1)
        out["cpu_port"] = {
            "kind": "port", "cores": os.cpu_count(), "vumps_iterations_timed": n_c, "seconds": c2 - c1,
            "seconds_per_vumps_iteration": (c2 - c1) / n_c, "gauge_seconds": c1 - c0,
            "galerkin_after": eps_c, "energy_per_site_after": envs.energy_per_site,
            "gpu_galerkin_after_same_iterations": float(log[min(n_c, len(log)) - 1, 0]),
            "gpu_energy_after_same_iterations": float(log[min(n_c, len(log)) - 1, 1]),
            "estimated_seconds_to_converge": (c2 - c1) / n_c * info2["iterations"],
            "note": "oracle restatement (numpy/OpenBLAS with all cores, python block loops), NOT MPSKit/Julia; bounded sample: "
                    "%d VUMPS iterations from the GPU arm's post-IDMRG2 state; the estimate assumes the GPU arm's iteration "
                    "count" % n_c,
        }
        out["vumps_iteration_speedup_vs_cpu_port"] = out["cpu_port"]["seconds_per_vumps_iteration"] / out["gpu_seconds_per_vumps_iteration"]
    return out


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.rows, self.proc = dev, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, in_timed = [], [], set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lo, hi = getattr(self, "window", (0.0, 1e30))
        t_lo, t_hi = getattr(self, "timed", (0.0, 0.0))
        for r in self.rows:
            try:
                if not (lo <= r[0] <= hi):
                    continue
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                in_timed += 1 if t_lo <= r[0] <= t_hi else 0
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": in_timed,
                "note": "nvidia-smi sampled every 20 ms while the H_AC apply loop runs: the timed region plus an untimed "
                        "continuation of the identical loop (the timed region alone is shorter than a few samples)"}


# ----------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        rc = reference_main(args)
        assert "hubbardtn_b200._lib" not in sys.modules, "the reference arm must not load libhtn.so"
        return rc

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        import faulthandler
        faulthandler.dump_traceback_later(240, exit=True)      # a rank stuck in a collective: dump where, and leave
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))

    from hubbardtn_b200 import device, synthetic
    from hubbardtn_b200 import sharding
    ctx = device.Context(local)
    strong = world > 1 and args.shard == "sector"
    # strong scaling: every rank holds the SAME site-0 problem and owns a shard of its terms (left symmetry sectors,
    # heavy ones split by MPO level: htn_plan_heff_ac_sharded); y = NCCL allreduce(sum) of the partial applies,
    # enqueued on the library's stream right behind the apply (no host synchronisation in between).
    # replicas ('--shard site'): rank r applies H_AC of unit-cell site r mod 4, no collective.
    case = synthetic.HeffCase(ctx, args.sym, D=args.D, chi=args.chi, site=0 if (strong or world == 1) else sharding.site_for_rank(rank, 4))
    full_stats = case.plan.stats
    if strong:
        case.plan = device.HeffAC(ctx, case.GL, case.W, case.GR, case.x, nshards=world, shard=rank)
    plan, x, y = case.plan, case.x, case.y
    st = plan.stats
    lib_stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", local))
    y_t = torch.as_tensor(y.device_array(), device="cuda")

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(k):
        """k steps on the library stream; returns device ms (CUDA events on that stream)."""
        if not strong:
            return plan.time(x, y, k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(lib_stream):
            e0.record(lib_stream)
            for _ in range(k):
                plan.apply(x, y)                              # library stream
                dist.all_reduce(y_t, op=dist.ReduceOp.SUM)    # ordered behind it (torch: current stream = library stream)
            e1.record(lib_stream)
        e1.synchronize()
        return e0.elapsed_time(e1)

    # ---- warm-up (>=3) -------------------------------------------------------------------
    W = max(args.warmup, 3)
    run_steps(W)
    # ---- FP64 peak of this GPU (roofline denominator; not in MEASURED_PEAKS.json; profiles/fp64_peak.json) ----
    dmma_peak = ctx.probe_fp64_peak(0)
    dfma_peak = ctx.probe_fp64_peak(1)
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    cublas_peak = best
    del a, b
    peak = max(dmma_peak, dfma_peak, cublas_peak)

    # ---- timed region: exactly K steps, device time on the launching stream, max over ranks -----------------
    sampler = ClockSampler(local)
    sampler.start()
    run_steps(max(W, 200))          # sampler start-up happens under load, outside the timed region
    barrier()
    t_lo = time.time()
    ms = run_steps(args.steps)
    barrier()
    t_hi = time.time()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    # ~1 s more of the same loop for the clock samples; the count must be the SAME on every rank (the loop holds one
    # collective per step), so it is derived from the max-over-ranks time, not from the local one
    reps_more = int(1.0 / max(ms_max / args.steps * 1e-3, 1e-6))
    run_steps(max(1, reps_more))
    t_end = time.time()
    sampler.window, sampler.timed = (t_lo, t_end), (t_lo, t_hi)
    clocks = sampler.stop()
    value = (1 if strong else world) * args.steps / (ms_max * 1e-3)
    checksum_dev = float(y_t.sum().item())                     # of the (reduced) y left by the timed loop

    # ---- per-stage times of this rank's plan ------------------------------------------------------------
    prof = plan.profile(x, y, reps=max(5, min(args.steps, 100)))
    gemm_ms = prof["stage_L_ms"] + prof["stage_R_ms"]
    achieved = st["flops"] / (gemm_ms * 1e-3) / 1e12
    shard_ms = plan.time(x, y, min(args.steps, 100)) / min(args.steps, 100)
    per_rank = torch.tensor([shard_ms, st["flops"]], dtype=torch.float64, device="cuda")
    gathered = [torch.zeros_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, per_rank)
    else:
        gathered = [per_rank]
    shard_table = [[float(g[0]), float(g[1]) / 1e9] for g in gathered]

    # ---- e2e: host buffers through the C ABI (pinned), H2D + D2H inside the timed region ----
    xh = torch.from_numpy(case.x_host.copy()).pin_memory()
    yh = torch.empty_like(xh).pin_memory()
    yh_t = None
    n = x.nelem
    if strong:
        # sharded e2e: every rank uploads x, applies its shard, the partial y is allreduced on the device and
        # downloaded (the call sequence a Julia host would issue per Krylov step)
        def e2e_step():
            x.upload_ptr(xh.data_ptr(), n)
            with torch.cuda.stream(lib_stream):
                plan.apply(x, y)
                dist.all_reduce(y_t, op=dist.ReduceOp.SUM)
            y.download_ptr(yh.data_ptr(), n)                  # library stream: ordered behind the allreduce
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        e2e_s = time.perf_counter() - t0
        checksum = float(yh.numpy().sum())
    else:
        for _ in range(3):
            plan.apply_host_ptr(xh.data_ptr(), yh.data_ptr(), n)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            plan.apply_host_ptr(xh.data_ptr(), yh.data_ptr(), n)   # returns after the D2H copy completed
        e2e_s = time.perf_counter() - t0
        checksum = float(yh.numpy().sum())
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = (1 if strong else world) * args.steps / float(te.item())

    # ---- ncu-measured DRAM traffic of the dominant kernels (committed capture; per apply) --------
    traffic = None
    traffic_note = None
    traffic_by_kernel = None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_roofline_inputs.json")) as f:
            ri = json.load(f)
        if args.D == 1024 and args.chi == 96 and args.sym == 0:
            traffic_by_kernel = ri["gemm_traffic_bytes_per_launch"]
            traffic = traffic_by_kernel["stack_gemm_kernel"]      # the dominant launch (stage L + W), bytes per launch
            traffic_note = ri.get("note")
    except Exception:
        traffic = None

    # ---- Krylov vector kernels against the HBM roofline (north star: "HBM GB/s for the Krylov kernels") ----
    kry = None
    if rank == 0:
        try:
            kry = device.probe_krylov(x, nvec=30, reps=20)
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs") if os.path.exists(
                os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
            kry["hbm_peak_GBs"] = hbm
            kry["note"] = ("30 basis vectors of len(AC) = %.1f MB: the whole basis (%.0f MB) fits the 126 MB L2, so rates above "
                           "the HBM peak are L2 hits" % (kry["vector_bytes"] / 1e6, 31 * kry["vector_bytes"] / 1e6))
        except Exception as e:          # keep the headline line alive
            kry = {"error": str(e)}

    # ---- ground-state time-to-converge (BASELINE metric part 2), rank 0 of a single-GPU run only ----------
    gs = None
    if rank == 0 and world == 1 and not args.no_groundstate:
        gs = groundstate_leg(ctx, args)
        if not args.no_groundstate_scale:
            # C2 at D_red <= 256 to a Galerkin error of 1e-10, CPU port beside it for 2 VUMPS iterations
            gs["C2_D256"] = groundstate_scale_leg(ctx, args, 256, 1e-9, 1e-10, 300, 30, 2)
            # the same model at D_red <= 1024: per-phase seconds where the contractions dominate (GPU only: one CPU-port
            # iteration at this size takes minutes)
            gs["C2_D1024"] = groundstate_scale_leg(ctx, args, 1024, 1e-11, 1e-10, 20, 10, 0)

    line = None
    if rank == 0:
        par = ("one GPU" if world == 1 else
               ("ONE apply sharded over the left symmetry sectors (heavy sectors split by MPO level) inside libhtn "
                "(htn_plan_heff_ac_sharded): GL/T/U rows and stage L are partitioned, NCCL allreduce(sum) of y (%d B) per "
                "apply on the library stream" % (y_t.numel() * 8) if strong else
                "site-parallel replicas (rank r -> unit-cell site r mod 4), no collective"))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if (strong or world == 1) else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, {"parallelism": par, "x_elems": n,
                                             "algorithmic_gflop_per_apply": full_stats["flops"] / 1e9,
                                             "executed_gflop_per_apply_padded": full_stats["padded_flops"] / 1e9,
                                             "workspace_MB": st["workspace_bytes"] / 1e6}),
            "tflops_fp64": value * full_stats["flops"] / 1e12,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
                    "checksum": checksum},
            "gpu_launches": int(st["launches_per_apply"]) * args.steps,
            "stages_ms": prof,
            "roofline": {
                "bound": "tensor", "kernel": "stack_gemm_kernel (stage L, the stage-W mix rides along) + grouped_gemm_kernel (stage R)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": traffic_note, "traffic_by_kernel": traffic_by_kernel,
                "launches": [
                    {"kernel": "stack_gemm_kernel (stage L: T = GL.x, the stage-W recoupling mix fused in)",
                     "algorithmic_flop": st["flops_L"], "ms": prof["stage_L_ms"],
                     "tflops": st["flops_L"] / (prof["stage_L_ms"] * 1e-3) / 1e12,
                     "frac": st["flops_L"] / (prof["stage_L_ms"] * 1e-3) / 1e12 / peak},
                    {"kernel": "grouped_gemm_kernel (stage R: y = U.GR, split-K)",
                     "algorithmic_flop": st["flops_R"], "ms": prof["stage_R_ms"],
                     "tflops": st["flops_R"] / (prof["stage_R_ms"] * 1e-3) / 1e12,
                     "frac": st["flops_R"] / (prof["stage_R_ms"] * 1e-3) / 1e12 / peak}],
                "peak_source": "peak measured by this repo's probe, live on this GPU: max(DMMA.8x8x4 issue loop %.2f, DFMA loop "
                               "%.2f, cuBLAS DGEMM 8192^3 %.2f) TFLOP/s; MEASURED_PEAKS.json has no FP64 entry "
                               "(profiles/fp64_peak.json holds the committed probe run)" % (dmma_peak, dfma_peak, cublas_peak),
                "whole_apply_frac": value * full_stats["flops"] / 1e12 / peak / (world if strong else 1) if strong
                else (value / world) * st["flops"] / 1e12 / peak,
            },
        }
        if world > 1:
            line["shards"] = {"per_rank_apply_ms_and_gflop": shard_table,
                              "collective": "NCCL allreduce(sum, f64) of y, %d bytes, stream-ordered behind the apply" % (y_t.numel() * 8)
                              if strong else "none",
                              "checksum_of_reduced_y": checksum_dev}
    # ---- CPU baseline beside it (rank 0, N=1 only; bounded sample) + checksum pin ----------------
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        ct = cpu_apply_timing(args, threads, args.cpu_seconds)
        flops_c, dt, nrep, cpu_checksum = ct["flops"], ct["seconds"], ct["reps"], ct["checksum"]
        line["cpu_baseline"] = {
            "value": 1.0 / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "oracle HeffACPlan through oracle/native (OpenBLAS dgemm per block with BLAS threads=1, %d native worker "
                      "threads over blocks): %d full applies of the same workload (all %d MPO levels, same seeded inputs); %.1f "
                      "GFLOP/s = %.2f GFLOP/s per core" % (threads, nrep, args.chi, flops_c / dt / 1e9, flops_c / dt / 1e9 / threads),
            "gflops_per_core": flops_c / dt / 1e9 / threads,
            "numpy_plan_applies_per_s": 1.0 / ct["numpy_plan_seconds"],
            "numpy_plan_note": "the same plan as pure numpy (python block loop under the interpreter lock, BLAS threads=1, "
                               "%d python threads): the round-1 baseline, kept for comparison" % threads}
        # the e2e result of the GPU arm is pinned on the oracle's value for the same inputs
        rel = abs(checksum - cpu_checksum) / max(abs(cpu_checksum), 1e-300)
        line["e2e"]["oracle_checksum"] = cpu_checksum
        line["e2e"]["checksum_rel_err"] = rel
        assert rel < 1e-9, "GPU e2e checksum %r differs from the oracle's %r" % (checksum, cpu_checksum)
    if rank == 0:
        if gs is not None:
            line["groundstate"] = gs
        line["krylov"] = kry
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        import faulthandler
        faulthandler.cancel_dump_traceback_later()
    return 0


if __name__ == "__main__":
    sys.exit(main())
