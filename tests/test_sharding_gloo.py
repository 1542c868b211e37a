"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank evaluates the H_AC partial sum over ITS MPO
level pairs with the oracle; allreduce(sum) must equal the unsharded apply (the identity the MPO-level
sharding of SURVEY.md 8(e) rests on), and the site / rate bookkeeping of bench.py must aggregate."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hubbardtn_b200 import sharding
    from oracle import sectors as S
    from oracle.heff import heff_ac_apply_naive
    from oracle.spaces import physical_space, synthetic_bond_space
    from oracle.tensors import EnvTensor, Legs, MPOTensor, MPSTensor
    kind = S.SU2U1
    rng = np.random.default_rng(7)                       # same inputs on every rank
    levels = [(0, 0, 0), (1, 1, 1), (1, 1, 1), (1, 1, -1), (0, 2, 0), (0, 0, 2), (0, 0, 0)]
    P = physical_space(kind, 1, 1)
    Va, Vb = synthetic_bond_space(kind, 16, 0), synthetic_bond_space(kind, 16, 1)
    Mleg = Legs(kind, levels)
    GL = EnvTensor("L", Va, Mleg, identity_levels=[0]).randomize(rng)
    GR = EnvTensor("R", Vb, Mleg, identity_levels=[6]).randomize(rng)
    W = MPOTensor(Mleg, P, Mleg).randomize(rng, pattern={(0, 0), (0, 1), (1, 2), (2, 6), (0, 3), (3, 6), (0, 4), (4, 6),
                                                         (0, 5), (5, 6), (0, 6), (6, 6)})
    x = MPSTensor(Va, P, Vb).randomize(rng)
    pairs = sorted({(k[0], k[3]) for k in W.entries})
    owner = sharding.partition_levels(pairs, len(levels), world)
    assert owner[(0, 1)] == owner[(1, 2)] == owner[(2, 6)]          # a chain stays on one rank
    mine = MPOTensor(Mleg, P, Mleg, {k: v for k, v in W.entries.items() if owner[(k[0], k[3])] == rank})
    y = heff_ac_apply_naive(GL, mine, GR, x)
    flat = torch.from_numpy(np.concatenate([y.blocks[k].ravel() for k in x.keys]))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    full = heff_ac_apply_naive(GL, W, GR, x)
    ref = np.concatenate([full.blocks[k].ravel() for k in x.keys])
    err = float(np.abs(flat.numpy() - ref).max() / np.abs(ref).max())
    # rate bookkeeping: gather per-rank times, aggregate as bench.py does
    ms = torch.tensor([10.0 + rank])
    gathered = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(gathered, ms)
    rate = sharding.aggregate_rate(world, 50, [float(g) for g in gathered])
    if rank == 0:
        out.put((err, rate, sharding.site_for_rank(5, 4), sorted(set(owner.values()))))
    dist.barrier()
    dist.destroy_process_group()


def test_mpo_level_sharding_sums_to_full_apply():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, rate, site, owners = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-13
    assert abs(rate - 2 * 50 / 11e-3) < 1e-6
    assert site == 1 and owners == [0, 1]
