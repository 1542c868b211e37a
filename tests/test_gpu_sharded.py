"""MPO-level sharding of one H_AC apply on the device (SURVEY.md 8(e) axis 2): the per-rank plans built from
`sharding.shard_mpo_entries` sum to the unsharded apply.  Runs on ONE GPU (the shards are evaluated one after
the other and added with the library's axpby); with N GPUs bench.py --shard mpo replaces the sum by an NCCL
allreduce of the same buffers (checksum equality is reported there)."""
import numpy as np
import pytest

from hubbardtn_b200 import device as dev, sectors as PS, sharding, synthetic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,world", [("rows", 2), ("rows", 4), ("chains", 2)])
def test_sharded_applies_sum_to_full_apply(ctx, mode, world):
    case = synthetic.HeffCase(ctx, PS.SU2U1, D=96, chi=14)
    case.plan.apply(case.x, case.y)
    full = case.y.download()
    acc = case.y.like()
    part = case.y.like()
    flops = 0.0
    owned = 0
    for rank in range(world):
        mine = sharding.shard_mpo_entries(case.w_entries, case.chi, world, rank, mode=mode)
        owned += len(mine)
        Wr = dev.Mpo(ctx, case.M, case.P, case.M, mine)
        plan = dev.HeffAC(ctx, case.GL, Wr, case.GR, case.x)
        plan.apply(case.x, part)
        acc.axpby(1.0, part, 1.0 if rank else 0.0)
        flops += plan.stats["flops"]
    assert owned == len(case.w_entries)                       # a partition: every entry exactly once
    got = acc.download()
    assert np.abs(got - full).max() < 1e-12 * np.abs(full).max()
    assert flops >= case.plan.stats["flops"]                  # duplication of shared stages is visible, never hidden
    arr = case.y.device_array().__cuda_array_interface__
    assert arr["typestr"] == "<f8" and arr["shape"][0] >= case.y.nelem


@pytest.mark.parametrize("world", [2, 4, 8])
def test_library_sector_shards_sum_to_full_apply(ctx, world):
    """htn_plan_heff_ac_sharded (sector / MPO-level units decided inside the library): the partial applies of all
    shards sum to the unsharded apply, every stage-L block is computed exactly once, and the shards are balanced."""
    case = synthetic.HeffCase(ctx, PS.SU2U1, D=256, chi=20)
    case.plan.apply(case.x, case.y)
    full = case.y.download()
    acc = case.y.like()
    part = case.y.like()
    flops, n_l = [], 0
    for rank in range(world):
        plan = dev.HeffAC(ctx, case.GL, case.W, case.GR, case.x, nshards=world, shard=rank)
        plan.apply(case.x, part)
        acc.axpby(1.0, part, 1.0 if rank else 0.0)
        flops.append(plan.stats["flops"])
        n_l += plan.stats["n_gemm_L"]
    got = acc.download()
    assert np.abs(got - full).max() < 1e-12 * np.abs(full).max()
    assert n_l == case.plan.stats["n_gemm_L"]                 # stage L is partitioned, never duplicated
    total = case.plan.stats["flops"]
    assert total <= sum(flops) <= 1.6 * total                 # stage R of split heavy sectors runs on each owner
    assert max(flops) <= 2.2 * sum(flops) / world             # no shard carries much more than a fair share
