"""Batch sharding of the loss path: 2 gloo ranks on the CPU, kernels stepped by tests/emu.

Each rank evaluates its slice of a mixed-baseline tri-min batch; the mean of the rank losses must
equal the single-process loss on the whole batch, and every rank's disparity gradients (scaled by
1/world) must equal the corresponding rows of the single-process gradients.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _slice_batch(inputs, outputs, noise, plan, rank, world):
    from baseboostdepth_b200.plan import build_plan, shard_ordering
    B = len(inputs["ordering"])
    per = B // world
    lo, hi = rank * per, (rank + 1) * per
    local_order, rows = shard_ordering(inputs["ordering"], rank, world)
    li = {"ordering": local_order}
    for k, v in inputs.items():
        if not torch.is_tensor(v):
            continue
        if isinstance(k, tuple) and k[0] == "color" and k[1] != 0:
            li[k] = v[rows[k[1]]]
        else:
            li[k] = v[lo:hi]
    lplan = build_plan(local_order, trimin=plan.trimin, decomp=plan.decomp, groups=[g for g in plan.groups])
    lo_out = {}
    for k, v in outputs.items():
        if k[0] == "disp":
            lo_out[k] = v[lo:hi].detach().clone().requires_grad_(True)
        elif k[0] in ("cam_T_cam", "cam_T_cam_error"):
            f = k[2]
            keep = [i for i, b in enumerate(plan.sel.get(f, [])) if lo <= b < hi]
            lo_out[k] = v[keep].detach().clone()
    lnoise = {}
    for g in lplan.groups:
        keep = [i for i, b in enumerate(plan.group_members[g]) if lo <= b < hi]
        lnoise[g] = noise[g][keep]
    return li, lo_out, lnoise, lplan


def _worker(rank, world, port, result):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from baseboostdepth_b200.plan import build_plan
    from baseboostdepth_b200.synthetic import make_batch, make_noise
    from baseboostdepth_b200.trainer import loss_step
    from fused_util import emu_backend
    from oracle.loss_path import default_opt

    cfg = dict(batch=4, height=32, width=64, baselines=[2, 1, "s", 2], trimin=True, decomp=False)
    opt = default_opt(height=32, width=64, trimin=True, batch_size=4)
    inputs, outputs, params = make_batch(seed=5, **cfg)
    plan = build_plan(inputs["ordering"], trimin=True)
    noise = make_noise(plan, 32, 64, seed=6)
    li, lo_out, lnoise, lplan = _slice_batch(inputs, outputs, noise, plan, rank, world)
    losses = loss_step(li, lo_out, opt, lplan, noise=lnoise, num_scales=4, backend=emu_backend())
    (losses["loss"] / world).backward()
    t = losses["loss"].detach().clone()
    dist.all_reduce(t)                              # optional scalar all-reduce for logging
    mean_loss = float(t) / world

    if rank == 0:
        full = loss_step(inputs, outputs, opt, plan, noise=noise, num_scales=4, backend=emu_backend())
        full["loss"].backward()
        per = 4 // world
        gerr = max(float((lo_out[("disp", s)].grad - params[("disp", s)].grad[:per]).abs().max()) for s in range(4))
        result.put((mean_loss, float(full["loss"]), gerr))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_equals_single_process():
    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, result)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    mean_loss, full_loss, gerr = result.get(timeout=10)
    assert abs(mean_loss - full_loss) <= 1e-6
    assert gerr <= 1e-9


def test_shard_rows_partition_the_stacks():
    from baseboostdepth_b200.plan import shard_ordering
    ordering = [[0, 3, -3], [0, "s"], [0, 1, -1], [0, 2, -2], [0, 3, -3], [0, 2, -2]]
    seen = {}
    for r in range(3):
        local, rows = shard_ordering(ordering, r, 3)
        assert local == ordering[2 * r:2 * r + 2]
        for f, idx in rows.items():
            seen.setdefault(f, []).extend(idx)
    assert seen[1] == list(range(5)) and seen[-3] == [0, 1] and seen["s"] == [0, 1, 2, 3]
