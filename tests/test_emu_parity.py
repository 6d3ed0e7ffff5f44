"""Kernel source vs oracle on the CPU.

tests/emu steps the *same* phase functions the sm_100a kernels run (one thread block at a
time, phase by phase), so tile/halo/reflection indexing, the candidate tables, the
per-pixel arithmetic and the analytic gradients are checked here without a GPU.  The
real kernels are checked against the oracle in test_gpu_parity.py.

Tolerances: loss 2e-6 relative; gradients 1e-5 relative L2 (north-star bar).
"""
import pytest
import torch

from baseboostdepth_b200.trainer import materialise_warps
from fused_util import emu_backend, run_fused
from helpers import Golden, assert_grad_parity, golden_cases, max_abs, rel_l2
from oracle import loss_path as O

CASES = golden_cases()


@pytest.mark.parametrize("case", CASES)
def test_fused_matches_oracle(case):
    g = Golden(case)
    ref, aux = O.run(g.inputs, g.outputs, g.opt(), g.noise, num_scales=g.num_scales)
    ref["loss"].backward()
    ref_grads = {k: v.grad.clone() for k, v in g.params.items() if v.grad is not None}

    h = Golden(case)
    losses, plan = run_fused(h.inputs, h.outputs, h.opt(), h.noise, h.num_scales, backend=emu_backend(),
                             groups=aux["groups"])
    for k, v in ref.items():
        assert abs(float(losses[k]) - float(v)) <= 2e-6 * max(1.0, abs(float(v))), k
        assert abs(float(losses[k]) - g.losses[k]) <= 2e-6, k   # and the reference's own number
    losses["loss"].backward()
    ref64 = None
    for k, gr in ref_grads.items():
        if rel_l2(h.params[k].grad, gr) > 1e-5 and ref64 is None:
            # a few hundred pixels with one ill-conditioned SSIM window: fp32 rounding noise of the
            # reference itself exceeds the bar -- fall back to the float64 yardstick (helpers.py)
            d = Golden(case, dtype=torch.float64)
            l64, _ = O.run(d.inputs, d.outputs, d.opt(), d.noise, num_scales=d.num_scales)
            l64["loss"].backward()
            ref64 = {kk: v.grad for kk, v in d.params.items() if v.grad is not None}
        assert_grad_parity(h.params[k].grad, gr, None if ref64 is None else ref64[k], k)

    # depth planes and the per-pixel selection
    for i, s in enumerate(h.scales):
        assert max_abs(h.outputs[("depth", 0, s)], g.outputs[("depth", 0, s)]) <= 2e-6 * 100
    win = h.outputs["argmin"]
    for i, s in enumerate(h.scales):
        vals = torch.cat(aux["to_optimise"][s], 0)
        args = torch.cat(aux["argmin"][s], 0)
        order = [b for grp in aux["groups"] for b in plan.group_members[grp]]
        mine = win[i][order].long()
        # selections must agree wherever the oracle's margin between best and runner-up is > 1e-6
        agree = (mine == args)
        if not bool(agree.all()):
            margin = _margin(g, aux, s)
            assert bool((agree | (margin <= 1e-6)).all()), (case, s, int((~agree).sum()))
        del vals


def _margin(g, aux, s):
    """best-vs-second-best gap of the oracle's candidate planes (recomputed, slow but small)."""
    # run the oracle once more keeping the concatenated candidate planes via torch.topk on -loss
    import torch.nn.functional as F  # noqa
    planes = aux.get("planes", {}).get(s)
    if planes is None:
        return torch.zeros_like(torch.cat(aux["to_optimise"][s], 0))
    out = []
    for p in planes:
        top = torch.topk(-p, 2, dim=1).values
        out.append(top[:, 0] - top[:, 1])
    return torch.cat(out, 0)


@pytest.mark.parametrize("case", ["plain_pm1", "trimin_decomp"])
def test_warps_match_reference(case):
    g = Golden(case)
    opt = g.opt()
    ref, aux = O.run(g.inputs, g.outputs, opt, g.noise, num_scales=g.num_scales)
    h = Golden(case)
    with torch.no_grad():
        losses, plan = run_fused(h.inputs, h.outputs, h.opt(), h.noise, h.num_scales, backend=emu_backend(),
                                 groups=aux["groups"])
        materialise_warps(h.inputs, h.outputs, opt, plan, backend=emu_backend())
    s0 = h.scales[0]
    for f in plan.frames:
        assert max_abs(h.outputs[("color", f, s0)], g.ref_out[("color", f, s0)]) <= 2e-5, f
        if g.decomp and f != "s":
            assert max_abs(h.outputs[("color_D", f, s0)], g.ref_out[("color_D", f, s0)]) <= 2e-5, f


def test_forward_only_mode():
    g = Golden("plain_pm1")
    with torch.no_grad():
        losses, _ = run_fused(g.inputs, g.outputs, g.opt(), g.noise, g.num_scales, backend=emu_backend())
    assert abs(float(losses["loss"]) - g.losses["loss"]) <= 2e-6
    assert not losses["loss"].requires_grad


def test_sql_mode_emulated():
    """opt.SQL (disp is depth, trainer.py:457-458) through the kernels' phase code on the CPU."""
    from baseboostdepth_b200.synthetic import make_batch, make_noise
    from baseboostdepth_b200.trainer import plan_for
    cfg = dict(batch=2, height=32, width=64, baselines=[1, 1], trimin=False, decomp=False)
    opt = O.default_opt(height=32, width=64, scales=[0], SQL=True, batch_size=2)
    res = []
    for mine in (False, True):
        inputs, outputs, params = make_batch(seed=12, device="cpu", scales=(0,), **cfg)
        with torch.no_grad():
            params[("disp", 0)].mul_(20.0).add_(1.0)
        plan = plan_for(inputs["ordering"], False, False, None)
        noise = make_noise(plan, 32, 64, seed=3)
        if mine:
            losses, _ = run_fused(inputs, outputs, opt, noise, 4, backend=emu_backend())
        else:
            losses, _ = O.run(inputs, outputs, opt, noise, num_scales=4)
        losses["loss"].backward()
        res.append((float(losses["loss"]), params[("disp", 0)].grad.clone()))
    assert abs(res[0][0] - res[1][0]) <= 2e-6
    assert rel_l2(res[1][1], res[0][1]) <= 1e-5


def test_step_leaves_no_reference_cycle():
    """The autograd graph of a step must die with its loss (refcount, no GC): a cycle through the fused
    node keeps AccumulateGrad nodes alive across steps, which breaks CUDA-graph capture later."""
    import gc
    import weakref
    g = Golden("plain_pm1")
    gc.collect()
    gc.disable()
    try:
        losses, _ = run_fused(g.inputs, g.outputs, g.opt(), g.noise, g.num_scales, backend=emu_backend())
        node = losses["loss"].grad_fn
        nodes = [node] + [fn for fn, _ in node.next_functions if fn is not None]
        refs = [weakref.ref(n) for n in nodes]
        del losses, node, nodes
        assert all(r() is None for r in refs)
    finally:
        gc.enable()


@pytest.mark.parametrize("name,baselines,trimin,decomp", [
    ("all_stereo_plain", ["s", "s"], False, False),
    ("all_stereo_trimin", ["s", "s", "s"], True, False),
    ("single_sample_stereo", ["s"], False, False),
    ("single_sample_b3_decomp", [3], True, True),
    ("one_of_each_trimin", ["s", 1, 2, 3], True, False),
])
def test_degenerate_batch_layouts(name, baselines, trimin, decomp):
    """Batch layouts at the edges of the candidate tables (no temporal source at all, a single sample, every
    baseline once): kernel source stepped on the CPU vs the oracle, losses / gradients / selections."""
    from baseboostdepth_b200.synthetic import make_batch, make_noise
    from baseboostdepth_b200.trainer import plan_for
    H, W, B = 32, 56, len(baselines)
    cfg = dict(batch=B, height=H, width=W, baselines=baselines, trimin=trimin, decomp=decomp, scales=(0, 1), seed=31)
    opt = O.default_opt(height=H, width=W, scales=[0, 1], trimin=trimin, decomp=decomp, pose_error=5.5, batch_size=B)

    gi, go, gp = make_batch(device="cpu", pose_error=5.5, **cfg)
    s_rows = gi[("color", "s", 0)].shape[0] if ("color", "s", 0) in gi else None
    plan = plan_for(gi["ordering"], trimin, decomp, s_rows)
    noise = make_noise(plan, H, W, seed=5)
    ref, aux = O.run(gi, go, opt, noise, num_scales=4)
    ref["loss"].backward()
    ref_grads = {k: v.grad.clone() for k, v in gp.items() if v.grad is not None}

    hi, ho, hp = make_batch(device="cpu", pose_error=5.5, **cfg)
    losses, plan2 = run_fused(hi, ho, opt, noise, 4, backend=emu_backend(), groups=aux["groups"])
    for k, v in ref.items():
        assert abs(float(losses[k].detach()) - float(v.detach())) <= 2e-6 * max(1.0, abs(float(v.detach()))), (name, k)
    losses["loss"].backward()
    for k, gr in ref_grads.items():
        assert hp[k].grad is not None, (name, k)
        assert rel_l2(hp[k].grad, gr) <= 2e-5, (name, k, rel_l2(hp[k].grad, gr))
    win = ho["argmin"]
    order = [b for grp in aux["groups"] for b in plan2.group_members[grp]]
    for i, s in enumerate([0, 1]):
        args = torch.cat(aux["argmin"][s], 0)
        mine = win[i][order].long()
        agree = mine == args
        if not bool(agree.all()):
            margins = []
            for p in aux["planes"][s]:
                top = torch.topk(-p, 2, dim=1).values
                margins.append(top[:, 0] - top[:, 1])
            assert bool((agree | (torch.cat(margins, 0) <= 1e-6)).all()), (name, s)


@pytest.mark.parametrize("H,W", [(16, 32), (24, 56), (48, 64), (40, 120), (16, 200)])
def test_fused_size_sweep(H, W):
    """Frame sizes around the tile geometry (one tile, exact multiples of 28 / 16, wide and flat, ragged both
    ways), tri-min with a mixed batch: kernel source stepped on the CPU vs the oracle."""
    from baseboostdepth_b200.synthetic import make_batch, make_noise
    from baseboostdepth_b200.trainer import plan_for
    baselines = [2, "s", 1]
    cfg = dict(batch=3, height=H, width=W, baselines=baselines, trimin=True, decomp=False, scales=(0, 1, 2, 3), seed=41)
    opt = O.default_opt(height=H, width=W, scales=[0, 1, 2, 3], trimin=True, decomp=False, batch_size=3)
    gi, go, gp = make_batch(device="cpu", **cfg)
    plan = plan_for(gi["ordering"], True, False, gi[("color", "s", 0)].shape[0])
    noise = make_noise(plan, H, W, seed=6)
    ref, aux = O.run(gi, go, opt, noise, num_scales=4)
    ref["loss"].backward()
    hi, ho, hp = make_batch(device="cpu", **cfg)
    losses, _ = run_fused(hi, ho, opt, noise, 4, backend=emu_backend(), groups=aux["groups"])
    for k, v in ref.items():
        assert abs(float(losses[k].detach()) - float(v.detach())) <= 2e-6 * max(1.0, abs(float(v.detach()))), k
    losses["loss"].backward()
    for k, v in gp.items():
        if v.grad is not None:
            assert rel_l2(hp[k].grad, v.grad) <= 2e-5, (k, rel_l2(hp[k].grad, v.grad))


@pytest.mark.parametrize("case", ["plain_pm1", "trimin_mixed"])
def test_streaming_kernel_with_short_segments(case):
    """The strip segments of the streaming kernel are taller than the golden frames; rebuild the harness with
    16-row segments so that segment seams (halo rows owned by the neighbour, TMA ring restarts, the per-pair
    sweeps of the many-candidate mode) are exercised on the CPU too."""
    g = Golden(case)
    ref, aux = O.run(g.inputs, g.outputs, g.opt(), g.noise, num_scales=g.num_scales)
    ref["loss"].backward()
    h = Golden(case)
    be = emu_backend("-DBBD_STREAM_RH=16 -DBBD_STREAM_RHM=16")
    losses, plan = run_fused(h.inputs, h.outputs, h.opt(), h.noise, h.num_scales, backend=be, groups=aux["groups"])
    losses["loss"].backward()
    for k, v in ref.items():
        assert abs(float(losses[k]) - float(v)) <= 2e-6 * max(1.0, abs(float(v))), k
    for k, v in g.params.items():
        if v.grad is not None:
            assert rel_l2(h.params[k].grad, v.grad) <= 1e-5, k


@pytest.mark.parametrize("flags", ["", "-DBBD_STREAM_RH=16 -DBBD_STREAM_RHM=16"])
def test_pipelined_form(flags, monkeypatch):
    """The opt-in three-warp form (bbd_pipe.cuh: gather / statistics / backward warps handing rows over through
    mbarrier-guarded rings) against the oracle, stepped with one fiber per thread of its 96-thread block; with
    16-row segments the rings wrap and the segment seams are crossed."""
    monkeypatch.setenv("BBD_PIPE", "1")
    g = Golden("plain_pm1")
    ref, aux = O.run(g.inputs, g.outputs, g.opt(), g.noise, num_scales=g.num_scales)
    ref["loss"].backward()
    h = Golden("plain_pm1")
    losses, plan = run_fused(h.inputs, h.outputs, h.opt(), h.noise, h.num_scales, backend=emu_backend(flags), groups=aux["groups"])
    losses["loss"].backward()
    for k, v in ref.items():
        assert abs(float(losses[k]) - float(v)) <= 2e-6 * max(1.0, abs(float(v))), k
    for k, v in g.params.items():
        if v.grad is not None:
            assert rel_l2(h.params[k].grad, v.grad) <= 1e-5, k
    # and bit-identical to the one-warp streaming form (same arithmetic, same order of adding)
    monkeypatch.setenv("BBD_PIPE", "0")
    s = Golden("plain_pm1")
    losses_s, _ = run_fused(s.inputs, s.outputs, s.opt(), s.noise, s.num_scales, backend=emu_backend(flags), groups=aux["groups"])
    losses_s["loss"].backward()
    for k in ref:
        assert float(losses[k]) == float(losses_s[k]), k


def test_far_baselines_that_never_win():
    """Tri-min batch whose +-2 / +-3 frames are far from the target (inverted intensities): their candidate pairs
    win no pixel, so the many-candidate kernel skips their sweeps of the gradient round (and zeroes their pose
    partials).  Loss, winners and gradients against the oracle on the same modified inputs."""
    def build():
        g = Golden("trimin_mixed")
        for k in list(g.inputs):
            if isinstance(k, tuple) and k[0] == "color" and k[1] in (2, -2, 3, -3):
                g.inputs[k] = g.inputs[k] + 3.0
        return g
    g = build()
    ref, aux = O.run(g.inputs, g.outputs, g.opt(), g.noise, num_scales=g.num_scales)
    ref["loss"].backward()
    h = build()
    be = emu_backend("-DBBD_STREAM_RHM=16")
    be.dll.emu_skipped_sweeps_total.restype = __import__("ctypes").c_long
    before = be.dll.emu_skipped_sweeps_total()
    losses, plan = run_fused(h.inputs, h.outputs, h.opt(), h.noise, h.num_scales, backend=be, groups=aux["groups"])
    assert be.dll.emu_skipped_sweeps_total() > before, "no sweep was skipped: the test does not reach the path it is for"
    losses["loss"].backward()
    for k, v in ref.items():
        assert abs(float(losses[k]) - float(v)) <= 2e-6 * max(1.0, abs(float(v))), k
    n_checked = 0
    for k, v in g.params.items():
        if v.grad is not None:
            assert rel_l2(h.params[k].grad, v.grad) <= 1e-5 or float(v.grad.abs().max()) == 0.0, k
            if float(v.grad.abs().max()) == 0.0:   # a pose nothing is warped with successfully: exactly zero here too
                assert h.params[k].grad is None or float(h.params[k].grad.abs().max()) == 0.0, k
            n_checked += 1
    assert n_checked > 0
