"""Tier B end to end: the reference's Trainer with FusedLossMixin mixed in (kernels stepped on the CPU).

Needs the reference checkout (build container); on the GPU box the same mixin is exercised through
loss_step in test_gpu_parity.py.
"""
import os
import sys
import types

import pytest
import torch

from baseboostdepth_b200.trainer import FusedLossMixin
from fused_util import emu_backend
from helpers import Golden, rel_l2
from oracle import loss_path as O

REF = "/root/reference"


def _reference_trainer():
    if not os.path.isdir(REF):
        pytest.skip("reference not mounted (GPU box)")
    for name in ("skimage", "skimage.transform", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.pyplot"].get_cmap = lambda *a, **k: None
    sys.dont_write_bytecode = True
    threads = torch.get_num_threads()
    for m in ("trainer", "layers"):
        sys.modules.pop(m, None)
    sys.path.insert(0, REF)
    try:
        import trainer as ref_trainer
    finally:
        sys.path.remove(REF)
        torch.set_num_threads(threads)
        for m in ("layers",):
            sys.modules.pop(m, None)
    mod = ref_trainer
    sys.modules.pop("trainer", None)
    return mod


@pytest.mark.parametrize("case", ["plain_mixed_s", "trimin_decomp"])
def test_reference_trainer_with_fused_mixin(case):
    ref_trainer = _reference_trainer()

    class FusedTrainer(FusedLossMixin, ref_trainer.Trainer):
        pass

    g = Golden(case)
    tr = FusedTrainer.__new__(FusedTrainer)
    tr.opt = types.SimpleNamespace(**vars(g.opt()))
    tr.device, tr.num_scales, tr.maxing_valid_frames = torch.device("cpu"), g.num_scales, False
    tr._bbd_backend = emu_backend()
    tr.collect_ident = True
    tr.early_phase = 0                                  # a logging step: warps must be materialised
    tr.opt.frame_ids = O.frame_ids_from_ordering(g.ordering)
    tr.valid_frames = O.initial_valid_frames(g.ordering)
    tr.valid_frames_trimin(g.inputs)                    # the reference's own bookkeeping

    groups = [f for f in O.initial_valid_frames(g.ordering) if f == "s" or f > 0]
    raw = iter([g.noise[k] / 0.00001 for k in groups])   # the kernel applies the 1e-5 itself
    real_randn = torch.randn
    torch.randn = lambda *a, **k: next(raw)
    try:
        outputs = tr.generate_images_pred(g.inputs, g.outputs)
        losses = tr.compute_losses(g.inputs, outputs)
    finally:
        torch.randn = real_randn
    for k, v in g.losses.items():
        assert abs(float(losses[k]) - v) <= 2e-6, (k, float(losses[k]), v)
    losses["loss"].backward()
    for k, ref in g.grads.items():
        assert rel_l2(g.params[k].grad, ref) <= 1e-5, k
    s0 = g.scales[0]
    for k, ref in g.ref_out.items():
        if k[0] in ("color", "color_D") and k[2] == s0:
            assert (outputs[k] - ref).abs().max() <= 2e-5, k
        if k[0] == "depth":
            assert (outputs[k] - ref).abs().max() <= 1e-4, k
    if g.trimin:
        norm, guide = tr.ident
        assert any(len(v) for v in norm.values())
