"""Mint golden vectors from the UNMODIFIED reference (build container only).

Imports ``trainer.Trainer`` and ``layers`` from ``/root/reference`` (read-only),
drives ``valid_frames_trimin`` -> ``generate_images_pred`` -> ``compute_losses``
-> ``backward`` on small seeded synthetic batches, and writes inputs + outputs
to ``tests/golden/<case>.npz``.  The reference has no tests or fixtures of its
own, so these files are what pins ``oracle/`` to the reference.

Recipe: ``skimage`` / ``matplotlib`` are not installed and are
not touched by the loss path, so empty stubs are placed in ``sys.modules``;
``Trainer.__new__`` skips ``__init__`` (which needs wandb, KITTI and
``gt_depths.npz``) and the handful of attributes the loss path reads are set by
hand.  The tie-break noise the reference draws with ``torch.randn``
(``trainer.py:518,522``) is captured and stored, because its draw order follows a
``set`` iteration order that is not stable across processes.

Run:  python tests/golden/make_golden.py      (needs /root/reference)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)

from baseboostdepth_b200.synthetic import make_batch  # noqa: E402

CASES = {
    # name: dict(baselines, trimin, decomp, no_ssim, scales, H, W, seed)
    "plain_pm1": dict(baselines=[1, 1, 1, 1], trimin=False, decomp=False, no_ssim=False,
                      scales=[0, 1, 2, 3], H=32, W=64, seed=11),
    "plain_mixed_s": dict(baselines=[2, "s", 1, 2, "s"], trimin=False, decomp=False, no_ssim=False,
                          scales=[0, 1, 2, 3], H=32, W=64, seed=12),
    "plain_nossim": dict(baselines=[1, 1, 1], trimin=False, decomp=False, no_ssim=True,
                         scales=[0, 1], H=32, W=64, seed=13),
    "trimin_mixed": dict(baselines=[3, 2, 1, "s", 3, 2], trimin=True, decomp=False, no_ssim=False,
                         scales=[0, 1, 2, 3], H=32, W=64, seed=14),
    "trimin_decomp": dict(baselines=[3, 2, 1, "s", 3, 4], trimin=True, decomp=True, no_ssim=False,
                          scales=[0, 2], H=32, W=64, seed=15),
    "trimin_all3": dict(baselines=[3, 3, 3], trimin=True, decomp=False, no_ssim=False,
                        scales=[0], H=32, W=64, seed=16),
    "stress_oob": dict(baselines=[1, 1], trimin=False, decomp=False, no_ssim=False,
                       scales=[0, 3], H=32, W=64, seed=17, stress=True),
    # ragged geometry: rows not a multiple of the 16-row tile, columns not a multiple of 28
    "ragged_40x72": dict(baselines=[2, 1, "s"], trimin=True, decomp=False, no_ssim=False,
                         scales=[0, 1, 2, 3], H=40, W=72, seed=18),
    # a frame narrower than one tile, a single sample
    "small_16x24_b1": dict(baselines=[1], trimin=False, decomp=False, no_ssim=False,
                           scales=[0, 1, 2], H=16, W=24, seed=19),
    # decomp on a ragged frame, out-of-bounds motion
    "ragged_decomp_24x40": dict(baselines=[1, 2], trimin=True, decomp=True, no_ssim=False,
                                scales=[0, 1], H=24, W=40, seed=20, stress=True),
}


def import_reference():
    for name in ("skimage", "skimage.transform", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.pyplot"].get_cmap = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    threads = torch.get_num_threads()
    sys.path.insert(0, REF)
    import layers as ref_layers  # noqa
    import trainer as ref_trainer  # noqa
    torch.set_num_threads(threads)
    return ref_layers, ref_trainer


def key_str(k):
    return "|".join(str(p) for p in k) if isinstance(k, tuple) else str(k)


def run_case(name, cfg, ref_layers, ref_trainer):
    H, W = cfg["H"], cfg["W"]
    inputs, outputs, params = make_batch(batch=len(cfg["baselines"]), height=H, width=W,
                                         baselines=cfg["baselines"], scales=cfg["scales"],
                                         trimin=cfg["trimin"], decomp=cfg["decomp"], seed=cfg["seed"],
                                         stress=cfg.get("stress", False))
    B = len(cfg["baselines"])
    tr = ref_trainer.Trainer.__new__(ref_trainer.Trainer)
    tr.opt = types.SimpleNamespace(
        height=H, width=W, scales=list(cfg["scales"]), min_depth=0.1, max_depth=100.0,
        disparity_smoothness=1e-3, no_ssim=cfg["no_ssim"], trimin=cfg["trimin"], decomp=cfg["decomp"],
        pose_error=5.5, SQL=False, incremental_skip=False, partial_skip=False, batch_size=B)
    tr.device = torch.device("cpu")
    tr.num_scales = 4  # frozen at __init__ from the default --scales (trainer.py:44)
    tr.ssim = ref_layers.SSIM()
    tr.backproject_depth = {0: ref_layers.BackprojectDepth(B, H, W)}
    tr.project_3d = {0: ref_layers.Project3D(B, H, W)}
    tr.maxing_valid_frames = False

    from oracle.loss_path import frame_ids_from_ordering, initial_valid_frames
    tr.opt.frame_ids = frame_ids_from_ordering(inputs["ordering"])
    tr.valid_frames = initial_valid_frames(inputs["ordering"])   # trainer.py:292 (set order made explicit)
    tr.valid_frames_trimin(inputs)

    drawn = []
    real_randn = torch.randn

    def recording_randn(*a, **k):
        t = real_randn(*a, **k)
        drawn.append(t)
        return t

    torch.manual_seed(cfg["seed"] + 1000)
    tr.generate_images_pred(inputs, outputs)
    torch.randn = recording_randn
    try:
        losses = tr.compute_losses(inputs, outputs)
    finally:
        torch.randn = real_randn
    losses["loss"].backward()

    # noise draw order = temp_positive (trainer.py:516-523)
    if cfg["trimin"]:
        it = list(set(el for sub in inputs["ordering"] for el in sub if el != 0))
        order = [f for f in it if f == "s" or f > 0]
    else:
        order = [f for f in tr.valid_frames if f == "s" or f > 0]
    assert len(order) == len(drawn), (order, len(drawn))

    blob = {"meta_baselines": np.array([str(m) for m in cfg["baselines"]]),
            "meta_scales": np.array(cfg["scales"]), "meta_flags": np.array(
                [int(cfg["trimin"]), int(cfg["decomp"]), int(cfg["no_ssim"])]),
            "meta_hw": np.array([H, W]), "meta_num_scales": np.array([tr.num_scales]),
            "meta_valid_frames": np.array([str(f) for f in tr.valid_frames])}
    for k, v in inputs.items():
        if torch.is_tensor(v):
            blob["in:" + key_str(k)] = v.detach().numpy()
    for k, v in params.items():
        blob["param:" + key_str(k)] = v.detach().numpy()
        blob["grad:" + key_str(k)] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
    for g, t in zip(order, drawn):
        blob[f"noise:{g}"] = (t * 0.00001).numpy()
    for k, v in losses.items():
        blob["loss:" + k] = v.detach().numpy()
    for k, v in outputs.items():
        if k[0] in ("color", "color_D", "depth") and (k[0] == "depth" or k[2] == cfg["scales"][0]):
            blob["out:" + key_str(k)] = v.detach().numpy()
    # masks (lists of bool) as written by the reference
    for attr in ("valid_mask_dict", "valid_mask", "valid_tri_mask_dict", "valid_tri_mask",
                 "valid_tri_mask_reverse"):
        d = getattr(tr, attr, None)
        if d:
            for k, v in d.items():
                blob[f"mask:{attr}:{k}"] = np.array(v, dtype=bool)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: loss={float(losses['loss']):.8f}  -> {os.path.relpath(path, ROOT)} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    ref_layers, ref_trainer = import_reference()
    which = sys.argv[1:] or list(CASES)
    for name in which:
        run_case(name, CASES[name], ref_layers, ref_trainer)


if __name__ == "__main__":
    main()
