"""Seeded synthetic KITTI-shaped batches for parity tests and ``bench.py``.

Distributions follow SURVEY.md 8(d): images U(0,1); disparity
0.01 + 0.29*U(0,1) per scale; KITTI-normalised intrinsics
(reference ``datasets/kitti_dataset.py:16-21``) with fx, fy jittered +-5 % per
batch; axis-angle 0.01*N(0,1), translation 0.02*N(0,1)*|f|; stereo
baseline tx = +-0.1 (reference ``datasets/mono_dataset.py:136-139``).  The batch
layout (compacted per-frame stacks, ``ordering``) is what the reference's
``custom_collate`` produces (``trainer.py:867-886``).

There is no network in the build environment, so these stand in for KITTI;
``bench.py`` says ``"data": "synthetic"``.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

from .geometry import transformation_from_parameters
from .plan import build_plan, frame_sort_key

WORKLOADS = {
    # name: (batch, height, width, baselines, trimin, decomp)   -- BASELINE.json configs
    "kitti_640x192_b12_pm1": (12, 192, 640, [1] * 12, False, False),
    "trimin_mixed_640x192_b12": (12, 192, 640, [3] * 6 + [2] * 3 + [1] * 2 + ["s"], True, False),
    "trimin_decomp_640x192_b12": (12, 192, 640, [3] * 6 + [2] * 3 + [1] * 2 + ["s"], True, True),
    "trimin_all3_640x192_b12": (12, 192, 640, [3] * 12, True, False),
    "hires_1024x320_b8_pm1": (8, 320, 1024, [1] * 8, False, False),
}


def ordering_from_baselines(baselines: Sequence) -> List[list]:
    return [[0, "s"] if m == "s" else [0, m, -m] for m in baselines]


def make_intrinsics(batch, height, width, gen, device, dtype, jitter=0.05):
    j = 1.0 + jitter * (2.0 * torch.rand(2, generator=gen, dtype=torch.float64) - 1.0)
    K = torch.tensor([[0.58 * width * j[0].item(), 0, 0.5 * width, 0],
                      [0, 1.92 * height * j[1].item(), 0.5 * height, 0],
                      [0, 0, 1, 0],
                      [0, 0, 0, 1]], dtype=torch.float32)
    inv_K = torch.linalg.pinv(K)
    K = K.to(device=device, dtype=dtype).unsqueeze(0).repeat(batch, 1, 1).contiguous()
    inv_K = inv_K.to(device=device, dtype=dtype).unsqueeze(0).repeat(batch, 1, 1).contiguous()
    return K, inv_K


def make_batch(batch=12, height=192, width=640, baselines=None, scales=(0, 1, 2, 3), trimin=False,
               decomp=False, pose_error=5.5, seed=1234, device="cpu", dtype=torch.float32,
               stress=False, requires_grad=True):
    """Build ``(inputs, outputs, params)`` dicts in the reference trainer's layout.

    ``params`` holds the leaf tensors gradients are taken against
    (``("disp", s)``, ``("axisangle", f)``, ``("translation", f)``).
    Random numbers are drawn on the CPU generator so that a given seed gives
    the same batch on every device.
    """
    if baselines is None:
        baselines = [1] * batch
    assert len(baselines) == batch
    gen = torch.Generator(device="cpu").manual_seed(seed)
    ordering = ordering_from_baselines(baselines)
    plan = build_plan(ordering, trimin=trimin, decomp=decomp)

    def U(*shape):
        return torch.rand(*shape, generator=gen, dtype=torch.float32).to(device=device, dtype=dtype)

    def N(*shape):
        return torch.randn(*shape, generator=gen, dtype=torch.float32).to(device=device, dtype=dtype)

    inputs: Dict = {"ordering": ordering}
    color0 = U(batch, 3, height, width)
    for s in scales:
        inputs[("color", 0, s)] = color0 if s == 0 else F.avg_pool2d(color0, 2 ** s)

    numeric = [m for m in baselines if m != "s"]
    top = max(numeric, default=0)
    for f in sorted([f for f in range(-top, top + 1) if f != 0], key=frame_sort_key):
        rows = sum(1 for m in numeric if m >= abs(f))
        inputs[("color", f, 0)] = U(rows, 3, height, width)
    if "s" in plan.frames or trimin:
        members = [m for m in baselines if (m in ("s", 1, 2) if trimin else m == "s")]
        inputs[("color", "s", 0)] = U(len(members), 3, height, width)

    K, inv_K = make_intrinsics(batch, height, width, gen, device, dtype)
    inputs[("K", 0)], inputs[("inv_K", 0)] = K, inv_K
    stereo_T = torch.eye(4, dtype=torch.float32).unsqueeze(0).repeat(batch, 1, 1)
    sign = torch.where(torch.rand(batch, generator=gen) > 0.5, 1.0, -1.0)
    stereo_T[:, 0, 3] = 0.1 * sign
    inputs["stereo_T"] = stereo_T.to(device=device, dtype=dtype)

    outputs: Dict = {}
    params: Dict = {}
    for s in scales:
        d = (U(batch, 1, height >> s, width >> s) if stress
             else 0.01 + 0.29 * U(batch, 1, height >> s, width >> s))
        d.requires_grad_(requires_grad)
        params[("disp", s)] = d
        outputs[("disp", s)] = d
    t_sigma = 0.1 if stress else 0.02
    for f in plan.frames:
        if f == "s":
            continue
        n = len(plan.sel[f])
        aa = (0.01 * N(n, 1, 3)).requires_grad_(requires_grad)
        tr = (t_sigma * abs(f) * N(n, 1, 3)).requires_grad_(requires_grad)
        params[("axisangle", f)], params[("translation", f)] = aa, tr
        T = transformation_from_parameters(aa, tr, invert=(f < 0))
        outputs[("cam_T_cam", 0, f)] = T
        if decomp:
            T_err = T.clone().detach()
            T_err[:, :3, 3:] /= pose_error
            outputs[("cam_T_cam_error", 0, f)] = T_err
    # frames the reference still iterates (its valid_frames list is extended downwards,
    # trainer.py:961-981) although no sample of this batch is warped from them: zero rows
    for f in range(-top, top + 1):
        if f != 0 and f not in plan.frames:
            outputs[("cam_T_cam", 0, f)] = torch.zeros(0, 4, 4, device=device, dtype=dtype)
            if decomp:
                outputs[("cam_T_cam_error", 0, f)] = torch.zeros(0, 4, 4, device=device, dtype=dtype)
    return inputs, outputs, params


def make_noise(plan, height, width, seed=4321, device="cpu", dtype=torch.float32):
    """Per-group tie-break noise planes ``randn * 1e-5`` (reference ``trainer.py:518,522``)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    noise = {}
    for g in plan.groups:
        n = len(plan.group_members[g])
        noise[g] = (torch.randn(n, 1, height, width, generator=gen, dtype=torch.float32) * 0.00001
                    ).to(device=device, dtype=dtype)
    return noise


def px_pairs(plan, height, width, num_scales):
    """Warped px-pairs per step: H*W * sum_scales sum_frames rows."""
    return height * width * num_scales * plan.warps_per_scale


def algorithmic_bytes(plan, height, width, scales):
    """Compulsory fp32 traffic per step as defined in SURVEY.md 8(d) / BASELINE.md 3."""
    hw = height * width
    reproj = 0
    ident = 0
    for b in range(plan.batch):
        d = len({f for f, _ in plan.rep[b]})
        reproj += (44 + 24 * d) * hw * len(scales)
        ident += (20 + 12 * d) * hw
    smooth = sum(36 * plan.batch * (height >> s) * (width >> s) for s in scales)
    # the kernels either side of the path: compulsory traffic of the disparity <-> depth steps
    pyr = sum(plan.batch * (height >> s) * (width >> s) for s in scales)
    full = plan.batch * hw * len(scales)
    d2d_fwd = 4 * pyr + 4 * full                 # read every disparity level, write S full-resolution depth planes
    d2d_bwd = 8 * full + 4 * pyr + 4 * pyr       # read gdepth + depth, read the smoothness gradient, write gdisp
    return {"reproj": reproj, "identity": ident, "smooth": smooth, "total": reproj + ident + smooth,
            "d2d_forward": d2d_fwd, "d2d_backward": d2d_bwd}
