"""Pose / disparity glue that sits either side of the view-synthesis loss.

These are the small PyTorch helpers the reference trainer imports from its
``layers`` module next to the hot-path layers (reference ``layers.py:13-100``
and ``:197-200``, ``:252-286``).  They are not on the CUDA hot path; they are
kept here so that ``baseboostdepth_b200.layers`` is a complete drop-in for
``from layers import ...`` (reference ``trainer.py:21-22``,
``networks/depth_decoder.py:8``).

Numerics: every element is produced by the same sequence of fp32 operations
as the reference so results are bit-identical on the same device; only the
way the 4x4 matrices are assembled differs (stack instead of ~10 indexed
writes into a zero tensor, which costs ~40 tiny kernels per call).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def disp_to_depth(disp, min_depth, max_depth):
    """Sigmoid disparity -> (scaled disparity, depth).  Reference ``layers.py:13-22``."""
    lo = 1 / max_depth
    hi = 1 / min_depth
    scaled = lo + (hi - lo) * disp
    return scaled, 1 / scaled


def rot_from_axisangle(vec):
    """Axis-angle (B,1,3) -> homogeneous rotation (B,4,4).  Reference ``layers.py:61-100``."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = (axis[..., i].unsqueeze(1) for i in range(3))
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    zero = torch.zeros_like(ca)
    one = torch.ones_like(ca)
    rows = [
        [x * xC + ca, xyC - zs, zxC + ys, zero],
        [xyC + zs, y * yC + ca, yzC - xs, zero],
        [zxC - ys, yzC + xs, z * zC + ca, zero],
        [zero, zero, zero, one],
    ]
    flat = torch.cat([e.reshape(-1, 1) for r in rows for e in r], dim=1)
    return flat.view(-1, 4, 4)


def get_translation_matrix(translation_vector):
    """Translation (B,1,3)/(B,3) -> homogeneous (B,4,4).  Reference ``layers.py:45-58``."""
    t = translation_vector.contiguous().view(-1, 3, 1)
    n = t.shape[0]
    eye = torch.eye(4, device=t.device, dtype=t.dtype).expand(n, 4, 4)
    top = torch.cat([eye[:, :3, :3], t], dim=2)
    return torch.cat([top, eye[:, 3:, :]], dim=1)


def transformation_from_parameters(axisangle, translation, invert=False):
    """(axis-angle, translation) -> 4x4 camera motion.  Reference ``layers.py:25-42``."""
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


class Conv3x3(nn.Module):
    """Pad (reflect or zero) then 3x3 conv.  Reference ``layers.py:118-133``."""

    def __init__(self, in_channels, out_channels, use_refl=True):
        super().__init__()
        self.pad = nn.ReflectionPad2d(1) if use_refl else nn.ZeroPad2d(1)
        self.conv = nn.Conv2d(int(in_channels), int(out_channels), 3)

    def forward(self, x):
        return self.conv(self.pad(x))


class ConvBlock(nn.Module):
    """Conv3x3 + ELU.  Reference ``layers.py:103-115``."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = Conv3x3(in_channels, out_channels)
        self.nonlin = nn.ELU(inplace=True)

    def forward(self, x):
        return self.nonlin(self.conv(x))


def upsample(x):
    """Nearest x2 upsample.  Reference ``layers.py:197-200``."""
    return F.interpolate(x, scale_factor=2, mode="nearest")


def compute_depth_errors(gt, pred, mask=None, SYNS=False):
    """Depth / edge metrics used by validation.  Reference ``layers.py:252-286``."""
    if SYNS:
        from scipy import ndimage

        mask = np.logical_and(mask, gt[:, :, 0])
        th_edges = 10
        d_target = ndimage.distance_transform_edt(1 - mask)
        d_pred = ndimage.distance_transform_edt(1 - pred[:, :, 0])
        pred_edges = pred[:, :, 0] & (d_target < th_edges)
        any_edge = bool(pred_edges.sum())
        edge_acc = d_target[pred_edges].mean() if any_edge else th_edges
        edge_comp = d_pred[mask].mean() if any_edge else th_edges
        return edge_acc, edge_comp

    ratio = torch.max(gt / pred, pred / gt)
    a1, a2, a3 = ((ratio < 1.25 ** k).float().mean() for k in (1, 2, 3))
    rmse = torch.sqrt(((gt - pred) ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = torch.mean(torch.abs(gt - pred) / gt)
    sq_rel = torch.mean((gt - pred) ** 2 / gt)
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3
