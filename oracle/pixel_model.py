"""Oracle, second layer: the ATen algorithms the reference leans on, restated per pixel in float64 numpy.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The arithmetic of the hot path lives in a third-party dependency of the reference that is not
under ``/root/reference``: PyTorch ATen (pinned ``pytorch=1.8.0``, ``environment.yml:162``).
``loss_path`` calls those ops at the reference's call sites; this module restates their published
algorithms so that the semantics the CUDA kernels implement are written down independently of
torch, and ``tests/test_oracle_golden.py`` checks the two layers against each other:

* ``grid_sample_border`` -- ``grid_sampler_2d`` bilinear, ``padding_mode="border"``,
  ``align_corners=True`` (ATen ``GridSampler.h``: unnormalise ``((g+1)/2)*(size-1)``, clip to
  ``[0, size-1]`` with zero gradient on/outside the border, taps at floor/floor+1, out-of-range taps
  contribute nothing), forward and gradient w.r.t. the grid.  Call sites: ``trainer.py:439,442``.
* ``ssim_map`` -- ``ReflectionPad2d(1)`` + five ``AvgPool2d(3, 1)`` + the SSIM algebra of
  ``layers.py:235-249``.
* ``backproject_project`` -- ``layers.py:160-167,181-195`` without any batching tricks.
* ``upsample_bilinear`` -- ``F.interpolate(mode="bilinear", align_corners=False)`` (ATen
  ``UpSample.h``: source index ``max(scale*(dst+0.5)-0.5, 0)``), ``trainer.py:456``.

Plain loops; meant for images of a few hundred pixels.
"""
from __future__ import annotations

import numpy as np


def reflect(i, n):
    """ReflectionPad2d(1) index: -1 -> 1, n -> n-2."""
    if i < 0:
        return -i
    if i >= n:
        return 2 * n - 2 - i
    return i


def grid_sample_border(img, grid, gout=None):
    """img (C,H,W), grid (Ho,Wo,2) in [-1,1]-normalised coordinates -> out (C,Ho,Wo).

    With ``gout`` (C,Ho,Wo) also returns d(sum(out*gout))/d(grid), shape (Ho,Wo,2)."""
    C, H, W = img.shape
    Ho, Wo, _ = grid.shape
    out = np.zeros((C, Ho, Wo))
    ggrid = np.zeros((Ho, Wo, 2))
    for i in range(Ho):
        for j in range(Wo):
            x = (grid[i, j, 0] + 1) / 2 * (W - 1)
            y = (grid[i, j, 1] + 1) / 2 * (H - 1)
            mx = 1.0 if 0 < x < W - 1 else 0.0          # clip_coordinates_set_grad
            my = 1.0 if 0 < y < H - 1 else 0.0
            x = min(max(x, 0.0), W - 1.0)
            y = min(max(y, 0.0), H - 1.0)
            x0, y0 = int(np.floor(x)), int(np.floor(y))
            taps = [(x0, y0, (x0 + 1 - x) * (y0 + 1 - y), -(y0 + 1 - y), -(x0 + 1 - x)),
                    (x0 + 1, y0, (x - x0) * (y0 + 1 - y), (y0 + 1 - y), -(x - x0)),
                    (x0, y0 + 1, (x0 + 1 - x) * (y - y0), -(y - y0), (x0 + 1 - x)),
                    (x0 + 1, y0 + 1, (x - x0) * (y - y0), (y - y0), (x - x0))]
            gx = gy = 0.0
            for tx, ty, w, dwx, dwy in taps:
                if 0 <= tx < W and 0 <= ty < H:
                    out[:, i, j] += img[:, ty, tx] * w
                    if gout is not None:
                        gx += float(np.dot(img[:, ty, tx], gout[:, i, j])) * dwx
                        gy += float(np.dot(img[:, ty, tx], gout[:, i, j])) * dwy
            ggrid[i, j, 0] = gx * mx * (W - 1) / 2
            ggrid[i, j, 1] = gy * my * (H - 1) / 2
    return (out, ggrid) if gout is not None else out


def ssim_map(x, y):
    """x, y (C,H,W) -> clamp((1 - SSIM)/2, 0, 1) with 3x3 reflection-padded mean pools."""
    C, H, W = x.shape
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    out = np.zeros_like(x)
    for c in range(C):
        for i in range(H):
            for j in range(W):
                wx = np.array([[x[c, reflect(i + di, H), reflect(j + dj, W)] for dj in (-1, 0, 1)] for di in (-1, 0, 1)])
                wy = np.array([[y[c, reflect(i + di, H), reflect(j + dj, W)] for dj in (-1, 0, 1)] for di in (-1, 0, 1)])
                mu_x, mu_y = wx.mean(), wy.mean()
                sig_x = (wx ** 2).mean() - mu_x ** 2
                sig_y = (wy ** 2).mean() - mu_y ** 2
                sig_xy = (wx * wy).mean() - mu_x * mu_y
                n = (2 * mu_x * mu_y + c1) * (2 * sig_xy + c2)
                d = (mu_x ** 2 + mu_y ** 2 + c1) * (sig_x + sig_y + c2)
                out[c, i, j] = min(max((1 - n / d) / 2, 0.0), 1.0)
    return out


def backproject_project(depth, inv_K, K, T, eps=1e-7):
    """depth (H,W), 4x4 matrices -> normalised sampling grid (H,W,2) (layers.py:160-195)."""
    H, W = depth.shape
    P = (K @ T)[:3, :]
    grid = np.zeros((H, W, 2))
    for i in range(H):
        for j in range(W):
            ray = inv_K[:3, :3] @ np.array([j, i, 1.0])
            cam = np.append(depth[i, j] * ray, 1.0)
            c = P @ cam
            px, py = c[0] / (c[2] + eps), c[1] / (c[2] + eps)
            grid[i, j, 0] = (px / (W - 1) - 0.5) * 2
            grid[i, j, 1] = (py / (H - 1) - 0.5) * 2
    return grid


def upsample_bilinear(d, H, W):
    """d (h,w) -> (H,W), align_corners=False."""
    h, w = d.shape
    out = np.zeros((H, W))

    def taps(o, n_in, n_out):
        src = max(n_in / n_out * (o + 0.5) - 0.5, 0.0)
        i0 = int(src)
        i1 = i0 + (1 if i0 < n_in - 1 else 0)
        return i0, i1, 1 - (src - i0), src - i0

    for i in range(H):
        y0, y1, wy0, wy1 = taps(i, h, H)
        for j in range(W):
            x0, x1, wx0, wx1 = taps(j, w, W)
            out[i, j] = wy0 * (wx0 * d[y0, x0] + wx1 * d[y0, x1]) + wy1 * (wx0 * d[y1, x0] + wx1 * d[y1, x1])
    return out


def reprojection_loss(pred, target, no_ssim=False):
    """(C,H,W) x2 -> (H,W): 0.85*mean_c(SSIM) + 0.15*mean_c(|target - pred|) (trainer.py:477-486)."""
    l1 = np.abs(target - pred).mean(0)
    if no_ssim:
        return l1
    return 0.85 * ssim_map(pred, target).mean(0) + 0.15 * l1
