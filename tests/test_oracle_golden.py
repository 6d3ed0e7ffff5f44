"""Pin the oracle: it must reproduce the reference's own outputs (tests/golden/*.npz).

The fixtures were minted by tests/golden/make_golden.py from the unmodified
reference on this container's CPU.  The oracle issues the same ATen ops, so on
the CPU the agreement is to the last few ulps; tolerances below leave room for
a different BLAS/threading on another host.
"""
import pytest
import torch

from helpers import Golden, golden_cases, max_abs, rel_l2
from oracle import loss_path as O

CASES = golden_cases()


def test_fixtures_present():
    assert len(CASES) >= 10


@pytest.mark.parametrize("case", CASES)
def test_masks_match_reference(case):
    g = Golden(case)
    fids = O.frame_ids_from_ordering(g.ordering)
    m = O.sub_batch_masks(g.ordering, fids, O.initial_valid_frames(g.ordering), g.trimin)
    assert m.valid_frames == g.valid_frames
    for attr, ref in g.masks.items():
        mine = getattr(m, attr)
        assert {str(k) for k in mine} == set(ref), (attr, sorted(map(str, mine)), sorted(ref))
        for k, v in mine.items():
            assert list(v) == ref[str(k)], (attr, k)


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference(case):
    g = Golden(case)
    out, aux = O.run(g.inputs, g.outputs, g.opt(), g.noise, num_scales=g.num_scales)
    for k, v in g.losses.items():
        assert abs(float(out[k]) - v) <= 2e-7 * max(1.0, abs(v)), (k, float(out[k]), v)
    # warped images / depth planes of the first scale
    for k, ref in g.ref_out.items():
        assert max_abs(g.outputs[k], ref) <= 1e-6, k
    out["loss"].backward()
    for k, ref in g.grads.items():
        got = g.params[k].grad
        got = torch.zeros_like(ref) if got is None else got
        assert rel_l2(got, ref) <= 1e-6, (k, rel_l2(got, ref))


def test_argmin_shapes():
    g = Golden("trimin_mixed")
    out, aux = O.run(g.inputs, g.outputs, g.opt(), g.noise, num_scales=g.num_scales)
    n = sum(a.shape[0] for a in aux["argmin"][0])
    assert n == len(g.baselines)


def test_pixel_model_agrees_with_aten_call_sites():
    """oracle.pixel_model (published ATen algorithms, float64 loops) vs oracle.loss_path (the ATen ops
    themselves) on a tiny frame: sampling grid, warp, its grid gradient, SSIM, photometric loss, upsample."""
    import numpy as np
    from oracle import pixel_model as M
    gen = torch.Generator().manual_seed(5)
    H, W = 7, 9
    img = torch.rand(1, 3, H, W, generator=gen, dtype=torch.float64)
    tgt = torch.rand(1, 3, H, W, generator=gen, dtype=torch.float64)
    depth = 1 + 4 * torch.rand(1, 1, H, W, generator=gen, dtype=torch.float64)
    K = torch.tensor([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]],
                     dtype=torch.float64).unsqueeze(0)
    inv_K = torch.linalg.pinv(K)
    T = torch.eye(4, dtype=torch.float64).unsqueeze(0)
    T[0, :3, 3] = torch.tensor([0.4, -0.1, 0.05], dtype=torch.float64)     # large motion: some samples leave the frame
    grid = O.project(O.backproject(depth, inv_K, H, W), K, T, H, W).detach().clone().requires_grad_(True)
    warped = O.warp(img, grid)
    gout = torch.rand(1, 3, H, W, generator=gen, dtype=torch.float64)
    (warped * gout).sum().backward()

    g_np = M.backproject_project(depth[0, 0].numpy(), inv_K[0].numpy(), K[0].numpy(), T[0].numpy())
    assert np.abs(g_np - grid.detach()[0].numpy()).max() < 1e-12
    w_np, gg_np = M.grid_sample_border(img[0].numpy(), g_np, gout[0].numpy())
    assert np.abs(w_np - warped.detach()[0].numpy()).max() < 1e-12
    assert np.abs(gg_np - grid.grad[0].numpy()).max() < 1e-10
    assert np.abs(M.ssim_map(w_np, tgt[0].numpy()) - O.ssim(warped.detach(), tgt)[0].numpy()).max() < 1e-12
    assert np.abs(M.reprojection_loss(w_np, tgt[0].numpy()) - O.reprojection_loss(warped.detach(), tgt)[0, 0].numpy()).max() < 1e-12
    small = torch.rand(1, 1, 4, 6, generator=gen, dtype=torch.float64)
    up = torch.nn.functional.interpolate(small, [8, 12], mode="bilinear", align_corners=False)
    assert np.abs(M.upsample_bilinear(small[0, 0].numpy(), 8, 12) - up[0, 0].numpy()).max() < 1e-12
