"""Projected coordinates of the streaming kernel's chain, stepped on the CPU, against the oracle's
Project3D grid and ATen's unnormalise + clip (bit-exact; the GPU twin is
tests/test_gpu_parity.py::test_projected_coordinates_bit_exact)."""
import ctypes as C

import pytest
import torch

from baseboostdepth_b200.synthetic import make_batch
from fused_util import emu_backend
from oracle import loss_path as O


@pytest.mark.parametrize("hw", [(48, 96), (40, 72), (17, 23)])
def test_stream_chain_coordinates_bit_exact(hw):
    H, W = hw
    be = emu_backend()
    batch = 3
    opt = O.default_opt(height=H, width=W, batch_size=batch, scales=[0])
    inputs, outputs, params = make_batch(seed=5, device="cpu", batch=batch, height=H, width=W, baselines=[1] * batch,
                                         scales=(0,), stress=(H == 17))
    ref_out = dict(outputs)
    with torch.no_grad():
        ordering = inputs["ordering"]
        masks = O.sub_batch_masks(ordering, O.frame_ids_from_ordering(ordering), O.initial_valid_frames(ordering), opt.trimin)
        O.view_synthesis(inputs, ref_out, opt, masks)
    depth = ref_out[("depth", 0, 0)].contiguous()
    inv_K = inputs[("inv_K", 0)].contiguous()
    for f in (1, -1):
        P = torch.matmul(inputs[("K", 0)], outputs[("cam_T_cam", 0, f)].detach())[:, :3, :].contiguous()
        want = ref_out[("grid", f, 0)]
        ix = ((want[..., 0] + 1) / 2 * (W - 1)).clamp(0, W - 1)
        iy = ((want[..., 1] + 1) / 2 * (H - 1)).clamp(0, H - 1)
        grid, pix = torch.empty(batch, 2, H, W), torch.empty(batch, 2, H, W)
        be.call("project_coords", batch, H, W, C.c_void_p(depth.data_ptr()), C.c_void_p(inv_K.data_ptr()),
                C.c_void_p(P.data_ptr()), C.c_void_p(grid.data_ptr()), C.c_void_p(pix.data_ptr()))
        assert torch.equal(grid.permute(0, 2, 3, 1), want)
        assert torch.equal(pix[:, 0], ix) and torch.equal(pix[:, 1], iy)
