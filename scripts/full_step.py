"""BASELINE.json configs[4]: one full optimisation step around the fused loss, batch-sharded.

``python bench.py --workload full_step_640x192_b12 [--gpus N]`` lands here.  The step follows the
reference's ``Trainer.process_batch`` + ``run_epoch`` body (``trainer.py:255-264,286-298``, plain
+-1 frames): pose network per source frame -> ``transformation_from_parameters`` -> depth network ->
view-synthesis loss -> ``backward`` -> Adam, with the encoder/decoder gradients all-reduced by
DistributedDataParallel over NCCL.  The networks are NOT the product: they are plain torch/cuDNN
modules of the Monodepth2 shape (ResNet-18 encoder, skip-connected decoder with reflection-padded
3x3 convolutions + ELU and four sigmoid disparity heads, a 6-channel ResNet-18 pose encoder and a
four-convolution pose head), random-initialised, there only so that the loss kernels are measured
inside the step they ship in.  The loss is ``baseboostdepth_b200.trainer.FusedLossMixin``.

Reported (one JSON line, printed by bench.py): training steps/s and examples/s over all ranks, the
step's breakdown (networks fwd / loss fwd / backward+all-reduce / Adam) and, beside it, the same
step with the loss computed by stock PyTorch ops on the same GPU (``--loss eager`` leg).
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn
import torch.nn.functional as F

from baseboostdepth_b200 import layers as L
from baseboostdepth_b200.trainer import FusedLossMixin


# ----------------------------------------------------------------------------- networks (plain torch)
class _Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.c1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.b1 = nn.BatchNorm2d(cout)
        self.c2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.b2 = nn.BatchNorm2d(cout)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.b1(self.c1(x)), inplace=True)
        y = self.b2(self.c2(y))
        return F.relu(y + (x if self.down is None else self.down(x)), inplace=True)


class Encoder18(nn.Module):
    """ResNet-18 trunk returning the five feature maps (strides 2..32)."""
    channels = (64, 64, 128, 256, 512)

    def __init__(self, num_input_images=1):
        super().__init__()
        self.stem = nn.Conv2d(3 * num_input_images, 64, 7, 2, 3, bias=False)
        self.bn = nn.BatchNorm2d(64)
        widths, stages, cin = (64, 128, 256, 512), [], 64
        for i, w in enumerate(widths):
            stages.append(nn.Sequential(_Block(cin, w, 1 if i == 0 else 2), _Block(w, w, 1)))
            cin = w
        self.stages = nn.ModuleList(stages)

    def forward(self, image):
        x = (image - 0.45) / 0.225
        feats = [F.relu(self.bn(self.stem(x)), inplace=True)]
        x = F.max_pool2d(feats[0], 3, 2, 1)
        for st in self.stages:
            x = st(x)
            feats.append(x)
        return feats


class DepthDecoder(nn.Module):
    dec = (16, 32, 64, 128, 256)

    def __init__(self, enc_channels, scales=(0, 1, 2, 3)):
        super().__init__()
        self.scales = tuple(scales)
        self.up0, self.up1, self.heads = nn.ModuleDict(), nn.ModuleDict(), nn.ModuleDict()
        for i in range(4, -1, -1):
            cin = enc_channels[-1] if i == 4 else self.dec[i + 1]
            self.up0[str(i)] = L.ConvBlock(cin, self.dec[i])
            self.up1[str(i)] = L.ConvBlock(self.dec[i] + (enc_channels[i - 1] if i > 0 else 0), self.dec[i])
        for s in self.scales:
            self.heads[str(s)] = L.Conv3x3(self.dec[s], 1)

    def forward(self, feats):
        out, x = {}, feats[-1]
        for i in range(4, -1, -1):
            x = L.upsample(self.up0[str(i)](x))
            if i > 0:
                x = torch.cat([x, feats[i - 1]], 1)
            x = self.up1[str(i)](x)
            if i in self.scales:
                out[("disp", i)] = torch.sigmoid(self.heads[str(i)](x))
        return out


class PoseHead(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.squeeze = nn.Conv2d(cin, 256, 1)
        self.c0 = nn.Conv2d(256, 256, 3, 1, 1)
        self.c1 = nn.Conv2d(256, 256, 3, 1, 1)
        self.c2 = nn.Conv2d(256, 12, 1)

    def forward(self, feats):
        x = F.relu(self.squeeze(feats[-1]))
        x = self.c2(F.relu(self.c1(F.relu(self.c0(x)))))
        x = 0.01 * x.mean((2, 3)).view(-1, 2, 1, 6)
        return x[..., :3], x[..., 3:]


# ----------------------------------------------------------------------------- the step
class StepTrainer(FusedLossMixin):
    """The attributes and methods of the reference ``Trainer`` that one training step touches."""

    def __init__(self, batch, height, width, device, scales=(0, 1, 2, 3), loss="fused", ddp=False, local_rank=0):
        self.opt = types.SimpleNamespace(
            height=height, width=width, scales=list(scales), min_depth=0.1, max_depth=100.0,
            disparity_smoothness=1e-3, no_ssim=False, trimin=False, decomp=False, pose_error=5.5, SQL=False,
            frame_ids=[0, -1, 1], batch_size=batch, learning_rate=1e-4)
        self.device, self.num_scales, self.loss_impl = device, len(scales), loss
        enc = Encoder18(1)
        self.models = {"encoder": enc, "depth": DepthDecoder(enc.channels, scales),
                       "pose_encoder": Encoder18(2), "pose": PoseHead(512)}
        for k in self.models:
            self.models[k] = self.models[k].to(device).train()
        self.net = nn.ModuleDict(self.models)
        self.n_params = sum(p.numel() for p in self.net.parameters())
        self.ddp = None
        if ddp:
            # one bucketed, overlapped NCCL all-reduce of every network gradient
            self.ddp = nn.parallel.DistributedDataParallel(_Whole(self), device_ids=[local_rank],
                                                           gradient_as_bucket_view=True)
        self.optimizer = torch.optim.Adam(self.net.parameters(), self.opt.learning_rate, fused=True)
        if loss == "eager":
            self._eager = _EagerLoss(batch, height, width, device)

    # trainer.py:889-900 (plain mode, every sample has both +-1): all rows valid
    def valid_frames_trimin(self, inputs):
        n = len(inputs["ordering"])
        self.valid_mask_dict = {f: [True] * n for f in self.valid_frames}
        self.valid_mask = {f: [True] * n for f in self.valid_frames}

    def predict_poses(self, inputs):       # trainer.py:393-407 (non-incremental branch)
        outputs = {}
        for f in self.valid_frames:
            mid, other = inputs[("color_aug", 0, 0)], inputs[("color_aug", f, 0)]
            pair = [other, mid] if f < 0 else [mid, other]
            axisangle, translation = self.models["pose"](self.models["pose_encoder"](torch.cat(pair, 1)))
            outputs[("cam_T_cam", 0, f)] = L.transformation_from_parameters(
                axisangle[:, 0], translation[:, 0], invert=(f < 0))
        return outputs

    def process_batch(self, inputs):       # trainer.py:286-298
        self.valid_frames = sorted({f for o in inputs["ordering"] for f in o if f != 0}, key=lambda f: (abs(f), f < 0))
        self.valid_frames_trimin(inputs)
        outputs = self.predict_poses(inputs)
        outputs.update(self.models["depth"](self.models["encoder"](inputs[("color_aug", 0, 0)])))
        if self.loss_impl == "fused":
            outputs.update(self.generate_images_pred(inputs, outputs))
            losses = self.compute_losses(inputs, outputs)
        elif self.loss_impl == "eager":
            losses = self._eager(self, inputs, outputs)
        else:                              # "none": networks only, to see the loss's share of the step
            losses = {"loss": sum(outputs[("disp", s)].mean() for s in self.opt.scales)
                      + sum(outputs[("cam_T_cam", 0, f)].sum() for f in self.valid_frames) * 0.0}
        return outputs, losses

    def step(self, inputs):                # trainer.py:255-264
        if self.ddp is not None:
            losses = self.ddp(inputs)
        else:
            _, losses = self.process_batch(inputs)
        self.optimizer.zero_grad(set_to_none=True)
        losses["loss"].backward()
        self.optimizer.step()
        return losses


class _Whole(nn.Module):
    """DDP wants one module whose forward is the whole differentiable step."""

    def __init__(self, trainer):
        super().__init__()
        self.net = trainer.net
        self._t = [trainer]

    def forward(self, inputs):
        return self._t[0].process_batch(inputs)[1]


class _EagerLoss:
    """The same loss out of stock PyTorch ops on the GPU (what the reference launches today): tier-A
    semantics spelled with ``torch`` only -- no kernels of ours -- as the on-GPU comparison leg."""

    def __init__(self, batch, height, width, device):
        ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32), torch.arange(width, dtype=torch.float32),
                                indexing="ij")
        self.pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width)], 0)[None].to(device)
        self.H, self.W = height, width

    @staticmethod
    def ssim(x, y):
        x, y = F.pad(x, (1, 1, 1, 1), mode="reflect"), F.pad(y, (1, 1, 1, 1), mode="reflect")
        mx, my = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
        sx = F.avg_pool2d(x * x, 3, 1) - mx * mx
        sy = F.avg_pool2d(y * y, 3, 1) - my * my
        sxy = F.avg_pool2d(x * y, 3, 1) - mx * my
        n = (2 * mx * my + 1e-4) * (2 * sxy + 9e-4)
        d = (mx * mx + my * my + 1e-4) * (sx + sy + 9e-4)
        return torch.clamp((1 - n / d) / 2, 0, 1)

    def photometric(self, pred, target):
        return 0.85 * self.ssim(pred, target).mean(1, True) + 0.15 * (target - pred).abs().mean(1, True)

    def __call__(self, tr, inputs, outputs):
        H, W, opt = self.H, self.W, tr.opt
        target = inputs[("color", 0, 0)]
        B = target.shape[0]
        K, inv_K = inputs[("K", 0)], inputs[("inv_K", 0)]
        ident = [self.photometric(inputs[("color", f, 0)], target) for f in tr.valid_frames]
        noise = torch.randn(ident[0].shape, device=target.device) * 1e-5
        losses, total = {}, 0
        for s in opt.scales:
            disp = outputs[("disp", s)]
            up = F.interpolate(disp, [H, W], mode="bilinear", align_corners=False)
            depth = 1 / (1 / opt.max_depth + (1 / opt.min_depth - 1 / opt.max_depth) * up)
            cam = (inv_K[:, :3, :3] @ self.pix.expand(B, -1, -1)) * depth.view(B, 1, -1)
            cam = torch.cat([cam, torch.ones(B, 1, H * W, device=cam.device)], 1)
            cands = []
            for f in tr.valid_frames:
                P = (K @ outputs[("cam_T_cam", 0, f)])[:, :3]
                c = P @ cam
                pix = (c[:, :2] / (c[:, 2:3] + 1e-7)).view(B, 2, H, W).permute(0, 2, 3, 1)
                pix = torch.stack([pix[..., 0] / (W - 1), pix[..., 1] / (H - 1)], -1)
                warped = F.grid_sample(inputs[("color", f, 0)], (pix - 0.5) * 2, padding_mode="border",
                                       align_corners=True)
                cands.append(self.photometric(warped, target))
            allc = torch.cat(cands + [i + noise for i in ident], 1)
            loss = allc.min(1)[0].mean()
            nd = disp / (disp.mean((2, 3), True) + 1e-7)
            img = inputs[("color", 0, s)]
            gx = (nd[..., :, :-1] - nd[..., :, 1:]).abs() * torch.exp(-(img[..., :, :-1] - img[..., :, 1:]).abs().mean(1, True))
            gy = (nd[..., :-1, :] - nd[..., 1:, :]).abs() * torch.exp(-(img[..., :-1, :] - img[..., 1:, :]).abs().mean(1, True))
            loss = loss + opt.disparity_smoothness * (gx.mean() + gy.mean()) / (2 ** s)
            losses[f"loss/{s}"] = loss
            total = total + loss
        losses["loss"] = total / tr.num_scales
        return losses


def make_inputs(batch, height, width, device, seed, scales=(0, 1, 2, 3)):
    """Synthetic KITTI-shaped batch in the reference loader's layout (``mono_dataset.py:140-204``)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    inputs = {"ordering": [[0, 1, -1]] * batch}
    for f in (0, -1, 1):
        img = torch.rand(batch, 3, height, width, generator=g)
        inputs[("color", f, 0)] = img
        inputs[("color_aug", f, 0)] = img
    for s in scales[1:]:
        inputs[("color", 0, s)] = F.interpolate(inputs[("color", 0, 0)], scale_factor=1 / (2 ** s), mode="area")
    K = torch.tensor([[0.58 * width, 0, 0.5 * width, 0], [0, 1.92 * height, 0.5 * height, 0],
                      [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32)
    inputs[("K", 0)] = K[None].repeat(batch, 1, 1)
    inputs[("inv_K", 0)] = torch.linalg.pinv(K)[None].repeat(batch, 1, 1)
    return {k: (v.to(device).contiguous() if torch.is_tensor(v) else v) for k, v in inputs.items()}
