"""scripts/full_step.py (BASELINE.json configs[4] harness) on the CPU: the step built on the fused loss
(kernel source stepped by tests/emu) gives the same losses and the same network gradients as the same
step with the loss spelled in stock PyTorch ops."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scripts"))

from baseboostdepth_b200 import layers as L
from fused_util import emu_backend


def test_full_step_fused_equals_stock_loss(monkeypatch):
    import full_step as FS
    monkeypatch.setattr(L, "_TEST_BACKEND", emu_backend())
    real_adam = torch.optim.Adam
    monkeypatch.setattr(torch.optim, "Adam", lambda p, lr, fused=True: real_adam(p, lr))   # fused Adam is CUDA-only
    dev = torch.device("cpu")
    torch.manual_seed(0)
    a = FS.StepTrainer(2, 64, 96, dev, loss="fused")
    a._bbd_backend = emu_backend()
    b = FS.StepTrainer(2, 64, 96, dev, loss="eager")
    b.net.load_state_dict(a.net.state_dict())
    inputs = FS.make_inputs(2, 64, 96, dev, seed=3)
    torch.manual_seed(5)
    la = a.process_batch(inputs)[1]
    torch.manual_seed(5)          # same tie-break noise draw
    lb = b.process_batch(inputs)[1]
    for k in lb:
        assert abs(float(la[k].detach()) - float(lb[k].detach())) < 2e-6, k
    la["loss"].backward()
    lb["loss"].backward()
    ga = torch.cat([p.grad.flatten() for p in a.net.parameters()])
    gb = torch.cat([p.grad.flatten() for p in b.net.parameters()])
    assert float((ga - gb).norm() / gb.norm()) < 1e-5
    # and one optimiser step runs
    a.step(inputs)
