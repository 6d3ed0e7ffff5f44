"""Where one full step (scripts/full_step.py) spends its time: GPU kernel time vs host time per leg."""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import full_step as FS  # noqa: E402

dev = torch.device("cuda", 0)
B, H, W = 12, 192, 640
inputs = FS.make_inputs(B, H, W, dev, seed=1)
for leg in ("none", "fused", "eager"):
    torch.manual_seed(1)
    tr = FS.StepTrainer(B, H, W, dev, loss=leg)
    for _ in range(5):
        tr.step(inputs)
    torch.cuda.synchronize()
    n = 5
    t0 = time.perf_counter()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            tr.step(inputs)
        torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    gpu_ms = sum(e.device_time for e in ev) / n / 1e3 if ev else float("nan")
    print(f"leg={leg}: wall {wall:.2f} ms/step (under profiler), GPU kernel time {gpu_ms:.2f} ms/step, "
          f"{len(ev) / n:.0f} device ops/step")
    if leg == "fused":
        mine = {}
        for e in ev:
            if "bbd" in e.name or "kernel" in e.name and any(s in e.name for s in ("reproj", "ident", "smooth", "d2d", "pose")):
                mine[e.name[:60]] = mine.get(e.name[:60], 0) + e.device_time / n
        for k, v in sorted(mine.items(), key=lambda kv: -kv[1]):
            print(f"    {v:8.1f} us  {k}")
    # host time of the loss alone (no sync inside)
    if leg == "fused":
        outputs = tr.predict_poses(inputs)
        outputs.update(tr.models["depth"](tr.models["encoder"](inputs[("color_aug", 0, 0)])))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            tr.generate_images_pred(inputs, outputs)
            losses = tr.compute_losses(inputs, outputs)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"    host time of generate_images_pred+compute_losses: {(t1 - t0) / 20 * 1e3:.3f} ms/call")
    del tr
    torch.cuda.empty_cache()
