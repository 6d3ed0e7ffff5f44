"""Kernel-level timings on the GPU box: fused loss forward-only vs forward+backward, per workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import workload, make_opt
from baseboostdepth_b200.synthetic import make_batch, make_noise
from baseboostdepth_b200.trainer import loss_step, plan_for

dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["kitti_640x192_b12_pm1"]:
    cfg = workload(name); opt = make_opt(cfg)
    inputs, outputs, params = make_batch(seed=1, device=dev, pose_error=5.5, **cfg)
    outputs = {k: (v.detach().requires_grad_(True) if v.numel() and k[0] in ("disp", "cam_T_cam") else v) for k, v in outputs.items()}
    plan = plan_for(inputs["ordering"], cfg["trimin"], cfg["decomp"], inputs[("color", "s", 0)].shape[0] if ("color", "s", 0) in inputs else None)
    noise = {g: n.to(dev) for g, n in make_noise(plan, cfg["height"], cfg["width"]).items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for grad in (True, False):
        ms = []
        for it in range(25):
            timers = {}
            flush.zero_()
            if grad:
                losses = loss_step(inputs, outputs, opt, plan, noise=noise, num_scales=4, timers=timers)
            else:
                with torch.no_grad():
                    losses = loss_step(inputs, outputs, opt, plan, noise=noise, num_scales=4, timers=timers)
            torch.cuda.synchronize()
            a, b = timers["reproj_fused"]
            if it >= 5: ms.append(a.elapsed_time(b))
        print(f"{name:28s} reproj kernel {'fwd+bwd' if grad else 'fwd only'}: {sum(ms)/len(ms)*1e3:8.1f} us")
