"""Diagnostic (GPU box): fused CUDA path vs the CPU emulation of the same kernel source vs the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import Golden, golden_cases, rel_l2, max_abs
from fused_util import emu_backend, run_fused
from oracle import loss_path as O

dev = torch.device("cuda:0")
for case in (sys.argv[1:] or golden_cases()):
    g = Golden(case)
    ref, aux = O.run(g.inputs, g.outputs, g.opt(), g.noise, num_scales=g.num_scales)
    ref["loss"].backward()
    e = Golden(case)
    le, plan = run_fused(e.inputs, e.outputs, e.opt(), e.noise, e.num_scales, backend=emu_backend(), groups=aux["groups"])
    le["loss"].backward()
    h = Golden(case, device=dev)
    noise = {k: v.to(dev) for k, v in h.noise.items()}
    lg, plan = run_fused(h.inputs, h.outputs, h.opt(), noise, h.num_scales, groups=aux["groups"])
    lg["loss"].backward()
    torch.cuda.synchronize()
    print(f"== {case}: loss gpu {float(lg['loss']):.9f} emu {float(le['loss']):.9f} oracle {float(ref['loss']):.9f}")
    wg, we = h.outputs["argmin"].cpu(), e.outputs["argmin"]
    print("   winner planes differ at", int((wg != we).sum()), "of", wg.numel())
    for i, s in enumerate(h.scales):
        print(f"   loss/{s} gpu-emu {float(lg[f'loss/{s}'])-float(le[f'loss/{s}']):+.2e}")
    for k, p in g.params.items():
        if p.grad is None:
            continue
        print(f"   {str(k):24s} gpu-vs-oracle {rel_l2(h.params[k].grad, p.grad):.2e}  gpu-vs-emu {rel_l2(h.params[k].grad, e.params[k].grad):.2e}  emu-vs-oracle {rel_l2(e.params[k].grad, p.grad):.2e}")
