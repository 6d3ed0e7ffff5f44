"""Diagnostic (GPU box): where do selections differ at full size, and do coordinates / warps agree?"""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import max_abs
from fused_util import run_fused
from oracle import loss_path as O
from baseboostdepth_b200 import _lib
from baseboostdepth_b200.synthetic import make_batch, make_noise
from baseboostdepth_b200.trainer import plan_for, materialise_warps

dev = torch.device("cuda:0")
cfg = dict(batch=4, height=192, width=640, baselines=[3, 2, 1, "s"], trimin=True, decomp=False)
opt = O.default_opt(height=192, width=640, trimin=True, batch_size=4)
inputs, outputs, params = make_batch(seed=21, device="cpu", **cfg)
plan = plan_for(inputs["ordering"], True, False, inputs[("color", "s", 0)].shape[0])
noise = make_noise(plan, 192, 640, seed=5)
for threads in (torch.get_num_threads(), 1):
    torch.set_num_threads(threads)
    with torch.no_grad():
        o2 = dict(outputs)
        ref, aux = O.run(inputs, o2, opt, noise, num_scales=4)
    gi, go, gp = make_batch(seed=21, device=dev, **cfg)
    with torch.no_grad():
        losses, plan = run_fused(gi, go, opt, {k: v.to(dev) for k, v in noise.items()}, 4, groups=aux["groups"])
        materialise_warps(gi, go, opt, plan)
    print("threads", threads, "loss diff", float(losses["loss"]) - float(ref["loss"]))
    order = [b for grp in aux["groups"] for b in plan.group_members[grp]]
    for i, s in enumerate([0, 1, 2, 3]):
        args = torch.cat(aux["argmin"][s], 0); mine = go["argmin"][i].cpu()[order].long()
        dis = mine != args
        margins = []
        for p in aux["planes"][s]:
            top = torch.topk(-p, 2, dim=1).values; margins.append(top[:, 0] - top[:, 1])
        margin = torch.cat(margins, 0)
        print(" scale", s, "disagree", int(dis.sum()), "margins", [f"{float(m):.2e}" for m in margin[dis][:6]])
        for f in plan.frames:
            d = max_abs(go[("color", f, s)], o2[("color", f, s)])
            print("   frame", f, "warp maxabs %.2e" % d, "depth maxabs %.2e" % max_abs(go[("depth", 0, s)], o2[("depth", 0, s)]))
