"""Small fused-loss run for compute-sanitizer (memcheck / racecheck / initcheck) on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import Golden
from fused_util import mirror_to_device, run_fused
from baseboostdepth_b200.trainer import materialise_warps

dev = torch.device("cuda:0")
for case in ("plain_mixed_s", "trimin_decomp", "ragged_40x72", "small_16x24_b1"):
    h = Golden(case)
    hi, ho, leaves = mirror_to_device(h.inputs, h.outputs, h.params, dev)
    noise = {k: v.to(dev) for k, v in h.noise.items()}
    losses, plan = run_fused(hi, ho, h.opt(), noise, h.num_scales)
    losses["loss"].backward()
    with torch.no_grad():
        materialise_warps(hi, ho, h.opt(), plan)
    torch.cuda.synchronize()
    print(case, float(losses["loss"]))
