"""Small, fixed workload for ncu captures of the cascade kernel: one persistent-grid wave of signals.
usage: python tools/ncu_target.py [cfg3] [waves]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wst_b200
CFG = {"cfg1": (32, 2), "cfg2": (64, 3), "cfg3": (128, 4), "repo": (128, 2), "cfg5": (512, 5)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
waves = int(sys.argv[2]) if len(sys.argv) > 2 else 1
M, J = CFG[name]
plan = wst_b200.get_plan(M, M, J, 8)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randint(0, 256, (148 * waves, 1, M, M), device="cuda", generator=g).float() / 255.0
for _ in range(3):
    f, _ = plan.forward(x)
torch.cuda.synchronize()
print("ok", tuple(f.shape), float(f.sum()))
