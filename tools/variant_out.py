"""Run one library variant on a seeded input and save its maps + a short throughput measurement (under gpurun):
WST_BUILD_LIB=libwst_b200_h1.so python tools/variant_out.py p256j4 gpurun_out/x_h1   -> x_h1_p256j4.npy, prints patches/s
Variants are then compared with each other off-line (the default build is the one the oracle tests hold)."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import wst_b200
CFG = {"cfg5": (512, 5, 4, 148), "p256j4": (256, 4, 3, 592), "p256j2": (256, 2, 3, 592), "p256j3": (256, 3, 3, 592), "cfg3": (128, 4, 3, 4096)}
name, prefix = sys.argv[1], sys.argv[2]
M, J, C, B = CFG[name]
plan = wst_b200.get_plan(M, M, J, 8)
rng = np.random.default_rng(5)
x = torch.from_numpy((rng.integers(0, 256, (2, C, M, M)) / 255.0).astype(np.float32)).cuda()
feats, maps = plan.forward(x, True, True)
if not os.environ.get("WST_NO_SAVE"):
    np.save("%s_%s.npy" % (prefix, name), maps.cpu().numpy())
g = torch.Generator(device="cuda").manual_seed(1)
xb = torch.randint(0, 256, (B, C, M, M), device="cuda", generator=g, dtype=torch.int32).float().div_(255.0)
for _ in range(2):
    plan.forward(xb)
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 3
for _ in range(n):
    plan.forward(xb)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
print("%s %s: %.1f patches/s (grid %d)" % (os.environ.get("WST_BUILD_LIB", "libwst_b200.so"), name, B / dt, plan.grid), flush=True)
