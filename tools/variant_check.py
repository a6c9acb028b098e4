"""A/B run of one library variant (under gpurun): parity of a few patches against the oracle, bench.py throughput
and optionally the per-phase cycle table.
usage: WST_BUILD_LIB=libwst_b200_v1.so python tools/variant_check.py cfg3 [--phases]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import wst_b200
from oracle import Scattering2D
from tests.parity import parity_report

CFG = {"cfg1": (32, 2), "cfg2": (64, 3), "cfg3": (128, 4), "repo": (128, 2), "cfg5": (512, 5), "p256j4": (256, 4)}
name = sys.argv[1]
M, J = CFG[name]
lib = os.environ.get("WST_BUILD_LIB", "libwst_b200.so")
rng = np.random.default_rng(3)
nb = 1 if M >= 256 else 2
x = (rng.integers(0, 256, (nb, 3, M, M)) / 255.0).astype(np.float32)
plan = wst_b200.get_plan(M, M, J, 8)
xs = torch.from_numpy(np.concatenate([x] * 80)).cuda() if M < 256 else torch.from_numpy(x).cuda()   # several signals per CTA: the prefetch path runs
feats, maps = plan.forward(xs, True, True)
ref = Scattering2D(J=J, shape=(M, M), L=8, precision="double", cache_filters=True)(x)
K = ref.shape[-3]
got = maps.cpu().numpy()
rep = parity_report(got[:nb].reshape(nb * 3, K, -1), ref.reshape(nb * 3, K, -1), J, 8)
same = all(np.array_equal(got[i * nb:(i + 1) * nb], got[:nb]) for i in range(got.shape[0] // nb))
worst = max(r["floored"] for r in rep.values())
out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", name, "--no-cpu", "--steps", "5", "--warmup", "3"],
                     capture_output=True, text=True).stdout.strip().splitlines()
try:
    d = json.loads(out[-1]); perf = "%.1f patches/s  e2e %.1f  fp32 %.4f clocks %s" % (d["value"], d["e2e"]["value"], d["fp32"]["frac"], d["clocks"])
except Exception as e:
    perf = "BENCH FAILED %r" % (e,)
print("%s %s: parity worst floored %.3g (%s) repeat-identical %s | %s" % (lib, name, worst, "OK" if worst <= 1e-4 else "FAIL", same, perf), flush=True)
if "--phases" in sys.argv:
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "phase_profile.py"), name])
