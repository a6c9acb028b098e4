"""Randomised parity sweep on the GPU (run under gpurun): random shapes / J / L / max_order / channel counts / batch sizes
through whichever engine the plan picks, float32 and uint8 ingest, small batches (split signals) and larger ones, against the
float64 oracle per order.  Prints one line per case and a summary; exits non-zero on any failure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import wst_b200
from oracle import Scattering2D
from tests.parity import parity_report

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fused = [(32, 1), (32, 2), (32, 3), (32, 4), (64, 2), (64, 3), (64, 4), (128, 2), (128, 3), (128, 4), (48, 3), (96, 2), (96, 4), (48, 2), (96, 3)]
worst, fails = 0.0, 0
for case in range(ncases):
    if case % 2 == 0:                                   # a compiled cascade, random batch geometry
        M, J = fused[rng.integers(len(fused))]
        H = W = M
        L = 8 if rng.random() < 0.7 else int(rng.integers(1, 9))
    else:                                               # any shape: DFT-matrix engine
        J = int(rng.integers(1, 4))
        H = int(rng.integers(2 ** J + 1, 90)); W = int(rng.integers(2 ** J + 1, 90))
        L = int(rng.integers(1, 11))
    mo = 1 if rng.random() < 0.2 else 2
    C = int(rng.integers(1, 5)); B = int(rng.choice([1, 2, 3, 7, 40, 130]))
    if H * W * B * C > 3e6:
        B = max(1, int(3e6 // (H * W * C)))
    u8 = rng.random() < 0.4
    px = rng.integers(0, 256, (B, H, W, C), dtype=np.uint8)
    chw = np.ascontiguousarray(np.transpose(px.astype(np.float32) / 255.0, (0, 3, 1, 2)))
    try:
        plan = wst_b200.get_plan(H, W, J, L, mo)
        x = torch.from_numpy(px).cuda() if u8 else torch.from_numpy(chw).cuda()
        feats, maps = plan.forward(x, True, True)
        f2 = plan.forward(x)[0]
        nchk = min(B, 2)
        ref = Scattering2D(J=J, shape=(H, W), L=L, max_order=mo, precision="double", cache_filters=True)(chw[:nchk].astype(np.float64))
        K = ref.shape[-3]
        got = maps[:nchk].cpu().numpy()
        rep = parity_report(got.reshape(nchk * C, K, -1), ref.reshape(nchk * C, K, -1), J, L, mo)
        w = max(r["floored"] for r in rep.values())
        fm = feats[:nchk].cpu().numpy().reshape(nchk * C, 2, K)
        repm = parity_report(fm[:, 0], ref.mean(axis=(-2, -1)).reshape(nchk * C, K), J, L, mo)
        w = max(w, max(r["floored"] for r in repm.values()))
        same = bool(torch.equal(feats, f2))
        ok = w <= 1e-4 and same and not np.isnan(got).any()
    except Exception as e:
        ok, w, same = False, float("nan"), False
        print("EXC", repr(e))
    worst = max(worst, w if w == w else 1.0); fails += not ok
    print("%s %3dx%-3d J=%d L=%-2d mo=%d C=%d B=%-3d %-5s engine=%-5s worst floored %.2e feats-stable %s" % (
        "ok  " if ok else "FAIL", H, W, J, L, mo, C, B, "u8" if u8 else "f32", plan.engine, w, same), flush=True)
print("cases %d  failures %d  worst floored rel err %.3g" % (ncases, fails, worst))
sys.exit(1 if fails else 0)
