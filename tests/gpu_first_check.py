"""Ad-hoc GPU bring-up script (run under gpurun): parity of filters / maps / features vs the oracle,
plus a rough timing of the cascade per configuration.  Not collected by pytest."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wst_b200
from oracle import Scattering2D as OScat

def metric(a, b):
    tau = 1e-3 * np.abs(b).max(axis=tuple(range(1, b.ndim)), keepdims=True)
    return float((np.abs(a - b) / np.maximum(np.abs(b), tau)).max())

print(torch.cuda.get_device_name(0), wst_b200._lib.load().wst2d_version())
for (M, J, L) in [(32, 2, 8), (64, 3, 8), (128, 4, 8), (128, 2, 8), (32, 3, 6)]:
    t = time.time(); plan = wst_b200.get_plan(M, M, J, L); torch.cuda.synchronize(); tp = time.time() - t
    S = OScat(J=J, shape=(M, M), L=L, precision='double', cache_filters=True)
    psi, phi = plan.filters()
    opsi = np.stack([p['levels'][0] for p in S.psi]); ophi = S.phi['levels'][0]
    print(f"[{M} J={J} L={L}] plan {tp:.3f}s  psi maxdiff {np.abs(psi-opsi).max():.2e} (max {np.abs(opsi).max():.3f})  phi maxdiff {np.abs(phi-ophi).max():.2e}")
    rng = np.random.default_rng(1)
    x = (rng.integers(0, 256, (4, 3, M, M)).astype(np.float32) / 255)
    xd = torch.from_numpy(x).cuda()
    feats, maps = plan.forward(xd, True, True); torch.cuda.synchronize()
    ref = S(x)
    m = maps.cpu().numpy()
    rm, rs = ref.mean(axis=(-2, -1)), ref.std(axis=(-2, -1))
    f = feats.cpu().numpy()
    print(f"   maps floored-rel {metric(m.reshape(12, -1), ref.reshape(12, -1)):.2e}  mean {metric(f[:, :, 0].reshape(12, -1), rm.reshape(12, -1)):.2e}  std {metric(f[:, :, 1].reshape(12, -1), rs.reshape(12, -1)):.2e}  nan {np.isnan(m).sum()}")
    fh = plan.forward_host(x)
    print(f"   host path == device path: {np.array_equal(fh, f)}")
    # timing
    B = 2048 if M <= 64 else 1024
    xb = torch.rand((B, 3, M, M), device='cuda')
    for _ in range(2): plan.forward(xb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(3): plan.forward(xb)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"   B={B}: {ms:.2f} ms/batch  -> {B/ms*1e3:.0f} patches/s")
