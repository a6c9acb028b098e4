"""Throughput of the two side kernels (advanced statistics, noise models) on resident data, CUDA events.
Run under gpurun:  python tools/aux_bench.py > gpurun_out/aux_bench.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wst_b200

def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = {}
B = 8192
u8 = torch.randint(0, 256, (B, 128, 128, 3), dtype=torch.uint8, device="cuda")           # 403 MB, larger than L2
f32 = (u8.permute(0, 3, 1, 2).float() / 255.0).contiguous()
ms = timed(lambda: wst_b200.advanced_stats(f32))
out["advanced_stats_f32_128x128x3"] = {"patches_per_s": B / ms * 1e3, "ms": ms, "GBps_in": f32.numel() * 4 / ms / 1e6}
ms = timed(lambda: wst_b200.advanced_stats(u8))
out["advanced_stats_u8_128x128x3"] = {"patches_per_s": B / ms * 1e3, "ms": ms, "GBps_in": u8.numel() / ms / 1e6}
for kind in wst_b200.NOISE_TYPES:
    ms = timed(lambda: wst_b200.add_noise(u8, kind, 25, seed=1))
    out["add_noise_%s_128x128x3" % kind] = {"patches_per_s": B / ms * 1e3, "ms": ms, "GBps_in_plus_out": 2 * u8.numel() / ms / 1e6}
plan = wst_b200.get_plan(128, 128, 2, 8)
ms = timed(lambda: torch.cat([wst_b200.advanced_stats(u8[:4096]).reshape(4096, -1), wst_b200.to_block(plan.forward(u8[:4096])[0])], dim=1), reps=3)
out["hybrid_u8_128x128x3_J2"] = {"patches_per_s": 4096 / ms * 1e3, "ms": ms}
print(json.dumps(out, indent=1))
