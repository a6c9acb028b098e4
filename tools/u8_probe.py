"""float32 vs uint8 ingest throughput of the cascade (same pixels), per config.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wst_b200
for M, J, B in ((64, 3, 8192), (128, 4, 4096), (32, 2, 32768)):
    plan = wst_b200.get_plan(M, M, J)
    u8 = torch.randint(0, 256, (B, M, M, 3), device="cuda", dtype=torch.uint8)
    f32 = (u8.float() / 255.0).permute(0, 3, 1, 2).contiguous()
    for name, x in (("f32", f32), ("u8", u8)):
        for _ in range(3):
            plan.forward(x)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            plan.forward(x)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print("%dx%d J=%d %s: %.0f patches/s" % (M, M, J, name, B / dt), flush=True)
