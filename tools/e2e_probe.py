"""Where does the host path spend its time?  Device forward vs forward_host at several chunk sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wst_b200
B, C, M, J = 4096, 3, 128, 4
plan = wst_b200.get_plan(M, M, J, 8)
x = torch.rand(B, C, M, M, device="cuda")
for _ in range(3): plan.forward(x)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3): plan.forward(x)
torch.cuda.synchronize(); print(f"device forward {(time.perf_counter()-t)/3*1e3:.1f} ms")
xh = torch.empty(B, C, M, M).pin_memory(); xh.copy_(x)
fh = torch.empty(B, C, 2, plan.K).pin_memory()
def run(tag):
    plan.forward_host(xh, out=fh)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): plan.forward_host(xh, out=fh)
    print(f"forward_host [{tag}] {(time.perf_counter()-t)/3*1e3:.1f} ms")
run("default")
for ch in (4, 2, 8, 12, 24):
    os.environ["WST_HOST_CHUNK_SIGNALS"] = str(148 * ch); run(f"chunk {ch} waves")
