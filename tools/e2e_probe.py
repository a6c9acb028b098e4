"""Where does the host path spend its time?  H2D bandwidth, and forward_host vs device forward."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wst_b200
B, C, M, J = 4096, 3, 128, 4
plan = wst_b200.get_plan(M, M, J, 8)
xh = torch.rand(B, C, M, M).pin_memory()
xd = torch.empty_like(xh, device="cuda")
for _ in range(2): xd.copy_(xh, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3): xd.copy_(xh, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
print(f"H2D pinned {xh.numel()*4/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms)")
fh = torch.empty(B, C, 2, plan.K).pin_memory()
for _ in range(2): plan.forward(xd)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3): plan.forward(xd)
torch.cuda.synchronize(); print(f"device forward {(time.perf_counter()-t)/3*1e3:.1f} ms")
plan.forward_host(xh, out=fh)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3): plan.forward_host(xh, out=fh)
print(f"forward_host {(time.perf_counter()-t)/3*1e3:.1f} ms")
