"""Per-phase SM-cycle breakdown of the cascade kernel (debug executor, clock64 per barrier-separated phase).
Run under gpurun:  python tools/phase_profile.py [cfg3] > gpurun_out/phases.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wst_b200

CFG = {"cfg1": (32, 2), "cfg2": (64, 3), "cfg3": (128, 4), "repo": (128, 2), "cfg5": (512, 5), "p256j2": (256, 2), "p256j3": (256, 3), "p256j4": (256, 4)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
M, J = CFG[name]
plan = wst_b200.get_plan(M, M, J, 8)
nsig_per_cta = 4 if M < 256 else 1
# the profiling twin runs one CTA per SM (launch bound 1), the production kernel plan.grid CTAs: a batch that gives CTA 0
# exactly nsig_per_cta signals on a grid of plan.grid CTAs
B = plan.grid * nsig_per_cta // 3
x = torch.rand(B, 3, M, M, device="cuda")
plan.phase_cycles(x)
cyc = plan.phase_cycles(x)
nsig0 = len(range(0, B * 3, min(B * 3, plan.grid)))        # signals CTA 0 processed
tot = sum(cyc.values())
print(f"{name}: CTA 0 processed {nsig0} signals, {tot / nsig0:.0f} cycles/signal")
bykind, bylevel = {}, {}
for (k, l), c in sorted(cyc.items(), key=lambda kv: -kv[1]):
    print(f"  {k:12s} level {l}: {c / nsig0:10.0f} cyc/signal  {100 * c / tot:5.1f}%")
    bykind[k] = bykind.get(k, 0) + c
    bylevel[l] = bylevel.get(l, 0) + c
print("by kind:")
for k, c in sorted(bykind.items(), key=lambda kv: -kv[1]):
    print(f"  {k:12s} {c / nsig0:10.0f}  {100 * c / tot:5.1f}%")
print("by level:", {l: f"{100 * c / tot:.1f}%" for l, c in sorted(bylevel.items())})
