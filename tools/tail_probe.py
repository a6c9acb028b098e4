"""Mid-size batches (a few persistent-grid waves): ms per forward call with the ragged last wave run as a split second
launch (default) and as a whole extra wave (WST_NO_TAIL_SPLIT=1), under gpurun.
usage: python tools/tail_probe.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wst_b200


def ms_per_call(plan, x, n=30):
    for _ in range(5):
        plan.forward(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        plan.forward(x)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = []
for (M, J, B) in [(128, 2, 50), (128, 2, 121), (128, 4, 50), (128, 4, 121), (64, 3, 200)]:
    plan = wst_b200.get_plan(M, M, J, 8)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randint(0, 256, (B, 3, M, M), device="cuda", generator=g, dtype=torch.int32).float().div_(255.0)
    a = ms_per_call(plan, x)
    os.environ["WST_NO_TAIL_SPLIT"] = "1"
    b = ms_per_call(plan, x)
    del os.environ["WST_NO_TAIL_SPLIT"]
    out.append({"patch": M, "J": J, "rgb_patches": B, "signals": 3 * B, "grid": plan.grid, "launches": plan.launch_count(B, 3),
                "ms_tail_split": round(a, 4), "ms_whole_wave": round(b, 4), "speedup": round(b / a, 3)})
    print(json.dumps(out[-1]), flush=True)
