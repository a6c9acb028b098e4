#!/usr/bin/env python
"""bench.py — WST patches/s on B200 (BASELINE.json metric) for the fused sm_100a Scattering2D path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg3]

A "step" is one pass of the hot path (Scattering2D(J, L=8, max_order=2) + mean/std pooling) over one
batch of synthetic RGB patches per GPU.  Default workload: BASELINE.json configs[2] — 128x128, J=4, the
configuration the metric ("WST patches/sec at 1/2/4/8 B200") and the north-star target are quoted on —
with a fixed per-GPU batch (weak scaling; the batch axis is sharded, no data-path collective).

One JSON line on stdout (rank 0):
  value        whole-job patches/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the public host API (pinned host buffers, H2D + D2H inside the timed region)
  roofline     dominant kernel (cascade) vs the HBM roofline: algorithmic bytes (SURVEY.md 8d bytes_min) per
               launch / live CUDA-event duration of that kernel, against MEASURED_PEAKS.json
  fp32         the binding roofline of this path: model flops / measured FMA peak
  cpu_baseline the oracle (NumPy/SciPy restatement of the reference's kymatio path) on the host cores,
               bounded sample, rank 0 at N=1
--impl reference times that CPU implementation alone on the same config/metric (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {   # name -> (M, J, L, max_order, C, default per-GPU batch)
    "cfg1": (32, 2, 8, 2, 3, 32768),
    "cfg2": (64, 3, 8, 2, 3, 16384),
    "cfg3": (128, 4, 8, 2, 3, 4096),
    "repo": (128, 2, 8, 2, 3, 4096),
    "cfg4": (64, 3, 8, 2, 3, 8192),
    "cfg5": (512, 5, 8, 2, 4, 148),
    "g100": (100, 2, 8, 2, 3, 2048),
    "p256j2": (256, 2, 8, 2, 3, 592),
    "p256j4": (256, 4, 8, 2, 3, 592),
    "p256j5": (256, 5, 8, 2, 3, 592),
    "p96j2": (96, 2, 8, 2, 3, 8192),
}
WORKLOAD_NAMES = {
    "cfg1": "Scattering2D J=2 L=8 max_order=2, 32x32 RGB patches (BASELINE configs[0])",
    "cfg2": "Scattering2D J=3 L=8 max_order=2, 64x64 RGB patches (BASELINE configs[1])",
    "cfg3": "Scattering2D J=4 L=8 max_order=2, 128x128 RGB patches, batch sharded across GPUs (BASELINE configs[2])",
    "repo": "Scattering2D J=2 L=8 max_order=2, 128x128 RGB patches (the reference's own setting)",
    "cfg4": "noise-robustness sweep: gaussian / salt_and_pepper / speckle / poisson / uniform noise on uint8 64x64 RGB patches, "
            "each followed by Scattering2D J=3 L=8 max_order=2 from the uint8 pixels (BASELINE configs[3]); a step = the five "
            "models over the batch, patches/s counts every noised patch",
    "cfg5": "whole-scene tiling: Scattering2D J=5 L=8 max_order=2 over the 512x512 tiles of a 4-band raster, tiles read in place "
            "by the cascade's input stage (BASELINE configs[4]; hybrid global-workspace cascade), tiles/s",
    "g100": "Scattering2D J=2 L=8 max_order=2, 100x100 RGB patches (not a BASELINE shape: padded side 108 = 4*27 has no "
            "compiled cascade, DFT-matrix engine)",
    "p256j2": "Scattering2D J=2 L=8 max_order=2, 256x256 RGB patches (not a BASELINE shape; global-workspace cascade)",
    "p256j4": "Scattering2D J=4 L=8 max_order=2, 256x256 RGB patches (not a BASELINE shape; global-workspace cascade)",
    "p256j5": "Scattering2D J=5 L=8 max_order=2, 256x256 RGB patches (not a BASELINE shape; hybrid global-workspace cascade)",
    "p96j2": "Scattering2D J=2 L=8 max_order=2, 96x96 RGB patches (not a BASELINE shape; shared-memory cascade, padded side 104 = 8*13)",
}


def num_coefficients(J, L, mo):
    return 1 + L * J + (L * L * J * (J - 1) // 2 if mo >= 2 else 0)


def flops_model(M, J, L, C):
    """SURVEY.md 8(d) flops_alg per patch: 10 n^2 log2 n per complex 2-D FFT + 2 flops per filter multiply."""
    import math
    N = ((M + 2 ** J) // 2 ** J + 1) * 2 ** J
    n = [N >> j for j in range(J + 1)]
    count = [1 + 2 * L] + [2 * L + 2 * L * L * j for j in range(1, J)] + [num_coefficients(J, L, 2)]
    fft = sum(c * 10 * m * m * math.log2(m) for c, m in zip(count, n))
    pm = n[0] ** 2
    for j1 in range(J):
        pm += L * n[0] ** 2 + L * n[j1] ** 2
        for j2 in range(j1 + 1, J):
            pm += L * L * (n[j1] ** 2 + n[j2] ** 2)
    return C * (fft + 2 * pm)


# ----------------------------------------------------------------------------- CPU baseline (oracle)
_CPU_STATE = {}


SWEEP = [("gaussian", 30), ("salt_and_pepper", 15), ("speckle", 35), ("poisson", 40), ("uniform", 25)]


def _cpu_worker(args):
    seed, n, M, J, L, mo, C = args[:7]
    sweep = len(args) > 7 and args[7]
    import numpy as np
    from oracle import extract_wst_features_training
    rng = np.random.default_rng(seed)
    if sweep:        # BASELINE configs[3]: the reference's noise models (add_noise.py:14-72) on uint8 pixels, then the extractor
        from oracle.add_noise import add_noise
        np.random.seed(42)
        u8 = rng.integers(0, 256, (n, M, M, C)).astype(np.uint8)
        done = 0
        for b in range(n):
            for kind, inten in SWEEP:
                noisy = add_noise(kind, u8[b], inten)
                chw = np.ascontiguousarray(np.transpose(noisy.astype(np.float32) / 255.0, (2, 0, 1)))
                extract_wst_features_training(chw, J=J, L=L, max_order=mo, cache_filters=True)
                done += 1
        return done
    x = (rng.integers(0, 256, (n, C, M, M)) / 255.0).astype(np.float32)
    for b in range(n):
        extract_wst_features_training(x[b], J=J, L=L, max_order=mo, cache_filters=True)
    return n


def cpu_all_cores(M, J, L, mo, C, per_core, sweep=False):
    """Amortised filter bank, one process per host core (fork; the bank is built once in the parent)."""
    import multiprocessing as mp
    from oracle import Scattering2D
    Scattering2D(J=J, shape=(M, M), L=L, max_order=mo, cache_filters=True)       # warm the shared cache pre-fork
    cores = len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(1000 + i, 1, M, J, L, mo, C, sweep) for i in range(cores)])   # warm-up
        t0 = time.perf_counter()
        done = sum(pool.map(_cpu_worker, [(i, per_core, M, J, L, mo, C, sweep) for i in range(cores)]))
        dt = time.perf_counter() - t0
    return done / dt, cores, done


def cpu_baseline(M, J, L, mo, C, sweep=False):
    """Three CPU numbers: A as-called (bank rebuilt per image, 1 thread), B amortised
    (1 thread), C amortised on all cores.  `value` is C, the strongest CPU arm."""
    import numpy as np
    from oracle import extract_wst_features_training
    rng = np.random.default_rng(0)
    big = M >= 128
    nA, nB = (1, 1) if M >= 512 else ((2, 8) if big else (4, 32))
    x = (rng.integers(0, 256, (max(nA, nB), C, M, M)) / 255.0).astype(np.float32)
    t0 = time.perf_counter()
    for b in range(nA):
        extract_wst_features_training(x[b], J=J, L=L, max_order=mo, cache_filters=False)
    a = nA / (time.perf_counter() - t0)
    extract_wst_features_training(x[0], J=J, L=L, max_order=mo, cache_filters=True)
    t0 = time.perf_counter()
    for b in range(nB):
        extract_wst_features_training(x[b], J=J, L=L, max_order=mo, cache_filters=True)
    bb = nB / (time.perf_counter() - t0)
    per_core = 1 if M >= 512 else (2 if big else (2 if sweep else 8))
    c, cores, done = cpu_all_cores(M, J, L, mo, C, per_core, sweep)
    return {
        "value": round(c, 3), "unit": "patches/s", "cores": cores, "kind": "port",
        "sample": "%d patches (%d per core) of the same workload%s, filter bank amortised, one process per core; "
                  "oracle = NumPy/SciPy restatement of kymatio 0.3.0 (reference engine not installable)" % (
                      done, done // cores, " (each source patch through the five noise models of add_noise.py, then the "
                      "extractor)" if sweep else ""),
        "oracle_engine": oracle_engine(),
        "as_called_1thread": round(a, 4), "as_called_sample": "%d patches, filter bank rebuilt per image "
        "(train_and_save_model.py:359)" % nA,
        "amortised_1thread": round(bb, 3), "amortised_sample": "%d patches" % nB,
    }


def oracle_engine():
    """the reference's real engine if this box has it, else the NumPy restatement."""
    try:
        import kymatio  # noqa: F401
        return "kymatio"
    except Exception:
        return "port"


def parity_sample(plan, M, J, L, mo, C, dev):
    """Per-order parity of the CUDA path against the float64 oracle on two seeded patches (tests/parity.py metrics)."""
    import numpy as np
    import torch
    from tests.parity import parity_report
    if oracle_engine() == "kymatio":
        from kymatio.numpy import Scattering2D as Ref
        ref_t = Ref(J=J, shape=(M, M), L=L, max_order=mo)
    else:
        from oracle import Scattering2D as Ref
        ref_t = Ref(J=J, shape=(M, M), L=L, max_order=mo, precision="double", cache_filters=True)
    rng = np.random.default_rng(42)
    n = 1 if M >= 256 else 2
    x = (rng.integers(0, 256, (n, C, M, M)) / 255.0).astype(np.float32)
    maps = plan.forward(torch.from_numpy(x).to(dev), False, True)[1].cpu().numpy()
    ref = ref_t(x.astype(np.float64))
    K = ref.shape[-3]
    rep = parity_report(maps.reshape(n * C, K, -1), ref.reshape(n * C, K, -1), J, L, mo)
    return {"oracle_engine": oracle_engine(), "patches": n, "tolerance": 1e-4,
            "per_order": {str(o): {k: float("%.3g" % v) for k, v in r.items()} for o, r in rep.items()},
            "ok": all(r["floored"] <= 1e-4 for r in rep.values())}


def latency_probe(dev):
    """The reference's real call pattern (train_and_save_model.py:486-488): one 3x128x128 image per call, J=2, L=8,
    through the drop-in extract_wst_features (host array in, host array out, plan cached)."""
    import numpy as np
    import torch
    import wst_b200
    rng = np.random.default_rng(1)
    img = (rng.integers(0, 256, (3, 128, 128)) / 255.0).astype(np.float32)

    def med(n=40):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter(); wst_b200.extract_wst_features(img); ts.append(time.perf_counter() - t0)
        ts.sort()
        return 1e3 * ts[len(ts) // 2]
    for _ in range(5):
        wst_b200.extract_wst_features(img)
    split = med()
    os.environ["WST_NO_SPLIT"] = "1"
    try:
        for _ in range(3):
            wst_b200.extract_wst_features(img)
        single = med()
    finally:
        del os.environ["WST_NO_SPLIT"]
    t0 = time.perf_counter()
    from oracle import extract_wst_features_training
    extract_wst_features_training(img, J=2, L=8, cache_filters=False)
    cpu_as_called = 1e3 * (time.perf_counter() - t0)
    return {"call": "extract_wst_features(img[3,128,128]) J=2 L=8, host in / host out, plan cached",
            "ms_per_image": round(split, 4), "ms_per_image_one_cta_per_signal": round(single, 4),
            "cpu_as_called_ms": round(cpu_as_called, 1),
            "note": "3 signals: units of (first-order group, child group) split over CTAs, the last CTA pools; CPU = oracle as the reference calls it "
                    "(filter bank rebuilt per image, train_and_save_model.py:359)"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- arms
def run_reference(args, cfg):
    """The reference's CPU implementation of the path (oracle port), all host cores, rank 0 only."""
    M, J, L, mo, C, _ = cfg
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    big = M >= 128
    sweep = args.config == "cfg4"
    per_core = 1 if (big or sweep) else 4
    rates, cores, done = [], 0, 0
    for i in range(args.warmup + args.steps):
        r, cores, done = cpu_all_cores(M, J, L, mo, C, per_core, sweep)
        if i >= args.warmup:
            rates.append((r, done))
    tot = sum(d for _, d in rates)
    secs = sum(d / r for r, d in rates)
    value = tot / secs
    out = {
        "impl": "reference", "metric": "WST patches/sec", "value": round(value, 3), "unit": "patches/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * secs / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAMES[args.config], "patch": [C, M, M], "J": J, "L": L, "max_order": mo,
                   "step": "%d patches (%d per host core) through the oracle's extract_wst_features%s, "
                           "filter bank amortised" % (done, done // max(cores, 1), " after the oracle's add_noise models" if sweep else ""),
                   "oracle_engine": oracle_engine()},
        "cpu_baseline": {"value": round(value, 3), "unit": "patches/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d patches, one process per core" % (args.steps, done)},
        "e2e": {"value": round(value, 3), "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def run_ours(args, cfg):
    import numpy as np
    import torch
    import torch.distributed as dist
    import wst_b200

    M, J, L, mo, C, default_batch = cfg
    B = args.batch or default_batch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    K = num_coefficients(J, L, mo)
    F = C * 2 * K
    plan = wst_b200.get_plan(M, M, J, L, mo, dev, engine=args.engine)
    sweep = args.config == "cfg4"
    scene = args.config == "cfg5"            # whole-scene tiling: the tiles are read straight from a [C, Himg, Wimg] raster
    nx_tiles = 16
    ny_tiles = (B + nx_tiles - 1) // nx_tiles

    def make_input(r):
        """The synthetic batch of rank r (seed 42 + 1000 r): uint8 HWC pixels for the noise sweep, a multispectral raster of
        ny x 16 tiles for the scene tiler, k/255 float32 CHW patches else."""
        gen = torch.Generator(device=dev).manual_seed(42 + 1000 * r)
        if sweep:
            return torch.randint(0, 256, (B, M, M, C), device=dev, generator=gen, dtype=torch.int32).to(torch.uint8)
        if scene:
            return torch.randint(0, 256, (C, ny_tiles * M, nx_tiles * M), device=dev, generator=gen, dtype=torch.int32).float().div_(255.0)
        return torch.randint(0, 256, (B, C, M, M), device=dev, generator=gen, dtype=torch.int32).float().div_(255.0)

    x = make_input(rank)
    gather = distributed and not args.no_gather
    NOISE = SWEEP
    per_step = B * (len(NOISE) if sweep else 1)           # patches one rank processes per step
    pending = []                                           # (work, gathered matrix) of gathers still in flight

    def compute():
        if scene:
            return plan.forward_scene(x, tile_range=(0, B))[0].view(B, F)
        if not sweep:
            return plan.forward(x)[0].view(B, F)
        outs = [plan.forward(wst_b200.add_noise(x, m, i, seed=42))[0].view(B, F) for m, i in NOISE]
        return torch.cat(outs, 0)

    def step():
        feats = compute()
        if gather:
            # all-gather straight into the [B_total, F] matrix on NCCL's stream; the next step's kernels overlap it
            while len(pending) >= 2:
                pending.pop(0)[0].wait()
            out, work = wst_b200.gather_features(feats, feats.shape[0] * world, async_op=True)
            pending.append((work, out, feats))
        return feats

    def drain():
        last = None
        while pending:
            w, last, _ = pending.pop(0)
            w.wait()
        return last

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    plan.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        feats = step()
    gathered = drain()                                     # every gather has landed before the clock stops
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    cas_ms, pool_ms, cas_n = plan.profile_read()
    plan.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * per_step * args.steps / (ms * 1e-3)

    # ---- the gathered matrix is in input order: rank 0 recomputes the first rows of every rank's shard
    gather_verified = None
    if gather and rank == 0:
        nchk = min(8, B)
        ok = True
        rows = feats.shape[0]
        for r in range(world):
            xr = make_input(r)
            if sweep:
                fr = plan.forward(wst_b200.add_noise(xr, NOISE[0][0], NOISE[0][1], seed=42)[:nchk].contiguous())[0].view(nchk, F)
            elif scene:
                fr = plan.forward_scene(xr, tile_range=(0, nchk))[0].view(nchk, F)
            else:
                fr = plan.forward(xr[:nchk].contiguous())[0].view(nchk, F)
            ok = ok and bool(torch.equal(gathered[r * rows:r * rows + nchk], fr))
            del xr
        gather_verified = ok
    if distributed:
        dist.barrier()

    # ---- e2e: host buffers through the public API, copies inside the timed region
    e2e_steps = max(1, min(args.steps, 5))
    if sweep:
        xh = torch.empty((B, M, M, C), dtype=torch.uint8).pin_memory()
        xh.copy_(x)
        fh = torch.empty((len(NOISE) * B, C, 2, K), dtype=torch.float32).pin_memory()

        def e2e_step():       # clean uint8 pixels from the host, noise + features on the device, features back to the host
            xd = xh.to(dev, non_blocking=True)
            for i, (m, inten) in enumerate(NOISE):
                f = plan.forward(wst_b200.add_noise(xd, m, inten, seed=42))[0]
                fh[i * B:(i + 1) * B].copy_(f, non_blocking=True)
            torch.cuda.synchronize()
        h2d, d2h = B * M * M * C, len(NOISE) * B * F * 4
        api = "torch H2D of the clean uint8 batch, wst_b200.add_noise + Plan.forward(uint8) per model, D2H of the features"
    elif scene:
        xh = torch.empty(tuple(x.shape), dtype=torch.float32).pin_memory()
        xh.copy_(x)
        fh = torch.empty((B, C, 2, K), dtype=torch.float32).pin_memory()

        def e2e_step():       # the raster from pinned host memory, tiled on the device, per-tile features back to the host
            f = plan.forward_scene(xh.to(dev, non_blocking=True), tile_range=(0, B))[0]
            fh.copy_(f, non_blocking=True)
            torch.cuda.synchronize()
        h2d, d2h = x.numel() * 4, B * F * 4
        api = "torch H2D of the [C, Himg, Wimg] raster, Plan.forward_scene (wst2d_forward_scene: tiles read in place), D2H of the features"
    else:
        xh = torch.empty((B, C, M, M), dtype=torch.float32).pin_memory()
        xh.copy_(x)
        fh = torch.empty((B, C, 2, K), dtype=torch.float32).pin_memory()

        def e2e_step():
            plan.forward_host(xh, out=fh)        # synchronous: H2D, kernels, D2H
        h2d, d2h = B * C * M * M * 4, B * F * 4
        api = "Plan.forward_host (wst2d_forward_host): pinned host in/out, chunked double-buffered copies"
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * per_step * e2e_steps / float(t.item())
    same = bool(torch.equal(fh.to(dev).view(-1, F), feats))

    # ---- literal BASELINE configs[2]: --stream-total patches streamed through the host API from uint8 pixels
    stream = None
    if args.stream_total and not sweep:
        chunk = min(args.stream_total, 16384)
        gen = torch.Generator(device="cpu").manual_seed(7 + rank)
        xu = torch.randint(0, 256, (chunk, M, M, C), generator=gen, dtype=torch.uint8).pin_memory()
        fo = torch.empty((chunk, C, 2, K), dtype=torch.float32).pin_memory()
        plan.forward_host(xu[:min(chunk, 1024)], out=fo[:min(chunk, 1024)])
        barrier()
        done, t0 = 0, time.perf_counter()
        while done < args.stream_total:
            n = min(chunk, args.stream_total - done)
            plan.forward_host(xu[:n], out=fo[:n])
            done += n
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if distributed:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        stream = {"patches_per_gpu": done, "patches_total": done * world, "seconds": round(float(tt.item()), 3),
                  "value": round(done * world / float(tt.item()), 2), "unit": "patches/s",
                  "api": "Plan.forward_host on uint8 [n, H, W, C] host chunks of %d patches (wst2d_forward_host_u8): H2D of the "
                         "pixels, /255 + transpose + cascade on the device, D2H of the features" % chunk,
                  "h2d_bytes": done * M * M * C, "d2h_bytes": done * F * 4}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        else:
            hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bytes_min = (C * M * M if sweep else 4 * C * M * M) + 4 * C * 2 * K
        launches_per_step = cas_n / args.steps if args.steps else 0
        avg_launch_ms = cas_ms / cas_n if cas_n else float("nan")
        patches_per_launch = per_step / launches_per_step if launches_per_step else 0
        achieved = bytes_min * patches_per_launch / (avg_launch_ms * 1e-3) / 1e9
        # DRAM bytes of one launch, from the committed ncu --set full capture of this kernel (profiles/): the kernel is
        # persistent and its traffic is per signal, so a capture of a different batch is scaled to this launch's batch
        traffic, traffic_note = None, None
        prof_json = os.path.join(ROOT, "profiles", "ncu_cascade_%s.json" % args.config)
        if os.path.exists(prof_json) and patches_per_launch and plan.engine == "fft":
            pj = json.load(open(prof_json))
            traffic = pj["dram_bytes_per_launch"] * patches_per_launch / pj["patches_per_launch"]
            traffic_note = "%s: %d patches captured, scaled to %d" % (pj.get("source", prof_json), pj["patches_per_launch"], patches_per_launch)
        dram_gbs = traffic / (avg_launch_ms * 1e-3) / 1e9 if traffic else None
        on_chip = traffic is not None and traffic < 2 * bytes_min * patches_per_launch
        note = ("HBM sees the input and the features once (traffic ~ algorithmic bytes); the fused path is fp32 / shared-memory "
                "bound (SURVEY.md F3), see fp32" if on_chip or traffic is None else
                "data region of this side lives in a global workspace: the kernel streams it through HBM (dram_achieved), "
                "traffic >> algorithmic bytes")
        if plan.engine != "fft":
            note = "DFT-matrix engine: one GEMM kernel per transform step, intermediates in an HBM/L2 workspace"
        fl = flops_model(M, J, L, C)
        try:
            fma_peak = wst_b200.fma_peak_tflops(local)
        except Exception:
            fma_peak = None
        n_launch = plan.launch_count(B, C) * (len(NOISE) if sweep else 1) + (2 * len(NOISE) if sweep else 0)
        out = {
            "metric": "WST patches/sec", "value": round(value, 2), "unit": "patches/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAMES[args.config], "patch": [C, M, M], "J": J, "L": L, "max_order": mo,
                       "batch_per_gpu": B, "global_batch": B * world, "features_per_patch": F, "engine": plan.engine,
                       "parallelism": "batch sharded x%d%s" % (world, ", NCCL all-gather of the feature matrix overlapped "
                                                               "with the next step" if gather else ""),
                       "l2": "inputs (%d MB per step) larger than L2; no flush needed" % (x.numel() * x.element_size() // 2 ** 20)
                             if x.numel() * x.element_size() > 126 * 2 ** 20 else
                             "inputs are %d MB per step: re-read from L2 across steps (the kernel is compute-bound: "
                             "HBM traffic is %.2f%% of its time at peak bandwidth)" % (
                                 x.numel() * x.element_size() // 2 ** 20,
                                 100 * (x.numel() * x.element_size() / (hbm_peak * 1e9)) / (ms / args.steps * 1e-3)),
                       "input_values": "k/255, k uniform in 0..255, generated on device, seed 42+1000*rank"},
            "e2e": {"value": round(e2e, 2), "unit": "patches/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "api": api, "matches_device_path": same},
            "gpu_launches": args.steps * n_launch,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 3), "peak": hbm_peak, "unit": "GB/s",
                         "frac": round(achieved / hbm_peak, 6), "traffic": traffic, "peak_source": peak_src,
                         "kernel": "cascade_kernel" if plan.engine == "fft" else "gemm_kernel (all launches of a forward call)",
                         "kernel_ms_per_launch": round(avg_launch_ms, 4),
                         "kernel_share_of_step": round(cas_ms / ms, 4) if ms else None,
                         "algorithmic_bytes_per_patch": bytes_min, "traffic_source": traffic_note,
                         "dram_achieved": round(dram_gbs, 1) if dram_gbs else None,
                         "dram_frac": round(dram_gbs / hbm_peak, 4) if dram_gbs else None, "note": note},
            "fp32": {"model_flops_per_patch": fl, "achieved_tflops": round(value / world * fl / 1e12, 3),
                     "peak_tflops": round(fma_peak, 2) if fma_peak else None,
                     "frac": round(value / world * fl / 1e12 / fma_peak, 4) if fma_peak else None,
                     "peak_source": "measured live: wst2d_fma_peak (FMA loop, all SMs)"},
            "clocks": clocks,
        }
        if gather:
            out["gather_verified"] = gather_verified
        if stream:
            out["stream"] = stream
        if not args.no_parity:
            try:
                out["parity"] = parity_sample(plan, M, J, L, mo, C, dev)
            except Exception as e:          # the bench line must not depend on the checker
                out["parity"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(M, J, L, mo, C, sweep)
            if not args.no_latency:
                out["latency"] = latency_probe(dev)
        print(json.dumps(out), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="patches per GPU per step (default per config)")
    ap.add_argument("--no-gather", action="store_true", help="skip the NCCL feature all-gather at N>1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-image latency probe")
    ap.add_argument("--no-parity", action="store_true", help="skip the per-order parity sample against the oracle")
    ap.add_argument("--engine", default="auto", choices=["auto", "fft", "gemm", "gemm_tf32x3"],
                    help="fused FFT cascade (auto where compiled) or the DFT-matrix engine on the fp32 / tensor pipe")
    ap.add_argument("--stream-total", type=int, default=0,
                    help="additionally stream this many patches per GPU through the uint8 host API (BASELINE configs[2]: 1000000 / N)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
