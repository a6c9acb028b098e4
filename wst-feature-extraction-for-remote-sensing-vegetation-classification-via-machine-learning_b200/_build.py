"""Build libwst_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

wst_lib.cu (C ABI, pooling, filter bank) and one translation unit per compiled cascade configuration
(wst_cfg_inst.cu with -DWST_CFG_N/-DWST_CFG_J, list in csrc/wst_configs.inc) are compiled in parallel and
linked into one shared library."""
import concurrent.futures
import os
import re
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
import hashlib

# Tuning builds (WST_BUILD_DEFS="-DWST_OPT_X=0 ..." with WST_BUILD_LIB naming the output) keep their objects apart
_DEFS = os.environ.get("WST_BUILD_DEFS", "").split()
OBJ = os.path.join(PKG_DIR, "_obj", hashlib.sha1(" ".join(_DEFS).encode()).hexdigest()[:8] if _DEFS else "default")
LIB_PATH = os.path.join(PKG_DIR, os.environ.get("WST_BUILD_LIB", "libwst_b200.so"))
NVCC_FLAGS = _DEFS + ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC"]
# The global-workspace variant (CFGG entries: sides whose arrays do not fit in shared memory): threads per CTA (two
# CTAs share an SM), CTAs per cluster (1: one CTA per signal; the cluster form is kept as a tuning knob, it measured
# slower), and the data-region budget in cfloats that sizes the per-signal working set (Cfg::BUDGET).
# The environment overrides exist for tuning runs: WST_BUILD_NTL / WST_BUILD_CL / WST_BUILD_BUDGET, and WST_BUILD_LIB
# names the output library (the same variable selects the library at run time, see tools/tune_variants.sh).
GLOBAL_VARIANT_THREADS = int(os.environ.get("WST_BUILD_NTL", 256))
GLOBAL_VARIANT_CLUSTER = int(os.environ.get("WST_BUILD_CL", 1))
GLOBAL_VARIANT_BUDGET = int(os.environ.get("WST_BUILD_BUDGET", 1 << 20))
# Hybrid form: cfloats of shared memory in which the levels that fit are processed like the shared-memory cascade (0: none)
GLOBAL_VARIANT_HYBRID = int(os.environ.get("WST_BUILD_HYBRID", 0))


# Shared-memory configurations that run with other than the default 640 threads / 27000-cfloat data region:
# {(N, J): (threads per CTA, data-region budget in cfloats)}; small sides leave room for several CTAs per SM, whose
# barrier waits then overlap.  WST_BUILD_SHARED="N:J:threads:budget,..." overrides for tuning runs.
SHARED_OVERRIDES = {(40, 2): (160, 6000), (80, 3): (320, 12000)}     # four / two narrower CTAs per SM (measured best)
# Per-configuration tuning knobs of csrc/wst_cascade.h that measured better for that configuration only
# (the sparse / dense-many products with hoisted spectrum rows: +2.4 % at 64x64 J=3, -1.2 % at 128x128 J=4)
# (the filter product fused into the column tile of the inverse transform: +1.9 % at 256x256 J=4, where only level 0 lives
# in the workspace; -1.8 % at 512x512 J=5, whose 288^2 / 144^2 children get four- and eight-column tiles)
CONFIG_DEFS = {(80, 3): ["-DWST_OPT_SPARSEROW=7"],
               (264, 2): ["-DWST_OPT_PRODTILE=1"], (272, 3): ["-DWST_OPT_PRODTILE=1"], (288, 4): ["-DWST_OPT_PRODTILE=1"],
               (320, 5): ["-DWST_OPT_PRODTILE=1"]}
for _item in filter(None, os.environ.get("WST_BUILD_SHARED", "").split(",")):
    _n, _j, _t, _b = (int(v) for v in _item.split(":"))
    SHARED_OVERRIDES[(_n, _j)] = (_t, _b)


# Global-workspace configurations built with other than GLOBAL_VARIANT_THREADS / no hybrid region:
# {(N, J): (threads per CTA, hybrid shared-memory budget in cfloats, workspace budget in cfloats)} — from measurements
GLOBAL_OVERRIDES = {
    # 256 x 256: one 768-thread CTA per SM, levels <= 144^2 in a 216 KB shared-memory region, and a 1 MB workspace per CTA
    # (one level-0 array at a time: the 148 workspaces stay in L2).  5.0 / 5.1 / 5.8 k -> 6.5 / 7.8 / 8.6 k patches/s at J=2/3/4.
    (264, 2): (768, 27000, 1 << 17), (272, 3): (768, 27000, 1 << 17), (288, 4): (768, 27000, 1 << 17),
    (320, 5): (640, 27000, 1 << 17), (240, 3): (768, 27000, 1 << 17),
    # 512 x 512 J=5: two 256-thread CTAs per SM, levels <= 72^2 in a 100 KB region each (+5 %; one wide CTA loses 10-16 %
    # on the 576^2 and 288^2 levels, which stay in the workspace and do not fit L2 at any budget)
    (576, 5): (256, 12500, 1 << 20),
}


def configs():
    """[(N, J, global_workspace)] from csrc/wst_configs.inc."""
    out = []
    only = os.environ.get("WST_BUILD_ONLY")          # tuning builds: "160:4,80:3" compiles those cascades only
    for line in open(os.path.join(CSRC, "wst_configs.inc")):
        m = re.match(r"\s*(CFGG?)\((\d+),\s*(\d+)\)", line)
        if m and (not only or "%s:%s" % (m.group(2), m.group(3)) in only.split(",")):
            out.append((int(m.group(2)), int(m.group(3)), m.group(1) == "CFGG"))
    return out


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(os.path.dirname(PKG_DIR), "include", "wst2d.h"))
    d.append(os.path.abspath(__file__))
    return d


def _stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(args):
    cmd, verbose = args
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def build_library(force=False, verbose=False):
    """Compile csrc/ -> libwst_b200.so.  Returns the library path."""
    if not force and not _stale(LIB_PATH):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    if os.environ.get("WST_BUILD_ONLY"):             # filtered copy of the configuration list for wst_lib.cu / wst_ops.h
        inc = os.path.join(OBJ, "wst_configs_only.inc")
        lines = ["%s(%d, %d)\n" % ("CFGG" if g else "CFG", n, j) for n, j, g in configs()]
        if not os.path.exists(inc) or open(inc).read() != "".join(lines):
            open(inc, "w").write("".join(lines))
        extra += ['-DWST_CONFIGS_FILE="%s"' % inc]
    jobs, objs = [], []
    o = os.path.join(OBJ, "wst_lib.o")
    objs.append(o)
    jobs.append(([nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, "wst_lib.cu"), "-o", o], verbose))
    for unit in ("wst_advstats", "wst_noise", "wst_generic"):
        o = os.path.join(OBJ, unit + ".o")
        objs.append(o)
        jobs.append(([nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, unit + ".cu"), "-o", o], verbose))
    for n, j, glob in configs():
        o = os.path.join(OBJ, "wst_cfg_%d_%d.o" % (n, j))
        defs = ["-DWST_CFG_N=%d" % n, "-DWST_CFG_J=%d" % j]
        if glob:
            cl, ntl, budget, hyb = GLOBAL_VARIANT_CLUSTER, GLOBAL_VARIANT_THREADS, GLOBAL_VARIANT_BUDGET, GLOBAL_VARIANT_HYBRID
            if (n, j) in GLOBAL_OVERRIDES and not any(k in os.environ for k in ("WST_BUILD_NTL", "WST_BUILD_HYBRID", "WST_BUILD_BUDGET")):
                ntl, hyb, budget = GLOBAL_OVERRIDES[(n, j)]
            o = os.path.join(OBJ, "wst_cfg_%d_%d_cl%d_t%d_b%d_h%d.o" % (n, j, cl, ntl, budget, hyb))
            defs += ["-DWST_CFG_GLOBAL=1", "-DWST_CFG_NT=%d" % (ntl * cl), "-DWST_CFG_CL=%d" % cl,
                     "-DWST_GLOBAL_BUDGET=%d" % budget, "-DWST_HYBRID_BUDGET=%d" % hyb]
        if (n, j) in SHARED_OVERRIDES:
            nt, budget = SHARED_OVERRIDES[(n, j)]
            o = os.path.join(OBJ, "wst_cfg_%d_%d_t%d_b%d.o" % (n, j, nt, budget))
            defs += ["-DWST_CFG_NT=%d" % nt, "-DWST_SMEM_BUDGET=%d" % budget]
        if (n, j) in CONFIG_DEFS and not _DEFS:
            defs += CONFIG_DEFS[(n, j)]
        objs.append(o)
        jobs.append(([nvcc] + NVCC_FLAGS + extra + defs + ["-c", os.path.join(CSRC, "wst_cfg_inst.cu"), "-o", o], verbose))
    if not force:                       # objects carry their variant in the name: recompile only what is out of date
        jobs = [jb for jb in jobs if _stale(jb[0][-1])]
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as ex:
        logs = list(ex.map(_compile, jobs))
    if verbose:
        print("\n".join(logs))
    _compile(([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB_PATH], verbose))
    return LIB_PATH
