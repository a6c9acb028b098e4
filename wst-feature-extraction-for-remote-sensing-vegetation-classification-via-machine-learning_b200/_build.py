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
OBJ = os.path.join(PKG_DIR, "_obj")
LIB_PATH = os.path.join(PKG_DIR, "libwst_b200.so")
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC"]
GLOBAL_VARIANT_THREADS = 256      # the global-workspace variant holds 2 x 24-point butterflies per thread


def configs():
    """[(N, J, global_workspace)] from csrc/wst_configs.inc."""
    out = []
    for line in open(os.path.join(CSRC, "wst_configs.inc")):
        m = re.match(r"\s*(CFGG?)\((\d+),\s*(\d+)\)", line)
        if m:
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
    jobs, objs = [], []
    o = os.path.join(OBJ, "wst_lib.o")
    objs.append(o)
    jobs.append(([nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, "wst_lib.cu"), "-o", o], verbose))
    for unit in ("wst_advstats", "wst_noise"):
        o = os.path.join(OBJ, unit + ".o")
        objs.append(o)
        jobs.append(([nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, unit + ".cu"), "-o", o], verbose))
    for n, j, glob in configs():
        o = os.path.join(OBJ, "wst_cfg_%d_%d.o" % (n, j))
        objs.append(o)
        defs = ["-DWST_CFG_N=%d" % n, "-DWST_CFG_J=%d" % j]
        if glob:
            defs += ["-DWST_CFG_GLOBAL=1", "-DWST_CFG_NT=%d" % GLOBAL_VARIANT_THREADS]
        jobs.append(([nvcc] + NVCC_FLAGS + extra + defs + ["-c", os.path.join(CSRC, "wst_cfg_inst.cu"), "-o", o], verbose))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        logs = list(ex.map(_compile, jobs))
    if verbose:
        print("\n".join(logs))
    _compile(([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB_PATH], verbose))
    return LIB_PATH
