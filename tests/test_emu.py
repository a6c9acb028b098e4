"""CPU replay of the CUDA kernels' phases (tests/emu) against the oracle: this is how the index
arithmetic of csrc/wst_cascade.h and the filter-bank formula of csrc/wst_filters.h are checked in the
GPU-less container.  The GPU itself is checked by tests/test_gpu_parity.py."""
import numpy as np
import pytest

from tests import emu
from oracle import Scattering2D, filter_bank


def floored_rel(a, b):
    a = a.reshape(a.shape[0], -1); b = b.reshape(b.shape[0], -1)
    tau = 1e-3 * np.abs(b).max(axis=1, keepdims=True)
    return float((np.abs(a - b) / np.maximum(np.abs(b), tau)).max())


@pytest.mark.parametrize("M", [6, 10, 12, 18, 20, 24, 34, 40, 48, 68, 80, 136, 160])
def test_fft_passes(M):
    rng = np.random.default_rng(M)
    x = (rng.standard_normal((M, M)) + 1j * rng.standard_normal((M, M))).astype(np.complex64)
    ref = np.fft.fft2(x.astype(np.complex128))
    assert np.abs(emu.fft2(x, -1) - ref).max() <= 1e-6 * np.abs(ref).max()
    ref = np.fft.ifft2(x.astype(np.complex128)) * M * M
    assert np.abs(emu.fft2(x, +1) - ref).max() <= 1e-6 * np.abs(ref).max()


@pytest.mark.parametrize("N,J,L", [(40, 2, 8), (48, 3, 6)])
def test_filter_bank_formula(N, J, L):
    psi, phi = emu.filter_bank(N, J, L)
    fb = filter_bank(N, N, J, L)
    opsi = np.stack([p["levels"][0] for p in fb["psi"]])
    assert np.abs(psi - opsi).max() < 1e-6
    assert np.abs(phi - fb["phi"]["levels"][0]).max() < 1e-6


@pytest.mark.parametrize("M,J,L,mo", [(32, 2, 8, 2), (32, 2, 8, 1), (32, 3, 6, 2), (64, 3, 8, 2), (128, 2, 8, 2)])
def test_cascade_vs_oracle(M, J, L, mo):
    rng = np.random.default_rng(7)
    x = (rng.integers(0, 256, (2, M, M)) / 255.0).astype(np.float32)
    S = Scattering2D(J=J, shape=(M, M), L=L, max_order=mo, precision="double", cache_filters=True)
    psi = np.stack([p["levels"][0] for p in S.psi]); phi = S.phi["levels"][0]
    got, feats = emu.forward(x, J, L, mo, psi, phi, with_features=True)
    assert not np.isnan(got).any() and not np.isnan(feats).any()
    ref = S(x)
    assert floored_rel(got, ref) <= 1e-4 / 4
    assert floored_rel(feats[:, 0], ref.mean(axis=(-2, -1))) <= 1e-4 / 4       # in-kernel pooling
    tau = 1e-3 * np.abs(ref.mean(axis=(-2, -1))).max(axis=1, keepdims=True)
    assert float((np.abs(feats[:, 1] - ref.std(axis=(-2, -1))) / np.maximum(ref.std(axis=(-2, -1)), tau)).max()) <= 1e-4 / 4


def test_kernels_stay_in_bounds_under_asan(tmp_path):
    """compute-sanitizer is not available on the GPU pool, so the bounds check is done here: the emulation is
    rebuilt with AddressSanitizer and exact-size shared-memory buffers (data region, twiddles, low-pass tables,
    reduction buffer are separate heap blocks) and the full cascade is replayed in a subprocess."""
    import os
    import subprocess
    import sys
    src = open(os.path.join(emu.HERE, "wst_emu.cpp")).read().replace("C::smem_cfloats() + 64", "C::smem_cfloats()")
    cpp = tmp_path / "wst_emu_asan.cpp"
    cpp.write_text(src)
    (tmp_path / "small_configs.inc").write_text("CFG(40, 2)\nCFG(48, 3)\nCFG(80, 3)\nCFG(36, 1)\nCFG(64, 4)\n")
    lib = tmp_path / "libwst_emu_asan.so"
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer", "-fPIC",
                    "-shared", "-DWST_EMU_CONFIG_FILE=\"small_configs.inc\"", "-I", str(tmp_path), "-I", emu.CSRC,
                    str(cpp), "-o", str(lib)], check=True)
    asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    code = (
        "import ctypes, numpy as np, sys\n"
        "sys.path.insert(0, %r)\n"
        "from oracle import Scattering2D\n"
        "lib = ctypes.CDLL(%r)\n"
        "lib.emu_forward.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]\n"
        "for M, J, L in [(32, 2, 8), (32, 3, 6), (64, 3, 8), (32, 1, 8), (32, 4, 8)]:\n"
        "    S = Scattering2D(J=J, shape=(M, M), L=L, cache_filters=True)\n"
        "    psi = np.ascontiguousarray(np.stack([p['levels'][0] for p in S.psi]), np.float32)\n"
        "    phi = np.ascontiguousarray(S.phi['levels'][0], np.float32)\n"
        "    N = S._M_padded; K = 1 + L * J + L * L * J * (J - 1) // 2; h = N // 2 ** J - 2\n"
        "    x = np.random.default_rng(0).random((1, M, M), dtype=np.float32)\n"
        "    out = np.empty((1, K, h, h), np.float32); f = np.empty((1, 2, K), np.float32)\n"
        "    rc = lib.emu_forward(N, J, L, 2, M, M, psi.ctypes.data, phi.ctypes.data, x.ctypes.data, 1, out.ctypes.data, f.ctypes.data)\n"
        "    assert rc == 0 and np.isfinite(out).all() and np.isfinite(f).all()\n"
        "print('clean')\n" % (emu.ROOT, str(lib)))
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "clean" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_shipped_build_variants_vs_oracle(tmp_path):
    """The small sides are built with fewer threads per CTA and a smaller data region than the defaults
    (wst_b200/_build.py: SHARED_OVERRIDES, several CTAs per SM).  Group sizes and every phase's work split depend on
    both, so the emulation is rebuilt with exactly those parameters and replayed against the oracle."""
    import ctypes
    import importlib
    import subprocess
    overrides = importlib.import_module("wst_b200._build").SHARED_OVERRIDES
    assert overrides, "no build variants to check"
    sizes = {(40, 2): 32, (80, 3): 64, (36, 1): 32, (72, 2): 64, (48, 3): 32, (64, 4): 32, (96, 4): 64}
    for (N, J), (nt, budget) in sorted(overrides.items()):
        M, L = sizes[(N, J)], 8
        inc = tmp_path / ("cfg_%d_%d.inc" % (N, J))
        inc.write_text("CFG(%d, %d)\n" % (N, J))
        lib_path = tmp_path / ("libwst_emu_%d_%d.so" % (N, J))
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-DWST_NT=%d" % nt,
                        "-DWST_SMEM_BUDGET=%d" % budget, "-DWST_EMU_CONFIG_FILE=\"%s\"" % inc.name,
                        "-I", str(tmp_path), "-I", emu.CSRC, "-I", emu.HERE, str(emu.HERE + "/wst_emu.cpp"),
                        "-o", str(lib_path)], check=True)
        lib = ctypes.CDLL(str(lib_path))
        lib.emu_forward.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        rng = np.random.default_rng(N)
        x = (rng.integers(0, 256, (2, M, M)) / 255.0).astype(np.float32)
        S = Scattering2D(J=J, shape=(M, M), L=L, precision="double", cache_filters=True)
        assert S._M_padded == N
        psi = np.ascontiguousarray(np.stack([p["levels"][0] for p in S.psi]), np.float32)
        phi = np.ascontiguousarray(S.phi["levels"][0], np.float32)
        K, h = S(x[:1]).shape[1], N // 2 ** J - 2
        out = np.full((2, K, h, h), np.nan, np.float32)
        feats = np.full((2, 2, K), np.nan, np.float32)
        rc = lib.emu_forward(N, J, L, 2, M, M, psi.ctypes.data, phi.ctypes.data, x.ctypes.data, 2, out.ctypes.data,
                             feats.ctypes.data)
        assert rc == 0
        ref = S(x)
        assert not np.isnan(out).any() and floored_rel(out, ref) <= 1e-4 / 4
        assert floored_rel(feats[:, 0], ref.mean(axis=(-2, -1))) <= 1e-4 / 4


def test_global_workspace_variant_vs_oracle(tmp_path):
    """The global-workspace variant of the cascade (sides whose arrays exceed shared memory: 256x256, 512x512) runs its
    inverse FFTs on staged shared-memory tiles.  Replayed here at small sides, where the oracle is quick: fused and
    unfused low-pass, Cooley-Tukey and Good-Thomas lengths, single-pass levels."""
    import ctypes
    import subprocess
    inc = tmp_path / "glob_configs.inc"
    inc.write_text("CFGG(40, 2)\nCFGG(80, 3)\nCFGG(136, 2)\nCFGG(48, 3)\n")
    lib_path = tmp_path / "libwst_emu_glob.so"
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-DWST_GLOBAL_BUDGET=32768",
                    "-DWST_EMU_CONFIG_FILE=\"%s\"" % inc.name, "-I", str(tmp_path), "-I", emu.CSRC, "-I", emu.HERE,
                    emu.HERE + "/wst_emu.cpp", "-o", str(lib_path)], check=True)
    lib = ctypes.CDLL(str(lib_path))
    lib.emu_forward.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    for M, J, L, N in [(32, 2, 8, 40), (64, 3, 8, 80), (128, 2, 8, 136), (32, 3, 6, 48)]:
        rng = np.random.default_rng(N)
        x = (rng.integers(0, 256, (2, M, M)) / 255.0).astype(np.float32)
        S = Scattering2D(J=J, shape=(M, M), L=L, precision="double", cache_filters=True)
        assert S._M_padded == N
        psi = np.ascontiguousarray(np.stack([p["levels"][0] for p in S.psi]), np.float32)
        phi = np.ascontiguousarray(S.phi["levels"][0], np.float32)
        ref = S(x)
        K, h = ref.shape[1], N // 2 ** J - 2
        out = np.full((2, K, h, h), np.nan, np.float32)
        feats = np.full((2, 2, K), np.nan, np.float32)
        rc = lib.emu_forward(N, J, L, 2, M, M, psi.ctypes.data, phi.ctypes.data, x.ctypes.data, 2, out.ctypes.data,
                             feats.ctypes.data)
        assert rc == 0
        assert not np.isnan(out).any() and floored_rel(out, ref) <= 1e-4 / 4, (M, J)
        assert floored_rel(feats[:, 0], ref.mean(axis=(-2, -1))) <= 1e-4 / 4


def test_hybrid_global_workspace_variant_vs_oracle(tmp_path):
    """Hybrid form of the global-workspace variant: level 0 in the workspace (staged passes), the levels that fit a small
    shared-memory budget processed there like the shared-memory cascade, children of workspace-level parents included.
    Replayed at small sides with budgets that put the split where the 256 x 256 / 512 x 512 builds have it."""
    import ctypes
    import subprocess
    inc = tmp_path / "hyb_configs.inc"
    inc.write_text("CFGG(160, 4)\nCFGG(80, 3)\nCFGG(48, 3)\n")
    lib_path = tmp_path / "libwst_emu_hyb.so"
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-DWST_GLOBAL_BUDGET=65536", "-DWST_HYBRID_BUDGET=7000",
                    "-DWST_EMU_CONFIG_FILE=\"%s\"" % inc.name, "-I", str(tmp_path), "-I", emu.CSRC, "-I", emu.HERE,
                    emu.HERE + "/wst_emu.cpp", "-o", str(lib_path)], check=True)
    lib = ctypes.CDLL(str(lib_path))
    lib.emu_forward.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.emu_query.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3
    for M, J, L, N in [(128, 4, 8, 160), (64, 3, 8, 80), (32, 3, 6, 48)]:
        rng = np.random.default_rng(N + 1)
        x = (rng.integers(0, 256, (1, M, M)) / 255.0).astype(np.float32)
        S = Scattering2D(J=J, shape=(M, M), L=L, precision="double", cache_filters=True)
        assert S._M_padded == N
        psi = np.ascontiguousarray(np.stack([p["levels"][0] for p in S.psi]), np.float32)
        phi = np.ascontiguousarray(S.phi["levels"][0], np.float32)
        ref = S(x)
        K, h = ref.shape[1], N // 2 ** J - 2
        out = np.full((1, K, h, h), np.nan, np.float32)
        feats = np.full((1, 2, K), np.nan, np.float32)
        rc = lib.emu_forward(N, J, L, 2, M, M, psi.ctypes.data, phi.ctypes.data, x.ctypes.data, 1, out.ctypes.data,
                             feats.ctypes.data)
        assert rc == 0
        assert not np.isnan(out).any() and floored_rel(out, ref) <= 1e-4 / 4, (M, J)
        assert floored_rel(feats[:, 0], ref.mean(axis=(-2, -1))) <= 1e-4 / 4


def test_fused_product_tile_vs_oracle(tmp_path):
    """WST_OPT_PRODTILE (the 256 x 256 builds): workspace-level arrays with fold factor <= 2 get their filter product
    computed straight into the shared-memory column tile of the inverse transform; the workspace then holds natural-order
    columns between the column and the row passes.  Replayed at small sides: first-order products without fold (80, 136,
    160) and with fold 2, second-order children with fold 2 (eight- and four-column tiles of four and eight arrays),
    Cooley-Tukey and Good-Thomas lengths (the unfused build of the same configurations is
    test_global_workspace_variant_vs_oracle; the two agree to rounding, profiles/r02_prodtile_ab.txt)."""
    import ctypes
    import subprocess

    def build(tag, cfgs, defs):
        inc = tmp_path / (tag + ".inc")
        inc.write_text("".join("CFGG(%d, %d)\n" % c for c in cfgs))
        lib_path = tmp_path / ("libwst_emu_%s.so" % tag)
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared"] + defs + ["-DWST_EMU_CONFIG_FILE=\"%s\"" % inc.name,
                        "-I", str(tmp_path), "-I", emu.CSRC, "-I", emu.HERE, emu.HERE + "/wst_emu.cpp", "-o", str(lib_path)],
                       check=True)
        lib = ctypes.CDLL(str(lib_path))
        lib.emu_forward.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        return lib

    def run(lib, M, J, L, N):
        rng = np.random.default_rng(N + 7)
        x = (rng.integers(0, 256, (1, M, M)) / 255.0).astype(np.float32)
        S = Scattering2D(J=J, shape=(M, M), L=L, precision="double", cache_filters=True)
        assert S._M_padded == N
        psi = np.ascontiguousarray(np.stack([p["levels"][0] for p in S.psi]), np.float32)
        phi = np.ascontiguousarray(S.phi["levels"][0], np.float32)
        ref = S(x)
        K, h = ref.shape[1], N // 2 ** J - 2
        out = np.full((1, K, h, h), np.nan, np.float32)
        feats = np.full((1, 2, K), np.nan, np.float32)
        rc = lib.emu_forward(N, J, L, 2, M, M, psi.ctypes.data, phi.ctypes.data, x.ctypes.data, 1, out.ctypes.data,
                             feats.ctypes.data)
        assert rc == 0 and not np.isnan(out).any()
        return out, ref

    glob = [(80, 3), (136, 2), (48, 3)]
    fused = build("pt1", glob, ["-DWST_OPT_PRODTILE=1", "-DWST_GLOBAL_BUDGET=32768"])
    for M, J, L, N in [(64, 3, 8, 80), (128, 2, 8, 136), (32, 3, 6, 48)]:
        a, ref = run(fused, M, J, L, N)
        assert floored_rel(a, ref) <= 1e-4 / 4, (M, J)
    hyb = build("pt1h", [(160, 4)], ["-DWST_OPT_PRODTILE=1", "-DWST_GLOBAL_BUDGET=65536", "-DWST_HYBRID_BUDGET=4000"])
    a, ref = run(hyb, 128, 4, 8, 160)
    assert floored_rel(a, ref) <= 1e-4 / 4


def test_shipped_global_workspace_builds_under_asan(tmp_path):
    """The 256 x 256 / 224 x 224 / 512 x 512 cascades exactly as `_build.py` ships them (threads per CTA, hybrid
    shared-memory budget, workspace budget, per-configuration knobs such as the fused product tile), replayed with
    AddressSanitizer and exact-size buffers (workspace, stage / hybrid region, tables are separate heap blocks) against
    the float64 oracle: the bounds check compute-sanitizer would do on the GPU pool, where it is not available."""
    import concurrent.futures
    import importlib
    import os
    import subprocess
    import sys
    b = importlib.import_module("wst_b200._build")
    sizes = {(264, 2): 256, (272, 3): 256, (288, 4): 256, (320, 5): 256, (240, 3): 224, (576, 5): 512}
    src = open(os.path.join(emu.HERE, "wst_emu.cpp")).read().replace("C::smem_cfloats() + 64", "C::smem_cfloats()")
    assert "Cfg<n, j, 256, true>" in src

    def build(item):
        (N, J), (nt, hyb, budget) = item
        cpp = tmp_path / ("emu_%d_%d.cpp" % (N, J))
        cpp.write_text(src.replace("Cfg<n, j, 256, true>", "Cfg<n, j, %d, true>" % nt))
        (tmp_path / ("c_%d_%d.inc" % (N, J))).write_text("CFGG(%d, %d)\n" % (N, J))
        lib = tmp_path / ("libwst_emu_asan_%d_%d.so" % (N, J))
        subprocess.run(["g++", "-std=c++17", "-O1", "-fsanitize=address", "-fno-omit-frame-pointer", "-fPIC", "-shared",
                        "-DWST_GLOBAL_BUDGET=%d" % budget, "-DWST_HYBRID_BUDGET=%d" % hyb] + b.CONFIG_DEFS.get((N, J), []) +
                       ["-DWST_EMU_CONFIG_FILE=\"c_%d_%d.inc\"" % (N, J), "-I", str(tmp_path), "-I", emu.CSRC, "-I", emu.HERE,
                        str(cpp), "-o", str(lib)], check=True)
        return (N, J), lib

    items = sorted(b.GLOBAL_OVERRIDES.items())
    assert set(k for k, _ in items) == set(sizes), "a shipped global-workspace configuration has no emulator case"
    if not os.environ.get("WST_SLOW_TESTS"):     # all six take 3.5 minutes (the sanitised compiles): by default the two
        items = [it for it in items if it[0] in ((288, 4), (320, 5))]   # builds with the widest CTAs and the fused product tile
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(items), os.cpu_count() or 2)) as ex:
        libs = list(ex.map(build, items))
    asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    # the sanitised child only replays the kernel: filter bank, input and reference are prepared here, without ASAN
    code = (
        "import ctypes, numpy as np, sys\n"
        "lib = ctypes.CDLL(sys.argv[1]); d = np.load(sys.argv[2]); N, J, L, M, K, h = (int(v) for v in d['geom'])\n"
        "lib.emu_forward.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]\n"
        "psi, phi, x = (np.ascontiguousarray(d[k], np.float32) for k in ('psi', 'phi', 'x'))\n"
        "out = np.empty((1, K, h, h), np.float32); f = np.empty((1, 2, K), np.float32)\n"
        "rc = lib.emu_forward(N, J, L, 2, M, M, psi.ctypes.data, phi.ctypes.data, x.ctypes.data, 1, out.ctypes.data, f.ctypes.data)\n"
        "assert rc == 0 and np.isfinite(out).all() and np.isfinite(f).all()\n"
        "np.save(sys.argv[3], out)\n"
        "print('clean')\n")
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    for (N, J), lib in libs:
        M, L = sizes[(N, J)], 8
        S = Scattering2D(J=J, shape=(M, M), L=L, precision="double", cache_filters=True)
        assert S._M_padded == N
        x = (np.random.default_rng(N).integers(0, 256, (1, M, M)) / 255.0).astype(np.float32)
        K, h = 1 + L * J + L * L * J * (J - 1) // 2, N // 2 ** J - 2
        inp, outp = tmp_path / ("in_%d_%d.npz" % (N, J)), tmp_path / ("out_%d_%d.npy" % (N, J))
        np.savez(inp, psi=np.stack([q["levels"][0] for q in S.psi]).astype(np.float32), phi=S.phi["levels"][0].astype(np.float32),
                 x=x, geom=np.array([N, J, L, M, K, h]))
        r = subprocess.run([sys.executable, "-c", code, str(lib), str(inp), str(outp)], capture_output=True, text=True, env=env)
        assert r.returncode == 0 and "clean" in r.stdout, "(%d, %d): " % (N, J) + r.stdout[-2000:] + r.stderr[-4000:]
        if M <= 256:                     # (512 x 512 parity is a GPU test; here it is the bounds that are checked)
            assert floored_rel(np.load(outp), S(x)) <= 1e-4 / 4, (N, J)
