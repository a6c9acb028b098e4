"""Generate tests/golden/*.npz — run in the build container:  python tests/golden/make_golden.py

Inputs: the reference's seven synthetic patterns (visualize_features.py:48-120; when /root/reference is
present the restated generators in tests/patterns.py are checked bit-for-bit against the reference's own
functions, extracted with `ast` because the module itself imports matplotlib/seaborn, absent here), plus
seeded uint8-grid noise patches.  Outputs: the oracle's results (float64 dataflow on float32 filters, and
the float32 dataflow) through the reference's wrappers.

kymatio itself is absent (PARITY UNPINNED, see oracle/__init__.py): these vectors pin the oracle and the
CUDA path to each other and to the closed-form checks in tests/test_oracle.py, not to kymatio output.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import patterns  # noqa: E402
from oracle import Scattering2D  # noqa: E402

REF = "/root/reference/src/visualization/visualize_features.py"


def check_patterns_against_reference():
    if not os.path.exists(REF):
        print("reference not present; skipping generator cross-check")
        return
    tree = ast.parse(open(REF).read())
    ns = {"np": np}
    wanted = set(patterns.REFERENCE_NAMES.values())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            exec(compile(ast.Module([node], []), REF, "exec"), ns)
    for size in (32, 64, 128):
        for name, fn in patterns.GENERATORS.items():
            ref = ns[patterns.REFERENCE_NAMES[name]](size)
            assert np.array_equal(ref, fn(size)), (name, size)
    print("tests/patterns.py == reference generators at 32/64/128")


CONFIGS = [  # (tag, M, J, L)
    ("cfg1_32_J2", 32, 2, 8), ("cfg2_64_J3", 64, 3, 8), ("cfg3_128_J4", 128, 4, 8),
    ("repo_128_J2", 128, 2, 8), ("compare_32_J3_L6", 32, 3, 6),
]


def main():
    check_patterns_against_reference()
    for tag, M, J, L in CONFIGS:
        x = patterns.all_patterns(M)                                 # float64 [7, M, M]
        rng = np.random.default_rng(42)
        noise = rng.integers(0, 256, (3, M, M)).astype(np.float64) / 255.0
        x = np.concatenate([x, noise]).astype(np.float32)            # the CUDA path consumes float32
        S64 = Scattering2D(J=J, shape=(M, M), L=L, precision="double", cache_filters=True)
        S32 = Scattering2D(J=J, shape=(M, M), L=L, precision="single", cache_filters=True)
        c64 = S64(x)
        c32 = S32(x)
        out = {
            "x": x,
            "mean64": c64.mean(axis=(-2, -1)), "std64": c64.std(axis=(-2, -1)),
            "mean32": c32.mean(axis=(-2, -1)), "std32": c32.std(axis=(-2, -1)),
        }
        if M == 32:
            out["maps64"] = c64
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
        print(tag, c64.shape, "fp32-vs-fp64 max abs", np.abs(c64 - c32).max())


if __name__ == "__main__":
    main()
