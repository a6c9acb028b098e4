#!/bin/bash
# Tuning run (under gpurun): bench.py for the given configs with every variant library built by the WST_BUILD_*
# overrides of _build.py.   usage: tools/tune_variants.sh cfg1 cfg2 cfg5
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/tune_variants.txt
for lib in wst-feature*/libwst_b200*.so; do
  name=$(basename "$lib")
  for cfg in "$@"; do
    out=$(WST_BUILD_LIB=$name timeout 300 python bench.py --config $cfg --no-cpu --steps 3 --warmup 3 2>/dev/null | tail -1)
    echo "$name $cfg $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"])
except Exception as e: print("FAILED", e)')" | tee -a gpurun_out/tune_variants.txt
  done
done
