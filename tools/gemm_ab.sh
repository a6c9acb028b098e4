#!/bin/bash
# Tensor-core decision with data (BASELINE.json north_star: "Tensor cores are used only if a DFT-as-GEMM variant for small
# patches beats the FFT path under 3xTF32 accuracy"): the same shapes through the fused FFT cascade, the DFT-matrix engine on
# the fp32 pipe and on the tensor cores (3xTF32 mma.sync), then ncu captures of the two heaviest GEMM launches of a forward
# call (order-2 filter product + partial inverse DFT, modulus epilogue) for both arithmetic engines.  Run under gpurun.
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r02}
for cfg in cfg1 cfg2; do
  for eng in fft gemm gemm_tf32x3; do
    B=$([ $eng = fft ] && echo 0 || echo 2048)
    timeout 300 python bench.py --config $cfg --engine $eng --batch $B --no-cpu --steps 3 --warmup 3 > $O/${T}_ab_${cfg}_${eng}.json 2> $O/${T}_ab_${cfg}_${eng}.err
    python - $O/${T}_ab_${cfg}_${eng}.json $cfg $eng <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], sys.argv[3], "%.0f patches/s" % d["value"], "parity", d["parity"].get("ok"), {o: r["floored"] for o, r in d["parity"].get("per_order", {}).items()}, "launches/step", d["gpu_launches"] // d["steps"])
except Exception as e:
    print(sys.argv[2], sys.argv[3], "FAILED", e)
PY
  done
done
timeout 300 python bench.py --config g100 --no-cpu --steps 3 > $O/${T}_bench_g100.json 2>/dev/null; cut -c1-100 $O/${T}_bench_g100.json
timeout 300 python bench.py --config g100 --engine gemm_tf32x3 --no-cpu --steps 3 > $O/${T}_bench_g100_tf32x3.json 2>/dev/null; cut -c1-100 $O/${T}_bench_g100_tf32x3.json
for eng in gemm gemm_tf32x3; do
  WST_ENGINE=$eng timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 10 -c 2 -f -o $O/${T}_prof_${eng}_cfg2 python tools/ncu_target.py cfg2 4 > $O/${T}_ncu_${eng}.log 2>&1
  tail -1 $O/${T}_ncu_${eng}.log
done
ls -la $O | head -30
