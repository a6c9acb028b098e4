#!/bin/bash
# The measurement pass committed under profiles/ (run under gpurun, one GPU): tests, bench lines for every named shape,
# the reference arm, the literal BASELINE configs[2] stream (1 M patches from uint8 host chunks), phase-cycle breakdowns,
# the ncu launch list and one full capture of the cascade kernel.
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r02}
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/${T}_gpu_tests.log; cat $O/${T}_gpu_tests.log; fi
python bench.py --steps 20 --warmup 3 --stream-total 1000000 > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err; cut -c1-150 $O/${T}_bench_cfg3.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_cfg3_reference.json 2>/dev/null; cut -c1-150 $O/${T}_bench_cfg3_reference.json
for c in cfg1 cfg2 repo cfg4; do python bench.py --config $c --steps 10 --warmup 3 --no-latency > $O/${T}_bench_$c.json 2>/dev/null; cut -c1-120 $O/${T}_bench_$c.json; done
python bench.py --config cfg5 --no-cpu --steps 3 --warmup 3 > $O/${T}_bench_cfg5.json 2>/dev/null; cut -c1-120 $O/${T}_bench_cfg5.json
python bench.py --config g100 --steps 5 --warmup 3 --no-latency > $O/${T}_bench_g100.json 2>/dev/null; cut -c1-120 $O/${T}_bench_g100.json
for c in cfg3 repo cfg2; do python tools/phase_profile.py $c > $O/${T}_phase_cycles_$c.txt 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${T}_ncu_launches_bench_cfg3.csv python bench.py --no-cpu --no-parity --steps 3 --warmup 3 > $O/${T}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cascade_kernel -s 3 -c 1 -f -o $O/${T}_prof_cfg3_bench python bench.py --no-cpu --no-parity --steps 1 --warmup 3 > $O/${T}_ncu_full.log 2>&1
tail -1 $O/${T}_ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:cascade_kernel -s 3 -c 1 -f -o $O/${T}_prof_cfg5_bench python bench.py --config cfg5 --no-cpu --no-parity --no-latency --steps 1 --warmup 3 > $O/${T}_ncu_full_cfg5.log 2>&1
tail -1 $O/${T}_ncu_full_cfg5.log
python bench.py --config p256j4 --no-cpu --steps 5 --warmup 3 --no-latency > $O/${T}_bench_p256j4.json 2>/dev/null; cut -c1-120 $O/${T}_bench_p256j4.json
python tools/tail_probe.py > $O/${T}_tail_probe.txt 2>&1
python tools/aux_bench.py > $O/${T}_aux_bench.json 2>/dev/null
ls -la $O | wc -l
