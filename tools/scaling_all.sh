#!/bin/bash
# Throughput of every named BASELINE shape at 2, 4 and 8 GPUs of one box (weak scaling, NCCL gather verified by rank 0);
# the N=1 lines come from single-GPU runs (tools/round_measure.sh).  Run under `gpurun --gpus 8`.  To keep the 8-GPU box
# time short, jobs with N < 8 run side by side on disjoint GPU sets (the ranks of different jobs share nothing but the
# host and the NVSwitch); each output line says so.
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-r02}
run() {  # run <cfg> <N> <gpu list> <port>
  CUDA_VISIBLE_DEVICES=$3 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 \
    --master-port $4 bench.py --gpus $2 --config $1 --steps 5 --warmup 3 --no-parity > $O/${T}_scale_$1_n$2.json 2> $O/${T}_scale_$1_n$2.err
  python - $O/${T}_scale_$1_n$2.json $1 $2 <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "N=%s" % sys.argv[3], "%.0f patches/s" % d["value"], "gather_verified", d.get("gather_verified"), "e2e %.0f" % d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], sys.argv[3], "FAILED", e)
PY
}
CFGS="cfg3 cfg1 cfg2 cfg4 cfg5"
for c in $CFGS; do run $c 8 0,1,2,3,4,5,6,7 29600; done
run cfg3 4 0,1,2,3 29601 & run cfg1 4 4,5,6,7 29602 & wait
run cfg2 4 0,1,2,3 29601 & run cfg4 4 4,5,6,7 29602 & wait
run cfg5 4 0,1,2,3 29601 & run cfg3 2 4,5 29602 & run cfg1 2 6,7 29603 & wait
run cfg2 2 0,1 29601 & run cfg4 2 2,3 29602 & run cfg5 2 4,5 29603 & wait
