#!/bin/bash
# 8-GPU lines of the named shapes on the current library (run under `gpurun --gpus 8`): tools/scale8.sh cfg1 cfg2 ...
cd "$(dirname "$0")/.."
O=gpurun_out
for c in "$@"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --config $c --steps 5 --warmup 3 --no-parity > $O/r02s8_scale_${c}_n8.json 2> $O/r02s8_scale_${c}_n8.err
  python -c "
import json; d=json.loads(open('$O/r02s8_scale_${c}_n8.json').read().strip().splitlines()[-1]); print('$c', d['n_gpus'], d['value'], d.get('gather_verified'), d['e2e']['value'])"
done
