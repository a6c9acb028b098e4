#!/bin/bash
# A/B of the WST_OPT_L2HINT variants (under gpurun, one GPU): throughput + maps of a seeded input per variant, and the
# DRAM traffic / L2 hit rate of one cascade launch per variant (ncu, four metrics).
cd "$(dirname "$0")/.."
O=gpurun_out; T=${T:-r02l2}
for v in "$@"; do
  for c in cfg5 p256j4; do
    WST_BUILD_LIB=libwst_b200_$v.so timeout 120 python tools/variant_out.py $c $O/${T}_$v 2>&1 | tail -1
  done
done | tee $O/${T}_throughput.txt
for v in ${NCU_V:-$@}; do
  WST_NO_SAVE=1 WST_BUILD_LIB=libwst_b200_$v.so timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
    --clock-control none -k regex:cascade_kernel -s 3 -c 1 --csv --log-file $O/${T}_ncu_$v.csv python tools/variant_out.py cfg5 $O/x > /dev/null 2>&1
  echo $v $(grep -E "dram__bytes|hit_rate|duration" $O/${T}_ncu_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tr '\n' ' ')
done | tee $O/${T}_ncu.txt
python - <<'P'
import numpy as np, glob, os
for c in ("cfg5", "p256j4"):
    fs = sorted(glob.glob("gpurun_out/%s_*_%s.npy" % (os.environ.get("T", "r02l2"), c)))
    if not fs: continue
    a = np.load(fs[0])
    for f in fs[1:]:
        print(c, os.path.basename(f), "identical to", os.path.basename(fs[0]), np.array_equal(a, np.load(f)))
    for f in fs: os.remove(f)
P
