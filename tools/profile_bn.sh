#!/bin/bash
# ncu --set full of the BatchNorm kernels at bottleneck shapes (stage 1: 64x56x56 relu, 256x56x56; stage 3: 256x14x14 relu, 1024x14x14)
TAG=${1:-r02_bn}
for S in "64 56 --relu" "256 56" "256 14 --relu" "1024 14"; do
  set -- $S
  N="${1}x${2}${3:+_relu}"
  timeout -s KILL 200 ncu --set full --clock-control none -k regex:'k_bn_' --launch-skip 10 --launch-count 5 \
    -o /tmp/${TAG}_$N -f python tools/bn_once.py --C $1 --HW $2 $3 --iters 3 > gpurun_out/${TAG}_ncu_$N.log 2>&1
  ncu -i /tmp/${TAG}_$N.ncu-rep --page raw --csv > /tmp/${TAG}_$N.csv 2>/dev/null
  python tools/ncu_raw_pick.py /tmp/${TAG}_$N.csv > gpurun_out/${TAG}_${N}_full.md
done
