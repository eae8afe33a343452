#!/bin/bash
# ncu --set full of the MRLA-base kernels (k_base_*) over the LAST block (t = T) of a stage, forward and backward, at the
# ResNet-50 stage-3 shape (256 x 1024 x 14 x 14 bf16, T = 6) and stage-1 shape (256 x 256 x 56 x 56, T = 3).
TAG=${1:-r02_base}
for S in "1024 14 6" "256 56 3"; do
  set -- $S
  N="${1}x${2}_T${3}"
  timeout -s KILL 300 ncu --set full --clock-control none -k regex:'k_base_(mix|apply|conv|mom_bwd|scatter|dx)' --launch-skip 0 --launch-count 400 \
    -o /tmp/${TAG}_$N -f python tools/base_once.py --C $1 --HW $2 --T $3 --iters 1 > gpurun_out/${TAG}_ncu_$N.log 2>&1
  ncu -i /tmp/${TAG}_$N.ncu-rep --page raw --csv > /tmp/${TAG}_$N.csv 2>/dev/null
  python tools/ncu_raw_pick.py /tmp/${TAG}_$N.csv > gpurun_out/${TAG}_${N}_full.md
done
