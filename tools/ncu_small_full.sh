#!/bin/bash
# ncu --set full of the four N-sized sweeps of the folded light tail at the small stage shapes (second iteration);
# only the raw-page CSV comes back (the .ncu-rep files exceed the gpurun_out size limit)
for s in "1024 14" "2048 7"; do set -- $s
  timeout 200 ncu --set full --clock-control none -k regex:k_light_nhwc --launch-skip 4 --launch-count 4 \
    -o /tmp/full_$1 -f python tools/tail_once.py --C $1 --HW $2 --iters 2 > gpurun_out/full_$1.log 2>&1
  ncu -i /tmp/full_$1.ncu-rep --page raw --csv > gpurun_out/full_$1.csv 2>/dev/null
done
