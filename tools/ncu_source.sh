#!/bin/bash
# per-instruction stall attribution (ncu source page) of sweep 1 (MODE 5) and sweep B (ring) at one stage shape
C=${1:-1024}; HW=${2:-14}
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_light_nhwc_tma --launch-skip 3 --launch-count 1 \
  -o /tmp/src_s1 -f python tools/tail_once.py --C $C --HW $HW --iters 2 > gpurun_out/src_s1.log 2>&1
ncu -i /tmp/src_s1.ncu-rep --page source --csv > gpurun_out/src_s1_$C.csv 2>/dev/null
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_light_nhwc_ring --launch-skip 1 --launch-count 1 \
  -o /tmp/src_sb -f python tools/tail_once.py --C $C --HW $HW --iters 2 > gpurun_out/src_sb.log 2>&1
ncu -i /tmp/src_sb.ncu-rep --page source --csv > gpurun_out/src_sb_$C.csv 2>/dev/null
ls -la gpurun_out | head -30
