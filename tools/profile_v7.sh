#!/bin/bash
# ncu --set full of the folded tail op at the stage shapes (v7 kernels) + CUDA-event timing of the op. Summaries -> gpurun_out/
TAG=${1:-r02}
python tools/tail_group.py > gpurun_out/${TAG}_tail_group.json 2> gpurun_out/${TAG}_tail_group.err
cat gpurun_out/${TAG}_tail_group.json
for S in "256 56" "512 28" "1024 14" "2048 7"; do
  set -- $S
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:'k_v7|k_light|k_bn_' --launch-skip 9 --launch-count 9 \
    -o /tmp/${TAG}_tail_$1 -f python tools/tail_once.py --C $1 --HW $2 --iters 2 > gpurun_out/${TAG}_ncu_$1.log 2>&1
  ncu -i /tmp/${TAG}_tail_$1.ncu-rep --page raw --csv > /tmp/tail_$1.csv 2>/dev/null
  python tools/ncu_raw_pick.py /tmp/tail_$1.csv > gpurun_out/${TAG}_tail_$1_full.md
done
# per-instruction hot spots of the slowest kernels (source page), stage 1
ncu -i /tmp/${TAG}_tail_256.ncu-rep --page source --csv > /tmp/src_256.csv 2>/dev/null
python tools/ncu_src_top.py /tmp/src_256.csv > gpurun_out/${TAG}_tail_256_src.md 2>&1 || true
