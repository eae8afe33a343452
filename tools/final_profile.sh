#!/bin/bash
# Round-end evidence on one B200: (1) ncu launch list of bench.py's timed region, (2) ncu --set full of the folded
# stage-1 tail op (DRAM bytes per kernel, stalls), (3) the bench line itself.  Only summaries come back.
set -x
timeout -s KILL 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file /tmp/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_bench.err
python tools/ncu_launch_shares.py /tmp/bench_launches.csv 50 > gpurun_out/final_bench_launches.md
timeout -s KILL 200 ncu --set full --clock-control none -k regex:'k_light|k_bn_' --launch-skip 8 --launch-count 8 \
  -o /tmp/tail1 -f python tools/tail_once.py --C 256 --HW 56 --iters 2 > gpurun_out/ncu_tail1.log 2>&1
ncu -i /tmp/tail1.ncu-rep --page raw --csv > /tmp/tail1.csv 2>/dev/null
python tools/ncu_raw_pick.py /tmp/tail1.csv > gpurun_out/final_tail_stage1_full.md
timeout -s KILL 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cat gpurun_out/bench_final.json | head -c 600
