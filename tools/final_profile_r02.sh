#!/bin/bash
# Round-2 evidence on one B200 (summaries only come back under gpurun_out/; copy them to profiles/):
#  (1) ncu launch list of bench.py's timed region (1 eager step), (2) ncu --set full of the tail op at the four stage shapes
#  + traffic.json, (3) bench lines: default, --config mrlab / deit_tail / effnet_tail.
COMMIT=${1:-unknown}
set -x
timeout -s KILL 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file /tmp/bench_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline > /dev/null 2> gpurun_out/r02_ncu_bench.err
python tools/ncu_launch_shares.py /tmp/bench_launches.csv 60 > gpurun_out/r02_final_bench_launches.md
for S in "256 56 1" "512 28 2" "1024 14 3" "2048 7 4"; do
  set -- $S
  timeout -s KILL 300 ncu --set full --clock-control none -k regex:'k_v7|k_light|k_bn_' --launch-skip 9 --launch-count 9 \
    -o /tmp/fin_$1 -f python tools/tail_once.py --C $1 --HW $2 --iters 2 > gpurun_out/r02_final_ncu_$1.log 2>&1
  ncu -i /tmp/fin_$1.ncu-rep --page raw --csv > /tmp/fin_$1.csv 2>/dev/null
  python tools/ncu_raw_pick.py /tmp/fin_$1.csv > gpurun_out/r02_final_tail_stage$3_full.md
done
python tools/make_traffic.py /tmp/fin_256.csv gpurun_out/traffic.json $COMMIT
python tools/tail_group.py > gpurun_out/r02_final_tail_group.json 2>/dev/null
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
timeout 600 python bench.py --config mrlab --no-cpu-baseline > gpurun_out/r02_bench_mrlab_1gpu.json 2> gpurun_out/r02_bench_mrlab.err
timeout 300 python bench.py --config deit_tail > gpurun_out/r02_bench_deit_tail.json 2> gpurun_out/r02_bench_deit.err
timeout 300 python bench.py --config effnet_tail > gpurun_out/r02_bench_effnet_tail.json 2> gpurun_out/r02_bench_effnet.err
head -c 400 gpurun_out/r02_bench_1gpu.json; echo; head -c 300 gpurun_out/r02_bench_mrlab_1gpu.json; echo; cat gpurun_out/r02_bench_deit_tail.json | head -c 700; echo; head -c 400 gpurun_out/r02_bench_effnet_tail.json
