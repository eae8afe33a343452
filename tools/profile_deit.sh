#!/bin/bash
TAG=${1:-r02_deit}
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:'k_deit_light' --launch-skip 2 --launch-count 2 \
  -o /tmp/${TAG} -f python tools/deit_once.py > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > /tmp/${TAG}.csv 2>/dev/null
python tools/ncu_raw_pick.py /tmp/${TAG}.csv > gpurun_out/${TAG}_full.md
ncu -i /tmp/${TAG}.ncu-rep --page source --csv > /tmp/${TAG}_src.csv 2>/dev/null
python tools/ncu_src_top.py /tmp/${TAG}_src.csv 25 > gpurun_out/${TAG}_src.md 2>&1
cat gpurun_out/${TAG}_full.md; head -70 gpurun_out/${TAG}_src.md
