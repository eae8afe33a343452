#!/bin/bash
# ncu --set full + source page of ONE kernel of the folded tail op.  usage: profile_one.sh <tag> <kernel regex> <C> <HW> [skip]
TAG=$1; KREG=$2; C=$3; HW=$4; SKIP=${5:-1}
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"$KREG" --launch-skip $SKIP --launch-count 1 \
  -o /tmp/${TAG} -f python tools/tail_once.py --C $C --HW $HW --iters 2 > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > /tmp/${TAG}.csv 2>/dev/null
python tools/ncu_raw_pick.py /tmp/${TAG}.csv > gpurun_out/${TAG}_full.md
ncu -i /tmp/${TAG}.ncu-rep --page source --csv > /tmp/${TAG}_src.csv 2>/dev/null
head -c 3000 /tmp/${TAG}_src.csv > gpurun_out/${TAG}_src_head.txt
python tools/ncu_src_top.py /tmp/${TAG}_src.csv 60 > gpurun_out/${TAG}_src.md 2>&1
