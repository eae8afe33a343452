M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
for s in "256 56" "512 28" "1024 14" "2048 7"; do set -- $s
  timeout 150 ncu --metrics $M --clock-control none --launch-skip 9 -c 9 --csv --log-file gpurun_out/l_$1.csv python tools/tail_once.py --C $1 --HW $2 --iters 2 > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/l_$1.csv gpurun_out/r01_v5_stage_$1.md
done
