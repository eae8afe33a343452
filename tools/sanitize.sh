#!/bin/bash
# compute-sanitizer (memcheck / racecheck / initcheck / synccheck) over a reduced -m gpu subset (SURVEY.md §5 "race detection").
# Usage (on the GPU box):  bash tools/sanitize.sh [tag]      -> gpurun_out/sanitize_<tag>_<tool>.log + summary
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
SUBSET="tests/test_light_gpu.py::test_light_tail_golden tests/test_light_gpu.py::test_light_tail_with_folded_add_relu tests/test_bn3_fold_gpu.py tests/test_deit_gpu.py tests/test_base_gpu.py::test_base_stage_golden tests/test_v7_gpu.py::test_virtual_x_matches_materialised_x tests/test_v7_gpu.py::test_plain_tail_on_v7_sweeps_vs_oracle"
KEXPR=${SANITIZE_K:-"not dtype2"}
for TOOL in ${SANITIZE_TOOLS:-memcheck racecheck synccheck}; do
  LOG=$OUT/sanitize_${TAG}_${TOOL}.log
  timeout ${SANITIZE_TIMEOUT:-600} compute-sanitizer --tool $TOOL --print-limit 20 --error-exitcode 0 \
    python -m pytest $SUBSET -x -q -m gpu -k "$KEXPR" -p no:cacheprovider > $LOG 2>&1
  echo "== $TOOL rc=$? ==" >> $OUT/sanitize_${TAG}_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $LOG | tail -5 >> $OUT/sanitize_${TAG}_summary.txt
done
cat $OUT/sanitize_${TAG}_summary.txt
