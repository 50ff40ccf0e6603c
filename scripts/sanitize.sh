#!/bin/bash
# compute-sanitizer passes over the small GPU parity tests: memcheck, racecheck, synccheck,
# initcheck.  Run on the GPU box:  gpurun -- bash scripts/sanitize.sh   → gpurun_out/san/*.log
# (summarised by hand into profiles/rNN_sanitizer_summary.txt)
set -u
OUT=gpurun_out/san
mkdir -p $OUT
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
SEL='(small_streams and T1) or pipeline_groups or replay_device or negative_dt'
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no"
  [ $tool = racecheck ] && extra="--racecheck-report all"
  start=$(date +%s)
  timeout ${SAN_TIMEOUT:-900} $CS --tool $tool $extra --error-exitcode 9 --log-file $OUT/$tool.log \
    python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > $OUT/${tool}_pytest.txt 2>&1
  echo "$tool exit=$? seconds=$(( $(date +%s) - start ))" | tee -a $OUT/summary.txt
  tail -3 $OUT/${tool}_pytest.txt | tee -a $OUT/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $OUT/$tool.log | sort | uniq -c | head -20 | tee -a $OUT/summary.txt
done
