#!/bin/bash
# On the GPU box: compute-sanitizer over small invocations of every kernel family; one summary file per tool under
# gpurun_out/ (copied to profiles/ afterwards).  Usage: tools/sanitize.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-r2}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  out=gpurun_out/${tag}_sanitizer_${tool}.log
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no"
  timeout 900 compute-sanitizer --tool $tool $extra --print-limit 20 python tools/sanitize_driver.py > $out 2>&1
  echo "== $tool: exit $? =="; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_driver done|Error:|Race reported|hazard" $out | sort | uniq -c | head -12
done
# which kernels those invocations launch (the driver under the ncu launch list)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_sanitizer_driver_launches.csv python tools/sanitize_driver.py > /dev/null 2>&1
echo "== driver launch list: $(wc -l < gpurun_out/${tag}_sanitizer_driver_launches.csv) lines =="
