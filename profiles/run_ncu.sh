#!/bin/bash
# Profiling passes of the B200 recipe (/opt/skills/guides/B200_PROFILING.md), run under gpurun:
#   gpurun --timeout 1500 -- 'bash profiles/run_ncu.sh r01'
# 1. the plain command must exit 0 first; 2. launch list (gpu__time_duration per launch);
# 3. one --set full capture of the two trace kernels.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
if [ "${LAUNCH_ONLY:-0}" != "0" ]; then exit 0; fi   # refresh the launch list only (kernel captures unchanged)
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_grid -s 18 -c 2 \
    -o gpurun_out/prof_grid_${TAG} -f $CMD > gpurun_out/ncu_grid_${TAG}.log 2>&1
echo "grid capture rc=$?"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_brute -s 0 -c 2 \
    -o gpurun_out/prof_brute_${TAG} -f $CMD > gpurun_out/ncu_brute_${TAG}.log 2>&1
echo "brute capture rc=$?"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_brute -s 14 -c 2 \
    -o gpurun_out/prof_brute_deep_${TAG} -f $CMD > gpurun_out/ncu_brute_deep_${TAG}.log 2>&1
echo "deep brute capture rc=$?"
if [ "${SKIP_SHADE:-0}" = "0" ]; then   # gpurun copies back at most 64 MiB: SKIP_SHADE=1 when the reports get large
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 9 -c 1 \
    -o gpurun_out/prof_shade_${TAG} -f $CMD > gpurun_out/ncu_shade_${TAG}.log 2>&1
echo "shade capture rc=$?"
fi
du -sh gpurun_out
ls -la gpurun_out/
