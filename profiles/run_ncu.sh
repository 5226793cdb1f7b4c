#!/bin/bash
# Profiling passes of the B200 recipe (/opt/skills/guides/B200_PROFILING.md), run under gpurun:
#   gpurun --timeout 1500 -- 'bash profiles/run_ncu.sh r02'
# 1. the plain command must exit 0 first; 2. launch list (gpu__time_duration per launch) of the bench command;
# 3. instruction counters of every k_trace_grid launch of two C4 frames (the issue-roofline numerator);
# 4. --set full captures of the trace kernels and k_shade.  Outputs land in gpurun_out/; profiles/summarize.py
# turns them into the committed summaries.
set -u
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs"
FRAME="python tools/prof_grid.py"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
if [ "${LAUNCH_ONLY:-0}" != "0" ]; then exit 0; fi   # refresh the launch list only (kernel captures unchanged)
$FRAME > /dev/null 2>&1 &&
ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_active,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -k regex:k_trace_grid -c 32 --csv --log-file gpurun_out/gridcounters_${TAG}.csv $FRAME > gpurun_out/ncu_gridcounters_${TAG}.log 2>&1
echo "grid counters rc=$?"
$FRAME > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_grid -s 16 -c 4 \
    -o gpurun_out/prof_grid_${TAG} -f $FRAME > gpurun_out/ncu_grid_${TAG}.log 2>&1
echo "grid capture rc=$?"
$FRAME > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 11 -c 1 \
    -o gpurun_out/prof_shade_${TAG} -f $FRAME > gpurun_out/ncu_shade_${TAG}.log 2>&1
echo "shade capture rc=$?"
if [ "${SKIP_BRUTE:-0}" != "0" ]; then du -sh gpurun_out; exit 0; fi   # (gpurun brings back at most 64 MiB: three captures can exceed it)
BRUTE="python tools/gpu_brute.py C4"
$BRUTE > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_brute -s 0 -c 2 \
    -o gpurun_out/prof_brute_${TAG} -f $BRUTE > gpurun_out/ncu_brute_${TAG}.log 2>&1
echo "brute capture rc=$?"
du -sh gpurun_out
ls -la gpurun_out/ | grep ${TAG}
