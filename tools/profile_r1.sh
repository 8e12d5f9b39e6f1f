#!/bin/bash
# Round-1 profiling recipe (run under gpurun on ONE B200): launch list + full capture of the named kernels.
# Usage: tools/profile_r1.sh <tag> <kernel-regex> [extra bench args]
set -u
TAG=${1:-r01}; KRE=${2:-head_}; shift 2 || true
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --cpu-steps 0 --no-graph $*"
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 450 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.log 2>&1
# full capture of the kernels matching the regex
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 6 -c 4 -f -o gpurun_out/prof_${TAG} $BENCH > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
