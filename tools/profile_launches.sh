#!/bin/bash
# ncu launch list of ONE steady-state training step (eager launches, same kernels as the CUDA-graph step).  One GPU, under gpurun.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
# bench: 3 warm-up + 1 timed + e2e 3+1 + classify 3 steps; ~150 of our launches per step + a few torch fills -> skip past warm-up
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 170 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --cpu-steps 0 --no-graph > gpurun_out/launches_${TAG}.log 2>&1
tail -2 gpurun_out/launches_${TAG}.log | cut -c1-300
