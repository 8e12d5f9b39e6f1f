#!/bin/bash
# ncu --set full capture of a few tcgen05 conv launches. One GPU, under gpurun.  Keeps each report small (<64 MiB total).
set -u
TAG=${1:-r01_tc}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --cpu-steps 0 --no-graph"
K='conv_tc_kernel|wgrad_tc_kernel'
# second warm-up step: filtered launches 69..137 = 23 fprop (layer order), then head/deconv/... backward
ncu --set full --clock-control none -k regex:"$K" -s 69 -c 4 -f -o gpurun_out/prof_${TAG}_fprop_layer1 $BENCH > gpurun_out/prof_${TAG}_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 88 -c 4 -f -o gpurun_out/prof_${TAG}_fprop_deconv $BENCH > gpurun_out/prof_${TAG}_b.log 2>&1
ncu --set full --clock-control none -k regex:"$K" -s 92 -c 8 -f -o gpurun_out/prof_${TAG}_bwd_top $BENCH > gpurun_out/prof_${TAG}_c.log 2>&1
ls -la gpurun_out | tail -8
