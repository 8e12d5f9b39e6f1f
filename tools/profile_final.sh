#!/bin/bash
# Round-1 final ncu evidence (one GPU, under gpurun): full-set captures of the tensor-core conv kernels and the head kernels.
set -u
TAG=${1:-r01_final}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --cpu-steps 0 --no-graph"
ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 136 -c 13 -f -o gpurun_out/prof_${TAG}_halo $BENCH > gpurun_out/prof_${TAG}_halo.log 2>&1
ncu --set full --clock-control none -k regex:wgrad_tc_kernel -s 116 -c 4 -f -o gpurun_out/prof_${TAG}_wgrad $BENCH > gpurun_out/prof_${TAG}_wgrad.log 2>&1
ncu --set full --clock-control none -k regex:head_ -s 8 -c 2 -f -o gpurun_out/prof_${TAG}_head $BENCH > gpurun_out/prof_${TAG}_head.log 2>&1
ls -la gpurun_out | grep ${TAG}
