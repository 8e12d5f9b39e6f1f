#!/bin/bash
# ncu evidence of the bandwidth-bound passes (one GPU, under gpurun): full-set captures of one step's BatchNorm / stem / head launches.
set -u
TAG=${1:-r01b}
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --steps 1 --warmup 3 --cpu-steps 0 --no-graph --no-parity"
timeout 300 ncu --set full --clock-control none -k regex:'bn_act_kernel|bn_bwd_reduce_kernel|bn_bwd_apply_kernel' -s 63 -c 63 -f -o $O/prof_${TAG}_bn $BENCH > $O/prof_${TAG}_bn.log 2>&1
ncu -i $O/prof_${TAG}_bn.ncu-rep --page raw --csv > $O/prof_${TAG}_bn.csv 2>/dev/null; rm -f $O/prof_${TAG}_bn.ncu-rep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'head_|stem_|maxpool|nchw_to|channel_stats|adam_kernel' -s 10 -c 10 -f -o $O/prof_${TAG}_misc $BENCH > $O/prof_${TAG}_misc.log 2>&1
ncu -i $O/prof_${TAG}_misc.ncu-rep --page raw --csv > $O/prof_${TAG}_misc.csv 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_${TAG}.csv $BENCH > $O/launches_${TAG}.log 2>&1
ls -la $O | grep ${TAG}
