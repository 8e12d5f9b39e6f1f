#!/bin/bash
# Round-2 ncu evidence (one GPU, under gpurun): launch list of one eager step + full-set captures of the tensor-core conv kernels, the
# fused head kernels, the large BatchNorm passes, the optimizer and the preprocessing kernel.  Summaries go to profiles/ via tools/summarize_ncu.py.
set -u
TAG=${1:-r02_final}
O=gpurun_out; mkdir -p $O
BENCH="python bench.py --steps 1 --warmup 3 --cpu-steps 0 --no-graph --no-parity --no-gpu-eager-baseline"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 720 --csv --log-file $O/launches_${TAG}.csv $BENCH > $O/launches_${TAG}.log 2>&1
timeout 170 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 136 -c 10 -f -o $O/prof_${TAG}_halo $BENCH > $O/prof_${TAG}_halo.log 2>&1
timeout 120 ncu --set full --clock-control none -k regex:'wgrad_tc_kernel|conv_tc_kernel' -s 150 -c 8 -f -o $O/prof_${TAG}_wgrad $BENCH > $O/prof_${TAG}_wgrad.log 2>&1
timeout 120 ncu --set full --clock-control none -k regex:'head_fwd_kernel|head_bwd_kernel|optim_kernel' -s 8 -c 3 -f -o $O/prof_${TAG}_head $BENCH > $O/prof_${TAG}_head.log 2>&1
timeout 150 ncu --set full --clock-control none -k regex:'bn_act_kernel|bn_bwd_reduce_kernel|bn_bwd_apply_kernel|pool_bn_bwd_reduce|stem_conv_tiled|stem_wgrad_pair|maxpool3' -s 69 -c 12 -f -o $O/prof_${TAG}_ew $BENCH > $O/prof_${TAG}_ew.log 2>&1
timeout 120 ncu --set full --clock-control none -k regex:preprocess_kernel -s 1 -c 2 -f -o $O/prof_${TAG}_pre python -m pytest tests/test_augment.py -m gpu -q -k "0-16" > $O/prof_${TAG}_pre.log 2>&1
for k in halo wgrad head ew pre; do ncu -i $O/prof_${TAG}_$k.ncu-rep --page raw --csv > $O/prof_${TAG}_$k.csv 2>/dev/null; done
rm -f $O/prof_${TAG}_wgrad.ncu-rep $O/prof_${TAG}_ew.ncu-rep $O/prof_${TAG}_pre.ncu-rep $O/prof_${TAG}_head.ncu-rep
ls -la $O | grep ${TAG}
