p=29700
run() { p=$((p+1)); env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 2 --steps 200 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('$*', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"; }
run AWR_B200_OPT_OVERLAP=1
run AWR_B200_OPT_OVERLAP=0
run AWR_B200_OPT_OVERLAP=1 AWR_B200_NCCL_SMS=8
run AWR_B200_OPT_OVERLAP=0 AWR_B200_NCCL_SMS=8
run AWR_B200_OPT_OVERLAP=1 AWR_B200_SM_RESERVE=8 AWR_B200_NCCL_SMS=16
run AWR_B200_OPT_OVERLAP=1
run AWR_B200_OPT_OVERLAP=0
